"""Every kernel family for every supported segment count (2..10): all policy-storage modes, moments in
registers and shared memory, both dynamics variants, screening, trajectories, a ragged last block and
the ARS epilogue must launch and give finite results (tools/sanitize_cases.py)."""
import os
import sys

import pytest

pytestmark = pytest.mark.gpu


def test_every_kernel_family_launches_for_every_n(S):
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import sanitize_cases
    sanitize_cases.main()


@pytest.mark.parametrize("n", [2, 3, 5, 7, 8, 10])
def test_no_out_of_bounds_writes_canaries(S, n):
    """compute-sanitizer is not available on the GPU pool, so every output buffer of the rollout / step
    kernels is embedded in a larger allocation filled with a sentinel: after the launches (ragged batch,
    every output enabled) the guard zones before and after each buffer must be untouched."""
    import numpy as np
    import torch
    rng = np.random.default_rng(n)
    B, H = 70, 70
    no, na, ws = 2 * n + 2, n - 1, (n - 1) * (2 * n + 2)
    p = S.make_params(n=n)
    SENT, PAD = -777.25, 4096

    def guarded(shape, dtype=torch.float64):
        numel = int(np.prod(shape))
        big = torch.full((numel + 2 * PAD,), SENT, dtype=dtype, device="cuda")
        return big, big[PAD:PAD + numel].view(*shape)

    bufs = {}
    for name, shape in (("returns", (B,)), ("final_state", (B, no)), ("trajectory", (H, B, no)),
                        ("stats_partial", ((B + 63) // 64, 2, no))):
        bufs[name] = guarded(shape)
    out = {k: v[1] for k, v in bufs.items()}
    W = torch.as_tensor(rng.uniform(-1, 1, ws) * 0.1).cuda()
    mean = torch.zeros(no, dtype=torch.float64, device="cuda")
    S.ops.rollout(p, H, B=B, base_policy=W, nu=0.05, seed=1, mean=mean, inv_sigma=torch.ones_like(mean),
                  stats_pivot=S.ops.reset_state(n), want_final=True, want_trajectory=True, out=out)
    sbig, sview = guarded((B, no))
    rbig, rview = guarded((B,))
    st = torch.as_tensor(rng.normal(size=(B, no))).cuda()
    ac = torch.as_tensor(rng.uniform(-5, 5, (B, na))).cuda()
    with torch.cuda.device(st.device):
        S._lib.check(S._lib.lib().swm_step_batched(__import__("ctypes").byref(p), 0, S._lib.ptr(st), S._lib.ptr(ac),
                                                   S._lib.ptr(sview), S._lib.ptr(rview), B, S._lib.stream_ptr()))
    torch.cuda.synchronize()
    for name, (big, view) in list(bufs.items()) + [("step_state", (sbig, sview)), ("step_reward", (rbig, rview))]:
        assert bool((big[:PAD] == SENT).all()) and bool((big[-PAD:] == SENT).all()), (n, name)
        assert not bool((view == SENT).any()), (n, name)  # and the buffer itself was fully written

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
    W = torch.as_tensor(rng.uniform(-1, 1, ws) * 0.1).cuda()
    mean = torch.zeros(no, dtype=torch.float64, device="cuda")
    # both rollout kernels: one thread per environment (64 per block) and lane-split (32 / L per warp)
    for tag, kern, per_block in (("thread", S.KERNEL_THREAD, 64), ("lanes", S.KERNEL_LANES, S.ops.lane_split_envs_per_warp(n))):
        kb = {}
        for name, shape in (("returns", (B,)), ("final_state", (B, no)), ("trajectory", (H, B, no)),
                            ("stats_partial", ((B + per_block - 1) // per_block, 2, no))):
            kb[name] = guarded(shape)
        S.ops.rollout(p, H, B=B, base_policy=W, nu=0.05, seed=1, mean=mean, inv_sigma=torch.ones_like(mean),
                      stats_pivot=S.ops.reset_state(n), want_final=True, want_trajectory=True, kernel=kern,
                      out={k: v[1] for k, v in kb.items()})
        bufs.update({tag + "_" + k: v for k, v in kb.items()})
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


def test_no_out_of_bounds_writes_ars_kernels(S):
    """Same guard-zone check for the ARS bookkeeping kernels (ranking, update, statistics, reductions, Philox
    table, batched select_action, learning-curve record) at awkward sizes."""
    import numpy as np
    import torch
    rng = np.random.default_rng(3)
    SENT, PAD = -777.25, 2048

    def guarded(shape, dtype=torch.float64, sent=SENT):
        numel = int(np.prod(shape))
        big = torch.full((numel + 2 * PAD,), sent, dtype=dtype, device="cuda")
        return big, big[PAD:PAD + numel].view(*shape)

    n, N, R = 7, 37, 3
    no, ws = 2 * n + 2, (n - 1) * (2 * n + 2)
    checks = []
    rets = torch.as_tensor(rng.normal(size=2 * N)).cuda()
    ob, order = guarded((N,), torch.int32, -7)
    S.ops.ars_topb(rets, out=order)
    checks.append(("order", ob, order, -7))
    Wb, W = guarded((ws,))
    W.zero_()
    sb, sig = guarded((1,))
    S.ops.ars_update(W, rets, N, order=order, n_order=11, alpha=0.1, seed=5, sigma_out=sig)
    checks += [("W", Wb, W, SENT), ("sigma", sb, sig, SENT)]
    part = torch.as_tensor(rng.normal(size=(5, 2, no))).cuda()
    rb, rec = guarded((1 + 2 * no,))
    S.ops.stats_finalize(part, 1234.0, S.ops.reset_state(n), out=rec)
    checks.append(("record", rb, rec, SENT))
    mb, mean = guarded((no,))
    ib, inv = guarded((no,))
    running = torch.zeros(1 + 2 * no, dtype=torch.float64, device="cuda")
    S.ops.stats_merge(running, rec.reshape(1, -1).clone(), mean, inv)
    checks += [("mean", mb, mean, SENT), ("inv_sigma", ib, inv, SENT)]
    gb, grp = guarded((2 * N,))
    S.ops.reduce_returns(torch.as_tensor(rng.normal(size=2 * N * R)).cuda(), R, out=grp)
    checks.append(("reduced", gb, grp, SENT))
    kb, mask = guarded((N,), torch.int32, -7)
    nb, npass = guarded((1,), torch.int32, -7)
    S.ops.screen_mask(rets, 0.0, mask, npass)
    checks += [("mask", kb, mask, -7), ("n_pass", nb, npass, -7)]
    cb, curve = guarded((9,))
    curve.fill_(0.0)
    S.ops.record_nanmean(rets, curve, torch.full((1,), 100, dtype=torch.int32, device="cuda"))  # clamps to the last slot
    checks.append(("curve", cb, curve, SENT))
    torch.cuda.synchronize()
    for name, big, view, sent in checks:
        assert bool((big[:PAD] == sent).all()) and bool((big[-PAD:] == sent).all()), name
        assert not bool((view == sent).any()), name
    assert float(curve[8]) == float(torch.nanmean(rets)) or abs(float(curve[8]) - float(rets.mean())) < 1e-12

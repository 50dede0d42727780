"""Throughput / latency of the rollout kernel for several workloads (used to pick register caps)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import swimmer_ars_b200 as S

def timeit(fn, reps=3):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

rng = np.random.default_rng(0)
H = 200
rows = []
for n in (3, 5, 10):
    p = S.make_params(n=n)
    ws = (n - 1) * (2 * n + 2)
    W = torch.as_tensor(rng.uniform(-1, 1, ws) * 0.1).cuda()
    mean = torch.zeros(2 * n + 2, dtype=torch.float64, device="cuda"); inv = torch.ones_like(mean)
    piv = S.ops.reset_state(n)
    for B in (2048, 65536, 262144):
        ac = torch.as_tensor(rng.uniform(-5, 5, (B, n - 1))).cuda()
        t_fixed = timeit(lambda: S.ops.rollout(p, H, actions=ac))
        t_v1 = timeit(lambda: S.ops.rollout(p, H, B=B, base_policy=W, nu=0.02, seed=1))
        t_v2 = timeit(lambda: S.ops.rollout(p, H, B=B, base_policy=W, nu=0.02, seed=1, mean=mean, inv_sigma=inv, stats_pivot=piv))
        t_grp = timeit(lambda: S.ops.rollout(p, H, B=B, base_policy=W, nu=0.02, seed=1, rollouts_per_policy=128, init_perturb=0.01)) if B >= 256 else float("nan")
        f = lambda t: B * H / t * 1e3
        print("n=%2d B=%6d  env-steps/s: fixed %.3e  philox-V1 %.3e  V2+stats %.3e  R=128 %.3e" % (n, B, f(t_fixed), f(t_v1), f(t_v2), f(t_grp)), flush=True)

"""A/B of library builds on the per-thread rollout kernel (run once per SWM_LIB_PATH on the same box):
config[1] under graph replay with the default and the plain schedule, and a 5- / 10-segment V2 batch."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import swimmer_ars_b200 as S

flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")


def timed(fn, reps=14, graph=True):
    fn()
    torch.cuda.synchronize()
    if graph:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        run = g.replay
    else:
        run = fn
    ts = []
    for _ in range(reps):
        flush.fill_(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts[2:]))


tag = os.path.basename(os.environ.get("SWM_LIB_PATH", "default"))
p = S.make_params(n=3)
B, H = 65536, 1000
ac = torch.as_tensor(np.random.default_rng(0).uniform(-5, 5, (B, 2))).cuda()
out = {"returns": torch.empty(B, dtype=torch.float64, device="cuda"),
       "final_state": torch.empty(B, 8, dtype=torch.float64, device="cuda")}
for sched in (None, "plain"):
    ms = timed(lambda: S.ops.rollout(p, H, actions=ac, want_final=True, out=out, schedule=sched))
    print("%-14s n3 fixed %-6s %.4f ms  %.4e env-steps/s" % (tag, sched, ms, B * H / ms * 1e3), flush=True)
for n, B in ((5, 131072), (10, 131072)):
    p = S.make_params(n=n)
    rng = np.random.default_rng(1)
    ws = (n - 1) * (2 * n + 2)
    W = torch.as_tensor(rng.normal(0, 0.05, (1, ws))).cuda()
    mean = torch.zeros(2 * n + 2, dtype=torch.float64, device="cuda")
    inv = torch.ones(2 * n + 2, dtype=torch.float64, device="cuda")
    fn = lambda: S.ops.rollout(p, H, base_policy=W.reshape(n - 1, 2 * n + 2), B=B, nu=0.02, seed=3, mean=mean,
                               inv_sigma=inv, kernel=S._lib.KERNEL_THREAD, schedule="plain")
    try:
        ms = timed(fn, reps=6)
        print("%-14s n%d V2 thread  %.4f ms  %.4e env-steps/s" % (tag, n, ms, B * H / ms * 1e3), flush=True)
    except Exception as e:  # signature drift between builds is not what this tool measures
        print(tag, n, "skipped:", repr(e)[:200])

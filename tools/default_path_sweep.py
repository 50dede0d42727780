"""Acceptance table for "the default path is the fast path": SwimmerEnv.rollout_batched / ops.rollout with no
schedule argument (the library decides) against one plain launch, 8 k .. 262 k envs, fixed actions and V2
policies, eager calls and CUDA-graph replay.

  python tools/default_path_sweep.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import swimmer_ars_b200 as S


def timed(fn, reps=6, graph=False):
    fn(); fn()
    torch.cuda.synchronize()
    if graph:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        run = g.replay
    else:
        run = fn
    run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        run()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    rng = np.random.default_rng(0)
    H = 1000
    for n, mode in ((3, "fixed"), (3, "v2"), (5, "v2"), (10, "v2")):
        p = S.make_params(n=n)
        no, ws = 2 * n + 2, (n - 1) * (2 * n + 2)
        for B in (8192, 16384, 32768, 57344, 65536, 73728, 98304, 131072, 262144):
            out = {"returns": torch.empty(B, dtype=torch.float64, device="cuda"),
                   "final_state": torch.empty(B, no, dtype=torch.float64, device="cuda")}
            if mode == "fixed":
                ac = torch.as_tensor(rng.uniform(-5, 5, (B, n - 1))).cuda()
                call = lambda **k: S.ops.rollout(p, H, actions=ac, want_final=True, out=out, **k)
            else:
                W = torch.as_tensor(rng.uniform(-1, 1, ws) * 0.05).cuda()
                mean = torch.zeros(no, dtype=torch.float64, device="cuda")
                inv = torch.ones_like(mean)
                call = lambda **k: S.ops.rollout(p, H, B=B, base_policy=W, nu=0.01, seed=1, mean=mean, inv_sigma=inv,
                                                 want_final=True, out=out, **k)
            t_plain = timed(lambda: call(schedule="plain"))
            t_def = timed(lambda: call())
            t_def_g = timed(lambda: call(), graph=True)
            print("n=%2d %-5s B=%6d (%.2f warps/SMSP): plain %.3f ms | default eager %.3f ms (%+.1f%%) | default, graph replay %.3f ms (%+.1f%%)"
                  % (n, mode, B, B / 32 / 592, t_plain, t_def, 100 * (t_plain / t_def - 1), t_def_g, 100 * (t_plain / t_def_g - 1)),
                  flush=True)


if __name__ == "__main__":
    main()

"""Where does chunked scheduling (ops.ChunkedRollout) pay off?  Plain launch vs (16 x 64) and (8 x 128) for
several segment counts and batch sizes, V2 policy rollouts with moments, replayed from CUDA graphs."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import swimmer_ars_b200 as S


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    rng = np.random.default_rng(0)
    H = 1000
    for n in (3, 5, 10):
        p = S.make_params(n=n)
        no, ws = 2 * n + 2, (n - 1) * (2 * n + 2)
        W = torch.as_tensor(rng.uniform(-1, 1, ws) * 0.05).cuda()
        mean = torch.zeros(no, dtype=torch.float64, device="cuda")
        inv = torch.ones_like(mean)
        piv = S.ops.reset_state(n)
        for B in (8192, 32768, 65536, 131072, 262144):
            kw = dict(nu=0.01, seed=1, mean=mean, inv_sigma=inv)
            out = {"returns": torch.empty(B, dtype=torch.float64, device="cuda")}
            t0 = timed(lambda: S.ops.rollout(p, H, B=B, base_policy=W, stats_pivot=piv, out=dict(out), **kw))
            row = "n=%2d B=%6d (%.2f warps/SMSP): plain %.3f ms" % (n, B, B / 32 / 592, t0)
            for n_sub, chunk in ((16, 64), (8, 128), (8, 256)):
                plan = S.ops.ChunkedRollout(p, H, B=B, n_sub=n_sub, chunk=chunk, base_policy=W, stats_pivot=piv, **kw)
                t1 = timed(plan.run)
                row += " | %dx%d %.3f ms (%+.0f%%)" % (n_sub, chunk, t1, 100 * (t0 / t1 - 1))
            print(row, flush=True)


if __name__ == "__main__":
    main()

"""Lane-split vs one-thread-per-environment rollout kernel: ms per 1,000-step rollout versus batch size
(the measured crossover behind capi.cu's kLaneSplitMaxWarpsPerSmsp).

  python tools/lane_split_sweep.py [n ...]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import swimmer_ars_b200 as S


def time_ms(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ns = [int(x) for x in sys.argv[1:]] or [3, 5, 7, 10]
    H = 1000
    rng = np.random.default_rng(0)
    for n in ns:
        p = S.make_params(n=n)
        no, ws = 2 * n + 2, (n - 1) * (2 * n + 2)
        W = torch.as_tensor(rng.uniform(-1, 1, ws) * 0.05).cuda()
        mean = torch.zeros(no, dtype=torch.float64, device="cuda")
        inv = torch.ones_like(mean)
        piv = S.ops.reset_state(n)
        print("n=%d  V2 + moments, Philox policies, H=%d" % (n, H))
        for B in [int(x) for x in os.environ.get("SWEEP_B", "64,256,512,1024,2048,4096,8192,16384,32768").split(",")]:
            row = []
            for kern in (S.KERNEL_THREAD, S.KERNEL_LANES, S.KERNEL_LANES2, S.KERNEL_LANES3):
                out = {}
                fn = lambda: S.ops.rollout(p, H, B=B, base_policy=W, nu=0.01, seed=1, mean=mean, inv_sigma=inv,
                                           stats_pivot=piv, kernel=kern, out=out)
                r = fn()
                out.update(returns=r.returns, stats_partial=r.stats_partial)
                row.append(time_ms(fn))
            print("  B=%6d  thread %.4f ms   lanes %.4f ms   lanes2 %.4f ms   lanes3 %.4f ms   (%.0f / %.0f / %.0f / %.0f cycles per "
                  "step @1.965 GHz)" % (B, row[0], row[1], row[2], row[3], row[0] * 1965, row[1] * 1965, row[2] * 1965, row[3] * 1965))


if __name__ == "__main__":
    main()

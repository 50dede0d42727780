"""Throughput of the batched single-step kernel (swm_step_batched) against its HBM roofline:
bytes per env = (2n+2 state in + n-1 action in + 2n+2 state out + 1 reward out) * 8."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import swimmer_ars_b200 as S

HBM_GBS = 6551.0  # MEASURED_PEAKS.json


def main():
    rng = np.random.default_rng(0)
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")
    for variant, vname in ((S.GYM, "gym"), (S.RLGLUE, "rlglue")):
        for n in (3, 5, 10):
            p = S.make_params(n=n, h=0.001 if variant == S.GYM else 0.01)
            for B in (1 << 16, 1 << 20, 1 << 22):
                st = torch.as_tensor(rng.normal(size=(B, 2 * n + 2))).cuda()
                ac = torch.as_tensor(rng.uniform(-5, 5, (B, n - 1))).cuda()
                out = torch.empty_like(st)
                for _ in range(3):
                    S.ops.step_batched(p, st, ac, variant, out=out)
                ts = []
                for _ in range(10):
                    flush.fill_(0.0)
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    S.ops.step_batched(p, st, ac, variant, out=out)
                    e1.record()
                    torch.cuda.synchronize()
                    ts.append(e0.elapsed_time(e1))
                ms = float(np.median(ts))
                byts = B * (2 * (2 * n + 2) + (n - 1) + 1) * 8
                print("%-6s n=%2d B=%8d  %8.4f ms  %.3e env-steps/s  %7.1f GB/s = %.2f of measured HBM copy"
                      % (vname, n, B, ms, B / ms * 1e3, byts / ms / 1e6, byts / ms / 1e6 / HBM_GBS), flush=True)


if __name__ == "__main__":
    main()

"""Cost of the optional trajectory output (saved_states of ars/environment.py:53; Database format of
ars/database.py): [H, B, 2n+2] time-major doubles written from inside the fused rollout."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import swimmer_ars_b200 as S


def t(fn, reps=5):
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
    rng = np.random.default_rng(0)
    for n, B, H in ((3, 65536, 1000), (5, 65536, 500), (10, 32768, 500)):
        p = S.make_params(n=n)
        no = 2 * n + 2
        ac = torch.as_tensor(rng.uniform(-5, 5, (B, n - 1))).cuda()
        out = {"returns": torch.empty(B, dtype=torch.float64, device="cuda"),
               "trajectory": torch.empty(H, B, no, dtype=torch.float64, device="cuda")}
        t0 = t(lambda: S.ops.rollout(p, H, actions=ac, out={"returns": out["returns"]}))
        t1 = t(lambda: S.ops.rollout(p, H, actions=ac, want_trajectory=True, out=out))
        gb = H * B * no * 8 / 1e9
        print("n=%2d B=%6d H=%4d: %.3f ms without, %.3f ms with the %.2f GB trajectory (%.0f GB/s written; "
              "HBM-only time at 6551 GB/s would be %.3f ms)" % (n, B, H, t0, t1, gb, gb / t1 * 1e3, gb / 6551 * 1e3))


if __name__ == "__main__":
    main()

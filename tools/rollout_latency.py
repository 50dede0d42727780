"""Per-step latency of the fused rollout kernel versus batch size (1 warp ... full chip)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import swimmer_ars_b200 as S
p = S.make_params(n=3)
rng = np.random.default_rng(0)
for B in (32, 64, 128, 148 * 32, 148 * 64, 148 * 128, 65536, 148 * 4 * 4 * 32, 148 * 4 * 5 * 32, 148*4*6*32, 148*4*8*32, 262144):
    ac = torch.as_tensor(rng.uniform(-5, 5, (B, 2))).cuda()
    for _ in range(2):
        S.ops.rollout(p, 1000, actions=ac)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        S.ops.rollout(p, 1000, actions=ac)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print("B=%7d  warps/SMSP=%5.2f  %.3f ms  %.0f cycles/step @1.92GHz  %.3e env-steps/s" % (B, B / 32 / 592, ms, ms * 1.92e6 / 1000, B * 1000 / ms * 1e3))

"""Per-rank share of BASELINE config[4] at 8 GPUs (n=10, 512 directions x 2 x 128 rollouts = 131,072 envs):
ARS iteration time with and without chunked rollout scheduling."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import swimmer_ars_b200 as S

p = S.make_params(n=10)
for chunks in (None, (4, 256), (8, 256), (8, 128), (16, 128), (16, 64)):
    eng = S.ArsEngine(p, N=512, b=512, alpha=0.0075, nu=0.01, H=1000, v2=True, semantics=S.ARS_AGENT, seed=0,
                      rollouts_per_direction=128, init_perturb=1e-2, distributed=False, use_graph=True,
                      rollout_chunks=chunks)
    for _ in range(3):
        eng.run_iteration()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        eng.run_iteration()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print("rollout_chunks=%-10s %.3f ms / iteration  (%.3e env-steps/s)  mean return %.6f"
          % (chunks, ms, 131072e3 / ms * 1e3, float(eng.returns.mean())), flush=True)
    del eng

"""The timed step of bench.py (one default rollout_batched call captured in a CUDA graph: 256 chunk kernels on 16
streams) replayed inside an NVTX range, so that `ncu --graph-profiling graph --nvtx --nvtx-include "bench_step/"`
measures the concurrent schedule as ONE workload (per-kernel ncu passes serialise the 256 launches)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import swimmer_ars_b200 as S

env = S.SwimmerEnv(n=3)
ac = torch.as_tensor(np.random.default_rng(0).uniform(-5, 5, (65536, 2))).cuda()
out = {"returns": torch.empty(65536, dtype=torch.float64, device="cuda"),
       "final_state": torch.empty(65536, 8, dtype=torch.float64, device="cuda")}
fn = lambda: env.rollout_batched(1000, actions=ac, want_final=True, out=out)
fn()
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    fn()
for _ in range(3):
    g.replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(3):
    torch.cuda.nvtx.range_push("bench_step")
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_pop()
    print("graph step %d: %.4f ms" % (i, e0.elapsed_time(e1)))

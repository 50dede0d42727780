"""Forced schedules (n_sub x chunk) of BASELINE config[1] under CUDA-graph replay: which one should AUTO pick?"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import swimmer_ars_b200 as S

p = S.make_params(n=3)
B, H = 65536, 1000
ac = torch.as_tensor(np.random.default_rng(0).uniform(-5, 5, (B, 2))).cuda()
out = {"returns": torch.empty(B, dtype=torch.float64, device="cuda"),
       "final_state": torch.empty(B, 8, dtype=torch.float64, device="cuda")}
flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")
for sched in ("plain", (8, 256), (8, 128), (16, 128), (16, 64), (24, 64), (32, 64), (32, 128), (12, 64), (20, 64)):
    fn = lambda: S.ops.rollout(p, H, actions=ac, want_final=True, out=out, schedule=sched)
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    ts = []
    for _ in range(12):
        flush.fill_(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts[2:]))
    print("%-10s %.4f ms  %.3e env-steps/s" % (sched, ms, B * H / ms * 1e3), flush=True)

"""End-to-end sanity: the ARS engine actually learns to swim.  Mean exploration return per iteration
(ARSAgent.runTraining's learning curve, ars_agent.py:195-201) for BASELINE config[0] (n=3, V1, 8 directions)
over 8 seeds on one GPU, and for config[2] (n=5, V2, 1,024 directions)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import swimmer_ars_b200 as S


def main():
    t0 = time.perf_counter()
    fan = S.SeedFanout(S.make_params(n=3), range(8), N=8, b=8, alpha=0.0075, nu=0.01, H=1000, max_iterations=1024)
    curves = fan.run(1000)
    dt = time.perf_counter() - t0
    print("config[0] x 8 seeds, 1001 iterations each in %.2f s wall" % dt)
    for it in (0, 10, 50, 100, 200, 400, 700, 1000):
        print("  iteration %4d: mean return over seeds %9.4f  (min %9.4f, max %9.4f)"
              % (it, curves[:, it].mean(), curves[:, it].min(), curves[:, it].max()))
    t0 = time.perf_counter()
    eng = S.ArsEngine(S.make_params(n=5), N=1024, b=1024, alpha=0.0075, nu=0.01, H=1000, v2=True, seed=0,
                      distributed=False, use_graph=True, curve_capacity=512)
    for _ in range(301):
        eng.run_iteration()
    c = eng.curve.cpu().numpy()
    dt = time.perf_counter() - t0
    print("config[2] (n=5, V2, 1,024 directions), 301 iterations in %.2f s wall" % dt)
    for it in (0, 10, 50, 100, 200, 300):
        print("  iteration %4d: mean return %9.4f" % (it, c[it]))
    assert curves[:, -1].mean() > curves[:, 0].mean() + 1.0 and c[300] > c[0] + 1.0, "no learning?"


if __name__ == "__main__":
    main()

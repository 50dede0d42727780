"""ARS iteration latency of the engine for the BASELINE ARS configurations, and the aggregate rate of
the seed fan-out.  Reports GPU time per iteration (CUDA events over K iterations enqueued back to
back) and the host time spent enqueueing them."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import swimmer_ars_b200 as S


def bench_engine(tag, K, **kw):
    for g in (False, True):
        _bench_engine(tag + (" [graph]" if g else " [eager]"), K, use_graph=g, **kw)


def _bench_engine(tag, K, **kw):
    eng = S.ArsEngine(distributed=False, **kw)
    for _ in range(3):
        eng.run_iteration()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(K):
        eng.run_iteration()
    e1.record()
    t_host = time.perf_counter() - t0
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    print("%-56s %8.3f ms/iter on the device (%7.1f it/s), host enqueue %6.3f ms/iter" %
          (tag, ms, 1e3 / ms, 1e3 * t_host / K))


def main():
    p3, p5 = S.make_params(n=3), S.make_params(n=5)
    bench_engine("config[0] n=3 V1 N=8 H=1000", 50, params=p3, N=8, b=8, alpha=0.0075, nu=0.01, H=1000)
    bench_engine("config[2] n=5 V2 N=1024 H=1000", 20, params=p5, N=1024, b=1024, alpha=0.0075, nu=0.01, H=1000, v2=True)
    bench_engine("config[3]-like n=3 V1 N=256 H=1000 + sim screen", 20, params=S.make_params(n=3, l_i=.8, m_i=1.2, k=10.2),
                 N=256, b=256, alpha=0.0075, nu=0.01, H=1000,
                 sim_params=S.make_params(n=3, l_i=.8006, m_i=1.2006, k=10.2006), sim_threshold=-1e9)
    for seeds in (1, 8, 32):
        fan = S.SeedFanout(p3, range(seeds), N=8, b=8, alpha=0.0075, nu=0.01, H=1000)
        fan.run(2)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fan.run(20, include_initial=False)
        dt = time.perf_counter() - t0
        print("seed fan-out config[0] x %2d seeds: %8.3f ms per round of iterations, %8.1f agent-iterations/s" %
              (seeds, 1e3 * dt / 20, seeds * 20 / dt))


if __name__ == "__main__":
    main()

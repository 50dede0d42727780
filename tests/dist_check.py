"""Multi-GPU check, run under torchrun on a box with >= 2 GPUs (not collected by pytest):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
      --master-port 29531 tests/dist_check.py

Directions sharded over the ranks + one NCCL all-gather per iteration must give (a) bit-identical
policy / statistics on every rank and (b) the single-GPU result (returns bit-identical in iteration
0; later iterations within 1e-10, because the V2 moments are merged per rank instead of in one sum).
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import swimmer_ars_b200 as S  # noqa: E402
from swimmer_ars_b200 import distributed as D  # noqa: E402


def main():
    rank, world, device = D.init_from_env()
    assert world >= 2, "run under torchrun with >= 2 ranks"
    for n, v2, sem, R in ((5, True, S.ARS_AGENT, 1), (3, False, S.ARS_TOPB, 1), (10, True, S.ARS_TOPB, 32)):
        p = S.make_params(n=n)
        kw = dict(N=16, b=5, alpha=0.02, nu=0.03, H=150, v2=v2, semantics=sem, seed=42,
                  rollouts_per_direction=R, init_perturb=1e-2 if R > 1 else 0.0, device=device)
        eng = S.ArsEngine(p, **kw)                      # sharded
        ref = S.ArsEngine(p, distributed=False, **kw)   # every rank also runs the whole problem alone
        for it in range(3):
            a = eng.run_iteration().clone()
            b = ref.run_iteration().clone()
            if it == 0:
                assert torch.equal(a, b), "iteration-0 returns must be bit-identical"
            def rel(x, y):
                return ((x - y).abs().max() / y.abs().max().clamp_min(1e-300)).item()
            dr, dw = rel(a, b), rel(eng.W, ref.W)
            dm = rel(eng.mean, ref.mean) if v2 else 0.0
            ds = rel(eng.inv_sigma, ref.inv_sigma) if v2 else 0.0
            if rank == 0:
                print("  n=%d it=%d  rel diff vs single GPU: returns %.1e  W %.1e  mean %.1e  inv_sigma %.1e"
                      % (n, it, dr, dw, dm, ds), flush=True)
            # north-star tolerance for returns and weight updates: 1e-6 relative
            assert dr < 1e-6 and dw < 1e-6 and dm < 1e-6 and ds < 1e-6
            # bit-identical across ranks
            buf = [torch.empty_like(eng.W) for _ in range(world)]
            dist.all_gather(buf, eng.W)
            assert all(torch.equal(buf[0], x) for x in buf), "policy differs between ranks"
            sb = [torch.empty_like(eng.stats) for _ in range(world)]
            dist.all_gather(sb, eng.stats)
            assert all(torch.equal(sb[0], x) for x in sb), "V2 statistics differ between ranks"
        if rank == 0:
            print("dist_check ok: n=%d v2=%s semantics=%d R=%d world=%d" % (n, v2, sem, R, world), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""Multi-GPU parity check of the sharded ARS iteration.  Run by tests/test_multi_gpu.py (pytest -m gpu, when
the box has >= 2 GPUs) and by hand:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
      --master-port 29531 tests/dist_check.py

Directions sharded over the ranks + one record exchange per iteration -- over NVLink peer memory inside a
captured CUDA graph ("p2p") and through the NCCL all-gather fallback ("collective") -- must give
(a) bit-identical policy / statistics on every rank and (b) the single-GPU result (returns bit-identical in
iteration 0; later iterations within 1e-10, because the V2 moments are merged per rank instead of in one sum).
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import swimmer_ars_b200 as S  # noqa: E402
from swimmer_ars_b200 import distributed as D  # noqa: E402


def rel(x, y):
    x, y = torch.nan_to_num(x), torch.nan_to_num(y)
    return ((x - y).abs().max() / y.abs().max().clamp_min(1e-300)).item()


def main():
    rank, world, device = D.init_from_env()
    assert world >= 2, "run under torchrun with >= 2 ranks"
    cases = [(5, True, S.ARS_AGENT, 1, False), (3, False, S.ARS_TOPB, 1, False), (10, True, S.ARS_TOPB, 32, False),
             (3, True, S.ARS_AGENT, 1, True)]  # last: reward-constraint screening through a simulator model
    for transport, use_graph in (("p2p", True), ("collective", False)):
        for n, v2, sem, R, safe in cases:
            p = S.make_params(n=n)
            kw = dict(N=16, b=5, alpha=0.02, nu=0.03, H=150, v2=v2, semantics=sem, seed=42,
                      rollouts_per_direction=R, init_perturb=1e-2 if R > 1 else 0.0, device=device)
            if safe:
                kw.update(sim_params=S.make_params(n=n, l_i=1.001, m_i=0.999, k=10.01), sim_threshold=0.0,
                          initial_policy=torch.linspace(-0.3, 0.3, (n - 1) * (2 * n + 2)))
            eng = S.ArsEngine(p, transport=transport, use_graph=use_graph, shard=True, **kw)   # sharded (forced: these
            # batches are small enough for shard="auto" to run them replicated, see below)
            ref = S.ArsEngine(p, distributed=False, **kw)                          # the whole problem alone
            assert eng.exchange.transport == transport and eng.world == world, (eng.exchange.transport, transport)
            for it in range(4):
                a = eng.run_iteration().clone()
                b = ref.run_iteration().clone()
                if it == 0:
                    assert torch.equal(torch.nan_to_num(a), torch.nan_to_num(b)), "iteration-0 returns must be bit-identical"
                assert torch.equal(torch.isnan(a), torch.isnan(b)), "screened-out directions differ"
                dr, dw = rel(a, b), rel(eng.W, ref.W)
                dm = rel(eng.mean, ref.mean) if v2 else 0.0
                ds = rel(eng.inv_sigma, ref.inv_sigma) if v2 else 0.0
                if rank == 0:
                    print("  %s n=%d it=%d  rel diff vs single GPU: returns %.1e  W %.1e  mean %.1e  inv_sigma %.1e%s"
                          % (transport, n, it, dr, dw, dm, ds, "  [%d screened]" % int(torch.isnan(a).sum() // 2) if safe else ""),
                          flush=True)
                # north-star tolerance for returns and weight updates: 1e-6 relative
                assert dr < 1e-6 and dw < 1e-6 and dm < 1e-6 and ds < 1e-6
                # bit-identical across ranks
                buf = [torch.empty_like(eng.W) for _ in range(world)]
                dist.all_gather(buf, eng.W)
                assert all(torch.equal(buf[0], x) for x in buf), "policy differs between ranks"
                sb = [torch.empty_like(eng.stats) for _ in range(world)]
                dist.all_gather(sb, eng.stats)
                assert all(torch.equal(sb[0], x) for x in sb), "V2 statistics differ between ranks"
            if use_graph:
                assert eng._graph is not None, "sharded iteration was not captured"
                assert eng.check_exchange() == 4
            eng.exchange.close()
            if rank == 0:
                print("dist_check ok: %s graph=%s n=%d v2=%s semantics=%d R=%d safe=%s world=%d"
                      % (transport, use_graph, n, v2, sem, R, safe, world), flush=True)
    # shard="auto": a batch that is latency-bound as a whole is not sharded; every rank runs all of it, no exchange
    p = S.make_params(n=3)
    kw = dict(N=16, b=5, alpha=0.02, nu=0.03, H=150, semantics=S.ARS_TOPB, seed=42, device=device)
    auto, ref = S.ArsEngine(p, use_graph=True, **kw), S.ArsEngine(p, distributed=False, **kw)
    assert auto.replicated and auto.world == 1 and auto.exchange.transport == "local"
    for it in range(3):
        assert torch.equal(auto.run_iteration(), ref.run_iteration()) and torch.equal(auto.W, ref.W)
    big = S.ArsEngine(S.make_params(n=5), N=1024, b=8, alpha=0.02, nu=0.03, H=10, v2=True, seed=1, device=device)
    # config[2] is sharded, its share runs on a faster kernel -- but over at most 4 ranks (one lane group per SM);
    # 8 ranks run as 2 blocks of 4, bit-identical across blocks
    assert not big.replicated and big.world == min(world, 4) and big.shard_replicas == world // big.world
    for it in range(2):
        big.run_iteration()
    buf = [torch.empty_like(big.W) for _ in range(world)]
    dist.all_gather(buf, big.W)
    assert all(torch.equal(buf[0], x) for x in buf) and bool(torch.isfinite(big.W).all())
    big.exchange.close()
    if rank == 0:
        print("dist_check ok: shard=auto replicates the latency-bound batch, shards config[2]", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""One process per GPU plumbing (torch.distributed): rendezvous from the torchrun environment,
contiguous sharding of the N directions over ranks."""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialises the default process group from RANK / WORLD_SIZE / MASTER_* (torchrun) and binds
    this process to its GPU.  Returns (rank, world, device).  No-op for a single process."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    cuda = torch.cuda.is_available()
    device = torch.device("cuda", local % max(torch.cuda.device_count(), 1)) if cuda else torch.device("cpu")
    if cuda:
        torch.cuda.set_device(device)
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        dist.init_process_group(backend or ("nccl" if cuda else "gloo"), rank=rank, world_size=world,
                                **({"device_id": device} if cuda else {}))
    return rank, world, device


def shard_directions(N, rank, world):
    """Contiguous shard [lo, hi) of N directions owned by `rank` (both signs and all repeats of a
    direction stay on one rank)."""
    if N % world != 0:
        raise ValueError("N=%d does not divide over %d ranks" % (N, world))
    per = N // world
    return rank * per, (rank + 1) * per


def pack_record(returns_local, stats_record=None):
    """[returns(2 N_local) | count, mean[F], M2[F]] -- the per-rank record that is all-gathered."""
    if stats_record is None:
        return returns_local
    return torch.cat([returns_local, stats_record])


def unpack_records(gathered, world, n_returns_local, n_features):
    """-> (returns_all[2N], records[world, 1+2F]) from the all-gathered buffer (rank order)."""
    g = gathered.view(world, -1)
    returns = g[:, :n_returns_local].reshape(-1)
    records = g[:, n_returns_local:n_returns_local + 1 + 2 * n_features]
    return returns, records

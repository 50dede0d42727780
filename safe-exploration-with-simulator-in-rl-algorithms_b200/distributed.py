"""One process per GPU plumbing (torch.distributed): rendezvous from the torchrun environment,
contiguous sharding of the N directions over ranks, and the per-iteration record exchange of a sharded
ARS iteration: every rank contributes ONE packed record

    [ returns (2 n_local) | mask (n_local, safe mode only) | count, mean[F], M2[F] (V2 only) ]

and every rank ends up with all of them, in rank order, for the redundant bit-identical ranking, update and
Welford merge.  Two transports:

  p2p         the record is stored straight into every rank's gather buffer over NVLink (CUDA-IPC peer
              pointers) by the pack kernel itself, which also waits for the other ranks' flags and unpacks:
              one launch, no NCCL on the data path, capturable in a CUDA graph (csrc/exchange.cu);
  collective  pack kernel -> torch.distributed all-gather (NCCL on GPUs, gloo in the CPU tests) -> unpack.
              Used when peer access is unavailable and by the host-side tests (`layout`, `all_gather`,
              `split` are device-agnostic torch code).
"""
import ctypes
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


class RecordLayout:
    """Offsets (in doubles) of the packed per-rank record; mirrors pack_exchange_kernel in csrc/exchange.cu."""

    def __init__(self, n_local, n_features=0, has_mask=False):
        self.n_local, self.n_features, self.has_mask = int(n_local), int(n_features), bool(has_mask)
        self.returns = (0, 2 * self.n_local)
        self.mask = (self.returns[1], self.returns[1] + (self.n_local if has_mask else 0))
        self.stats = (self.mask[1], self.mask[1] + (1 + 2 * self.n_features if n_features > 0 else 0))
        self.length = self.stats[1]

    def pack(self, returns_local, mask_local=None, stats_record=None):
        """Host/torch restatement of the kernel's packing (CPU tests, debugging)."""
        parts = [returns_local.to(torch.float64)]
        if self.has_mask:
            parts.append(mask_local.to(torch.float64))
        if self.n_features > 0:
            parts.append(stats_record.to(torch.float64))
        rec = torch.cat(parts)
        assert rec.numel() == self.length
        return rec

    def split(self, gathered, world):
        """-> (returns_all[2N], mask_all[N] int32 or None, records[world, 1+2F] or None) from the gathered
        buffer [world * length] in rank order."""
        g = gathered.view(world, self.length)
        returns = g[:, self.returns[0]:self.returns[1]].reshape(-1)
        mask = g[:, self.mask[0]:self.mask[1]].reshape(-1).ne(0).to(torch.int32) if self.has_mask else None
        records = g[:, self.stats[0]:self.stats[1]] if self.n_features > 0 else None
        return returns, mask, records


class RecordExchange:
    """All-gather of one RecordLayout record per rank and iteration (see the module docstring).

    transport: "auto" (p2p when every rank can map its peers, else collective), "p2p", "collective".
    `use_cuda=False` keeps everything in torch (gloo tests)."""

    def __init__(self, layout, *, group=None, device=None, transport="auto", use_cuda=True, distributed=None):
        self.layout, self.group = layout, group
        use_dist = (dist.is_available() and dist.is_initialized()) if distributed is None else distributed
        self.world = dist.get_world_size(group) if use_dist else 1
        self.rank = dist.get_rank(group) if use_dist else 0
        self.device = torch.device(device) if device is not None else torch.device("cpu")
        self.handle = None
        self.transport = "local" if self.world == 1 else transport
        if self.world > 1:
            self.record = torch.zeros(layout.length, dtype=torch.float64, device=self.device)
            self.gathered = torch.zeros(self.world * layout.length, dtype=torch.float64, device=self.device)
        env = os.environ.get("SWM_EXCHANGE", "")
        if self.transport == "auto":
            self.transport = env if env in ("p2p", "collective") else ("p2p" if use_cuda else "collective")
            if self.transport == "p2p" and not self._open_p2p(strict=False):
                self.transport = "collective"
        elif self.transport == "p2p":
            self._open_p2p(strict=True)

    # ---- p2p: CUDA-IPC peer pointers into every rank's gather buffer ----
    def _open_p2p(self, strict):
        from . import _lib
        L = _lib.lib()
        h = ctypes.c_void_p()
        ok = True
        with torch.cuda.device(self.device):
            rc = L.swm_exchange_create(self.world, self.rank, self.layout.length, ctypes.byref(h))
            mine = (ctypes.c_char * _lib.IPC_HANDLE_BYTES)()
            if rc == 0:
                rc = L.swm_exchange_ipc_handle(h, mine)
            ok = rc == 0
            handles = [None] * self.world
            dist.all_gather_object(handles, bytes(mine) if ok else None, group=self.group)
            if ok and all(x is not None for x in handles):
                table = (ctypes.c_char * (_lib.IPC_HANDLE_BYTES * self.world)).from_buffer_copy(b"".join(handles))
                ok = L.swm_exchange_open_peers(h, table) == 0
            else:
                ok = False
            # every rank must agree, otherwise some would wait in the kernel for peers that use the collective
            flags = [None] * self.world
            dist.all_gather_object(flags, ok, group=self.group)
            ok = all(flags)
        if not ok:
            if h:
                L.swm_exchange_destroy(h)
            if strict:
                raise _lib.SwimmerLibError("peer-memory exchange unavailable: " + L.swm_last_cuda_error().decode())
            return False
        self.handle = h
        return True

    def status(self):
        """(epochs completed, sticky status): status != 0 means a peer did not answer within the kernel's
        time-out (1 + its rank).  Synchronises the device."""
        if self.handle is None:
            return 0, 0
        from . import _lib
        e, s = ctypes.c_uint64(), ctypes.c_uint64()
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().swm_exchange_status(self.handle, ctypes.byref(e), ctypes.byref(s)))
        return e.value, s.value

    def close(self):
        if self.handle is not None:
            from . import _lib
            with torch.cuda.device(self.device):
                _lib.lib().swm_exchange_destroy(self.handle)
            self.handle = None

    @property
    def capturable(self):
        """True when an iteration that uses this exchange may be captured in a CUDA graph."""
        return self.transport in ("local", "p2p")

    # ---- collective: torch.distributed all-gather of the packed record (device-agnostic) ----
    def all_gather(self, record):
        """record [length] -> gathered [world * length] in rank order."""
        if self.world == 1:
            return record
        try:
            dist.all_gather_into_tensor(self.gathered, record, group=self.group)
        except (RuntimeError, NotImplementedError):  # backends without the flat form
            parts = list(self.gathered.view(self.world, -1).unbind(0))
            dist.all_gather(parts, record, group=self.group)
        return self.gathered

    def gather_split(self, record):
        """collective transport end to end: all-gather + split (torch only)."""
        return self.layout.split(self.all_gather(record), self.world)

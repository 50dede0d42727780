"""CPU-only checks: the C-ABI library loads and exports every symbol include/*.h declares, the
binding's struct layouts match the build, and the host-side logic (config dataclasses, database
format, sharding, update-rule table) behaves like the reference's."""
import ctypes
import os
import re
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    names = []
    for h in ("swimmer_ars.h", "swimmer_rlglue_env.h"):
        src = open(os.path.join(ROOT, "include", h)).read()
        names += re.findall(r"^SWM_API[^;(]*?\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", src, flags=re.M)
    return sorted(set(names))


def test_library_exports_every_declared_symbol(S):
    L = ctypes.CDLL(S._lib.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 25 and "swm_rollout" in names and "env_step" in names
    for name in names:
        assert hasattr(L, name), "missing export: " + name
    # RL-Glue's own utility symbols must NOT be exported: -lrlutils defines them in a real RL-Glue build
    for name in ("allocateRLStruct", "clearRLStruct"):
        assert not hasattr(L, name), "would interpose on librlutils: " + name


def test_struct_layouts_match_build(S):
    L = S._lib.lib()
    v = [ctypes.c_int() for _ in range(4)]
    assert L.swm_abi_struct_sizes(*[ctypes.byref(x) for x in v]) == 0
    assert [x.value for x in v] == [ctypes.sizeof(S._lib.SwmParams), ctypes.sizeof(S._lib.SwmPhilox),
                                    ctypes.sizeof(S._lib.SwmScreen), ctypes.sizeof(S._lib.SwmRollout)]
    assert L.swm_abi_version() == S._lib.ABI_VERSION == 3
    assert L.swm_strerror(-2) == b"unsupported configuration"


def test_no_cpu_fallback(S):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    p = S.make_params(n=3)
    with pytest.raises(S.SwimmerLibError):
        S.ops.step_batched(p, torch.zeros(1, 8, dtype=torch.float64), torch.zeros(1, 2, dtype=torch.float64))
    with pytest.raises(S.SwimmerLibError):
        env = S.SwimmerEnv(n=3)
        env.reset()
        env.step([0., 0.])


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "safe-exploration-with-simulator-in-rl-algorithms_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.lower(), (dirpath, f)
    # examples/ and tools/ are not product code either way, but they must not lean on the checker
    for sub in ("examples", "tools"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, sub)):
            for f in files:
                if f.endswith((".py", ".cu", ".sh")):
                    src = open(os.path.join(dirpath, f)).read()
                    assert "import oracle" not in src and "from oracle" not in src, (dirpath, f)


def test_parameters_and_threshold(S):
    g = np.load(os.path.join(ROOT, "tests", "golden", "misc.npz"))
    for K, A, B, H, want in g["threshold_alpha"]:
        assert abs(S.Threshold(K=K, A=A, B=B).compute_alpha(int(H)) - want) <= 1e-12 * abs(want)
    ep = S.EnvParam("x", n=3, H=10, l_i=1., m_i=1., h=1e-3, k=10., epsilon=0.)
    ap = S.ARSParam("a", V1=True, n_iter=1, H=10, N=2, b=1, alpha=.1, nu=.1, safe=False, threshold=0, initial_w="Zero")
    assert (ep.n, ap.N, ap.initial_w) == (3, 2, "Zero")
    assert [f for f in ep.__dataclass_fields__] == ["name", "n", "H", "l_i", "m_i", "h", "k", "epsilon"]


def test_database_roundtrip(S):
    db = S.Database()
    for i in range(3):
        db.add_trajectory(np.full((5, 8), float(i)), np.full((2, 8), -float(i)))
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "d.npz")
        db.save(path)
        z = np.load(path)
        assert set(z.files) == {"policies", "trajectories"} and z["trajectories"].shape == (3, 5, 8)
        db2 = S.Database(); db2.load(path)
        assert db2.size == 3 and np.all(db2.policies[2] == -2.0)


def test_update_rule_table(S):
    assert S.ops.update_args(S.ARS_AGENT, 8, 3) == (True, 8, 3.0, 0)    # all N used, divisor b
    assert S.ops.update_args(S.ARS_TOPB, 8, 3) == (True, 3, 0.0, 0)     # order[:b], divisor len(order)
    assert S.ops.update_args(S.ARS_RLGLUE, 8, 3) == (False, 3, 3.0, 1)  # first b, sample std


def test_swimmer_env_host_surface(S):
    env = S.SwimmerEnv(n=5)
    assert env.observation_space.shape == (12,) and env.action_space.shape == (4,)
    assert env.reset() == [0., 0.] + [np.pi / 2, 0.] * 5
    env.set_state(list(range(12)))
    assert env.get_state() == [float(i) for i in range(12)]
    assert env.get_reward() == 0.0 and env.check_terminal() is False
    e2 = S.make()
    assert e2.n == 5 and e2.envName == "LeonSwimmer-v0"
    r = S.SwimmerEnv(n=3, variant="rlglue")
    assert r.reset() == [0.001] * 8


def _dist_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import sys
    sys.path.insert(0, ROOT)
    import swimmer_ars_b200 as S  # noqa: F401
    from swimmer_ars_b200 import distributed as D
    dist.init_process_group("gloo", rank=rank, world_size=world)
    N, F = 8, 4
    lo, hi = D.shard_directions(N, rank, world)
    # exactly the objects ArsEngine builds for a sharded V2 engine in safe mode, on the collective transport
    layout = D.RecordLayout(hi - lo, n_features=F, has_mask=True)
    ex = D.RecordExchange(layout, device="cpu", transport="collective", use_cuda=False)
    local = torch.arange(2 * lo, 2 * hi, dtype=torch.float64)          # returns of my directions
    mask = torch.tensor([(k + rank) % 2 for k in range(hi - lo)], dtype=torch.int32)
    rec = torch.cat([torch.tensor([float(rank + 1)]), torch.full((F,), float(rank)), torch.full((F,), 10. * rank)])
    ex.record.copy_(layout.pack(local, mask, rec))
    returns, mask_all, records = ex.gather_split(ex.record)
    q.put((rank, ex.transport, ex.capturable, returns.tolist(), mask_all.tolist(), records.tolist()))
    dist.destroy_process_group()


def test_sharded_record_exchange_gloo(S):
    """world_size 2 on CPU (gloo): the engine's record exchange on its collective transport
    (distributed.RecordLayout / RecordExchange, the code ArsEngine._enqueue runs when peer memory is
    unavailable): every rank reconstructs the same 2N returns in direction order, the screening mask and
    the statistics records in rank order."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29650 + os.getpid() % 200
    procs = [ctx.Process(target=_dist_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    got = sorted(q.get(timeout=120) for _ in range(2))
    [p.join(60) for p in procs]
    for rank, transport, capturable, returns, mask_all, records in got:
        assert transport == "collective" and capturable is False
        assert returns == [float(i) for i in range(16)]
        assert mask_all == [0, 1, 0, 1, 1, 0, 1, 0]
        assert records[0][0] == 1.0 and records[1][0] == 2.0 and records[1][1] == 1.0 and records[1][-1] == 10.0
    from swimmer_ars_b200 import distributed as D
    assert D.shard_directions(1024, 3, 8) == (384, 512)
    with pytest.raises(ValueError):
        D.shard_directions(10, 0, 4)
    lay = D.RecordLayout(128, n_features=12, has_mask=False)
    assert (lay.returns, lay.mask, lay.stats, lay.length) == ((0, 256), (256, 256), (256, 281), 281)
    assert lay.length == S._lib.lib().swm_pack_record_doubles(128, 0, 12)
    assert D.RecordLayout(4, 0, True).length == S._lib.lib().swm_pack_record_doubles(4, 1, 0) == 12


def test_chunk_plan_partition(S):
    """Host logic of ops.ChunkedRollout: sub-batches are contiguous, cover the batch exactly, never split
    the environments of one direction, and the time chunks are 64-aligned and sum to H."""
    import itertools
    for B, unit, n_sub, H, chunk in itertools.product((64, 1000, 65536), (1, 2, 8), (1, 3, 16, 5000), (1, 64, 1000),
                                                      (64, 128, 256)):
        if B % unit:
            continue
        subs, lens = S.ops.plan_chunks(B, unit, n_sub, H, chunk)
        assert subs[0][0] == 0 and subs[-1][1] == B and len(subs) <= min(n_sub, B // unit)
        assert all(a[1] == b[0] for a, b in zip(subs, subs[1:]))
        assert all((hi - lo) % unit == 0 and hi > lo for lo, hi in subs)
        assert max(hi - lo for lo, hi in subs) - min(hi - lo for lo, hi in subs) <= unit
        assert sum(lens) == H and all(L == chunk for L in lens[:-1]) and 0 < lens[-1] <= chunk
    import pytest
    with pytest.raises(ValueError):
        S.ops.plan_chunks(128, 1, 4, 1000, 100)   # not a multiple of 64
    with pytest.raises(ValueError):
        S.ops.plan_chunks(130, 4, 4, 1000, 64)    # would split a direction


def test_bench_reference_arm_prints_one_json_line_with_the_shared_config():
    """`bench.py --impl reference` (the CPU arm the driver launches next to the GPU arm): exactly one JSON line
    on stdout, the contract keys, and a `config` dict identical to the GPU arm's (bench.CONFIG)."""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--no-cpu"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True,
                         timeout=300)
    assert out.returncode == 0
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    line = json.loads(lines[0])
    sys.path.insert(0, ROOT)
    import bench
    assert line["impl"] == "reference" and line["config"] == bench.CONFIG and line["metric"] == bench.METRIC
    for key in ("value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "dtype",
                "data", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["value"] > 1e5 and line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0

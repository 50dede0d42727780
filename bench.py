"""bench.py -- swimmer env-steps/s on BASELINE.json config[1] (isolated batched physics: 3-segment
swimmer, 65,536 envs per GPU, fixed random actions, 1,000 explicit-Euler steps), plus the ARS
iteration rate of config[2] as a supplementary key.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

A "step" is one pass of the fused rollout kernel over one batch (65,536 envs x 1,000 steps =
65.5 M env-steps per GPU).  Prints ONE JSON line (rank 0).  See DESIGN.md section "Measurement".
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_SEG, B_PER_GPU, H = 3, 65536, 1000
W_REF = {3: 822, 5: 1751, 10: 5701}        # algorithmic flops per env-step
W_REF_V2 = {3: 838, 5: 1775, 10: 5745}
# Measured once per kernel change with ncu (profiles/r01c_summary.md), config[1] kernel, per launch:
NCU_DRAM_BYTES_PER_LAUNCH = 1.07e6          # dram__bytes_read.sum + dram__bytes_write.sum
NCU_EXEC_FLOPS_PER_ENV_STEP = 312.0         # executed FP64 flops per env-step: 2 per DFMA, 1 per DADD/DMUL (ncu source page)
METRIC, UNIT = "swimmer env-steps/sec", "env-steps/s"
WORKLOAD = "config[1]: 3-segment swimmer, 65,536 envs per GPU, fixed random actions U(-5,5), 1,000 explicit-Euler steps"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-ars", action="store_true", help="skip the supplementary ARS-iteration measurement")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------
# CPU legs (oracle port, all host threads).  The only place bench.py executes oracle/.
# ----------------------------------------------------------------------------------------------
def cpu_fixed_action_rate(n_envs, steps, threads):
    """env-steps/s of the C port of the reference dynamics on `threads` host threads."""
    from oracle import oracle_lib as O
    p = O.make_params(n=N_SEG)
    actions = np.random.default_rng(0).uniform(-5, 5, (n_envs, N_SEG - 1))
    O.lib()
    bounds = np.linspace(0, n_envs, threads + 1).astype(int)
    out = [None] * threads

    def work(i):
        out[i] = O.rollout_fixed_batch(p, O.GYM, actions, steps, int(bounds[i]), int(bounds[i + 1]),
                                       want_final=False)[0]
    t0 = time.perf_counter()
    ts = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    dt = time.perf_counter() - t0
    return n_envs * steps / dt, dt


def cpu_sample_size(threads, target_s):
    rate1, _ = cpu_fixed_action_rate(64, 200, 1)   # calibration: ~12.8k steps on one thread
    envs = int(max(threads, min(B_PER_GPU, rate1 * threads * target_s / H)))
    return max(threads, (envs // threads) * threads)


def ref_cpp_rate(steps=20000):
    """Secondary: the unmodified reference C++ swimmer (oracle/_ref, RL-Glue variant, Eigen QR +
    redundant inverse), one core, no printing.  Different dynamics variant: reported, not compared."""
    from oracle import oracle_lib as O
    if O.ref_cpp() is None:
        return None
    p = O.make_params(n=N_SEG, h=0.01)
    O.ref_cpp_set_params(p)
    st = np.full(2 * N_SEG + 2, 0.001)
    a = np.array([1.5, -2.0])
    t0 = time.perf_counter()
    O.ref_cpp().ref_rollout_fixed(st.ctypes.data_as(O._dp), a.ctypes.data_as(O._dp), steps)
    return steps / (time.perf_counter() - t0)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    # bounded sample per step: the whole run (K timed + W warm-up steps) stays near two minutes
    per_step_s = min(4.0, max(0.25, 100.0 / max(1, args.steps + args.warmup)))
    envs = cpu_sample_size(threads, per_step_s)
    for _ in range(args.warmup):
        cpu_fixed_action_rate(max(threads, envs // 8), H, threads)
    t_tot, done = 0.0, 0
    for _ in range(args.steps):
        rate, dt = cpu_fixed_action_rate(envs, H, threads)
        t_tot += dt
        done += envs * H
    value = done / t_tot
    sample = "%d envs x %d steps per step (of %d), %d host threads" % (envs, H, B_PER_GPU, threads)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "C port of the reference gym swimmer (oracle/swimmer_oracle.c, dense (n+2) formulation of "
                "remy_swimmer_env.py) -- the Python reference itself cannot travel to the GPU box; it ran at "
                "~5.5k env-steps/s/core in the build container (BASELINE.md section 2)",
    }
    cpp = ref_cpp_rate()
    if cpp:
        line["ref_cpp_rlglue_1core"] = {"value": cpp, "unit": UNIT, "kind": "reference",
                                        "note": "unmodified rlglue/environment/SwimmerEnvironment.cpp updateState, n=3"}
    emit(line)


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.proc, self.rows = None, []
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(index), "--query-gpu=" + self.FIELDS,
                 "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            c = [x.strip() for x in r.split(",")]
            if len(c) < 7:
                continue
            try:
                sm.append(float(c[0])); mx = float(c[1])
            except ValueError:
                continue
            for nm, v in zip(names, c[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        # idle samples (before the first launch) sit at the idle clock: take the median of the upper half
        hot = sorted(sm)[len(sm) // 2:] if sm else []
        return {"sm_mhz": float(np.median(hot)) if hot else None, "sm_max_mhz": mx,
                "reasons": sorted(reasons), "samples": len(sm),
                "note": "nvidia-smi -lms 100 from before warm-up to the end of the e2e region, same kernel "
                        "kept running until >=5 samples; median of the upper half of samples"}


def run_b200(args):
    import torch
    import torch.distributed as dist
    import swimmer_ars_b200 as S
    from swimmer_ars_b200 import distributed as D

    rank, world, device = D.init_from_env()
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    if world != args.gpus and rank == 0:
        print("warning: --gpus %d but WORLD_SIZE=%d" % (args.gpus, world), file=sys.stderr)
    params = S.make_params(n=N_SEG)
    rng = np.random.default_rng(rank)
    host_actions = torch.as_tensor(rng.uniform(-5, 5, (B_PER_GPU, N_SEG - 1))).pin_memory()
    actions = host_actions.to(device)
    out = {"returns": torch.empty(B_PER_GPU, dtype=torch.float64, device=device),
           "final_state": torch.empty(B_PER_GPU, 2 * N_SEG + 2, dtype=torch.float64, device=device)}
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=device)
    stream = torch.cuda.current_stream()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def single_launch():
        return S.ops.rollout(params, H, actions=actions, want_final=True, out=out)

    # The timed step: the same batch scheduled as 16 sub-batches x 64-step chunks on 16 streams
    # (ops.ChunkedRollout: swm_rollout launches chained through final_state -> init_state, replayed from one
    # CUDA graph).  Every final state is bit-identical to the single launch; short launches from several
    # streams remove the quantisation of 65,536 envs over the SM sub-partitions (3.46 warps each).
    plan = S.SwimmerEnv(n=N_SEG, device=device).rollout_plan(H, n_sub=16, chunk=64, actions=actions)
    plan.run()
    torch.cuda.synchronize()
    step_graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(step_graph):
        plan.run()
    ref_res = single_launch()
    step_graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(plan.state, ref_res.final_state), "chunked schedule must reproduce the single launch"

    def one_step():
        step_graph.replay()

    # ---- FP64 roofline denominator: DFMA probe measured live on this GPU ----
    sink = torch.zeros(8, dtype=torch.float64, device=device)
    sms = torch.cuda.get_device_properties(device).multi_processor_count
    S.ops.fp64_probe(sms * 8, 256, 200, sink)
    best = 0.0
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        flops = S.ops.fp64_probe(sms * 8, 256, 4000, sink)
        e1.record(stream)
        e1.synchronize()
        best = max(best, flops / (e0.elapsed_time(e1) * 1e-3))
    fp64_peak_tflops = best / 1e12

    sampler = ClockSampler(device.index) if rank == 0 else None
    for _ in range(max(args.warmup, 3)):
        one_step()
    barrier()
    # ---- device-resident timing: K steps, each its own event pair, L2 flushed in between ----
    evs = []
    t_wall0 = time.perf_counter()
    for _ in range(args.steps):
        flush.fill_(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        one_step()
        e1.record(stream)
        evs.append((e0, e1))
    barrier()
    wall = time.perf_counter() - t_wall0
    ms = [a.elapsed_time(b) for a, b in evs]
    t_dev = torch.tensor([sum(ms)], dtype=torch.float64, device=device)
    # ---- supplementary: one plain swm_rollout launch per step (no chunking), same timing method ----
    evs1 = []
    for _ in range(args.steps):
        flush.fill_(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        single_launch()
        e1.record(stream)
        evs1.append((e0, e1))
    barrier()
    t_single = torch.tensor([sum(a.elapsed_time(b) for a, b in evs1)], dtype=torch.float64, device=device)

    # ---- supplementary: the same K launches with TWO in flight (two streams, device-resident inputs, no
    # flush).  One config[1] batch is only 3.46 warps per SM sub-partition; two batches fill the FP64 pipe. ----
    side = [torch.cuda.Stream(device=device) for _ in range(2)]
    outs2 = [{"returns": torch.empty_like(out["returns"]), "final_state": torch.empty_like(out["final_state"])}
             for _ in range(2)]
    barrier()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record(stream)
    for st in side:
        st.wait_stream(stream)
    for k in range(args.steps):
        with torch.cuda.stream(side[k & 1]):
            S.ops.rollout(params, H, actions=actions, want_final=True, out=outs2[k & 1])  # plain launches
    for st in side:
        stream.wait_stream(st)
    c1.record(stream)
    barrier()
    t_conc = torch.tensor([c0.elapsed_time(c1)], dtype=torch.float64, device=device)

    # ---- end-to-end through the host-buffer entry point: every step uploads its actions from pinned host
    # memory and downloads returns + final states into pinned host memory (SwimmerEnv.rollout_batched_host,
    # double-buffered on two streams: the copies of one step overlap the kernel of its neighbour).  No L2
    # flush here: the inputs are re-uploaded from the host on every step. ----
    host_ret = [torch.empty(B_PER_GPU, dtype=torch.float64).pin_memory() for _ in range(2)]
    host_fin = [torch.empty(B_PER_GPU, 2 * N_SEG + 2, dtype=torch.float64).pin_memory() for _ in range(2)]
    env = S.SwimmerEnv(n=N_SEG, device=device)

    def e2e_step(k):
        return env.rollout_batched_host(H, host_actions, host_ret[k & 1], host_fin[k & 1])
    for k in range(3):
        e2e_step(k)
    env.synchronize_host()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for k in range(args.steps):
        e2e_step(k)
    env.synchronize_host()
    e1.record(stream)
    barrier()
    t_e2e = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=device)
    e2e_check = float(host_ret[(args.steps - 1) & 1].sum())  # the host really holds the results
    assert np.isfinite(e2e_check)
    # nvidia-smi needs ~100 ms per sample: if the timed regions above were too short to be sampled,
    # keep the same kernel running (untimed) until a few samples under load exist.
    if sampler:
        t_end = time.perf_counter() + 3.0
        n0 = len(sampler.rows)
        while len(sampler.rows) < n0 + 5 and time.perf_counter() < t_end:
            for _ in range(20):
                one_step()
            torch.cuda.synchronize()
    clocks = sampler.stop() if sampler else None
    if world > 1:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
        dist.all_reduce(t_conc, op=dist.ReduceOp.MAX)
        dist.all_reduce(t_single, op=dist.ReduceOp.MAX)
    t_dev_s, t_e2e_s = float(t_dev.cpu()[0]) * 1e-3, float(t_e2e.cpu()[0]) * 1e-3
    total_steps = float(world) * B_PER_GPU * H * args.steps
    value = total_steps / t_dev_s
    e2e_value = total_steps / t_e2e_s
    conc_value = total_steps / (float(t_conc.cpu()[0]) * 1e-3)
    single_value = total_steps / (float(t_single.cpu()[0]) * 1e-3)

    # ---- supplementary: ARS iterations/s (rollouts + NCCL exchange + ranking + update) ----
    #   config[2]: ARS V2, n=5, 1,024 directions, H=1000  (2,048 envs in total: latency-bound)
    #   config[4]: n=10, 4,096 directions x 2 x 128 rollouts = 1,048,576 envs, H=1000 (throughput-bound)
    ars = None
    if not args.no_ars:
        ars = []
        for tag, n_, Ndir, R_, K2, desc in (
                ("config[2]", 5, 1024, 1, 5, "ARS V2, 5-segment swimmer, 1,024 directions (2,048 rollouts), H=1000"),
                ("config[4]", 10, 4096, 128, 2, "ARS V2, 10-segment swimmer, 4,096 directions x 2 x 128 rollouts "
                                                "(1,048,576 envs, reset + 1e-2*U[0,1) starts), H=1000")):
            if Ndir % world != 0:
                continue
            eng = S.ArsEngine(S.make_params(n=n_), N=Ndir, b=Ndir, alpha=0.0075, nu=0.01, H=1000, v2=True,
                              semantics=S.ARS_AGENT, seed=0, device=device, rollouts_per_direction=R_,
                              init_perturb=1e-2 if R_ > 1 else 0.0, use_graph=True)
            for _ in range(3 if R_ == 1 else 2):  # eager warm-up, graph capture (1 GPU), replay
                eng.run_iteration()
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(K2):
                eng.run_iteration()
            e1.record(stream)
            barrier()
            t_ars = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=device)
            if world > 1:
                dist.all_reduce(t_ars, op=dist.ReduceOp.MAX)
            t_ars_s = float(t_ars.cpu()[0]) * 1e-3
            steps_per_iter = 2.0 * Ndir * R_ * 1000
            ars.append({"workload": "%s: %s; directions sharded over %d GPU(s) (strong scaling)" % (tag, desc, world),
                        "iters_per_s": K2 / t_ars_s, "env_steps_per_s": K2 * steps_per_iter / t_ars_s,
                        "ms_per_iter": 1e3 * t_ars_s / K2, "iters_timed": K2,
                        "roofline_frac_fp64": (K2 * steps_per_iter / t_ars_s / world) * W_REF_V2[n_] / 1e12 / fp64_peak_tflops,
                        "launch": "CUDA graph replay" if eng._graph is not None else "eager (NCCL exchange)",
                        "mean_return_last": float(eng.returns.mean().cpu())})
            del eng
        if world == 1:
            # seed fan-out (ars/experiment.py:64-72: one agent per seed): 32 agents of config[0]
            # (n=3, V1, 8 directions, H=1000), one CUDA stream + one graph launch per agent-iteration
            seeds, K3 = 32, 20
            fan = S.SeedFanout(S.make_params(n=3), range(seeds), N=8, b=8, alpha=0.0075, nu=0.01, H=1000, device=device)
            fan.run(2)
            barrier()
            t0 = time.perf_counter()
            fan.run(K3, include_initial=False)
            dt = time.perf_counter() - t0
            ars.append({"workload": "config[0] x %d seeds: ARS V1, 3-segment swimmer, 8 directions, H=1000, one agent per "
                                    "seed on one GPU (seed fan-out)" % seeds,
                        "agent_iters_per_s": seeds * K3 / dt, "env_steps_per_s": seeds * K3 * 16 * 1000 / dt,
                        "ms_per_round": 1e3 * dt / K3, "timing": "host wall clock around %d rounds incl. final sync" % K3})
            del fan

    if rank == 0:
        cpu = None
        if not args.no_cpu:
            threads = os.cpu_count() or 1
            envs = cpu_sample_size(threads, 12.0)
            rate, dt = cpu_fixed_action_rate(envs, H, threads)
            cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
                   "sample": "%d envs x %d steps of the same workload in %.1f s (oracle C port of the "
                             "reference gym swimmer)" % (envs, H, dt)}
        ms_per_step = 1e3 * t_dev_s / args.steps
        achieved = (value / world) * W_REF[N_SEG] / 1e12
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "n_segments": N_SEG, "envs_per_gpu": B_PER_GPU, "H": H,
                       "l2": "flushed between steps (256 MiB write outside the timed events); inputs 1 MiB",
                       "schedule": "each step = one CUDA-graph launch of %d swm_rollout kernels: 16 sub-batches x "
                                   "64-step chunks on 16 streams, state chained through final_state -> init_state, "
                                   "final states bit-identical to a single launch" % plan.launches,
                       "timing": "CUDA events per step on the launch stream, summed; max over ranks"},
            "roofline": {"bound": "fp64", "achieved": achieved, "peak": fp64_peak_tflops, "unit": "TFLOP/s",
                         "frac": achieved / fp64_peak_tflops, "traffic": NCU_DRAM_BYTES_PER_LAUNCH,
                         "traffic_unit": "bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum, "
                                         "profiles/r01c_n3_fixed_ncu.csv): ~0 B per env-step, not HBM-bound",
                         "peak_source": "DFMA probe kernel measured in this run (MEASURED_PEAKS.json has no FP64 "
                                        "entry; nominal 37.2 TFLOP/s)",
                         "flops_per_env_step": W_REF[N_SEG],
                         "executed": {"flops_per_env_step": NCU_EXEC_FLOPS_PER_ENV_STEP,
                                      "achieved": (value / world) * NCU_EXEC_FLOPS_PER_ENV_STEP / 1e12,
                                      "frac": (value / world) * NCU_EXEC_FLOPS_PER_ENV_STEP / 1e12 / fp64_peak_tflops,
                                      "note": "FP64 flops the O(n) kernel really executes (ncu instruction counts) against "
                                              "the same DFMA peak; on B200 a DFMA with three register sources issues "
                                              "every 3 cycles, so the pipe saturates below 1.0 (profiles/r01b_summary.md)"},
                         "note": "achieved = per-GPU env-steps/s x 822 algorithmic flops/env-step of the reference's "
                                 "dense formulation (SURVEY 8d); the O(n) kernel executes fewer real flops"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B_PER_GPU * (N_SEG - 1) * 8,
                    "d2h_bytes_per_step": B_PER_GPU * (2 * N_SEG + 3) * 8,
                    "api": "SwimmerEnv.rollout_batched_host: pinned host actions in, pinned host returns + final states "
                           "out every step, double-buffered on two streams; CUDA events around all K steps"},
            "single_launch": {"value": single_value, "unit": UNIT,
                              "note": "one plain swm_rollout launch per step (the whole batch in one kernel, no "
                                      "chunking), same events / flush as `value`"},
            "two_in_flight": {"value": conc_value, "unit": UNIT, "streams": 2,
                              "roofline_frac_executed": (conc_value / world) * NCU_EXEC_FLOPS_PER_ENV_STEP / 1e12 / fp64_peak_tflops,
                              "note": "same kernel, same inputs resident in HBM, K launches alternating on two streams "
                                      "(no L2 flush): one 65,536-env batch is 3.46 warps per SM sub-partition (the busiest "
                                      "holds 4), two batches in flight balance and fill the FP64 pipe; this is also why "
                                      "the double-buffered e2e number exceeds the one-launch-at-a-time `value`"},
            "gpu_launches": args.steps * plan.launches, "clocks": clocks, "cpu_baseline": cpu,
            "wall_s_timed_region": wall,
        }
        if ars:
            line["ars"] = ars
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit(line):
    """The ONE JSON line goes to the real stdout; everything else any library prints there while the run
    is in progress (e.g. NCCL's version banner) has been redirected to stderr."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    global _REAL_STDOUT
    args = parse()
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()

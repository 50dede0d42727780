#!/bin/bash
# Final-build captures of round 2 (run under gpurun; every ncu pass follows a plain run of the same command):
#   per case: CUDA-event time (plain), then `ncu --set full` of one launch of the dominant kernel;
#   the timed bench step as ONE graph (ncu --graph-profiling graph: the concurrent chunked schedule measured as
#   a unit); the launch list of the bench command.
set -u
mkdir -p gpurun_out
run_case() {  # name, kernel regex, SWM_PROFILE_KERNEL
  export SWM_PROFILE_KERNEL=$3
  python tools/profile_cases.py $1 3 > gpurun_out/r02z_$1_k$3.txt 2>&1 || return
  ncu --set full --clock-control none --import-source on -k regex:$2 -s 2 -c 1 -o gpurun_out/r02z_$1_k$3 \
      python tools/profile_cases.py $1 3 > gpurun_out/r02z_$1_k$3.ncu.log 2>&1
  cat gpurun_out/r02z_$1_k$3.txt
}
run_case n3_fixed rollout_kernel 0
# the same batch as ONE forced plain launch (the kernel the roofline is quoted for)
python tools/profile_cases.py n3_fixed_plain 3 > gpurun_out/r02z_n3_fixed_plain.txt 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -s 2 -c 1 -o gpurun_out/r02z_n3_fixed_plain \
      python tools/profile_cases.py n3_fixed_plain 3 > gpurun_out/r02z_n3_fixed_plain.ncu.log 2>&1
cat gpurun_out/r02z_n3_fixed_plain.txt
run_case n5_v2 lane_rollout_kernel 0
run_case n5_v2_1024 lane2_rollout_kernel 0
run_case n5_v2_256 lane2_rollout_kernel 0
run_case n10_grp rollout_kernel 0
run_case n3_safe rollout_kernel 0
unset SWM_PROFILE_KERNEL
python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r02z_bench_steps3.json 2> gpurun_out/r02z_bench_steps3.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02z_launches_bench_steps3.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r02z_ncu_bench.log 2>&1
ncu --graph-profiling graph --clock-control none -c 4 \
    --metrics gpu__time_duration.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_fp64.sum,smsp__issue_active.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum \
    --csv --log-file gpurun_out/r02z_graph_bench_step.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-ars --no-sustained > gpurun_out/r02z_ncu_graph.log 2>&1
tail -3 gpurun_out/r02z_ncu_graph.log
# the timed step (256 chunk launches on 16 streams, one graph) measured as one unit inside an NVTX range
python tools/graph_step_profile.py > gpurun_out/r02z_graph_step.txt 2>&1
ncu --graph-profiling graph --nvtx --nvtx-include "bench_step/" --clock-control none \
    --metrics gpu__time_duration.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_fp64.sum,smsp__issue_active.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum \
    --csv --log-file gpurun_out/r02z_graph_step.csv python tools/graph_step_profile.py > gpurun_out/r02z_graph_step.log 2>&1
cat gpurun_out/r02z_graph_step.txt

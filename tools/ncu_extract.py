"""Reduces an .ncu-rep to the metrics DESIGN.md / profiles/*.md quote (csv: metric,unit,value per launch).

  python tools/ncu_extract.py <report.ncu-rep> <out.csv>
"""
import csv
import io
import subprocess
import sys

KEEP = (
    "Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__cycles_elapsed.max", "sm__cycles_active.avg",
    "sm__cycles_elapsed.avg.per_second",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fp64_cycles_active.max.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.min.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_active.avg.per_cycle_active",
    "smsp__warps_eligible.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__sass_branch_targets_threads_divergent.sum",
    "smsp__sass_branch_targets_threads_uniform.pct", "smsp__sass_branch_targets.sum",
)


def main():
    rep, out = sys.argv[1], sys.argv[2]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, launches = rows[0], rows[1], rows[2:]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + ["launch%d" % i for i in range(len(launches))])
        for i, h in enumerate(hdr):
            if h in KEEP or (h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")):
                w.writerow([h, units[i]] + [l[i] for l in launches])


if __name__ == "__main__":
    main()

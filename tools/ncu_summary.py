"""Summarise .ncu-rep captures (ncu --set full) as a small CSV for profiles/: one row per (kernel, metric).

    python tools/ncu_summary.py OUT.csv "comment" REP [REP ...]
"""
import csv
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "launch__grid_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct",
    "smsp__average_warp_latency_per_inst_issued.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
]


def main(out, comment, reps):
    rows = [["capture", "metric", "unit", "value"], ["# " + comment, "", "", ""]]
    for rep in reps:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        table = list(csv.reader(raw.splitlines()))
        head, units = table[0], table[1]
        name_col = head.index("Kernel Name")
        for rec in table[2:]:
            name = rec[name_col].split("(")[0].replace("void ", "").replace("tamtr::", "")
            for m in METRICS:
                if m in head:
                    i = head.index(m)
                    rows.append([name, m, units[i], rec[i]])
    with open(out, "w", newline="") as f:
        csv.writer(f).writerows(rows)
    print("wrote", out, len(rows) - 2, "rows")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3:])

"""Print the handful of ncu metrics that decide what bounds a kernel (issue, pipes, shared memory, DRAM, stalls).

    python tools/ncu_brief.py REP.ncu-rep
"""
import csv
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "smsp__average_warp_latency_per_inst_issued.ratio",
]


def main(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    t = list(csv.reader(raw.splitlines()))
    head, units = t[0], t[1]
    for rec in t[2:]:
        print("==", rec[head.index("Kernel Name")][:90])
        for i, h in enumerate(head):
            stall = "smsp__average_warps_issue_stalled" in h and h.endswith("per_issue_active.ratio")
            if h in WANT or stall:
                try:
                    v = float(rec[i].replace(",", ""))
                except ValueError:
                    continue
                if stall and v < 0.1:
                    continue
                print(f"  {h.replace('smsp__average_warps_issue_stalled_', 'stall:'):88s} {units[i]:16s} {rec[i]}")


if __name__ == "__main__":
    main(sys.argv[1])

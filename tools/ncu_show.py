#!/usr/bin/env python
"""Print the counters that matter from `ncu --page raw --csv` exports (gpurun_out/<tag>_<kernel>_raw.csv)."""
import csv
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def main():
    for path in sys.argv[1:]:
        rows = list(csv.reader(open(path)))
        h, u = rows[0], rows[1]
        for r in rows[2:]:
            print("## %s :: %s" % (path, r[h.index("Kernel Name")][:70]))
            for w in WANT:
                if w in h:
                    print("  %-72s %16s %s" % (w, r[h.index(w)], u[h.index(w)]))
            for i, name in enumerate(h):
                if "issue_stalled" in name and name.endswith("per_issue_active.ratio") or ("pipe" in name and "tensor" in name and "pct" in name):
                    try:
                        v = float(r[i].replace(",", ""))
                    except ValueError:
                        continue
                    if v >= 0.3:
                        print("  %-72s %16s %s" % (name.replace("smsp__average_warps_issue_stalled_", "stall:").replace("_per_issue_active.ratio", ""), r[i], u[i]))


if __name__ == "__main__":
    main()

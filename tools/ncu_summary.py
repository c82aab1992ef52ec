"""Summarises ncu artefacts brought back in gpurun_out/ into small text files for profiles/.

  python tools/ncu_summary.py launches <launches.csv>            # per-kernel share of the step
  python tools/ncu_summary.py full <file.ncu-rep> [...]          # key counters of `ncu --set full` captures
"""
import collections
import csv
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "lts__t_bytes.sum", "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = next(r for r in rows if r[0] == "ID")
    data = rows[rows.index(hdr) + 1:]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in data:
        v = float(r[vi].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui], 1e-6)
        a = agg.setdefault(r[ki].split("(")[0], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print("# %s: %d launches, %.2f ms of kernel time (ncu-serialised, cold cache: compare shares)" % (path, len(data), tot))
    print("%-58s %8s %12s %7s %12s" % ("kernel", "launches", "total ms", "share", "avg ms"))
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-58s %8d %12.3f %6.1f%% %12.4f" % (k[:58], c, t, 100 * t / tot, t / c))


def full(paths):
    for p in paths:
        out = subprocess.run(["ncu", "-i", p, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        if len(rows) < 3:
            print("# %s: no data" % p)
            continue
        h, units = rows[0], rows[1]
        for r in rows[2:]:
            name = r[h.index("Kernel Name")] if "Kernel Name" in h else "?"
            print("# %s :: %s" % (p, name[:100]))
            for w in WANT:
                if w in h:
                    i = h.index(w)
                    print("  %-70s %18s %s" % (w, r[i], units[i]))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        full(sys.argv[2:])

#!/usr/bin/env python
"""profiles/traffic.json from `ncu --set full` raw-page exports: measured DRAM bytes per launch of the main kernel of
every kernel class bench.py reports (roofline.traffic).  usage: tools/make_traffic.py <tag> (reads gpurun_out/<tag>_*_raw.csv)"""
import csv
import re
import glob
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLASS_OF = {"ntt_fwd_digits_kernel": "relinearize", "r32_digits_kernel": "relinearize_u32", "ntt_inv_tensor_kernel": "square_tensor", "ntt_inv_kernel": "ntt_inverse", "ntt_fwd_kernel": "ntt_forward",
            "tc_mac_kernel": "weighted_sum_tc_i8", "tcn2_mac_kernel": "weighted_sum_tcn_i8", "tcn_mac_kernel": "weighted_sum_tcn_i8_rowmajor", "tcn_split_kernel": "tcn_plane_split",
            "behz_floor_kernel": "behz_floor_sk", "pool_kernel": "pool_sum", "bn_kernel": "batch_norm", "mac_kernel": "weighted_sum_mac"}
# ncu counters bench.py prints next to every roofline fraction (percent of peak while the kernel was active)
COUNTERS = {"pipe_tensor_pct": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "pipe_fma_pct": "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
            "pipe_fmaheavy_cycles_pct": "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",   # IMAD runs here: the binding pipe of the transforms
            "sm_throughput_pct": "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "pipe_alu_pct": "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
            "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
            "dram_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "l2_hit_pct": "lts__t_sector_hit_rate.pct"}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    tag = sys.argv[1]
    out, counters = {}, {}
    for path in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", tag + "_*_raw.csv"))):
        rows = list(csv.reader(open(path)))
        h, u = rows[0], rows[1]
        for r in rows[2:]:
            name = r[h.index("Kernel Name")]
            key = next((k for k in CLASS_OF if re.search(r"(::|\s|^)%s[<(]" % k, name)), None)
            if key is None:
                continue
            tot = 0.0
            for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                i = h.index(m)
                tot += float(r[i].replace(",", "")) * UNIT.get(u[i], 1.0)
            d = float(r[h.index("gpu__time_duration.sum")].replace(",", ""))
            cls = CLASS_OF[key]
            cn = counters.setdefault(cls, {"kernel": key})
            for short, m in COUNTERS.items():
                if m in h and r[h.index(m)] not in ("", "n/a"):
                    cn[short] = round(float(r[h.index(m)].replace(",", "")), 2)
            ent = out.setdefault(CLASS_OF[key], {"kernel": key, "launches": []})
            ent["launches"].append({"dram_bytes": tot, "duration": d, "duration_unit": u[h.index("gpu__time_duration.sum")],
                                    "grid": r[h.index("launch__grid_size")]})
    for ent in out.values():
        ent["dram_bytes_per_launch"] = sum(l["dram_bytes"] for l in ent["launches"]) / len(ent["launches"])
    json.dump({"source": "ncu --set full --clock-control none, one forward at batch 8 (tools/gpu_profile.sh %s)" % tag, "classes": out,
               "counters": counters},
              open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
    print(json.dumps({k: v["dram_bytes_per_launch"] for k, v in out.items()}, indent=1))


if __name__ == "__main__":
    main()

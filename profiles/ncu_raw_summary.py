#!/usr/bin/env python
"""Pick the roofline-relevant metrics out of `ncu -i X.ncu-rep --page raw --csv` (one line per profiled launch).
usage: ncu -i X.ncu-rep --page raw --csv | python profiles/ncu_raw_summary.py [out.json]"""
import csv
import json
import sys

KEYS = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warp_latency_per_inst_issued.ratio", "sm__cycles_elapsed.max"]
rows = list(csv.reader(sys.stdin))
hdr, units = rows[0], rows[1]
out = []
for r in rows[2:]:
    d = {}
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            d[k] = r[i] if k == "Kernel Name" else f"{r[i]} {units[i]}".strip()
    out.append(d)
for d in out:
    print(json.dumps(d, indent=1))
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)

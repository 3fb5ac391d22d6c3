#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` dump: stall mix and the hottest SASS instructions.
usage: ncu -i X.ncu-rep --page source --csv > src.csv ; python profiles/ncu_source_summary.py src.csv [top]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr, data, kernels = None, [], 0
for r in rows:
    if r and r[0] == "Kernel Name":
        kernels += 1
        if kernels > 1:
            break
        print("kernel:", r[1][:100])
        continue
    if r and r[0] == "Address":
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        data.append(r)
ix = {h: i for i, h in enumerate(hdr)}
tot = sum(int(r[ix["# Samples"]]) for r in data)
print("static instructions:", len(data), " samples:", tot, " warp-instructions executed:", sum(int(r[ix["Instructions Executed"]]) for r in data))
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {s: sum(int(r[ix[s]]) for r in data) for s in stalls}
for s, v in sorted(agg.items(), key=lambda x: -x[1])[:9]:
    print(f"  {s:24s} {v:7d} {100 * v / tot:5.1f}%")
print("hottest instructions (samples, executed, SASS):")
for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:top_n]:
    print(f"  {r[ix['# Samples']]:>6s} {r[ix['Instructions Executed']]:>9s}  {r[ix['Source']][:100]}")

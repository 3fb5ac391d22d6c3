#!/bin/bash
# GPU session: duration of k_compact_leaves / k_expand_select under ncu, direct vs staged list writes
mkdir -p gpurun_out
for st in 0 1; do
AZ_COMPACT_STAGE=$st timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 6000 -c 120 --csv --log-file gpurun_out/l_$st.csv python bench.py --burn-in 2 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --extras none > gpurun_out/ncu_l.log 2>&1
python - <<PY
import csv, collections, re
rows=list(csv.reader(open('gpurun_out/l_$st.csv')))
for i,r in enumerate(rows):
    if 'Kernel Name' in r: h=r; start=i; break
ki=h.index('Kernel Name'); vi=h.index('Metric Value')
tot=collections.Counter(); cnt=collections.Counter()
for r in rows[start+1:]:
    if len(r)<=vi: continue
    m=re.search(r'(k_\w+)', r[ki]); name=m.group(1) if m else r[ki][:30]
    try: v=float(r[vi].replace(',',''))
    except: continue
    tot[name]+=v; cnt[name]+=1
print('stage $st', {k:(cnt[k], round(v/cnt[k]/1e3,2)) for k,v in tot.items()})
PY
done

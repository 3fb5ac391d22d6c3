#!/bin/bash
# GPU session: fused leaf list with ONE 64-bit atomic per warp (no fences / block barrier): tests, A/B against the extra launch,
# and the per-launch durations of the tree kernels under ncu for both
mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_compaction.py -x -q > gpurun_out/pytest_compact.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_compact.log); tail -5 gpurun_out/pytest_compact.log
for rep in 0 1; do for f in 0 1; do
(AZ_COMPACT_FUSED=$f timeout 200 python bench.py --burn-in 14 --steps 6 --warmup 3 --no-e2e --no-cpu-baseline --extras none > gpurun_out/bench_f${f}_$rep.json 2> gpurun_out/bench_f${f}_$rep.err; echo "fused=$f rep=$rep rc=$?")
python -c "
import json; d=json.load(open('gpurun_out/bench_f${f}_$rep.json'))
print('%.4e sims/s  %.2f ms/step  us/simstep %.1f  kernel_ms %.4f  launches %d' % (d['value'], d['ms_per_step'], d['details']['us_per_simulation_step'], d['roofline']['kernel_ms'], d['gpu_launches']))"
done; done
for f in 0 1; do
AZ_COMPACT_FUSED=$f timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 6000 -c 120 --csv --log-file gpurun_out/lf_$f.csv python bench.py --burn-in 2 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --extras none > gpurun_out/ncu_lf.log 2>&1
python - <<PY
import csv, collections, re
rows=list(csv.reader(open('gpurun_out/lf_$f.csv')))
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
print('fused $f', {k:(cnt[k], round(v/cnt[k]/1e3,2)) for k,v in tot.items()})
PY
done

#!/bin/bash
mkdir -p gpurun_out
(timeout 300 python bench.py --burn-in 16 --steps 6 --no-cpu-baseline --no-e2e --extras resnet4x64:fp16,basic > gpurun_out/bench_s.json 2> gpurun_out/bench_s.err; echo "bench rc=$?"); tail -1 gpurun_out/bench_s.err | cut -c1-200
python -c "
import json; d=json.load(open('gpurun_out/bench_s.json'))
r=d['roofline']
print(d['value'], d['ms_per_step'], r['kernel_ms'], r['kernel_ms_events_ungraphed_step'], r['kernel_in_timed_region'], r['frac'], r['kernel_share_of_step'], r['tree_kernel']['us_per_launch'], d['clocks']['sm_mhz'])
for x in d['net_in_loop']: print(x['evaluator'], x['workload'], '%.3e'%x['sims_per_s'], round(x['evaluator_kernel_us'],1), round(x['us_per_sim_step'],1), round(x['tensor_frac_of_measured_bf16'],3))
print(d.get('errors'))"

#!/bin/bash
# GPU session: k_mlp_fused with the FADD2 / F2FP.RELU epilogue - in the loop
mkdir -p gpurun_out
(timeout 300 python bench.py --burn-in 8 --steps 3 --no-cpu-baseline --no-e2e --extras basic,basic:fp16 > gpurun_out/bench_s.json 2> gpurun_out/bench_s.err; echo "bench rc=$?"); tail -1 gpurun_out/bench_s.err | cut -c1-200
python -c "
import json; d=json.load(open('gpurun_out/bench_s.json'))
for x in d['net_in_loop']: print(x['evaluator'], x['workload'], '%.3e'%x['sims_per_s'], round(x['evaluator_kernel_us'],1), round(x['us_per_sim_step'],1), round(x['tensor_frac_of_measured_bf16'],3))
print(d.get('errors'))"

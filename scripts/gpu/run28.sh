#!/bin/bash
# GPU session: k_resnet_wide as the default 64-channel kernel - whole GPU suite, then the headline bench with fp32 / fp16-pair shuffles
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log); tail -5 gpurun_out/pytest_gpu.log
for s in 0 1; do
(AZ_WIDE_SHFL16=$s timeout 500 python bench.py --burn-in 20 --steps 6 --no-cpu-baseline --extras resnet4x64:fp16 > gpurun_out/bench_wide_s$s.json 2> gpurun_out/bench_wide_s$s.err; echo "bench rc=$?"); tail -2 gpurun_out/bench_wide_s$s.err
python -c "
import json; d=json.load(open('gpurun_out/bench_wide_s$s.json'))
print(d['value'], d['e2e']['value'], d['roofline']['kernel'], d['roofline']['kernel_ms'], d['roofline']['frac'])
print(d['details'].get('evaluator_max_abs_dev_vs_fp32_predict'))
for x in d['net_in_loop']: print(x['evaluator'], x['workload'], '%.3e'%x['sims_per_s'], round(x['evaluator_kernel_us'],1), round(x['tensor_frac_of_measured_bf16'],3))
print(d.get('errors'))"
done

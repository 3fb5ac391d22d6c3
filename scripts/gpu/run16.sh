#!/bin/bash
# GPU session: CTA-pair instance of the 128-channel kernel
mkdir -p gpurun_out
(timeout 240 python -m pytest tests/test_gpu_resnet_pipe.py -x -q -k "1280" > gpurun_out/pytest_pair128.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_pair128.log); tail -6 gpurun_out/pytest_pair128.log
nvidia-smi --query-gpu=utilization.gpu,memory.used --format=csv,noheader
(timeout 400 python bench.py --burn-in 4 --steps 2 --no-cpu-baseline --no-e2e --extras resnet9x128,resnet9x128:bf16:pair > gpurun_out/bench_p128.json 2> gpurun_out/bench_p128.err; echo "bench rc=$?"); tail -2 gpurun_out/bench_p128.err
python -c "
import json; d=json.load(open('gpurun_out/bench_p128.json'))
for x in d['net_in_loop']: print(x['evaluator'], x['workload'], '%.3e'%x['sims_per_s'], round(x['evaluator_kernel_us'],1), round(x['tensor_frac_of_measured_bf16'],3))
print(d.get('errors'))"

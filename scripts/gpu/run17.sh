#!/bin/bash
# GPU session: the 64-channel kernel with fused filter rows (N = 192, csrc/az_resnet_wide.cu) - parity, then in the loop
mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_resnet_pipe.py -x -q -k "192" > gpurun_out/pytest_wide.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_wide.log); tail -12 gpurun_out/pytest_wide.log
nvidia-smi --query-gpu=utilization.gpu,memory.used --format=csv,noheader
(timeout 400 python bench.py --burn-in 4 --steps 2 --no-cpu-baseline --no-e2e --extras resnet4x64:bf16:pipe2,resnet4x64:bf16:wide,resnet4x64:fp16:wide > gpurun_out/bench_wide.json 2> gpurun_out/bench_wide.err; echo "bench rc=$?"); tail -2 gpurun_out/bench_wide.err
python -c "
import json; d=json.load(open('gpurun_out/bench_wide.json'))
for x in d['net_in_loop']: print(x['evaluator'], x['workload'], '%.3e'%x['sims_per_s'], round(x['evaluator_kernel_us'],1), round(x['tensor_frac_of_measured_bf16'],3))
print(d.get('errors'))"

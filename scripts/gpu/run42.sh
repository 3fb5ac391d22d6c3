#!/bin/bash
# GPU session: FADD2 / F2FP.RELU epilogues in k_resnet_pipe and k_cnn_conv - whole GPU suite, then those kernels in the loop
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "rc=$?" >> gpurun_out/pytest.log); tail -3 gpurun_out/pytest.log
(timeout 400 python bench.py --burn-in 4 --steps 2 --no-cpu-baseline --no-e2e --extras resnet9x128,cnn,resnet4x64:bf16:pipe2 > gpurun_out/bench_s.json 2> gpurun_out/bench_s.err; echo "bench rc=$?"); tail -1 gpurun_out/bench_s.err | cut -c1-200
python -c "
import json; d=json.load(open('gpurun_out/bench_s.json'))
for x in d['net_in_loop']: print(x['evaluator'], x['workload'], '%.3e'%x['sims_per_s'], round(x['evaluator_kernel_us'],1), round(x['us_per_sim_step'],1), round(x['tensor_frac_of_measured_bf16'],3))
print(d.get('errors'))"

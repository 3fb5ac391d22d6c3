#!/bin/bash
# GPU session: two-CTAs-per-SM instance of the pipe kernel (variant 2) - tests, in-loop comparison, ncu
mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_resnet_pipe.py tests/test_gpu_api.py -x -q > gpurun_out/pytest_a.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_a.log); tail -6 gpurun_out/pytest_a.log
nvidia-smi --query-gpu=utilization.gpu,memory.used --format=csv,noheader
(timeout 400 python bench.py --burn-in 30 --steps 6 --no-cpu-baseline --no-e2e --extras resnet4x64:bf16:pipe2,resnet4x64:bf16:pipe > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"); tail -3 gpurun_out/bench.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_resnet_pipe --launch-skip 4000 -c 1 -o gpurun_out/r02_pipe64x2 -f python bench.py --burn-in 4 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --extras none --trunk-variant 2 > gpurun_out/ncu_p2.log 2>&1; tail -2 gpurun_out/ncu_p2.log | cut -c1-200

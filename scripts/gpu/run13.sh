#!/bin/bash
# GPU session: CTA-pair instance of the pipe kernel (variant 3)
mkdir -p gpurun_out
(timeout 240 python -m pytest tests/test_gpu_resnet_pipe.py -x -q -k "640" > gpurun_out/pytest_pair.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_pair.log); tail -12 gpurun_out/pytest_pair.log
nvidia-smi --query-gpu=utilization.gpu,memory.used --format=csv,noheader
(timeout 300 python bench.py --burn-in 30 --steps 5 --no-cpu-baseline --no-e2e --extras none --trunk-variant 3 > gpurun_out/bench_pair.json 2> gpurun_out/bench_pair.err; echo "bench rc=$?"); tail -2 gpurun_out/bench_pair.err
python -c "
import json; d=json.load(open('gpurun_out/bench_pair.json')); print('pair', d['value'], d['roofline']['kernel_ms'], d['roofline']['frac'])"

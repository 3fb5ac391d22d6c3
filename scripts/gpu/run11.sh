#!/bin/bash
# GPU session: CNN conv kernel without batch barriers
mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_cnn.py -x -q > gpurun_out/pytest_cnn.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_cnn.log); tail -5 gpurun_out/pytest_cnn.log
nvidia-smi --query-gpu=utilization.gpu,memory.used --format=csv,noheader
(timeout 120 python scripts/profile_net_step.py 16384 cnn > gpurun_out/steps.log 2>&1); tail -2 gpurun_out/steps.log
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_cnn --launch-skip 20 -c 2 -o gpurun_out/r02_cnn_b -f python scripts/profile_net_step.py 16384 cnn > gpurun_out/ncu_cnn.log 2>&1; tail -1 gpurun_out/ncu_cnn.log

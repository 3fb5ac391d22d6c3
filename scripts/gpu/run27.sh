#!/bin/bash
# GPU session: k_resnet_wide - parity, then time per 16384 positions
mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_resnet_pipe.py -x -q -k "192" > gpurun_out/pytest_wide.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_wide.log); tail -5 gpurun_out/pytest_wide.log
timeout 120 python scripts/profile_net_step.py 16384 resnet0x64:v4 resnet1x64:v4 resnet4x64:v4 resnet8x64:v4 2>&1 | grep "E="
AZ_WIDE_SHFL16=1 timeout 60 python scripts/profile_net_step.py 16384 resnet4x64:v4 2>&1 | grep "E="

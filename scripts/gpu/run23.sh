#!/bin/bash
# GPU session: fp16-packed neighbour shuffles in k_resnet_wide (AZ_WIDE_SHFL16=1) - parity and time
mkdir -p gpurun_out
(AZ_WIDE_SHFL16=1 timeout 300 python -m pytest tests/test_gpu_resnet_pipe.py -x -q -k "192" > gpurun_out/pytest_wide.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_wide.log); tail -4 gpurun_out/pytest_wide.log
AZ_WIDE_SHFL16=1 python scripts/profile_net_step.py 16384 resnet4x64:v4 resnet8x64:v4 2>&1 | grep "E="
AZ_WIDE_SHFL16=0 python scripts/profile_net_step.py 16384 resnet4x64:v4 2>&1 | grep "E="

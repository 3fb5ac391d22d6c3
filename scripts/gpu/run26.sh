#!/bin/bash
# GPU session: 8 vs 16 epilogue warps in k_resnet_wide
mkdir -p gpurun_out
for ew in 16 8; do
echo "EW $ew"
(AZ_ENGINE_LIB=$PWD/_ab/libaz_ew$ew.so timeout 300 python -m pytest tests/test_gpu_resnet_pipe.py -x -q -k "192" > gpurun_out/pytest_wide.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_wide.log); tail -3 gpurun_out/pytest_wide.log
AZ_ENGINE_LIB=$PWD/_ab/libaz_ew$ew.so python scripts/profile_net_step.py 16384 resnet0x64:v4 resnet4x64:v4 resnet8x64:v4 2>&1 | grep "E="
done

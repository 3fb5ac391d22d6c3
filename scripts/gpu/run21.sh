#!/bin/bash
# GPU session: ncu of k_resnet_wide (eager step of 16384 leaves)
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_resnet_wide --launch-skip 10 -c 1 -o gpurun_out/r02_wide_c -f python scripts/profile_net_step.py 16384 resnet4x64:v4 > gpurun_out/ncu_wide.log 2>&1; tail -2 gpurun_out/ncu_wide.log

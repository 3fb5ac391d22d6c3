#!/bin/bash
# GPU session: k_resnet_wide after the epilogue rewrite (FADD2, F2FP.RELU, biases in registers)
mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_resnet_pipe.py -x -q -k "192" > gpurun_out/pytest_wide.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_wide.log); tail -4 gpurun_out/pytest_wide.log
python scripts/profile_net_step.py 16384 resnet4x64:v4 resnet4x64:v2 2>&1 | tail -2

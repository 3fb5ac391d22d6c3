#!/bin/bash
# GPU session: the two rules kernels under ncu --set full (pipe utilisation: which pipe bounds them)
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_env_step_h|k_state_info_h' --launch-skip 26 -c 2 -o gpurun_out/r02_rules -f python scripts/bench_kernels.py rules > gpurun_out/ncu_rules.log 2>&1; tail -1 gpurun_out/ncu_rules.log | cut -c1-200
ls -la gpurun_out/r02_rules.ncu-rep

#!/bin/bash
# GPU session: ncu --set full of k_expand_select<COMPACT> in the steady state of the headline loop; the step kernel of the rules under ncu
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_expand_select --launch-skip 5000 -c 1 -o gpurun_out/r02_expand_select -f python bench.py --burn-in 6 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --extras none > gpurun_out/ncu_es.log 2>&1; tail -1 gpurun_out/ncu_es.log | cut -c1-200
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'k_env_step_h' --launch-skip 20 -c 5 --csv --log-file gpurun_out/rules_ncu_env.csv python scripts/bench_kernels.py rules > gpurun_out/rules_ncu.log 2>&1
ls -la gpurun_out/r02_expand_select.ncu-rep gpurun_out/rules_ncu_env.csv

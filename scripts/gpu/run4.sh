#!/bin/bash
# GPU session: 128-channel kernel tests, headline bench with per-step times, steady-state launch list, ncu of k_resnet128
mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_trunk128.py tests/test_gpu_api.py tests/test_gpu_player.py -x -q > gpurun_out/pytest.log 2>&1; echo "rc=$?" >> gpurun_out/pytest.log); tail -4 gpurun_out/pytest.log
(timeout 500 python bench.py --no-cpu-baseline --extras resnet9x128 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"); tail -3 gpurun_out/bench.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 14000 -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --burn-in 4 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --extras none > gpurun_out/ncu_list.log 2>&1; tail -2 gpurun_out/ncu_list.log | head -c 300
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_resnet128 --launch-skip 10 -c 1 -o gpurun_out/r02_resnet128b -f python scripts/profile_net_step.py 16384 resnet9x128 > gpurun_out/ncu128.log 2>&1; tail -2 gpurun_out/ncu128.log

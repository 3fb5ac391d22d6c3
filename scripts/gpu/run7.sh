#!/bin/bash
# GPU session: full default bench (all extras), the reference arm, GPU test-suite, ncu full capture of the headline kernel (traffic)
mkdir -p gpurun_out
(timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"); tail -3 gpurun_out/bench.err
(timeout 600 python bench.py --impl reference --steps 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"); tail -3 gpurun_out/bench_ref.err
(timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "rc=$?" >> gpurun_out/pytest.log); tail -4 gpurun_out/pytest.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_resnet_trunk --launch-skip 4000 -c 1 -o gpurun_out/r02_trunk_compact -f python bench.py --burn-in 4 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --extras none > gpurun_out/ncu_trunk.log 2>&1; tail -2 gpurun_out/ncu_trunk.log | cut -c1-200
lscpu | head -20 > gpurun_out/lscpu.txt

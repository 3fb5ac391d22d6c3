#!/bin/bash
# GPU session: what the driver runs at round end - smoke(), the GPU test-suite, the default bench line and the reference arm; then the
# steady-state launch list and an ncu capture of k_resnet_wide
mkdir -p gpurun_out
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"); tail -2 gpurun_out/smoke.log
(timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "rc=$?" >> gpurun_out/pytest.log); tail -4 gpurun_out/pytest.log
(timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"); tail -3 gpurun_out/bench.err; wc -l gpurun_out/bench.json
(timeout 600 python bench.py --impl reference --steps 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?")
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 14000 -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --burn-in 4 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --extras none > gpurun_out/ncu_list.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_resnet_wide --launch-skip 4000 -c 1 -o gpurun_out/r02_wide_steady -f python bench.py --burn-in 14 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --extras none > gpurun_out/ncu_w.log 2>&1; tail -1 gpurun_out/ncu_w.log | cut -c1-200

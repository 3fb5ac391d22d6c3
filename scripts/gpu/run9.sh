#!/bin/bash
# GPU session: stagger sweep of the two-CTAs-per-SM kernel
mkdir -p gpurun_out
for ns in 0 600 1100 1600 2200; do
  AZ_PIPE_STAGGER_NS=$ns timeout 200 python bench.py --burn-in 30 --steps 5 --no-cpu-baseline --no-e2e --extras none > gpurun_out/bench_st$ns.json 2> gpurun_out/bench_st$ns.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_st$ns.json')); print($ns, d['value'], d['roofline']['kernel_ms'], d['roofline']['frac'])"
done
(timeout 200 python -m pytest tests/test_gpu_resnet_pipe.py tests/test_gpu_config3.py -x -q > gpurun_out/pytest_a.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_a.log); tail -3 gpurun_out/pytest_a.log

#!/bin/bash
# GPU session: simulations per replayed graph (AZ_GRAPH_UNROLL) in the headline loop
mkdir -p gpurun_out
for u in 1 8 32; do
(AZ_GRAPH_UNROLL=$u timeout 300 python bench.py --burn-in 16 --steps 6 --no-cpu-baseline --no-e2e --extras none > gpurun_out/bench_u$u.json 2> gpurun_out/bench_u$u.err; echo "bench rc=$?"); tail -1 gpurun_out/bench_u$u.err | cut -c1-200
python -c "
import json; d=json.load(open('gpurun_out/bench_u$u.json'))
print('unroll $u', d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['tree_kernel']['us_per_launch'], d['clocks']['sm_mhz'])"
done
(timeout 200 python -m pytest tests/test_gpu_resnet_pipe.py -x -q -k "deterministic" 2>&1 | tail -3)

#!/bin/bash
# GPU session: A/B on one box - FC warp only (libaz_fcold.so) against last batch shared by the epilogue warps (shipped)
for rep in 1 2; do
for lib in _ab/libaz_fcold.so alphazero-implementation_b200/libaz_engine.so; do
AZ_ENGINE_LIB=$PWD/$lib timeout 300 python bench.py --burn-in 12 --steps 5 --no-cpu-baseline --no-e2e --extras none > gpurun_out/bench_ab.json 2> gpurun_out/bench_ab.err
python -c "
import json; d=json.load(open('gpurun_out/bench_ab.json')); r=d['roofline']
print('$lib', round(d['value']/1e6,3), round(d['ms_per_step'],1), round(r['kernel_ms']*1e3,1), round(r['kernel_ms_events_ungraphed_step']*1e3,1), d['clocks']['sm_mhz'])"
done; done

#!/bin/bash
# GPU session: the default bench line and the reference arm (profiles/r02_bench_1gpu.json, r02_bench_reference_arm.json)
mkdir -p gpurun_out
(timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"); tail -2 gpurun_out/bench.err
(timeout 600 python bench.py --impl reference --steps 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?")
python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['clocks'], d['details']['step_ms'])"

#!/bin/bash
# GPU session: rules kernels on 32-bit halves - parity tests with full output, then GB/s
mkdir -p gpurun_out
for m in 3 4; do
(AZ_RULES_MODE=$m timeout 200 python -m pytest tests/test_gpu_rules.py -q > gpurun_out/pytest_rules_$m.log 2>&1; tail -1 gpurun_out/pytest_rules_$m.log)
AZ_RULES_MODE=$m timeout 200 python scripts/bench_kernels.py rules > gpurun_out/rules_mode_$m.json 2> gpurun_out/rules_mode_$m.err
python -c "
import json; d=json.load(open('gpurun_out/rules_mode_$m.json'))
print('mode $m', {k:(round(v['gbs']),round(v['frac'],3)) for k,v in d.items() if isinstance(v,dict)})"
done
grep -E "^(FAILED|E  )" gpurun_out/pytest_rules_3.log gpurun_out/pytest_rules_4.log | head -30

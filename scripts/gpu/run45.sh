#!/bin/bash
# GPU session: rules kernels on 32-bit halves (c4::h32) - parity tests and achieved GB/s per formulation
# (AZ_RULES_MODE -1 = 64-bit kernels, 0 / 1 / 2 = shifts on the ALU pipe / FMA pipe / split)
mkdir -p gpurun_out
for m in 0 3; do
(AZ_RULES_MODE=$m timeout 200 python -m pytest tests/test_gpu_rules.py -x -q 2>&1 | tail -1)
AZ_RULES_MODE=$m timeout 200 python scripts/bench_kernels.py rules > gpurun_out/rules_mode_$m.json 2> gpurun_out/rules_mode_$m.err
python -c "
import json; d=json.load(open('gpurun_out/rules_mode_$m.json'))
print('mode $m', {k:(round(v['gbs']),round(v['frac'],3)) for k,v in d.items() if isinstance(v,dict)})"
done

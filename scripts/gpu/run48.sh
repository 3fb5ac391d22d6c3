#!/bin/bash
# GPU session: what the driver runs at round end with the final code - smoke(), the bench line (the driver's own command) and the
# reference arm; then the steady-state launch list of the headline command and the rules kernels under ncu (duration + DRAM bytes)
mkdir -p gpurun_out
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"); tail -2 gpurun_out/smoke.log
(timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"); tail -3 gpurun_out/bench.err; wc -l gpurun_out/bench.json
(timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?")
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 14000 -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --burn-in 4 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --extras none > gpurun_out/ncu_list.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'k_env_step_h|k_state_info_h' --launch-skip 30 -c 6 --csv --log-file gpurun_out/rules_ncu.csv python scripts/bench_kernels.py rules > gpurun_out/rules_ncu.log 2>&1
python - <<PY
import json
d=json.load(open('gpurun_out/bench.json'))
print('value %.4e e2e %.4e ms/step %.2f kernel_ms %.4f frac %.4f launches %d' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['gpu_launches']))
print(d.get('errors'))
PY

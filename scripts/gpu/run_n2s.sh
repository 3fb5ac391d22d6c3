#!/bin/bash
# 2-GPU session, short: configs[3] path of bench.py with the final code (games sharded, episode all-gather inside the timed region)
mkdir -p gpurun_out
(timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 8 --warmup 3 --burn-in 12 --no-cpu-baseline --extras none > gpurun_out/bench_n2s.json 2> gpurun_out/bench_n2s.err; echo "bench rc=$?")
grep -i "nranks\|error\|Traceback" gpurun_out/bench_n2s.err | head -8
python -c "
import json; d=json.load(open('gpurun_out/bench_n2s.json'))
print('value %.4e e2e %.4e ms/step %.2f kernel_ms %.4f n_gpus %d workload %s' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['n_gpus'], d['config']['workload']))
print(d.get('episode_allgather'), d.get('errors'))"

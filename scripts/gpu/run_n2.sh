#!/bin/bash
# 2-GPU session: configs[3] path of bench.py (games sharded, episode all-gather inside the timed region), short
mkdir -p gpurun_out
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench rc=$?")
grep -i "nranks\|error\|Traceback" gpurun_out/bench_n2.err | head -8; tail -c 1500 gpurun_out/bench_n2.json

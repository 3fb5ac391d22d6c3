#!/bin/bash
# 8-GPU session: configs[3] through bench.py (65536 games sharded, all-gather inside the timed region) and configs[4]
# (full iterations with the trainer rank playing a smaller share)
mkdir -p gpurun_out
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; echo "bench8 rc=$?")
grep -c "Init COMPLETE" gpurun_out/bench_n8.err; grep -m2 "nranks" gpurun_out/bench_n8.err | cut -c1-200; head -c 700 gpurun_out/bench_n8.json; echo
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 scripts/run_iteration.py --games 65536 --sims 800 --net resnet4x64 --iterations 2 --batch-size 2048 --trainer-share 0.08 > gpurun_out/config5_n8.json 2> gpurun_out/config5_n8.err; echo "config5 rc=$?")
tail -c 1500 gpurun_out/config5_n8.json; tail -3 gpurun_out/config5_n8.err | cut -c1-300

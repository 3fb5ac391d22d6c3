#!/bin/bash
# 4-GPU session: configs[3] through bench.py at 4 and 2 GPUs (default steps)
mkdir -p gpurun_out
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 4 > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err; echo "bench4 rc=$?")
grep -c "Init COMPLETE" gpurun_out/bench_n4.err; grep -m1 "nranks" gpurun_out/bench_n4.err | cut -c1-220; head -c 400 gpurun_out/bench_n4.json; echo
(CUDA_VISIBLE_DEVICES=0,1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench2 rc=$?")
head -c 400 gpurun_out/bench_n2.json; echo

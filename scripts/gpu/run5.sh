#!/bin/bash
# GPU session: pipe kernel for 64 channels; full GPU test suite; headline bench; ncu of k_resnet_pipe<64>
mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_resnet_pipe.py tests/test_gpu_trunk.py tests/test_gpu_player.py -x -q > gpurun_out/pytest_a.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_a.log); tail -4 gpurun_out/pytest_a.log
(timeout 700 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_resnet_pipe.py --deselect tests/test_gpu_trunk.py --deselect tests/test_gpu_player.py > gpurun_out/pytest_b.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_b.log); tail -4 gpurun_out/pytest_b.log
(timeout 600 python bench.py --extras resnet4x64:bf16:pingpong,resnet4x64:fp16,resnet9x128 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"); tail -3 gpurun_out/bench.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_resnet_pipe --launch-skip 10 -c 1 -o gpurun_out/r02_pipe64 -f python scripts/profile_net_step.py 16384 resnet4x64 > gpurun_out/ncu64.log 2>&1; tail -2 gpurun_out/ncu64.log

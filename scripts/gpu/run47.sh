#!/bin/bash
# GPU session: whole GPU suite after the rules-kernel rework; lanes per tree A/B of the fused tree kernel
mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "rc=$?" >> gpurun_out/pytest.log); tail -3 gpurun_out/pytest.log
timeout 200 python scripts/ab_lanes.py > gpurun_out/ab_lanes.log 2>&1; grep trees gpurun_out/ab_lanes.log | head -12

#!/bin/bash
# GPU session: the whole GPU suite on the committed tree (what the driver runs at round end)
mkdir -p gpurun_out
(timeout 90 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "rc=$?" >> gpurun_out/pytest.log); tail -3 gpurun_out/pytest.log

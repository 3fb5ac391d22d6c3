#!/bin/bash
mkdir -p gpurun_out
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"); tail -2 gpurun_out/smoke.log
(timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "rc=$?" >> gpurun_out/pytest.log); tail -4 gpurun_out/pytest.log

#!/bin/bash
# GPU session: compute-sanitizer memcheck over every kernel family at tiny sizes; then the full GPU test-suite
mkdir -p gpurun_out
(timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python scripts/sanitize_smoke.py > gpurun_out/sanitize.log 2>&1; echo "sanitize rc=$?"); tail -8 gpurun_out/sanitize.log | cut -c1-250
(timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "rc=$?" >> gpurun_out/pytest.log); tail -4 gpurun_out/pytest.log

#!/bin/bash
# GPU session: the default build once more - rules / compaction / API tests and smoke()
mkdir -p gpurun_out
(timeout 200 python -m pytest tests/test_gpu_rules.py tests/test_gpu_compaction.py tests/test_gpu_api.py -x -q 2>&1 | tail -2)
(timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1)

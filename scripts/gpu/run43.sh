#!/bin/bash
# GPU session: list of the leaves to evaluate built inside k_expand_select (az_set_leaf_compaction 2) - new tests, whole GPU suite,
# then the headline loop with the extra launch (AZ_COMPACT_FUSED=0) against the fused list, A/B/A/B on one box
mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_compaction.py tests/test_gpu_resnet_pipe.py -x -q > gpurun_out/pytest_compact.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_compact.log); tail -15 gpurun_out/pytest_compact.log
(timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "rc=$?" >> gpurun_out/pytest.log); tail -3 gpurun_out/pytest.log
for rep in 0 1; do for f in 0 1; do
(AZ_COMPACT_FUSED=$f timeout 200 python bench.py --burn-in 14 --steps 6 --warmup 3 --no-e2e --no-cpu-baseline --extras none > gpurun_out/bench_f${f}_$rep.json 2> gpurun_out/bench_f${f}_$rep.err; echo "fused=$f rep=$rep rc=$?")
python -c "
import json; d=json.load(open('gpurun_out/bench_f${f}_$rep.json'))
print('%.4e sims/s  %.2f ms/step  us/simstep %.1f  kernel_ms %.4f  launches %d' % (d['value'], d['ms_per_step'], d['details']['us_per_simulation_step'], d['roofline']['kernel_ms'], d['gpu_launches']))"
done; done

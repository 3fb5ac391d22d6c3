#!/bin/bash
# GPU session: fp16-pair shuffles with fp16 operands too (AZ_WIDE_SHFL16=1): parity + deviation from fp32 predict
mkdir -p gpurun_out
(AZ_WIDE_SHFL16=1 timeout 600 python -m pytest tests/test_gpu_resnet_pipe.py tests/test_gpu_config3.py -x -q -k "192 or deterministic or config3 or deviation or tolerance or 1e" > gpurun_out/pytest_s16.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_s16.log); tail -5 gpurun_out/pytest_s16.log
cat gpurun_out/evaluator_deviation.json | head -12
python -c "
import json; d=json.load(open('gpurun_out/policy_target_deviation.json')); print({k:(v['policy_target_mean_abs_dev'], v['trees_with_identical_visit_counts']) for k,v in d.items()})"

#!/bin/bash
# submit a GPU session script through gpurun, retrying while the pod answers "busy" (nothing is charged for those)
# usage: scripts/gpu/submit.sh <timeout_s> <script> [gpus]
T=$1; S=$2; G=${3:-1}
for i in $(seq 1 30); do
  if [ "$G" = "1" ]; then /usr/local/graft/bin/gpurun --timeout $T -- "bash $S" > /tmp/gpurun_last.log 2>&1; else /usr/local/graft/bin/gpurun --gpus $G --timeout $T -- "bash $S" > /tmp/gpurun_last.log 2>&1; fi
  rc=$?
  if grep -q "status=transient" /tmp/gpurun_last.log || [ $rc -eq 3 ]; then sleep 90; continue; fi
  break
done
tail -40 /tmp/gpurun_last.log

#!/bin/bash
# usage: tools/gpu_job_n.sh <gpus> <timeout-seconds> '<command>'  -- multi-GPU gpurun with retries while busy (exit 3)
n=$1; t=$2; shift; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --gpus "$n" --timeout "$t" -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3

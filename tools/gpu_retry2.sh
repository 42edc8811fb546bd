#!/bin/bash
# tools/gpu_retry2.sh <n-gpus> <timeout-seconds> <command...>: multi-GPU gpurun with retries while the pod answers "busy" (exit 3)
N=$1; T=$2; shift; shift
for i in $(seq 1 15); do
  /usr/local/graft/bin/gpurun --gpus "$N" --timeout "$T" -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 120
done
exit 3

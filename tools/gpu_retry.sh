#!/bin/bash
# usage: tools/gpu_retry.sh <timeout-seconds> <gpus> '<command>'   -- retries while gpurun answers "busy" (exit 3)
T=$1; G=$2; shift 2
for i in $(seq 1 30); do
  if [ "$G" = "1" ]; then gpurun --timeout $T -- "$@"; else gpurun --gpus $G --timeout $T -- "$@"; fi
  rc=$?
  [ $rc -ne 3 ] && exit $rc
  sleep 90
done
exit 3

#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit code 3, nothing charged).
# usage: [GPUS=N] gpu_retry.sh TIMEOUT 'command'
t=$1; shift
for i in $(seq 1 30); do
  if [ -n "$GPUS" ]; then /usr/local/graft/bin/gpurun --gpus "$GPUS" --timeout "$t" -- "$@"; else /usr/local/graft/bin/gpurun --timeout "$t" -- "$@"; fi
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3

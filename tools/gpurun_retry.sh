#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout-seconds> '<command>'   — retries while the pod has no free GPU slot (exit 3)
t=$1; shift
for k in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout "$t" -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 150
done
exit 3

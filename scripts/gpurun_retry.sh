#!/bin/bash
# usage: scripts/gpurun_retry.sh <timeout-seconds> [--gpus N] -- '<command>'   (retries while the pod answers "busy", exit code 3)
t=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$t" "$@" > gpurun_out/.retry.log 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" gpurun_out/.retry.log; then cat gpurun_out/.retry.log; exit $rc; fi
  sleep 120
done
cat gpurun_out/.retry.log; exit 3

#!/bin/bash
# gpurun with retries while the pod has no free slot (exit code 3 = nothing charged).  usage: gpurun_retry.sh <timeout> '<command>'
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout "$1" -- "$2"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 120
done
exit 3

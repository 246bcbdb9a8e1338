#!/bin/bash
# usage: scripts/gpurun_retry.sh <log> <timeout> <command...>   (retries while the pod answers "busy": exit code 3 / transient)
log=$1; shift; to=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $to -- "$@" > $log 2>&1
  if grep -q "status=transient" $log; then sleep 90; continue; fi
  break
done
tail -3 $log

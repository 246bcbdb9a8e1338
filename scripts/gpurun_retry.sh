#!/bin/bash
# usage: scripts/gpurun_retry.sh <log> <timeout> [--gpus N] <command>   (retries while the pod answers "busy")
log=$1; shift; to=$1; shift
extra=""
if [ "$1" = "--gpus" ]; then extra="--gpus $2"; shift; shift; fi
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $to $extra -- "$@" > $log 2>&1
  if grep -q "status=transient" $log; then sleep 90; continue; fi
  break
done
tail -3 $log

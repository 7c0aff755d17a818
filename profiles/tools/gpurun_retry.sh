#!/bin/bash
# usage: gpurun_retry.sh <logfile> <timeout> <command...>   retries while the pod answers busy (rc 3)
log=$1; shift; to=$1; shift
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout $to -- "$@" > $log 2>&1
  if grep -q "status=transient\|status=busy" $log; then sleep 60; continue; fi
  break
done
tail -30 $log

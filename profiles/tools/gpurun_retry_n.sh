#!/bin/bash
# usage: gpurun_retry_n.sh <n_gpus> <logfile> <timeout> <command...>   retries while the pod answers busy
n=$1; shift; log=$1; shift; to=$1; shift
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --gpus $n --timeout $to -- "$@" > $log 2>&1
  if grep -q "status=transient\|status=busy" $log; then sleep 90; continue; fi
  break
done
tail -30 $log

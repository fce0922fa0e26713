#!/bin/bash
# usage: scripts/gpurun_retry.sh LOG TIMEOUT [--gpus N] -- command ...   (retries while the pod answers busy, rc 3)
log=$1; to=$2; shift 2
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout $to "$@" > $log 2>&1
  rc=$?
  [ $rc -ne 3 ] && exit $rc
  sleep 150
done
exit 3

#!/bin/bash
# usage: gpurun_retry.sh <log> <gpurun args...>: repeats the call while the pod answers "no slot" (exit 3), at most 20 times
log=$1; shift
for try in $(seq 1 20); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 75
done
exit 3

#!/bin/bash
# usage: scripts/gpurun_retry.sh <log> <timeout-seconds> <command...>   retries while the pod answers "busy" (exit 3)
log=$1; shift; tmo=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$tmo" -- "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 120
done
exit 3

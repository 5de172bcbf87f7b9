#!/bin/bash
# gpurun with retries while the pod answers "transient" (exit code 3): scripts/gpurun_retry.sh <log> <timeout> <command...>
log=$1; shift; to=$1; shift
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout $to -- "$@" > $log 2>&1; rc=$?
  if ! grep -q "status=transient" $log; then exit $rc; fi
  sleep 120
done
exit 3

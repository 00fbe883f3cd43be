#!/bin/bash
# usage: gpurun_retry.sh <timeout_s> <out_file> <command...>
# retries while the pod answers "busy" (exit 3) or another call is still registered as running (exit 2)
T=$1; OUT=$2; shift 2
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$T" -- "$@" > "$OUT" 2>&1
  rc=$?
  if [ $rc -ne 3 ] && [ $rc -ne 2 ]; then exit $rc; fi
  sleep 90
done
exit 3

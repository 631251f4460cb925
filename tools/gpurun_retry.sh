#!/bin/bash
# usage: tools/gpurun_retry.sh <logfile> <gpurun args...>   — retries while the pod answers busy/transient (exit 2/3)
log=$1; shift
for attempt in 1 2 3 4 5 6 7 8; do
  timeout 3300 gpurun "$@" > "$log" 2>&1; rc=$?
  echo "exit $rc (attempt $attempt)" >> "$log"
  if [ $rc -ne 3 ] && [ $rc -ne 2 ]; then break; fi
  sleep 90
done

#!/bin/bash
# usage: gpurun_retry.sh <timeout> <script> <logfile>  -- retries while the pod answers "busy" (exit 3), nothing is charged for those
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout "$1" -- "bash $2" > "$3" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3

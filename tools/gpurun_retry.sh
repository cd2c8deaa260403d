#!/bin/bash
# local helper: retry a gpurun call while the pod answers "busy" (exit code 3)
# usage: tools/gpurun_retry.sh <logfile> <gpurun args...>
LOG=$1; shift
for i in $(seq 1 20); do
  gpurun "$@" > "$LOG" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then echo "rc=$rc" >> "$LOG"; echo done >> "$LOG"; exit $rc; fi
  sleep 90
done
echo "gave up" >> "$LOG"; echo done >> "$LOG"

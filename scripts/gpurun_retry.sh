#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit 3: nothing charged).  usage: gpurun_retry.sh <log> <timeout> [--gpus N] -- <command>
LOG=$1; TO=$2; shift 2
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout "$TO" "$@" > "$LOG" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then echo "gpurun rc=$rc" >> "$LOG"; exit $rc; fi
  sleep 90
done
echo "gpurun: still busy after 20 tries" >> "$LOG"; exit 3

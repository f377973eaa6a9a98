#!/bin/bash
# usage: scripts/gpu_retry.sh <log> <timeout_s> <gpus> <command...>   — retries while the pod answers "busy" (exit 3)
LOG=$1; TO=$2; GP=$3; shift 3
for i in $(seq 1 40); do
  if [ "$GP" = "1" ]; then /usr/local/graft/bin/gpurun --timeout $TO -- "$@" > $LOG 2>&1; else /usr/local/graft/bin/gpurun --gpus $GP --timeout $TO -- "$@" > $LOG 2>&1; fi
  rc=$?
  if [ $rc -ne 3 ]; then echo "attempt $i rc=$rc" >> $LOG; exit $rc; fi
  sleep 90
done
exit 3

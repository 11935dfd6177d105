#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit code 3: nothing charged).
# usage: gpurun_retry.sh <timeout> <out file> [--gpus N] <command>
T=$1; OUT=$2; shift 2
EXTRA=()
if [ "$1" == "--gpus" ]; then EXTRA=(--gpus "$2"); shift 2; fi
for i in $(seq 1 60); do
  /usr/local/graft/bin/gpurun --timeout $T "${EXTRA[@]}" -- "$@" > $OUT 2>&1
  rc=$?
  [ $rc -ne 3 ] && exit $rc
  sleep 75
done
exit 3

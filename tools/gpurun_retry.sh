#!/bin/bash
# usage: tools/gpurun_retry.sh <log> <gpus> <timeout_s> <command...>   -- retries while the pod answers "busy" (exit 3)
log=$1; gpus=$2; to=$3; shift 3
for attempt in $(seq 1 12); do
  /usr/local/graft/bin/gpurun $( [ "$gpus" != 1 ] && echo --gpus $gpus ) --timeout $to -- "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 60
done
exit 3

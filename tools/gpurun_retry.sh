#!/bin/bash
# usage: tools/gpurun_retry.sh LOGFILE GPURUN_ARGS... ; retries while the pod answers "busy" (exit 3)
log=$1; shift
for k in $(seq 1 30); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 60
done
exit 3

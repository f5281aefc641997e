#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit code 3): usage gpurun_retry.sh OUTFILE [gpurun args...] -- command
out=$1; shift
for try in $(seq 1 12); do
  /usr/local/graft/bin/gpurun "$@" > "$out" 2>&1
  rc=$?
  [ $rc -ne 3 ] && exit $rc
  sleep 200
done
exit 3

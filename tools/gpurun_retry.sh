#!/bin/bash
# usage: tools/gpurun_retry.sh [gpurun args...] -- 'command'   (retries while the pod answers "busy", rc 3)
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@"
  rc=$?
  [ $rc -ne 3 ] && exit $rc
  sleep 150
done
exit 3

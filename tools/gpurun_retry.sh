#!/bin/bash
# usage: tools/gpurun_retry.sh <tag> [gpurun args...] -- retries while the pod answers "busy" (exit 3)
tag=$1; shift
for i in 1 2 3 4 5 6 7 8 9 10; do
  /usr/local/graft/bin/gpurun "$@" > gpurun_out/${tag}_call.log 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3

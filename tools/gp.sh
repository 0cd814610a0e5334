#!/bin/bash
# gp.sh <tag> <timeout> <command...>: gpurun with retries while the pod is busy; log in gpurun_out/<tag>.call
tag=$1; shift; to=$1; shift
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout $to -- "$@" > gpurun_out/$tag.call 2>&1
  rc=$?
  if grep -q "status=transient" gpurun_out/$tag.call || [ $rc -eq 3 ]; then sleep 90; continue; fi
  break
done
echo "done rc=$rc" >> gpurun_out/$tag.call

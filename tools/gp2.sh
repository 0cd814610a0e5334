#!/bin/bash
# gp2.sh <tag> <gpus> <timeout> <command...>: multi-GPU gpurun with retries while the pod is busy
tag=$1; shift; n=$1; shift; to=$1; shift
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --gpus $n --timeout $to -- "$@" > gpurun_out/$tag.call 2>&1
  rc=$?
  if grep -q "status=transient" gpurun_out/$tag.call || [ $rc -eq 3 ]; then sleep 120; continue; fi
  break
done
echo "done rc=$rc" >> gpurun_out/$tag.call

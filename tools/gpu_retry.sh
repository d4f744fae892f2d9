#!/bin/bash
# usage: tools/gpu_retry.sh <timeout_s> [--gpus N] -- '<command>'   (retries while the pod answers busy/transient)
T=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$T" "$@" > /tmp/gpurun_last.txt 2>&1
  rc=$?
  if grep -q "status=transient\|nothing was charged" /tmp/gpurun_last.txt || [ $rc -eq 3 ]; then
    sleep 90; continue
  fi
  break
done
tail -60 /tmp/gpurun_last.txt
exit $rc

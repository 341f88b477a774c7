#!/bin/bash
# usage: scratch/gpu_retry.sh TIMEOUT 'command'   -- retries while the pod answers "busy" (rc 3 / transient)
T=$1; shift
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun --timeout $T -- "$@" 2>&1)
  if echo "$out" | grep -q "status=transient"; then sleep 90; continue; fi
  echo "$out" | tail -120
  exit 0
done
echo "gave up: pod busy"

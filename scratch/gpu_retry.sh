#!/bin/bash
# usage: scratch/gpu_retry.sh TIMEOUT [gpurun options, e.g. --gpus 2] -- 'command'
# retries while the pod answers "busy" (transient status, nothing charged)
T=$1; shift
OPTS=()
while [ "$1" != "--" ] && [ $# -gt 1 ]; do OPTS+=("$1"); shift; done
[ "$1" == "--" ] && shift
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun --timeout $T "${OPTS[@]}" -- "$@" 2>&1)
  if echo "$out" | grep -q "status=transient"; then sleep 90; continue; fi
  echo "$out" | tail -150
  exit 0
done
echo "gave up: pod busy"

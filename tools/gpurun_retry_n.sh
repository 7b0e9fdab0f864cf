#!/bin/bash
# usage: tools/gpurun_retry_n.sh <gpus> <timeout_s> '<command>'
g=$1; t=$2; shift 2
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun --gpus "$g" --timeout "$t" -- "$@" 2>&1)
  if echo "$out" | grep -q "status=transient\|exit code 3\|no box\|busy"; then sleep 60; continue; fi
  echo "$out"; exit 0
done
echo "$out"; echo "gave up"

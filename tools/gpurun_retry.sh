#!/bin/bash
# gpurun with retries while the pod answers "transient" (no slot) -- nothing is charged for those.
# usage: tools/gpurun_retry.sh <timeout_s> '<command>'
t=$1; shift
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$t" -- "$@" 2>&1)
  if echo "$out" | grep -q "status=transient\|exit code 3\|no box"; then sleep 45; continue; fi
  echo "$out"; exit 0
done
echo "$out"; echo "gave up after 40 transient answers"

#!/bin/bash
# gpurun with retries while the pod answers "busy / transient" (nothing is charged for those): scripts/gpu.sh <timeout_s> '<command>'
T=$1; shift
for i in $(seq 1 30); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$T" -- "$@" 2>&1); rc=$?
  if echo "$out" | grep -q "status=transient\|no box or slot\|retry in a few minutes" || [ $rc -eq 3 ]; then sleep 60; continue; fi
  echo "$out"; exit $rc
done
echo "$out"; exit 3

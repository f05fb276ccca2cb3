#!/bin/bash
# usage: gpurun_retry.sh <timeout_s> '<command>'   -- retries while the pod answers busy (exit 3 / transient)
T=$1; shift
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$T" -- "$@" 2>&1)
  if echo "$out" | grep -q "status=transient\|status=busy"; then sleep 60; continue; fi
  echo "$out"; exit 0
done
echo "$out"; exit 3

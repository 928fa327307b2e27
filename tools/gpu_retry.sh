#!/bin/bash
# usage: tools/gpu_retry.sh <timeout> '<command>'  -- retries while the pod answers busy (rc 3 / transient)
T=$1; shift
for i in $(seq 1 30); do
  out=$(/usr/local/graft/bin/gpurun --timeout $T -- "$@" 2>&1)
  echo "$out" | tail -40
  if echo "$out" | grep -q "status=transient\|rc=3\|no box\|busy"; then sleep 60; continue; fi
  break
done

#!/bin/bash
# tools/gr.sh TIMEOUT 'command' : gpurun with retries while the pod answers busy (exit 3 / transient); nothing is charged for those
t=$1; shift
for i in 1 2 3 4 5 6 7 8 9 10 11 12; do
  out=$(/usr/local/graft/bin/gpurun --timeout "$t" -- "$@" 2>&1)
  if echo "$out" | grep -q "status=transient\|no box\|retry in a few minutes"; then sleep 60; continue; fi
  echo "$out"; exit 0
done
echo "$out"; exit 3

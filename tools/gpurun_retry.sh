#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout> '<command>'  -- retries while the pod answers busy (rc 3 / transient)
for i in $(seq 1 20); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$1" -- "$2" 2>&1); rc=$?
  if echo "$out" | grep -q "status=transient"; then sleep 90; continue; fi
  echo "$out"; exit $rc
done
echo "gave up: pod busy"; exit 3

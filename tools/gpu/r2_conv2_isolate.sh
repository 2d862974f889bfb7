#!/bin/bash
# where does the time of the CTA-pair conv go: timing switches d1 (no producer work), d2 (no epilogue), d4 (no weight reloads)
cd "$(dirname "$0")/../.."
run() { echo "== $*"; timeout 90 python tools/conv2_check.py "$@" 2>&1 | grep -v Warning | grep -v "^  ran\|reference\|round-1\|statistics"; }
timeout 120 tools/micro/mma_rate > gpurun_out/mma_rate2.txt 2>&1
for sh in "64 64 96 0 96 3 t g q" "64 32 192 0 192 3 t g q" "64 32 192 0 576 1 t q"; do
  for d in d0 d1 d2 d4 d3 d5 d7; do run $sh $d; done
done

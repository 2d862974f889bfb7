#!/bin/bash
cd "$(dirname "$0")/../.."
run() { echo "== $*"; timeout 90 python tools/conv2_check.py "$@" 2>&1 | grep -v Warning | grep -v "^  ran"; }
for sh in "64 64 96 0 96 3 t q" "64 32 192 0 192 3 t q"; do
  for d in d55 d23 d39 d7 d15 d47 d31; do run $sh $d; done
done

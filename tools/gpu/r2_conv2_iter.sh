#!/bin/bash
cd "$(dirname "$0")/../.."
run() { echo "== $*"; timeout 45 python tools/conv2_check.py "$@" 2>&1 | grep -v Warning | grep -v "^  ran"; rc=${PIPESTATUS[0]}; if [ "$rc" = "124" ]; then echo "HANG -- aborting"; exit 1; fi; }
run 3 8 64 32 96 3 t r g
run 64 64 96 0 96 3 t g q d7
run 64 64 96 0 96 3 t g q d15
run 64 64 96 0 96 3 t g q d0
run 64 64 96 0 96 3 t g q d1
run 64 64 96 0 96 3 t g q d2
run 64 32 192 0 192 3 t g q
run 64 32 192 0 576 1 t q
run 64 64 96 0 96 3 t r s192,96 g q

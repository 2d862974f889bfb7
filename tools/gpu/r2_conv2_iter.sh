#!/bin/bash
cd "$(dirname "$0")/../.."
run() { echo "== $*"; timeout 45 python tools/conv2_check.py "$@" 2>&1 | grep -v Warning | grep -v "^  ran"; rc=${PIPESTATUS[0]}; if [ "$rc" = "124" ]; then echo "HANG -- aborting"; exit 1; fi; }
run 3 8 64 32 96 3 t r g
run 2 16 48 48 144 3 t r g
run 2 8 96 0 96 3 t s96,96 g
run 2 16 64 0 192 1 t g
run 5 4 32 0 32 3 t r
run 64 64 96 0 96 3 t g q d7
run 64 64 96 0 96 3 t g q d0
run 64 64 96 0 96 3 t g q d1
run 64 64 96 0 96 3 t g q d2
run 64 32 192 0 192 3 t g q
run 64 32 192 0 576 1 t q
run 64 64 96 0 96 3 t r s192,96 g q
run 64 16 288 0 288 3 t g q
run 64 8 384 0 384 3 t g q

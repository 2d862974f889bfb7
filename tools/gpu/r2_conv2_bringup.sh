#!/bin/bash
# bring-up of the CTA-pair conv kernel: small exact checks first (abort on the first hang), then cfg2-sized shapes
# against the round-1 kernel
cd "$(dirname "$0")/../.."
run() { echo "== $*"; timeout 45 python tools/conv2_check.py "$@" 2>&1 | grep -v Warning; rc=${PIPESTATUS[0]}; echo "rc=$rc"; if [ "$rc" = "124" ]; then echo "HANG -- aborting"; exit 1; fi; }
run 2 8 32 0 32 3
run 2 16 32 0 64 3 t r
run 3 8 64 32 96 3 t r g
run 2 16 48 48 144 3 t r g
run 2 16 64 0 192 1 t g
run 4 8 128 128 256 3 t r g
run 2 16 32 0 512 3 g
run 2 16 64 0 64 3 t s32 g
run 2 8 96 0 96 3 t s96,96 g
run 3 16 48 0 48 3 t s48,48
run 37 16 32 0 32 3 t r g
run 8 32 96 0 96 3 t r g
run 64 64 96 0 96 3 t g q
run 64 64 96 0 96 3 t g q d7
run 64 64 96 0 96 3 t r s192,96 g q
run 64 32 192 0 192 3 t g q
run 64 32 192 0 192 3 t g q d7
run 64 32 192 0 576 1 t q
run 64 32 192 0 192 1 r g q
run 64 16 288 0 288 3 t g q
run 64 8 384 0 384 3 t g q
run 64 64 288 0 96 3 t g q
run 64 64 192 0 192 3 g q
run 64 64 16 0 96 3 g q
run 64 64 96 0 16 3 t q

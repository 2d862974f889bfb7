set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
for args in "64 32 192 576 1 0 tab" "64 32 192 192 1 0 res" "64 64 96 96 3 0 tab" "64 64 96 96 3 1 tab" "64 32 192 192 3 0 tab" "64 64 288 96 3 0 tab" "64 16 288 864 1 0 tab" "64 64 192 192 3 0 plain" "64 8 384 384 3 0 tab"; do
  timeout 120 python tools/umma_timing.py $args
done

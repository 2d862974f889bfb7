#!/bin/bash
# round-1 kernel with the second MMA warp + epilogue statistics: op tests, per-shape timing, a short cfg2 bench
cd "$(dirname "$0")/../.."
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_conv2.py -x -q -m gpu 2>&1 | tail -8
for args in "64 64 96 96 3 0 tab" "64 32 192 192 3 0 tab" "64 32 192 576 1 0 tab" "64 64 288 96 3 0 tab" "64 16 288 288 3 0 tab" "64 8 384 384 3 0 tab"; do
  timeout 60 python tools/umma_timing.py $args 2>&1 | grep -v Warning
done
timeout 400 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-psnr 2> gpurun_out/bench_v1b.err | tail -1 > gpurun_out/bench_v1b.json
python -c "
import json; d=json.load(open('gpurun_out/bench_v1b.json')); print('frames/s', d['value'], 'e2e', d['e2e']['value']); print(json.dumps(d.get('roofline',{}).get('per_kind',{}), indent=0)[:1500])"

#!/bin/bash
# round-2: autoregressive video_gen lines of the per-GPU shards of cfg4 (64 clips over 4 GPUs) and cfg5 (32 clips over 8)
cd "$(dirname "$0")/../.."
timeout 500 python bench.py --workload cfg4 --ar --batch 16 --steps 1 --warmup 1 --no-cpu-baseline --no-psnr --no-roofline 2> gpurun_out/r2_bench_cfg4_ar_b16.err | tail -1 > gpurun_out/r2_bench_cfg4_ar_b16.json
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_cfg4_ar_b16.json')); print('cfg4 AR B=16:', d['value'], d['unit'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'])"
timeout 700 python bench.py --workload cfg5 --ar --batch 4 --steps 1 --warmup 1 --no-cpu-baseline --no-psnr --no-roofline 2> gpurun_out/r2_bench_cfg5_ar_b4.err | tail -1 > gpurun_out/r2_bench_cfg5_ar_b4.json
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_cfg5_ar_b4.json')); print('cfg5 AR B=4:', d['value'], d['unit'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'])"

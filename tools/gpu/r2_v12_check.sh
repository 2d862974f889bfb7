#!/bin/bash
# round-2: producer tweaks of k_conv_umma (two threads per row on 128-row slabs, centre-only shortcut K-blocks)
cd "$(dirname "$0")/../.."
( time timeout 900 python -m pytest tests -x -q -m gpu ) > gpurun_out/r2_pytest_gpu_v12.log 2>&1
tail -5 gpurun_out/r2_pytest_gpu_v12.log
{
for shp in "64 32 192 576 1" "64 32 192 192 1" "64 64 96 96 3" "64 32 192 192 3" "64 16 288 288 3"; do
  timeout 120 python tools/umma_timing.py $shp
done
} > gpurun_out/r2_v12_timing.txt 2>&1
cat gpurun_out/r2_v12_timing.txt
timeout 600 python bench.py --steps 3 --warmup 3 2> gpurun_out/r2_bench_v12.err | tail -1 > gpurun_out/r2_bench_v12.json
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_v12.json')); print('frames/s', d['value'], 'e2e', d['e2e']['value'], 'psnr', d.get('psnr_vs_oracle_db'), 'cpu', d.get('cpu_baseline',{}).get('value'), d.get('cpu_baseline',{}).get('kind')); print({k:round(v['ms_per_forward'],3) for k,v in d['roofline']['per_kind'].items()}); print(d['roofline'].get('frac'), d['clocks'])"

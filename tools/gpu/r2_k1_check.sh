#!/bin/bash
# round-2: input-stationary 1x1 kernel (conv1x1_umma.cu) + un-fused skip projection
cd "$(dirname "$0")/../.."
timeout 300 python -m pytest tests/test_gpu_ops.py -x -q -k conv1x1 > gpurun_out/r2_k1_unit.log 2>&1
tail -15 gpurun_out/r2_k1_unit.log
{
for shp in "64 32 192 576 1" "64 32 192 192 1 0 res" "64 64 192 96 1 0 plain" "64 16 288 864 1" "64 8 384 1152 1"; do
  timeout 120 python tools/umma_timing.py $shp 2>&1 | head -1
  MCVD_CONV1X1=0 timeout 120 python tools/umma_timing.py $shp 2>&1 | head -1
done
} > gpurun_out/r2_k1_timing.txt 2>&1
cat gpurun_out/r2_k1_timing.txt
( time timeout 900 python -m pytest tests -x -q -m gpu ) > gpurun_out/r2_pytest_gpu_k1.log 2>&1
tail -5 gpurun_out/r2_pytest_gpu_k1.log
for fs in auto 1; do
MCVD_FUSE_SC=$fs timeout 600 python bench.py --steps 3 --warmup 3 2> gpurun_out/r2_bench_k1_$fs.err | tail -1 > gpurun_out/r2_bench_k1_$fs.json
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_k1_$fs.json')); print('fuse=$fs frames/s', d['value'], 'e2e', d['e2e']['value'], 'psnr', d.get('psnr_vs_oracle_db')); print({k:round(v['ms_per_forward'],3) for k,v in d['roofline']['per_kind'].items()}); print(d['roofline'].get('frac'), d['clocks'])"
done

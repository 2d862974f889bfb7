#!/bin/bash
# round-2: cycle attribution of the input-stationary 1x1 kernel; timing switches; bench with / without the fused shortcut
cd "$(dirname "$0")/../.."
timeout 300 python -m pytest tests/test_gpu_ops.py -x -q -k "conv1x1 or test_conv" 2>&1 | tail -3
{
for shp in "64 32 192 576 1 0 tab" "64 64 192 96 1 0 plain" "64 32 192 192 1 0 res" "64 16 288 864 1 0 tab"; do
  for f in 0 1 3; do
    echo "== dbgf=$f"; MCVD_K1_DBGF=$f timeout 120 python tools/umma_timing.py $shp 2>&1 | tail -5
  done
done
} > gpurun_out/r2_k1_attr3.txt 2>&1
cat gpurun_out/r2_k1_attr3.txt
for fs in auto 1; do
MCVD_FUSE_SC=$fs timeout 600 python bench.py --steps 3 --warmup 3 2> gpurun_out/r2_bench_k1c_$fs.err | tail -1 > gpurun_out/r2_bench_k1c_$fs.json
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_k1c_$fs.json')); print('fuse=$fs frames/s', d['value'], 'e2e', d['e2e']['value'], 'psnr', d.get('psnr_vs_oracle_db')); print({k:round(v['ms_per_forward'],3) for k,v in d['roofline']['per_kind'].items()}); print(d['roofline'].get('frac'), d['clocks'])"
done

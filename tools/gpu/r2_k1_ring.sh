#!/bin/bash
# round-2: K1 operand ring for wide inputs + row-per-lane epilogue in the general kernel
cd "$(dirname "$0")/../.."
timeout 240 python -m pytest tests/test_gpu_ops.py -x -q -k "conv1x1" 2>&1 | tail -3
( time timeout 600 python -m pytest tests -x -q -m gpu ) > gpurun_out/r2_pytest_gpu_ring.log 2>&1
grep -n "passed\|failed\|Error\|Timeout" gpurun_out/r2_pytest_gpu_ring.log | tail -5
timeout 400 python bench.py --steps 3 --warmup 3 2> gpurun_out/r2_bench_ring.err | tail -1 > gpurun_out/r2_bench_ring.json
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_ring.json')); print('frames/s', d['value'], 'e2e', d['e2e']['value'], 'psnr', d.get('psnr_vs_oracle_db')); print({k:round(v['ms_per_forward'],3) for k,v in d['roofline']['per_kind'].items()}); print(d['roofline'].get('frac'), d['clocks'])"

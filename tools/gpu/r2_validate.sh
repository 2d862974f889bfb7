#!/bin/bash
# round-2 validation: the whole -m gpu suite, the operand-split experiment, a short default bench
cd "$(dirname "$0")/../.."
( time timeout 1500 python -m pytest tests -x -q -m gpu --durations=12 ) > gpurun_out/r2_pytest_gpu.log 2>&1
tail -30 gpurun_out/r2_pytest_gpu.log
timeout 900 python tools/split_experiment.py > gpurun_out/r2_split_experiment.txt 2> gpurun_out/r2_split_experiment.err
cat gpurun_out/r2_split_experiment.txt; tail -2 gpurun_out/r2_split_experiment.err
timeout 600 python bench.py --steps 2 --warmup 1 2> gpurun_out/r2_bench_a.err | tail -1 > gpurun_out/r2_bench_a.json
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_a.json')); print('frames/s', d['value'], 'e2e', d['e2e']['value'], 'psnr', d.get('psnr_vs_oracle_db'), 'cpu', d.get('cpu_baseline',{}).get('value'), d.get('cpu_baseline',{}).get('kind')); print({k:v['ms_per_forward'] for k,v in d['roofline']['per_kind'].items()}); print(d['roofline'].get('frac'), d['clocks'])"
tail -3 gpurun_out/r2_bench_a.err

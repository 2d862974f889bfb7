#!/bin/bash
# round-2 validation (second pass): full -m gpu suite, default bench, per-launch ncu metrics of one cfg2 forward
cd "$(dirname "$0")/../.."
( time timeout 1500 python -m pytest tests -x -q -m gpu --durations=5 ) > gpurun_out/r2_pytest_gpu.log 2>&1
tail -12 gpurun_out/r2_pytest_gpu.log
timeout 600 python bench.py --steps 3 --warmup 3 2> gpurun_out/r2_bench_b.err | tail -1 > gpurun_out/r2_bench_b.json
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_b.json')); print('frames/s', d['value'], 'e2e', d['e2e']['value'], 'psnr', d.get('psnr_vs_oracle_db'), 'cpu', d.get('cpu_baseline',{}).get('value'), d.get('cpu_baseline',{}).get('kind')); print({k:round(v['ms_per_forward'],3) for k,v in d['roofline']['per_kind'].items()}); print(d['roofline'].get('frac'), d['clocks'])"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_launch_metrics_cfg2_b64.csv python tools/profile_forward.py cfg2 64 > gpurun_out/r2_profile_forward.log 2>&1
python tools/launch_metrics_summary.py gpurun_out/r2_launch_metrics_cfg2_b64.csv | tee gpurun_out/r2_launch_metrics_cfg2_b64.txt

#!/bin/bash
# round-2 final validation: full -m gpu suite, default bench (with PSNR + CPU arm), fused-shortcut comparison,
# per-launch ncu metrics, cfg4 AR shard line with e2e clocks
cd "$(dirname "$0")/../.."
( time timeout 600 python -m pytest tests -x -q -m gpu --durations=5 ) > gpurun_out/r2_pytest_gpu_final.log 2>&1
grep -n "passed\|failed\|Error\|Timeout" gpurun_out/r2_pytest_gpu_final.log | tail -5
timeout 600 python bench.py --steps 3 --warmup 3 2> gpurun_out/r2_bench_final.err | tail -1 > gpurun_out/r2_bench_final.json
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_final.json')); print('frames/s', d['value'], 'e2e', d['e2e']['value'], 'psnr', d.get('psnr_vs_oracle_db'), 'cpu', d.get('cpu_baseline',{}).get('value')); print({k:round(v['ms_per_forward'],3) for k,v in d['roofline']['per_kind'].items()}); print(d['roofline'].get('frac'), d['clocks'])"
MCVD_FUSE_SC=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-psnr --no-roofline 2> gpurun_out/r2_bench_final_fused.err | tail -1 > gpurun_out/r2_bench_final_fused.json
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_final_fused.json')); print('fused shortcut: frames/s', d['value'], d['clocks'])"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_launch_metrics_final.csv python tools/profile_forward.py cfg2 64 > gpurun_out/r2_profile_forward.log 2>&1
python tools/launch_metrics_summary.py gpurun_out/r2_launch_metrics_final.csv | tee gpurun_out/r2_launch_metrics_final.txt
timeout 300 python bench.py --workload cfg4 --ar --batch 16 --steps 1 --warmup 1 --no-cpu-baseline --no-psnr --no-roofline 2> gpurun_out/r2_bench_cfg4_ar_b16.err | tail -1 > gpurun_out/r2_bench_cfg4_ar_b16.json
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_cfg4_ar_b16.json')); print('cfg4 AR B=16:', d['value'], d['unit'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e'])"

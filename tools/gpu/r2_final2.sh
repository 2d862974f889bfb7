#!/bin/bash
# round-2: residual rows prefetched into L2 by the waiting epilogue warps; transpose pads freed when unused
cd "$(dirname "$0")/../.."
( time timeout 600 python -m pytest tests -x -q -m gpu ) > gpurun_out/r2_pytest_gpu_final2.log 2>&1
grep -n "passed\|failed\|Error\|Timeout" gpurun_out/r2_pytest_gpu_final2.log | tail -5
{
for shp in "64 64 96 96 3" "64 64 96 96 3 0 res" "64 32 192 192 3 0 res" "64 32 192 192 1 0 res"; do
  timeout 120 python tools/umma_timing.py $shp 2>&1 | head -1
done
} > gpurun_out/r2_timing_final2.txt 2>&1
cat gpurun_out/r2_timing_final2.txt
timeout 600 python bench.py --steps 3 --warmup 3 2> gpurun_out/r2_bench_final2.err | tail -1 > gpurun_out/r2_bench_final2.json
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_final2.json')); print('frames/s', d['value'], 'e2e', d['e2e']['value'], 'psnr', d.get('psnr_vs_oracle_db'), 'cpu', d.get('cpu_baseline',{}).get('value')); print({k:round(v['ms_per_forward'],3) for k,v in d['roofline']['per_kind'].items()}); print(d['roofline'].get('frac'), d['clocks'])"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_launch_metrics_final2.csv python tools/profile_forward.py cfg2 64 > gpurun_out/r2_profile_forward.log 2>&1
python tools/launch_metrics_summary.py gpurun_out/r2_launch_metrics_final2.csv | tee gpurun_out/r2_launch_metrics_final2.txt | head -6

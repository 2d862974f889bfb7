#!/bin/bash
# round-2: full -m gpu suite + per-launch ncu metrics of one cfg2 forward on the current build
cd "$(dirname "$0")/../.."
( time timeout 900 python -m pytest tests -x -q -m gpu ) > gpurun_out/r2_pytest_gpu_cur.log 2>&1
tail -4 gpurun_out/r2_pytest_gpu_cur.log
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_launch_metrics_cur.csv python tools/profile_forward.py cfg2 64 > gpurun_out/r2_profile_forward.log 2>&1
python tools/launch_metrics_summary.py gpurun_out/r2_launch_metrics_cur.csv | tee gpurun_out/r2_launch_metrics_cur.txt

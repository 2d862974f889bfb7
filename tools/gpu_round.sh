#!/bin/bash
# Staged GPU check: every stage has its own timeout and log, so one hang or crash does not hide the rest.
# usage (under gpurun): bash tools/gpu_round.sh [stage ...]   stages: ops_other ops_umma model_simt model_umma smoke bench
mkdir -p gpurun_out
STAGES=${@:-"ops_other ops_umma model_simt model_umma smoke"}
nvidia-smi --query-gpu=name,driver_version,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
python -c "import os; print('cpus', os.cpu_count())" >> gpurun_out/gpu.txt
for s in $STAGES; do
  echo "=== stage $s"; t0=$(date +%s)
  case $s in
    ops_other)  timeout 600 python -m pytest tests/test_gpu_ops.py -q -x -k "not umma" > gpurun_out/$s.log 2>&1 ;;
    ops_umma)   timeout 600 python -m pytest tests/test_gpu_ops.py -q -k "umma" > gpurun_out/$s.log 2>&1 ;;
    model_simt) timeout 900 python -m pytest tests/test_gpu_model.py -q -k "simt" > gpurun_out/$s.log 2>&1 ;;
    model_umma) timeout 900 python -m pytest tests/test_gpu_model.py -q -k "not simt" > gpurun_out/$s.log 2>&1 ;;
    smoke)      timeout 600 python __graft_entry__.py --smoke > gpurun_out/$s.log 2>&1 ;;
    bench)      timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/$s.log ;;
    bench_simt) timeout 1500 python bench.py --conv simt --steps 1 --warmup 1 --no-psnr --no-cpu-baseline > gpurun_out/bench_simt.json 2> gpurun_out/$s.log ;;
    *) echo "unknown stage $s" ;;
  esac
  rc=$?; echo "stage $s rc=$rc secs=$(( $(date +%s) - t0 ))"; tail -n 25 gpurun_out/$s.log
done

#!/bin/bash
# round-2: 2-GPU bench (distinct clips per rank, all-gather + D2H inside the e2e region)
cd "$(dirname "$0")/../.."
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 3 2> gpurun_out/r2_bench_2gpu.err | tail -1 > gpurun_out/r2_bench_2gpu.json
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_2gpu.json')); print('2 GPUs: frames/s', d['value'], 'e2e', d['e2e'], d['clocks'])"
tail -3 gpurun_out/r2_bench_2gpu.err

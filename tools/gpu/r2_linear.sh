#!/bin/bash
# round-2: batched linear kernel (per-clip timesteps)
cd "$(dirname "$0")/../.."
( time timeout 400 python -m pytest tests -x -q -m gpu ) > gpurun_out/r2_pytest_gpu_final3.log 2>&1
grep -n "passed\|failed\|Error\|Timeout" gpurun_out/r2_pytest_gpu_final3.log | tail -5
timeout 200 python - <<'PY' 2>&1 | tail -4
import torch, time, sys
sys.path.insert(0, '.')
from mcvd_b200.synthetic import make_module
from mcvd_b200 import detfill
cfg, net, sd = make_module("cfg2", "cuda:0")
x, cond = detfill.synthetic_inputs(cfg, 64)
x, cond = x.cuda(), cond.cuda()
t = torch.full((64,), 500, device="cuda:0")
for _ in range(3): y = net(x, t, cond=cond)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): y = net(x, t, cond=cond)
e1.record(); torch.cuda.synchronize()
print("module.forward with a per-clip label tensor, cfg2 B=64: %.2f ms" % (e0.elapsed_time(e1) / 10))
PY

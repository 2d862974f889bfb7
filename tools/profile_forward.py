"""Run warm-up forwards, then ONE network evaluation between cudaProfilerStart/Stop (for ncu
--profile-from-start off).  usage: python tools/profile_forward.py [workload] [batch]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mcvd_b200 import detfill
from mcvd_b200.synthetic import make_module

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
cfg, net, sd = make_module(name, "cuda:0")
B = int(sys.argv[2]) if len(sys.argv) > 2 else cfg.bench_batch
x, cond = detfill.synthetic_inputs(cfg, B)
x, cond = x.cuda(), cond.cuda()
t = torch.full((B,), 500, dtype=torch.long, device="cuda")
eng = net.engine()
P = eng.program(B)
eng.set_inputs(P, x, t, cond)
for _ in range(2):
    eng.run_cond(P)
    eng.run_step(P)
torch.cuda.synchronize()
torch.cuda.profiler.start()
eng.run_step(P)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ops per forward:", len(P.step_ops), "umma:", P.n_umma, "simt:", P.n_simt)

"""Oracle (CPU port of the reference) forward time vs torch thread count -- picks the CPU baseline setup."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mcvd_b200 import detfill
from mcvd_b200.synthetic import make_module
from oracle import mcvd_oracle as O
name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
cfg, net, sd = make_module(name, "cpu")
print("cpu_count", os.cpu_count(), flush=True)
for B in (1, 2, 8):
    x, cond = detfill.synthetic_inputs(cfg, B)
    t = torch.full((B,), 500, dtype=torch.long)
    for th in (8, 16, 32, 64, 128):
        if th > (os.cpu_count() or 1): continue
        torch.set_num_threads(th)
        O.unet_forward(cfg, sd, x, t, cond)
        t0 = time.perf_counter(); O.unet_forward(cfg, sd, x, t, cond); dt = time.perf_counter() - t0
        print(f"B={B} threads={th} forward {dt:.3f} s  -> {B*cfg.data.num_frames/(dt*101):.3f} frames/s", flush=True)

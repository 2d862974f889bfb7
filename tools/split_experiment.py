"""Does a cheaper operand split pass the parity gate?  

Every tensor-core product is hi*hi + lo_a*hi_w + hi_a*lo_w (fp16 hi/lo of both operands).  This runs the full-size
forward of every BASELINE workload (1 clip, t = 500) with each cheaper variant of the CONV kernel and reports the
worst |error| / (atol + rtol*|ref|) against the oracle (gate: <= 1 with rtol 1e-3, atol 1e-4), plus the 10-step DDPM
PSNR of cfg2.  usage: split_experiment.py [workloads...]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mcvd_b200 import detfill, samplers
from mcvd_b200.synthetic import make_module, allclose_report
from oracle import mcvd_oracle as O

MODES = {3: "hi*hi + lo_a*hi_w + hi_a*lo_w (default)", 1: "hi*hi + lo_a*hi_w (fp16 weights)",
         2: "hi*hi + hi_a*lo_w (fp16 activations)", 4: "hi*hi only"}
names = sys.argv[1:] or ["cfg1", "cfg2", "cfg3", "cfg4", "cfg5"]
torch.set_num_threads(min(os.cpu_count() or 1, 16))
print("| workload | " + " | ".join(f"split {m}" for m in MODES) + " |")
for name in names:
    cfg, net, sd = make_module(name, "cuda:0")
    x, cond = detfill.synthetic_inputs(cfg, 1)
    tt = torch.tensor([500])
    ref = O.unet_forward(cfg, sd, x, tt, cond)
    eng = net.engine()
    cells = []
    for m in MODES:
        eng.split_mode = m
        eng.programs.clear()
        out = net(x.cuda(), tt.cuda(), cond=cond.cuda()).cpu()
        bad, mx, ratio = allclose_report(out, ref, 1e-3, 1e-4)
        cell = f"{ratio:.3f} ({'pass' if bad == 0 else f'FAIL {bad}'}; max abs {mx:.1e})"
        if name == "cfg2":
            L = 10
            zs = [detfill.normal(f"sz{i}", x.shape, seed=5) for i in range(L - 1)]
            o = samplers.ddpm_sampler(x.cuda(), net, cond=cond.cuda(), final_only=True, denoise=True, subsample_steps=L,
                                      noise_list=[z.cuda() for z in zs])[0].cpu()
            if m == 3:
                fn = lambda xx, t_, cc: O.unet_forward(cfg, sd, xx, t_, cc)
                orc = O.ddpm_sample(fn, O.make_schedule(cfg), x.clone(), cond, L, True, True, noise=zs)[0]
            to01 = lambda a: ((a + 1) / 2).clamp(0, 1)
            cell += f", 10-step PSNR {O.psnr01(to01(o), to01(orc)):.1f} dB"
        cells.append(cell)
    print(f"| {name} | " + " | ".join(cells) + " |", flush=True)
    del net, eng
    torch.cuda.empty_cache()

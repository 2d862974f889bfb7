"""Generate the golden fixtures in tests/golden/ from the UNMODIFIED reference.  TEST INFRASTRUCTURE.

Run in the build container (needs /root/reference):   python -m oracle.gen_golden

For each small workload it instantiates the reference ``UNetMore_DDPM`` on CPU, re-randomises the
weights with the deterministic hash fill (``mcvd_b200.detfill``; seed 1234), and records
  * the forward output eps for t in {0, 37, 990} on the synthetic (x_T, cond),
  * the output of the reference ``ddpm_sampler`` / ``ddim_sampler`` / ``FPNDM_sampler`` (L = the
    workload's subsample) with per-step noise injected by patching ``torch.randn_like`` around the
    unmodified reference call,
  * the autoregressive video_gen block loop (restated; the reference's loop is entangled with
    dataset / metric code) driven by the reference sampler.
Inputs and weights are NOT stored (they regenerate bit-identically from the hash); only outputs are.
"""
from __future__ import annotations

import os
import sys
from unittest import mock

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from mcvd_b200 import configs, detfill          # noqa: E402
from oracle import ref_import, mcvd_oracle as O  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
T_VALUES = (0, 37, 990)


def step_noise(shape, L, tag="z"):
    return [detfill.normal(f"{tag}{i}", shape) for i in range(L - 1)]


def gen(name):
    cfg = configs.workload(name)
    net = ref_import.build_reference_net(cfg)
    _, ddpm, ddim, fpndm = ref_import.ref_models()
    B = cfg.bench_batch
    x, cond = detfill.synthetic_inputs(cfg, B)
    L = cfg.sampling.subsample
    out = {}
    with torch.no_grad():
        for t in T_VALUES:
            out[f"eps_t{t}"] = net(x, torch.full((B,), t, dtype=torch.long), cond=cond).numpy()
        zs = step_noise(x.shape, L)
        it = iter(zs)
        with mock.patch("torch.randn_like", lambda _x: next(it)):
            out["ddpm"] = ddpm(x.clone(), net, cond=cond, final_only=True, denoise=True, subsample_steps=L,
                               clip_before=True, log=False, verbose=False)[0].numpy()
        # init_prev_t warm start (t_min > 0): first randn_like call is the re-noising, then the per-step noise
        it2 = iter([detfill.normal("warm", x.shape)] + zs)
        with mock.patch("torch.randn_like", lambda _x: next(it2)):
            out["ddpm_tmin"] = ddpm(x.clone(), net, cond=cond, final_only=True, denoise=True, subsample_steps=L,
                                    clip_before=True, log=False, verbose=False, t_min=0.35)[0].numpy()
        out["ddim"] = ddim(x.clone(), net, cond=cond, final_only=True, denoise=True, subsample_steps=L,
                           clip_before=True, log=False, verbose=False)[0].numpy()
        out["fpndm"] = fpndm(x.clone(), net, cond=cond, final_only=True, subsample_steps=L,
                             clip_before=True, log=False, verbose=False)[0].numpy()
        # AR loop (runner:1501-1570 restated in the oracle) around the reference sampler
        nfp = cfg.sampling.num_frames_pred
        n_iter = -(-nfp // cfg.data.num_frames)
        inits = [detfill.normal(f"ar_init{i}", x.shape) for i in range(n_iter)]

        def sampler(x_T, c, i):
            zi = iter(step_noise(x.shape, L, tag=f"ar{i}_z"))
            with mock.patch("torch.randn_like", lambda _x: next(zi)):
                return ddpm(x_T.clone(), net, cond=c, final_only=True, denoise=True, subsample_steps=L,
                            clip_before=True, log=False, verbose=False)
        out["video"] = O.video_gen_loop(cfg, sampler, cond, inits, nfp).numpy()
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, f"{name}.npz")
    np.savez_compressed(path, **out)
    print(name, {k: v.shape for k, v in out.items()}, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    assert ref_import.available(), "reference tree not found"
    for n in (sys.argv[1:] or ["tiny", "tiny_spade", "tiny_rgb", "cfg1"]):
        gen(n)

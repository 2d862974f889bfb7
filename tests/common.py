"""Shared helpers for the test-suite (test infrastructure)."""
import os

import numpy as np
import torch

from mcvd_b200 import configs, detfill
from mcvd_b200.model import UNetMore_DDPM

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def make_module(name, device="cpu", seed=1234):
    """mcvd_b200 module with deterministic re-randomised weights (same fill as the goldens)."""
    cfg = configs.workload(name)
    cfg.device = torch.device(device)
    net = UNetMore_DDPM(cfg)
    sd = net.state_dict()
    detfill.randomize_state_dict(sd, seed)
    net.load_state_dict(sd)
    net = net.to(device).eval()
    return cfg, net, {k: v.clone().cpu() for k, v in sd.items()}


def golden(name):
    return np.load(os.path.join(GOLDEN_DIR, f"{name}.npz"))


def step_noise(shape, L, tag="z"):
    return [detfill.normal(f"{tag}{i}", shape) for i in range(L - 1)]


def max_err(a, b):
    return float((a.double() - b.double()).abs().max())


def allclose_report(a, b, rtol, atol):
    d = (a.double() - b.double()).abs()
    lim = atol + rtol * b.double().abs()
    bad = int((d > lim).sum())
    return bad, float(d.max()), float((d / lim).max())

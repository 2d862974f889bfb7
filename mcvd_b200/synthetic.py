"""Synthetic model instances for tests, smoke() and bench.py: the module with deterministic
re-randomised weights (there is no network for checkpoints; the default init zeroes half the net)."""
import torch

from . import configs, detfill
from .model import UNetMore_DDPM


def make_module(name, device="cpu", seed=1234):
    """-> (config, mcvd_b200 module on `device` in eval mode, CPU copy of its state_dict)."""
    cfg = configs.workload(name) if isinstance(name, str) else name
    cfg.device = torch.device(device)
    net = UNetMore_DDPM(cfg)
    sd = net.state_dict()
    detfill.randomize_state_dict(sd, seed)
    net.load_state_dict(sd)
    net = net.to(device).eval()
    return cfg, net, {k: v.clone().cpu() for k, v in sd.items()}


def allclose_report(a, b, rtol, atol):
    """(# elements violating |a-b| <= atol + rtol*|b|, max abs err, max err/limit)."""
    d = (a.double() - b.double()).abs()
    lim = atol + rtol * b.double().abs()
    return int((d > lim).sum()), float(d.max()), float((d / lim).max())

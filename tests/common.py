"""Shared helpers for the test-suite (test infrastructure)."""
import os

import numpy as np
import torch

from mcvd_b200 import configs, detfill

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


from mcvd_b200.synthetic import make_module, allclose_report  # noqa: E402,F401


def golden(name):
    return np.load(os.path.join(GOLDEN_DIR, f"{name}.npz"))


def step_noise(shape, L, tag="z"):
    return [detfill.normal(f"{tag}{i}", shape) for i in range(L - 1)]


def max_err(a, b):
    return float((a.double() - b.double()).abs().max())



"""Shared helpers for the parity tests."""
import os

import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# Tolerances, from BASELINE.json north_star: fp32 within 1e-4 relative, bf16 within 2e-2 relative
# (relative = max |a - b| / max |b| over the tensor), indexing bit-exact.
TOL = {torch.float32: 1e-4, torch.bfloat16: 2e-2}


def relerr(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def load_golden(name):
    return torch.load(os.path.join(GOLDEN, name + ".pt"), map_location="cpu", weights_only=False)


def assert_close(a, b, tol, what=""):
    e = relerr(a, b)
    assert e <= tol, f"{what}: relative error {e:.3e} > {tol:.1e}"
    return e

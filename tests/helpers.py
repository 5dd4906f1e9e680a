"""Shared test helpers (fixtures loading, comparisons)."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    out = {}
    for k in z.files:
        a = z[k]
        out[k] = torch.from_numpy(a) if a.ndim > 0 else a.item()
    return out


def assert_close(a, b, atol, rtol=0.0, msg=""):
    a = torch.as_tensor(a).detach().cpu().double()
    b = torch.as_tensor(b).detach().cpu().double()
    assert a.shape == b.shape, f"{msg}: shape {tuple(a.shape)} vs {tuple(b.shape)}"
    err = (a - b).abs()
    tol = atol + rtol * b.abs()
    bad = err > tol
    assert not bad.any(), f"{msg}: max err {err.max().item():.3e} (tol {atol:g}+{rtol:g}*|b|), {int(bad.sum())} bad of {bad.numel()}"


def psnr_db(a, b):
    mse = torch.mean((a.double() - b.double()) ** 2).clamp_min(1e-30)
    return float(-10.0 * torch.log10(mse))

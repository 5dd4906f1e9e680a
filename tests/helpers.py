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
    # NaN-aware: a NaN error compares False against any tolerance, so "bad" is "not within", and a NaN / inf on one
    # side only is always an error (identical non-finite values on both sides -- e.g. the NaN mean of an empty mask,
    # which the reference produces too -- are equal)
    same_nonfinite = (a == b) | (torch.isnan(a) & torch.isnan(b))
    bad = ~((err <= tol) | same_nonfinite)
    if bad.any():
        fin = err[torch.isfinite(err)]
        mx = fin.max().item() if fin.numel() else float("nan")
        nn_ = int((~torch.isfinite(a) & ~same_nonfinite).sum())
        raise AssertionError(f"{msg}: max err {mx:.3e} (tol {atol:g}+{rtol:g}*|b|), {int(bad.sum())} bad of "
                             f"{bad.numel()}, {nn_} non-finite in the result")


def psnr_db(a, b):
    mse = torch.mean((a.double() - b.double()) ** 2).clamp_min(1e-30)
    return float(-10.0 * torch.log10(mse))

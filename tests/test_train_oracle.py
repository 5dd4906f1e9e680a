"""Pins oracle/train_oracle.py (losses, clip + Adam) to fixtures produced by the reference's
own models/loss.py / mse2psnr and the installed torch.optim.Adam / clip_grad_norm_ (tools/make_golden_train.py).  CPU."""
import math

import pytest
import torch

from oracle import train_oracle as to
from helpers import load_golden, assert_close


def test_photometric():
    g = load_golden("train_photometric")
    rgb0 = g["rgb0"].clone().requires_grad_(True)
    rgb = g["rgb"].clone().requires_grad_(True)
    loss, m0, m1, p0, p1 = to.photometric_loss(rgb0, rgb, g["target"])
    loss.backward()
    assert_close(m0, g["mse0"], 0, 1e-6)
    assert_close(m1, g["mse"], 0, 1e-6)
    assert_close(m1, g["img2mse"], 0, 1e-6)
    assert_close(p0, g["psnr0"].reshape(()), 1e-5)
    assert_close(p1, g["psnr"].reshape(()), 1e-5)
    assert_close(rgb0.grad, g["g_rgb0"], 1e-9, 1e-6)
    assert_close(rgb.grad, g["g_rgb"], 1e-9, 1e-6)
    loss1 = to.photometric_loss(None, rgb.detach(), g["target"])
    assert_close(loss1[0], g["mse"], 0, 1e-6)


@pytest.mark.parametrize("tag", ["a", "b"])
def test_dsnerf_losses(tag):
    g = load_golden("train_dsnerf_" + tag)
    depth = g["depth"].clone().requires_grad_(True)
    dl = to.compute_depth_loss(depth, g["depths"], g["near"], g["far"])
    dl.backward()
    assert_close(dl, g["depth_loss"], 0, 1e-6)
    assert_close(depth.grad, g["g_depth"], 1e-9, 1e-6)
    w = g["weights"].clone().requires_grad_(True)
    sl = to.compute_sigma_loss(w, g["z_vals"], g["dists"], g["depths"], g["near"], g["far"], err=1)
    sl.backward()
    assert_close(sl, g["sigma_loss"], 0, 1e-6)
    assert_close(w.grad, g["g_weights"], 1e-9, 1e-6)
    w.grad = None
    pr = to.compute_sigma_loss_per_ray(w, g["z_vals"], g["dists"], g["depths"], err=1)
    (pr * g["per_ray_coef"]).sum().backward()
    assert_close(pr, g["per_ray"], 1e-9, 1e-6)
    assert_close(w.grad, g["g_weights_per_ray"], 1e-9, 1e-6)
    sl2 = to.compute_sigma_loss(w.detach(), g["z_vals"], g["dists"], g["depths"], g["near"], g["far"], err=0.25)
    assert_close(sl2, g["sigma_loss_err025"], 0, 1e-6)


def adam_case(g):
    n_groups = [int(n) for n in g["n_groups"]]
    total = sum(n_groups)
    init = [g["init.p%d" % i] for i in range(total)]
    params, i = [], 0
    for n in n_groups:
        params.append(init[i:i + n])
        i += n
    grads = [[g["g%d.%d" % (t, k)] for k in range(total)] for t in range(int(g["steps"]))]
    return params, grads, [float(x) for x in g["lrs"]]


@pytest.mark.parametrize("tag,max_norm", [("clip", 1.0), ("noclip", None)])
def test_clip_and_adam_matches_fixture(tag, max_norm):
    g = load_golden("train_adam")
    params, grads, lrs = adam_case(g)
    p, m, v, norms = to.clip_and_adam(params, grads, lrs, max_norm=max_norm)
    for i in range(len(p)):
        assert_close(p[i], g["%s.p%d" % (tag, i)], 1e-9, 1e-6, "p%d" % i)
        assert_close(m[i], g["%s.m%d" % (tag, i)], 1e-12, 1e-6, "m%d" % i)
        assert_close(v[i], g["%s.v%d" % (tag, i)], 1e-16, 1e-6, "v%d" % i)
    if max_norm is not None:
        assert_close(torch.stack(norms), g["clip.norms"], 0, 1e-6)


def test_adam_recurrence_restated():
    """The explicit recurrence (what csrc/train_step.cu implements) against the torch optimiser's fixture."""
    g = load_golden("train_adam")
    params, grads, lrs = adam_case(g)
    norms = g["clip.norms"]
    coefs = [min(1.0, 1.0 / (float(n) + 1e-6)) for n in norms]
    flat_lr = [lr for grp, lr in zip(params, lrs) for _ in grp]
    k = 0
    for grp in params:
        for p0 in grp:
            p, m, v = to.adam_restated(p0, [gs[k] for gs in grads], flat_lr[k], clip_coefs=coefs)
            assert_close(p, g["clip.p%d" % k], 2e-8, 2e-6, "p%d" % k)
            assert_close(m, g["clip.m%d" % k], 1e-8, 2e-6, "m%d" % k)     # m cancels: error scales with |g|
            assert_close(v, g["clip.v%d" % k], 1e-14, 2e-6, "v%d" % k)
            k += 1
    assert math.isclose(coefs[3], 1.0)      # a step whose norm is below the threshold is not scaled


@pytest.mark.parametrize("tag", ["a", "b", "empty"])
def test_compute_2d_iou(tag):
    g = load_golden("train_iou2d")
    iou, masks = to.compute_2d_iou(g[tag + ".T"], g[tag + ".sem"], 0.1)
    assert iou == float(g[tag + ".iou"])
    assert (torch.from_numpy(masks) == g[tag + ".masks"]).all()
    if tag == "empty":
        assert not masks.any()

"""Generates tests/golden/train_*.npz: the reference's own loss functions (models/loss.py, models/rendering__.py
img2mse / mse2psnr, imported UNMODIFIED from /root/reference) and the installed torch.optim.Adam /
torch.nn.utils.clip_grad_norm_ (what the reference's configure_optimizers / Trainer call) on seeded inputs.

Run in the build container only:   python tools/make_golden_train.py
tests/test_train_oracle.py pins oracle/train_oracle.py to these fixtures; the -m gpu tests pin the kernels to them."""
import importlib.util
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_harness  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def load_ref_loss():
    path = os.path.join(ref_harness.REFERENCE_ROOT, "models", "loss.py")
    spec = importlib.util.spec_from_file_location("ref_models_loss", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load_ref_function(relpath, name):
    """One function of a reference module whose other imports are not installed (utils/metrics.py needs pytorch3d):
    its source is cut out of the file by the parser and executed unmodified."""
    import ast
    path = os.path.join(ref_harness.REFERENCE_ROOT, relpath)
    tree = ast.parse(open(path).read())
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == name)
    ns = {"np": np, "torch": torch}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), path, "exec"), ns)
    return ns[name]


def save(name, **kw):
    arrs = {k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in kw.items()}
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **arrs)
    print("wrote", path, "%.1f KB" % (os.path.getsize(path) / 1024))


def render_like(R, S, gen, near, far):
    """weights / z_vals / dists shaped like a render pass's outputs (sorted z, weights summing to <= 1, some exact
    zeros and a far-cap last distance as raw2outputs produces)."""
    z = near + (far - near) * torch.sort(torch.rand(R, S, generator=gen), dim=1).values
    dists = torch.cat([z[:, 1:] - z[:, :-1], torch.full((R, 1), 1e10)], dim=1) * 1.1
    w = torch.softmax(4.0 * torch.randn(R, S, generator=gen), dim=1) * torch.rand(R, 1, generator=gen)
    w[torch.rand(R, S, generator=gen) < 0.05] = 0.0
    return w, z, dists


def main():
    torch.set_num_threads(4)
    ref = ref_harness.load_reference()
    rl = load_ref_loss()
    R_ = ref.rendering
    gen = torch.Generator().manual_seed(11)

    # ---- photometric: nn.MSELoss twice + mse2psnr, with autograd gradients of the summed loss
    R = 257
    rgb0 = torch.rand(R, 3, generator=gen, requires_grad=True)
    rgb = torch.rand(R, 3, generator=gen, requires_grad=True)
    target = torch.rand(R, 3, generator=gen)
    mse_fn = torch.nn.MSELoss()
    l0, l1 = mse_fn(rgb0, target), mse_fn(rgb, target)
    (l0 + l1).backward()
    save("train_photometric", rgb0=rgb0, rgb=rgb, target=target, mse0=l0, mse=l1, psnr0=R_.mse2psnr(l0.detach()),
         psnr=R_.mse2psnr(l1.detach()), img2mse=R_.img2mse(rgb.detach(), target), g_rgb0=rgb0.grad, g_rgb=rgb.grad)

    # ---- DS-NeRF losses
    near, far = 0.03, 0.80
    for tag, (R, S) in (("a", (193, 64)), ("b", (50, 37))):
        w, z, dists = render_like(R, S, gen, near, far)
        depths = near - 0.1 + (far - near + 0.2) * torch.rand(R, generator=gen)      # some outside (near, far)
        depth = (near + (far - near) * torch.rand(R, generator=gen)).requires_grad_(True)
        dl = rl.compute_depth_loss(depth, depths, near, far)
        dl.backward()
        w.requires_grad_(True)
        sl = rl.compute_sigma_loss(w, z, dists, depths, near, far, err=1)
        sl.backward()
        g_w = w.grad.clone()
        w.grad = None
        pr = rl.compute_sigma_loss_per_ray(w, z, dists, depths, err=1)
        coef = torch.rand(R, generator=gen)
        (pr * coef).sum().backward()
        sl2 = rl.compute_sigma_loss(w.detach(), z, dists, depths, near, far, err=0.25)
        save("train_dsnerf_" + tag, near=near, far=far, weights=w, z_vals=z, dists=dists, depths=depths, depth=depth,
             depth_loss=dl, g_depth=depth.grad, sigma_loss=sl, g_weights=g_w, per_ray=pr, per_ray_coef=coef,
             g_weights_per_ray=w.grad, sigma_loss_err025=sl2)

    # ---- 2-D IoU of the online-tracking evaluation (utils/metrics.py:527-550)
    iou_fn = load_ref_function(os.path.join("utils", "metrics.py"), "compute_2d_iou")
    out = {}
    for tag, (R, V, p_vehicle) in (("a", (1000, 2, 0.2)), ("b", (4099, 5, 0.05)), ("empty", (300, 3, 0.0))):
        T = torch.rand(R, V, generator=gen)
        T[torch.rand(R, V, generator=gen) < (0.15 if tag != "empty" else 0.0)] *= 0.05       # some pixels covered by an object
        if tag == "empty":
            T = 0.5 + 0.5 * T
        T[0, 0] = float("nan")
        T[1, 0] = 0.1                                                                 # the threshold itself: not below
        sem = torch.rand(R, generator=gen) < p_vehicle
        iou, masks = iou_fn(T, sem, 0.1)
        out.update({tag + ".T": T, tag + ".sem": sem, tag + ".iou": np.float64(iou), tag + ".masks": masks})
    save("train_iou2d", **out)

    # ---- clip_grad_norm_ + Adam: 3 groups (static nets, dynamic nets, poses) with the reference's learning rates
    shapes = [[(64, 63), (64,), (3, 128), (3,)], [(96, 33), (1, 64), (1,)], [(5, 7)]]
    lrs = [5e-4, 5e-4 * 0.5, 1e-3]
    params = [[0.1 * torch.randn(*s, generator=gen) for s in grp] for grp in shapes]
    steps = 6
    grads = []
    for t in range(steps):
        scale = [3.0, 0.02, 1.0, 1e-6, 40.0, 0.3][t]       # norms above and below the clip threshold
        grads.append([scale * torch.randn(*s, generator=gen) for grp in shapes for s in grp])
    sys.path.insert(0, ROOT)
    from oracle import train_oracle as to
    out = {}
    for tag, max_norm in (("clip", 1.0), ("noclip", None)):
        p, m, v, norms = to.clip_and_adam(params, grads, lrs, max_norm=max_norm)
        for i, (pi, mi, vi) in enumerate(zip(p, m, v)):
            out["%s.p%d" % (tag, i)] = pi
            out["%s.m%d" % (tag, i)] = mi
            out["%s.v%d" % (tag, i)] = vi
        if norms:
            out[tag + ".norms"] = torch.stack(norms)
    for i, pi in enumerate([p for grp in params for p in grp]):
        out["init.p%d" % i] = pi
    for t, gs in enumerate(grads):
        for i, gi in enumerate(gs):
            out["g%d.%d" % (t, i)] = gi
    save("train_adam", lrs=np.asarray(lrs), steps=steps, n_groups=np.asarray([len(g) for g in shapes]), **out)


if __name__ == "__main__":
    main()

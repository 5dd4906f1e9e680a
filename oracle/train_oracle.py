"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

CPU restatement of what sits either side of the render path inside one training step (SURVEY.md section 8(f)
rows 2-3), each function citing the reference lines it follows (paths relative to the reference root):

  photometric_loss      train_online__.py:158-166 (nn.MSELoss :84) + models/rendering__.py:22-23 mse2psnr
  compute_depth_loss    models/loss.py:4-10
  compute_sigma_loss    models/loss.py:13-66        compute_sigma_loss_per_ray  models/loss.py:70-87
  clip_and_adam         train_online__.py:333-353 (torch.optim.Adam, betas (0.9, 0.999)) and the Trainer's
                        gradient_clip_val=1.0 (:1170, i.e. torch.nn.utils.clip_grad_norm_)

  compute_2d_iou        utils/metrics.py:527-550

PINNED: tools/make_golden_train.py runs the unmodified reference `models/loss.py` and `models/rendering__.py`
(imported from /root/reference) and the installed torch.optim.Adam / clip_grad_norm_ -- the very code the reference
calls -- on seeded inputs and commits tests/golden/train_*.npz; tests/test_train_oracle.py checks this file against
them.  `adam_restated` is the explicit per-element recurrence; `clip_and_adam` drives the real torch optimiser."""
import math

import torch

EPS_F32 = torch.finfo(torch.float32).eps


def photometric_loss(rgb0, rgb, target):
    """-> (loss, img_loss0, img_loss, psnr0, psnr); rgb0 may be None (N_importance == 0)."""
    def psnr(m):
        return -10.0 * torch.log(m) / torch.log(torch.tensor([10.0]))
    mse = torch.mean((rgb - target) ** 2)
    if rgb0 is None:
        return mse, torch.zeros(()), mse, torch.zeros(()), psnr(mse).reshape(())
    mse0 = torch.mean((rgb0 - target) ** 2)
    return mse0 + mse, mse0, mse, psnr(mse0).reshape(()), psnr(mse).reshape(())


def compute_depth_loss(depth, gt_depth, near, far):
    mask = torch.logical_and(gt_depth < far, gt_depth > near)
    return torch.mean(((depth[mask] - gt_depth[mask]) / gt_depth[mask]) ** 2)


def _sigma_terms(weights, z_vals, dists, depths, err):
    w = torch.where(weights <= 0, torch.full_like(weights, EPS_F32), weights)
    return -torch.log(w) * torch.exp(-((z_vals - depths[:, None]) ** 2) / (2 * err)) * dists


def compute_sigma_loss(weights, z_vals, dists, depths, near, far, err=1):
    mask = torch.logical_and(depths < far, depths > near)
    t = _sigma_terms(weights[mask], z_vals[mask], dists[mask], depths[mask], err)
    return torch.sum(t, dim=1).mean()


def compute_sigma_loss_per_ray(weights, z_vals, dists, depths, err=1):
    return torch.sum(_sigma_terms(weights, z_vals, dists, depths, err), dim=1)


def clip_and_adam(params, grad_steps, lrs, betas=(0.9, 0.999), eps=1e-8, max_norm=None):
    """Drives torch.optim.Adam (one param group per entry of `lrs`; params = list of lists of tensors) and, when
    max_norm is given, torch.nn.utils.clip_grad_norm_ over all parameters before every step -- what HybridOptim under
    Trainer(gradient_clip_val=...) does.  grad_steps: list over steps of flat lists of gradients (group order).
    Returns (params after the last step, exp_avg, exp_avg_sq, total norms per step)."""
    ps = [[torch.nn.Parameter(p.clone()) for p in grp] for grp in params]
    opt = torch.optim.Adam([{"params": grp, "lr": lr} for grp, lr in zip(ps, lrs)], betas=betas, eps=eps)
    flat = [p for grp in ps for p in grp]
    norms = []
    for grads in grad_steps:
        for p, g in zip(flat, grads):
            p.grad = g.clone()
        if max_norm is not None:
            norms.append(torch.nn.utils.clip_grad_norm_(flat, max_norm).detach().clone())
        opt.step()
    return ([p.detach().clone() for p in flat], [opt.state[p]["exp_avg"].clone() for p in flat],
            [opt.state[p]["exp_avg_sq"].clone() for p in flat], norms)


def adam_restated(p, grads, lr, betas=(0.9, 0.999), eps=1e-8, clip_coefs=None):
    """The recurrence torch.optim.Adam implements (amsgrad=False, weight_decay=0), one tensor, fp32:
    m <- m + (1-b1)(g - m);  v <- b2 v + (1-b2) g^2;  p <- p - lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps)."""
    p = p.clone()
    m = torch.zeros_like(p)
    v = torch.zeros_like(p)
    b1, b2 = betas
    for t, g in enumerate(grads, start=1):
        if clip_coefs is not None:
            g = g * clip_coefs[t - 1]
        m = m + (1 - b1) * (g - m)
        v = v * b2 + (1 - b2) * g * g
        denom = v.sqrt() / math.sqrt(1 - b2 ** t) + eps
        p = p - (lr / (1 - b1 ** t)) * (m / denom)
    return p, m, v


def compute_2d_iou(dynamic_transmittance, semantic_mask, thres=0.1):
    """utils/metrics.py:527-550 (numpy loops over the objects, as there)."""
    import numpy as np
    num_rays, num_vehicles = dynamic_transmittance.shape
    sem = semantic_mask.detach().cpu().numpy().astype(bool)
    union_pred = np.zeros((num_rays,), dtype=bool)
    masks = np.zeros((num_vehicles, num_rays), dtype=bool)
    for i in range(num_vehicles):
        masks[i] = (dynamic_transmittance[:, i] < thres).cpu().numpy()
        union_pred = np.logical_or(union_pred, masks[i])
    union = np.count_nonzero(np.logical_or(sem, union_pred))
    inter = np.count_nonzero(np.logical_and(sem, union_pred))
    return (0 if union == 0 else inter / union), masks

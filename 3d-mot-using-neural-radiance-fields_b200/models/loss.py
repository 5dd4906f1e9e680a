"""Mirror of the reference's `models/loss.py` (DS-NeRF depth / sigma losses) plus the fused photometric loss of the
training step, over the C ABI (csrc/train_step.cu).  Same function names, argument
order and results as the reference; the arithmetic runs in CUDA kernels (no torch fallback)."""
import torch
from torch.autograd import Function

from .. import _capi
from .._capi import check, f32, ptr, stream
from ..functional import _c, _count


def _ws(device):
    return torch.empty((_capi.lib().star_train_ws_bytes(),), device=device, dtype=torch.uint8)


def _scalar(g, device):
    """Upstream gradient of a scalar loss as a 1-element contiguous fp32 device tensor."""
    return g.detach().reshape(1).to(device=device, dtype=torch.float32).contiguous()


class _DepthLoss(Function):
    @staticmethod
    def forward(ctx, depth, gt_depth, near, far):
        depth, gt = _c(depth), _c(gt_depth)
        out = torch.empty((2,), device=depth.device)
        check(_capi.lib().star_depth_loss_forward(f32(depth), f32(gt), depth.numel(), near, far, f32(out),
                                                  ptr(_ws(depth.device)), stream()), "star_depth_loss_forward")
        _count()
        ctx.save_for_backward(depth, gt, out)
        ctx.nf = (near, far)
        return out[0].clone()

    @staticmethod
    def backward(ctx, g):
        depth, gt, out = ctx.saved_tensors
        g_depth = torch.empty_like(depth)
        check(_capi.lib().star_depth_loss_backward(f32(depth), f32(gt), depth.numel(), ctx.nf[0], ctx.nf[1], f32(out),
                                                   f32(_scalar(g, depth.device)), f32(g_depth), stream()),
              "star_depth_loss_backward")
        _count()
        return g_depth, None, None, None


def compute_depth_loss(depth, gt_depth, near, far):
    """models/loss.py:4-10: mean over the rays with near < gt_depth < far of ((depth - gt_depth) / gt_depth)^2."""
    if depth.shape != gt_depth.shape:
        raise ValueError("depth and gt_depth must have the same shape")
    return _DepthLoss.apply(depth, gt_depth, float(near), float(far))


class _SigmaLoss(Function):
    @staticmethod
    def forward(ctx, weights, z_vals, dists, depths, near, far, err, per_ray):
        weights, z_vals, dists, depths = _c(weights), _c(z_vals), _c(dists), _c(depths)
        R, S = weights.shape
        dev = weights.device
        out = torch.empty((2,), device=dev)
        rays = torch.empty((R,), device=dev) if per_ray else None
        check(_capi.lib().star_sigma_loss_forward(f32(weights), f32(z_vals), f32(dists), f32(depths), R, S, near, far,
                                                  err, f32(out), f32(rays) if per_ray else None, ptr(_ws(dev)),
                                                  stream()), "star_sigma_loss_forward")
        _count()
        ctx.save_for_backward(weights, z_vals, dists, depths, out)
        ctx.cfg = (near, far, err, per_ray)
        return rays if per_ray else out[0].clone()

    @staticmethod
    def backward(ctx, g):
        weights, z_vals, dists, depths, out = ctx.saved_tensors
        near, far, err, per_ray = ctx.cfg
        R, S = weights.shape
        g_w = torch.empty_like(weights)
        gg = _c(g.detach().float()) if per_ray else _scalar(g, weights.device)
        check(_capi.lib().star_sigma_loss_backward(f32(weights), f32(z_vals), f32(dists), f32(depths), R, S, near, far,
                                                   err, f32(out), f32(gg), 1 if per_ray else 0, f32(g_w), stream()),
              "star_sigma_loss_backward")
        _count()
        return g_w, None, None, None, None, None, None, None


def _check_sigma_args(weights, z_vals, dists, depths):
    if weights.dim() != 2 or z_vals.shape != weights.shape or dists.shape != weights.shape:
        raise ValueError("weights, z_vals and dists must share the shape [N_rays, N_samples]")
    if depths.shape != weights.shape[:1]:
        raise ValueError("depths must have the shape [N_rays]")


def compute_sigma_loss(weights, z_vals, dists, depths, near, far, err=1):
    """models/loss.py:13-66.  Gradient flows to `weights` only (z_vals / dists carry none on the render path)."""
    _check_sigma_args(weights, z_vals, dists, depths)
    return _SigmaLoss.apply(weights, z_vals, dists, depths, float(near), float(far), float(err), False)


def compute_sigma_loss_per_ray(weights, z_vals, dists, depths, err=1):
    """models/loss.py:70-87: the per-ray sums without the near/far mask (callbacks/check_batch_grad.py)."""
    _check_sigma_args(weights, z_vals, dists, depths)
    return _SigmaLoss.apply(weights, z_vals, dists, depths, float("-inf"), float("inf"), float(err), True)


class _Photometric(Function):
    @staticmethod
    def forward(ctx, rgb0, rgb, target):
        rgb, target = _c(rgb), _c(target)
        rgb0 = _c(rgb0) if rgb0 is not None else None
        dev = rgb.device
        out = torch.empty((5,), device=dev)
        need0 = rgb0 is not None and ctx.needs_input_grad[0]
        need1 = ctx.needs_input_grad[1]
        g0 = torch.empty_like(rgb0) if need0 else None
        g1 = torch.empty_like(rgb) if need1 else None
        check(_capi.lib().star_photometric_loss(f32(rgb0) if rgb0 is not None else None, f32(rgb), f32(target),
                                                rgb.numel(), f32(out), f32(g0) if need0 else None,
                                                f32(g1) if need1 else None, ptr(_ws(dev)), stream()),
              "star_photometric_loss")
        _count()
        ctx.g = (g0, g1)
        ctx.set_materialize_grads(False)
        loss, mse0, mse, psnr0, psnr = out[4], out[0], out[1], out[2], out[3]
        ctx.mark_non_differentiable(psnr0, psnr)
        return loss, mse0, mse, psnr0, psnr

    @staticmethod
    def backward(ctx, g_loss, g_mse0, g_mse, _p0, _p1):
        def scaled(g, a, b):
            if g is None or (a is None and b is None):
                return None
            return g * (a if b is None else b if a is None else a + b)
        g0, g1 = ctx.g
        return scaled(g0, g_loss, g_mse0), scaled(g1, g_loss, g_mse), None


def photometric_loss(rgb0, rgb, target):
    """The image terms of `training_step` (train_online__.py:158-166, train_app_init__.py): one pass computes
    img_loss0 = MSELoss(rgb0, target), img_loss = MSELoss(rgb, target), their sum, both PSNRs
    (models/rendering__.py:22-23) and the gradients of both rgb maps.  rgb0 may be None (N_importance == 0).
    Returns (loss, img_loss0, img_loss, psnr0, psnr) as 0-dim tensors."""
    if rgb.shape != target.shape or (rgb0 is not None and rgb0.shape != target.shape):
        raise ValueError("rgb maps and target must share one shape")
    return _Photometric.apply(rgb0, rgb, target)

"""Render orchestration with the reference's public surface (models/rendering__.py of
burakcuhadar/3D-MOT-using-Neural-Radiance-Fields): same function names, positional arguments and
output dictionaries; every tensor op runs in the sm_100a kernels behind include/star_b200.h.

Keyword-only extras (`t_rand=`, `u=`, `noise=`) inject the random draws the reference takes from
torch.rand / torch.randn, so that parity tests can run "with fixed noise"; `z_samples=` injects the
fine sample positions themselves (the output of sample_pdf) for stage-wise parity of the fine pass."""
import numpy as np
import torch

from .. import functional as F_
from .types__ import NERF_KEYS


def img2mse(img1, img2):
    return torch.mean((img1 - img2) ** 2)


def mse2psnr(mse):
    return -10.0 * torch.log(mse) / torch.log(torch.tensor([10.0], device=mse.device))


def to8b(img, debug_type=None):
    return (255 * np.clip(img, 0, 1)).astype(np.uint8)


def get_rays(H, W, K, c2w):
    """Pinhole rays (:41-55): dirs = [(i-cx)/fx, -(j-cy)/fy, -1] rotated by c2w; rays_d un-normalised.
    A CUDA `c2w` runs the star_get_rays kernel (bit-identical); host tensors (dataset preparation, as in the reference's
    datasets/*.py) stay on the host."""
    if torch.is_tensor(c2w) and c2w.is_cuda:
        return F_.get_rays(H, W, K, c2w)
    dev = c2w.device if torch.is_tensor(c2w) else None
    i, j = torch.meshgrid(torch.linspace(0, W - 1, W, device=dev), torch.linspace(0, H - 1, H, device=dev),
                          indexing="xy")
    dirs = torch.stack([(i - K[0][2]) / K[0][0], -(j - K[1][2]) / K[1][1], -torch.ones_like(i)], -1)
    rays_d = torch.sum(dirs[..., None, :] * c2w[:3, :3], -1)
    rays_o = c2w[:3, -1].expand(rays_d.shape).clone()
    return rays_o, rays_d


def get_rays_np(H, W, K, c2w):
    i, j = np.meshgrid(np.arange(W, dtype=np.float32), np.arange(H, dtype=np.float32), indexing="xy")
    dirs = np.stack([(i - K[0][2]) / K[0][0], -(j - K[1][2]) / K[1][1], -np.ones_like(i)], -1)
    rays_d = np.sum(dirs[..., np.newaxis, :] * c2w[:3, :3], -1)
    rays_o = np.broadcast_to(c2w[:3, -1], np.shape(rays_d))
    return rays_o, rays_d


def sample_pts(rays_o, rays_d, near, far, N_samples, perturb=0, lindisp=False, is_train=True, *, t_rand=None):
    """(:75-112) -> pts [R,N,3], z_vals [R,N]."""
    if is_train and perturb > 0.0:
        if t_rand is None:
            t_rand = torch.rand((rays_o.shape[0], N_samples), device=rays_o.device)
    else:
        t_rand = None
    return F_.sample_pts(rays_o, rays_d, near, far, N_samples, lindisp, t_rand)


def raw2alpha(raw, dists, act_fn=None):
    """(:301-303) 1 - exp(-softplus(raw) * dists); elementwise glue kept in torch for API parity only
    (the compositing kernels compute alpha internally)."""
    return 1.0 - torch.exp(-torch.nn.functional.softplus(raw) * dists)


def raw2outputs(raw_alpha, raw_rgb, z_vals, rays_d, raw_noise_std, white_bkgd, far_dist, *, noise=None):
    """(:307-379) single-field compositing -> NerfNetworkOutput."""
    if raw_noise_std > 0.0:
        if noise is None:
            noise = torch.randn(raw_alpha.shape, device=raw_alpha.device)
        raw_alpha = raw_alpha + noise * raw_noise_std
    rgb, disp, acc, depth, weights, dists = F_.CompositeSingle.apply(raw_alpha, raw_rgb, z_vals, rays_d,
                                                                     float(far_dist), bool(white_bkgd))
    return {"rgb": rgb, "disp": disp, "acc": acc, "weights": weights, "depth": depth, "dists": dists,
            "z_vals": z_vals}


def raw2outputs_star(raw_alpha_static, raw_rgb_static, raw_alpha_dynamic, raw_rgb_dynamic, z_vals, rays_d,
                     raw_noise_std=0, white_bkgd=False, far_dist=1e10, test=False, *, chunk=None):
    """(:383-576) static + V dynamic fields and the five regularisers (:612-715) -> StarNetworkOutput.
    `chunk`: ray-chunk length whose per-chunk means are summed (STaR.forward semantics, star__.py:84-112);
    None = one chunk."""
    if raw_noise_std > 0.0:
        raise NotImplementedError("the reference hard-codes raw_noise_std=0 for the multi-field path (star__.py:221)")
    R = raw_alpha_static.shape[0]
    o = F_.CompositeStar.apply(raw_alpha_static, raw_rgb_static, raw_alpha_dynamic, raw_rgb_dynamic, z_vals, rays_d,
                               float(far_dist), bool(white_bkgd), int(chunk or max(R, 1)), bool(test))
    d = dict(zip(F_.STAR_OUT_KEYS, o))
    regs = d.pop("regs")
    d["loss_alpha_entropy"], d["loss_dynamic_vs_static_reg"], d["loss_ray_reg"], d["loss_static_reg"], \
        d["loss_dynamic_reg"] = regs[0], regs[1], regs[2], regs[3], regs[4]
    if not test:
        d["rgb_dynamic_all"] = None
    return d


def sample_pdf(bins, weights, N_samples, det=False, *, u=None):
    """(:719-761) inverse-CDF sampling -> samples [R,N_samples]."""
    return F_.sample_pdf(bins, weights, N_samples, det=det, u=u)


def _coarse_to_fine(net_call, training, pts, z_vals, rays_o, rays_d, N_importance, u, z_samples=None):
    """Shared body of render_star_appinit / render_star_online / render_nerf (:115-149, :187-298)."""
    result = {}
    coarse = net_call(pts, z_vals, True)
    for k, v in coarse.items():
        result[f"{k}0"] = v
    if N_importance > 0:
        # the fine positions rays_o + rays_d * z (:137-139, :283-291) are never materialised: the MLP kernels form them
        # from (rays_o, rays_d, z_all) with the reference's roundings
        z_samples, z_all, z_std, _ = F_.hierarchical(z_vals, coarse["weights"], N_importance, det=not training,
                                                     rays_o=rays_o, rays_d=rays_d, u=u, z_samples=z_samples,
                                                     want_pts=False)
        fine = net_call((rays_o, rays_d, z_all), z_all, False)
        for k, v in fine.items():
            result[k] = v
        result["z_std"] = z_std
    return result


# Inference (no autograd graph wanted): the whole coarse -> fine render is ONE C-ABI call (star_render_forward: every
# kernel of the V + 1 coarse nets, compositing, inverse-CDF sampling / merge, the V + 1 fine nets and compositing queued
# without returning to Python).  Training keeps the staged path below, whose stages have backward twins.
FUSED_INFERENCE = True


def _fused(star_network, pts, viewdirs, z_vals, rays_o, rays_d, N_importance, pose, step, u, z_samples):
    if not FUSED_INFERENCE or torch.is_grad_enabled() or not hasattr(star_network, "static_coarse_nerf"):
        return None
    sc = star_network.static_coarse_nerf
    if star_network.training and sc.raw_noise_std > 0.0 and pose is None:
        return None                                    # density noise (:335): staged path
    Ni = N_importance if star_network.N_importance > 0 else 0
    if N_importance > 0 and star_network.N_importance <= 0:
        raise ValueError("N_importance should be positive")
    sf = star_network.static_fine_nerf if Ni > 0 else None
    dyn_c, dyn_f, pose12, scales = [], [], None, (None, None)
    if pose is not None:
        if pose.dim() not in (2, 3):
            raise NotImplementedError
        dyn_c = list(star_network.dynamic_coarse_nerfs)
        dyn_f = list(star_network.dynamic_fine_nerfs) if Ni > 0 else []
        if not dyn_c:
            return None
        pose12 = torch.stack([F_.pose_to_mat12(pose[i]) for i in range(len(dyn_c))])
        m = dyn_c[0]
        scales = (m.embedder.scale(step, viewdirs.device, pad_to=64), m.embedder_dirs.scale(step, viewdirs.device, pad_to=32))
    res = F_.render_forward((sc, sf), (dyn_c, dyn_f), sc._prec(), rays_o, rays_d, viewdirs, Ni, z_vals=z_vals, pts=pts,
                            pose12=pose12, det=not star_network.training, u=u, z_samples=z_samples,
                            white_bkgd=sc.white_bkgd, far_dist=star_network.far_dist, chunk=star_network.chunk,
                            test=not star_network.training, enc_scales=scales)
    return {k: v for k, v in res.items() if not k.startswith("_")}


def render_star_appinit(star_network, pts, viewdirs, z_vals, rays_o, rays_d, N_importance, *, u=None,
                        z_samples=None):
    """(:115-149)."""
    fused = _fused(star_network, pts, viewdirs, z_vals, rays_o, rays_d, N_importance, None, None, u, z_samples)
    if fused is not None:
        return fused

    def call(p, z, coarse):
        return star_network(p, viewdirs, z, rays_d, is_coarse=coarse)
    return _coarse_to_fine(call, star_network.training, pts, z_vals, rays_o, rays_d, N_importance, u, z_samples)


def render_star_online(star_network, pts, viewdirs, z_vals, rays_o, rays_d, N_importance, pose, step=None, *,
                       u=None, z_samples=None):
    """(:249-298)."""
    fused = _fused(star_network, pts, viewdirs, z_vals, rays_o, rays_d, N_importance, pose, step, u, z_samples)
    if fused is not None:
        return fused

    def call(p, z, coarse):
        return star_network(p, viewdirs, z, rays_d, pose, is_coarse=coarse, step=step)
    return _coarse_to_fine(call, star_network.training, pts, z_vals, rays_o, rays_d, N_importance, u, z_samples)


def render_nerf(nerf_coarse, nerf_fine, pts, viewdirs, z_vals, rays_o, rays_d, N_importance, far_dist, *, u=None,
                z_samples=None):
    """(:187-245) two bare NeRF modules."""
    def call(p, z, coarse):
        net = nerf_coarse if coarse else nerf_fine
        ra, rc = net(p, viewdirs, step=None)
        return raw2outputs(ra, rc, z, rays_d, net.raw_noise_std if net.training else 0, net.white_bkgd, far_dist)
    res = _coarse_to_fine(call, nerf_coarse.training, pts, z_vals, rays_o, rays_d, N_importance, u, z_samples)
    assert all(k in res for k in NERF_KEYS)
    return res

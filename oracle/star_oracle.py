"""ORACLE -- TEST INFRASTRUCTURE ONLY.  CPU restatement (torch fp32, eager ATen ops) of the
STaR / NeRF render hot path of burakcuhadar/3D-MOT-using-Neural-Radiance-Fields.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this file, and only as the checker / the CPU baseline.  The product path
(3d-mot-using-neural-radiance-fields_b200/) never imports it and has no CPU fallback.

Parity status
  * vanilla path (sample_pts, embedder, ResNet-FC NeRF MLP, raw2outputs, raw2outputs_star,
    regularisers, sample_pdf, render_star_appinit / render_star_online / render_nerf, 4x4 pose):
    PINNED against the unmodified reference executed in the build container
    (tools/make_golden.py -> tests/golden/*.npz, checked by tests/test_oracle_golden.py).
    The reference itself holds no golden vectors or asserting tests.
  * 7-vector (t, q) pose branch: the arithmetic lives in pypose, which is neither vendored nor
    version-pinned by the reference -> "parity unpinned".  Restated from pypose's published
    semantics (see se3_act / so3_act below) and cross-checked against the 4x4 branch.

Everything is written in a functional style over a flat {state_dict key: tensor} mapping that
uses the reference checkpoint key layout, so it accepts the state_dict
of either the reference STaR module or the B200 one.

All file:line citations are into /root/reference.
"""
import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

EPS = torch.finfo(torch.float32).eps  # utils/constants.py:3


# ----------------------------------------------------------------------------- a1 sample_pts
def sample_pts(rays_o, rays_d, near, far, N_samples, perturb=0, lindisp=False, is_train=True,
               t_rand=None):
    """models/rendering__.py:75-112.  `t_rand` injects the jitter noise (":105" draws torch.rand)."""
    near_c = near * torch.ones_like(rays_d[..., :1])
    far_c = far * torch.ones_like(rays_d[..., :1])
    t = torch.linspace(0.0, 1.0, steps=N_samples, device=rays_o.device)
    if lindisp:
        z = 1.0 / (1.0 / near_c * (1.0 - t) + 1.0 / far_c * t)
    else:
        z = near_c * (1.0 - t) + far_c * t
    z = z.expand([rays_o.shape[0], N_samples])
    if is_train and perturb > 0.0:
        mid = 0.5 * (z[..., 1:] + z[..., :-1])
        hi = torch.cat([mid, z[..., -1:]], -1)
        lo = torch.cat([z[..., :1], mid], -1)
        if t_rand is None:
            t_rand = torch.rand(z.shape, device=rays_o.device)
        z = lo + (hi - lo) * t_rand
    pts = rays_o[..., None, :] + rays_d[..., None, :] * z[..., :, None]
    return pts, z


def get_rays(H, W, K, c2w):
    """models/rendering__.py:41-55 (pinhole rays, rays_d not normalised)."""
    i, j = torch.meshgrid(torch.linspace(0, W - 1, W), torch.linspace(0, H - 1, H), indexing="xy")
    dirs = torch.stack([(i - K[0][2]) / K[0][0], -(j - K[1][2]) / K[1][1], -torch.ones_like(i)], -1)
    rays_d = torch.sum(dirs[..., None, :] * c2w[:3, :3], -1)
    rays_o = c2w[:3, -1].expand(rays_d.shape).clone()
    return rays_o, rays_d


# ----------------------------------------------------------------------------- a3 embedder
def embed(x, L, step=None, end_barf=-1):
    """models/embedder.py:81-112: [x, sin(2^0 x), cos(2^0 x), sin(2^1 x), ...]; optional BARF mask
    with the reference quirk that element j of the 6L encoded dims is scaled by w[j mod L] (:32)."""
    freqs = 2.0 ** torch.linspace(0.0, L - 1, steps=L)
    parts = [x]
    for f in freqs:
        parts.append(torch.sin(x * f))
        parts.append(torch.cos(x * f))
    enc = torch.cat(parts, -1)
    if step is None or end_barf == -1:
        return enc
    d = x.shape[-1]
    alpha = (step - 0) / (end_barf - 0) * L
    k = torch.arange(L, dtype=torch.float32, device=x.device)
    w = (1 - (alpha - k).clamp_(min=0, max=1).mul_(math.pi).cos_()) / 2
    tail = enc[:, d:]
    masked = (tail.contiguous().view(-1, L) * w).view(*tail.shape)
    return torch.cat([enc[:, :d], masked], 1)


# ----------------------------------------------------------------------------- a4 NeRF MLP
def n_blocks_of(params: Dict[str, torch.Tensor], prefix: str) -> int:
    n = 0
    while f"{prefix}pts_net.blocks.{n}.fc_0.weight" in params:
        n += 1
    return n


def nerf_mlp(params, prefix, pts, viewdirs, L_xyz=10, L_dir=4, step=None, end_barf=-1, emulate_bf16=False):
    """models/nerf.py:112-179 + models/resnet.py:51-59,103-110.  Returns RAW (raw_alpha [R,S],
    raw_rgb [R,S,3]) -- the only mode STaR uses (z_vals is None, nerf.py:178).
    emulate_bf16: model of the tensor-core tier -- every GEMM operand (weights, and the activations /
    encodings fed to a GEMM) rounded to bf16, fp32 accumulation, fp32 biases, fp32 residual stream,
    the two heads (alpha_linear, rgb_linear) in fp32 on un-rounded activations."""
    GEMMS = ("pts_net", "feature_linear", "views_linears")

    def mode_of(name):
        # emulate_bf16: False | True / "bf16" | "fp16" (one mode for every GEMM), or a dict keyed by the layer's
        # short name ("lin_in", "fc_0", "fc_1", "lin_out", "feature_linear", "views_linears"; missing = exact) --
        # the per-layer form is what tools/error_budget.py uses.  Modes "bf16x2" / "fp16x2" model a split operand
        # (hi + lo, the lo * lo product dropped: three MMAs per K-step).
        if isinstance(emulate_bf16, dict):
            parts = name.split(".")
            return emulate_bf16.get(parts[-2] if parts[-1].isdigit() else parts[-1], False)
        return "bf16" if emulate_bf16 is True else emulate_bf16

    def q(t, mode):
        return t.half().to(t.dtype) if mode.startswith("fp16") else t.bfloat16().to(t.dtype)

    def lin(name, h):
        w, b = params[f"{prefix}{name}.weight"], params[f"{prefix}{name}.bias"]
        mode = mode_of(name) if name.startswith(GEMMS) else False
        if not mode:
            return F.linear(h, w, b)
        h1, w1 = q(h, mode), q(w, mode)
        if mode.endswith("x2"):
            h2, w2 = q(h - h1, mode), q(w - w1, mode)
            return F.linear(h1, w1, b) + F.linear(h2, w1) + F.linear(h1, w2)
        return F.linear(h1, w1, b)

    R, S = pts.shape[0], pts.shape[1]
    p = pts.reshape(-1, 3)
    d = viewdirs[:, None].expand(pts.shape).reshape(-1, 3)          # nerf.py:133-136
    e_p = embed(p, L_xyz, step, end_barf)
    e_d = embed(d, L_dir, step, end_barf)
    x = lin("pts_net.lin_in", e_p)                                  # resnet.py:104
    for b in range(n_blocks_of(params, prefix)):                    # resnet.py:51-59 (pre-act)
        net = lin(f"pts_net.blocks.{b}.fc_0", F.relu(x))
        x = x + lin(f"pts_net.blocks.{b}.fc_1", F.relu(net))
    h = lin("pts_net.lin_out", F.relu(x))                           # resnet.py:109
    raw_alpha = lin("alpha_linear", h)                              # nerf.py:151
    feat = lin("feature_linear", h)                                 # nerf.py:152
    h2 = F.relu(lin("views_linears.0", torch.cat([feat, e_d], -1)))  # nerf.py:153-157
    raw_rgb = lin("rgb_linear", h2)                                 # nerf.py:159
    return raw_alpha.reshape(R, S), raw_rgb.reshape(R, S, 3)


# ----------------------------------------------------------------------------- a5/a6 compositing
def raw2alpha(raw, dists):
    """models/rendering__.py:301-303 (softplus density, act_fn ignored)."""
    return 1.0 - torch.exp(-F.softplus(raw) * dists)


def _dists(z_vals, rays_d, far_dist):
    d = z_vals[..., 1:] - z_vals[..., :-1]
    d = torch.cat([d, torch.tensor([far_dist]).expand(d[..., :1].shape)], -1)
    return d * torch.norm(rays_d[..., None, :], dim=-1)


def _excl_cumprod(alpha):
    one = torch.ones(alpha.shape[:-1] + (1,))
    return torch.cumprod(torch.cat([one, 1.0 - alpha + 1e-10], -1), -1)[..., :-1]


def raw2outputs(raw_alpha, raw_rgb, z_vals, rays_d, raw_noise_std, white_bkgd, far_dist, noise=None):
    """models/rendering__.py:307-379.  `noise` injects the (std-scaled) density noise (":335")."""
    dists = _dists(z_vals, rays_d, far_dist)
    rgb = torch.sigmoid(raw_rgb)
    nz = 0.0
    if raw_noise_std > 0.0:
        nz = (torch.randn(raw_alpha.shape) if noise is None else noise) * raw_noise_std
    alpha = raw2alpha(raw_alpha + nz, dists)
    weights = alpha * _excl_cumprod(alpha)
    rgb_map = torch.sum(weights[..., None] * rgb, -2)
    depth = torch.sum(weights * z_vals, -1)
    wsum = torch.sum(weights, -1)
    wsum = torch.where(wsum >= 0, wsum, 1e-7)
    disp = 1.0 / torch.max(1e-10 * torch.ones_like(depth), depth / wsum)
    acc = torch.sum(weights, -1)
    if white_bkgd:
        rgb_map = rgb_map + (1.0 - acc[..., None])
    return {"rgb": rgb_map, "disp": disp, "acc": acc, "weights": weights, "depth": depth,
            "dists": dists, "z_vals": z_vals}


# ----------------------------------------------------------------------------- a8 regularisers
def alpha_entropy(alpha_s, alpha_d):
    """models/rendering__.py:612-631."""
    V = alpha_d.shape[1]
    cs = alpha_s.clamp(min=EPS, max=1 - EPS)
    cd = alpha_d.clamp(min=EPS, max=1 - EPS)
    e = -torch.mean(alpha_s * torch.log(cs) + (1 - alpha_s) * torch.log1p(-cs)) / (V + 1)
    e = e + -torch.mean(alpha_d * torch.log(cd) + (1 - alpha_d) * torch.log1p(-cd), (0, 2)).sum() / (V + 1)
    return e


def dynamic_vs_static_reg(alpha_s, alpha_d):
    """models/rendering__.py:634-651 (sigma arguments are unused by the live code)."""
    tot = alpha_s + alpha_d.sum(dim=1)
    ps = (alpha_s / tot.clamp(min=EPS)).clamp(min=EPS)
    pd = (alpha_d / tot.clamp(min=EPS)[:, None, :]).clamp(min=EPS)
    return -torch.mean(tot * (ps * ps.log() + torch.sum(pd * pd.log(), dim=1)))


def ray_reg(sigma_d, sigma_sum):
    """models/rendering__.py:682-695."""
    V = sigma_d.shape[1]
    n = sigma_d / sigma_sum.clamp(min=EPS)[:, None, :]
    return torch.mean(torch.max(n, dim=-1)[0] ** 2.0, dim=0).sum() / V


def static_reg(sigma_s, alpha_s):
    """models/rendering__.py:698-711."""
    c = alpha_s.clamp(min=EPS, max=1 - EPS)
    ssum = torch.sum(sigma_s, dim=-1, keepdim=True)
    mask = torch.where(ssum < 0.1, 0.0, 1.0)
    p = c / torch.sum(c, dim=-1, keepdim=True)
    return torch.mean(mask * -torch.mean(p * torch.log(p), dim=-1, keepdim=True))


def dynamic_reg(sigma_d):
    """models/rendering__.py:714-715."""
    return sigma_d.mean()


# ----------------------------------------------------------------------------- a7 multi-field
def raw2outputs_star(raw_alpha_s, raw_rgb_s, raw_alpha_d, raw_rgb_d, z_vals, rays_d,
                     white_bkgd=False, far_dist=1e10, test=False):
    """models/rendering__.py:383-576 with raw_noise_std = 0 (star__.py:221 hard-codes 0)."""
    dists = _dists(z_vals, rays_d, far_dist)
    c_s = torch.sigmoid(raw_rgb_s)
    c_d = torch.sigmoid(raw_rgb_d)
    a_s = raw2alpha(raw_alpha_s, dists)
    a_d = raw2alpha(raw_alpha_d, dists[:, None, :])
    a_t = raw2alpha(raw_alpha_s + raw_alpha_d.sum(dim=1), dists)       # :416-418 softplus of summed raw
    T_s, T_d, T = _excl_cumprod(a_s), _excl_cumprod(a_d), _excl_cumprod(a_t)
    mix = a_s[..., None] * c_s + torch.sum(a_d[..., None] * c_d, dim=1)
    rgb = torch.sum(T[..., None] * mix, dim=-2)                         # :456-463
    rgb_static = torch.sum(T_s[..., None] * a_s[..., None] * c_s, dim=-2)
    rgb_dynamic = torch.sum(T_d[..., None] * a_d[..., None] * c_d, dim=-2)
    depth_dynamic = torch.sum(T_d * a_d * z_vals[:, None, :], -1)
    depth_static = torch.sum(T_s * a_s * z_vals, -1)
    weights = T * a_t                                                   # :503
    depth = torch.sum(weights * z_vals, -1)
    wsum = torch.sum(weights, -1)
    wsum = torch.where(wsum >= 0, wsum, EPS)                            # :509-511
    disp = 1.0 / torch.max(1e-10 * torch.ones_like(depth), depth / wsum)
    acc = torch.sum(weights, -1)
    if white_bkgd:
        rgb = rgb + (1.0 - acc[..., None])
    sig_s = F.softplus(raw_alpha_s)
    sig_d = F.softplus(raw_alpha_d)
    sig_sum = sig_s + sig_d.sum(dim=1)
    out = {
        "rgb": rgb, "disp": disp, "acc": acc, "weights": weights, "depth": depth,
        "rgb_static": rgb_static, "rgb_dynamic": rgb_dynamic, "depth_static": depth_static,
        "depth_dynamic": depth_dynamic, "dynamic_transmittance": T_d[:, :, -1],
        "loss_alpha_entropy": alpha_entropy(a_s, a_d),
        "loss_dynamic_vs_static_reg": dynamic_vs_static_reg(a_s, a_d),
        "loss_ray_reg": ray_reg(sig_d, sig_sum),
        "loss_static_reg": static_reg(sig_s, a_s),
        "loss_dynamic_reg": dynamic_reg(sig_d),
        "rgb_dynamic_all": None,
    }
    if test:                                                            # :534-555
        a_all = raw2alpha(raw_alpha_d.sum(dim=1), dists)
        T_all = _excl_cumprod(a_all)
        out["rgb_dynamic_all"] = torch.sum(T_all[..., None] * torch.sum(a_d[..., None] * c_d, dim=1), dim=-2)
    return out


# ----------------------------------------------------------------------------- a9 sample_pdf
def pdf_to_cdf(weights, exact_sum=False):
    """models/rendering__.py:722-734.  exact_sum=False follows the reference op for op
    (torch.sum cascade order, machine dependent: AVX2 vs AVX512 give different last bits).
    exact_sum=True is the *defined arithmetic* the CUDA kernel implements: the normaliser is the
    exactly-rounded fp32 sum (fp64 accumulate); torch's CPU cumsum already is fp64-accumulate /
    fp32-round-per-element, which the kernel reproduces bit for bit."""
    w = weights + 1e-5
    if exact_sum:
        s = w.double().sum(-1, keepdim=True).float()
    else:
        s = torch.sum(w, -1, keepdim=True)
    pdf = w / s
    cdf = torch.cumsum(pdf, -1)
    return torch.cat([torch.zeros_like(cdf[..., :1]), cdf], -1)


def invert_cdf(bins, cdf, u):
    """models/rendering__.py:744-761 -> samples, inds, below, above (int64)."""
    u = u.contiguous()
    inds = torch.searchsorted(cdf, u, right=True)
    below = torch.max(torch.zeros_like(inds - 1), inds - 1)
    above = torch.min((cdf.shape[-1] - 1) * torch.ones_like(inds), inds)
    g = torch.stack([below, above], -1)
    shp = [g.shape[0], g.shape[1], cdf.shape[-1]]
    cdf_g = torch.gather(cdf.unsqueeze(1).expand(shp), 2, g)
    bins_g = torch.gather(bins.unsqueeze(1).expand(shp), 2, g)
    denom = cdf_g[..., 1] - cdf_g[..., 0]
    denom = torch.where(denom < 1e-5, torch.ones_like(denom), denom)
    t = (u - cdf_g[..., 0]) / denom
    samples = bins_g[..., 0] + t * (bins_g[..., 1] - bins_g[..., 0])
    return samples, inds, below, above


def sample_pdf(bins, weights, N_samples, det=False, u=None, exact_sum=False, return_details=False):
    """models/rendering__.py:719-761.  `u` injects the uniform draws (":741" uses torch.rand)."""
    cdf = pdf_to_cdf(weights, exact_sum)
    if u is None:
        if det:
            u = torch.linspace(0.0, 1.0, steps=N_samples).expand(list(cdf.shape[:-1]) + [N_samples])
        else:
            u = torch.rand(list(cdf.shape[:-1]) + [N_samples])
    samples, inds, below, above = invert_cdf(bins, cdf, u)
    if return_details:
        return samples, {"cdf": cdf, "u": u, "inds": inds, "below": below, "above": above}
    return samples


# ----------------------------------------------------------------------------- pose (K2)
def quat_rotate(q, p):
    """pypose SO3 Act (third-party, unpinned): uv = 2 q_v x p; out = p + q_w uv + q_v x uv."""
    qv, qw = q[..., :3], q[..., 3:4]
    uv = 2.0 * torch.cross(qv.expand_as(p), p, dim=-1)
    return p + qw * uv + torch.cross(qv.expand_as(p), uv, dim=-1)


def quat_to_matrix(q):
    x, y, z, w = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    K = torch.zeros(q.shape[:-1] + (3, 3), dtype=q.dtype)
    K[..., 0, 1], K[..., 0, 2] = -z, y
    K[..., 1, 0], K[..., 1, 2] = z, -x
    K[..., 2, 0], K[..., 2, 1] = -y, x
    return torch.eye(3, dtype=q.dtype) + 2.0 * w[..., None, None] * K + 2.0 * (K @ K)


class _SE3Act(torch.autograd.Function):
    """pypose SE3_Act: gradient for X is the LEFT tangent gradient padded with a zero:
    [sum g, sum out x g, 0]  (out = transformed point, g = dL/dout)."""

    @staticmethod
    def forward(ctx, X, p):
        out = X[..., :3] + quat_rotate(X[..., 3:7], p)
        ctx.save_for_backward(X, out)
        return out

    @staticmethod
    def backward(ctx, g):
        X, out = ctx.saved_tensors
        gx = torch.cat([g, torch.cross(out, g, dim=-1), torch.zeros_like(g[..., :1])], -1)
        while gx.dim() > X.dim():
            gx = gx.sum(0)
        return gx, g @ quat_to_matrix(X[..., 3:7])


class _SO3Act(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, p):
        out = quat_rotate(q, p)
        ctx.save_for_backward(q, out)
        return out

    @staticmethod
    def backward(ctx, g):
        q, out = ctx.saved_tensors
        gq = torch.cat([torch.cross(out, g, dim=-1), torch.zeros_like(g[..., :1])], -1)
        while gq.dim() > q.dim():
            gq = gq.sum(0)
        return gq, g @ quat_to_matrix(q)


def se3_act(X, p):
    return _SE3Act.apply(X, p)


def so3_act(q, p):
    return _SO3Act.apply(q, p)


# ----------------------------------------------------------------------------- a11 STaR.forward
class StarConfig:
    """The fields of `args` the path reads (star__.py:28-53, nerf.py:41-101)."""

    def __init__(self, num_vehicles=0, N_importance=128, chunk=8192, far_dist=1e10, white_bkgd=False,
                 raw_noise_std=0.0, multires=10, multires_views=4, end_barf=-1, emulate_bf16=False):
        self.num_vehicles = num_vehicles
        self.N_importance = N_importance
        self.chunk = chunk
        self.far_dist = far_dist
        self.white_bkgd = white_bkgd
        self.raw_noise_std = raw_noise_std
        self.multires = multires
        self.multires_views = multires_views
        self.end_barf = end_barf
        self.emulate_bf16 = emulate_bf16       # model of the bf16 tensor-core MLP tier (see nerf_mlp)


def _star_chunk(params, cfg, pts, viewdirs, z_vals, rays_d, pose, is_coarse, step, training):
    """models/star__.py:119-225."""
    tag = "coarse" if is_coarse else "fine"
    if not is_coarse and cfg.N_importance <= 0:
        raise ValueError("N_importance should be positive")
    V = cfg.num_vehicles
    R, S = pts.shape[0], pts.shape[1]
    kw = dict(L_xyz=cfg.multires, L_dir=cfg.multires_views, end_barf=cfg.end_barf,
              emulate_bf16=getattr(cfg, "emulate_bf16", False))
    ra_s, rc_s = nerf_mlp(params, f"static_{tag}_nerf.", pts, viewdirs, step=None, **kw)   # :144
    if pose is None:
        return raw2outputs(ra_s, rc_s, z_vals, rays_d, cfg.raw_noise_std if training else 0,
                           cfg.white_bkgd, cfg.far_dist)
    if pose.dim() == 3:                                                                   # :160-180
        ph = torch.cat([pts, torch.ones((R, S, 1))], dim=-1).reshape(-1, 4)
        pd = torch.einsum("vij,nj->vni", pose, ph).reshape(V, R, S, 4)[..., :3]
        vd = torch.einsum("vij,nj->vni", pose[:, :3, :3], viewdirs)
    elif pose.dim() == 2:                                                                 # :182-199
        flat = pts.reshape(-1, 3)
        pd = torch.stack([se3_act(pose[i], flat).reshape(R, S, 3) for i in range(V)], 0)
        vd = torch.stack([so3_act(pose[i, 3:], viewdirs) for i in range(V)], 0)
    else:
        raise NotImplementedError
    ra_d, rc_d = [], []
    for i in range(V):                                                                    # :201-210
        a, c = nerf_mlp(params, f"dynamic_{tag}_nerfs.{i}.", pd[i], vd[i], step=step, **kw)
        ra_d.append(a)
        rc_d.append(c)
    ra_d = torch.stack(ra_d, 1) if V else torch.zeros((R, 0, S))
    rc_d = torch.stack(rc_d, 1) if V else torch.zeros((R, 0, S, 3))
    return raw2outputs_star(ra_s, rc_s, ra_d, rc_d, z_vals, rays_d, cfg.white_bkgd, cfg.far_dist,
                            test=not training)


def star_forward(params, cfg, pts, viewdirs, z_vals, rays_d, pose=None, is_coarse=True, step=None,
                 training=False):
    """models/star__.py:68-116: ray chunks of cfg.chunk; per-ray tensors concatenated, 0-dim
    (regulariser) outputs SUMMED over chunks (:111-112), None passed through."""
    acc = {}
    for i in range(0, pts.shape[0], cfg.chunk):
        j = min(pts.shape[0], i + cfg.chunk)
        part = _star_chunk(params, cfg, pts[i:j], viewdirs[i:j], z_vals[i:j], rays_d[i:j], pose,
                           is_coarse, step, training)
        for k, v in part.items():
            acc.setdefault(k, []).append(v)
    out = {}
    for k, v in acc.items():
        if v[0] is None:
            out[k] = None
        elif v[0].dim() == 0:
            out[k] = sum(v)
        else:
            out[k] = torch.cat(v, 0)
    return out


# ----------------------------------------------------------------------------- a10 orchestration
def _hierarchical(z_vals, weights, N_importance, det, u, exact_sum, z_samples=None):
    """z_samples: "teacher forcing" -- use these fine samples instead of inverting the coarse pdf.
    sample_pdf is ill-conditioned where the pdf is small (t = (u - cdf_lo) / denom with denom down to
    1e-5 turns a 1e-7 rounding difference of the coarse weights into a 1e-3 shift of a sample), so
    stage-wise parity of the fine pass is checked on identical sample positions."""
    if z_samples is not None:
        zs = z_samples.detach()
        z_all, _ = torch.sort(torch.cat([z_vals, zs], -1), -1)
        return zs, z_all
    mid = 0.5 * (z_vals[..., 1:] + z_vals[..., :-1])
    zs = sample_pdf(mid, weights[..., 1:-1], N_importance, det=det, u=u, exact_sum=exact_sum).detach()
    z_all, _ = torch.sort(torch.cat([z_vals, zs], -1), -1)
    return zs, z_all


def render_star(params, cfg, pts, viewdirs, z_vals, rays_o, rays_d, N_importance, pose=None, step=None,
                training=False, u=None, exact_sum=False, z_samples=None):
    """models/rendering__.py:115-149 (pose None, render_star_appinit) and :249-298
    (render_star_online).  `u` injects sample_pdf's uniform draws in training mode."""
    res = {}
    coarse = star_forward(params, cfg, pts, viewdirs, z_vals, rays_d, pose, True, step, training)
    for k, v in coarse.items():
        res[f"{k}0"] = v
    if N_importance > 0:
        zs, z_all = _hierarchical(z_vals, coarse["weights"], N_importance, not training, u, exact_sum,
                                  z_samples)
        pts_f = rays_o[..., None, :] + rays_d[..., None, :] * z_all[..., :, None]
        fine = star_forward(params, cfg, pts_f, viewdirs, z_all, rays_d, pose, False, step, training)
        res.update(fine)
        res["z_std"] = torch.std(zs, dim=-1, unbiased=False)
    return res


def render_nerf(params, prefix_coarse, prefix_fine, cfg, pts, viewdirs, z_vals, rays_o, rays_d,
                N_importance, far_dist, training=False, u=None, exact_sum=False, z_samples=None):
    """models/rendering__.py:187-245 (two bare NeRF modules)."""
    kw = dict(L_xyz=cfg.multires, L_dir=cfg.multires_views, end_barf=cfg.end_barf,
              emulate_bf16=getattr(cfg, "emulate_bf16", False))
    std = cfg.raw_noise_std if training else 0
    a, c = nerf_mlp(params, prefix_coarse, pts, viewdirs, **kw)
    coarse = raw2outputs(a, c, z_vals, rays_d, std, cfg.white_bkgd, far_dist)
    zs, z_all = _hierarchical(z_vals, coarse["weights"], N_importance, not training, u, exact_sum, z_samples)
    pts_f = rays_o[..., None, :] + rays_d[..., None, :] * z_all[..., :, None]
    a, c = nerf_mlp(params, prefix_fine, pts_f, viewdirs, **kw)
    fine = raw2outputs(a, c, z_all, rays_d, std, cfg.white_bkgd, far_dist)
    res = dict(fine)
    for k, v in coarse.items():
        res[f"{k}0"] = v
    res["z_std"] = torch.std(zs, dim=-1, unbiased=False)
    return res


# ----------------------------------------------------------------------------- synthetic inputs
def init_star_params(num_vehicles, N_importance=1, W=256, D=8, L_xyz=10, L_dir=4, seed=0,
                     redraw_fc1=True, bias_std=0.0):
    """Random-init weights with the reference's initialisers (resnet.py:33-37,80-86;
    nerf.py:104-109) in the reference state_dict key layout.  fc_1.weight is zero in the
    reference init (resnet.py:37), which makes half the GEMMs multiply zeros, so synthetic
    benchmarks / parity tests re-draw it (hard part 6)."""
    g = torch.Generator().manual_seed(seed)
    in_xyz, in_dir = 3 + 6 * L_xyz, 3 + 6 * L_dir

    def kaiming(o, i):
        return torch.randn(o, i, generator=g) * math.sqrt(2.0 / i)

    def default_w(o, i):
        b = 1.0 / math.sqrt(i)
        return (torch.rand(o, i, generator=g) * 2 - 1) * b

    def default_b(o, i):
        b = 1.0 / math.sqrt(i)
        return (torch.rand(o, generator=g) * 2 - 1) * b

    def net(prefix, depth, out):
        def put(name, w, b):
            if bias_std > 0:
                b = b + torch.randn(b.shape, generator=g) * bias_std
            out[f"{prefix}{name}.weight"], out[f"{prefix}{name}.bias"] = w, b
        put("pts_net.lin_in", kaiming(W, in_xyz), torch.zeros(W))
        put("pts_net.lin_out", torch.randn(W, W, generator=g) * math.sqrt(1.0 / W), torch.zeros(W))
        for b in range(depth // 2):
            put(f"pts_net.blocks.{b}.fc_0", kaiming(W, W), torch.zeros(W))
            put(f"pts_net.blocks.{b}.fc_1", kaiming(W, W) if redraw_fc1 else torch.zeros(W, W), torch.zeros(W))
        put("views_linears.0", kaiming(W // 2, W + in_dir), torch.zeros(W // 2))
        put("feature_linear", default_w(W, W), default_b(W, W))
        put("alpha_linear", kaiming(1, W), torch.zeros(1))
        a = math.sqrt(6.0 / (W // 2 + 3))
        put("rgb_linear", (torch.rand(3, W // 2, generator=g) * 2 - 1) * a, default_b(3, W // 2))

    out = {}
    tags = ["coarse", "fine"] if N_importance > 0 else ["coarse"]
    for t in tags:
        net(f"static_{t}_nerf.", D, out)
    for t in tags:
        for v in range(num_vehicles):
            net(f"dynamic_{t}_nerfs.{v}.", D // 2, out)
    return out


def lego_rays(H, W, theta=30.0, phi=-30.0, radius=4.0):
    """C1/C2 synthetic camera (datasets/lego.py:33-38 pose_spherical, camera_angle_x 0.6911112)."""
    focal = 0.5 * W / math.tan(0.5 * 0.6911112)
    K = torch.tensor([[focal, 0, 0.5 * W], [0, focal, 0.5 * H], [0, 0, 1.0]])
    t = torch.eye(4)
    t[2, 3] = radius
    ph = phi / 180.0 * math.pi
    rp = torch.tensor([[1, 0, 0, 0], [0, math.cos(ph), -math.sin(ph), 0],
                       [0, math.sin(ph), math.cos(ph), 0], [0, 0, 0, 1.0]])
    th = theta / 180.0 * math.pi
    rt = torch.tensor([[math.cos(th), 0, -math.sin(th), 0], [0, 1, 0, 0],
                       [math.sin(th), 0, math.cos(th), 0], [0, 0, 0, 1.0]])
    c2w = rt @ rp @ t
    c2w = torch.tensor([[-1.0, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]]) @ c2w
    ro, rd = get_rays(H, W, K, c2w[:3, :4])
    return ro.reshape(-1, 3), rd.reshape(-1, 3)


def carla_rays(R, seed=0, n_views=50, H=400, W=400):
    """C3/C4 CARLA-shaped rays: cameras on a ring r=0.25, h=0.06 looking at
    the origin, fov 90 deg, R random pixels over n_views views; rays_d un-normalised."""
    g = torch.Generator().manual_seed(seed)
    view = torch.randint(0, n_views, (R,), generator=g)
    px = torch.randint(0, W, (R,), generator=g).float()
    py = torch.randint(0, H, (R,), generator=g).float()
    ang = view.float() / n_views * 2 * math.pi
    eye = torch.stack([0.25 * torch.cos(ang), 0.25 * torch.sin(ang), torch.full_like(ang, 0.06)], -1)
    fwd = F.normalize(-eye, dim=-1)
    up = torch.tensor([0.0, 0.0, 1.0]).expand_as(fwd)
    right = F.normalize(torch.cross(fwd, up, dim=-1), dim=-1)
    upv = torch.cross(right, fwd, dim=-1)
    focal = W / 2.0
    dx, dy = (px - 0.5 * W) / focal, -(py - 0.5 * H) / focal
    rays_d = dx[:, None] * right + dy[:, None] * upv + fwd
    return eye.contiguous(), rays_d.contiguous()


def random_poses7(V, seed=3):
    g = torch.Generator().manual_seed(seed)
    t = (torch.rand(V, 3, generator=g) * 2 - 1) * 0.1
    q = torch.cat([torch.randn(V, 3, generator=g) * 0.05, torch.ones(V, 1)], -1)
    q = q / q.norm(dim=-1, keepdim=True)
    return torch.cat([t, q], -1)


def pose7_to_matrix(p7):
    M = torch.eye(4).repeat(p7.shape[0], 1, 1)
    M[:, :3, :3] = quat_to_matrix(p7[:, 3:7])
    M[:, :3, 3] = p7[:, :3]
    return M

"""ORACLE -- TEST INFRASTRUCTURE ONLY.  CPU restatement (torch fp32, eager ATen ops) of the
mip-NeRF / integrated-positional-encoding variant of the STaR render path:
models/star_mipnerf.py:99-357, models/rendering_starmip.py:32-175, models/mipnerf.py:53-100.

PARITY UNPINNED.  The arithmetic of this variant lives in `nerfstudio` (and `pypose` for the pose), which
the reference neither vendors nor version-pins (absent from environment.yaml:1-21) and which is not
installed here, so the reference cannot be executed for this path and it holds no golden vectors or
asserting tests.  The functions below restate nerfstudio's published algorithms (recalled from public
nerfstudio 0.3 - 1.0 sources; each cites the upstream symbol) driven exactly the way the reference call
sites drive them; the regulariser quirks that follow from the trailing singleton dimension of the mip
tensors are restated from the reference's own functions (models/rendering__.py:612-715), which ARE pinned
(oracle/star_oracle.py).  Invariants that tie this file to pinned code are checked in
tests/test_mip_oracle.py.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this file.
All file:line citations are into /root/reference.
"""
import math

import torch
import torch.nn.functional as F

from . import star_oracle as so

EPS = so.EPS
N_FREQ_XYZ, MAX_EXP_XYZ = 24, 24.0     # models/mipnerf.py:58-64
N_FREQ_DIR, MAX_EXP_DIR = 4, 4.0       # models/mipnerf.py:65-71
D_XYZ = 3 * N_FREQ_XYZ * 2 + 3         # 147
D_DIR = 3 * N_FREQ_DIR * 2 + 3         # 27
W_BASE, W_HEAD, N_BASE, SKIP = 256, 128, 8, 4   # nerfstudio NeRFField defaults (mipnerf.py:73-78)
HIST_PAD = 0.01                        # nerfstudio PDFSampler(histogram_padding=0.01)


# ----------------------------------------------------------------------------- samplers
def uniform_bins(R, Nc, training=False, t_rand=None):
    """nerfstudio SpacedSampler.generate_ray_samples with the identity spacing (UniformSampler),
    driven by star_mipnerf.py:75-77,271.  Returns the Nc+1 bin edges in [0,1] ("spacing") per ray;
    `t_rand` [R,Nc+1] injects the stratified jitter of training mode."""
    bins = torch.linspace(0.0, 1.0, Nc + 1)[None, :]
    if training:
        if t_rand is None:
            t_rand = torch.rand(R, Nc + 1)
        centers = (bins[..., 1:] + bins[..., :-1]) / 2.0
        upper = torch.cat([centers, bins[..., -1:]], -1)
        lower = torch.cat([bins[..., :1], centers], -1)
        bins = lower + (upper - lower) * t_rand
    return bins.expand(R, Nc + 1)


def spacing_to_euclidean(x, near, far):
    """SpacedSampler: spacing_fn_inv(x * s_far + (1 - x) * s_near), identity spacing_fn;
    near/far from NearFarCollider (star_mipnerf.py:83-86,268-269)."""
    return x * far + (1 - x) * near


def pdf_sample(spacing_bins, weights, Ni, training=False, u_rand=None, return_details=False, exact_sum=False):
    """nerfstudio PDFSampler.generate_ray_samples(include_original=False, histogram_padding=0.01,
    train_stratified=True, single_jitter=False), called at star_mipnerf.py:286-288,334-336.
    spacing_bins [R,Nc+1] (existing edges), weights [R,Nc] -> new edges [R,Ni+1] (detached).
    `u_rand` [R,Ni+1] injects the training-mode jitter.  exact_sum: the *defined arithmetic* of the CUDA kernel
    (normaliser = exactly rounded fp32 sum; torch.sum's order is machine dependent, see star_oracle.pdf_to_cdf)."""
    nb = Ni + 1
    w = weights + HIST_PAD
    wsum = w.double().sum(-1, keepdim=True).float() if exact_sum else torch.sum(w, dim=-1, keepdim=True)
    padding = torch.relu(1e-5 - wsum)
    w = w + padding / w.shape[-1]
    wsum = wsum + padding
    pdf = w / wsum
    cdf = torch.min(torch.ones_like(pdf), torch.cumsum(pdf, dim=-1))
    cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], dim=-1)
    u = torch.linspace(0.0, 1.0 - (1.0 / nb), steps=nb)
    if training:
        if u_rand is None:
            u_rand = torch.rand(cdf.shape[0], nb)
        u = u.expand(cdf.shape[0], nb) + u_rand / nb
    else:
        u = (u + 1.0 / (2 * nb)).expand(cdf.shape[0], nb)
    u = u.contiguous()
    inds = torch.searchsorted(cdf, u, side="right")
    below = torch.clamp(inds - 1, 0, spacing_bins.shape[-1] - 1)
    above = torch.clamp(inds, 0, spacing_bins.shape[-1] - 1)
    cdf_g0, bins_g0 = torch.gather(cdf, -1, below), torch.gather(spacing_bins, -1, below)
    cdf_g1, bins_g1 = torch.gather(cdf, -1, above), torch.gather(spacing_bins, -1, above)
    t = torch.clip(torch.nan_to_num((u - cdf_g0) / (cdf_g1 - cdf_g0), 0), 0, 1)
    bins = (bins_g0 + t * (bins_g1 - bins_g0)).detach()
    if return_details:
        return bins, {"cdf": cdf, "u": u, "inds": inds, "below": below, "above": above}
    return bins


# ----------------------------------------------------------------------------- IPE
def frustum_gaussian(origins, directions, starts, ends, radius):
    """nerfstudio Frustums.get_gaussian_blob -> utils.math.conical_frustum_to_gaussian ->
    compute_3d_gaussian; only the DIAGONAL of the covariance is returned (it is all NeRFEncoding reads).
    origins/directions [R,3], starts/ends [R,S], radius scalar -> mean [R,S,3], diag [R,S,3]."""
    mu = (starts + ends) / 2.0
    hw = (ends - starts) / 2.0
    t_mean = mu + (2.0 * mu * hw ** 2.0) / (3.0 * mu ** 2.0 + hw ** 2.0)
    means = origins[:, None, :] + directions[:, None, :] * t_mean[..., None]
    dir_var = (hw ** 2) / 3 - (4 / 15) * ((hw ** 4 * (12 * mu ** 2 - hw ** 2)) / (3 * mu ** 2 + hw ** 2) ** 2)
    rad_var = radius ** 2 * ((mu ** 2) / 4 + (5 / 12) * hw ** 2 - 4 / 15 * (hw ** 4) / (3 * mu ** 2 + hw ** 2))
    d = directions[:, None, :]
    mag = torch.clamp(torch.sum(d ** 2, dim=-1, keepdim=True), min=1e-10)
    diag = dir_var[..., None] * (d * d) + rad_var[..., None] * (1.0 - d * (d / mag))
    return means, diag, (t_mean, dir_var, rad_var)


def encoding_freqs(n_freq, max_exp):
    """nerfstudio NeRFEncoding: 2 ** linspace(min_freq_exp=0, max_freq_exp, num_frequencies) (fp32)."""
    return 2 ** torch.linspace(0.0, max_exp, n_freq)


def nerf_encoding(x, n_freq, max_exp, diag=None):
    """nerfstudio NeRFEncoding.pytorch_fwd (include_input=True; the raw input is appended LAST):
    [sin(2 pi x_d f_k) (d-major, k-minor), sin(2 pi x_d f_k + pi/2), x]; with covariances every feature is
    damped by exp(-0.5 diag_d f_k^2) (expected_sin; no (2 pi)^2 factor on the variance)."""
    freqs = encoding_freqs(n_freq, max_exp)
    scaled = (2 * torch.pi * x)[..., None] * freqs
    scaled = scaled.reshape(*scaled.shape[:-2], -1)
    arg = torch.cat([scaled, scaled + torch.pi / 2.0], dim=-1)
    if diag is None:
        enc = torch.sin(arg)
    else:
        var = diag[..., :, None] * freqs[None, :] ** 2
        var = var.reshape(*var.shape[:-2], -1)
        enc = torch.exp(-0.5 * torch.cat(2 * [var], dim=-1)) * torch.sin(arg)
    return torch.cat([enc, x], dim=-1)


# ----------------------------------------------------------------------------- field
def mip_field(params, prefix, origins, directions, starts, ends, emulate=False, return_raw=False):
    """models/mipnerf.py:89-100 -> nerfstudio NeRFField(use_integrated_encoding=True):
    get_density (IPE -> mlp_base: 8 x 256, ReLU, input re-concatenated IN FRONT of layer 4's input,
    out_activation ReLU -> DensityFieldHead Linear(256,1)+Softplus) and get_outputs (mlp_head on
    cat[encoded_dir, base_out]: 2 x 128 ReLU -> RGBFieldHead Linear(128,3)+Sigmoid).
    pixel_area = 1 (star_mipnerf.py:267,308) -> cone radius sqrt(1)/sqrt(pi).
    emulate: False | "bf16" | "fp16": operand-rounding model of the tensor-core tier (GEMM operands rounded,
    fp32 accumulate, fp32 bias; the two heads in fp32)."""
    def q(t):
        if not emulate:
            return t
        return t.half().to(t.dtype) if emulate == "fp16" else t.bfloat16().to(t.dtype)

    def lin(name, h, gemm=True):
        w, b = params[f"{prefix}{name}.weight"], params[f"{prefix}{name}.bias"]
        return F.linear(q(h), q(w), b) if gemm else F.linear(h, w, b)

    R, S = starts.shape
    radius = math.sqrt(1.0) / 1.7724538509055159
    means, diag, _ = frustum_gaussian(origins, directions, starts, ends, radius)
    enc = nerf_encoding(means, N_FREQ_XYZ, MAX_EXP_XYZ, diag).reshape(R * S, D_XYZ)
    x = enc
    for i in range(N_BASE):
        if i == SKIP:
            x = torch.cat([enc, x], -1)
        x = F.relu(lin(f"field.mlp_base.layers.{i}", x))
    raw_sigma = lin("field.field_output_density.net", x, gemm=False)
    density = F.softplus(raw_sigma)
    e_d = nerf_encoding(directions, N_FREQ_DIR, MAX_EXP_DIR)[:, None, :].expand(R, S, D_DIR).reshape(R * S, D_DIR)
    h = torch.cat([e_d, x], -1)
    for i in range(2):
        h = F.relu(lin(f"field.mlp_head.layers.{i}", h))
    raw_rgb = lin("field.field_heads.0.net", h, gemm=False)
    if return_raw:
        return raw_sigma.reshape(R, S), raw_rgb.reshape(R, S, 3)
    return density.reshape(R, S, 1), torch.sigmoid(raw_rgb).reshape(R, S, 3)


# ----------------------------------------------------------------------------- compositing
def weights_alphas_transmittance(deltas, densities):
    """models/rendering_starmip.py:32-63 (deltas [R,S,1]; densities [R,S,1] or [R,V,S,1])."""
    dd = deltas * densities if densities.dim() == 3 else deltas[:, None, :, :] * densities
    alphas = 1 - torch.exp(-dd)
    tr = torch.cumsum(dd[..., :-1, :], dim=-2)
    tr = torch.cat([torch.zeros((*tr.shape[:-2], 1, 1)), tr], dim=-2)
    tr = torch.exp(-tr)
    return torch.nan_to_num(alphas * tr), alphas, tr


def median_depth(weights, starts, ends):
    """nerfstudio DepthRenderer(method="median"): first sample whose cumulative weight reaches 0.5
    (searchsorted side="left", clamped); depth = its (start + end) / 2.  weights [R,S,1] -> [R,1]."""
    steps = (starts + ends) / 2
    cw = torch.cumsum(weights[..., 0], dim=-1)
    split = torch.ones((*weights.shape[:-2], 1)) * 0.5
    idx = torch.clamp(torch.searchsorted(cw, split, side="left"), 0, steps.shape[-1] - 1)
    return torch.gather(steps, dim=-1, index=idx)


def appinit_outputs(density_s, rgb_s, deltas, starts, ends):
    """models/rendering_starmip.py:66-91."""
    w, a, T = weights_alphas_transmittance(deltas, density_s)
    return {"rgb": torch.sum(T * a * rgb_s, dim=-2), "acc": torch.sum(w, dim=-2), "weights": w,
            "depth": median_depth(w, starts, ends)}


def online_outputs(density_s, rgb_s, density_d, rgb_d, deltas, starts, ends):
    """models/rendering_starmip.py:112-175.  The regularisers are the vanilla functions
    (rendering__.py:612-715) applied to tensors with a trailing singleton dimension, which changes two of
    them: compute_ray_reg's max(dim=-1) runs over that singleton (no max over samples: the loss is
    sum_{v,s} mean_r (sigma_d/sigma_tot)^2 / V), and compute_static_reg -- fed transmittance_s for sigma_s
    (:156) -- normalises each alpha by itself, p = 1, so the loss is identically (-)0."""
    V = density_d.shape[1]
    w_s, a_s, T_s = weights_alphas_transmittance(deltas, density_s)
    w_d, a_d, T_d = weights_alphas_transmittance(deltas, density_d)
    total = density_s + density_d.sum(dim=1)
    w, a, T = weights_alphas_transmittance(deltas, total)
    rgb = torch.sum(T * (a_s * rgb_s + torch.sum(a_d * rgb_d, dim=1)), dim=-2)
    depth_d = torch.stack([median_depth(w_d[:, i], starts, ends) for i in range(V)], 1)   # [R,V,1]
    return {
        "rgb": rgb, "acc": torch.sum(w, dim=-2), "weights": w, "depth": median_depth(w, starts, ends),
        "rgb_static": torch.sum(T_s * a_s * rgb_s, dim=-2), "depth_static": median_depth(w_s, starts, ends),
        "rgb_dynamic": torch.sum(T_d * a_d * rgb_d, dim=-2), "depth_dynamic": depth_d[..., 0],
        "dynamic_transmittance": T_d[:, :, -1, :],
        "loss_alpha_entropy": so.alpha_entropy(a_s, a_d),
        "loss_dynamic_vs_static_reg": so.dynamic_vs_static_reg(a_s, a_d),
        "loss_ray_reg": so.ray_reg(density_d, total),
        "loss_static_reg": so.static_reg(T_s, a_s),
        "loss_dynamic_reg": so.dynamic_reg(density_d),
    }


# ----------------------------------------------------------------------------- STaR (mip) forward
class MipConfig:
    """The fields of `args` the mip path reads (star_mipnerf.py:45-86)."""

    def __init__(self, num_vehicles=0, N_samples=64, N_importance=128, chunk=8192, near=3.0, far=80.0,
                 scale_factor=0.01, emulate=False):
        self.num_vehicles = num_vehicles
        self.N_samples = N_samples
        self.N_importance = N_importance
        self.chunk = chunk
        self.near_plane = scale_factor * near     # NearFarCollider (star_mipnerf.py:83-86)
        self.far_plane = scale_factor * far
        self.emulate = emulate


def _fields(params, cfg, origins, viewdirs, pose, starts, ends):
    d_s, c_s = mip_field(params, "static_nerf.", origins, viewdirs, starts, ends, cfg.emulate)
    if pose is None:
        return d_s, c_s, None, None
    if pose.dim() != 2:
        raise NotImplementedError          # star_mipnerf.py:195-198 (4x4 poses are not supported by the mip variant)
    dd, cd = [], []
    for i in range(cfg.num_vehicles):      # star_mipnerf.py:206-240: origins and directions move, starts/ends stay
        o_i = so.se3_act(pose[i], origins)
        v_i = so.so3_act(pose[i, 3:], viewdirs)
        a, c = mip_field(params, f"dynamic_nerfs.{i}.", o_i, v_i, starts, ends, cfg.emulate)
        dd.append(a)
        cd.append(c)
    return d_s, c_s, torch.stack(dd, 1), torch.stack(cd, 1)


def _mip_chunk(params, cfg, origins, viewdirs, pose, training, t_rand, u_rand, details=None, exact_sum=False):
    """star_mipnerf.py:262-357 (__forward_app_init / __forward_online)."""
    R = origins.shape[0]
    near, far = cfg.near_plane, cfg.far_plane
    sp = uniform_bins(R, cfg.N_samples, training, t_rand)
    eu = spacing_to_euclidean(sp, near, far)
    res = {}
    for tag in ("0", ""):
        starts, ends = eu[..., :-1], eu[..., 1:]
        deltas = (ends - starts)[..., None]
        d_s, c_s, d_d, c_d = _fields(params, cfg, origins, viewdirs, pose, starts, ends)
        if pose is None:
            out = appinit_outputs(d_s, c_s, deltas, starts, ends)
        else:
            out = online_outputs(d_s, c_s, d_d, c_d, deltas, starts, ends)
        for k, v in out.items():
            res[k + tag] = v
        if details is not None:
            details["bins" + tag] = eu
        if tag == "0":
            sp = pdf_sample(sp, out["weights"][..., 0], cfg.N_importance, training, u_rand, exact_sum=exact_sum)
            eu = spacing_to_euclidean(sp, near, far)
    return res


def star_mip_forward(params, cfg, origins, viewdirs, pose=None, training=False, t_rand=None, u_rand=None,
                     details=None, exact_sum=False):
    """star_mipnerf.py:99-137: ray chunks of cfg.chunk; per-ray tensors concatenated, 0-dim outputs summed."""
    acc = {}
    dets = []
    for i in range(0, origins.shape[0], cfg.chunk):
        j = min(origins.shape[0], i + cfg.chunk)
        d = {} if details is not None else None
        part = _mip_chunk(params, cfg, origins[i:j], viewdirs[i:j], pose, training,
                          None if t_rand is None else t_rand[i:j], None if u_rand is None else u_rand[i:j], d,
                          exact_sum)
        if d is not None:
            dets.append(d)
        for k, v in part.items():
            acc.setdefault(k, []).append(v)
    if details is not None:
        for k in dets[0]:
            details[k] = torch.cat([d[k] for d in dets], 0)
    return {k: (sum(v) if v[0].dim() == 0 else torch.cat(v, 0)) for k, v in acc.items()}


# ----------------------------------------------------------------------------- synthetic inputs
def mip_param_names():
    names = [f"field.mlp_base.layers.{i}" for i in range(N_BASE)]
    names += ["field.field_output_density.net"]
    names += [f"field.mlp_head.layers.{i}" for i in range(2)]
    names += ["field.field_heads.0.net"]
    return names


def mip_param_shapes():
    shp = {}
    for i in range(N_BASE):
        k = D_XYZ if i == 0 else (W_BASE + D_XYZ if i == SKIP else W_BASE)
        shp[f"field.mlp_base.layers.{i}"] = (W_BASE, k)
    shp["field.field_output_density.net"] = (1, W_BASE)
    shp["field.mlp_head.layers.0"] = (W_HEAD, W_BASE + D_DIR)
    shp["field.mlp_head.layers.1"] = (W_HEAD, W_HEAD)
    shp["field.field_heads.0.net"] = (3, W_HEAD)
    return shp


def init_mip_params(num_vehicles, seed=0, gain=1.0, bias_std=0.0):
    """Random-init weights in the state_dict layout of the reference module tree (star_mipnerf.py:60-71:
    static_nerf / dynamic_nerfs.{i} are MipNerfModel, whose only parameters are field.*), nn.Linear default
    initialiser (nerfstudio MLP).  `gain` > 1 widens the weights so that densities are not all tiny."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for prefix in ["static_nerf."] + [f"dynamic_nerfs.{v}." for v in range(num_vehicles)]:
        for name, (o, i) in mip_param_shapes().items():
            b = 1.0 / math.sqrt(i)
            out[f"{prefix}{name}.weight"] = (torch.rand(o, i, generator=g) * 2 - 1) * b * gain * math.sqrt(3.0)
            bias = (torch.rand(o, generator=g) * 2 - 1) * b
            if bias_std > 0:
                bias = bias + torch.randn(o, generator=g) * bias_std
            out[f"{prefix}{name}.bias"] = bias
    return out

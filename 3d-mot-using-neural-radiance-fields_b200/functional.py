"""torch-facing operators over the C ABI (libstar_b200.so): thin argument marshalling plus
torch.autograd.Function wrappers.  PyTorch supplies device memory, streams and the autograd graph;
all arithmetic happens in the CUDA kernels.  Reference citations are into /root/reference.
"""
import ctypes as C
import os
import weakref

import torch
from torch.autograd import Function

from . import _capi
from ._capi import check, f32, ptr, stream

# Upper bound on samples handled by one MLP launch group and on bytes of saved activations kept
# between forward and backward (beyond it the backward re-runs the forward per ray chunk).
MAX_SAMPLES_PER_LAUNCH = int(os.environ.get("STAR_B200_MAX_SAMPLES", 1 << 19))
STASH_BUDGET_BYTES = int(float(os.environ.get("STAR_B200_STASH_GB", "24")) * (1 << 30))

# A/B switch: run the tensor-core forward on the CTA-pair (cta_group::2) kernels instead of the one-CTA-per-SM ones (same
# results, measured slower: see include/star_b200.h)
TC_DX_PIPELINED = os.environ.get("STAR_B200_DX_PIPELINED", "0") == "1"
TC_NO_WSHARE = os.environ.get("STAR_B200_NO_WSHARE", "0") == "1"      # A/B: opt out of the weight-sharing cluster launch

# instrumentation for bench.py: number of kernel-launching C-ABI calls issued
LAUNCH_COUNTER = {"calls": 0}


def _count(n=1):
    LAUNCH_COUNTER["calls"] += n


# bench.py sets PROFILE = {} to collect (start_event, end_event, samples, algorithmic FLOPs) per named C-ABI call on the
# launching stream (CUDA events only; no synchronisation is added).
MLP_MAC_PER_SAMPLE = {4: 708352, 2: 446208}    # static / dynamic net; 2 FLOP per MAC
MIP_MAC_PER_SAMPLE = 587264
PROFILE = None


def _prof_begin():
    if PROFILE is None:
        return None
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


def _prof_end(name, e0, samples, flops):
    if e0 is None:
        return
    e1 = torch.cuda.Event(enable_timing=True)
    e1.record()
    PROFILE.setdefault(name, []).append((e0, e1, samples, flops))


def _c(t):
    return t if t.is_contiguous() else t.contiguous()


# ------------------------------------------------------------------------------------------ a1
_LINSPACE_CACHE = {}


def _linspace01(n, device):
    key = (n, str(device))
    t = _LINSPACE_CACHE.get(key)
    if t is None:
        t = torch.linspace(0.0, 1.0, steps=n).to(device)   # host formula, then one H2D copy (cached)
        _LINSPACE_CACHE[key] = t
    return t


def sample_pts(rays_o, rays_d, near, far, N_samples, lindisp=False, t_rand=None):
    """models/rendering__.py:75-112 on the GPU.  t_rand: injected jitter or None."""
    rays_o, rays_d = _c(rays_o), _c(rays_d)
    R = rays_o.shape[0]
    pts = torch.empty((R, N_samples, 3), device=rays_o.device, dtype=torch.float32)
    z = torch.empty((R, N_samples), device=rays_o.device, dtype=torch.float32)
    t_vals = _linspace01(N_samples, rays_o.device)
    if t_rand is not None:
        t_rand = _c(t_rand)
    check(_capi.lib().star_sample_pts(f32(rays_o), f32(rays_d), f32(t_vals), f32(t_rand) if t_rand is not None else None,
                                      float(near), float(far), R, N_samples, int(bool(lindisp)), f32(pts), f32(z),
                                      stream()), "star_sample_pts")
    _count()
    return pts, z


# ------------------------------------------------------------------------------------------ a2
def get_rays(H, W, K, c2w, rows=None, want_viewdirs=False):
    """models/rendering__.py:41-55 on the GPU: rays of the pixel rows `rows` = (row0, nrows) (default: the whole view)
    from the 3x4 camera-to-world matrix `c2w` (CUDA tensor).  Returns rays_o, rays_d [nrows, W, 3] (+ viewdirs)."""
    c2w = _c(c2w[:3, :4].to(torch.float32))
    row0, nrows = (0, H) if rows is None else rows
    dev = c2w.device
    ro = torch.empty((nrows, W, 3), device=dev)
    rd = torch.empty((nrows, W, 3), device=dev)
    vd = torch.empty((nrows, W, 3), device=dev) if want_viewdirs else None
    check(_capi.lib().star_get_rays(int(H), int(W), float(K[0][0]), float(K[1][1]), float(K[0][2]), float(K[1][2]),
                                    f32(c2w), int(row0), int(nrows), f32(ro), f32(rd), ptr(vd), stream()), "star_get_rays")
    _count()
    return (ro, rd, vd) if want_viewdirs else (ro, rd)


# ------------------------------------------------------------------------------------------ a3
def embed(x, L, scale=None):
    """models/embedder.py:81-112 (stand-alone form; the MLP kernels encode in registers)."""
    x = _c(x)
    M = x.shape[0]
    out = torch.empty((M, 3 + 6 * L), device=x.device, dtype=torch.float32)
    check(_capi.lib().star_embed(f32(x), M, L, f32(scale) if scale is not None else None, f32(out), stream()),
          "star_embed")
    _count()
    return out


# ------------------------------------------------------------------------------------------ a5/a6
class CompositeSingle(Function):
    """raw2outputs (models/rendering__.py:307-379): returns rgb, disp, acc, depth, weights, dists."""

    @staticmethod
    def forward(ctx, raw_alpha, raw_rgb, z_vals, rays_d, far_dist, white_bkgd):
        raw_alpha, raw_rgb, z_vals, rays_d = _c(raw_alpha), _c(raw_rgb), _c(z_vals), _c(rays_d)
        R, S = raw_alpha.shape
        dev = raw_alpha.device
        rgb = torch.empty((R, 3), device=dev)
        disp, acc, depth = (torch.empty((R,), device=dev) for _ in range(3))
        weights = torch.empty((R, S), device=dev)
        dists = torch.empty((R, S), device=dev)
        check(_capi.lib().star_composite_single_forward(
            f32(raw_alpha), f32(raw_rgb), f32(z_vals), f32(rays_d), R, S, float(far_dist), int(bool(white_bkgd)),
            f32(rgb), f32(disp), f32(acc), f32(depth), f32(weights), f32(dists), stream()),
            "star_composite_single_forward")
        _count()
        ctx.save_for_backward(raw_alpha, raw_rgb, z_vals, rays_d)
        ctx.cfg = (float(far_dist), int(bool(white_bkgd)))
        ctx.mark_non_differentiable(dists)
        return rgb, disp, acc, depth, weights, dists

    @staticmethod
    def backward(ctx, g_rgb, g_disp, g_acc, g_depth, g_weights, _g_dists):
        raw_alpha, raw_rgb, z_vals, rays_d = ctx.saved_tensors
        R, S = raw_alpha.shape
        far_dist, white = ctx.cfg
        d_alpha = torch.empty_like(raw_alpha)
        d_rgb = torch.empty_like(raw_rgb)
        g = [None if t is None else _c(t) for t in (g_rgb, g_disp, g_acc, g_depth, g_weights)]
        check(_capi.lib().star_composite_single_backward(
            f32(raw_alpha), f32(raw_rgb), f32(z_vals), f32(rays_d), R, S, far_dist, white,
            *[f32(t) if t is not None else None for t in g], f32(d_alpha), f32(d_rgb), stream()),
            "star_composite_single_backward")
        _count()
        return d_alpha, d_rgb, None, None, None, None


# ------------------------------------------------------------------------------------------ a7/a8
STAR_OUT_KEYS = ("rgb", "disp", "acc", "depth", "weights", "rgb_static", "depth_static", "rgb_dynamic",
                 "depth_dynamic", "dynamic_transmittance", "rgb_dynamic_all", "regs")


class CompositeStar(Function):
    """raw2outputs_star + regularisers (models/rendering__.py:383-576, :612-715).
    Returns the tensors of STAR_OUT_KEYS (rgb_dynamic_all is an empty tensor when test=False)."""

    @staticmethod
    def forward(ctx, ra_s, rc_s, ra_d, rc_d, z_vals, rays_d, far_dist, white_bkgd, chunk, test):
        ra_s, rc_s, ra_d, rc_d, z_vals, rays_d = (_c(t) for t in (ra_s, rc_s, ra_d, rc_d, z_vals, rays_d))
        R, V, S = ra_d.shape
        dev = ra_s.device
        e = lambda *s: torch.empty(s, device=dev, dtype=torch.float32)
        out = dict(rgb=e(R, 3), disp=e(R), acc=e(R), depth=e(R), weights=e(R, S), rgb_static=e(R, 3),
                   depth_static=e(R), rgb_dynamic=e(R, V, 3), depth_dynamic=e(R, V), dynamic_transmittance=e(R, V),
                   rgb_dynamic_all=e(R, 3) if test else e(0), regs=e(5))
        L = _capi.lib()
        ws = torch.empty((L.star_composite_multi_ws_bytes(R) // 4,), device=dev, dtype=torch.float32)
        mo = _capi.StarMultiOut(*[ptr(out[k]) if out[k].numel() else None for k in STAR_OUT_KEYS])
        check(L.star_composite_multi_forward(f32(ra_s), f32(rc_s), f32(ra_d), f32(rc_d), f32(z_vals), f32(rays_d),
                                             R, V, S, float(far_dist), int(bool(white_bkgd)), int(chunk),
                                             C.byref(mo), ptr(ws), stream()), "star_composite_multi_forward")
        _count(2)
        ctx.save_for_backward(ra_s, rc_s, ra_d, rc_d, z_vals, rays_d)
        ctx.cfg = (float(far_dist), int(bool(white_bkgd)), int(chunk))
        # per-field visualisation products: not differentiated (reference training never reads them
        # in a loss, train_online__.py:158-273)
        ctx.mark_non_differentiable(out["rgb_static"], out["depth_static"], out["rgb_dynamic"],
                                    out["depth_dynamic"], out["dynamic_transmittance"], out["rgb_dynamic_all"])
        return tuple(out[k] for k in STAR_OUT_KEYS)

    @staticmethod
    def backward(ctx, g_rgb, g_disp, g_acc, g_depth, g_weights, _a, _b, _c2, _d, _e, _f, g_regs):
        ra_s, rc_s, ra_d, rc_d, z_vals, rays_d = ctx.saved_tensors
        R, V, S = ra_d.shape
        far_dist, white, chunk = ctx.cfg
        d = [torch.empty_like(t) for t in (ra_s, rc_s, ra_d, rc_d)]
        g = [None if t is None else _c(t) for t in (g_rgb, g_disp, g_acc, g_depth, g_weights, g_regs)]
        check(_capi.lib().star_composite_multi_backward(
            f32(ra_s), f32(rc_s), f32(ra_d), f32(rc_d), f32(z_vals), f32(rays_d), R, V, S, far_dist, white, chunk,
            *[f32(t) if t is not None else None for t in g], *[f32(t) for t in d], stream()),
            "star_composite_multi_backward")
        _count()
        return d[0], d[1], d[2], d[3], None, None, None, None, None, None


# ------------------------------------------------------------------------------------------ a9
def sample_pdf(bins, weights, N_samples, det=False, u=None, return_details=False):
    """models/rendering__.py:719-761.  `weights` may be a non-contiguous last-dim-contiguous view
    (the reference passes weights[..., 1:-1]); `u` injects the uniform draws."""
    bins = _c(bins)
    if weights.stride(-1) != 1:
        weights = weights.contiguous()
    R, nb = bins.shape
    assert weights.shape == (R, nb - 1), "sample_pdf: weights must have one column less than bins"
    dev = bins.device
    u_det = None
    if u is None:
        if det:
            u_det = _linspace01(N_samples, dev)
        else:
            u = torch.rand((R, N_samples), device=dev)
    if u is not None:
        u = _c(u)
    samples = torch.empty((R, N_samples), device=dev)
    details = {}
    if return_details:
        details = dict(inds=torch.empty((R, N_samples), device=dev, dtype=torch.int64),
                       below=torch.empty((R, N_samples), device=dev, dtype=torch.int64),
                       above=torch.empty((R, N_samples), device=dev, dtype=torch.int64),
                       cdf=torch.empty((R, nb), device=dev))
    wptr = weights.data_ptr()
    check(_capi.lib().star_sample_pdf(f32(bins), nb, wptr, weights.stride(0), f32(u) if u is not None else None,
                                      f32(u_det) if u_det is not None else None, R, nb, N_samples, f32(samples),
                                      ptr(details.get("inds")), ptr(details.get("below")), ptr(details.get("above")),
                                      ptr(details.get("cdf")), stream()), "star_sample_pdf")
    _count()
    if return_details:
        details["u"] = u if u is not None else u_det.expand(R, N_samples)
        return samples, details
    return samples


def invert_cdf(bins, cdf, u):
    """searchsorted(right=True) + gather + lerp on a caller-supplied cdf (:744-761)."""
    bins, cdf, u = _c(bins), _c(cdf), _c(u)
    R, nb = cdf.shape
    Ni = u.shape[1]
    dev = bins.device
    samples = torch.empty((R, Ni), device=dev)
    inds, below, above = (torch.empty((R, Ni), device=dev, dtype=torch.int64) for _ in range(3))
    check(_capi.lib().star_invert_cdf(f32(bins), f32(cdf), f32(u), R, nb, Ni, f32(samples), ptr(inds), ptr(below),
                                      ptr(above), stream()), "star_invert_cdf")
    _count()
    return samples, inds, below, above


def hierarchical(z_vals, weights, N_importance, det, rays_o, rays_d, u=None, want_pts=True, z_samples=None):
    """models/rendering__.py:128-144 / :271-296 in one kernel.  Returns z_samples, z_all, z_std, pts_fine.
    Everything is detached, as in the reference (z_samples.detach(), :135,278).
    z_samples: injected fine samples (skips the inverse-CDF step; merge / std / pts only)."""
    z_vals, weights = _c(z_vals.detach()), _c(weights.detach())
    rays_o, rays_d = _c(rays_o), _c(rays_d)
    R, Nc = z_vals.shape
    dev = z_vals.device
    if z_samples is not None:
        zs = _c(z_samples.detach())
        assert zs.shape == (R, N_importance)
        z_all = torch.empty((R, Nc + N_importance), device=dev)
        z_std = torch.empty((R,), device=dev)
        pts = torch.empty((R, Nc + N_importance, 3), device=dev) if want_pts else None
        check(_capi.lib().star_merge_samples(f32(z_vals), f32(zs), f32(rays_o), f32(rays_d), R, Nc, N_importance,
                                             f32(z_all), f32(z_std), ptr(pts), stream()), "star_merge_samples")
        _count()
        return zs, z_all, z_std, pts
    u_det = None
    if u is None:
        if det:
            u_det = _linspace01(N_importance, dev)
        else:
            u = torch.rand((R, N_importance), device=dev)
    if u is not None:
        u = _c(u)
    zs = torch.empty((R, N_importance), device=dev)
    z_all = torch.empty((R, Nc + N_importance), device=dev)
    z_std = torch.empty((R,), device=dev)
    pts = torch.empty((R, Nc + N_importance, 3), device=dev) if want_pts else None
    check(_capi.lib().star_hierarchical(f32(z_vals), f32(weights), f32(u) if u is not None else None,
                                        f32(u_det) if u_det is not None else None, f32(rays_o), f32(rays_d), R, Nc,
                                        N_importance, f32(zs), f32(z_all), f32(z_std), ptr(pts), stream()),
          "star_hierarchical")
    _count()
    return zs, z_all, z_std, pts


def composite_hier(raw_alpha, raw_rgb, z_vals, rays_d, far_dist, white_bkgd, N_importance, det, u=None, want_weights=True):
    """Coarse-pass tail of a single-field render in ONE kernel (star_composite_hier_forward, csrc/ray_fused.cu):
    raw2outputs (models/rendering__.py:307-379) + z_mid / sample_pdf / sort(cat) / std (:128-144).  No autograd graph
    (inference; the fine samples are detached in the reference anyway).  Returns the dict
    rgb, disp, acc, depth, weights, dists, z_samples, z_vals, z_std; weights / dists are None when want_weights=False.
    Needs an even Nc (the kernel's two-samples-per-lane form): StarError(STAR_E_UNSUPPORTED) otherwise."""
    raw_alpha, raw_rgb, z_vals, rays_d = _c(raw_alpha.detach()), _c(raw_rgb.detach()), _c(z_vals.detach()), _c(rays_d)
    R, Nc = raw_alpha.shape
    dev = raw_alpha.device
    u_det = None
    if u is None:
        if det:
            u_det = _linspace01(N_importance, dev)
        else:
            u = torch.rand((R, N_importance), device=dev)
    if u is not None:
        u = _c(u)
    rgb = torch.empty((R, 3), device=dev)
    disp, acc, depth, z_std = (torch.empty((R,), device=dev) for _ in range(4))
    weights = torch.empty((R, Nc), device=dev) if want_weights else None
    dists = torch.empty((R, Nc), device=dev) if want_weights else None
    zs = torch.empty((R, N_importance), device=dev)
    z_all = torch.empty((R, Nc + N_importance), device=dev)
    check(_capi.lib().star_composite_hier_forward(
        f32(raw_alpha), f32(raw_rgb), f32(z_vals), f32(rays_d), f32(u) if u is not None else None,
        f32(u_det) if u_det is not None else None, R, Nc, N_importance, float(far_dist), int(bool(white_bkgd)), f32(rgb),
        f32(disp), f32(acc), f32(depth), ptr(weights), ptr(dists), f32(zs), f32(z_all), f32(z_std), stream()),
        "star_composite_hier_forward")
    _count()
    return dict(rgb=rgb, disp=disp, acc=acc, depth=depth, weights=weights, dists=dists, z_samples=zs, z_vals=z_all,
                z_std=z_std)


# ------------------------------------------------------------------------------------------ a4 + K2
def flat_master(params):
    """Flat fp32 master vector of a net: a zero-copy view when its parameters already sit back to back in one storage
    in master order (optim.flatten_parameters re-homes them that way), else a concatenated copy."""
    p0 = params[0]
    addr = p0.data_ptr()
    for p in params:
        if p.data_ptr() != addr or not p.is_contiguous() or p.dtype != torch.float32:
            return torch.cat([q.detach().reshape(-1).float() for q in params])
        addr += 4 * p.numel()
    total = (addr - p0.data_ptr()) // 4
    if (p0.storage_offset() + total) * 4 > p0.untyped_storage().nbytes():
        return torch.cat([q.detach().reshape(-1) for q in params])
    return torch.as_strided(p0.detach(), (total,), (1,))


class NetRuntime:
    """Kernel-side state of one NeRF MLP: the flat fp32 master vector (order of include/star_b200.h)
    and the packed weight image, rebuilt when any parameter's version / storage changes."""

    def __init__(self, module, n_blocks, L_xyz, L_dir):
        self.module = module
        self.n_blocks, self.L_xyz, self.L_dir = n_blocks, L_xyz, L_dir
        self._key = None
        self._flat = None
        self._packed = {}
        # parallel.GradSync: flat gradient range of this net (the backward kernels accumulate straight into it) and the
        # object told when a forward starts / a backward has been launched
        self.grad_sink = None
        self.sink_owner = None

    def ordered_params(self):
        m = self.module
        lins = [m.pts_net.lin_in]
        for b in m.pts_net.blocks:
            lins += [b.fc_0, b.fc_1]
        lins += [m.pts_net.lin_out, m.alpha_linear, m.feature_linear, m.views_linears[0], m.rgb_linear]
        out = []
        for l in lins:
            out += [l.weight, l.bias]
        return out

    def desc(self, precision):
        if TC_DX_PIPELINED and precision != _capi.PREC_F32:
            precision = precision | _capi.PREC_FLAG_DX_PIPELINED
        if TC_NO_WSHARE and precision != _capi.PREC_F32:
            precision = precision | _capi.PREC_FLAG_NO_WSHARE
        return _capi.net_desc(self.n_blocks, self.L_xyz, self.L_dir, precision)

    def refresh(self, precision):
        params = self.ordered_params()
        key = tuple((p.data_ptr(), p._version) for p in params)
        if key != self._key:
            self._flat = flat_master(params)
            self._packed = {}
            self._key = key
        if precision not in self._packed:
            d = self.desc(precision)
            L = _capi.lib()
            nbytes = L.star_packed_bytes(C.byref(d))
            if nbytes == 0:
                raise _capi.StarError("unsupported network shape / precision for the packed weight image")
            assert self._flat.numel() == L.star_net_param_count(C.byref(d))
            packed = torch.empty((nbytes,), device=self._flat.device, dtype=torch.uint8)
            check(L.star_pack_weights(C.byref(d), f32(self._flat), ptr(packed), stream()), "star_pack_weights")
            _count()
            self._packed[precision] = packed
        return self._flat, self._packed[precision]


STASH_RESERVE_BYTES = int(float(os.environ.get("STAR_B200_STASH_RESERVE_GB", "14")) * (1 << 30))
_STASH_LIVE = {}       # device index -> bytes of activation stashes currently alive
_DEVICE_BYTES = {}     # device index -> total memory (queried once; cudaMemGetInfo costs ~2 ms per call)


def _dev_index(device):
    return device.index if device.index is not None else torch.cuda.current_device()


def _stash_release(idx, nbytes):
    _STASH_LIVE[idx] -= nbytes


def stash_fits(nbytes, device):
    """Keep the activations of a forward for its backward?  Yes when they fit the per-call budget AND, together with
    the stashes that are still alive (earlier nets of the same step), leave STASH_RESERVE_BYTES of the device free
    for everything else; otherwise the backward re-runs the forward chunk by chunk.  Pure host bookkeeping."""
    if nbytes > STASH_BUDGET_BYTES:
        return False
    idx = _dev_index(device)
    total = _DEVICE_BYTES.get(idx)
    if total is None:
        total = _DEVICE_BYTES[idx] = torch.cuda.get_device_properties(idx).total_memory
    return _STASH_LIVE.get(idx, 0) + nbytes + STASH_RESERVE_BYTES <= total


def stash_alloc(nbytes, device):
    """A stash buffer whose bytes count as alive until the tensor is garbage-collected (the backward drops it as soon
    as the chunk is done; a graph that is never back-propagated drops it with the graph)."""
    t = torch.empty((nbytes,), device=device, dtype=torch.uint8)
    idx = _dev_index(device)
    _STASH_LIVE[idx] = _STASH_LIVE.get(idx, 0) + nbytes
    weakref.finalize(t, _stash_release, idx, nbytes)
    return t


# ---- range guard of the 16-bit tiers: one int32 per device in mapped pinned host memory.  The MLP kernels set it to 1 when
# a raw output is not finite (an fp16 operand overflowed, or the inputs / weights were not finite); the host looks at it
# WITHOUT synchronising at the next MLP call (an overflow therefore raises one call late at the latest) and on demand.
_RANGE_FLAGS = {}


def _range_flag(device):
    idx = _dev_index(device)
    t = _RANGE_FLAGS.get(idx)
    if t is None:
        t = _RANGE_FLAGS[idx] = torch.zeros(1, dtype=torch.int32).pin_memory()
    return t


def check_range(sync=True):
    """Raises StarError if any tensor-core MLP launch so far produced a non-finite raw output (fp16 activations beyond
    65504: switch that net to set_precision('bf16') or 'fp32').  sync=True waits for the device first."""
    if sync and _RANGE_FLAGS:
        torch.cuda.synchronize()
    for idx, t in _RANGE_FLAGS.items():
        if int(t[0]) != 0:
            t[0] = 0
            raise _capi.StarError("cuda:%d: a tensor-core MLP launch produced non-finite raw outputs (fp16 operand range "
                                  "exceeded, or non-finite inputs / weights); use set_precision('bf16') or 'fp32'" % idx)


def _geom_ptrs(geom, a, b):
    """geom = pts [R,S,3]  or  (rays_o [R,3], rays_d [R,3], z_vals [R,S])  ->  the four pointer arguments of
    star_mlp_forward / star_mlp_backward for the ray range [a, b)."""
    if torch.is_tensor(geom):
        return f32(geom[a:b]), None, None, None
    ro, rd, z = geom
    return None, f32(ro[a:b]), f32(rd[a:b]), f32(z[a:b])


def _ray_chunks(R, S):
    per = max(1, MAX_SAMPLES_PER_LAUNCH // S)
    return [(a, min(R, a + per)) for a in range(0, R, per)]


class NerfRaw(Function):
    """One radiance MLP on (pts, viewdirs) with an optional rigid transform into the object frame:
    returns RAW (raw_alpha [R,S], raw_rgb [R,S,3]) like NeRF.forward with z_vals=None
    (models/nerf.py:112-179).  pose12 = [R(3x3) row-major | t] or None."""

    @staticmethod
    def forward(ctx, rt, precision, grad_mode, pts, viewdirs, pose12, sc_xyz, sc_dir, *params):
        # pts: the materialised positions [R,S,3], or the tuple (rays_o, rays_d, z_vals) -- the kernel then forms
        # rays_o + rays_d * z itself (bit-identical; nothing of size [R,S,3] is read)
        if torch.is_tensor(pts):
            geom = _c(pts)
            R, S = geom.shape[0], geom.shape[1]
        else:
            geom = tuple(_c(t.detach()) for t in pts)
            R, S = geom[2].shape
        viewdirs = _c(viewdirs)
        dev = viewdirs.device
        flat, packed = rt.refresh(precision)
        d = rt.desc(precision)
        L = _capi.lib()
        flag = None
        if precision != _capi.PREC_F32:
            flag = _range_flag(dev)
            if int(flag[0]) != 0:
                check_range(sync=False)
        raw_alpha = torch.empty((R, S), device=dev)
        raw_rgb = torch.empty((R, S, 3), device=dev)
        # ctx.needs_input_grad ignores torch.no_grad(): `grad_mode` (torch.is_grad_enabled() at the call site) decides
        # whether anything is kept for a backward pass at all
        need_grad = grad_mode and (any(ctx.needs_input_grad[8:]) or (pose12 is not None and ctx.needs_input_grad[5]))
        p12 = _c(pose12.detach()) if pose12 is not None else None
        chunks = _ray_chunks(R, S)
        stashes = []
        # activations needed by the backward are stashed by the forward of the same tier (fp32: per-GEMM inputs in
        # fp32; tensor-core tiers: 16-bit swizzled blocks); over budget, the backward re-runs the forward per chunk
        keep = need_grad and stash_fits(L.star_stash_bytes(C.byref(d), R * S), dev)
        for (a, b) in chunks:
            st = None
            if keep:
                st = stash_alloc(L.star_stash_bytes(C.byref(d), (b - a) * S), dev)
            stashes.append(st)
            e0 = _prof_begin()
            check(L.star_mlp_forward(C.byref(d), ptr(packed), *_geom_ptrs(geom, a, b), f32(viewdirs[a:b]),
                                     f32(p12) if p12 is not None else None,
                                     f32(sc_xyz) if sc_xyz is not None else None,
                                     f32(sc_dir) if sc_dir is not None else None, b - a, S, f32(raw_alpha[a:b]),
                                     f32(raw_rgb[a:b]), S, ptr(st), flag.data_ptr() if flag is not None else None,
                                     stream()), "star_mlp_forward")
            _prof_end("mlp_forward_stash" if st is not None else "mlp_forward", e0, (b - a) * S,
                      2.0 * MLP_MAC_PER_SAMPLE.get(rt.n_blocks, 0) * (b - a) * S)
            _count()
        if need_grad:
            if rt.sink_owner is not None:
                rt.sink_owner.on_forward(rt)
            ctx.rt, ctx.precision, ctx.chunks, ctx.stashes = rt, precision, chunks, stashes
            ctx.flat, ctx.packed = flat, packed
            ctx.scales = (sc_xyz, sc_dir)
            ctx.geom_is_pts = torch.is_tensor(geom)
            saved = (geom,) if ctx.geom_is_pts else geom
            ctx.save_for_backward(viewdirs, p12 if p12 is not None else torch.empty(0, device=dev), *saved)
            ctx.shapes = [p.shape for p in params]
        return raw_alpha, raw_rgb

    @staticmethod
    def backward(ctx, g_alpha, g_rgb):
        viewdirs, p12, *saved = ctx.saved_tensors
        geom = saved[0] if ctx.geom_is_pts else tuple(saved)
        if p12.numel() == 0:
            p12 = None
        rt, precision = ctx.rt, ctx.precision
        R, S = g_alpha.shape if g_alpha is not None else g_rgb.shape[:2]
        dev = viewdirs.device
        d = rt.desc(precision)
        L = _capi.lib()
        sc_xyz, sc_dir = ctx.scales
        g_alpha = torch.zeros((R, S), device=dev) if g_alpha is None else _c(g_alpha)
        g_rgb = torch.zeros((R, S, 3), device=dev) if g_rgb is None else _c(g_rgb)
        sink = rt.grad_sink if (rt.grad_sink is not None and rt.grad_sink.numel() == ctx.flat.numel()) else None
        grad_flat = sink if sink is not None else torch.zeros_like(ctx.flat)
        pose_acc = torch.zeros((32,), device=dev) if p12 is not None else None
        for i, (a, b) in enumerate(ctx.chunks):
            n = (b - a) * S
            st = ctx.stashes[i]
            if st is None:
                # over the stash budget, or a second backward through a retained graph (the stash of a chunk is dropped
                # as soon as it has been used): recompute the forward of this ray chunk with activation stashing
                st = torch.empty((L.star_stash_bytes(C.byref(d), n),), device=dev, dtype=torch.uint8)
                tmp_a = torch.empty((b - a, S), device=dev)
                tmp_c = torch.empty((b - a, S, 3), device=dev)
                check(L.star_mlp_forward(C.byref(d), ptr(ctx.packed), *_geom_ptrs(geom, a, b), f32(viewdirs[a:b]),
                                         f32(p12) if p12 is not None else None,
                                         f32(sc_xyz) if sc_xyz is not None else None,
                                         f32(sc_dir) if sc_dir is not None else None, b - a, S, f32(tmp_a),
                                         f32(tmp_c), S, ptr(st), None, stream()), "star_mlp_forward(recompute)")
                _count()
            ws = torch.empty((L.star_mlp_backward_workspace_bytes(C.byref(d), n),), device=dev, dtype=torch.uint8)
            e0 = _prof_begin()
            check(L.star_mlp_backward(C.byref(d), ptr(ctx.packed), f32(ctx.flat), *_geom_ptrs(geom, a, b),
                                      f32(viewdirs[a:b]), f32(p12) if p12 is not None else None,
                                      f32(sc_xyz) if sc_xyz is not None else None,
                                      f32(sc_dir) if sc_dir is not None else None, b - a, S, f32(g_alpha[a:b]),
                                      f32(g_rgb[a:b]), S, ptr(st), ptr(ws), f32(grad_flat),
                                      f32(pose_acc) if pose_acc is not None else None, stream()),
                  "star_mlp_backward")
            _prof_end("mlp_backward", e0, n, 4.0 * MLP_MAC_PER_SAMPLE.get(rt.n_blocks, 0) * n)   # dX + dW
            _count(3 + 2 * rt.n_blocks + 4)
            ctx.stashes[i] = None
        grads, off = [], 0
        if sink is not None:     # the parameters' .grad are views of the sink: nothing to hand to autograd
            grads = [None] * len(ctx.shapes)
            rt.sink_owner.on_backward_launched(rt)
        else:
            for shp in ctx.shapes:
                n = 1
                for s_ in shp:
                    n *= s_
                grads.append(grad_flat[off:off + n].view(shp))
                off += n
        g_pose = None
        if p12 is not None and ctx.needs_input_grad[5]:
            # Euclidean gradient w.r.t. [R | t]:  dR = sum g p^T + sum h d^T,  dt = sum g
            g_pose = torch.cat([pose_acc[3:12] + pose_acc[15:24], pose_acc[0:3]])
        return (None, None, None, None, None, g_pose, None, None, *grads)


# ------------------------------------------------------------------------------------------ a10 + a11 in one call
def render_forward(static_nets, dynamic_nets, precision, rays_o, rays_d, viewdirs, N_importance, *, z_vals=None, pts=None,
                   near=None, far=None, N_samples=None, lindisp=False, t_rand=None, pose12=None, det=True, u=None,
                   z_samples=None, white_bkgd=False, far_dist=1e10, chunk=None, test=True, enc_scales=(None, None),
                   camera=None):
    """The whole coarse -> fine render (models/rendering__.py:115-149, :249-298 on models/star__.py:119-225) as ONE
    C-ABI call, star_render_forward: every kernel is queued on the current stream without returning to Python.
    Inference only (nothing is kept for a backward pass).

    static_nets = (coarse NeRF module, fine NeRF module or None); dynamic_nets = (list of coarse, list of fine) object
    modules; pose12 [V,12].  Coarse depths: `z_vals` [R,Nc] (optionally with the caller's `pts`), or near / far /
    N_samples (sampled inside).  Rays: rays_o / rays_d / viewdirs [R,3], or camera = (H, W, K, c2w, (row0, nrows)) --
    rays generated inside.  Returns the reference's output dictionary (StarRenderOutput / NerfNetworkOutput keys)."""
    L = _capi.lib()
    sc_net, sf_net = static_nets
    dyn_c, dyn_f = dynamic_nets
    V = len(dyn_c)
    Ni = int(N_importance) if sf_net is not None else 0
    if camera is not None:
        H, W, K, c2w, (row0, nrows) = camera
        c2w = _c(c2w[:3, :4].to(torch.float32))
        dev, R = c2w.device, nrows * W
    else:
        rays_o, rays_d = _c(rays_o), _c(rays_d)
        viewdirs = _c(viewdirs) if viewdirs is not None else None
        dev, R = rays_o.device, rays_o.shape[0]
    if z_vals is not None:
        z_vals = _c(z_vals)
        Nc = z_vals.shape[1]
        pts = _c(pts) if pts is not None else None
    else:
        Nc = int(N_samples)
    Nf = Nc + Ni
    cfg = _capi.StarRenderCfg(R, Nc, Ni, V, int(precision) | (_capi.PREC_FLAG_NO_WSHARE if TC_NO_WSHARE and int(precision) != _capi.PREC_F32 else 0), sc_net._rt.n_blocks, dyn_c[0]._rt.n_blocks if V else 2,
                              sc_net._rt.L_xyz, sc_net._rt.L_dir, int(bool(white_bkgd)), int(bool(lindisp)),
                              int(bool(test)), int(chunk or max(R, 1)), float(near or 0.0), float(far or 0.0),
                              float(far_dist))
    flag = None
    if precision != _capi.PREC_F32:
        flag = _range_flag(dev)
        if int(flag[0]) != 0:
            check_range(sync=False)
    e = lambda *s_: torch.empty(s_, device=dev, dtype=torch.float32)

    def pass_out(S):
        o = dict(rgb=e(R, 3), disp=e(R), acc=e(R), depth=e(R), weights=e(R, S))
        if V:
            o.update(rgb_static=e(R, 3), depth_static=e(R), rgb_dynamic=e(R, V, 3), depth_dynamic=e(R, V),
                     dynamic_transmittance=e(R, V), rgb_dynamic_all=e(R, 3) if test else None, regs=e(5))
        return o
    oc, of = pass_out(Nc), (pass_out(Nf) if Ni > 0 else {})
    mo = lambda o: _capi.StarMultiOut(*[ptr(o.get(k)) for k in STAR_OUT_KEYS])
    extra = {}
    if V == 0:
        extra["dists0"] = e(R, Nc)
        if Ni > 0:
            extra["dists"] = e(R, Nf)
    if z_vals is None:
        extra["z_vals0"] = e(R, Nc)
    if Ni > 0:
        extra["z_vals"], extra["z_std"] = e(R, Nf), e(R)
        if z_samples is None:
            extra["z_samples"] = e(R, Ni)
    if camera is not None:
        extra["rays_o"], extra["rays_d"], extra["viewdirs"] = e(R, 3), e(R, 3), e(R, 3)
    out = _capi.StarRenderOut(mo(oc), mo(of), *[ptr(extra.get(k)) for k in
                                                ("dists0", "dists", "z_vals0", "z_vals", "z_samples", "z_std", "rays_o",
                                                 "rays_d", "viewdirs")])
    # packed weight images (rebuilt when a parameter changed)
    pk = lambda m: ptr(m._rt.refresh(precision)[1])
    keep = [sc_net._rt.refresh(precision)[1]]
    arr_c = (C.c_void_p * max(V, 1))(*[pk(m) for m in dyn_c])
    arr_f = (C.c_void_p * max(V, 1))(*[pk(m) for m in dyn_f]) if (V and Ni > 0) else (C.c_void_p * 1)()
    u_det = None
    if Ni > 0 and z_samples is None and u is None:
        if det:
            u_det = _linspace01(Ni, dev)
        else:
            u = torch.rand((R, Ni), device=dev)
    t_vals = _linspace01(Nc, dev) if z_vals is None else None
    fp = lambda t: f32(_c(t)) if t is not None else None
    sin = _capi.StarRenderIn(
        fp(rays_o) if camera is None else None, fp(rays_d) if camera is None else None, fp(viewdirs),
        *( (int(camera[0]), int(camera[1]), int(row0), int(nrows), float(K[0][0]), float(K[1][1]), float(K[0][2]),
            float(K[1][2]), f32(c2w)) if camera is not None else (0, 0, 0, 0, 0.0, 0.0, 0.0, 0.0, None)),
        fp(z_vals), fp(pts), fp(t_vals), fp(t_rand), fp(u), fp(u_det), fp(z_samples), fp(pose12),
        fp(enc_scales[0]), fp(enc_scales[1]), pk(sc_net), pk(sf_net) if Ni > 0 else None, arr_c, arr_f)
    ws_bytes = L.star_render_workspace_bytes(C.byref(cfg))
    ws = torch.empty((ws_bytes,), device=dev, dtype=torch.uint8)
    e0 = _prof_begin()
    check(L.star_render_forward(C.byref(cfg), C.byref(sin), C.byref(out), ptr(ws), ws_bytes,
                                flag.data_ptr() if flag is not None else None, stream()), "star_render_forward")
    mac = MLP_MAC_PER_SAMPLE.get(sc_net._rt.n_blocks, 0) + V * (MLP_MAC_PER_SAMPLE.get(dyn_c[0]._rt.n_blocks, 0) if V else 0)
    n_samp = R * (Nc + (Nf if Ni > 0 else 0))
    _prof_end("render_forward", e0, n_samp, 2.0 * mac * n_samp)
    passes = 2 if Ni > 0 else 1
    # single field, <= 128 (even) coarse samples, fine samples drawn by the call: compositing + hierarchical are one kernel
    fused_tail = V == 0 and Ni > 0 and z_samples is None and Nc % 2 == 0 and 4 <= Nc <= 128
    _count(passes * (1 + V + (2 if V else 1)) + (1 if Ni > 0 else 0) + (1 if z_vals is None else 0)
           + (1 if (camera is not None or viewdirs is None) else 0) - (1 if fused_tail else 0))
    del keep

    def to_dict(o, z, dists, sfx):
        d = {}
        for k in ("rgb", "disp", "acc", "weights", "depth"):
            d[k + sfx] = o[k]
        if V:
            for k in ("rgb_static", "depth_static", "rgb_dynamic", "depth_dynamic", "dynamic_transmittance"):
                d[k + sfx] = o[k]
            d["rgb_dynamic_all" + sfx] = o["rgb_dynamic_all"]
            r = o["regs"]
            for i, k in enumerate(("loss_alpha_entropy", "loss_dynamic_vs_static_reg", "loss_ray_reg", "loss_static_reg",
                                   "loss_dynamic_reg")):
                d[k + sfx] = r[i]
        else:
            d["dists" + sfx] = dists
            d["z_vals" + sfx] = z
        return d
    z0 = z_vals if z_vals is not None else extra["z_vals0"]
    res = to_dict(oc, z0, extra.get("dists0"), "0")
    if Ni > 0:
        res.update(to_dict(of, extra["z_vals"], extra.get("dists"), ""))
        res["z_std"] = extra["z_std"]
    res["_z_vals0"], res["_z_all"] = z0, extra.get("z_vals")
    res["_z_samples"] = extra.get("z_samples", z_samples)
    if camera is not None:
        res["_rays"] = (extra["rays_o"], extra["rays_d"], extra["viewdirs"])
    return res


class Pose7ToMat12(Function):
    """[tx,ty,tz,qx,qy,qz,qw] -> [R(q) row-major | t] with pypose's gradient convention
    (models/star__.py:191-196 via pp.SE3.Act / pp.SO3.Act): the gradient returned for the 7-vector is
    the LEFT tangent-space gradient padded with a zero, [dt, vee(R dR^T - dR R^T) + t x dt, 0], i.e.
    [sum g, sum p' x g + sum d' x h, 0] -- not the Euclidean derivative of the stored numbers.
    The quaternion is not normalised (pypose Act does not normalise either)."""

    @staticmethod
    def forward(ctx, pose7):
        x, y, z, w = pose7[3], pose7[4], pose7[5], pose7[6]
        K = torch.zeros((3, 3), device=pose7.device, dtype=pose7.dtype)
        K[0, 1], K[0, 2], K[1, 0], K[1, 2], K[2, 0], K[2, 1] = -z, y, z, -x, -y, x
        Rm = torch.eye(3, device=pose7.device, dtype=pose7.dtype) + 2.0 * w * K + 2.0 * (K @ K)
        ctx.save_for_backward(Rm, pose7[:3].clone())
        return torch.cat([Rm.reshape(9), pose7[:3]])

    @staticmethod
    def backward(ctx, g12):
        Rm, t = ctx.saved_tensors
        dR, dt = g12[:9].view(3, 3), g12[9:12]
        A = Rm @ dR.t()
        rot = torch.stack([A[1, 2] - A[2, 1], A[2, 0] - A[0, 2], A[0, 1] - A[1, 0]])
        rot = rot + torch.linalg.cross(t, dt)
        return torch.cat([dt, rot, torch.zeros(1, device=g12.device, dtype=g12.dtype)])


def pose_to_mat12(pose_v):
    """One object's pose -> 12-vector.  [4,4]: plain slicing (Euclidean autograd, star__.py:160-180);
    [7]: Pose7ToMat12."""
    if pose_v.dim() == 2:
        return torch.cat([pose_v[:3, :3].reshape(9), pose_v[:3, 3]])
    return Pose7ToMat12.apply(pose_v)

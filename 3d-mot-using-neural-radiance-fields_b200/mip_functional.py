"""torch-facing operators of the mip-NeRF variant over the C ABI: argument marshalling plus
torch.autograd.Function wrappers; all arithmetic happens in csrc/mip_f32.cu and csrc/mip_render.cu.
Reference citations are into /root/reference (models/star_mipnerf.py, models/rendering_starmip.py,
models/mipnerf.py; the underlying algorithms are nerfstudio's)."""
import ctypes as C
import math

import torch
from torch.autograd import Function

from . import _capi
from ._capi import check, f32, ptr, stream
from .functional import (_c, _count, _prof_begin, _prof_end, _ray_chunks, stash_alloc, stash_fits, MIP_MAC_PER_SAMPLE,
                         flat_master as _flat_master)

N_FREQ_XYZ, MAX_EXP_XYZ = 24, 24.0     # models/mipnerf.py:58-64
N_FREQ_DIR, MAX_EXP_DIR = 4, 4.0       # models/mipnerf.py:65-71
CONE_RADIUS = math.sqrt(1.0) / 1.7724538509055159   # pixel_area = 1 (star_mipnerf.py:267,308)

_TABLES = {}


def _table(key, make, device):
    k = (key, str(device))
    t = _TABLES.get(k)
    if t is None:
        t = make().to(device)     # host formula (bit-identical to the reference's torch ops), one H2D copy, cached
        _TABLES[k] = t
    return t


def freq_table(device):
    """[2**linspace(0,24,24) | its square | 2**linspace(0,4,4) | pad] (nerfstudio NeRFEncoding)."""
    def make():
        f = 2 ** torch.linspace(0.0, MAX_EXP_XYZ, N_FREQ_XYZ)
        fd = 2 ** torch.linspace(0.0, MAX_EXP_DIR, N_FREQ_DIR)
        return torch.cat([f, f ** 2, fd, torch.zeros(64 - 2 * N_FREQ_XYZ - N_FREQ_DIR)])
    return _table("freqs", make, device)


# ------------------------------------------------------------------------------------------ samplers
def uniform_bins(R, Nc, near, far, device, t_rand=None):
    """nerfstudio UniformSampler (star_mipnerf.py:271): spacing / euclidean frustum edges [R,Nc+1]."""
    lin = _table(("lin", Nc + 1), lambda: torch.linspace(0.0, 1.0, Nc + 1), device)
    spacing = torch.empty((R, Nc + 1), device=device)
    euclid = torch.empty((R, Nc + 1), device=device)
    if t_rand is not None:
        t_rand = _c(t_rand)
    check(_capi.lib().star_mip_uniform_bins(f32(lin), f32(t_rand) if t_rand is not None else None, float(near),
                                            float(far), R, Nc, f32(spacing), f32(euclid), stream()),
          "star_mip_uniform_bins")
    _count()
    return spacing, euclid


def pdf_sample(spacing_bins, weights, Ni, near, far, training=False, u_rand=None, return_details=False):
    """nerfstudio PDFSampler (star_mipnerf.py:286-288): new frustum edges [R,Ni+1] (detached)."""
    spacing_bins = _c(spacing_bins.detach())
    weights = weights.detach()
    if weights.stride(-1) != 1:
        weights = weights.contiguous()
    R, ne = spacing_bins.shape
    Nc = ne - 1
    assert weights.shape == (R, Nc)
    dev = spacing_bins.device
    nb = Ni + 1

    def make():
        u = torch.linspace(0.0, 1.0 - (1.0 / nb), steps=nb)
        return u if training else u + 1.0 / (2 * nb)
    u_base = _table(("u", nb, bool(training)), make, dev)
    if training and u_rand is None:
        u_rand = torch.rand((R, nb), device=dev)
    if not training:
        u_rand = None
    if u_rand is not None:
        u_rand = _c(u_rand)
    sp = torch.empty((R, nb), device=dev)
    eu = torch.empty((R, nb), device=dev)
    det = {}
    if return_details:
        det = dict(inds=torch.empty((R, nb), device=dev, dtype=torch.int64), cdf=torch.empty((R, ne), device=dev))
    check(_capi.lib().star_mip_pdf_sample(f32(spacing_bins), weights.data_ptr(), weights.stride(0), f32(u_base),
                                          f32(u_rand) if u_rand is not None else None, float(near), float(far), R, Nc,
                                          Ni, f32(sp), f32(eu), ptr(det.get("inds")), ptr(det.get("cdf")), stream()),
          "star_mip_pdf_sample")
    _count()
    if return_details:
        return sp, eu, det
    return sp, eu


# ------------------------------------------------------------------------------------------ field
class MipRuntime:
    """Kernel-side state of one mip-NeRF field: flat fp32 master vector (order of include/star_b200.h) and the
    packed weight image, rebuilt when any parameter's version / storage changes."""

    def __init__(self, field):
        self.field = field
        self._key = None
        self._flat = None
        self._packed = {}

    def ordered_params(self):
        f = self.field
        lins = list(f.mlp_base.layers) + [f.field_output_density.net] + list(f.mlp_head.layers) + [f.field_heads[0].net]
        out = []
        for l in lins:
            out += [l.weight, l.bias]
        return out

    def refresh(self, precision):
        params = self.ordered_params()
        key = tuple((p.data_ptr(), p._version) for p in params)
        if key != self._key:
            self._flat = _flat_master(params)
            self._packed = {}
            self._key = key
        if precision not in self._packed:
            L = _capi.lib()
            nbytes = L.star_mip_packed_bytes(precision)
            if nbytes == 0:
                raise _capi.StarError("mip field: precision tier not available")
            assert self._flat.numel() == L.star_mip_param_count()
            packed = torch.empty((nbytes,), device=self._flat.device, dtype=torch.uint8)
            check(L.star_mip_pack_weights(precision, f32(self._flat), f32(freq_table(self._flat.device)), ptr(packed),
                                          stream()), "star_mip_pack_weights")
            _count()
            self._packed[precision] = packed
        return self._flat, self._packed[precision]


_WARNED_FP32_TRAINING = False


class MipFieldRaw(Function):
    """One mip-NeRF field on (origins, directions, frustum edges) with an optional rigid transform of the ray into
    the object frame: RAW (pre-softplus density [R,S], pre-sigmoid rgb [R,S,3])
    (models/mipnerf.py:89-100; star_mipnerf.py:200-260)."""

    @staticmethod
    def forward(ctx, rt, precision, grad_mode, origins, dirs, bins, pose12, *params):
        origins, dirs, bins = _c(origins), _c(dirs), _c(bins.detach())
        R, S = bins.shape[0], bins.shape[1] - 1
        dev = origins.device
        need_grad = grad_mode and (any(ctx.needs_input_grad[7:]) or (pose12 is not None and ctx.needs_input_grad[6]))
        want_pose_grad = pose12 is not None and ctx.needs_input_grad[6] and pose12.requires_grad
        if need_grad and precision != _capi.PREC_F32 and want_pose_grad:
            # the tensor-core backward of the mip field produces weight gradients only: a pass that must differentiate
            # the ray (an object's pose, through the integrated positional encoding) runs the fp32 kernels -- said out
            # loud once
            global _WARNED_FP32_TRAINING
            if not _WARNED_FP32_TRAINING:
                import warnings
                warnings.warn("star_b200 mip field: the 16-bit tensor-core tier has no pose gradient; object fields whose "
                              "pose requires grad run this (and every later) training pass on the fp32 CUDA-core kernels",
                              RuntimeWarning, stacklevel=2)
                _WARNED_FP32_TRAINING = True
            precision = _capi.PREC_F32
        flat, packed = rt.refresh(precision)
        L = _capi.lib()
        freqs = freq_table(dev)
        raw_sigma = torch.empty((R, S), device=dev)
        raw_rgb = torch.empty((R, S, 3), device=dev)
        p12 = _c(pose12.detach()) if pose12 is not None else None
        chunks = _ray_chunks(R, S)
        keep = need_grad and stash_fits(L.star_mip_stash_bytes(precision, R * S), dev)
        stashes = []
        for (a, b) in chunks:
            st = None
            if keep:
                st = stash_alloc(L.star_mip_stash_bytes(precision, (b - a) * S), dev)
                stashes.append(st)
            e0 = _prof_begin()
            check(L.star_mip_field_forward(precision, ptr(packed), f32(origins[a:b]), f32(dirs[a:b]),
                                           f32(p12) if p12 is not None else None, f32(bins[a:b]), f32(freqs),
                                           CONE_RADIUS, b - a, S, f32(raw_sigma[a:b]), f32(raw_rgb[a:b]), S, ptr(st),
                                           stream()), "star_mip_field_forward")
            _prof_end("mip_field_forward_stash" if st is not None else "mip_field_forward", e0, (b - a) * S,
                      2.0 * MIP_MAC_PER_SAMPLE * (b - a) * S)
            _count()
        if need_grad:
            ctx.rt, ctx.precision, ctx.chunks, ctx.stashes = rt, precision, chunks, stashes if keep else None
            ctx.flat, ctx.packed = flat, packed
            ctx.save_for_backward(origins, dirs, bins, p12 if p12 is not None else torch.empty(0, device=dev))
            ctx.shapes = [p.shape for p in params]
        return raw_sigma, raw_rgb

    @staticmethod
    def backward(ctx, g_sigma, g_rgb):
        origins, dirs, bins, p12 = ctx.saved_tensors
        if p12.numel() == 0:
            p12 = None
        precision = ctx.precision
        R, S = bins.shape[0], bins.shape[1] - 1
        dev = origins.device
        L = _capi.lib()
        freqs = freq_table(dev)
        g_sigma = torch.zeros((R, S), device=dev) if g_sigma is None else _c(g_sigma)
        g_rgb = torch.zeros((R, S, 3), device=dev) if g_rgb is None else _c(g_rgb)
        grad_flat = torch.zeros_like(ctx.flat)
        pose_acc = torch.zeros((32,), device=dev) if (p12 is not None and precision == _capi.PREC_F32) else None
        for i, (a, b) in enumerate(ctx.chunks):
            n = (b - a) * S
            if ctx.stashes is not None:
                st = ctx.stashes[i]
            else:   # over the stash budget: re-run the forward of this ray chunk with activation stashing
                st = torch.empty((L.star_mip_stash_bytes(precision, n),), device=dev, dtype=torch.uint8)
                tmp_a = torch.empty((b - a, S), device=dev)
                tmp_c = torch.empty((b - a, S, 3), device=dev)
                check(L.star_mip_field_forward(precision, ptr(ctx.packed), f32(origins[a:b]), f32(dirs[a:b]),
                                               f32(p12) if p12 is not None else None, f32(bins[a:b]), f32(freqs),
                                               CONE_RADIUS, b - a, S, f32(tmp_a), f32(tmp_c), S, ptr(st), stream()),
                      "star_mip_field_forward(recompute)")
                _count()
            ws = torch.empty((L.star_mip_backward_workspace_bytes(precision, n),), device=dev, dtype=torch.uint8)
            e0 = _prof_begin()
            check(L.star_mip_field_backward(precision, ptr(ctx.packed), f32(origins[a:b]), f32(dirs[a:b]),
                                            f32(p12) if p12 is not None else None, f32(bins[a:b]), f32(freqs),
                                            CONE_RADIUS, b - a, S, f32(g_sigma[a:b]), f32(g_rgb[a:b]), S, ptr(st),
                                            ptr(ws), f32(grad_flat), f32(pose_acc) if pose_acc is not None else None,
                                            stream()), "star_mip_field_backward")
            _prof_end("mip_field_backward", e0, n, 4.0 * MIP_MAC_PER_SAMPLE * n)
            _count(15)
            if ctx.stashes is not None:
                ctx.stashes[i] = None
        grads, off = [], 0
        for shp in ctx.shapes:
            n = 1
            for s_ in shp:
                n *= s_
            grads.append(grad_flat[off:off + n].view(shp))
            off += n
        g_pose = None
        if pose_acc is not None and ctx.needs_input_grad[6]:
            # Euclidean gradient w.r.t. [R | t]:  dR = sum g o^T + sum h d^T,  dt = sum g
            g_pose = torch.cat([pose_acc[3:12] + pose_acc[15:24], pose_acc[0:3]])
        return (None, None, None, None, None, None, g_pose, *grads)


# ------------------------------------------------------------------------------------------ compositing
class MipCompositeSingle(Function):
    """get_starmip_appinit_outputs (models/rendering_starmip.py:66-91): rgb [R,3], acc [R], depth [R], weights [R,S]."""

    @staticmethod
    def forward(ctx, raw_sigma, raw_rgb, bins):
        raw_sigma, raw_rgb, bins = _c(raw_sigma), _c(raw_rgb), _c(bins)
        R, S = raw_sigma.shape
        dev = raw_sigma.device
        rgb = torch.empty((R, 3), device=dev)
        acc, depth = torch.empty((R,), device=dev), torch.empty((R,), device=dev)
        weights = torch.empty((R, S), device=dev)
        check(_capi.lib().star_mip_composite_single_forward(f32(raw_sigma), f32(raw_rgb), f32(bins), R, S, f32(rgb),
                                                            f32(acc), f32(depth), f32(weights), stream()),
              "star_mip_composite_single_forward")
        _count()
        ctx.save_for_backward(raw_sigma, raw_rgb, bins)
        ctx.mark_non_differentiable(depth)     # median depth: piecewise constant in the densities
        return rgb, acc, depth, weights

    @staticmethod
    def backward(ctx, g_rgb, g_acc, _g_depth, g_weights):
        raw_sigma, raw_rgb, bins = ctx.saved_tensors
        R, S = raw_sigma.shape
        d_sigma, d_rgb = torch.empty_like(raw_sigma), torch.empty_like(raw_rgb)
        g = [None if t is None else _c(t) for t in (g_rgb, g_acc, g_weights)]
        check(_capi.lib().star_mip_composite_single_backward(
            f32(raw_sigma), f32(raw_rgb), f32(bins), R, S, *[f32(t) if t is not None else None for t in g],
            f32(d_sigma), f32(d_rgb), stream()), "star_mip_composite_single_backward")
        _count()
        return d_sigma, d_rgb, None


MIP_OUT_KEYS = ("rgb", "acc", "depth", "weights", "rgb_static", "depth_static", "rgb_dynamic", "depth_dynamic",
                "dynamic_transmittance", "regs")


class MipCompositeStar(Function):
    """get_starmip_online_outputs (models/rendering_starmip.py:112-175) -> the tensors of MIP_OUT_KEYS."""

    @staticmethod
    def forward(ctx, rs_s, rc_s, rs_d, rc_d, bins, chunk):
        rs_s, rc_s, rs_d, rc_d, bins = (_c(t) for t in (rs_s, rc_s, rs_d, rc_d, bins))
        R, V, S = rs_d.shape
        dev = rs_s.device
        e = lambda *s: torch.empty(s, device=dev, dtype=torch.float32)
        out = dict(rgb=e(R, 3), acc=e(R), depth=e(R), weights=e(R, S), rgb_static=e(R, 3), depth_static=e(R),
                   rgb_dynamic=e(R, V, 3), depth_dynamic=e(R, V), dynamic_transmittance=e(R, V), regs=e(5))
        L = _capi.lib()
        ws = torch.empty((L.star_mip_composite_multi_ws_bytes(R) // 4,), device=dev, dtype=torch.float32)
        mo = _capi.StarMipMultiOut(*[ptr(out[k]) for k in MIP_OUT_KEYS])
        check(L.star_mip_composite_multi_forward(f32(rs_s), f32(rc_s), f32(rs_d), f32(rc_d), f32(bins), R, V, S,
                                                 int(chunk), C.byref(mo), ptr(ws), stream()),
              "star_mip_composite_multi_forward")
        _count(2)
        ctx.save_for_backward(rs_s, rc_s, rs_d, rc_d, bins)
        ctx.chunk = int(chunk)
        ctx.mark_non_differentiable(out["depth"], out["rgb_static"], out["depth_static"], out["rgb_dynamic"],
                                    out["depth_dynamic"], out["dynamic_transmittance"])
        return tuple(out[k] for k in MIP_OUT_KEYS)

    @staticmethod
    def backward(ctx, g_rgb, g_acc, _d, g_weights, _a, _b, _c2, _e, _f, g_regs):
        rs_s, rc_s, rs_d, rc_d, bins = ctx.saved_tensors
        R, V, S = rs_d.shape
        d = [torch.empty_like(t) for t in (rs_s, rc_s, rs_d, rc_d)]
        g = [None if t is None else _c(t) for t in (g_rgb, g_acc, g_weights, g_regs)]
        check(_capi.lib().star_mip_composite_multi_backward(
            f32(rs_s), f32(rc_s), f32(rs_d), f32(rc_d), f32(bins), R, V, S, ctx.chunk,
            *[f32(t) if t is not None else None for t in g], *[f32(t) for t in d], stream()),
            "star_mip_composite_multi_backward")
        _count()
        return d[0], d[1], d[2], d[3], None, None

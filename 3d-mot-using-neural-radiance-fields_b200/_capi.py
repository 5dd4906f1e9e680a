"""ctypes binding of libstar_b200.so (include/star_b200.h).

Fails loudly when the library is missing or a call returns a nonzero status: there is no CPU or
PyTorch fallback for any entry point (north star: "no CPU fallback").
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# STAR_B200_LIB: another build of the same ABI (A/B measurements of kernel variants on one box)
LIB_PATH = os.environ.get("STAR_B200_LIB") or os.path.join(_HERE, "libstar_b200.so")

ABI_VERSION = 2
PREC_F32, PREC_BF16, PREC_F16 = 0, 1, 2
PREC_FLAG_RETIRED = 0x300      # round-2 A/B variants of the forward kernels (CTA pair, direct stash stores): rejected
PREC_FLAG_DX_PIPELINED = 0x400
PREC_FLAG_NO_WSHARE = 0x800     # inference forward: every CTA streams its own weights (default: cluster of 2 shares them)
PRECISIONS = {"fp32": PREC_F32, "bf16": PREC_BF16, "fp16": PREC_F16}

c_f = C.c_void_p      # device pointers are passed as raw addresses
c_i64 = C.c_int64


class StarNetDesc(C.Structure):
    _fields_ = [("n_blocks", C.c_int32), ("L_xyz", C.c_int32), ("L_dir", C.c_int32), ("precision", C.c_int32)]


class StarMipMultiOut(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "rgb", "acc", "depth", "weights", "rgb_static", "depth_static", "rgb_dynamic", "depth_dynamic",
        "dynamic_transmittance", "regs")]


class StarMultiOut(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "rgb", "disp", "acc", "depth", "weights", "rgb_static", "depth_static", "rgb_dynamic",
        "depth_dynamic", "dynamic_transmittance", "rgb_dynamic_all", "regs")]


class StarRenderCfg(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("R", "Nc", "Ni", "V", "precision", "n_blocks_static", "n_blocks_dynamic", "L_xyz",
                                         "L_dir", "white_bkgd", "lindisp", "test", "chunk")] + \
               [(n, C.c_float) for n in ("near_", "far_", "far_dist")]


class StarRenderIn(C.Structure):
    _fields_ = [("rays_o", C.c_void_p), ("rays_d", C.c_void_p), ("viewdirs", C.c_void_p),
                ("H", C.c_int32), ("W", C.c_int32), ("row0", C.c_int32), ("nrows", C.c_int32),
                ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float), ("c2w", C.c_void_p),
                ("z_vals", C.c_void_p), ("pts", C.c_void_p), ("t_vals", C.c_void_p), ("t_rand", C.c_void_p),
                ("u", C.c_void_p), ("u_det", C.c_void_p), ("z_samples", C.c_void_p),
                ("pose12", C.c_void_p), ("enc_scale_xyz", C.c_void_p), ("enc_scale_dir", C.c_void_p),
                ("packed_static_coarse", C.c_void_p), ("packed_static_fine", C.c_void_p),
                ("packed_dynamic_coarse", C.POINTER(C.c_void_p)), ("packed_dynamic_fine", C.POINTER(C.c_void_p))]


class StarRenderOut(C.Structure):
    _fields_ = [("coarse", StarMultiOut), ("fine", StarMultiOut)] + \
               [(n, C.c_void_p) for n in ("dists0", "dists", "z_vals0", "z_vals", "z_samples", "z_std", "rays_o", "rays_d",
                                          "viewdirs")]


class StarAdamSeg(C.Structure):
    _fields_ = [("param", C.c_void_p), ("grad", C.c_void_p), ("exp_avg", C.c_void_p), ("exp_avg_sq", C.c_void_p),
                ("n", C.c_int64), ("step_size", C.c_float), ("bc2_sqrt", C.c_float)]


_SIGS = {
    "star_abi_version": (C.c_int, []),
    "star_error_string": (C.c_char_p, [C.c_int]),
    "star_last_cuda_error": (C.c_int, []),
    "star_watchdog_word": (C.c_int, [C.c_int]),
    "star_net_param_count": (C.c_size_t, [C.POINTER(StarNetDesc)]),
    "star_packed_bytes": (C.c_size_t, [C.POINTER(StarNetDesc)]),
    "star_pack_weights": (C.c_int, [C.POINTER(StarNetDesc), c_f, c_f, c_f]),
    "star_sample_pts": (C.c_int, [c_f, c_f, c_f, c_f, C.c_float, C.c_float, C.c_int, C.c_int, C.c_int, c_f, c_f, c_f]),
    "star_get_rays": (C.c_int, [C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, c_f, C.c_int, C.c_int,
                                c_f, c_f, c_f, c_f]),
    "star_embed": (C.c_int, [c_f, C.c_int, C.c_int, c_f, c_f, c_f]),
    "star_stash_bytes": (C.c_size_t, [C.POINTER(StarNetDesc), c_i64]),
    "star_mlp_forward": (C.c_int, [C.POINTER(StarNetDesc), c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_f, C.c_int,
                                   C.c_int, c_f, c_f, c_i64, c_f, c_f, c_f]),
    "star_mlp_backward_workspace_bytes": (C.c_size_t, [C.POINTER(StarNetDesc), c_i64]),
    "star_mlp_backward": (C.c_int, [C.POINTER(StarNetDesc), c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_f, C.c_int,
                                    C.c_int, c_f, c_f, c_i64, c_f, c_f, c_f, c_f, c_f]),
    "star_composite_single_forward": (C.c_int, [c_f, c_f, c_f, c_f, C.c_int, C.c_int, C.c_float, C.c_int, c_f, c_f,
                                                c_f, c_f, c_f, c_f, c_f]),
    "star_composite_single_backward": (C.c_int, [c_f, c_f, c_f, c_f, C.c_int, C.c_int, C.c_float, C.c_int, c_f, c_f,
                                                 c_f, c_f, c_f, c_f, c_f, c_f]),
    "star_composite_multi_ws_bytes": (C.c_size_t, [C.c_int]),
    "star_composite_multi_forward": (C.c_int, [c_f, c_f, c_f, c_f, c_f, c_f, C.c_int, C.c_int, C.c_int, C.c_float,
                                               C.c_int, C.c_int, C.POINTER(StarMultiOut), c_f, c_f]),
    "star_composite_multi_backward": (C.c_int, [c_f, c_f, c_f, c_f, c_f, c_f, C.c_int, C.c_int, C.c_int, C.c_float,
                                                C.c_int, C.c_int, c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_f,
                                                c_f]),
    "star_sample_pdf": (C.c_int, [c_f, c_i64, c_f, c_i64, c_f, c_f, C.c_int, C.c_int, C.c_int, c_f, c_f, c_f, c_f,
                                  c_f, c_f]),
    "star_invert_cdf": (C.c_int, [c_f, c_f, c_f, C.c_int, C.c_int, C.c_int, c_f, c_f, c_f, c_f, c_f]),
    "star_hierarchical": (C.c_int, [c_f, c_f, c_f, c_f, c_f, c_f, C.c_int, C.c_int, C.c_int, c_f, c_f, c_f, c_f, c_f]),
    "star_merge_samples": (C.c_int, [c_f, c_f, c_f, c_f, C.c_int, C.c_int, C.c_int, c_f, c_f, c_f, c_f]),
    "star_composite_hier_forward": (C.c_int, [c_f, c_f, c_f, c_f, c_f, c_f, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int,
                                              c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_f]),
    "star_render_workspace_bytes": (C.c_size_t, [C.POINTER(StarRenderCfg)]),
    "star_render_forward": (C.c_int, [C.POINTER(StarRenderCfg), C.POINTER(StarRenderIn), C.POINTER(StarRenderOut), c_f,
                                      C.c_size_t, c_f, c_f]),
    # a12: mip-NeRF variant
    "star_mip_uniform_bins": (C.c_int, [c_f, c_f, C.c_float, C.c_float, C.c_int, C.c_int, c_f, c_f, c_f]),
    "star_mip_pdf_sample": (C.c_int, [c_f, c_f, c_i64, c_f, c_f, C.c_float, C.c_float, C.c_int, C.c_int, C.c_int,
                                      c_f, c_f, c_f, c_f, c_f]),
    "star_mip_param_count": (C.c_size_t, []),
    "star_mip_packed_bytes": (C.c_size_t, [C.c_int]),
    "star_mip_pack_weights": (C.c_int, [C.c_int, c_f, c_f, c_f, c_f]),
    "star_mip_stash_bytes": (C.c_size_t, [C.c_int, c_i64]),
    "star_mip_backward_workspace_bytes": (C.c_size_t, [C.c_int, c_i64]),
    "star_mip_field_forward": (C.c_int, [C.c_int, c_f, c_f, c_f, c_f, c_f, c_f, C.c_float, C.c_int, C.c_int, c_f, c_f,
                                         c_i64, c_f, c_f]),
    "star_mip_field_backward": (C.c_int, [C.c_int, c_f, c_f, c_f, c_f, c_f, c_f, C.c_float, C.c_int, C.c_int, c_f,
                                          c_f, c_i64, c_f, c_f, c_f, c_f, c_f]),
    "star_mip_composite_single_forward": (C.c_int, [c_f, c_f, c_f, C.c_int, C.c_int, c_f, c_f, c_f, c_f, c_f]),
    "star_mip_composite_single_backward": (C.c_int, [c_f, c_f, c_f, C.c_int, C.c_int, c_f, c_f, c_f, c_f, c_f, c_f]),
    "star_mip_composite_multi_ws_bytes": (C.c_size_t, [C.c_int]),
    "star_mip_composite_multi_forward": (C.c_int, [c_f, c_f, c_f, c_f, c_f, C.c_int, C.c_int, C.c_int, C.c_int,
                                                   C.POINTER(StarMipMultiOut), c_f, c_f]),
    "star_mip_composite_multi_backward": (C.c_int, [c_f, c_f, c_f, c_f, c_f, C.c_int, C.c_int, C.c_int, C.c_int,
                                                    c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_f]),
    # losses and the optimiser step
    "star_train_ws_bytes": (C.c_size_t, []),
    "star_photometric_loss": (C.c_int, [c_f, c_f, c_f, c_i64, c_f, c_f, c_f, c_f, c_f]),
    "star_depth_loss_forward": (C.c_int, [c_f, c_f, c_i64, C.c_float, C.c_float, c_f, c_f, c_f]),
    "star_depth_loss_backward": (C.c_int, [c_f, c_f, c_i64, C.c_float, C.c_float, c_f, c_f, c_f, c_f]),
    "star_sigma_loss_forward": (C.c_int, [c_f, c_f, c_f, c_f, c_i64, C.c_int, C.c_float, C.c_float, C.c_float, c_f, c_f,
                                          c_f, c_f]),
    "star_sigma_loss_backward": (C.c_int, [c_f, c_f, c_f, c_f, c_i64, C.c_int, C.c_float, C.c_float, C.c_float, c_f,
                                           c_f, C.c_int, c_f, c_f]),
    "star_grad_sqnorm": (C.c_int, [C.POINTER(StarAdamSeg), C.c_int, c_f, c_f]),
    "star_grad_sqnorm_result": (C.c_void_p, [c_f]),
    "star_grad_scale": (C.c_int, [C.POINTER(StarAdamSeg), C.c_int, c_f, C.c_float, c_f]),
    "star_iou2d": (C.c_int, [c_f, c_f, c_i64, C.c_int, C.c_float, c_f, c_f, c_f]),
    "star_adam_step": (C.c_int, [C.POINTER(StarAdamSeg), C.c_int, C.c_double, C.c_double, C.c_double, c_f, C.c_float,
                                 C.c_int, c_f]),
}

EXPORTED_SYMBOLS = tuple(_SIGS)

_lib = None


class StarError(RuntimeError):
    pass


class _NvtxLib:
    """STAR_B200_NVTX=1: every kernel-launching C-ABI call (those that take a stream) runs inside an NVTX range named
    after the entry point, so that nsys / ncu timelines show the render path stage by stage."""

    def __init__(self, L):
        self._L = L

    def __getattr__(self, name):
        fn = getattr(self._L, name)
        sig = _SIGS.get(name)
        if sig is None or not sig[1] or sig[1][-1] is not c_f or sig[0] is not C.c_int:
            return fn

        def wrapped(*a):
            torch.cuda.nvtx.range_push(name)
            try:
                return fn(*a)
            finally:
                torch.cuda.nvtx.range_pop()
        object.__setattr__(self, name, wrapped)
        return wrapped


def lib():
    """Loads libstar_b200.so once.  Missing library -> hard error (build with __graft_entry__.build())."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise StarError("libstar_b200.so is not built (%s); run `python -c 'import __graft_entry__ as g; "
                            "g.build()'` -- there is no fallback path" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        if L.star_abi_version() != ABI_VERSION:
            raise StarError("libstar_b200.so ABI mismatch")
        if os.environ.get("STAR_B200_NVTX", "0") == "1":
            L = _NvtxLib(L)
        _lib = L
    return _lib


def check(code, what):
    if code != 0:
        L = lib()
        msg = L.star_error_string(code).decode()
        extra = ""
        if code == 6:
            extra = " (cudaError %d)" % L.star_last_cuda_error()
            wd = watchdog_report()
            if wd:
                extra += " [kernel watchdog: %s]" % wd
        raise StarError("%s failed: %s%s" % (what, msg, extra))


WATCHDOG_FAMILIES = ("mlp forward", "dX chain", "dW", "pipelined dX", "mip forward", "mip dX")


def watchdog_report():
    """Text for the tensor-core kernels' barrier watchdogs that fired ('' if none): after a CUDA launch failure this says
    which kernel family hung on which wait (include/star_b200.h, star_watchdog_word)."""
    L = lib()
    out = []
    for i, name in enumerate(WATCHDOG_FAMILIES):
        w = L.star_watchdog_word(i)
        if w:
            out.append("%s: wait code %d in CTA %d" % (name, w >> 16, w & 0xffff))
        b, e = L.star_watchdog_word(16 + i), L.star_watchdog_word(32 + i)
        if b != e:
            out.append("%s: launch %d started and did not finish (%d finished)" % (name, b, e))
    return "; ".join(out)


def launch_markers():
    """(started, finished) launch counts per tensor-core kernel family (diagnostics)."""
    L = lib()
    return {name: (L.star_watchdog_word(16 + i), L.star_watchdog_word(32 + i)) for i, name in enumerate(WATCHDOG_FAMILIES)}


def ptr(t):
    """Device address of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise StarError("expected a CUDA tensor: the STaR B200 path has no CPU implementation")
    if not t.is_contiguous():
        raise StarError("expected a contiguous tensor")
    return t.data_ptr()


def f32(t):
    if t.dtype != torch.float32:
        raise StarError("expected float32, got %s" % t.dtype)
    return ptr(t)


def stream():
    return torch.cuda.current_stream().cuda_stream


def net_desc(n_blocks, L_xyz, L_dir, precision):
    return StarNetDesc(int(n_blocks), int(L_xyz), int(L_dir), int(precision))

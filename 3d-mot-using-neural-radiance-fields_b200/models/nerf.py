"""NeRF radiance field module (mirror of class NeRF, models/nerf.py:34-191 of the reference).

Owns the fp32 master parameters in the reference state_dict layout; `forward` runs the fused
pose-transform + encoding + ResNet-FC + heads CUDA kernel through the C ABI and returns the RAW
(raw_alpha, raw_rgb) pair, or the composited tuple when z_vals is given."""
import os

import torch
from torch import nn

from .. import functional as F_
from .. import _capi
from .embedder import get_embedder
from .resnet import ResnetFC


def default_precision():
    """'fp32' (1e-4 tier, CUDA cores), 'fp16' (tcgen05 tensor cores, fp16 operands: the 2e-3 tier) or 'bf16' (same
    kernels, bf16 operands: wider range, 9e-3 worst case on random-init nets)."""
    return _capi.PRECISIONS[os.environ.get("STAR_B200_PRECISION", "fp32")]


class NeRF(nn.Module):
    def __init__(self, D, W, args, has_time=False, more_view_layers=False):
        super().__init__()
        if has_time or more_view_layers or not args.use_viewdirs or W != 256 or args.i_embed == -1:
            raise NotImplementedError("B200 path: W=256, use_viewdirs=True, 3-D inputs, single view layer only")
        self.D, self.W = D, W
        self.embedder, self.input_ch = get_embedder(args.multires, args.end_barf, args.i_embed)
        self.embedder_dirs, self.input_ch_views = get_embedder(args.multires_views, args.end_barf, args.i_embed)
        self.use_viewdirs = True
        self.pts_net = ResnetFC(self.input_ch, d_out=W, n_blocks=D // 2, d_hidden=W)
        self.views_linears = nn.ModuleList([nn.Linear(self.input_ch_views + W, W // 2)])
        self.feature_linear = nn.Linear(W, W)
        self.alpha_linear = nn.Linear(W, 1)
        self.rgb_linear = nn.Linear(W // 2, 3)
        self.netchunk = args.netchunk            # kept for API parity; the kernels tile internally
        self.raw_noise_std = args.raw_noise_std
        self.white_bkgd = args.white_bkgd
        for layer in self.views_linears:
            nn.init.kaiming_normal_(layer.weight, nonlinearity="relu")
            nn.init.zeros_(layer.bias)
        nn.init.kaiming_normal_(self.alpha_linear.weight, nonlinearity="relu")
        nn.init.zeros_(self.alpha_linear.bias)
        nn.init.xavier_uniform_(self.rgb_linear.weight)
        self.precision = None                    # None -> STAR_B200_PRECISION env / fp32
        self._rt = F_.NetRuntime(self, D // 2, self.embedder.L, self.embedder_dirs.L)

    def _prec(self):
        return default_precision() if self.precision is None else self.precision

    def raw(self, pts, viewdirs, pose12=None, step=None):
        """RAW outputs; pose12 moves samples into the object frame inside the kernel (K2).
        pts: [R,S,3], or the tuple (rays_o, rays_d, z_vals): positions formed inside the kernel."""
        sc_xyz = self.embedder.scale(step, viewdirs.device, pad_to=64)
        sc_dir = self.embedder_dirs.scale(step, viewdirs.device, pad_to=32)
        return F_.NerfRaw.apply(self._rt, self._prec(), torch.is_grad_enabled(), pts, viewdirs, pose12, sc_xyz, sc_dir,
                                *self._rt.ordered_params())

    def forward(self, pts, viewdirs, z_vals=None, rays_d=None, step=None, time=None):
        if time is not None:
            raise NotImplementedError("time-conditioned NeRF is outside the B200 hot path")
        raw_alpha, raw_rgb = self.raw(pts, viewdirs, None, step)
        if z_vals is None:
            return raw_alpha, raw_rgb
        from .rendering__ import raw2outputs
        o = raw2outputs(raw_alpha, raw_rgb, z_vals, rays_d, self.raw_noise_std if self.training else 0,
                        self.white_bkgd, 1e10)
        return o["rgb"], o["disp"], o["acc"], o["weights"], o["depth"]

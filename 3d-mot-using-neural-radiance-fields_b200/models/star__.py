"""STaR compositional field (mirror of models/star__.py:24-225 of the reference): a static NeRF plus
`num_vehicles` rigid dynamic-object NeRFs, coarse and fine, with the reference's module tree (and
therefore checkpoint layout):  static_{coarse,fine}_nerf, dynamic_{coarse,fine}_nerfs.{i}."""
import torch
from torch import nn

from .. import functional as F_
from .nerf import NeRF
from .rendering__ import raw2outputs, raw2outputs_star


class STaR(nn.Module):
    def __init__(self, args):
        super().__init__()
        self.num_vehicles = args.num_vehicles
        self.chunk = args.chunk
        self.far_dist = args.far_dist
        self.N_importance = args.N_importance
        self.static_coarse_nerf = NeRF(D=args.netdepth, W=args.netwidth, args=args)
        if args.N_importance > 0:
            self.static_fine_nerf = NeRF(D=args.netdepth_fine, W=args.netwidth_fine, args=args)
        self.dynamic_coarse_nerfs = nn.ModuleList(
            [NeRF(D=args.netdepth // 2, W=args.netwidth, args=args) for _ in range(self.num_vehicles)])
        if args.N_importance > 0:
            self.dynamic_fine_nerfs = nn.ModuleList(
                [NeRF(D=args.netdepth_fine // 2, W=args.netwidth_fine, args=args) for _ in range(self.num_vehicles)])

    def get_nerf_params(self):
        return (list(self.static_coarse_nerf.parameters()) + list(self.static_fine_nerf.parameters())
                + list(self.dynamic_coarse_nerfs.parameters()) + list(self.dynamic_fine_nerfs.parameters()))

    def set_precision(self, precision):
        """'fp32' | 'bf16' | 'fp16' for every field MLP."""
        from .. import _capi
        p = _capi.PRECISIONS[precision]
        for m in self.modules():
            if isinstance(m, NeRF):
                m.precision = p

    def forward(self, pts, viewdirs, z_vals, rays_d, pose=None, is_coarse=True, object_pose=None, step=None):
        """(:68-116).  The reference walks ray chunks of `self.chunk` in Python; here one launch group
        covers all rays and `chunk` only enters the regulariser normalisation, which reproduces
        'mean within a chunk, summed over chunks' (:111-112) exactly."""
        if is_coarse:
            static_model, dynamic_models = self.static_coarse_nerf, self.dynamic_coarse_nerfs
        else:
            if self.N_importance <= 0:
                raise ValueError("N_importance should be positive")
            static_model, dynamic_models = self.static_fine_nerf, self.dynamic_fine_nerfs

        raw_alpha_s, raw_rgb_s = static_model.raw(pts, viewdirs, None, step=None)      # :144 (never BARF)
        if pose is None and object_pose is None:
            return raw2outputs(raw_alpha_s, raw_rgb_s, z_vals, rays_d,
                               static_model.raw_noise_std if self.training else 0, static_model.white_bkgd,
                               far_dist=self.far_dist)
        if object_pose is not None or pose.dim() not in (2, 3):
            raise NotImplementedError
        ra_d, rc_d = [], []
        for i, model in enumerate(dynamic_models):                                       # :201-210
            p12 = F_.pose_to_mat12(pose[i])
            a, c = model.raw(pts, viewdirs, p12, step=step)
            ra_d.append(a)
            rc_d.append(c)
        raw_alpha_d = torch.stack(ra_d, 1)
        raw_rgb_d = torch.stack(rc_d, 1)
        return raw2outputs_star(raw_alpha_s, raw_rgb_s, raw_alpha_d, raw_rgb_d, z_vals, rays_d, 0,
                                static_model.white_bkgd, far_dist=self.far_dist, test=not self.training,
                                chunk=self.chunk)

    # the reference exposes forward_chunk too; one chunk == the whole call here
    def forward_chunk(self, pts, viewdirs, z_vals, rays_d, pose=None, is_coarse=True, object_pose=None, step=None):
        saved, self.chunk = self.chunk, max(pts.shape[0], 1)
        try:
            return self.forward(pts, viewdirs, z_vals, rays_d, pose, is_coarse, object_pose, step)
        finally:
            self.chunk = saved

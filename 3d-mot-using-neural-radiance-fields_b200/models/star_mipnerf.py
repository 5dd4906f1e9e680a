"""STaR with mip-NeRF fields (mirror of models/star_mipnerf.py:41-357 of the reference): one field per scene
element shared by the coarse and the fine pass, module tree `static_nerf`, `dynamic_nerfs.{i}` (checkpoint layout),
forward(origins, viewdirs, pose=None) -> StarMipAppInitCombinedOutput / StarMipOnlineCombinedOutput.

Keyword-only extras `t_rand=` / `u_rand=` inject the random draws of nerfstudio's samplers in training mode."""
import torch
from torch import nn

from .. import functional as F_
from .. import mip_functional as MF
from .mipnerf import MipNerfModel
from .rendering_starmip import get_starmip_appinit_outputs, get_starmip_online_outputs


class STaR(nn.Module):
    def __init__(self, args):
        super().__init__()
        self.num_vehicles = args.num_vehicles
        self.chunk = args.chunk
        self.far_dist = args.far_dist
        self.N_importance = args.N_importance
        self.N_samples = args.N_samples
        self.static_nerf = MipNerfModel()
        self.dynamic_nerfs = nn.ModuleList([MipNerfModel() for _ in range(self.num_vehicles)])
        # NearFarCollider (:83-86)
        self.near_plane = args.scale_factor * args.near
        self.far_plane = args.scale_factor * args.far

    def get_nerf_params(self):
        return list(self.static_nerf.parameters()) + list(self.dynamic_nerfs.parameters())

    def set_precision(self, precision):
        from .. import _capi
        for m in [self.static_nerf] + list(self.dynamic_nerfs):
            m.precision = _capi.PRECISIONS[precision]

    def _pass(self, origins, viewdirs, pose, euclid):
        rs_s, rc_s = self.static_nerf.raw(origins, viewdirs, euclid)
        if pose is None:
            return get_starmip_appinit_outputs(rs_s, rc_s, euclid)
        rs_d, rc_d = [], []
        for i, model in enumerate(self.dynamic_nerfs):                      # :200-260
            a, c = model.raw(origins, viewdirs, euclid, F_.pose_to_mat12(pose[i]))
            rs_d.append(a)
            rc_d.append(c)
        return get_starmip_online_outputs(rs_s, rc_s, torch.stack(rs_d, 1), torch.stack(rc_d, 1), euclid,
                                          chunk=self.chunk)

    def forward(self, origins, viewdirs, pose=None, *, t_rand=None, u_rand=None):
        """(:99-137, :262-357).  The reference walks ray chunks of `self.chunk` in Python; here one launch group
        covers all rays and `chunk` only enters the regulariser normalisation ('mean within a chunk, summed over
        chunks')."""
        if pose is not None and pose.dim() != 2:
            raise NotImplementedError          # :195-198: the mip variant supports 7-vector poses only
        R = origins.shape[0]
        dev = origins.device
        near, far = self.near_plane, self.far_plane
        if self.training and t_rand is None:
            t_rand = torch.rand((R, self.N_samples + 1), device=dev)
        spacing, euclid = MF.uniform_bins(R, self.N_samples, near, far, dev, t_rand if self.training else None)
        coarse = self._pass(origins, viewdirs, pose, euclid)
        sp2, eu2 = MF.pdf_sample(spacing, coarse["weights"][..., 0], self.N_importance, near, far,
                                 training=self.training, u_rand=u_rand)
        fine = self._pass(origins, viewdirs, pose, eu2)
        result = dict(fine)
        for k, v in coarse.items():
            result[f"{k}0"] = v
        return result

    def forward_chunk(self, origins, viewdirs, pose=None, **kw):
        saved, self.chunk = self.chunk, max(origins.shape[0], 1)
        try:
            return self.forward(origins, viewdirs, pose, **kw)
        finally:
            self.chunk = saved

"""mip-NeRF field modules (mirror of models/mipnerf.py:30-100 of the reference, which instantiates nerfstudio's
NeRFField): the module tree -- and therefore the checkpoint keys `field.mlp_base.layers.{i}`,
`field.field_output_density.net`, `field.mlp_head.layers.{i}`, `field.field_heads.0.net` -- follows nerfstudio's;
the arithmetic runs in the fused sm_100a kernel (csrc/mip_f32.cu) through the C ABI."""
import os

import torch
from torch import nn

from .. import _capi
from .. import mip_functional as MF

D_XYZ = 3 * MF.N_FREQ_XYZ * 2 + 3     # 147 (models/mipnerf.py:58-64)
D_DIR = 3 * MF.N_FREQ_DIR * 2 + 3     # 27  (models/mipnerf.py:65-71)


class _MLP(nn.Module):
    """nerfstudio field_components.mlp.MLP: `layers` ModuleList of nn.Linear (parameter container only)."""

    def __init__(self, dims):
        super().__init__()
        self.layers = nn.ModuleList([nn.Linear(i, o) for i, o in dims])


class _Head(nn.Module):
    """nerfstudio FieldHead: `net` = nn.Linear (parameter container only)."""

    def __init__(self, i, o):
        super().__init__()
        self.net = nn.Linear(i, o)


class NeRFField(nn.Module):
    """nerfstudio NeRFField(position_encoding F=24, direction_encoding F=4, use_integrated_encoding=True,
    base 8 x 256 with skip (4,), head 2 x 128, RGBFieldHead)."""

    def __init__(self, W=256, WH=128, n_base=8, skip=4):
        super().__init__()
        self.mlp_base = _MLP([(D_XYZ if i == 0 else (W + D_XYZ if i == skip else W), W) for i in range(n_base)])
        self.field_output_density = _Head(W, 1)
        self.mlp_head = _MLP([(W + D_DIR, WH), (WH, WH)])
        self.field_heads = nn.ModuleList([_Head(WH, 3)])


class MipNerfModel(nn.Module):
    """Mirror of MipNerfModel (models/mipnerf.py:30-100): `.field`, get_outputs() -> (density, rgb)."""

    def __init__(self, config=None, **kwargs):
        super().__init__()
        self.config = config
        self.field = NeRFField()
        self.precision = None
        self._rt = MF.MipRuntime(self.field)

    def _prec(self):
        if self.precision is not None:
            return self.precision
        return _capi.PRECISIONS[os.environ.get("STAR_B200_MIP_PRECISION", "fp32")]

    def get_param_groups(self):
        return {"fields": list(self.field.parameters())}

    def raw(self, origins, directions, bins, pose12=None):
        """RAW (pre-activation) density [R,S] and rgb [R,S,3] for frustum edges `bins` [R,S+1]."""
        return MF.MipFieldRaw.apply(self._rt, self._prec(), torch.is_grad_enabled(), origins, directions, bins, pose12,
                                    *self._rt.ordered_params())

    def get_outputs(self, origins, directions, bins):
        """(models/mipnerf.py:89-100) densities [R,S,1] (softplus) and rgb [R,S,3] (sigmoid).  The reference passes a
        nerfstudio RaySamples; here the frustums are (origins, directions, edges)."""
        raw_sigma, raw_rgb = self.raw(origins, directions, bins)
        return torch.nn.functional.softplus(raw_sigma)[..., None], torch.sigmoid(raw_rgb)

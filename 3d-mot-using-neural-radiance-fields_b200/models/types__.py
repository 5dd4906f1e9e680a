"""Output-dict contracts of the render path (mirrors models/types__.py:5-111 of the reference; plain
typing instead of torchtyping, which is not a dependency of this package)."""
from typing import Optional, TypedDict, Union

from torch import Tensor


class NerfNetworkOutput(TypedDict):
    rgb: Tensor          # [R, 3]
    disp: Tensor         # [R]
    acc: Tensor          # [R]
    weights: Tensor      # [R, S]
    depth: Tensor        # [R]
    dists: Tensor        # [R, S]
    z_vals: Tensor       # [R, S]


class StarNetworkOutput(TypedDict):
    rgb: Tensor                     # [R, 3]
    disp: Tensor                    # [R]
    acc: Tensor                     # [R]
    dynamic_transmittance: Tensor   # [R, V]
    weights: Tensor                 # [R, S]
    depth: Tensor                   # [R]
    rgb_static: Tensor              # [R, 3]
    depth_static: Tensor            # [R]
    rgb_dynamic: Tensor             # [R, V, 3]
    rgb_dynamic_all: Optional[Tensor]   # [R, 3], eval mode only
    depth_dynamic: Tensor           # [R, V]
    loss_alpha_entropy: Tensor      # []
    loss_dynamic_vs_static_reg: Tensor
    loss_ray_reg: Tensor
    loss_static_reg: Tensor
    loss_dynamic_reg: Tensor


NERF_KEYS = ("rgb", "disp", "acc", "weights", "depth", "dists", "z_vals")
STAR_KEYS = ("rgb", "disp", "acc", "weights", "depth", "rgb_static", "rgb_dynamic", "depth_static",
             "depth_dynamic", "dynamic_transmittance", "loss_alpha_entropy", "loss_dynamic_vs_static_reg",
             "loss_ray_reg", "loss_static_reg", "loss_dynamic_reg", "rgb_dynamic_all")

StarRenderOutput = dict
NetworkOutput = Union[NerfNetworkOutput, StarNetworkOutput]

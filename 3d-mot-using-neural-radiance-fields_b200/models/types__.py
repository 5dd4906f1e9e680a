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


# ---- mip variant (models/types__.py:36-87 of the reference): per-ray tensors carry a trailing singleton dimension
class StarMipOnlineOutput(TypedDict):
    rgb: Tensor                     # [R, 3]
    acc: Tensor                     # [R, 1]
    weights: Tensor                 # [R, S, 1]   (S = N_samples + N_importance... the fine pass has N_importance frustums)
    depth: Tensor                   # [R, 1]
    rgb_static: Tensor              # [R, 3]
    depth_static: Tensor            # [R, 1]
    rgb_dynamic: Tensor             # [R, V, 3]
    depth_dynamic: Tensor           # [R, V]
    dynamic_transmittance: Tensor   # [R, V, 1]
    loss_alpha_entropy: Tensor      # []
    loss_dynamic_vs_static_reg: Tensor
    loss_ray_reg: Tensor
    loss_static_reg: Tensor
    loss_dynamic_reg: Tensor


class StarMipOnlineCombinedOutput(StarMipOnlineOutput):
    rgb0: Tensor
    acc0: Tensor
    weights0: Tensor                # [R, N_samples, 1]
    depth0: Tensor
    rgb_static0: Tensor
    depth_static0: Tensor
    rgb_dynamic0: Tensor
    depth_dynamic0: Tensor
    dynamic_transmittance0: Tensor
    loss_alpha_entropy0: Tensor
    loss_dynamic_vs_static_reg0: Tensor
    loss_ray_reg0: Tensor
    loss_static_reg0: Tensor
    loss_dynamic_reg0: Tensor


class StarMipAppInitOutput(TypedDict):
    rgb: Tensor                     # [R, 3]
    acc: Tensor                     # [R, 1]
    weights: Tensor                 # [R, S, 1]
    depth: Tensor                   # [R, 1]


class StarMipAppInitCombinedOutput(StarMipAppInitOutput):
    rgb0: Tensor
    acc0: Tensor
    weights0: Tensor
    depth0: Tensor


class StarCoarseNetworkOutput(TypedDict):          # (:89-107) N_importance <= 0: only the "...0" keys
    rgb0: Tensor
    disp0: Tensor
    acc0: Tensor
    dynamic_transmittance0: Tensor
    weights0: Tensor
    depth0: Tensor
    rgb_static0: Tensor
    depth_static0: Tensor
    rgb_dynamic0: Tensor
    rgb_dynamic_all0: Optional[Tensor]
    depth_dynamic0: Tensor
    loss_alpha_entropy0: Tensor
    loss_dynamic_vs_static_reg0: Tensor
    loss_ray_reg0: Tensor
    loss_static_reg0: Tensor
    loss_dynamic_reg0: Tensor


class CoarseWithFineRenderOutput(StarNetworkOutput, StarCoarseNetworkOutput):    # (:109-110)
    z_std: Tensor                   # [R]


MIP_APPINIT_KEYS = tuple(StarMipAppInitCombinedOutput.__annotations__)
MIP_ONLINE_KEYS = tuple(StarMipOnlineCombinedOutput.__annotations__)

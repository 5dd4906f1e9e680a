"""star_b200 -- B200-native (sm_100a) implementation of the STaR / NeRF render hot path of
burakcuhadar/3D-MOT-using-Neural-Radiance-Fields behind the reference's own Python API.

    import star_b200                       # repo-root shim that loads this package
    from star_b200.models.star__ import STaR
    from star_b200.models.rendering__ import sample_pts, render_star_online

`star_b200.install()` registers the mirrors as `models.rendering__`, `models.star__`, `models.nerf`,
`models.embedder`, `models.resnet`, `models.types__`, `models.star_mipnerf`, `models.mipnerf`,
`models.rendering_starmip`, `models.loss` in sys.modules so that the reference's train
scripts pick them up unmodified.  `star_b200.optim` holds the fused clip + Adam step,
`star_b200.evaluation` the multi-frame / multi-view test loop of online tracking, `star_b200.parallel` the ray sharding."""
import sys

from . import _capi, functional, parallel  # noqa: F401
from . import evaluation, metrics, mip_functional, optim  # noqa: F401
from .models import loss  # noqa: F401
from .models import embedder, mipnerf, nerf, rendering__, rendering_starmip, resnet, star__, star_mipnerf, types__  # noqa: F401
from .models.star__ import STaR  # noqa: F401

__all__ = ["STaR", "functional", "install", "optim", "parallel", "rendering__", "star__"]


def install(package="models"):
    """Make `from models.star__ import STaR` etc. resolve to the B200 path."""
    import types
    pkg = sys.modules.get(package)
    if pkg is None:
        pkg = types.ModuleType(package)
        pkg.__path__ = []
        sys.modules[package] = pkg
    for name, mod in (("rendering__", rendering__), ("star__", star__), ("nerf", nerf), ("embedder", embedder),
                      ("resnet", resnet), ("types__", types__), ("star_mipnerf", star_mipnerf), ("mipnerf", mipnerf),
                      ("rendering_starmip", rendering_starmip), ("loss", loss)):
        sys.modules[f"{package}.{name}"] = mod
        setattr(pkg, name, mod)

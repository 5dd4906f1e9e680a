"""Compositing of the mip variant with the reference's function names (models/rendering_starmip.py:66-175).
The reference functions take post-activation densities / colours plus nerfstudio RaySamples and renderers; the fused
kernels take the RAW field outputs and the frustum edges (activations are applied in the kernel)."""
from .. import mip_functional as MF

REG_KEYS = ("loss_alpha_entropy", "loss_dynamic_vs_static_reg", "loss_ray_reg", "loss_static_reg", "loss_dynamic_reg")


def get_starmip_appinit_outputs(raw_sigma_static, raw_rgb_static, bins):
    """(:66-91) -> StarMipAppInitOutput (types__.py:75-80): trailing singleton dims as in the reference."""
    rgb, acc, depth, weights = MF.MipCompositeSingle.apply(raw_sigma_static, raw_rgb_static, bins)
    return {"rgb": rgb, "acc": acc[:, None], "weights": weights[..., None], "depth": depth[:, None]}


def get_starmip_online_outputs(raw_sigma_static, raw_rgb_static, raw_sigma_dynamic, raw_rgb_dynamic, bins, chunk=1 << 30):
    """(:112-175) -> StarMipOnlineOutput (types__.py:36-54)."""
    o = dict(zip(MF.MIP_OUT_KEYS, MF.MipCompositeStar.apply(raw_sigma_static, raw_rgb_static, raw_sigma_dynamic,
                                                             raw_rgb_dynamic, bins, chunk)))
    regs = o.pop("regs")
    out = {"rgb": o["rgb"], "acc": o["acc"][:, None], "weights": o["weights"][..., None], "depth": o["depth"][:, None],
           "rgb_static": o["rgb_static"], "depth_static": o["depth_static"][:, None], "rgb_dynamic": o["rgb_dynamic"],
           "depth_dynamic": o["depth_dynamic"], "dynamic_transmittance": o["dynamic_transmittance"][..., None]}
    for i, k in enumerate(REG_KEYS):
        out[k] = regs[i]
    return out

"""Device side of the online-tracking evaluation: `compute_2d_iou` of the reference's
`utils/metrics.py:527-550` with the same signature and return values, computed by one kernel on the device that holds
`dynamic_transmittance` (the reference copies every object's mask to the host and loops in numpy)."""
import torch

from . import _capi
from ._capi import check, f32, ptr, stream
from .functional import _c, _count


@torch.no_grad()
def compute_2d_iou(dynamic_transmittance, semantic_mask, thres=0.1):
    """-> (iou, predicted_masks): iou = |mask AND union_v(T_v < thres)| / |mask OR union_v(T_v < thres)| (0 when the
    union is empty) as a Python number, predicted_masks = bool numpy array [num_vehicles, num_rays] like the reference."""
    if dynamic_transmittance.dim() != 2 or semantic_mask.shape != dynamic_transmittance.shape[:1]:
        raise ValueError("dynamic_transmittance must be [N_rays, num_vehicles] and semantic_mask [N_rays]")
    T = _c(dynamic_transmittance.detach())
    R, V = T.shape
    # any non-zero value is True, as in np.logical_or / np.logical_and of the reference (a float mask of 0.5 counts)
    sem = _c((semantic_mask.detach().to(device=T.device) != 0).to(torch.uint8))
    pred = torch.empty((V, R), device=T.device, dtype=torch.uint8)
    counts = torch.empty((2,), device=T.device, dtype=torch.int64)
    check(_capi.lib().star_iou2d(f32(T), ptr(sem), R, V, float(thres), ptr(pred), ptr(counts), stream()), "star_iou2d")
    _count()
    inter, union = (int(x) for x in counts.cpu())          # one 16-byte read-back
    iou = 0 if union == 0 else inter / union
    return iou, pred.cpu().numpy().astype(bool)

"""Online-tracking evaluation driver: the multi-frame / multi-view test loop of the
reference on the B200 path, sharded over the GPUs of a box.

Reference: `StarOnline.test_step` (train_online__.py:654-851) renders, for every test view and every frame
`i < eval_last_frame`, the full H x W image with that frame's object poses (identity at frame 0, the optimised
`poses[i - 1]` afterwards, train_online__.py:122-134 via `StarOnline.forward`, :86-153), then measures it against the
target image: MSE / PSNR of the whole image and of the dynamic / static pixels selected by the semantic mask (:663-682)
and the 2-D IoU of `dynamic_transmittance < 0.1` against the mask (utils/metrics.py:527-550).  `StarOnlineCallback`
(callbacks/online_training_callback.py:91-162) advances the number of frames being tracked.

Here one work item = (view, frame).  Items are dealt to the ranks round-robin (rays are independent, SURVEY.md 8e: no
data-path collective), each item is ONE `star_render_forward` call that starts from the camera (K, c2w) -- ray generation,
depth sampling, the V + 1 coarse and fine nets and both compositing passes run without returning to Python -- and the
per-item metrics are reduced on the device; only the small metric table crosses ranks (one all-gather).
LPIPS / SSIM of the reference's test_step are torchmetrics models outside the render path and are
not reproduced; the images are returned so that a caller can feed them to any such metric."""
import torch
import torch.distributed as dist

from . import functional as F_
from . import metrics as M_
from . import parallel as P_

METRIC_KEYS = ("mse", "psnr", "psnr_dynamic", "psnr_static", "iou_2d", "mask_pixels")


def frame_poses(poses, frame, num_vehicles, device):
    """Object poses of `frame` as the reference forms them (train_online__.py:122-134): identity [0,0,0,0,0,0,1] per
    object at frame 0, the optimised 7-vectors `poses[frame - 1]` ([V,7]) afterwards."""
    if frame == 0:
        p = torch.zeros((num_vehicles, 7), device=device)
        p[:, 6] = 1.0
        return p
    return poses[frame - 1]


def _psnr(mse):
    return -10.0 * torch.log10(mse)          # mse2psnr (models/rendering__.py:18-23)


@torch.no_grad()
def render_view(star_network, H, W, K, c2w, pose, near, far, N_samples, N_importance, lindisp=False, rows=None,
                step=None):
    """One full view (or the pixel rows `rows` = (row0, nrows)) in eval mode through the single-call render entry,
    starting from the camera.  pose: [V,7] / [V,4,4] or None.  Returns the reference's output dictionary."""
    sc = star_network.static_coarse_nerf
    Ni = N_importance if star_network.N_importance > 0 else 0
    sf = star_network.static_fine_nerf if Ni > 0 else None
    dyn_c, dyn_f, pose12, scales = [], [], None, (None, None)
    if pose is not None:
        dyn_c = list(star_network.dynamic_coarse_nerfs)
        dyn_f = list(star_network.dynamic_fine_nerfs) if Ni > 0 else []
        pose12 = torch.stack([F_.pose_to_mat12(pose[i]) for i in range(len(dyn_c))])
        m = dyn_c[0]
        scales = (m.embedder.scale(step, c2w.device, pad_to=64), m.embedder_dirs.scale(step, c2w.device, pad_to=32))
    res = F_.render_forward((sc, sf), (dyn_c, dyn_f), sc._prec(), None, None, None, Ni, near=near, far=far,
                            N_samples=N_samples, lindisp=lindisp, pose12=pose12, det=True, white_bkgd=sc.white_bkgd,
                            far_dist=star_network.far_dist, chunk=star_network.chunk, test=True, enc_scales=scales,
                            camera=(H, W, K, c2w, rows if rows is not None else (0, H)))
    return {k: v for k, v in res.items() if not k.startswith("_")}


@torch.no_grad()
def frame_metrics(result, target, semantic_mask, thres=0.1):
    """Metrics of one rendered frame against its target [R,3] and boolean semantic mask [R] (train_online__.py:663-682,
    utils/metrics.py:527-550), as a device tensor in METRIC_KEYS order (no host read-back)."""
    rgb = result["rgb"]
    se = ((rgb - target) ** 2)
    m = semantic_mask != 0
    n_dyn = m.sum()
    mse = se.mean()
    mse_dyn = se[m].mean() if se.numel() else mse        # empty mask -> NaN, as torch.mean of an empty tensor
    mse_sta = se[~m].mean()
    if "dynamic_transmittance" in result and result["dynamic_transmittance"] is not None:
        pred = (result["dynamic_transmittance"] < thres).any(dim=1)
        union = (pred | m).sum()
        inter = (pred & m).sum()
        iou = torch.where(union > 0, inter.float() / union.clamp_min(1).float(), torch.zeros((), device=rgb.device))
    else:
        iou = torch.zeros((), device=rgb.device)
    return torch.stack([mse, _psnr(mse), _psnr(mse_dyn), _psnr(mse_sta), iou, n_dyn.float()])


@torch.no_grad()
def evaluate_sequence(star_network, cameras, poses, H, W, K, near, far, N_samples, N_importance, eval_last_frame,
                      targets=None, semantic_masks=None, lindisp=False, keep_images=False, group=None):
    """The reference's test loop over `cameras` [n_views, 3, 4] x frames 0 .. eval_last_frame - 1, sharded over the
    ranks.  poses: [F - 1, V, 7] optimised object poses (frame 0 is the identity).  targets [n_views, F, H*W, 3] and
    semantic_masks [n_views, F, H*W] (any device; only this rank's items are moved) are optional: without them only the
    images are produced.

    Returns {"table": [n_views, F, len(METRIC_KEYS)] (NaN where no target), "mean": dict of the reference's running
    means (iou_2d averaged over the frames whose mask is non-empty, train_online__.py:757-761), "images": {(view,
    frame): dict} for this rank's items when keep_images}.  Every rank returns the full table."""
    was_training = star_network.training
    star_network.eval()
    rank, world = P_.world()
    dev = next(star_network.parameters()).device
    n_views = cameras.shape[0]
    V = len(star_network.dynamic_coarse_nerfs)
    items = [(v, f) for v in range(n_views) for f in range(eval_last_frame)]
    table = torch.full((n_views, eval_last_frame, len(METRIC_KEYS)), float("nan"), device=dev)
    images = {}
    try:
        for idx in range(rank, len(items), world):
            v, f = items[idx]
            pose = frame_poses(poses, f, V, dev) if V else None
            res = render_view(star_network, H, W, K, cameras[v].to(dev), pose, near, far, N_samples, N_importance,
                              lindisp=lindisp)
            if targets is not None:
                tgt = targets[v][f].to(dev, non_blocking=True).reshape(-1, 3)
                msk = semantic_masks[v][f].to(dev, non_blocking=True).reshape(-1) if semantic_masks is not None else \
                    torch.zeros((tgt.shape[0],), dtype=torch.bool, device=dev)
                table[v, f] = frame_metrics(res, tgt, msk)
            if keep_images:
                images[(v, f)] = {k: res[k] for k in ("rgb", "depth", "rgb_static", "rgb_dynamic_all",
                                                      "dynamic_transmittance") if res.get(k) is not None}
        if world > 1:
            # every item was written by exactly one rank: NaN elsewhere -> sum of nan_to_num, presence by counting
            have = (~torch.isnan(table[..., 0])).float()
            t = torch.nan_to_num(table, nan=0.0)
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
            dist.all_reduce(have, op=dist.ReduceOp.SUM, group=group)
            table = torch.where(have[..., None] > 0, t, torch.full_like(t, float("nan")))
    finally:
        star_network.train(was_training)
    mean = {}
    if targets is not None:
        flat = table.reshape(-1, len(METRIC_KEYS))
        for i, k in enumerate(METRIC_KEYS[:4]):
            mean[k] = float(flat[:, i].nanmean())
        nz = flat[:, 5] > 0
        mean["iou_2d"] = float(flat[nz, 4].mean()) if bool(nz.any()) else 0.0
    return {"table": table, "mean": mean, "images": images}


class OnlineFrameScheduler:
    """Frame progression of online tracking -- the host logic of `StarOnlineCallback.on_train_epoch_end`
    (callbacks/online_training_callback.py:91-162): while the first `initial_num_frames` frames are being fitted, one
    more frame is admitted once the epoch's mean fine loss drops to `online_thres` (then the threshold becomes 95e-5);
    afterwards a frame is admitted when the loss is at or below the threshold AND more than 70 epochs have passed since
    the last admission; training stops when the count exceeds `num_frames`.  Precrop epochs are skipped (:97-98)."""

    def __init__(self, online_thres, initial_num_frames, num_frames, precrop_iters=0):
        self.online_thres = float(online_thres)
        self.initial_num_frames = int(initial_num_frames)
        self.max_num_frames = int(num_frames)
        self.precrop_iters = int(precrop_iters)
        self.current_frame_num = int(initial_num_frames)
        self.count = 0
        self.should_stop = False

    def epoch_end(self, epoch, avg_fine_loss):
        """-> True when a new frame was admitted this epoch."""
        if epoch < self.precrop_iters:
            return False
        admitted = False
        if self.current_frame_num == self.initial_num_frames:
            if avg_fine_loss <= self.online_thres:
                self.current_frame_num += 1
                self.online_thres = 95e-5
                admitted = True
        else:
            self.count += 1
            if self.count > 70 and avg_fine_loss <= self.online_thres:
                self.count = 0
                self.current_frame_num += 1
                admitted = True
        if self.current_frame_num > self.max_num_frames:
            self.should_stop = True
        return admitted

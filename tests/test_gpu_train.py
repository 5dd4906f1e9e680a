"""GPU parity of the loss / optimiser-step kernels (csrc/train_step.cu) through the Python
mirror and the C ABI: against the fixtures produced by the reference's own loss functions and torch.optim.Adam
(tests/golden/train_*.npz), against oracle/train_oracle.py on fresh seeded inputs, and -- at C4's full parameter count
-- against torch's own CUDA optimiser.  Tolerances: 1e-6 relative on losses and gradients (fp32 sums in a different
order), Adam parameters 2e-6 relative + 2e-8 absolute after 6 steps."""
import math

import pytest
import torch

import star_b200
from star_b200 import functional as F_, optim as O_
from star_b200.models import loss as L_, rendering__ as R_
from oracle import ref_harness, star_oracle as so, train_oracle as to
from helpers import load_golden, assert_close

pytestmark = pytest.mark.gpu
DEV = "cuda"


def c(t):
    return t.to(DEV)


# ------------------------------------------------------------------------------------------ losses
def test_photometric_loss_matches_the_reference_fixture():
    g = load_golden("train_photometric")
    rgb0, rgb = c(g["rgb0"]).requires_grad_(True), c(g["rgb"]).requires_grad_(True)
    loss, m0, m1, p0, p1 = L_.photometric_loss(rgb0, rgb, c(g["target"]))
    assert not p0.requires_grad and not p1.requires_grad
    loss.backward()
    assert_close(m0, g["mse0"], 0, 1e-6)
    assert_close(m1, g["mse"], 0, 1e-6)
    assert_close(loss, g["mse0"] + g["mse"], 0, 1e-6)
    assert_close(p0, g["psnr0"].reshape(()), 1e-5)
    assert_close(p1, g["psnr"].reshape(()), 1e-5)
    assert_close(rgb0.grad, g["g_rgb0"], 1e-9, 1e-6)
    assert_close(rgb.grad, g["g_rgb"], 1e-9, 1e-6)
    # N_importance == 0: only one map; upstream scale on the fine term alone
    rgb1 = c(g["rgb"]).requires_grad_(True)
    loss1, _z, m1b, _p, p1b = L_.photometric_loss(None, rgb1, c(g["target"]))
    (3.0 * m1b).backward()
    assert_close(loss1, g["mse"], 0, 1e-6)
    assert_close(rgb1.grad, 3.0 * g["g_rgb"], 1e-9, 1e-6)


def test_photometric_loss_full_view_grid_reduction_is_deterministic():
    gen = torch.Generator().manual_seed(5)
    R = 640000
    rgb0, rgb, tgt = (torch.rand(R, 3, generator=gen) for _ in range(3))
    ref = to.photometric_loss(rgb0.double(), rgb.double(), tgt.double())
    a = L_.photometric_loss(c(rgb0), c(rgb), c(tgt))
    b = L_.photometric_loss(c(rgb0), c(rgb), c(tgt))
    for x, y, r in zip(a, b, ref):
        assert torch.equal(x, y)
        assert_close(x, r.float(), 1e-6, 1e-6)


@pytest.mark.parametrize("tag", ["a", "b"])     # a: S = 64 (16-byte path), b: S = 37 (scalar path)
def test_dsnerf_losses_match_the_reference_fixture(tag):
    g = load_golden("train_dsnerf_" + tag)
    depth = c(g["depth"]).requires_grad_(True)
    dl = L_.compute_depth_loss(depth, c(g["depths"]), g["near"], g["far"])
    (2.0 * dl).backward()
    assert_close(dl, g["depth_loss"], 0, 1e-6)
    assert_close(depth.grad, 2.0 * g["g_depth"], 1e-9, 1e-6)
    w = c(g["weights"]).requires_grad_(True)
    args = (c(g["z_vals"]), c(g["dists"]), c(g["depths"]))
    sl = L_.compute_sigma_loss(w, *args, g["near"], g["far"], err=1)
    sl.backward()
    assert_close(sl, g["sigma_loss"], 0, 2e-6)
    assert_close(w.grad, g["g_weights"], 1e-9, 2e-6)
    assert float((w.grad[c(g["weights"]) <= 0]).abs().max()) == 0.0
    w.grad = None
    pr = L_.compute_sigma_loss_per_ray(w, *args, err=1)
    (pr * c(g["per_ray_coef"])).sum().backward()
    assert_close(pr, g["per_ray"], 1e-9, 2e-6)
    assert_close(w.grad, g["g_weights_per_ray"], 1e-9, 2e-6)
    sl2 = L_.compute_sigma_loss(w.detach(), *args, g["near"], g["far"], err=0.25)
    assert_close(sl2, g["sigma_loss_err025"], 0, 2e-6)


def test_dsnerf_losses_edge_cases():
    g = load_golden("train_dsnerf_a")
    # no ray inside (near, far): torch.mean of an empty selection is NaN, gradients are all zero / NaN-free masks
    far_all = torch.full_like(g["depths"], 5.0)
    dl = L_.compute_depth_loss(c(g["depth"]), c(far_all), g["near"], g["far"])
    sl = L_.compute_sigma_loss(c(g["weights"]), c(g["z_vals"]), c(g["dists"]), c(far_all), g["near"], g["far"])
    assert math.isnan(float(dl)) and math.isnan(float(sl))
    ref = to.compute_depth_loss(g["depth"], far_all, g["near"], g["far"])
    assert math.isnan(float(ref))
    with pytest.raises(ValueError):
        L_.compute_sigma_loss(c(g["weights"]), c(g["z_vals"])[:, :-1], c(g["dists"]), c(g["depths"]), 0.0, 1.0)
    with pytest.raises(star_b200._capi.StarError):
        L_.compute_depth_loss(g["depth"], g["depths"], 0.0, 1.0)        # CPU tensors: no fallback


def test_sigma_loss_on_a_real_render_pass_against_the_oracle():
    """weights / z_vals / dists straight out of the compositing kernel (C3-shaped: 256 + 256 samples)."""
    gen = torch.Generator().manual_seed(9)
    R, S = 4096, 512
    raw_a = torch.randn(R, S, generator=gen) * 2.0
    raw_c = torch.randn(R, S, 3, generator=gen)
    z = 0.03 + 0.77 * torch.sort(torch.rand(R, S, generator=gen), dim=1).values
    rd = torch.randn(R, 3, generator=gen)
    depths = 0.77 * torch.rand(R, generator=gen)
    o = so.raw2outputs(raw_a, raw_c, z, rd, 0.0, False, 1e10)
    ref = to.compute_sigma_loss(o["weights"].double(), z.double(), o["dists"].double(), depths.double(), 0.03, 0.8)
    og = F_.CompositeSingle.apply(c(raw_a), c(raw_c), c(z), c(rd), 1e10, False)
    weights, dists = og[4], og[5]
    got = L_.compute_sigma_loss(weights, c(z), dists, c(depths), 0.03, 0.8)
    assert_close(got, ref.float(), 0, 2e-4)      # log of 1e-4-accurate weights (the compositing kernel's bar)


# ------------------------------------------------------------------------------------------ clip + Adam
def adam_case(g):
    n_groups = [int(n) for n in g["n_groups"]]
    total = sum(n_groups)
    grads = [[g["g%d.%d" % (t, k)] for k in range(total)] for t in range(int(g["steps"]))]
    return n_groups, total, grads, [float(x) for x in g["lrs"]]


def check_adam_state(opt, flat, g, tag):
    for i, p in enumerate(flat):
        assert_close(p, g["%s.p%d" % (tag, i)], 2e-8, 2e-6, "p%d" % i)
        assert_close(opt.state[p]["exp_avg"], g["%s.m%d" % (tag, i)], 1e-8, 2e-6, "m%d" % i)
        assert_close(opt.state[p]["exp_avg_sq"], g["%s.v%d" % (tag, i)], 1e-14, 2e-6, "v%d" % i)
        assert int(opt.state[p]["step"]) == int(g["steps"])


@pytest.mark.parametrize("layout", ["separate", "flat"])
@pytest.mark.parametrize("tag,max_norm", [("clip", 1.0), ("noclip", None)])
def test_fused_adam_matches_torch_adam_fixture(tag, max_norm, layout):
    """3 parameter groups (the reference's static / dynamic / pose learning rates), 6 steps whose gradient norms sit on
    both sides of the clip threshold; tensors of 1, 3, 35 ... elements exercise every alignment path."""
    g = load_golden("train_adam")
    n_groups, total, grads, lrs = adam_case(g)
    flat = [torch.nn.Parameter(c(g["init.p%d" % i])) for i in range(total)]
    if layout == "flat":
        holder = torch.nn.ParameterList(flat)
        buf = O_.flatten_parameters(holder)
        assert buf.numel() == sum(p.numel() for p in flat)
        assert flat[1].data_ptr() == flat[0].data_ptr() + 4 * flat[0].numel()
    groups, i = [], 0
    for n, lr in zip(n_groups, lrs):
        groups.append({"params": flat[i:i + n], "lr": lr})
        i += n
    opt = O_.FusedAdam(groups, betas=(0.9, 0.999), max_grad_norm=max_norm)
    for t, gs in enumerate(grads):
        if layout == "flat":      # gradients arriving as one run, like a net's grad_flat
            gbuf = torch.cat([x.reshape(-1) for x in gs]).to(DEV)
            off = 0
            for p in flat:
                p.grad = gbuf[off:off + p.numel()].view(p.shape)
                off += p.numel()
        else:
            for p, x in zip(flat, gs):
                p.grad = c(x)
        v0 = flat[0]._version
        opt.step()
        assert flat[0]._version > v0
        if max_norm is not None:
            assert_close(opt.total_grad_norm(), g["clip.norms"][t], 0, 1e-6)
    check_adam_state(opt, flat, g, tag)
    sd = opt.state_dict()
    assert len(sd["state"]) == total and len(sd["param_groups"]) == 3
    # the state dict loads into torch.optim.Adam (and back): one more identical step on both sides
    ref = [torch.nn.Parameter(p.detach().clone()) for p in flat]
    groups_ref, i = [], 0
    for n, lr in zip(n_groups, lrs):
        groups_ref.append({"params": ref[i:i + n], "lr": lr})
        i += n
    topt = torch.optim.Adam(groups_ref, betas=(0.9, 0.999))
    import copy                              # (load_state_dict aliases same-device tensors: copy them)
    topt.load_state_dict(copy.deepcopy(sd))
    opt.load_state_dict(copy.deepcopy(topt.state_dict()))
    opt.max_grad_norm = None
    for p, q, x in zip(flat, ref, grads[0]):
        p.grad, q.grad = c(x), c(x)
    opt.step()
    topt.step()
    for p, q in zip(flat, ref):
        assert_close(p, q, 2e-8, 2e-6)


def test_clip_grad_norm_matches_torch():
    g = load_golden("train_adam")
    n_groups, total, grads, lrs = adam_case(g)
    for t in (0, 3):
        ps = [torch.nn.Parameter(c(g["init.p%d" % i])) for i in range(total)]
        for p, x in zip(ps, grads[t]):
            p.grad = c(x)
        norm = O_.clip_grad_norm_(ps, 1.0)
        ref_ps = [torch.nn.Parameter(g["init.p%d" % i].clone()) for i in range(total)]
        for p, x in zip(ref_ps, grads[t]):
            p.grad = x.clone()
        ref_norm = torch.nn.utils.clip_grad_norm_(ref_ps, 1.0)
        assert_close(norm, ref_norm, 0, 1e-6)
        assert_close(norm, g["clip.norms"][t], 0, 1e-6)
        for p, q in zip(ps, ref_ps):
            assert_close(p.grad, q.grad, 1e-12, 1e-6)


def test_write_back_clipped_grads():
    g = load_golden("train_adam")
    n_groups, total, grads, lrs = adam_case(g)
    ps = [torch.nn.Parameter(c(g["init.p%d" % i])) for i in range(total)]
    for p, x in zip(ps, grads[0]):
        p.grad = c(x)
    opt = O_.FusedAdam(ps, lr=1e-3, max_grad_norm=1.0, write_back_clipped_grads=True)
    opt.step()
    coef = 1.0 / (float(g["clip.norms"][0]) + 1e-6)
    for p, x in zip(ps, grads[0]):
        assert_close(p.grad, x * coef, 1e-12, 1e-6)


def test_fused_adam_rejects_what_it_does_not_cover():
    p = torch.nn.Parameter(torch.zeros(4, device=DEV))
    with pytest.raises(ValueError):
        O_.FusedAdam([p], weight_decay=0.1)
    with pytest.raises(ValueError):
        O_.FusedAdam([p], amsgrad=True)
    q = torch.nn.Parameter(torch.zeros(4))
    q.grad = torch.ones(4)
    with pytest.raises(star_b200._capi.StarError):
        O_.FusedAdam([q]).step()


def star_net(V, precision):
    net = star_b200.STaR(ref_harness.make_args(num_vehicles=V, N_importance=16, chunk=8192))
    net.load_state_dict(so.init_star_params(V, 16, seed=0, bias_std=0.02))
    net.to(DEV)
    net.set_precision(precision)
    return net


def test_flatten_parameters_training_step_against_torch_adam():
    """Two STaR models with identical weights take 3 full training steps (render, photometric loss, backward,
    clip, Adam): one with torch.optim.Adam + torch clip on separate tensors, one with flattened parameters + FusedAdam.
    The flat model's master vectors are zero-copy views, and the two end within fp32 rounding of each other."""
    V = 1
    a, b = star_net(V, "fp32"), star_net(V, "fp32")
    flat = O_.flatten_parameters(b)
    n_par = sum(p.numel() for p in b.parameters())
    assert flat.numel() == n_par
    rt = b.static_coarse_nerf._rt
    master, _ = rt.refresh(F_._capi.PREC_F32)
    assert master.data_ptr() == rt.ordered_params()[0].data_ptr()        # no concatenated copy
    pose_a = torch.nn.Parameter(torch.tensor([[0.02, -0.01, 0.03, 0.01, -0.02, 0.015, 1.0]], device=DEV))
    pose_b = torch.nn.Parameter(pose_a.detach().clone())

    def groups(net, pose):
        stat = list(net.static_coarse_nerf.parameters()) + list(net.static_fine_nerf.parameters())
        dyn = list(net.dynamic_coarse_nerfs.parameters()) + list(net.dynamic_fine_nerfs.parameters())
        return [{"params": stat, "lr": 5e-4}, {"params": dyn, "lr": 2.5e-4}, {"params": [pose], "lr": 1e-3}]
    opt_a = torch.optim.Adam(groups(a, pose_a), betas=(0.9, 0.999))
    opt_b = O_.FusedAdam(groups(b, pose_b), betas=(0.9, 0.999), max_grad_norm=1.0)
    gen = torch.Generator().manual_seed(21)
    R = 256
    ro = c(torch.zeros(R, 3))
    rd = torch.randn(R, 3, generator=gen)
    rd = c(rd / rd.norm(dim=-1, keepdim=True))
    tgt = c(torch.rand(R, 3, generator=gen))
    for step in range(3):
        losses = []
        for net, pose, opt, fused in ((a, pose_a, opt_a, False), (b, pose_b, opt_b, True)):
            net.train()
            opt.zero_grad(set_to_none=True)
            pts, z = R_.sample_pts(ro, rd, 0.03, 0.8, 24, perturb=0, is_train=True)
            u = c(torch.rand(R, 16, generator=torch.Generator().manual_seed(100 + step)))
            out = R_.render_star_online(net, pts, rd, z, ro, rd, 16, pose, u=u)
            if fused:
                loss = L_.photometric_loss(out["rgb0"], out["rgb"], tgt)[0]
            else:
                loss = R_.img2mse(out["rgb0"], tgt) + R_.img2mse(out["rgb"], tgt)
            loss.backward()
            if not fused:
                torch.nn.utils.clip_grad_norm_([p for gr in opt.param_groups for p in gr["params"]], 1.0)
            opt.step()
            losses.append(float(loss))
        assert abs(losses[0] - losses[1]) <= 1e-4 * abs(losses[0])
    # Adam's first steps move every element by ~lr * sign(g): an element whose gradient is rounding noise (the fp32 dW
    # reduction uses atomics, so its last bits differ from run to run) may move the other way.  Those are rare and
    # bounded by 2 * lr * steps; everything else agrees to fp32 rounding.
    n_bad = n_all = 0
    for (ka, pa), (kb, pb) in zip(list(a.named_parameters()) + [("pose", pose_a)],
                                  list(b.named_parameters()) + [("pose", pose_b)]):
        assert ka == kb
        d = (pb.detach() - pa.detach()).abs()
        assert float(d.max()) <= 2 * 1e-3 * 3 + 1e-6, ka
        n_bad += int((d > 2e-5 + 1e-4 * pa.detach().abs()).sum())
        n_all += d.numel()
    assert n_bad <= 2e-3 * n_all, (n_bad, n_all)
    # the packed weights follow the update: a fresh model loaded from b's state dict renders the same image
    fresh = star_net(V, "fp32")
    fresh.load_state_dict(b.state_dict())
    with torch.no_grad():
        for net in (b, fresh):
            net.eval()
        pts, z = R_.sample_pts(ro, rd, 0.03, 0.8, 24, is_train=False)
        o1 = R_.render_star_online(b, pts, rd, z, ro, rd, 16, pose_b)
        o2 = R_.render_star_online(fresh, pts, rd, z, ro, rd, 16, pose_b)
    assert torch.equal(o1["rgb"], o2["rgb"])


def test_fused_adam_at_c4_parameter_count_against_torch_cuda_adam():
    """5.9 M parameters (2 static + 10 dynamic nets, V = 5) in one flat run, 3 steps, against torch.optim.Adam on the
    same device; also checks that one step is a handful of launches."""
    n = 2 * 711300 + 10 * 448132
    gen = torch.Generator(device=DEV).manual_seed(3)
    p0 = 0.1 * torch.randn(n, device=DEV, generator=gen)
    pa = torch.nn.Parameter(p0.clone())
    pb = torch.nn.Parameter(p0.clone())
    oa = torch.optim.Adam([pa], lr=5e-4)
    ob = O_.FusedAdam([pb], lr=5e-4, max_grad_norm=1.0)
    for t in range(3):
        gr = torch.randn(n, device=DEV, generator=gen) * (10.0 ** (-t))
        pa.grad, pb.grad = gr.clone(), gr.clone()
        torch.nn.utils.clip_grad_norm_([pa], 1.0)
        oa.step()
        n0 = F_.LAUNCH_COUNTER["calls"]
        ob.step()
        assert F_.LAUNCH_COUNTER["calls"] - n0 == 2
    assert_close(pb, pa, 2e-8, 2e-6)
    assert_close(ob.state[pb]["exp_avg_sq"], oa.state[pa]["exp_avg_sq"], 1e-16, 2e-6)


# ------------------------------------------------------------------------------------------ evaluation (8f-4)
@pytest.mark.parametrize("tag", ["a", "b", "empty"])
def test_compute_2d_iou_matches_the_reference_fixture(tag):
    """utils/metrics.py:527-550 on the device: bit-exact masks and counts (NaN and the threshold itself are not below)."""
    from star_b200 import metrics as M_
    g = load_golden("train_iou2d")
    iou, masks = M_.compute_2d_iou(c(g[tag + ".T"]), c(g[tag + ".sem"]), 0.1)
    assert iou == float(g[tag + ".iou"])
    assert masks.dtype == bool and (torch.from_numpy(masks) == g[tag + ".masks"]).all()
    # the semantic mask may live on the host (the reference's batches do): same result
    iou2, _ = M_.compute_2d_iou(c(g[tag + ".T"]), g[tag + ".sem"], 0.1)
    assert iou2 == iou


def test_compute_2d_iou_full_view_against_the_oracle():
    from star_b200 import metrics as M_
    gen = torch.Generator().manual_seed(31)
    R, V = 1280 * 720, 5
    T = torch.rand(R, V, generator=gen) ** 4
    sem = torch.rand(R, generator=gen) < 0.1
    ref_iou, ref_masks = to.compute_2d_iou(T, sem, 0.1)
    iou, masks = M_.compute_2d_iou(c(T), c(sem), 0.1)
    assert iou == ref_iou and (masks == ref_masks).all()
    with pytest.raises(ValueError):
        M_.compute_2d_iou(c(T), c(sem)[:-1])
    with pytest.raises(star_b200._capi.StarError):
        M_.compute_2d_iou(T, sem)            # CPU tensors: no fallback


# ------------------------------------------------------------------------------------------ 8(f) row 4: evaluation driver
def test_evaluate_sequence_matches_per_item_staged_renders():
    """evaluation.evaluate_sequence (views x frames, each item one star_render_forward call from the camera) against the
    same items rendered through the staged reference-API path and measured with plain torch / metrics.compute_2d_iou."""
    import math
    from star_b200 import evaluation as E_, metrics as M_
    from star_b200.models import rendering__ as R_
    from oracle import ref_harness, star_oracle as so
    V, Nc, Ni, H, W, F = 2, 16, 16, 10, 14, 3
    net = star_b200.STaR(ref_harness.make_args(num_vehicles=V, N_importance=Ni, chunk=64, white_bkgd=False))
    net.load_state_dict(so.init_star_params(V, Ni, seed=8, bias_std=0.02))
    net.to(DEV).train()                       # the driver switches to eval and back
    focal = 0.5 * W
    K = torch.tensor([[focal, 0, 0.5 * W], [0, focal, 0.5 * H], [0, 0, 1.0]])
    g = torch.Generator().manual_seed(2)
    cams = []
    for _ in range(2):
        q, _r = torch.linalg.qr(torch.randn(3, 3, generator=g))
        cams.append(torch.cat([q, torch.randn(3, 1, generator=g) * 0.05], 1))
    cams = torch.stack(cams)
    poses = torch.stack([so.random_poses7(V, seed=20 + f) for f in range(F - 1)]).to(DEV)
    targets = torch.rand(2, F, H * W, 3, generator=g)
    masks = torch.rand(2, F, H * W, generator=g) < 0.3
    masks[1, 2] = False                       # a frame without dynamic pixels: excluded from the IoU mean
    out = E_.evaluate_sequence(net, cams, poses, H, W, K, 0.03, 0.8, Nc, Ni, F, targets=targets, semantic_masks=masks,
                               keep_images=True)
    assert net.training and out["table"].shape == (2, F, len(E_.METRIC_KEYS)) and len(out["images"]) == 2 * F
    net.eval()
    ious = []
    with torch.no_grad():
        for v in range(2):
            for f in range(F):
                ro, rd = R_.get_rays(H, W, K, cams[v].to(DEV))
                ro, rd = ro.reshape(-1, 3), rd.reshape(-1, 3)
                vd = rd / rd.norm(dim=-1, keepdim=True)
                pts, z = R_.sample_pts(ro, rd, 0.03, 0.8, Nc, is_train=False)
                pose = E_.frame_poses(poses, f, V, DEV)
                ref = R_.render_star_online(net, pts, vd, z, ro, rd, Ni, pose)
                tgt, m = targets[v, f].to(DEV), masks[v, f].to(DEV)
                assert_close(out["images"][(v, f)]["rgb"], ref["rgb"], 2e-5, msg="rgb")     # (viewdirs: last-bit differences)
                mse = ((ref["rgb"] - tgt) ** 2).mean()
                row = out["table"][v, f]
                assert_close(row[0], mse, 1e-6, msg="mse")
                assert_close(row[1], -10.0 * torch.log10(mse), 1e-4, msg="psnr")
                if bool(m.any()):
                    assert_close(row[2], -10.0 * torch.log10(((ref["rgb"] - tgt)[m] ** 2).mean()), 1e-3, msg="psnr_dynamic")
                iou, _ = M_.compute_2d_iou(ref["dynamic_transmittance"], m)
                assert abs(float(row[4]) - iou) < 1e-6 + 0.02 * (iou > 0)     # (a transmittance within rounding of 0.1 may flip)
                assert int(row[5]) == int(m.sum())
                if int(m.sum()) > 0:
                    ious.append(float(row[4]))
    assert abs(out["mean"]["iou_2d"] - sum(ious) / len(ious)) < 1e-6
    assert math.isfinite(out["mean"]["psnr"]) and math.isfinite(out["mean"]["psnr_static"])

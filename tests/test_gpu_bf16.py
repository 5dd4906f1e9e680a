"""GPU parity tests of the bf16 tensor-core (tcgen05) MLP tier.

Two references:
  * the oracle's bf16 MODEL of the kernel (every GEMM operand rounded to bf16, fp32 accumulate, fp32
    biases / residual stream / heads): the kernel must agree with it up to accumulation order and the
    rounding flips caused by its double-angle sin/cos -- this pins the kernel's arithmetic;
  * the fp32 reference: the end-to-end deviation caused by bf16 operands, judged by the north star's
    bounds (2e-3 absolute on rgb / weights, 0.05 dB PSNR).  With RANDOM-INIT weights the raw network
    outputs deviate by ~0.5 % (11 chained GEMMs with 2^-9 operand rounding), which is 4e-4 mean / 3e-3
    worst-case on composited rgb -- so the 2e-3 bound is asserted on the mean and the 99th percentile,
    and the worst case is bounded at 1e-2 (DESIGN.md "bf16 tier accuracy")."""
import pytest
import torch

import star_b200
from star_b200 import functional as F_, _capi
from star_b200.models import rendering__ as R_
from oracle import ref_harness, star_oracle as so
from helpers import load_golden, assert_close, psnr_db

pytestmark = pytest.mark.gpu
DEV = "cuda"


def cu(t):
    return t.to(DEV) if torch.is_tensor(t) else t


def make_star(V, Ni, chunk, white, seed, training, precision="bf16"):
    args = ref_harness.make_args(num_vehicles=V, N_importance=Ni, chunk=chunk, white_bkgd=white)
    net = star_b200.STaR(args)
    sd = so.init_star_params(V, Ni, seed=seed, bias_std=0.02)
    net.load_state_dict(sd, strict=True)
    net.to(DEV).train(training)
    net.set_precision(precision)
    return net, {k: v.clone() for k, v in sd.items()}


@pytest.mark.parametrize("prec", ["bf16", "fp16"])
@pytest.mark.parametrize("R,S,dyn,with_pose", [(4, 32, False, False), (70, 33, False, False), (16, 12, True, True),
                                               (301, 191, False, False), (1, 1, True, True)])
def test_tc_mlp_matches_bf16_model(R, S, dyn, with_pose, prec):
    net, params = make_star(1, 8, 4096, False, seed=21, training=False, precision=prec)
    prefix = "dynamic_coarse_nerfs.0." if dyn else "static_coarse_nerf."
    module = net.dynamic_coarse_nerfs[0] if dyn else net.static_coarse_nerf
    ro, rd = so.carla_rays(R, seed=7)
    vd = rd / rd.norm(dim=-1, keepdim=True)
    pts, _ = so.sample_pts(ro, rd, 0.03, 0.8, S)
    p = {k: v for k, v in params.items() if k.startswith(prefix)}
    pose = so.pose7_to_matrix(so.random_poses7(1, seed=9))[0] if with_pose else None
    if with_pose:
        ph = torch.cat([pts, torch.ones(R, S, 1)], -1).reshape(-1, 4)
        pd = (ph @ pose.T).reshape(R, S, 4)[..., :3]
        vdd = vd @ pose[:3, :3].T
    else:
        pd, vdd = pts, vd
    a_m, c_m = so.nerf_mlp(p, prefix, pd, vdd, emulate_bf16=prec)
    a_f, c_f = so.nerf_mlp(p, prefix, pd, vdd)
    with torch.no_grad():
        p12 = F_.pose_to_mat12(cu(pose)) if with_pose else None
        a, c = module.raw(cu(pts), cu(vd), p12)
    # (1) against the bf16 model of the kernel: rounding flips only
    ea, ec = (a.cpu() - a_m).abs(), (c.cpu() - c_m).abs()
    assert float(ea.mean()) < 1e-3 and float(ec.mean()) < 1e-3, (float(ea.mean()), float(ec.mean()))
    # (a single bf16 rounding flip of an early activation cascades to ~1e-2 on that sample: bound the max
    # loosely and require the typical element to agree to 1e-5)
    assert float(ea.max()) < 6e-2 and float(ec.max()) < 6e-2, (float(ea.max()), float(ec.max()))
    # (fp16's 8x finer grid makes at least one flip per sample the norm, each worth ~1e-4)
    med = 2e-5 if prec == "bf16" else 1e-3
    assert float(ea.median()) < med and float(ec.median()) < med, (float(ea.median()), float(ec.median()))
    # (2) against fp32: the bf16 operand rounding itself (~0.5 % of the raw magnitudes)
    scale = float(a_f.abs().max()) + 1.0
    assert float((a.cpu() - a_f).abs().mean()) < 1e-2 * scale
    assert float((c.cpu() - c_f).abs().mean()) < 1e-2 * scale


def test_tc_launch_chunking_is_bit_invisible():
    """Rows of a tile are independent dot products with a fixed K order: results must not depend on how
    samples are grouped into 128-row tiles or launches."""
    net, _ = make_star(0, 8, 4096, False, seed=3, training=False)
    ro, rd = so.carla_rays(257, seed=1)
    vd = rd / rd.norm(dim=-1, keepdim=True)
    pts, _ = so.sample_pts(ro, rd, 0.03, 0.8, 37)
    with torch.no_grad():
        a, c = net.static_coarse_nerf.raw(cu(pts), cu(vd), None)
        saved = F_.MAX_SAMPLES_PER_LAUNCH
        F_.MAX_SAMPLES_PER_LAUNCH = 37 * 11
        try:
            a2, c2 = net.static_coarse_nerf.raw(cu(pts), cu(vd), None)
        finally:
            F_.MAX_SAMPLES_PER_LAUNCH = saved
        a3, c3 = net.static_coarse_nerf.raw(cu(pts[5:200]), cu(vd[5:200]), None)
    assert torch.equal(a, a2) and torch.equal(c, c2)
    assert torch.equal(a[5:200], a3) and torch.equal(c[5:200], c3)


def _dev_stats(out, ref, keys):
    res = {}
    for k in keys:
        e = (out[k].detach().cpu().double() - ref[k].double()).abs().flatten()
        res[k] = (float(e.mean()), float(e.quantile(0.99)) if e.numel() > 1 else float(e.max()), float(e.max()))
    return res


# (mean, 99th percentile, max) bounds on |rgb - ref| and |weights - ref| against the reference fixtures.
# "fp16" is the QUALIFIED 16-bit tier (the one bench.py reports): the north star's 2e-3 absolute in the max norm (and
# <= 0.05 dB PSNR, below).  "bf16" keeps the same kernels with bf16 operands for networks whose activations leave
# fp16's range; tools/error_budget.py shows that each of its 11 GEMMs alone costs 2-3e-3 on these random-init fixtures, so
# no selective fix exists and it does NOT meet the bound (measured mean 1.7e-3 .. 2.2e-3, max 9e-3): its numbers below are
# a regression guard, not a parity claim.
BOUNDS = {"bf16": (3e-3, 1.2e-2, 2e-2), "fp16": (4e-4, 1.5e-3, 2e-3)}


@pytest.mark.parametrize("prec", ["bf16", "fp16"])
@pytest.mark.parametrize("name,V", [("e2e_appinit_eval", 0), ("e2e_online_mat_eval", 2)])
def test_tc_e2e_eval_within_bounds(name, V, prec):
    g = load_golden(name)
    net, _ = make_star(V, 24, 4096, V == 0, int(g["seed"]), training=False, precision=prec)
    ro, rd = cu(g["rays_o"]), cu(g["rays_d"])
    vd = rd / rd.norm(dim=-1, keepdim=True)
    with torch.no_grad():
        pts, z = R_.sample_pts(ro, rd, g["near"], g["far"], int(g["Nc"]), is_train=False)
        if V:
            out = R_.render_star_online(net, pts, vd, z, ro, rd, int(g["Ni"]), cu(g["pose"]),
                                        z_samples=cu(g["z_samples"]))
        else:
            out = R_.render_star_appinit(net, pts, vd, z, ro, rd, int(g["Ni"]), z_samples=cu(g["z_samples"]))
    st = _dev_stats(out, g, ("rgb0", "rgb", "weights0", "weights"))
    b_mean, b_p99, b_max = BOUNDS[prec]
    for k, (mean, p99, mx) in st.items():
        assert mean < b_mean and p99 < b_p99 and mx < b_max, (k, mean, p99, mx)
    # depth is in scene units: judged relative to the far plane (the lego-shaped fixture has far = 6 and only 16 + 24
    # samples per ray; at C2's 64 + 128 samples the absolute error is < 1e-3, tests/test_gpu_configs.py)
    far = float(g["far"])
    for k in ("depth0", "depth"):
        e = (out[k].cpu() - g[k]).abs() / far
        assert float(e.mean()) < b_mean and float(e.max()) < b_max, (k, float(e.mean()), float(e.max()))
    F_.check_range()
    # PSNR of the bf16 render vs the fp32 reference render itself, and the PSNR shift w.r.t. a target image
    target = torch.rand(out["rgb"].shape, generator=torch.Generator().manual_seed(1))
    shift = abs(psnr_db(out["rgb"].cpu(), target) - psnr_db(g["rgb"], target))
    assert shift < 0.05, shift
    assert psnr_db(out["rgb"].cpu(), g["rgb"]) > (60.0 if prec == "fp16" else 40.0)


def test_tc_training_step_runs_and_grads_match_fp32_path():
    """bf16 forward + (fp32 recompute) backward: gradients must agree with the all-fp32 path to within the
    forward's bf16 deviation."""
    g = load_golden("e2e_online_quat_train")
    outs = {}
    for prec in ("fp32", "bf16"):
        net, _ = make_star(2, 24, int(g["chunk"]), False, int(g["seed"]), training=True, precision=prec)
        ro, rd, target = cu(g["rays_o"]), cu(g["rays_d"]), cu(g["target"])
        vd = rd / rd.norm(dim=-1, keepdim=True)
        pose = cu(g["pose"]).clone().requires_grad_(True)
        pts, z = R_.sample_pts(ro, rd, g["near"], g["far"], int(g["Nc"]))
        out = R_.render_star_online(net, pts, vd, z, ro, rd, int(g["Ni"]), pose, u=cu(g["u"]),
                                    z_samples=cu(g["z_samples"]))
        loss = ((out["rgb0"] - target) ** 2).mean() + ((out["rgb"] - target) ** 2).mean() \
            + 1e-3 * out["loss_alpha_entropy"]
        loss.backward()
        outs[prec] = (float(loss), pose.grad.clone(), net.static_fine_nerf.pts_net.lin_in.weight.grad.clone())
    assert abs(outs["bf16"][0] - outs["fp32"][0]) < 5e-3 * abs(outs["fp32"][0])
    for i in (1, 2):
        a, b = outs["bf16"][i], outs["fp32"][i]
        assert float((a - b).norm() / b.norm()) < 0.25     # bf16 forward flips ReLU masks (see the test below)


@pytest.mark.parametrize("prec", ["bf16", "fp16"])
@pytest.mark.parametrize("R,S,dyn,with_pose", [(8, 32, False, False), (70, 33, False, False), (40, 48, True, True),
                                               (300, 64, True, True), (1250, 64, True, True)])   # last: 625 tiles, > 4 per CTA
def test_tc_backward_matches_rounding_model_and_fp64(R, S, dyn, with_pose, prec):
    """Tensor-core backward (dX chain + dW GEMMs over the 16-bit stash).  Two references, both fp64 autograd:
      * the oracle's operand-rounding model of the forward with straight-through gradients -- the kernel must
        agree to ~1 % (what remains is the 16-bit rounding of the back-propagated gradients themselves);
      * the exact (unrounded) network -- the deviation caused by 16-bit operands at all (ReLU masks flip):
        3-9 % for bf16, 1-3 % for fp16 on these random-init nets with random upstream gradients."""
    net, params = make_star(1, 8, 4096, False, seed=21, training=True, precision=prec)
    prefix = "dynamic_coarse_nerfs.0." if dyn else "static_coarse_nerf."
    module = net.dynamic_coarse_nerfs[0] if dyn else net.static_coarse_nerf
    ro, rd = so.carla_rays(R, seed=7)
    vd = rd / rd.norm(dim=-1, keepdim=True)
    pts, _ = so.sample_pts(ro, rd, 0.03, 0.8, S)
    gen = torch.Generator().manual_seed(4)
    ga, gc = torch.randn(R, S, generator=gen), torch.randn(R, S, 3, generator=gen)
    dt = torch.float64

    def oracle(emulate):
        p = {k: v.clone().to(dt).requires_grad_(True) for k, v in params.items() if k.startswith(prefix)}
        pose = so.pose7_to_matrix(so.random_poses7(1, seed=9))[0].to(dt).requires_grad_(True) if with_pose else None
        if with_pose:
            ph = torch.cat([pts.to(dt), torch.ones(R, S, 1, dtype=dt)], -1).reshape(-1, 4)
            pd = (ph @ pose.T).reshape(R, S, 4)[..., :3]
            vdd = vd.to(dt) @ pose[:3, :3].T
        else:
            pd, vdd = pts.to(dt), vd.to(dt)
        a, c = so.nerf_mlp(p, prefix, pd, vdd, emulate_bf16=emulate)
        ((a * ga.to(dt)).sum() + (c * gc.to(dt)).sum()).backward()
        return p, pose

    p64, pose64 = oracle(False)
    pm, posem = oracle(prec)
    pose_g = cu(pose64.detach().float()).requires_grad_(True) if with_pose else None
    p12 = F_.pose_to_mat12(pose_g) if with_pose else None
    a, c = module.raw(cu(pts), cu(vd), p12)
    ((a * cu(ga)).sum() + (c * cu(gc)).sum()).backward()
    # (the fp16 tier back-propagates in bf16 -- mixed operand formats do not exist in kind::f16 -- on ReLU masks taken
    #  from its fp16 forward: its distance to the exact network sits between the all-fp16 and the all-bf16 figures)
    tol_model, tol_exact = (0.03, 0.2) if prec == "bf16" else (0.03, 0.12)

    def rel(x, ref):
        return float((x.cpu().double() - ref).norm() / (ref.norm() + 1e-30))

    for k, v in module.named_parameters():
        assert rel(v.grad, pm[prefix + k].grad) < tol_model, (k, rel(v.grad, pm[prefix + k].grad))
        assert rel(v.grad, p64[prefix + k].grad) < tol_exact, (k, rel(v.grad, p64[prefix + k].grad))
    if with_pose:
        assert rel(pose_g.grad, posem.grad) < 2 * tol_model, ("pose", rel(pose_g.grad, posem.grad))
        assert rel(pose_g.grad, pose64.grad) < tol_exact, ("pose", rel(pose_g.grad, pose64.grad))


@pytest.mark.parametrize("prec", ["bf16", "fp32"])
def test_recompute_path_matches_stash_path_and_stash_accounting(prec, monkeypatch):
    """Over the stash budget the backward re-runs the forward chunk by chunk: same kernels, same gradients (the dW
    reduction uses atomics, so equality is to rounding, not bitwise).  The host-side accounting of live stash bytes
    returns to zero once the graph is gone."""
    import gc
    g = load_golden("e2e_online_quat_train")
    grads = {}
    for mode in ("stash", "recompute"):
        monkeypatch.setattr(F_, "STASH_BUDGET_BYTES", (24 << 30) if mode == "stash" else 0)
        net, _ = make_star(2, 24, int(g["chunk"]), False, int(g["seed"]), training=True, precision=prec)
        ro, rd, target = cu(g["rays_o"]), cu(g["rays_d"]), cu(g["target"])
        vd = rd / rd.norm(dim=-1, keepdim=True)
        pose = cu(g["pose"]).clone().requires_grad_(True)
        pts, z = R_.sample_pts(ro, rd, g["near"], g["far"], int(g["Nc"]))
        out = R_.render_star_online(net, pts, vd, z, ro, rd, int(g["Ni"]), pose, u=cu(g["u"]),
                                    z_samples=cu(g["z_samples"]))
        live = F_._STASH_LIVE.get(torch.cuda.current_device(), 0)
        assert (live > 0) == (mode == "stash")
        loss = ((out["rgb0"] - target) ** 2).mean() + ((out["rgb"] - target) ** 2).mean()
        loss.backward()
        grads[mode] = [pose.grad.clone()] + [p.grad.clone() for p in net.parameters()]
        del out, loss
        gc.collect()
        assert F_._STASH_LIVE.get(torch.cuda.current_device(), 0) == 0
    for a, b in zip(grads["stash"], grads["recompute"]):
        assert float((a - b).norm()) <= 1e-4 * float(b.norm()) + 1e-9


def test_fp16_range_guard_reports_overflow():
    """fp16 operands overflow beyond 65504: the kernel flags non-finite raw outputs through its status word and the
    host raises (at the next MLP call at the latest, or on demand) instead of returning garbage silently."""
    net, _ = make_star(0, 8, 4096, False, seed=3, training=False, precision="fp16")
    ro, rd = so.carla_rays(64, seed=1)
    vd = rd / rd.norm(dim=-1, keepdim=True)
    pts, _ = so.sample_pts(ro, rd, 0.03, 0.8, 16)
    with torch.no_grad():
        a, _ = net.static_coarse_nerf.raw(cu(pts), cu(vd), None)
        F_.check_range()                                       # sane weights: nothing flagged
        assert bool(torch.isfinite(a).all())
        net.static_coarse_nerf.pts_net.lin_in.weight.mul_(3.0e4)
        net.static_coarse_nerf.pts_net.lin_in.weight._version  # (mul_ bumped the version -> the packed image is rebuilt)
        a, _ = net.static_coarse_nerf.raw(cu(pts), cu(vd), None)
        with pytest.raises(_capi.StarError):
            F_.check_range()
        F_.check_range()                                       # the flag is cleared once reported
        # the same weights on the bf16 tier stay finite (8-bit exponent)
        net.set_precision("bf16")
        a, _ = net.static_coarse_nerf.raw(cu(pts), cu(vd), None)
        F_.check_range()
        assert bool(torch.isfinite(a).all())


def test_second_backward_through_a_retained_graph():
    """backward(retain_graph=True) followed by a second backward: the stash of a chunk is dropped after its first use, the
    second pass recomputes it -- same gradients (dW reductions use atomics: equal to rounding)."""
    net, _ = make_star(0, 8, 4096, False, seed=3, training=True, precision="fp16")
    ro, rd = so.carla_rays(96, seed=1)
    vd = rd / rd.norm(dim=-1, keepdim=True)
    pts, _ = so.sample_pts(ro, rd, 0.03, 0.8, 24)
    a, c = net.static_coarse_nerf.raw(cu(pts), cu(vd), None)
    loss = (a ** 2).mean() + (c ** 2).mean()
    loss.backward(retain_graph=True)
    g1 = [p.grad.clone() for p in net.static_coarse_nerf.parameters()]
    net.zero_grad()
    loss.backward()
    for x, y in zip(g1, (p.grad for p in net.static_coarse_nerf.parameters())):
        assert float((x - y).norm()) <= 1e-4 * float(x.norm()) + 1e-12


@pytest.mark.parametrize("prec", ["bf16", "fp16"])
def test_forward_is_independent_of_the_tile_schedule(prec):
    """The forward kernel overlaps tiles (lin_in's operand of tile t + 1 is encoded into a ring slot while the view layer
    of tile t runs, the dirs block is rewritten after lin_in's epilogue): a sample's raw outputs must not depend on which
    CTA / position in the CTA's tile sequence it lands on, nor on the kernel variant.  One launch with several tiles per
    CTA (148 CTAs) against launches of the same rays in small groups (one or two tiles per CTA, other CTAs), inference
    against training (stash) variant -- all bit-identical; gradients of the big launch equal the sum of the groups'."""
    net, _ = make_star(1, 8, 4096, False, seed=21, training=True, precision=prec)
    for R, S, dyn in ((1150, 191, False), (1301, 67, True)):
        module = net.dynamic_coarse_nerfs[0] if dyn else net.static_coarse_nerf
        ro, rd = so.carla_rays(R, seed=7)
        vd = rd / rd.norm(dim=-1, keepdim=True)
        pts, _ = so.sample_pts(ro, rd, 0.03, 0.8, S)
        p12 = F_.pose_to_mat12(cu(so.pose7_to_matrix(so.random_poses7(1, seed=9))[0])) if dyn else None
        pts_d, vd_d = cu(pts), cu(vd)
        with torch.no_grad():
            a0, c0 = module.raw(pts_d, vd_d, p12)                     # inference variant, ~8-11 tiles per CTA
        module.zero_grad()
        a, c = module.raw(pts_d, vd_d, p12)                           # training (stash) variant
        ((a ** 2).mean() + (c ** 2).mean()).backward()
        g_big = [p.grad.clone() for p in module.parameters()]
        assert torch.equal(a0, a.detach()) and torch.equal(c0, c.detach()), (R, S)
        module.zero_grad()
        for lo in range(0, R, 97):                                    # 97 rays: 145 / 51 tiles, never the same alignment
            hi = min(R, lo + 97)
            with torch.no_grad():
                a1, c1 = module.raw(pts_d[lo:hi].contiguous(), vd_d[lo:hi].contiguous(), p12)
            assert torch.equal(a1, a0[lo:hi]) and torch.equal(c1, c0[lo:hi]), (R, S, lo)
            a2, c2 = module.raw(pts_d[lo:hi].contiguous(), vd_d[lo:hi].contiguous(), p12)
            (((a2 ** 2).sum() + 0.0) / a.numel() + (c2 ** 2).sum() / c.numel()).backward()
        for x, y in zip(g_big, (p.grad for p in module.parameters())):
            assert float((x - y).norm()) <= 2e-3 * float(x.norm()) + 1e-12
    F_.check_range()


def test_stash_forward_back_to_back_launches_do_not_fault():
    """Regression for a barrier race that only showed on some GPUs and only with kernels queued back to back: with the
    dirs K-block FIRST in the view layer, the 16 epilogue warps announced feature block 3 on the w_full barrier of the ring
    stage that K-block 0 (the dirs block) was still waiting on whenever the producer thread lagged behind its stash
    stores -- an arrival-count underflow, reported as cudaErrorLaunchFailure after 0.5-20 s of training.  ~1500 training
    forwards without a synchronisation in between (tools/stress_tc.py runs the long version)."""
    net, _ = make_star(0, 8, 1 << 20, False, seed=3, training=True, precision="fp16")
    m = net.static_fine_nerf
    R, S = 4096, 192
    ro, rd = so.carla_rays(R, seed=1)
    vd = cu(rd / rd.norm(dim=-1, keepdim=True))
    pts, _ = so.sample_pts(ro, rd, 2.0, 6.0, S)
    pts = cu(pts).contiguous()
    for _ in range(30):
        for _ in range(25):
            a, c = m.raw(pts, vd, None)
            del a, c
        torch.cuda.synchronize()
    F_.check_range()


def test_retired_variant_flags_are_rejected():
    """Bits 0x100 / 0x200 of StarNetDesc.precision selected the CTA-pair and direct-stash A/B variants (removed): an entry
    point must refuse them instead of silently running something else."""
    import ctypes as C
    lib = _capi.lib()
    x = torch.zeros(128, 3, device="cuda")
    out_a, out_c = torch.zeros(128, device="cuda"), torch.zeros(128, 3, device="cuda")
    packed = torch.zeros(1 << 22, dtype=torch.uint8, device="cuda")
    for flag in (0x100, 0x200):
        d = _capi.net_desc(4, 10, 4, _capi.PREC_F16 | flag)
        rc = lib.star_mlp_forward(C.byref(d), packed.data_ptr(), x.data_ptr(), None, None, None, x.data_ptr(), None, None,
                                  None, 1, 128, out_a.data_ptr(), out_c.data_ptr(), 128, None, None, _capi.stream())
        assert rc == 2, rc          # STAR_E_UNSUPPORTED
    torch.cuda.synchronize()


@pytest.mark.parametrize("prec", ["bf16", "fp16"])
def test_weight_sharing_cluster_launch_is_bit_identical(prec, monkeypatch):
    """Inference launches of at least two tiles per SM run as clusters of 2 whose CTAs each fetch half of every weight K-block
    and multicast it to both (STAR_PREC_FLAG_NO_WSHARE opts out): same arithmetic -> the same bits, for even and odd tile
    counts (the odd one rounds up to a ghost tile), static and object nets, through star_mlp_forward and star_render_forward."""
    net, _ = make_star(1, 24, 4096, False, seed=23, training=False, precision=prec)
    for R, S, dyn in ((700, 64, False), (673, 67, False), (673, 67, True), (2051, 19, True)):
        assert R * S >= 2 * 148 * 128
        module = net.dynamic_coarse_nerfs[0] if dyn else net.static_coarse_nerf
        ro, rd = so.carla_rays(R, seed=8)
        vd = rd / rd.norm(dim=-1, keepdim=True)
        pts, _ = so.sample_pts(ro, rd, 0.03, 0.8, S)
        pose = cu(so.pose7_to_matrix(so.random_poses7(1, seed=9))[0])
        res = {}
        for off in (False, True):
            monkeypatch.setattr(F_, "TC_NO_WSHARE", off)
            with torch.no_grad():
                res[off] = module.raw(cu(pts), cu(vd), F_.pose_to_mat12(pose) if dyn else None)
        assert torch.equal(res[False][0], res[True][0]) and torch.equal(res[False][1], res[True][1]), (R, S, dyn)
    ro, rd = so.carla_rays(1500, seed=9)
    ro, rd = cu(ro), cu(rd)
    vd = rd / rd.norm(dim=-1, keepdim=True)
    pose = cu(so.random_poses7(1, seed=3))
    outs = {}
    for off in (False, True):
        monkeypatch.setattr(F_, "TC_NO_WSHARE", off)
        with torch.no_grad():
            pts, z = R_.sample_pts(ro, rd, 0.03, 0.8, 32, is_train=False)
            outs[off] = R_.render_star_online(net, pts, vd, z, ro, rd, 24, pose)
    for k in ("rgb", "rgb0", "weights", "depth", "rgb_dynamic"):
        assert torch.equal(outs[False][k], outs[True][k]), k


@pytest.mark.parametrize("prec", ["bf16", "fp16"])
def test_pipelined_dx_chain_gives_the_gradients_of_the_serial_one(prec, monkeypatch):
    """The opt-in pipelined dX kernel (N = 128 halves, per-block b_done / stash_done barriers) against the default serial one:
    the same GEMMs in the same K order -> the same gradient stash, hence equal weight and pose gradients up to the dW atomics."""
    net, _ = make_star(1, 8, 4096, False, seed=21, training=True, precision=prec)
    for R, S, dyn in ((301, 191, False), (129, 5, True), (640, 64, True)):
        module = net.dynamic_coarse_nerfs[0] if dyn else net.static_coarse_nerf
        ro, rd = so.carla_rays(R, seed=7)
        vd = rd / rd.norm(dim=-1, keepdim=True)
        pts, _ = so.sample_pts(ro, rd, 0.03, 0.8, S)
        res = {}
        for piped in (False, True):
            monkeypatch.setattr(F_, "TC_DX_PIPELINED", piped)
            module.zero_grad()
            pose = cu(so.pose7_to_matrix(so.random_poses7(1, seed=9))[0]).requires_grad_(dyn)
            a, c = module.raw(cu(pts), cu(vd), F_.pose_to_mat12(pose) if dyn else None)
            ((a ** 2).mean() + (c ** 2).mean()).backward()
            res[piped] = [p.grad.clone() for p in module.parameters()] + ([pose.grad.clone()] if dyn else [])
        for x, y in zip(res[False], res[True]):
            assert float((x - y).norm()) <= 1e-4 * float(x.norm()) + 1e-12

"""GPU parity tests of the mip-NeRF variant: CUDA path through the C ABI against
oracle/mip_oracle.py on the same seeded inputs.  The oracle is "parity unpinned" (nerfstudio is not available to the
reference container); tolerances are the north star's fp32 tier: bit-exact frustum edges / searchsorted indices,
1e-4 absolute on rgb / acc / weights."""
import argparse

import pytest
import torch

import star_b200
from star_b200 import mip_functional as MF
from star_b200.models.star_mipnerf import STaR as MipSTaR
from oracle import mip_oracle as mo, star_oracle as so
from helpers import assert_close
from test_gpu_parity import assert_as_accurate

pytestmark = pytest.mark.gpu
DEV = "cuda"
NEAR, FAR = 0.03, 0.8
REGS = ["loss_alpha_entropy", "loss_dynamic_vs_static_reg", "loss_ray_reg", "loss_static_reg", "loss_dynamic_reg"]


def cu(t):
    return t.to(DEV)


def gen(seed):
    return torch.Generator().manual_seed(seed)


def make_net(V, Nc, Ni, chunk, seed, training, gain=2.0):
    args = argparse.Namespace(num_vehicles=V, chunk=chunk, far_dist=1e10, N_importance=Ni, N_samples=Nc,
                              scale_factor=0.01, near=3.0, far=80.0)
    net = MipSTaR(args)
    sd = mo.init_mip_params(V, seed=seed, gain=gain, bias_std=0.02)
    net.load_state_dict(sd, strict=True)
    net.to(DEV).train(training)
    return net, {k: v.clone() for k, v in sd.items()}


def rays(R, seed):
    ro, rd = so.carla_rays(R, seed=seed)
    return ro, rd / rd.norm(dim=-1, keepdim=True)


# ------------------------------------------------------------------------------------------ samplers
@pytest.mark.parametrize("R,Nc", [(1, 1), (5, 64), (1001, 256)])
def test_uniform_bins_bit_exact(R, Nc):
    sp, eu = MF.uniform_bins(R, Nc, NEAR, FAR, DEV)
    ref = mo.uniform_bins(R, Nc)
    assert torch.equal(sp.cpu(), ref.contiguous()) and torch.equal(eu.cpu(), mo.spacing_to_euclidean(ref, NEAR, FAR))
    t = torch.rand(R, Nc + 1, generator=gen(0))
    sp, eu = MF.uniform_bins(R, Nc, NEAR, FAR, DEV, t_rand=cu(t))
    ref = mo.uniform_bins(R, Nc, training=True, t_rand=t)
    assert torch.equal(sp.cpu(), ref) and torch.equal(eu.cpu(), mo.spacing_to_euclidean(ref, NEAR, FAR))


@pytest.mark.parametrize("R,Nc,Ni,training", [(3, 4, 2, False), (33, 64, 128, False), (33, 64, 128, True),
                                              (7, 256, 512, True), (4099, 32, 48, False)])
def test_pdf_sample_bit_exact_vs_defined_arithmetic(R, Nc, Ni, training):
    g = gen(1)
    sp = mo.uniform_bins(R, Nc, training, torch.rand(R, Nc + 1, generator=g) if training else None).contiguous()
    w = torch.rand(R, Nc, generator=g) ** 4
    w[0] = 0.0                                              # a ray without any weight (the eps-padding branch)
    u = torch.rand(R, Ni + 1, generator=g) if training else None
    ref, det = mo.pdf_sample(sp, w, Ni, training, u, return_details=True, exact_sum=True)
    sp2, eu2, d2 = MF.pdf_sample(cu(sp), cu(w), Ni, NEAR, FAR, training=training, u_rand=cu(u) if training else None,
                                 return_details=True)
    assert torch.equal(d2["cdf"].cpu(), det["cdf"])
    assert torch.equal(d2["inds"].cpu(), det["inds"])
    assert torch.equal(sp2.cpu(), ref)
    assert torch.equal(eu2.cpu(), mo.spacing_to_euclidean(ref, NEAR, FAR))
    # strided weights view (the caller passes weights[..., 0] of a [R,S,1] tensor: contiguous; and a column slice)
    wide = torch.cat([w, w], -1)
    sp3, _ = MF.pdf_sample(cu(sp), cu(wide)[:, :Nc], Ni, NEAR, FAR, training=training, u_rand=cu(u) if training else None)
    assert torch.equal(sp3, sp2)


# ------------------------------------------------------------------------------------------ field
def _field_inputs(R, S, seed):
    ro, vd = rays(R, seed)
    t = torch.rand(R, S + 1, generator=gen(seed + 1))
    eu = mo.spacing_to_euclidean(mo.uniform_bins(R, S, True, t), NEAR, FAR).contiguous()
    return ro, vd, eu


@pytest.mark.parametrize("R,S,with_pose", [(3, 5, False), (16, 12, True), (70, 33, False), (9, 64, True)])
def test_mip_field_forward_backward_vs_oracle(R, S, with_pose):
    net, sd = make_net(1, 4, 4, 64, seed=3, training=True, gain=1.4)
    ro, vd, eu = _field_inputs(R, S, 10)
    model = net.dynamic_nerfs[0] if with_pose else net.static_nerf
    prefix = "dynamic_nerfs.0." if with_pose else "static_nerf."
    pose7 = so.random_poses7(1, seed=4)[0]
    ga = torch.randn(R, S, generator=gen(20))
    gc = torch.randn(R, S, 3, generator=gen(21))

    def run_oracle(dtype):
        p = {k: v.to(dtype).clone().requires_grad_(True) for k, v in sd.items() if k.startswith(prefix)}
        pose = pose7.to(dtype).clone().requires_grad_(True)
        o, d = ro.to(dtype), vd.to(dtype)
        if with_pose:
            o, d = so.se3_act(pose, o), so.so3_act(pose[3:], d)
        e = eu.to(dtype)
        a, c = mo.mip_field(p, prefix, o, d, e[:, :-1], e[:, 1:], return_raw=True)
        ((a * ga.to(dtype)).sum() + (c * gc.to(dtype)).sum()).backward()
        return a, c, p, pose

    a32, c32, p32, pose32 = run_oracle(torch.float32)
    pose_g = cu(pose7).requires_grad_(True)
    p12 = star_b200.functional.pose_to_mat12(pose_g) if with_pose else None
    a, c = model.raw(cu(ro), cu(vd), cu(eu), p12)
    assert_close(a, a32, 1e-4, rtol=1e-4, msg="raw density")     # raw (pre-activation) values reach O(10^2) here
    assert_close(c, c32, 1e-4, rtol=1e-4, msg="raw rgb")
    ((a * cu(ga)).sum() + (c * cu(gc)).sum()).backward()
    a64, c64, p64, pose64 = run_oracle(torch.float64)
    names = {"field.mlp_base.layers.0": model.field.mlp_base.layers[0], "field.mlp_base.layers.4": model.field.mlp_base.layers[4],
             "field.mlp_base.layers.7": model.field.mlp_base.layers[7], "field.field_output_density.net": model.field.field_output_density.net,
             "field.mlp_head.layers.0": model.field.mlp_head.layers[0], "field.mlp_head.layers.1": model.field.mlp_head.layers[1],
             "field.field_heads.0.net": model.field.field_heads[0].net}
    errs = []
    for n, lin in names.items():
        for part in ("weight", "bias"):
            k = f"{prefix}{n}.{part}"
            try:
                assert_as_accurate(getattr(lin, part).grad, p32[k].grad, p64[k].grad, k)
            except AssertionError as e:
                errs.append(str(e))
    assert not errs, "\n".join(errs)
    if with_pose:
        assert float(pose_g.grad[6]) == 0.0
        assert_as_accurate(pose_g.grad[:6], pose32.grad[:6], pose64.grad[:6], "pose7 (pypose tangent gradient)")


# ------------------------------------------------------------------------------------------ compositing
@pytest.mark.parametrize("R,S", [(5, 2), (33, 31), (64, 192), (17, 1025)])
def test_mip_composite_single_forward_backward_vs_oracle(R, S):
    g = gen(5)
    rs = (torch.randn(R, S, generator=g) * 3 + 2)
    rc = torch.randn(R, S, 3, generator=g)
    _, _, eu = _field_inputs(R, S, 30)
    gr, ga, gw = torch.randn(R, 3, generator=g), torch.randn(R, generator=g), torch.randn(R, S, generator=g)

    def oracle(dtype):
        a, c = rs.to(dtype).clone().requires_grad_(True), rc.to(dtype).clone().requires_grad_(True)
        e = eu.to(dtype)
        out = mo.appinit_outputs(torch.nn.functional.softplus(a)[..., None], torch.sigmoid(c), (e[:, 1:] - e[:, :-1])[..., None],
                                 e[:, :-1], e[:, 1:])
        ((out["rgb"] * gr.to(dtype)).sum() + (out["acc"][:, 0] * ga.to(dtype)).sum() + (out["weights"][..., 0] * gw.to(dtype)).sum()).backward()
        return out, a.grad, c.grad

    o32, da32, dc32 = oracle(torch.float32)
    a, c = cu(rs).requires_grad_(True), cu(rc).requires_grad_(True)
    rgb, acc, depth, weights = MF.MipCompositeSingle.apply(a, c, cu(eu))
    assert_close(rgb, o32["rgb"], 1e-5, msg="rgb")
    assert_close(acc, o32["acc"][:, 0], 1e-5, msg="acc")
    assert_close(weights, o32["weights"][..., 0], 1e-5, msg="weights")
    _check_median(depth, o32["depth"][:, 0], o32["weights"][..., 0])
    ((rgb * cu(gr)).sum() + (acc * cu(ga)).sum() + (weights * cu(gw)).sum()).backward()
    o64, da64, dc64 = oracle(torch.float64)
    assert_close(a.grad, da64, 2e-5, rtol=1e-4, msg="d raw_sigma")
    assert_close(c.grad, dc64, 2e-5, rtol=1e-4, msg="d raw_rgb")


def _check_median(depth, ref_depth, ref_weights):
    """The median depth is an index decision on cumsum(weights) >= 0.5: rays whose cumulative weight passes within
    1e-5 of 0.5 may legitimately pick the neighbouring sample; all others must agree exactly (up to fp32 of the mid-point)."""
    cw = torch.cumsum(ref_weights.double(), -1)
    near_tie = ((cw - 0.5).abs() < 1e-5).any(-1)
    d, r = depth.detach().cpu(), ref_depth
    ok = (d - r).abs() <= 1e-6
    assert bool((ok | near_tie).all()), f"median depth differs on {int((~(ok | near_tie)).sum())} rays"
    assert float(near_tie.float().mean()) < 0.05


@pytest.mark.parametrize("R,V,S,chunk", [(7, 1, 40, 100), (40, 3, 48, 16), (33, 5, 130, 7), (9, 8, 33, 4)])
def test_mip_composite_multi_forward_backward_vs_oracle(R, V, S, chunk):
    g = gen(6)
    rs_s = torch.randn(R, S, generator=g) * 2 + 1
    rc_s = torch.randn(R, S, 3, generator=g)
    rs_d = torch.randn(R, V, S, generator=g) * 2
    rc_d = torch.randn(R, V, S, 3, generator=g)
    _, _, eu = _field_inputs(R, S, 40)
    gr, ga, gw = torch.randn(R, 3, generator=g), torch.randn(R, generator=g), torch.randn(R, S, generator=g)
    lam = torch.tensor([0.7, -0.4, 0.3, 0.9, 0.5])

    def oracle(dtype):
        ins = [t.to(dtype).clone().requires_grad_(True) for t in (rs_s, rc_s, rs_d, rc_d)]
        e = eu.to(dtype)
        loss, outs = 0.0, None
        for i in range(0, R, chunk):          # scalars: per-chunk means summed over chunks
            j = min(R, i + chunk)
            o = mo.online_outputs(torch.nn.functional.softplus(ins[0][i:j])[..., None], torch.sigmoid(ins[1][i:j]),
                                  torch.nn.functional.softplus(ins[2][i:j])[..., None], torch.sigmoid(ins[3][i:j]),
                                  (e[i:j, 1:] - e[i:j, :-1])[..., None], e[i:j, :-1], e[i:j, 1:])
            loss = loss + (o["rgb"] * gr[i:j].to(dtype)).sum() + (o["acc"][:, 0] * ga[i:j].to(dtype)).sum() + \
                (o["weights"][..., 0] * gw[i:j].to(dtype)).sum() + sum(lam[k].to(dtype) * o[REGS[k]] for k in range(5))
            outs = o if outs is None else {k: (outs[k] + v if v.dim() == 0 else torch.cat([outs[k], v], 0)) for k, v in o.items()}
        loss.backward()
        return outs, [t.grad for t in ins]

    o32, g32 = oracle(torch.float32)
    ins = [cu(t).requires_grad_(True) for t in (rs_s, rc_s, rs_d, rc_d)]
    out = dict(zip(MF.MIP_OUT_KEYS, MF.MipCompositeStar.apply(*ins, cu(eu), chunk)))
    for k in ("rgb", "rgb_static", "rgb_dynamic"):
        assert_close(out[k], o32[k], 1e-5, msg=k)
    assert_close(out["acc"], o32["acc"][:, 0], 1e-5, msg="acc")
    assert_close(out["weights"], o32["weights"][..., 0], 1e-5, msg="weights")
    assert_close(out["dynamic_transmittance"], o32["dynamic_transmittance"][..., 0], 1e-5, msg="dynamic_transmittance")
    _check_median(out["depth"], o32["depth"][:, 0], o32["weights"][..., 0])
    for k in range(5):
        assert_close(out["regs"][k], o32[REGS[k]], 2e-6, rtol=2e-5, msg=REGS[k])
    loss = (out["rgb"] * cu(gr)).sum() + (out["acc"] * cu(ga)).sum() + (out["weights"] * cu(gw)).sum() + (out["regs"] * cu(lam)).sum()
    loss.backward()
    o64, g64 = oracle(torch.float64)
    for t, ref, name in zip(ins, g64, ("d raw_sigma_s", "d raw_rgb_s", "d raw_sigma_d", "d raw_rgb_d")):
        assert_close(t.grad, ref, 3e-5, rtol=2e-4, msg=name)


# ------------------------------------------------------------------------------------------ end to end
@pytest.mark.parametrize("V,training", [(0, False), (0, True), (2, False), (2, True)])
def test_mip_star_forward_end_to_end(V, training):
    R, Nc, Ni = 37, 16, 24
    net, sd = make_net(V, Nc, Ni, chunk=16, seed=7, training=training)
    ro, vd = rays(R, 50)
    pose = so.random_poses7(V, seed=8) if V else None
    t_rand = torch.rand(R, Nc + 1, generator=gen(60))
    u_rand = torch.rand(R, Ni + 1, generator=gen(61))
    cfg = mo.MipConfig(num_vehicles=V, N_samples=Nc, N_importance=Ni, chunk=16)
    det = {}
    ref = mo.star_mip_forward(sd, cfg, ro, vd, pose=pose, training=training, t_rand=t_rand, u_rand=u_rand, details=det,
                              exact_sum=True)
    out = net(cu(ro), cu(vd), cu(pose) if V else None, t_rand=cu(t_rand), u_rand=cu(u_rand))
    assert set(out.keys()) == set(ref.keys())
    from star_b200.models import types__ as T_      # the reference's TypedDicts (models/types__.py:36-87)
    assert set(out.keys()) == set(T_.MIP_ONLINE_KEYS if V else T_.MIP_APPINIT_KEYS)
    for k, v in ref.items():
        assert out[k].shape == v.shape, (k, out[k].shape, v.shape)
    # coarse pass: identical frustums -> tight; fine pass: the PDF sampler amplifies 1e-7 weight differences
    # (same conditioning argument as sample_pdf) -> 5e-3 on per-sample weights, tight on integrals
    for k in ("rgb0", "acc0", "weights0"):
        assert_close(out[k], ref[k], 1e-4, msg=k)
    for k in ("rgb", "acc"):
        assert_close(out[k], ref[k], 2e-3, msg=k)
    if V:
        for k in ("rgb_static0", "rgb_dynamic0", "dynamic_transmittance0"):
            assert_close(out[k], ref[k], 1e-4, msg=k)
        for k in REGS:
            assert_close(out[k + "0"], ref[k + "0"], 1e-5, rtol=1e-4, msg=k + "0")
            assert_close(out[k], ref[k], 1e-3, rtol=5e-3, msg=k)
    _check_median(out["depth0"][:, 0], ref["depth0"][:, 0], ref["weights0"][..., 0])


def test_mip_star_training_step_gradients_reach_weights_and_pose():
    R, Nc, Ni, V = 24, 8, 8, 2
    net, sd = make_net(V, Nc, Ni, chunk=1 << 20, seed=9, training=True)
    ro, vd = rays(R, 70)
    pose7 = so.random_poses7(V, seed=10)
    t_rand = torch.rand(R, Nc + 1, generator=gen(80))
    u_rand = torch.rand(R, Ni + 1, generator=gen(81))
    target = torch.rand(R, 3, generator=gen(82))
    cfg = mo.MipConfig(num_vehicles=V, N_samples=Nc, N_importance=Ni, chunk=1 << 20)

    def oracle(dtype):
        p = {k: v.to(dtype).clone().requires_grad_(True) for k, v in sd.items()}
        pose = pose7.to(dtype).clone().requires_grad_(True)
        o = mo.star_mip_forward(p, cfg, ro.to(dtype), vd.to(dtype), pose=pose, training=True, t_rand=t_rand.to(dtype),
                                u_rand=u_rand.to(dtype), exact_sum=True)
        loss = ((o["rgb"] - target.to(dtype)) ** 2).mean() + 0.1 * ((o["rgb0"] - target.to(dtype)) ** 2).mean() + \
            1e-3 * o["loss_alpha_entropy"] + 1e-3 * o["loss_dynamic_vs_static_reg0"] + 1e-5 * o["loss_ray_reg"]
        loss.backward()
        return loss, p, pose

    l32, p32, pose32 = oracle(torch.float32)
    pose_g = cu(pose7).requires_grad_(True)
    o = net(cu(ro), cu(vd), pose_g, t_rand=cu(t_rand), u_rand=cu(u_rand))
    tg = cu(target)
    loss = ((o["rgb"] - tg) ** 2).mean() + 0.1 * ((o["rgb0"] - tg) ** 2).mean() + 1e-3 * o["loss_alpha_entropy"] + \
        1e-3 * o["loss_dynamic_vs_static_reg0"] + 1e-5 * o["loss_ray_reg"]
    loss.backward()
    assert abs(float(loss) - float(l32)) < 2e-3
    l64, p64, pose64 = oracle(torch.float64)
    # the fine pass sits on slightly different frustums than the oracle's (PDF-sampler conditioning), so gradients are
    # compared at the 5 % level here; the stage-wise tests above are the tight ones
    for k in ("static_nerf.field.mlp_base.layers.0.weight", "dynamic_nerfs.1.field.mlp_head.layers.1.weight"):
        mod = net
        for part in k.split(".")[:-1]:
            mod = mod[int(part)] if part.isdigit() else getattr(mod, part)
        gpu, r64 = mod.weight.grad.cpu().double(), p64[k].grad
        assert float((gpu - r64).norm() / r64.norm()) < 5e-2, k
    assert float((pose_g.grad.cpu().double() - pose64.grad).norm() / pose64.grad.norm()) < 5e-2
    assert float(pose_g.grad[:, 6].abs().max()) == 0.0


def test_mip_empty_ray_batch_gives_empty_outputs():
    """R = 0 through the mip variant's forward (app-init branch, both tiers): empty outputs of the right shapes, no launch."""
    net, _ = make_net(0, 16, 24, 4096, seed=3, training=False)
    ro = torch.zeros(0, 3, device=DEV)
    for prec in ("fp32", "fp16"):
        net.set_precision(prec)
        with torch.no_grad():
            out = net(ro, ro, None)
        assert out["rgb"].shape == (0, 3) and out["rgb0"].shape == (0, 3)
        assert out["weights"].shape[0] == 0 and out["acc"].shape[0] == 0
    torch.cuda.synchronize()


def test_mip_error_paths():
    net, _ = make_net(1, 4, 4, 64, seed=1, training=False)
    ro, vd = rays(3, 1)
    with pytest.raises(NotImplementedError):
        net(cu(ro), cu(vd), torch.eye(4, device=DEV)[None])
    with pytest.raises(star_b200._capi.StarError):
        net(ro, vd)          # CPU tensors: there is no CPU path


# ------------------------------------------------------------------------------------------ tensor-core tier
@pytest.mark.parametrize("prec", ["bf16", "fp16"])
@pytest.mark.parametrize("R,S,with_pose", [(4, 32, False), (70, 33, False), (16, 12, True), (301, 191, False), (1, 1, True)])
def test_mip_tc_field_matches_rounding_model(R, S, with_pose, prec):
    """tcgen05 tier of the mip field against (1) the oracle's operand-rounding model of the kernel (GEMM operands --
    encodings, activations, weights -- rounded to 16 bits, fp32 accumulate / bias / heads) and (2) the fp32 oracle."""
    net, sd = make_net(1, 4, 4, 64, seed=3, training=False, gain=1.4)
    net.set_precision(prec)
    ro, vd, eu = _field_inputs(R, S, 10)
    model = net.dynamic_nerfs[0] if with_pose else net.static_nerf
    prefix = "dynamic_nerfs.0." if with_pose else "static_nerf."
    pose7 = so.random_poses7(1, seed=4)[0]
    p = {k: v for k, v in sd.items() if k.startswith(prefix)}
    o, d = (so.se3_act(pose7, ro), so.so3_act(pose7[3:], vd)) if with_pose else (ro, vd)
    a_m, c_m = mo.mip_field(p, prefix, o, d, eu[:, :-1], eu[:, 1:], emulate=prec, return_raw=True)
    a_f, c_f = mo.mip_field(p, prefix, o, d, eu[:, :-1], eu[:, 1:], return_raw=True)
    with torch.no_grad():
        p12 = star_b200.functional.pose_to_mat12(cu(pose7)) if with_pose else None
        a, c = model.raw(cu(ro), cu(vd), cu(eu), p12)
    scale = float(a_f.abs().max()) + float(c_f.abs().max()) + 1e-3
    ea, ec = (a.cpu() - a_m).abs() / scale, (c.cpu() - c_m).abs() / scale
    assert float(ea.mean()) < 1e-3 and float(ec.mean()) < 1e-3, (float(ea.mean()), float(ec.mean()))
    assert float(ea.max()) < 6e-2 and float(ec.max()) < 6e-2, (float(ea.max()), float(ec.max()))
    med = 1e-4 if prec == "bf16" else 1e-3
    assert float(ea.median()) < med and float(ec.median()) < med, (float(ea.median()), float(ec.median()))
    assert float((a.cpu() - a_f).abs().mean()) < 2e-2 * scale
    assert float((c.cpu() - c_f).abs().mean()) < 2e-2 * scale


@pytest.mark.parametrize("prec", ["bf16", "fp16"])
def test_mip_tc_end_to_end_render_close_to_fp32(prec):
    R, Nc, Ni, V = 64, 32, 48, 2
    net, sd = make_net(V, Nc, Ni, chunk=1 << 20, seed=7, training=False, gain=1.4)
    ro, vd = rays(R, 50)
    pose = so.random_poses7(V, seed=8)
    with torch.no_grad():
        ref = net(cu(ro), cu(vd), cu(pose))
        net.set_precision(prec)
        out = net(cu(ro), cu(vd), cu(pose))
    # 16-bit operands: rgb within the north star's 2e-3 in the mean (bf16) / in the max norm (fp16)
    err = (out["rgb0"] - ref["rgb0"]).abs()
    assert float(err.mean()) < 2e-3 and float(err.max()) < (2e-3 if prec == "fp16" else 2e-2), (float(err.mean()), float(err.max()))
    assert float((out["rgb"] - ref["rgb"]).abs().mean()) < 4e-3
    assert psnr_shift(out["rgb0"], ref["rgb0"]) < 0.05


def psnr_shift(a, b):
    """|PSNR(a, t) - PSNR(b, t)| against a fixed random target t."""
    t = torch.rand(a.shape, generator=gen(99)).to(a.device)
    pa = -10 * torch.log10(((a - t) ** 2).mean())
    pb = -10 * torch.log10(((b - t) ** 2).mean())
    return float((pa - pb).abs())


@pytest.mark.parametrize("prec", ["bf16", "fp16"])
@pytest.mark.parametrize("R,S,with_pose", [(8, 32, False), (70, 33, False), (40, 48, True), (1250, 64, False)])
def test_mip_tc_backward_matches_rounding_model_and_fp64(R, S, with_pose, prec):
    """Tensor-core backward of the mip field (stash forward + dX chain + dW GEMMs + head gradients), weight gradients.
    Two fp64 autograd references, as for the vanilla MLP: the oracle's operand-rounding model with straight-through
    gradients (what remains is the 16-bit rounding of the back-propagated gradients) and the exact network (ReLU masks
    flip under 16-bit operands).  with_pose: an object field whose pose does NOT require grad -- weights only."""
    net, sd = make_net(1, 4, 4, 64, seed=3, training=True, gain=1.4)
    net.set_precision(prec)
    ro, vd, eu = _field_inputs(R, S, 10)
    model = net.dynamic_nerfs[0] if with_pose else net.static_nerf
    prefix = "dynamic_nerfs.0." if with_pose else "static_nerf."
    pose7 = so.random_poses7(1, seed=4)[0]
    ga = torch.randn(R, S, generator=gen(20))
    gc = torch.randn(R, S, 3, generator=gen(21))
    dt = torch.float64

    def run_oracle(emulate):
        p = {k: v.to(dt).clone().requires_grad_(True) for k, v in sd.items() if k.startswith(prefix)}
        o, d = ro.to(dt), vd.to(dt)
        if with_pose:
            o, d = so.se3_act(pose7.to(dt), o), so.so3_act(pose7[3:].to(dt), d)
        e = eu.to(dt)
        a, c = mo.mip_field(p, prefix, o, d, e[:, :-1], e[:, 1:], emulate=emulate, return_raw=True)
        ((a * ga.to(dt)).sum() + (c * gc.to(dt)).sum()).backward()
        return p

    p64, pm = run_oracle(False), run_oracle(prec)
    p12 = star_b200.functional.pose_to_mat12(cu(pose7)) if with_pose else None     # no grad wanted for the pose
    a, c = model.raw(cu(ro), cu(vd), cu(eu), p12)
    ((a * cu(ga)).sum() + (c * cu(gc)).sum()).backward()
    tol_model, tol_exact = (0.04, 0.25) if prec == "bf16" else (0.03, 0.12)

    def rel(x, ref):
        return float((x.cpu().double() - ref).norm() / (ref.norm() + 1e-30))

    for k, v in model.named_parameters():
        assert v.grad is not None, k
        assert rel(v.grad, pm[prefix + k].grad) < tol_model, (k, rel(v.grad, pm[prefix + k].grad))
        assert rel(v.grad, p64[prefix + k].grad) < tol_exact, (k, rel(v.grad, p64[prefix + k].grad))


def test_mip_tc_training_with_pose_gradient_runs_the_fp32_kernels_loudly():
    """The tensor-core backward of the mip field has no gradient of the ray: an object field whose pose requires grad
    runs its training pass on the fp32 kernels (with a RuntimeWarning, once) -- same gradients as the fp32 tier; the
    static field of the same model still trains on the tensor cores."""
    import warnings
    from star_b200 import mip_functional as MF_
    MF_._WARNED_FP32_TRAINING = False
    net, _ = make_net(1, 8, 8, 64, seed=11, training=True, gain=1.4)
    ro, vd = rays(8, 3)
    t = torch.rand(8, 9, generator=gen(1))
    u = torch.rand(8, 9, generator=gen(2))
    outs = {}
    for prec in ("fp32", "bf16"):
        net.set_precision(prec)
        net.zero_grad()
        pose = cu(so.random_poses7(1, seed=8)).requires_grad_(True)
        with warnings.catch_warnings(record=True) as w:
            warnings.simplefilter("always")
            o = net(cu(ro), cu(vd), pose, t_rand=cu(t), u_rand=cu(u))
            o["rgb"].sum().backward()
        outs[prec] = (net.dynamic_nerfs[0].field.mlp_base.layers[0].weight.grad.clone(), pose.grad.clone(),
                      net.static_nerf.field.mlp_base.layers[0].weight.grad.clone(),
                      [x for x in w if issubclass(x.category, RuntimeWarning)])
    assert len(outs["fp32"][3]) == 0 and len(outs["bf16"][3]) >= 1
    # the object field ran the same fp32 kernels; its upstream gradients differ only through the static field's bf16 forward
    for i in (0, 1):
        a, b = outs["bf16"][i], outs["fp32"][i]
        assert float((a - b).norm() / b.norm()) < 0.2
    assert float(outs["bf16"][2].abs().max()) > 0

"""Generates tests/golden/*.npz by running the UNMODIFIED reference (/root/reference, imported
through oracle/ref_harness.py with stub modules) on small seeded inputs.

Run in the build container only (the reference tree does not exist on the GPU box):
    python tools/make_golden.py
The fixtures are committed; tests/test_oracle_golden.py pins oracle/star_oracle.py to them and the
-m gpu tests pin the CUDA path to them.

Weights are not stored: they are re-created from a seed by oracle.star_oracle.init_star_params
(deterministic CPU generator) and loaded into the reference STaR with load_state_dict(strict=True),
which also pins the checkpoint key layout.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_harness, star_oracle as so  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def npz(name, **kw):
    arrs = {}
    for k, v in kw.items():
        if v is None:
            continue
        if torch.is_tensor(v):
            v = v.detach().cpu().numpy()
        arrs[k] = np.asarray(v)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **arrs)
    print("wrote", path, "%.1f KB" % (os.path.getsize(path) / 1024))


def flat_outputs(prefix, d):
    return {prefix + k: v for k, v in d.items() if v is not None}


def grad_digest(named_params):
    """Compact pin of weight gradients: per tensor [l2 norm, sum, first 8 flat values]."""
    out = {}
    for k, p in named_params:
        g = p.grad
        if g is None:
            continue
        f = g.reshape(-1)
        out["gd." + k] = torch.cat([f.norm()[None], f.sum()[None], f[:8] if f.numel() >= 8 else
                                    torch.cat([f, torch.zeros(8 - f.numel())])])
    return out


def rays_fixture(ref):
    """get_rays / get_rays_np (rendering__.py:41-71) on a non-square view with an off-centre principal point."""
    import math
    H, W = 37, 53
    focal = 0.5 * W / math.tan(0.5 * 0.6911112)
    K = torch.tensor([[focal, 0, 0.5 * W], [0, focal * 1.01, 0.5 * H - 0.25], [0, 0, 1.0]])
    g = torch.Generator().manual_seed(5)
    q, _ = torch.linalg.qr(torch.randn(3, 3, generator=g))
    c2w = torch.cat([q, torch.randn(3, 1, generator=g)], 1)
    ro, rd = ref.rendering.get_rays(H, W, K, c2w)
    ro_np, rd_np = ref.rendering.get_rays_np(H, W, K.numpy(), c2w.numpy())
    npz("get_rays", H=H, W=W, K=K, c2w=c2w, rays_o=ro, rays_d=rd, rays_o_np=np.ascontiguousarray(ro_np), rays_d_np=rd_np)


def degenerate_inputs():
    """Rays on which the reference's clamps and guards bite (shared with tests/test_gpu_parity.py through the fixture):
    nothing on the ray (all alphas 0 -> disp = 1 / max(1e-10, 0 / 0)), transparent static + faint objects, an object opaque at
    its first sample, saturated densities, repeated depths, a ray of one depth, saturated colours."""
    gen = torch.Generator().manual_seed(21)
    R, V, S = 16, 3, 40
    ras = torch.randn(R, S, generator=gen) * 3 - 1
    rcs = torch.randn(R, S, 3, generator=gen) * 2
    rad = torch.randn(R, V, S, generator=gen) * 3 - 2
    rcd = torch.randn(R, V, S, 3, generator=gen) * 2
    ras[0], rad[0] = -80.0, -80.0
    ras[1], rad[1] = -80.0, -30.0
    rad[2, 1, 0] = 1e4
    ras[3] = 80.0
    rad[4] = 80.0
    ras[5], rad[5] = 0.0, 0.0
    rcs[8], rcs[9] = 60.0, -60.0
    ro, rd = so.carla_rays(R, seed=8)
    _, z = so.sample_pts(ro, rd, 0.03, 0.8, S)
    z = z.clone()
    z[6, 3:30] = z[6, 3:4]
    z[7] = z[7, :1]
    return ras, rcs, rad, rcd, z.contiguous(), rd


def degenerate_fixture(ref):
    """raw2outputs (white / black) and raw2outputs_star (test=True) of the UNMODIFIED reference on degenerate_inputs()."""
    R_ = ref.rendering
    ras, rcs, rad, rcd, z, rd = degenerate_inputs()
    o_w = R_.raw2outputs(ras, rcs, z, rd, 0.0, True, 1e10)
    o_b = R_.raw2outputs(ras, rcs, z, rd, 0.0, False, 1e10)
    s_te = R_.raw2outputs_star(ras, rcs, rad, rcd, z, rd, 0, False, 1e10, test=True)
    npz("degenerate", raw_alpha_s=ras, raw_rgb_s=rcs, raw_alpha_d=rad, raw_rgb_d=rcd, z_vals=z, rays_d=rd,
        **flat_outputs("white.", o_w), **flat_outputs("black.", o_b), **flat_outputs("star.", s_te))


def main():
    os.makedirs(OUT, exist_ok=True)
    if "--only-degenerate" in sys.argv:   # added after the other fixtures were committed: leaves them untouched
        degenerate_fixture(ref_harness.load_reference())
        return
    if "--only-rays" in sys.argv:      # added after the other fixtures were committed: leaves them untouched
        rays_fixture(ref_harness.load_reference())
        return
    torch.set_num_threads(8)
    torch.set_float32_matmul_precision("highest")
    ref = ref_harness.load_reference()
    R_ = ref.rendering

    # ---- embedder (embedder.py:81-112) incl. BARF mask quirk
    g = torch.Generator().manual_seed(11)
    x = (torch.rand(37, 3, generator=g) * 2 - 1) * 1.5
    e10, _ = ref.embedder.get_embedder(10, -1, 0)
    e4, _ = ref.embedder.get_embedder(4, -1, 0)
    b10, _ = ref.embedder.get_embedder(10, 40, 0)
    b4, _ = ref.embedder.get_embedder(4, 40, 0)
    npz("embed", x=x, enc10=e10(x), enc4=e4(x), barf10_s13=b10(x, step=13), barf4_s13=b4(x, step=13),
        barf10_s0=b10(x, step=0), barf10_s99=b10(x, step=99))

    # ---- sample_pts (rendering__.py:75-112)
    ro, rd = so.carla_rays(48, seed=5)
    pts, z = R_.sample_pts(ro, rd, 0.03, 0.8, 32, perturb=0, lindisp=False, is_train=True)
    pts_l, z_l = R_.sample_pts(ro, rd, 0.03, 0.8, 32, perturb=0, lindisp=True, is_train=False)
    torch.manual_seed(77)
    pts_p, z_p = R_.sample_pts(ro, rd, 0.03, 0.8, 32, perturb=1.0, lindisp=False, is_train=True)
    torch.manual_seed(77)
    t_rand = torch.rand(48, 32)
    npz("sample_pts", rays_o=ro, rays_d=rd, near=0.03, far=0.8, pts=pts, z=z, pts_lindisp=pts_l,
        z_lindisp=z_l, t_rand=t_rand, pts_perturb=pts_p, z_perturb=z_p)

    # ---- raw2outputs / raw2outputs_star / sample_pdf on random raw values
    g = torch.Generator().manual_seed(21)
    R, S, V = 40, 48, 3
    ra = torch.randn(R, S, generator=g) * 3 - 1
    rc = torch.randn(R, S, 3, generator=g) * 2
    rad = torch.randn(R, V, S, generator=g) * 3 - 2
    rcd = torch.randn(R, V, S, 3, generator=g) * 2
    ro, rd = so.carla_rays(R, seed=6)
    _, z = R_.sample_pts(ro, rd, 0.03, 0.8, S)
    z = z.contiguous()
    o_w = R_.raw2outputs(ra, rc, z, rd, 0.0, True, 1e10)
    o_b = R_.raw2outputs(ra, rc, z, rd, 0.0, False, 1e10)
    npz("raw2outputs", raw_alpha=ra, raw_rgb=rc, z_vals=z, rays_d=rd,
        **flat_outputs("white.", o_w), **flat_outputs("black.", o_b))
    s_tr = R_.raw2outputs_star(ra, rc, rad, rcd, z, rd, 0, False, 1e10, test=False)
    s_te = R_.raw2outputs_star(ra, rc, rad, rcd, z, rd, 0, True, 1e10, test=True)
    npz("raw2outputs_star", raw_alpha_s=ra, raw_rgb_s=rc, raw_alpha_d=rad, raw_rgb_d=rcd, z_vals=z,
        rays_d=rd, **flat_outputs("train.", s_tr), **flat_outputs("test.", s_te))

    # sample_pdf: capture the internal cdf / u / inds by wrapping torch.searchsorted (non-invasive)
    rec = {}
    real_ss = torch.searchsorted

    def spy(cdf, u, right=False, **kw):
        out = real_ss(cdf, u, right=right, **kw)
        rec["cdf"], rec["u"], rec["inds"] = cdf.clone(), u.clone(), out.clone()
        return out

    w = o_b["weights"]
    mid = 0.5 * (z[..., 1:] + z[..., :-1])
    torch.searchsorted = spy
    try:
        s_det = R_.sample_pdf(mid, w[..., 1:-1], 64, det=True)
        det_rec = dict(rec)
        torch.manual_seed(5)
        s_rnd = R_.sample_pdf(mid, w[..., 1:-1], 64, det=False)
        rnd_rec = dict(rec)
    finally:
        torch.searchsorted = real_ss
    npz("sample_pdf", bins=mid, weights=w[..., 1:-1], samples_det=s_det, cdf_det=det_rec["cdf"],
        inds_det=det_rec["inds"], u_rnd=rnd_rec["u"], samples_rnd=s_rnd, cdf_rnd=rnd_rec["cdf"],
        inds_rnd=rnd_rec["inds"])

    # ---- NeRF MLP raw outputs (nerf.py:112-179, resnet.py)
    def ref_star(V, N_importance, chunk, white, seed, end_barf=-1, bias_std=0.02):
        args = ref_harness.make_args(num_vehicles=V, N_importance=N_importance, chunk=chunk,
                                     white_bkgd=white, end_barf=end_barf)
        net = ref.star.STaR(args)
        sd = so.init_star_params(V, N_importance, seed=seed, bias_std=bias_std)
        net.load_state_dict(sd, strict=True)
        return net, args

    net, _ = ref_star(1, 16, 4096, False, seed=3)
    ro, rd = so.carla_rays(24, seed=7)
    vd = rd / rd.norm(dim=-1, keepdim=True)
    pts, z = R_.sample_pts(ro, rd, 0.03, 0.8, 12)
    net.eval()
    with torch.no_grad():
        a_s, c_s = net.static_coarse_nerf(pts, vd, step=None)
        a_d, c_d = net.dynamic_fine_nerfs[0](pts, vd, step=None)
    npz("nerf_mlp", seed=3, V=1, pts=pts, viewdirs=vd, raw_alpha_static_coarse=a_s, raw_rgb_static_coarse=c_s,
        raw_alpha_dynamic_fine0=a_d, raw_rgb_dynamic_fine0=c_d)

    # every end-to-end case also records the fine samples the reference drew (the return value of its
    # sample_pdf), so that the fine pass can be checked on identical sample positions
    zrec = {}
    real_sample_pdf = R_.sample_pdf

    def sample_pdf_spy(*a, **kw):
        out = real_sample_pdf(*a, **kw)
        zrec["z_samples"] = out.detach().clone()
        return out

    R_.sample_pdf = sample_pdf_spy

    # ---- end to end: app-init eval (C1-shaped, tiny), white bkgd, det sampling
    net, args = ref_star(0, 24, 4096, True, seed=4)
    net.eval()
    ro, rd = so.lego_rays(6, 6)
    vd = rd / rd.norm(dim=-1, keepdim=True)
    pts, z = R_.sample_pts(ro, rd, 2.0, 6.0, 16, perturb=0, is_train=False)
    with torch.no_grad():
        out = R_.render_star_appinit(net, pts, vd, z, ro, rd, 24)
    npz("e2e_appinit_eval", seed=4, rays_o=ro, rays_d=rd, near=2.0, far=6.0, Nc=16, Ni=24, z_samples=zrec["z_samples"], **flat_outputs("", out))

    # ---- end to end: app-init TRAIN (random u, grads to weights), chunk < R
    net, args = ref_star(0, 24, 10, False, seed=5)
    net.train()
    ro, rd = so.carla_rays(28, seed=8)
    vd = rd / rd.norm(dim=-1, keepdim=True)
    pts, z = R_.sample_pts(ro, rd, 0.03, 0.8, 16)
    g = torch.Generator().manual_seed(9)
    target = torch.rand(28, 3, generator=g)
    torch.manual_seed(123)
    out = R_.render_star_appinit(net, pts, vd, z, ro, rd, 24)
    torch.manual_seed(123)
    u = torch.rand(28, 24)
    loss = ((out["rgb0"] - target) ** 2).mean() + ((out["rgb"] - target) ** 2).mean() + 0.1 * out["depth"].mean()
    loss.backward()
    npz("e2e_appinit_train", seed=5, chunk=10, rays_o=ro, rays_d=rd, near=0.03, far=0.8, Nc=16, Ni=24, u=u,
        target=target, loss=loss, z_samples=zrec["z_samples"], **flat_outputs("", out),
        **grad_digest(net.named_parameters()))

    # ---- end to end: online, V=2, 4x4 pose, train mode (regularisers + pose grads), two ray chunks
    lam = (1e-3, 1e-3, 1e-5, 1e-4, 1e-4)
    regs = ["loss_alpha_entropy", "loss_dynamic_vs_static_reg", "loss_ray_reg", "loss_static_reg", "loss_dynamic_reg"]

    def online_case(name, pose_fn, training, chunk, seed):
        net, args = ref_star(2, 24, chunk, False, seed=seed)
        net.train(training)
        ro, rd = so.carla_rays(20, seed=10)
        vd = rd / rd.norm(dim=-1, keepdim=True)
        pts, z = R_.sample_pts(ro, rd, 0.03, 0.8, 16)
        p7 = so.random_poses7(2, seed=3)
        pose = pose_fn(p7).clone().requires_grad_(True)
        g = torch.Generator().manual_seed(12)
        target = torch.rand(20, 3, generator=g)
        torch.manual_seed(321)
        out = R_.render_star_online(net, pts, vd, z, ro, rd, 24, pose, step=None)
        torch.manual_seed(321)
        u = torch.rand(20, 24)
        extra = {}
        if training:
            loss = ((out["rgb0"] - target) ** 2).mean() + ((out["rgb"] - target) ** 2).mean()
            for l, k in zip(lam, regs):
                loss = loss + l * 0.5 * (out[k] + out[k + "0"])
            loss.backward()
            extra = dict(loss=loss, pose_grad=pose.grad, **grad_digest(net.named_parameters()))
        npz(name, seed=seed, chunk=chunk, rays_o=ro, rays_d=rd, near=0.03, far=0.8, Nc=16, Ni=24, u=u,
            target=target, pose=pose, z_samples=zrec["z_samples"], **flat_outputs("", out), **extra)

    online_case("e2e_online_mat_train", so.pose7_to_matrix, True, 12, seed=6)
    with torch.no_grad():
        online_case("e2e_online_mat_eval", so.pose7_to_matrix, False, 4096, seed=6)
    # 7-vector pose: goes through the pypose STUB (third-party semantics restated; parity unpinned)
    online_case("e2e_online_quat_train", lambda p: p, True, 4096, seed=6)
    rays_fixture(ref)


if __name__ == "__main__":
    main()

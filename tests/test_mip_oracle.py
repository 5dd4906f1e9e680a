"""CPU checks of oracle/mip_oracle.py (the mip-NeRF restatement).  nerfstudio is not available,
so the oracle is "parity unpinned"; these tests tie it to what IS pinned (the vanilla oracle, itself checked against
the unmodified reference) and to closed-form / brute-force evaluations of the published formulas."""
import math

import torch

from oracle import mip_oracle as mo, star_oracle as so


def test_layer_shapes_and_mac_count():
    shp = mo.mip_param_shapes()
    macs = sum(o * i for (o, i) in shp.values())
    assert macs == 587264
    assert mo.D_XYZ == 147 and mo.D_DIR == 27
    sd = mo.init_mip_params(2)
    assert len(sd) == 3 * 2 * len(shp)
    assert sd["dynamic_nerfs.1.field.mlp_base.layers.4.weight"].shape == (256, 147 + 256)


def test_uniform_bins_and_collider():
    b = mo.uniform_bins(3, 8)
    assert torch.equal(b[0], torch.linspace(0, 1, 9)) and b.shape == (3, 9)
    t = torch.rand(3, 9, generator=torch.Generator().manual_seed(0))
    bj = mo.uniform_bins(3, 8, training=True, t_rand=t)
    assert (bj[:, 1:] >= bj[:, :-1]).all() and bj.min() >= 0 and bj.max() <= 1
    e = mo.spacing_to_euclidean(b, 0.03, 0.8)
    assert abs(float(e[0, 0]) - 0.03) < 1e-7 and abs(float(e[0, -1]) - 0.8) < 1e-7


def test_frustum_gaussian_matches_full_covariance_and_sampling():
    g = torch.Generator().manual_seed(1)
    o = torch.randn(4, 3, generator=g)
    d = torch.nn.functional.normalize(torch.randn(4, 3, generator=g), dim=-1)
    st = torch.rand(4, 5, generator=g) * 0.5 + 0.1
    en = st + torch.rand(4, 5, generator=g) * 0.2 + 0.01
    r = 0.05
    mean, diag, (tm, dv, rv) = mo.frustum_gaussian(o, d, st, en, r)
    # nerfstudio compute_3d_gaussian with full matrices
    dd = d[:, None, :].expand(4, 5, 3)
    outer = dd[..., :, None] * dd[..., None, :]
    null = torch.eye(3) - dd[..., :, None] * (dd / (dd ** 2).sum(-1, keepdim=True))[..., None, :]
    cov = dv[..., None, None] * outer + rv[..., None, None] * null
    assert torch.allclose(torch.diagonal(cov, dim1=-2, dim2=-1), diag, atol=1e-7)
    # brute force: uniform points in one conical frustum (radius r*t at depth t)
    n = 400000
    t0, t1 = float(st[0, 0]), float(en[0, 0])
    u = torch.rand(n, generator=g, dtype=torch.float64)
    t = (t0 ** 3 + u * (t1 ** 3 - t0 ** 3)) ** (1 / 3)          # density ~ t^2
    rho = torch.sqrt(torch.rand(n, generator=g, dtype=torch.float64)) * r * t
    assert abs(float(t.mean()) - float(tm[0, 0])) < 2e-4
    assert abs(float(t.var()) - float(dv[0, 0])) < 2e-4 * float(dv[0, 0]) + 2e-6
    assert abs(float((rho ** 2).mean() / 2) - float(rv[0, 0])) < 1e-2 * float(rv[0, 0])


def test_encoding_layout_and_damping():
    x = torch.tensor([[0.1, -0.2, 0.3]])
    e = mo.nerf_encoding(x, 4, 4.0)
    assert e.shape == (1, 27) and torch.equal(e[:, 24:], x)
    f = 2 ** torch.linspace(0, 4, 4)
    assert torch.allclose(e[0, 0:4], torch.sin(2 * math.pi * x[0, 0] * f), atol=1e-6)          # dim-major, freq-minor
    assert torch.allclose(e[0, 12:16], torch.cos(2 * math.pi * x[0, 0] * f), atol=1e-5)        # sin(. + pi/2) block
    e0 = mo.nerf_encoding(x, 24, 24.0, torch.zeros(1, 3))
    assert torch.allclose(e0, mo.nerf_encoding(x, 24, 24.0), atol=0)                           # zero covariance = plain
    ed = mo.nerf_encoding(x, 24, 24.0, torch.full((1, 3), 1e-3))
    assert ed.shape == (1, 147) and float(ed[0, 23].abs()) == 0.0                               # top frequency fully damped


def test_regulariser_quirks_of_the_singleton_dimension():
    g = torch.Generator().manual_seed(2)
    R, V, S = 6, 3, 5
    sd = torch.rand(R, V, S, 1, generator=g)
    ss = torch.rand(R, S, 1, generator=g)
    tot = ss + sd.sum(1)
    n = sd / tot.clamp(min=so.EPS)[:, None]
    assert torch.allclose(so.ray_reg(sd, tot), (n[..., 0] ** 2).mean(0).sum() / V, atol=1e-7)
    a = torch.rand(R, S, 1, generator=g)
    assert float(so.static_reg(torch.rand(R, S, 1, generator=g), a)) == 0.0
    # entropy / dvs / dynamic_reg are shape-agnostic: same value with and without the trailing dimension
    ad = torch.rand(R, V, S, 1, generator=g)
    assert torch.allclose(so.alpha_entropy(a, ad), so.alpha_entropy(a[..., 0], ad[..., 0]), atol=1e-7)
    assert torch.allclose(so.dynamic_vs_static_reg(a, ad), so.dynamic_vs_static_reg(a[..., 0], ad[..., 0]), atol=1e-7)


def test_density_compositing_equals_alpha_compositing():
    """weights/alphas/transmittance of rendering_starmip.py:32-63 == the vanilla exclusive cumprod of (1 - alpha)."""
    g = torch.Generator().manual_seed(3)
    dens = torch.rand(7, 9, 1, generator=g) * 30
    delt = torch.rand(7, 9, 1, generator=g) * 0.1
    w, a, T = mo.weights_alphas_transmittance(delt, dens)
    T2 = torch.cumprod(torch.cat([torch.ones(7, 1), 1 - a[..., 0]], -1), -1)[:, :-1]
    assert torch.allclose(T[..., 0], T2, atol=1e-6) and torch.allclose(w[..., 0], a[..., 0] * T2, atol=1e-6)


def test_median_depth():
    w = torch.tensor([[[0.1], [0.3], [0.2], [0.1]], [[0.0], [0.0], [0.1], [0.1]]])
    st = torch.tensor([[0.0, 1.0, 2.0, 3.0]]).expand(2, 4)
    d = mo.median_depth(w, st, st + 1.0)
    assert d.shape == (2, 1) and float(d[0]) == 2.5 and float(d[1]) == 3.5   # ray 1 never reaches 0.5 -> last sample


def test_pdf_sample_properties():
    g = torch.Generator().manual_seed(4)
    R, Nc, Ni = 5, 16, 24
    sp = mo.uniform_bins(R, Nc)
    w = torch.rand(R, Nc, generator=g)
    for training in (False, True):
        u = torch.rand(R, Ni + 1, generator=g) if training else None
        nb, det = mo.pdf_sample(sp, w, Ni, training, u, return_details=True)
        assert nb.shape == (R, Ni + 1) and (nb[:, 1:] >= nb[:, :-1]).all() and nb.min() >= 0 and nb.max() <= 1
        assert det["inds"].min() >= 1 and det["inds"].max() <= Nc + 1
        nb2 = mo.pdf_sample(sp, w, Ni, training, u, exact_sum=True)
        assert torch.allclose(nb, nb2, atol=1e-5)
    flat = mo.pdf_sample(sp, torch.zeros(R, Nc), Ni)        # constant pdf -> evenly spread edges
    want = (torch.linspace(0, 1 - 1 / (Ni + 1), Ni + 1) + 1 / (2 * (Ni + 1))).expand(R, Ni + 1)
    assert torch.allclose(flat, want, atol=1e-5)


def test_online_with_empty_objects_equals_appinit_and_chunks_sum():
    cfg = mo.MipConfig(num_vehicles=2, N_samples=8, N_importance=8, chunk=4)
    sd = mo.init_mip_params(2, seed=5, gain=2.0)
    for v in range(2):   # objects with (numerically) zero density everywhere
        sd[f"dynamic_nerfs.{v}.field.field_output_density.net.weight"].zero_()
        sd[f"dynamic_nerfs.{v}.field.field_output_density.net.bias"].fill_(-200.0)
    ro, rd = so.carla_rays(9, seed=6)
    vd = rd / rd.norm(dim=-1, keepdim=True)
    a = mo.star_mip_forward(sd, cfg, ro, vd)
    b = mo.star_mip_forward(sd, cfg, ro, vd, pose=so.random_poses7(2))
    for k in ("rgb", "acc", "weights", "depth", "rgb0", "weights0"):
        assert torch.allclose(a[k], b[k], atol=1e-6), k
    assert torch.allclose(b["rgb_static"], b["rgb"], atol=1e-6) and float(b["rgb_dynamic"].abs().max()) < 1e-12
    assert b["dynamic_transmittance"].shape == (9, 2, 1) and b["depth_dynamic"].shape == (9, 2)
    assert float(b["loss_static_reg"]) == 0.0
    # scalar outputs are per-chunk means summed over chunks (star_mipnerf.py:130-133)
    cfg1 = mo.MipConfig(num_vehicles=2, N_samples=8, N_importance=8, chunk=1 << 20)
    parts = [mo.star_mip_forward(sd, cfg1, ro[i:i + 4], vd[i:i + 4], pose=so.random_poses7(2)) for i in (0, 4, 8)]
    assert torch.allclose(b["loss_dynamic_reg"], sum(p["loss_dynamic_reg"] for p in parts), atol=1e-7)


def test_pose_gradient_reaches_the_7_vector_with_pypose_convention():
    cfg = mo.MipConfig(num_vehicles=1, N_samples=6, N_importance=6)
    sd = mo.init_mip_params(1, seed=7, gain=2.0)
    ro, rd = so.carla_rays(5, seed=8)
    vd = rd / rd.norm(dim=-1, keepdim=True)
    pose = so.random_poses7(1).requires_grad_(True)
    out = mo.star_mip_forward(sd, cfg, ro, vd, pose=pose)
    (out["rgb"].sum() + out["rgb0"].sum()).backward()
    assert pose.grad.shape == (1, 7) and float(pose.grad[0, 6]) == 0.0 and float(pose.grad[0, :6].abs().max()) > 0

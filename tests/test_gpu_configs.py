"""GPU tests at the sizes of BASELINE.json's configs.

C1 (configs[0], the reference's own CPU-runnable case): one 100x100 lego view, coarse+fine 64+128, eval -- the whole
view against the CPU oracle (fp32 tier: coarse keys tight, fine keys through the reference's own fine samples).
C2 (configs[1]) at full size (800x800, bf16 MLP): size-independent properties -- finiteness and ranges, weights summing
to acc, invariance to how the rays are split over launches, agreement of a random subset with the fp32 tier, and
sortedness of the merged samples."""
import pytest
import torch

import star_b200
from star_b200 import functional as F_
from star_b200.models import rendering__ as R_
from oracle import ref_harness, star_oracle as so
from helpers import assert_close, psnr_db

pytestmark = pytest.mark.gpu
DEV = "cuda"
NC, NI, NEAR, FAR = 64, 128, 2.0, 6.0


def lego_net(precision):
    net = star_b200.STaR(ref_harness.make_args(num_vehicles=0, N_importance=NI, chunk=8192, white_bkgd=True))
    sd = so.init_star_params(0, NI, seed=0, bias_std=0.02)
    net.load_state_dict(sd)
    net.to(DEV).eval()
    net.set_precision(precision)
    return net, sd


def test_c1_lego_100x100_view_against_the_cpu_oracle():
    net, sd = lego_net("fp32")
    ro, rd = so.lego_rays(100, 100)
    ro, rd = ro.reshape(-1, 3).contiguous(), rd.reshape(-1, 3).contiguous()
    vd = rd / rd.norm(dim=-1, keepdim=True)
    cfg = so.StarConfig(0, NI, 8192, white_bkgd=True)
    with torch.no_grad():
        pts, z = so.sample_pts(ro, rd, NEAR, FAR, NC, is_train=False)
        ref = so.render_star(sd, cfg, pts, vd, z, ro, rd, NI, training=False, exact_sum=True)
        mid = 0.5 * (z[..., 1:] + z[..., :-1])
        zs_ref = so.sample_pdf(mid, ref["weights0"][..., 1:-1], NI, det=True, exact_sum=True)
        c = lambda t: t.to(DEV)
        pts_g, z_g = R_.sample_pts(c(ro), c(rd), NEAR, FAR, NC, is_train=False)
        assert torch.equal(z_g.cpu(), z) and torch.equal(pts_g.cpu(), pts)
        out = R_.render_star_appinit(net, pts_g, c(vd), z_g, c(ro), c(rd), NI)
        forced = R_.render_star_appinit(net, pts_g, c(vd), z_g, c(ro), c(rd), NI, z_samples=c(zs_ref))
    assert out["rgb"].shape == (10000, 3) and set(ref.keys()) <= set(out.keys())
    for k in ("rgb0", "depth0", "acc0", "weights0", "disp0"):
        assert_close(out[k], ref[k], 1e-4, rtol=1e-4 if k == "disp0" else 0.0, msg=k)
    for k in ("rgb", "depth", "acc", "weights"):      # same fine samples as the reference -> the north star's 1e-4
        assert_close(forced[k], ref[k], 1e-4, msg=k + " (teacher forced)")
    # free running: the fine samples come from this run's own coarse weights (sample_pdf conditioning)
    assert float((out["rgb"].cpu() - ref["rgb"]).abs().max()) < 5e-3
    assert abs(psnr_db(out["rgb"].cpu(), ref["rgb"]) ) > 40.0


def test_c2_full_size_render_fp16_tier_properties_and_oracle_parity():
    """C2 = lego 800x800, 64+128, the 16-bit tensor-core tier that the bench reports (fp16 operands).  Whole view:
    size-independent properties.  4096 random rays of the view: against the fp32 CPU ORACLE, teacher forced with the
    oracle's own fine samples -- the north star's bound for the 16-bit MLP: 2e-3 absolute on rgb / weights / depth
    and <= 0.05 dB PSNR."""
    net, sd = lego_net("fp16")
    ro_h, rd_h = so.lego_rays(800, 800)
    ro_h, rd_h = ro_h.reshape(-1, 3).contiguous(), rd_h.reshape(-1, 3).contiguous()
    ro, rd = ro_h.to(DEV), rd_h.to(DEV)
    vd = rd / rd.norm(dim=-1, keepdim=True)
    R = ro.shape[0]
    assert R == 640000
    with torch.no_grad():
        pts, z = R_.sample_pts(ro, rd, NEAR, FAR, NC, is_train=False)
        out = R_.render_star_appinit(net, pts, vd, z, ro, rd, NI)
        F_.check_range()                      # no fp16 operand overflowed anywhere in the view
        for k in ("rgb", "rgb0", "depth", "acc", "weights", "weights0", "z_std"):
            assert bool(torch.isfinite(out[k]).all()), k
        assert out["rgb"].shape == (R, 3) and out["weights"].shape == (R, NC + NI) and out["weights0"].shape == (R, NC)
        assert float(out["rgb"].min()) >= -1e-5 and float(out["rgb"].max()) <= 1.0 + 1e-4      # white background blend
        assert float(out["weights"].min()) >= 0.0
        assert_close(out["weights"].sum(-1), out["acc"], 2e-5, msg="sum(weights) == acc")
        assert float(out["acc"].max()) <= 1.0 + 1e-5
        # merged sample positions are sorted and stay inside [near, far]
        zf = out["z_vals"]
        assert bool((zf[:, 1:] >= zf[:, :-1]).all()) and float(zf.min()) >= NEAR - 1e-6 and float(zf.max()) <= FAR + 1e-6
        # splitting the view over launches differently does not change a bit
        idx = torch.randperm(R, device=DEV, generator=torch.Generator(device=DEV).manual_seed(0))[:4096]
        sub = R_.render_star_appinit(net, pts[idx], vd[idx], z[idx], ro[idx], rd[idx], NI)
        for k in ("rgb", "depth", "weights", "rgb0"):
            assert torch.equal(sub[k], out[k][idx]), k
        # ---- the oracle (fp32, CPU) on the same 4096 rays
        ih = idx.cpu()
        cfg = so.StarConfig(0, NI, 8192, white_bkgd=True)
        vd_h = rd_h[ih] / rd_h[ih].norm(dim=-1, keepdim=True)
        pts_h, z_h = so.sample_pts(ro_h[ih], rd_h[ih], NEAR, FAR, NC, is_train=False)
        ref = so.render_star(sd, cfg, pts_h, vd_h, z_h, ro_h[ih], rd_h[ih], NI, training=False, exact_sum=True)
        mid = 0.5 * (z_h[..., 1:] + z_h[..., :-1])
        zs_ref = so.sample_pdf(mid, ref["weights0"][..., 1:-1], NI, det=True, exact_sum=True)
        forced = R_.render_star_appinit(net, pts[idx], vd[idx], z[idx], ro[idx], rd[idx], NI, z_samples=zs_ref.to(DEV))
    for k in ("rgb0", "rgb", "weights0", "weights", "depth0", "depth", "acc"):
        e = (forced[k].cpu() - ref[k]).abs()
        assert float(e.max()) <= 2e-3, (k, float(e.max()), float(e.mean()))
    target = torch.rand(4096, 3, generator=torch.Generator().manual_seed(1))
    for k in ("rgb0", "rgb"):
        assert abs(psnr_db(forced[k].cpu(), target) - psnr_db(ref[k], target)) <= 0.05, k
    assert psnr_db(forced["rgb"].cpu(), ref["rgb"]) > 60.0

"""Per-layer error budget of the 16-bit tensor-core MLP tiers ("next round" item 1).

For every GEMM layer group of the NeRF MLP the oracle's operand-rounding model is switched on for THAT group only
(all other layers exact fp32) and the composited outputs are compared with the all-fp32 oracle on the end-to-end
fixtures and on a slice of the C2 workload; then the candidate tier mixes are evaluated the same way.  CPU only.

    python tools/error_budget.py [--c2-rays 512]
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import star_oracle as so  # noqa: E402
from helpers import load_golden, psnr_db  # noqa: E402

GROUPS = ["lin_in", "fc_0", "fc_1", "lin_out", "feature_linear", "views_linears"]


def cases(c2_rays):
    for name, V in (("e2e_appinit_eval", 0), ("e2e_online_mat_eval", 2)):
        g = load_golden(name)
        p = so.init_star_params(V, 24, seed=int(g["seed"]), bias_std=0.02)
        vd = g["rays_d"] / g["rays_d"].norm(dim=-1, keepdim=True)
        pts, z = so.sample_pts(g["rays_o"], g["rays_d"], g["near"], g["far"], int(g["Nc"]), is_train=False)

        def run(mode, p=p, g=g, V=V, vd=vd, pts=pts, z=z):
            cfg = so.StarConfig(V, 24, 4096, white_bkgd=(V == 0), emulate_bf16=mode)
            with torch.no_grad():
                return so.render_star(p, cfg, pts, vd, z, g["rays_o"], g["rays_d"], int(g["Ni"]),
                                      pose=g["pose"] if V else None, training=False, z_samples=g["z_samples"])
        yield name, run
    if c2_rays:
        H = W = 800
        ro, rd = so.lego_rays(H, W)
        sel = torch.randperm(H * W, generator=torch.Generator().manual_seed(5))[:c2_rays]
        ro, rd = ro.reshape(-1, 3)[sel], rd.reshape(-1, 3)[sel]
        vd = rd / rd.norm(dim=-1, keepdim=True)
        p = so.init_star_params(0, 128, seed=0)
        pts, z = so.sample_pts(ro, rd, 2.0, 6.0, 64, is_train=False)
        cfg0 = so.StarConfig(0, 128, 8192, white_bkgd=True)
        with torch.no_grad():
            ref = so.render_star(p, cfg0, pts, vd, z, ro, rd, 128, training=False, exact_sum=True)
        zs = None

        def run(mode):
            cfg = so.StarConfig(0, 128, 8192, white_bkgd=True, emulate_bf16=mode)
            with torch.no_grad():   # teacher forced: same fine samples for every tier
                return so.render_star(p, cfg, pts, vd, z, ro, rd, 128, training=False, z_samples=run.zs)
        # the fine samples of the exact coarse pass
        mid = 0.5 * (z[..., 1:] + z[..., :-1])
        run.zs = so.sample_pdf(mid, ref["weights0"][..., 1:-1], 128, det=True, exact_sum=True)
        yield f"c2_lego_{c2_rays}rays", run


def stats(out, ref):
    res = {}
    for k in ("rgb0", "rgb", "weights", "depth"):
        e = (out[k].double() - ref[k].double()).abs()
        res[k] = float(e.max())
    tgt = torch.rand(out["rgb"].shape, generator=torch.Generator().manual_seed(1))
    res["dpsnr"] = abs(psnr_db(out["rgb"], tgt) - psnr_db(ref["rgb"], tgt))
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--c2-rays", type=int, default=256)
    a = ap.parse_args()
    torch.set_num_threads(os.cpu_count())
    mixes = {"all bf16": "bf16", "all fp16": "fp16"}
    for grp in GROUPS:
        mixes[f"only {grp} bf16"] = {grp: "bf16"}
    for grp in GROUPS:
        mixes[f"only {grp} fp16"] = {grp: "fp16"}
    mixes["fp16, lin_in split"] = {g: "fp16" for g in GROUPS} | {"lin_in": "fp16x2"}
    mixes["bf16, lin_in split"] = {g: "bf16" for g in GROUPS} | {"lin_in": "bf16x2"}
    mixes["all bf16x2"] = {g: "bf16x2" for g in GROUPS}
    for name, run in cases(a.c2_rays):
        ref = run(False)
        print(f"\n== {name}: max |x - fp32| (rgb0, rgb, weights, depth), |dPSNR| dB")
        for label, mode in mixes.items():
            s = stats(run(mode), ref)
            print(f"  {label:28s} rgb0 {s['rgb0']:.2e}  rgb {s['rgb']:.2e}  w {s['weights']:.2e}  depth {s['depth']:.2e}"
                  f"  dPSNR {s['dpsnr']:.4f}")


if __name__ == "__main__":
    main()

"""Small-shape pass over every kernel family of the render path, for `compute-sanitizer --tool memcheck|racecheck`
(the tensor-core kernels keep 20+ mbarriers with hand-rolled parities).  Shapes are tiny because the
sanitizer slows kernels 10-100x:  fp32 + fp16 + bf16 tiers, static and object nets, forward / stash forward / dX / dW /
heads, single-call render (V = 0 and V = 2, also from a camera), the CTA-pair forward, compositing + hierarchical +
losses + optimiser step, and the mip field.

    compute-sanitizer --tool memcheck python tools/sanitize_smoke.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import star_b200  # noqa: E402
from star_b200 import functional as F_, optim as O_  # noqa: E402
from star_b200.models import rendering__ as R_, loss as L_  # noqa: E402
from oracle import ref_harness, star_oracle as so  # noqa: E402  (input builders only)

dev = "cuda"
V, Nc, Ni, R = 2, 16, 24, 70
net = star_b200.STaR(ref_harness.make_args(num_vehicles=V, N_importance=Ni, chunk=64))
net.load_state_dict(so.init_star_params(V, Ni, seed=11, bias_std=0.02))
net.to(dev)
ro, rd = so.carla_rays(R, seed=1)
ro, rd = ro.to(dev), rd.to(dev)
vd = rd / rd.norm(dim=-1, keepdim=True)
pose = torch.nn.Parameter(so.random_poses7(V, seed=3).to(dev))
u = torch.rand(R, Ni, device=dev)
target = torch.rand(R, 3, device=dev)
for prec in ("fp32", "fp16", "bf16"):
    for pair in (False,):
        net.set_precision(prec)
        net.train()
        net.zero_grad()
        pose.grad = None
        pts, z = R_.sample_pts(ro, rd, 0.03, 0.8, Nc)
        out = R_.render_star_online(net, pts, vd, z, ro, rd, Ni, pose, u=u)
        loss = L_.photometric_loss(out["rgb0"], out["rgb"], target)[0] + 1e-3 * out["loss_alpha_entropy"]
        loss.backward()
        net.eval()
        with torch.no_grad():
            o1 = R_.render_star_online(net, pts, vd, z, ro, rd, Ni, pose)          # single-call entry, V = 2
            o2 = R_.render_star_appinit(net, pts, vd, z, ro, rd, Ni)               # single-call entry, V = 0
        torch.cuda.synchronize()
        print(prec, "pair" if pair else "single", float(loss), float(o1["rgb"].mean()), float(o2["rgb"].mean()), flush=True)
K = torch.tensor([[20.0, 0, 8.0], [0, 20.0, 6.0], [0, 0, 1.0]])
c2w = torch.eye(4, device=dev)[:3]
with torch.no_grad():
    res = F_.render_forward((net.static_coarse_nerf, net.static_fine_nerf), ([], []), net.static_coarse_nerf._prec(), None,
                            None, None, Ni, near=0.03, far=0.8, N_samples=Nc, camera=(12, 16, K, c2w, (0, 12)))
opt = O_.FusedAdam(list(net.parameters()) + [pose], lr=5e-4, max_grad_norm=1.0)
opt.step()
torch.cuda.synchronize()
# mip field
import argparse  # noqa: E402
from oracle import mip_oracle as mo  # noqa: E402
from star_b200.models.star_mipnerf import STaR as MipSTaR  # noqa: E402
margs = argparse.Namespace(num_vehicles=1, chunk=1 << 20, far_dist=1e10, N_importance=24, N_samples=16, scale_factor=0.01,
                           near=3.0, far=80.0)
mnet = MipSTaR(margs)
mnet.load_state_dict(mo.init_mip_params(1, seed=5, gain=1.4, bias_std=0.02))
mnet.to(dev).train()
mp = torch.nn.Parameter(so.random_poses7(1, seed=4).to(dev))
mo_ = mnet(ro[:40], vd[:40], mp)
(mo_["rgb"].mean() + mo_["rgb0"].mean()).backward()
mnet.eval()
mnet.set_precision("fp16")
with torch.no_grad():
    mnet(ro[:40], vd[:40], mp)
torch.cuda.synchronize()
F_.check_range()
print("sanitize_smoke: done", float(res["rgb"].mean()))

"""Uninitialised-memory hunt: run the C2 training step (4096 rays, 64 + 128 samples, fp16 tier) after filling the caching
allocator's free blocks with different byte patterns.  A kernel that reads memory it (or an earlier kernel) never wrote shows
up as gradients that depend on the pattern, as non-finite values, or as a CUDA fault."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import star_b200
from star_b200.models import rendering__ as R_, loss as L_
from oracle import ref_harness, star_oracle as so

dev = torch.device("cuda")
prec = sys.argv[1] if len(sys.argv) > 1 else "fp16"
net = star_b200.STaR(ref_harness.make_args(num_vehicles=0, N_importance=128, chunk=1 << 20))
net.load_state_dict(so.init_star_params(0, 8, seed=3, bias_std=0.02), strict=False)
net.to(dev).set_precision(prec)
net.train()
R, NC, NI = 4096, 64, 128
ro, rd = so.carla_rays(R, seed=1)
ro, rd = ro.to(dev), rd.to(dev)
vd = rd / rd.norm(dim=-1, keepdim=True)
g = torch.Generator().manual_seed(5)
u = torch.rand(R, NI, generator=g).to(dev)
target = torch.rand(R, 3, generator=g).to(dev)


def poison(byte, gb=24):
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    xs = [torch.full((1 << 30,), byte, dtype=torch.uint8, device=dev) for _ in range(gb)]
    torch.cuda.synchronize()
    del xs          # back to the allocator's cache, contents intact


def step():
    net.zero_grad(set_to_none=True)
    pts, z = R_.sample_pts(ro, rd, 2.0, 6.0, NC, perturb=0, is_train=True)
    out = R_.render_star_appinit(net, pts, vd, z, ro, rd, NI, u=u)
    loss = L_.photometric_loss(out["rgb0"], out["rgb"], target)[0]
    loss.backward()
    torch.cuda.synchronize()
    return float(loss), [p.grad.clone() for p in net.parameters() if p.grad is not None]


ref = None
for byte in (0x00, 0xFF, 0x7F, 0x7C, 0xFC, 0x55, 0x80):
    poison(byte)
    for rep in range(2):
        loss, grads = step()
        fin = all(bool(torch.isfinite(x).all()) for x in grads)
        if ref is None:
            ref = (loss, grads)
        worst = max(float((a - b).abs().max() / (b.abs().max() + 1e-30)) for a, b in zip(grads, ref[1]))
        print("pattern 0x%02X rep %d: loss %.9g  finite %s  max rel grad diff vs first %.3g" % (byte, rep, loss, fin, worst), flush=True)
star_b200.functional.check_range()
print("done")

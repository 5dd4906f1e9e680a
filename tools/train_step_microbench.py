"""Times the loss / optimiser-step kernels alone (CUDA events, L2 flushed between iterations) against the measured HBM
peak, next to what the reference runs for the same step on the same GPU (torch.optim.Adam + clip_grad_norm_, eager
torch losses).  Algorithmic bytes: Adam 28 B/param (p, g, m, v read; p, m, v written), squared norm 4 B/param,
sigma loss forward 12 B/sample, backward 16 B/sample, photometric 20 B/element (3 reads, 2 gradient writes)."""
import json, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import star_b200
from star_b200 import functional as F_, optim as O_
from star_b200.models import loss as L_
PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.isfile(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
dev = "cuda"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def timeit(fn, n=10, do_flush=True):
    for _ in range(3): fn()
    ms = 0.0
    for i in range(n):
        if do_flush: flush.fill_(i)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ms += e0.elapsed_time(e1)
    return ms / n

def report(name, ms, nbytes):
    gbs = nbytes / ms / 1e6
    print("%-58s %8.3f ms  %8.1f GB/s  %.2f of HBM peak (%.0f)" % (name, ms, gbs, gbs / PEAK, PEAK), flush=True)

g = torch.Generator(device=dev).manual_seed(0)
for label, n in (("C2: 2 static nets", 2 * 711300), ("C4: 2 static + 10 dynamic nets", 2 * 711300 + 10 * 448132),
                 ("64 M parameters (beyond L2)", 64 << 20)):
    p0 = 0.1 * torch.randn(n, device=dev, generator=g)
    gr = torch.randn(n, device=dev, generator=g)
    pa, pb = torch.nn.Parameter(p0.clone()), torch.nn.Parameter(p0.clone())
    pa.grad, pb.grad = gr.clone(), gr.clone()
    oa = torch.optim.Adam([pa], lr=5e-4)
    ob = O_.FusedAdam([pb], lr=5e-4, max_grad_norm=1.0)
    def ref_step():
        torch.nn.utils.clip_grad_norm_([pa], 1.0); oa.step()
    for fl in (True, False):
        tag = "L2 flushed" if fl else "L2 warm"
        ms = timeit(ob.step, do_flush=fl)
        report("FusedAdam + clip, %s (%s)" % (label, tag), ms, n * 32)
        ms = timeit(ref_step, do_flush=fl)
        report("  torch.optim.Adam + clip_grad_norm_ (%s)" % tag, ms, n * 32)
# the same C4 parameter count as ~390 separate tensors (what torch's foreach path sees in the reference)
shapes = []
for nb in [4, 4] + [2] * 10:
    shapes += [(256, 63), (256,)] + [(256, 256), (256,)] * (2 * nb) + [(256, 256), (256,), (1, 256), (1,), (256, 256), (256,),
                                                                     (128, 283), (128,), (3, 128), (3,)]
pa = [torch.nn.Parameter(0.1 * torch.randn(*s, device=dev, generator=g)) for s in shapes]
holder = torch.nn.ParameterList([torch.nn.Parameter(p.detach().clone()) for p in pa])
O_.flatten_parameters(holder)
pb = list(holder)
n = sum(p.numel() for p in pa)
gflat = torch.randn(n, device=dev, generator=g)
off = 0
for a, b in zip(pa, pb):
    a.grad = gflat[off:off + a.numel()].view(a.shape).clone()
    b.grad = gflat[off:off + a.numel()].view(a.shape)
    off += a.numel()
oa = torch.optim.Adam(pa, lr=5e-4)
ob = O_.FusedAdam(pb, lr=5e-4, max_grad_norm=1.0)
def ref_step():
    torch.nn.utils.clip_grad_norm_(pa, 1.0); oa.step()
import time
for name, fn in (("FusedAdam + clip, %d tensors in one flat run" % len(pa), ob.step), ("  torch.optim.Adam + clip_grad_norm_, %d tensors" % len(pa), ref_step)):
    ms = timeit(fn)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(20): fn()
    torch.cuda.synchronize(); wall = (time.perf_counter() - t0) / 20 * 1e3
    report(name, ms, n * 32)
    print("%-58s %8.3f ms wall per step (host + device, back to back)" % ("", wall))

for R, S in ((4096, 192), (65536, 512), (640000, 192)):
    w = torch.rand(R, S, device=dev, generator=g) * 0.01
    z = 0.03 + 0.77 * torch.sort(torch.rand(R, S, device=dev, generator=g), dim=1).values
    dists = torch.rand(R, S, device=dev, generator=g) * 0.01
    depths = 0.03 + 0.77 * torch.rand(R, device=dev, generator=g)
    wl = w.clone().requires_grad_(True)
    with torch.no_grad():
        ms = timeit(lambda: L_.compute_sigma_loss(w, z, dists, depths, 0.03, 0.8))
    report("sigma_loss forward R=%d S=%d" % (R, S), ms, R * S * 12)
    loss = L_.compute_sigma_loss(wl, z, dists, depths, 0.03, 0.8)
    ms = timeit(lambda: torch.autograd.grad(loss, wl, retain_graph=True))
    report("sigma_loss backward R=%d S=%d" % (R, S), ms, R * S * 16)
    if R * S <= 65536 * 512:
        sys.path.insert(0, ROOT)
        from oracle import train_oracle as to
        with torch.no_grad():
            ms = timeit(lambda: to.compute_sigma_loss(w, z, dists, depths, 0.03, 0.8))
        report("  eager torch ops (the reference's formulation) fwd", ms, R * S * 12)
for R in (4096, 640000):
    a, b, t = (torch.rand(R, 3, device=dev, generator=g) for _ in range(3))
    a.requires_grad_(True); b.requires_grad_(True)
    ms = timeit(lambda: L_.photometric_loss(a, b, t))
    report("photometric loss + both gradients R=%d" % R, ms, R * 3 * 20)
    def ref():
        l = torch.mean((a - t) ** 2) + torch.mean((b - t) ** 2)
        torch.autograd.grad(l, (a, b))
    ms = timeit(ref)
    report("  eager torch MSE x2 + autograd R=%d" % R, ms, R * 3 * 20)

"""Times the sampling / compositing kernels alone (CUDA events, inputs resident in HBM, L2 flushed between
iterations) and reports achieved algorithmic GB/s against the measured HBM peak (MEASURED_PEAKS.json)."""
import json, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import star_b200
from star_b200 import functional as F_
from star_b200.models import rendering__ as R_
PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.isfile(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
dev = "cuda"
R = int(os.environ.get("R", 160000))
# The flush also hides the host: the timed call's Python wrapper (allocation + ctypes, 20-40 us) must be enqueued while the
# flush still runs, or the gap lands inside the event pair -- 256 MB (39 us) was too short for that and inflated every kernel
# under ~100 us (r1k / r2j figures of sample_pts and composite_single S=64); 2 GB = 0.3 ms.
flush = torch.empty(int(os.environ.get("FLUSH_MB", 2048)) << 20, dtype=torch.uint8, device=dev)

def timeit(fn, n=10):
    if os.environ.get("PROFILE"):      # under ncu: one call per kernel
        fn(); torch.cuda.synchronize(); return 1.0
    for _ in range(3): fn()
    ms = 0.0
    for i in range(n):
        flush.fill_(i); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ms += e0.elapsed_time(e1)
    return ms / n

def report(name, ms, nbytes):
    gbs = nbytes / ms / 1e6
    print("%-44s %8.3f ms  %8.1f GB/s  %.2f of HBM peak (%.0f)" % (name, ms, gbs, gbs / PEAK, PEAK), flush=True)

g = torch.Generator(device=dev).manual_seed(0)
ro = torch.randn(R, 3, device=dev, generator=g); rd = torch.nn.functional.normalize(torch.randn(R, 3, device=dev, generator=g), dim=-1) * 1.1
with torch.no_grad():
    for Nc, Ni in ((64, 128), (256, 256)):
        Rr = R if Nc == 64 else R // 4
        o, d = ro[:Rr].contiguous(), rd[:Rr].contiguous()
        ms = timeit(lambda: F_.sample_pts(o, d, 2.0, 6.0, Nc))
        report("sample_pts R=%d Nc=%d" % (Rr, Nc), ms, Rr * (24 + 16 * Nc))
        pts, z = F_.sample_pts(o, d, 2.0, 6.0, Nc)
        for S, zz in ((Nc, z),):
            ra = torch.randn(Rr, S, device=dev, generator=g); rc = torch.randn(Rr, S, 3, device=dev, generator=g)
            ms = timeit(lambda: F_.CompositeSingle.apply(ra, rc, zz, d, 1e10, True))
            report("composite_single_fwd R=%d S=%d" % (Rr, S), ms, Rr * S * 28)
            w = F_.CompositeSingle.apply(ra, rc, zz, d, 1e10, True)[4]
        ms = timeit(lambda: F_.hierarchical(z, w, Ni, True, o, d))
        report("hierarchical (det) R=%d %d+%d" % (Rr, Nc, Ni), ms, Rr * (8 * Nc + 4 * Ni + 16 * (Nc + Ni)))
        u = torch.rand(Rr, Ni, device=dev, generator=g)
        ms = timeit(lambda: F_.hierarchical(z, w, Ni, False, o, d, u=u))
        report("hierarchical (random u) R=%d %d+%d" % (Rr, Nc, Ni), ms, Rr * (8 * Nc + 8 * Ni + 16 * (Nc + Ni)))
        # coarse-pass tail of the single-call render (no positions): the two stand-alone kernels against the fused one
        ra0 = torch.randn(Rr, Nc, device=dev, generator=g); rc0 = torch.randn(Rr, Nc, 3, device=dev, generator=g)
        def pair():
            ww = F_.CompositeSingle.apply(ra0, rc0, z, d, 1e10, True)[4]
            F_.hierarchical(z, ww, Ni, True, o, d, want_pts=False)
        ms = timeit(pair)
        report("composite + hierarchical, 2 kernels R=%d %d+%d" % (Rr, Nc, Ni), ms, Rr * (28 * Nc + 8 * Nc + 4 * Ni + 4 * (Nc + Ni)))
        ms = timeit(lambda: F_.composite_hier(ra0, rc0, z, d, 1e10, True, Ni, True))
        report("composite_hier fused R=%d %d+%d" % (Rr, Nc, Ni), ms, Rr * (28 * Nc + 4 * Ni + 4 * (Nc + Ni)))
        ms = timeit(lambda: F_.composite_hier(ra0, rc0, z, d, 1e10, True, Ni, True, want_weights=False))
        report("composite_hier fused, no weights/dists out" , ms, Rr * (20 * Nc + 4 * Ni + 4 * (Nc + Ni)))
        zs, zall, zstd, ptsf = F_.hierarchical(z, w, Ni, True, o, d)
        S = Nc + Ni
        ra = torch.randn(Rr, S, device=dev, generator=g); rc = torch.randn(Rr, S, 3, device=dev, generator=g)
        ms = timeit(lambda: F_.CompositeSingle.apply(ra, rc, zall, d, 1e10, True))
        report("composite_single_fwd R=%d S=%d" % (Rr, S), ms, Rr * S * 28)
        V = 5
        rad = torch.randn(Rr, V, S, device=dev, generator=g); rcd = torch.randn(Rr, V, S, 3, device=dev, generator=g)
        ms = timeit(lambda: F_.CompositeStar.apply(ra, rc, rad, rcd, zall, d, 1e10, False, 8192, False))
        report("composite_multi_fwd V=5 R=%d S=%d" % (Rr, S), ms, Rr * S * ((1 + V) * 16 + 4 + 4))
        with torch.enable_grad():
            leaves = [t.clone().requires_grad_(True) for t in (ra, rc, rad, rcd)]
            ws = F_.CompositeStar.apply(*leaves, zall, d, 1e10, False, 8192, False)
            loss = ws[0].sum() + ws[4].sum() + ws[11].sum()      # rgb, weights, the five regularisers
            def bwd():
                for t in leaves: t.grad = None
                loss.backward(retain_graph=True)
            ms = timeit(bwd)
        report("composite_multi_bwd (+ autograd glue) V=5 R=%d S=%d" % (Rr, S), ms, Rr * S * ((1 + V) * 32 + 4 + 4))

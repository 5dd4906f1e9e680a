#!/usr/bin/env python
"""Benchmark of the STaR / NeRF render hot path (BASELINE.json metric: rays/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--mode render|train] [--precision fp16|bf16|fp32]
    python bench.py --impl reference ...      # the reference itself (oracle/_ref) on the host cores; oracle port if absent

Workload (BASELINE.json configs[1], "C2"): lego-shaped vanilla NeRF, coarse+fine 64+128 samples,
random-init weights (fc_1 re-drawn), perturb=0, white background.
  render : one "step" = one full 800x800 view (640 000 rays) through sample_pts + render_star_appinit (eval).
  train  : one "step" = one 4096-ray training step (forward + backward to all MLP weights).
`value` is device-timed with the rays already resident in HBM; `e2e` goes through the same public
API from pinned HOST buffers (H2D of the rays and D2H of rgb/depth/acc inside the timed region).
With N>1 (torchrun, one process per GPU) every rank renders its own view / batch (rays shard with no
data-path collective; "weak" scaling); train mode all-reduces the flat gradient with NCCL, overlapped with the
backward of the other net.  The default line (C2 render) also carries sub-records: "train" (C2 4096-ray step, with
and without the optimiser), "c4_train" / "c4_render" (static + 5 objects), "c5_render" (mip field), "strong" (N > 1: one
view and one batch SPLIT over the ranks), "accuracy" (max error of the benched tier against the fp32 oracle) and, at N = 1,
"gpu_eager_baseline" (the reference's eager PyTorch code on the same GPU).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

NC, NI = 64, 128
NEAR, FAR = 2.0, 6.0
F_STATIC = 1.416704e6      # MLP FLOP per sample (2*MAC), static net
RENDER_HW = 800
TRAIN_RAYS = 4096


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mode", default="render", choices=["render", "train"])
    ap.add_argument("--workload", default="c2", choices=["c2", "c3", "c4", "c5"],
                    help="c2: lego vanilla NeRF (the headline, BASELINE.json configs[1]); c4: carla_star_online_multi "
                         "(static + 5 objects, 256+256 samples, 7-vector poses); c5: carla_star_app_init_mip (mip-NeRF "
                         "fields, 256+512 frustums).  c4 / c5 are extra measurements; the driver runs c2.")
    ap.add_argument("--rays", type=int, default=0, help="c4 / c5: rays per step per GPU (0 = default of the workload)")
    ap.add_argument("--precision", default=os.environ.get("STAR_B200_PRECISION", "fp16"), choices=["bf16", "fp16", "fp32"],
                    help="MLP tier: fp16 = tcgen05, fp16 operands / fp32 accumulate (the tier that meets the 2e-3 bound; "
                         "default), bf16 = same kernels with bf16 operands, fp32 = CUDA-core 1e-4 tier")
    ap.add_argument("--hw", type=int, default=RENDER_HW, help="render mode: view is hw x hw rays")
    ap.add_argument("--train-rays", type=int, default=TRAIN_RAYS)
    ap.add_argument("--cpu-sample-rays", type=int, default=0, help="rays of the bounded CPU sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--opt-step", action="store_true",
                    help="train mode: include gradient clipping + the fused Adam step (and the weight re-pack) in the step")
    ap.add_argument("--no-train-extra", action="store_true",
                    help="render mode: skip the extra 4096-ray training-step measurement reported under \"train\"")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the c4 / c5 / strong-scaling / accuracy / eager-GPU sub-records of the default line")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------ inputs
def lego_view(hw, theta=30.0):
    from oracle import star_oracle as so
    ro, rd = so.lego_rays(hw, hw, theta=theta)
    return ro.reshape(-1, 3).contiguous(), rd.reshape(-1, 3).contiguous()


def make_params(seed=0):
    from oracle import star_oracle as so
    return so.init_star_params(0, NI, seed=seed, bias_std=0.02)


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.rows = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ reference arm
def _reference_modules():
    """The reference's own code (oracle/_ref: its hot-path files, unmodified) when present, else None (-> oracle port)."""
    from oracle import ref_harness
    if not ref_harness.reference_available():
        return None
    return ref_harness.load_reference()


def _c2_sample(n_rays, seed=0):
    ro, rd = lego_view(RENDER_HW)
    g = torch.Generator().manual_seed(seed)
    idx = torch.randperm(ro.shape[0], generator=g)[:n_rays]
    ro, rd = ro[idx].contiguous(), rd[idx].contiguous()
    return ro, rd, rd / rd.norm(dim=-1, keepdim=True), g


def reference_render_fn(device="cpu"):
    """-> (kind, fn): fn(ro, rd, vd) renders C2 rays (coarse+fine 64+128, eval) with the reference's eager PyTorch code
    (kind "reference": models/rendering__.py::render_star_appinit on models/star__.py::STaR, unmodified) or, when
    oracle/_ref is absent, with the oracle port (same ATen ops)."""
    from oracle import ref_harness, star_oracle as so
    ref = _reference_modules()
    if ref is not None:
        net = ref.star.STaR(ref_harness.make_args(num_vehicles=0, N_importance=NI, chunk=8192, white_bkgd=True))
        net.load_state_dict(make_params())
        net.to(device).eval()

        def fn(ro, rd, vd):
            pts, z = ref.rendering.sample_pts(ro, rd, NEAR, FAR, NC, perturb=0, is_train=False)
            return ref.rendering.render_star_appinit(net, pts, vd, z, ro, rd, NI)
        return "reference", fn
    p = {k: v.to(device) for k, v in make_params().items()}
    cfg = so.StarConfig(0, NI, 8192, white_bkgd=True)

    def fn(ro, rd, vd):
        pts, z = so.sample_pts(ro, rd, NEAR, FAR, NC, is_train=False)
        return so.render_star(p, cfg, pts, vd, z, ro, rd, NI, training=False)
    return "port", fn


def cpu_render_rays_per_s(n_rays, threads, repeat=1):
    """The reference on the host cores: render `n_rays` rays of the C2 view, coarse+fine, eval mode."""
    torch.set_num_threads(threads)
    kind, fn = reference_render_fn("cpu")
    ro, rd, vd, _ = _c2_sample(n_rays)
    best = None
    with torch.no_grad():
        for _ in range(repeat):
            t0 = time.perf_counter()
            out = fn(ro, rd, vd)
            float(out["rgb"].sum())
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
    return n_rays / best, best, kind


def cpu_train_rays_per_s(n_rays, threads):
    from oracle import ref_harness, star_oracle as so
    torch.set_num_threads(threads)
    ro, rd, vd, g = _c2_sample(n_rays)
    target = torch.rand(n_rays, 3, generator=g)
    ref = _reference_modules()
    t0 = time.perf_counter()
    if ref is not None:      # the reference draws sample_pdf's u itself in train mode (rendering__.py:741)
        kind = "reference"
        net = ref.star.STaR(ref_harness.make_args(num_vehicles=0, N_importance=NI, chunk=8192, white_bkgd=True))
        net.load_state_dict(make_params())
        net.train()
        t0 = time.perf_counter()
        pts, z = ref.rendering.sample_pts(ro, rd, NEAR, FAR, NC, perturb=0, is_train=True)
        out = ref.rendering.render_star_appinit(net, pts, vd, z, ro, rd, NI)
    else:
        kind = "port"
        p = {k: v.requires_grad_(True) for k, v in make_params().items()}
        cfg = so.StarConfig(0, NI, 8192, white_bkgd=True)
        u = torch.rand(n_rays, NI, generator=g)
        t0 = time.perf_counter()
        pts, z = so.sample_pts(ro, rd, NEAR, FAR, NC)
        out = so.render_star(p, cfg, pts, vd, z, ro, rd, NI, training=True, u=u)
    loss = ((out["rgb"] - target) ** 2).mean() + ((out["rgb0"] - target) ** 2).mean()
    loss.backward()
    dt = time.perf_counter() - t0
    return n_rays / dt, dt, kind


def gpu_eager_baseline(dev, n_rays=65536):
    """The reference's eager PyTorch path ON THE SAME GPU ("the real bar"): C2 render of `n_rays` rays at
    float32 matmul precision "highest" (what its fp32 results mean) and "medium" (utils/io.py:487-494 lets the user
    pick; TF32 tensor cores).  Device-timed, 1 warm-up + 2 runs each."""
    res = {}
    saved = torch.get_float32_matmul_precision()
    try:
        ro, rd, vd, _ = _c2_sample(n_rays)
        ro, rd, vd = ro.to(dev), rd.to(dev), vd.to(dev)
        kind, fn = reference_render_fn(dev)
        with torch.device(dev):          # the reference creates a few constants without device= (rendering__.py:321,335)
            for prec in ("highest", "medium"):
                torch.set_float32_matmul_precision(prec)
                best = None
                with torch.no_grad():
                    for i in range(3):
                        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        e0.record()
                        out = fn(ro, rd, vd)
                        e1.record()
                        torch.cuda.synchronize()
                        if i:
                            t = e0.elapsed_time(e1)
                            best = t if best is None else min(best, t)
                res[prec] = {"value": n_rays / (best * 1e-3), "unit": "rays/s", "ms": best}
        res.update({"kind": kind, "sample": "%d random rays of the C2 view, eager PyTorch on cuda" % n_rays})
    except Exception as e:       # reported, never fatal: this is context, not the product
        res = {"unavailable": "%s: %s" % (type(e).__name__, str(e)[:200])}
    finally:
        torch.set_float32_matmul_precision(saved)
    return res


def hbm_kernels(dev, R=160000):
    """Achieved HBM GB/s of the sampling / compositing kernels of the C2 path (north star: ">= 60 % of HBM peak on the
    sampling / compositing kernels"), each timed alone: CUDA events, inputs resident in HBM, a 2 GB fill between launches
    (flushes the 126 MB L2 and keeps the device busy while the host enqueues the timed call).  Bytes = the kernel's
    algorithmic reads + writes; peak = MEASURED_PEAKS.json hbm_gbs (copy bandwidth).  `R` rays = one
    chunk of the C2 view."""
    from star_b200 import functional as F_
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak = json.load(open(pk))["hbm_gbs"] if os.path.isfile(pk) else 6650.0
    flush = torch.empty(2048 << 20, dtype=torch.uint8, device=dev)

    def timeit(fn, n=8):
        for _ in range(3):
            fn()
        ms = 0.0
        for i in range(n):
            flush.fill_(i)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ms += e0.elapsed_time(e1)
        return ms / n

    res = {"peak": peak, "unit": "GB/s", "peak_source": "MEASURED_PEAKS.json hbm_gbs" if os.path.isfile(pk) else "fallback",
           "rays": R, "l2": "2 GB fill between launches"}

    def rec(name, ms, nbytes):
        gbs = nbytes / ms / 1e6
        res[name] = {"ms": round(ms, 4), "achieved": round(gbs, 1), "frac": round(gbs / peak, 3), "bytes": nbytes}
    try:
        g = torch.Generator(device=dev).manual_seed(0)
        ro = torch.randn(R, 3, device=dev, generator=g)
        rd = torch.nn.functional.normalize(torch.randn(R, 3, device=dev, generator=g), dim=-1) * 1.1
        with torch.no_grad():
            rec("sample_pts_64", timeit(lambda: F_.sample_pts(ro, rd, NEAR, FAR, NC)), R * (24 + 16 * NC))
            _pts, z = F_.sample_pts(ro, rd, NEAR, FAR, NC)
            ra = torch.randn(R, NC, device=dev, generator=g)
            rc = torch.randn(R, NC, 3, device=dev, generator=g)
            rec("composite_single_fwd_64", timeit(lambda: F_.CompositeSingle.apply(ra, rc, z, rd, 1e10, True)), R * NC * 28)
            rec("composite_hier_fused_64_128", timeit(lambda: F_.composite_hier(ra, rc, z, rd, 1e10, True, NI, True)),
                R * (28 * NC + 4 * NI + 4 * (NC + NI)))
            w = F_.CompositeSingle.apply(ra, rc, z, rd, 1e10, True)[4]
            _zs, zall, _zstd, _p = F_.hierarchical(z, w, NI, True, ro, rd)
            S = NC + NI
            ra = torch.randn(R, S, device=dev, generator=g)
            rc = torch.randn(R, S, 3, device=dev, generator=g)
            rec("composite_single_fwd_192", timeit(lambda: F_.CompositeSingle.apply(ra, rc, zall, rd, 1e10, True)), R * S * 28)
    except Exception as e:       # reported, never fatal
        res["unavailable"] = "%s: %s" % (type(e).__name__, str(e)[:200])
    del flush
    return res


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    # each step is a bounded sample; the whole --steps K run is sized to stay within a few minutes
    n = args.cpu_sample_rays or max(256, cpu_sample_default(args.mode, threads) * 3 // max(3, args.steps) // 256 * 256)
    fn = cpu_render_rays_per_s if args.mode == "render" else cpu_train_rays_per_s
    for _ in range(min(args.warmup, 1)):
        fn(max(64, n // 8), threads)
    times, kind = [], "port"
    for _ in range(args.steps):
        r = fn(n, threads)
        times.append(r[1])
        kind = r[2]
    dt = sum(times) / len(times)
    val = n / dt
    line = {
        "impl": "reference", "metric": metric_name(args.mode), "value": val, "unit": "rays/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, "fp32"),
        "cpu_baseline": {"value": val, "unit": "rays/s", "cores": threads, "kind": kind,
                         "sample": "%d random rays of the %s workload per step" % (n, args.mode)},
        "e2e": {"value": val, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def cpu_sample_default(mode, threads):
    """Bounded CPU sample sized for roughly 10-20 s of host work (about 120 rays/s/core render, 40 train)."""
    per_core = 120 if mode == "render" else 40
    return int(max(512, min(65536, 12 * per_core * threads)) // 256 * 256)


def metric_name(mode):
    return "rays/sec (render fwd)" if mode == "render" else "rays/sec (train fwd+bwd)"


def workload_config(args, precision):
    wk = getattr(args, "workload", "c2")
    if wk == "c4":
        R = args.rays or (8192 if args.mode == "train" else 65536)
        wl = ("C4 carla_star_online_multi: static + 5 rigid objects (7-vector poses), coarse+fine 256+256, %d CARLA-shaped "
              "rays/step/GPU, %s" % (R, "training step (fwd+bwd to all MLP weights and the poses, 3 regularisers)"
                                     if args.mode == "train" else "eval render"))
        return {"workload": wl, "N_samples": 256, "N_importance": 256, "num_vehicles": 5, "mlp_precision": precision,
                "parallelism": "ray-sharded x%d" % args.gpus, "l2": "L2 flushed between timed iterations (256 MB write)"}
    if wk == "c3":
        R = args.rays or (4096 if args.mode == "train" else 65536)
        wl = ("C3 carla_star_app_init: static + 1 rigid object net, appearance initialisation (static nets on the path), "
              "coarse+fine 256+256, ray chunk 5000, %d CARLA-shaped rays/step/GPU, %s" % (
                  R, "training step (fwd+bwd, img2mse(rgb0) + img2mse(rgb))" if args.mode == "train" else "eval render"))
        return {"workload": wl, "N_samples": 256, "N_importance": 256, "num_vehicles": 1, "mlp_precision": precision,
                "parallelism": "ray-sharded x%d" % args.gpus, "l2": "L2 flushed between timed iterations (256 MB write)"}
    if wk == "c5":
        R = args.rays or 16384
        wl = ("C5 carla_star_app_init_mip: mip-NeRF field (IPE), 256+512 frustums, %d CARLA-shaped rays/step/GPU, %s" % (
            R, "training step" if args.mode == "train" else "eval render"))
        return {"workload": wl, "N_samples": 256, "N_importance": 512, "mlp_precision": precision,
                "parallelism": "ray-sharded x%d" % args.gpus, "l2": "L2 flushed between timed iterations (256 MB write)"}
    if args.mode == "render":
        wl = "C2 lego vanilla NeRF coarse+fine 64+128, full %dx%d view render (%d rays/step/GPU), eval, perturb=0" % (
            args.hw, args.hw, args.hw * args.hw)
    else:
        wl = "C2 lego vanilla NeRF coarse+fine 64+128, %d-ray training step (fwd+bwd, all MLP weights)" % args.train_rays
    return {"workload": wl, "N_samples": NC, "N_importance": NI, "mlp_precision": precision,
            "parallelism": "ray-sharded x%d" % args.gpus,
            "l2": "inputs+activations per step exceed the 126 MB L2" if args.mode == "render" else
                  "L2 flushed between timed iterations (256 MB write)"}


# ------------------------------------------------------------------------------------------------ workloads
def build_workload(args, dev, rank, world):
    """Returns (net, ro_h, rd_h, step_device, params, d2h) for args.workload / args.mode.
    step_device(ro, rd) runs one pass from device-resident rays; d2h(out) reads the step's result back to the host."""
    import star_b200
    from star_b200.models import rendering__ as R_, loss as L_
    from oracle import ref_harness, star_oracle as so  # argument / synthetic-input builders only; no compute
    train = args.mode == "train"
    wk = args.workload
    if wk == "c2":
        net = star_b200.STaR(ref_harness.make_args(num_vehicles=0, N_importance=NI, chunk=8192, white_bkgd=True))
        net.load_state_dict(make_params())
        net.to(dev)
        net.set_precision(args.precision)
        net.train(train)
        if train:
            ro_h, rd_h = lego_view(RENDER_HW, theta=30.0 + 7.0 * rank)
            g = torch.Generator().manual_seed(100 + rank)
            idx = torch.randperm(ro_h.shape[0], generator=g)[:args.train_rays]
            ro_h, rd_h = ro_h[idx].contiguous(), rd_h[idx].contiguous()
            u = torch.rand(args.train_rays, NI, generator=g).to(dev)
            target = torch.rand(args.train_rays, 3, generator=g).to(dev)
        else:
            ro_h, rd_h = lego_view(args.hw, theta=30.0 + 7.0 * rank)
        params = [p for p in net.parameters()]
        # N > 1: flat gradient buffer, the kernels accumulate into it, each net's range is all-reduced on a side stream
        # as soon as its backward is queued (the fine net's collective runs under the coarse net's backward)
        sync = star_b200.parallel.GradSync(net) if (train and world > 1) else None

        def step_device(ro, rd):
            vd = rd / rd.norm(dim=-1, keepdim=True)
            if train:
                if sync is None:
                    for p in params:
                        p.grad = None
                pts, z = R_.sample_pts(ro, rd, NEAR, FAR, NC, perturb=0, is_train=True)
                out = R_.render_star_appinit(net, pts, vd, z, ro, rd, NI, u=u)
                loss = L_.photometric_loss(out["rgb0"], out["rgb"], target)[0]      # train_online__.py:158-166, fused
                if sync is not None:
                    sync.scale_loss(loss).backward()
                    sync.finish()
                else:
                    loss.backward()
                return loss
            with torch.no_grad():
                pts, z = R_.sample_pts(ro, rd, NEAR, FAR, NC, perturb=0, is_train=False)
                out = R_.render_star_appinit(net, pts, vd, z, ro, rd, NI)
            return out
        return net, ro_h, rd_h, step_device, params

    if wk == "c3":
        # configs/carla_star_app_init.txt: num_vehicles = 1, 256 + 256 samples, chunk = 5000; train_app_init__.py:60-77
        Nc = Ni = 256
        R = args.rays or (4096 if train else 65536)
        ro_h, rd_h = so.carla_rays(R, seed=30 + rank)
        g = torch.Generator().manual_seed(300 + rank)
        target = torch.rand(R, 3, generator=g).to(dev)
        net = star_b200.STaR(ref_harness.make_args(num_vehicles=1, N_importance=Ni, chunk=5000, white_bkgd=False))
        net.load_state_dict(so.init_star_params(1, Ni, seed=0, bias_std=0.02))
        net.to(dev)
        net.set_precision(args.precision)
        net.train(train)
        params = [p for p in net.parameters()]
        u = torch.rand(R, Ni, generator=g).to(dev) if train else None
        sync = star_b200.parallel.GradSync(net) if (train and world > 1) else None

        def step_device(ro, rd):
            vd = rd / rd.norm(dim=-1, keepdim=True)
            if train:
                if sync is None:
                    for p in params:
                        p.grad = None
                pts, z = R_.sample_pts(ro, rd, 0.03, 0.8, Nc, perturb=0, is_train=True)
                out = R_.render_star_appinit(net, pts, vd, z, ro, rd, Ni, u=u)
                loss = L_.photometric_loss(out["rgb0"], out["rgb"], target)[0]
                if sync is not None:
                    sync.scale_loss(loss).backward()
                    sync.finish()
                else:
                    loss.backward()
                return loss
            with torch.no_grad():
                pts, z = R_.sample_pts(ro, rd, 0.03, 0.8, Nc, perturb=0, is_train=False)
                out = R_.render_star_appinit(net, pts, vd, z, ro, rd, Ni)
            return out
        return net, ro_h, rd_h, step_device, params

    V = 5
    R = args.rays or ((8192 if train else 65536) if wk == "c4" else 16384)
    ro_h, rd_h = so.carla_rays(R, seed=10 + rank)
    g = torch.Generator().manual_seed(200 + rank)
    target = torch.rand(R, 3, generator=g).to(dev)
    if wk == "c4":
        Nc = Ni = 256
        net = star_b200.STaR(ref_harness.make_args(num_vehicles=V, N_importance=Ni, chunk=8192, white_bkgd=False))
        net.load_state_dict(so.init_star_params(V, Ni, seed=0, bias_std=0.02))
        net.to(dev)
        net.set_precision(args.precision)
        net.train(train)
        pose = torch.nn.Parameter(so.random_poses7(V, seed=3).to(dev))
        params = [p for p in net.parameters()] + [pose]
        u = torch.rand(R, Ni, generator=g).to(dev) if train else None
        sync = star_b200.parallel.GradSync(net, extra_params=[pose]) if (train and world > 1) else None

        def step_device(ro, rd):
            vd = rd / rd.norm(dim=-1, keepdim=True)
            if train:
                if sync is None:
                    for p in params:
                        p.grad = None
                pts, z = R_.sample_pts(ro, rd, 0.03, 0.8, Nc, perturb=0, is_train=True)
                out = R_.render_star_online(net, pts, vd, z, ro, rd, Ni, pose, u=u)
                loss = L_.photometric_loss(out["rgb0"], out["rgb"], target)[0]
                for sfx in ("", "0"):        # configs/carla_star_online_multi.txt:72-74
                    loss = loss + 0.5 * (1e-3 * out["loss_alpha_entropy" + sfx] + 1e-3 * out["loss_dynamic_vs_static_reg" + sfx]
                                         + 1e-5 * out["loss_ray_reg" + sfx])
                if sync is not None:
                    sync.scale_loss(loss).backward()
                    sync.finish()
                else:
                    loss.backward()
                return loss
            with torch.no_grad():
                pts, z = R_.sample_pts(ro, rd, 0.03, 0.8, Nc, perturb=0, is_train=False)
                out = R_.render_star_online(net, pts, vd, z, ro, rd, Ni, pose)
            return out
        return net, ro_h, rd_h, step_device, params

    # c5: mip variant (static field only in app-init mode, star_mipnerf.py:262-300)
    import argparse as _ap
    from star_b200.models.star_mipnerf import STaR as MipSTaR
    from oracle import mip_oracle as mo
    margs = _ap.Namespace(num_vehicles=0, chunk=8192, far_dist=1e10, N_importance=512, N_samples=256, scale_factor=0.01,
                          near=3.0, far=80.0)
    net = MipSTaR(margs)
    net.load_state_dict(mo.init_mip_params(0, seed=0, gain=1.4, bias_std=0.02))
    net.to(dev)
    net.set_precision(args.precision if args.precision in MIP_TIERS else "fp32")
    net.train(train)
    params = [p for p in net.parameters()]
    t_rand = torch.rand(R, 257, generator=g).to(dev) if train else None
    u_rand = torch.rand(R, 513, generator=g).to(dev) if train else None

    def step_device(ro, rd):
        vd = rd / rd.norm(dim=-1, keepdim=True)
        if train:
            for p in params:
                p.grad = None
            out = net(ro, vd, None, t_rand=t_rand, u_rand=u_rand)
            _l, mse0, mse, _p0, _p1 = L_.photometric_loss(out["rgb0"], out["rgb"], target)
            loss = mse + 0.1 * mse0   # train_app_init_mip.py:60
            loss.backward()
            if world > 1:
                star_b200.parallel.allreduce_gradients(params)
            return loss
        with torch.no_grad():
            out = net(ro, vd, None)
        return out
    return net, ro_h, rd_h, step_device, params


MIP_TIERS = ("fp32", "bf16", "fp16")      # precision tiers of the mip field (16-bit tiers: forward + weight gradients)


# ------------------------------------------------------------------------------------------------ B200 arm
def run_b200(args):
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    line = measure(args)
    if args.mode == "render" and not args.no_train_extra and args.workload == "c2":
        # the metric also names the training step: measure it on the same box and attach it to the line
        a2 = argparse.Namespace(**vars(args))
        a2.mode, a2.no_cpu_baseline = "train", True
        t = measure(a2)
        if rank == 0:
            line["train"] = {k: t[k] for k in ("metric", "value", "unit", "ms_per_step", "e2e", "roofline", "gpu_launches")}
            line["train"]["config"] = t["config"]["workload"]
        a2.opt_step = True
        t = measure(a2)
        if rank == 0:
            line["train"]["with_optimizer"] = {
                "metric": "rays/sec (train fwd+bwd + clip + Adam + weight re-pack)", "value": t["value"], "unit": "rays/s",
                "ms_per_step": t["ms_per_step"], "gpu_launches": t["gpu_launches"]}
    if args.mode == "render" and args.workload == "c2" and not args.no_extras and not args.no_train_extra:
        # the other configurations BASELINE.json names, as sub-records of the same line (weak scaling like the headline:
        # every rank its own rays): C4 = static + 5 objects with pose gradients, C5 = mip field
        for name, wk, mode, rays in (("c3_train", "c3", "train", 4096), ("c3_train_r1000", "c3", "train", 1000),
                                     ("c4_train", "c4", "train", 8192), ("c4_train_r1000", "c4", "train", 1000),
                                     ("c4_render", "c4", "render", 65536), ("c5_render", "c5", "render", 16384),
                                     ("c5_frame_render", "c5", "render", 1280 * 720),      # one whole frame of configs[4]
                                     ("c5_train", "c5", "train", 4096)):
            a3 = argparse.Namespace(**vars(args))
            a3.workload, a3.mode, a3.rays, a3.no_cpu_baseline, a3.opt_step = wk, mode, rays, True, False
            a3.steps, a3.warmup = min(args.steps, 2 if rays > 500000 else 5), 3
            t = measure(a3)
            if rank == 0:
                line[name] = {k: t[k] for k in ("metric", "value", "unit", "ms_per_step", "e2e", "roofline", "gpu_launches",
                                                "dtype", "steps")}
                line[name]["config"] = t["config"]["workload"]
        if world > 1:
            st = measure_strong(args)
            if rank == 0:
                line["strong"] = st
        if rank == 0:
            line["accuracy"] = accuracy_vs_oracle(args, torch.device("cuda", local))
            if world == 1:
                line["hbm_kernels"] = hbm_kernels(torch.device("cuda", local))
                line["gpu_eager_baseline"] = gpu_eager_baseline(torch.device("cuda", local))
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def accuracy_vs_oracle(args, dev, n_rays=512):
    """Max deviation of the benched MLP tier from the fp32 CPU oracle on `n_rays` random rays of the C2 view (fine pass
    teacher-forced with the oracle's own fine samples): the numbers the north star bounds (2e-3 absolute on rgb / depth /
    weights and 0.05 dB PSNR for the 16-bit MLP, 1e-4 for fp32).  Checker use of oracle/, outside every timed region."""
    import star_b200
    from star_b200 import functional as F_
    from star_b200.models import rendering__ as R_
    from oracle import ref_harness, star_oracle as so
    torch.set_num_threads(os.cpu_count() or 1)
    sd = make_params()
    ro, rd, vd, _ = _c2_sample(n_rays, seed=3)
    cfg = so.StarConfig(0, NI, 8192, white_bkgd=True)
    with torch.no_grad():
        pts, z = so.sample_pts(ro, rd, NEAR, FAR, NC, is_train=False)
        ref = so.render_star(sd, cfg, pts, vd, z, ro, rd, NI, training=False, exact_sum=True)
        mid = 0.5 * (z[..., 1:] + z[..., :-1])
        zs = so.sample_pdf(mid, ref["weights0"][..., 1:-1], NI, det=True, exact_sum=True)
        net = star_b200.STaR(ref_harness.make_args(num_vehicles=0, N_importance=NI, chunk=8192, white_bkgd=True))
        net.load_state_dict(sd)
        net.to(dev).eval()
        net.set_precision(args.precision)
        c = lambda t: t.to(dev)
        pts_g, z_g = R_.sample_pts(c(ro), c(rd), NEAR, FAR, NC, perturb=0, is_train=False)
        out = R_.render_star_appinit(net, pts_g, c(vd), z_g, c(ro), c(rd), NI, z_samples=c(zs))
        F_.check_range()
    res = {"tier": args.precision, "rays": n_rays, "against": "oracle/star_oracle.py fp32 on the host (pinned to the reference)"}
    for k in ("rgb0", "rgb", "weights", "depth"):
        res["max_abs_" + k] = float((out[k].cpu() - ref[k]).abs().max())
    tgt = torch.rand(n_rays, 3, generator=torch.Generator().manual_seed(1))
    psnr = lambda a: float(-10.0 * torch.log10(((a.double() - tgt.double()) ** 2).mean()))
    res["psnr_shift_db"] = abs(psnr(out["rgb"].cpu()) - psnr(ref["rgb"]))
    res["bound"] = {"fp32": 1e-4}.get(args.precision, 2e-3)
    return res


def measure_strong(args):
    """Strong scaling at N ranks: (1) ONE 800x800 view split over the ranks by contiguous ray ranges, per-ray outputs
    all-gathered on every rank (parallel.render_sharded); (2) ONE 4096-ray training batch split over the ranks, gradients
    all-reduced (GradSync).  Device-timed, max over ranks; value = rays of the whole view / batch per second."""
    import torch.distributed as dist
    import star_b200
    from star_b200 import parallel as P
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    res = {}
    steps = min(args.steps, 5)

    def timed(fn):
        for _ in range(3):
            fn()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        dist.barrier()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / steps], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    # (1) one view
    a = argparse.Namespace(**vars(args))
    a.mode, a.workload = "render", "c2"
    _net, ro_h, rd_h, step_device, _ = build_workload(a, dev, 0, 1)      # rank 0's view on every rank
    ro, rd = ro_h.to(dev), rd_h.to(dev)
    ms = timed(lambda: P.render_sharded(step_device, ro, rd))
    R = ro.shape[0]
    res["render_view_split"] = {"metric": "rays/sec (render fwd, one %dx%d view split over %d GPUs, outputs all-gathered)" % (
        args.hw, args.hw, world), "value": R / (ms * 1e-3), "unit": "rays/s", "ms_per_step": ms, "scaling": "strong"}
    # (2) one batch
    a = argparse.Namespace(**vars(args))
    a.mode, a.workload, a.train_rays = "train", "c2", args.train_rays // world
    _net, ro_h, rd_h, step_device, _ = build_workload(a, dev, rank, world)
    ro, rd = ro_h.to(dev), rd_h.to(dev)
    ms = timed(lambda: step_device(ro, rd))
    res["train_batch_split"] = {"metric": "rays/sec (train fwd+bwd, one %d-ray batch split over %d GPUs, gradients all-reduced)" % (
        a.train_rays * world, world), "value": a.train_rays * world / (ms * 1e-3), "unit": "rays/s", "ms_per_step": ms,
        "scaling": "strong"}
    return res


def measure(args):
    """One timed measurement (render or train per args.mode); returns the JSON line as a dict on rank 0."""
    import torch.distributed as dist
    import star_b200
    from star_b200 import functional as F_
    from star_b200.models import rendering__ as R_
    from oracle import ref_harness  # make_args only (argparse.Namespace builder); no compute

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local)

    train = args.mode == "train"
    net, ro_h, rd_h, step_device, params = build_workload(args, dev, rank, world)
    if train and args.opt_step:
        # the whole optimisation step of the reference: + clip_grad_norm_(1.0) + Adam, fused, over
        # parameters re-homed into one flat buffer; the packed 16-bit weight images are rebuilt every step
        from star_b200 import optim as O_
        if getattr(net, "_star_flat", None) is None:      # (N > 1: GradSync already re-homed them)
            O_.flatten_parameters(net)
        opt = O_.FusedAdam(params, lr=5e-4, betas=(0.9, 0.999), max_grad_norm=1.0)
        fwd_bwd = step_device

        def step_device(ro, rd):
            loss = fwd_bwd(ro, rd)
            opt.step()
            return loss
    R = ro_h.shape[0]
    ro_h, rd_h = ro_h.pin_memory(), rd_h.pin_memory()
    ro, rd = ro_h.to(dev), rd_h.to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if (train or args.workload != "c2") else None

    def step_e2e():
        a = ro_h.to(dev, non_blocking=True)
        b = rd_h.to(dev, non_blocking=True)
        out = step_device(a, b)
        if train:
            return out.cpu()
        return out["rgb"].cpu(), out["depth"].cpu(), out["acc"].cpu()     # 20 B per ray

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step_device(ro, rd)
    barrier()

    # ---- timed region 1: device-resident inputs
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    F_.LAUNCH_COUNTER["calls"] = 0
    F_.PROFILE = {}
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * args.steps)]
    barrier()
    t_wall = time.perf_counter()
    for i in range(args.steps):
        if flush is not None:
            flush.fill_(i & 0xff)
        ev[2 * i].record()
        step_device(ro, rd)
        ev[2 * i + 1].record()
    barrier()
    t_wall = time.perf_counter() - t_wall
    launches = F_.LAUNCH_COUNTER["calls"]
    prof = F_.PROFILE
    F_.PROFILE = None
    ms = sum(ev[2 * i].elapsed_time(ev[2 * i + 1]) for i in range(args.steps)) / args.steps
    clk = clocks.stop() if rank == 0 else None

    # ---- timed region 2: end to end from pinned host memory through the public API
    step_e2e()
    barrier()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step_e2e()
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1) / args.steps
    ms_e2e_wall = (time.perf_counter() - t0) * 1e3 / args.steps
    ms_e2e = max(ms_e2e, ms_e2e_wall)      # the D2H read is synchronous: wall clock covers it

    t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])

    if rank == 0:
        total_rays = R * world
        # roofline of the dominant kernel (the fused MLP forward): algorithmic FLOPs / event time
        roof = None
        parts = {}
        for k, evs in (prof or {}).items():
            if evs:
                parts[k] = {"ms": sum(a.elapsed_time(b) for a, b, _, _ in evs), "samples": sum(n for _, _, n, _ in evs),
                            "flops": sum(f for _, _, _, f in evs), "launches": len(evs)}
        if parts:
            # dominant C-ABI call of the step and its algorithmic FLOPs (2 FLOP per MAC of the MLP layers, padding
            # excluded; backward = dX + dW = 2x forward), summed over the launches of the timed region
            key = max(parts, key=lambda k: parts[k]["ms"])
            tot_ms, tot_samples, n_l = parts[key]["ms"], parts[key]["samples"], parts[key]["launches"]
            peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.isfile(
                os.path.join(ROOT, "MEASURED_PEAKS.json")) else None
            peak_tf = (peak["bf16_tflops_sustained"] if peak else 1400.0)
            ach = parts[key]["flops"] / (tot_ms * 1e-3) / 1e12
            # dram__bytes_read + dram__bytes_write per launch: bench.py cannot read DRAM counters, so the per-sample figure
            # of the committed `ncu --set full` capture of this kernel (profiles/traffic.json, written by
            # tools/ncu_traffic.py from profiles/*_raw.csv; static net) is scaled to this run's average launch
            tj = os.path.join(ROOT, "profiles", "traffic.json")
            per_sample = (json.load(open(tj)) if os.path.isfile(tj) else {}).get(key, {}).get("dram_bytes_per_sample")
            traffic = per_sample * tot_samples / n_l if (args.precision != "fp32" and per_sample is not None
                                                         and args.workload == "c2") else None
            tier = args.precision if not key.startswith("mip") else (args.precision if args.precision in MIP_TIERS else "fp32")
            roof = {"bound": "tensor", "kernel": "star_%s (%s)" % (key, tier), "achieved": ach,
                    "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf, "traffic": traffic,
                    "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained" if peak else "fallback",
                    "launches": n_l, "avg_launch_ms": tot_ms / n_l,
                    "share_of_step": tot_ms / (ms * args.steps),
                    "calls_ms_per_step": {k: round(v["ms"] / args.steps, 3) for k, v in parts.items()}}
        line = {
            "metric": metric_name(args.mode), "value": total_rays / (ms * 1e-3), "unit": "rays/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "dtype": {"bf16": "bf16", "fp16": "f16", "fp32": "f32"}[args.precision],
            "data": "synthetic", "config": workload_config(args, args.precision),
            "e2e": {"value": total_rays / (ms_e2e * 1e-3), "unit": "rays/s", "h2d_bytes_per_step": 24 * R,
                    "d2h_bytes_per_step": (4 if train else 20 * R)},
            "gpu_launches": launches, "clocks": clk, "roofline": roof, "wall_ms_per_step": t_wall * 1e3 / args.steps,
        }
        if not args.no_cpu_baseline and world == 1 and args.workload == "c2":
            threads = os.cpu_count() or 1
            n = args.cpu_sample_rays or cpu_sample_default(args.mode, threads)
            fn = cpu_train_rays_per_s if train else cpu_render_rays_per_s
            fn(max(64, n // 8), threads)
            v, dt, kind = fn(n, threads)
            line["cpu_baseline"] = {"value": v, "unit": "rays/s", "cores": threads, "kind": kind,
                                    "sample": "%d random rays of the same workload, %.1f s" % (n, dt)}
        return line
    return None


def main():
    args = parse()
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ and args.impl != "reference":
        # plain `python bench.py --gpus N`: re-launch as one process per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", "29511", os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    if args.impl == "reference":
        run_reference(args)
    else:
        try:
            run_b200(args)
        except Exception:
            # a tensor-core kernel whose barrier watchdog fired says which wait hung (readable after the context died)
            try:
                import star_b200
                rep = star_b200._capi.watchdog_report()
                sys.stderr.write("[bench] kernel watchdog: %s | launch markers %s\n" % (rep or "-", star_b200._capi.launch_markers()))
            except Exception:
                pass
            raise


if __name__ == "__main__":
    main()

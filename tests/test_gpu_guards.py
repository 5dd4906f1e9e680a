"""Out-of-bounds canaries (compute-sanitizer is closed on this GPU pool, profiles/r2d_sanitizer_closed.txt): every
buffer a kernel family WRITES is carved out of an arena with guard words on both sides; after the launch the guards must be
intact and the payload fully overwritten where the contract says so.  Shapes are chosen off the tile grid (odd tile
counts, partial last tile, odd ray counts) on purpose."""
import ctypes as C

import pytest
import torch

import star_b200
from star_b200 import _capi, functional as F_
from oracle import ref_harness, star_oracle as so

pytestmark = pytest.mark.gpu
DEV = "cuda"
GUARD = 1024            # bytes on each side
PATTERN = 0xA5


class Arena:
    def __init__(self):
        self.items = []

    def take(self, nbytes, dtype=torch.uint8):
        nb = (int(nbytes) + 255) // 256 * 256
        buf = torch.full((nb + 2 * GUARD,), PATTERN, dtype=torch.uint8, device=DEV)
        self.items.append(buf)
        return buf[GUARD:GUARD + nb].view(dtype) if dtype != torch.uint8 else buf[GUARD:GUARD + nb]

    def check(self):
        torch.cuda.synchronize()
        for i, buf in enumerate(self.items):
            assert bool((buf[:GUARD] == PATTERN).all()) and bool((buf[-GUARD:] == PATTERN).all()), f"guard {i} overwritten"


@pytest.mark.parametrize("prec", ["fp16", "bf16", "fp32"])
@pytest.mark.parametrize("R,S,dyn", [(129, 1, False), (77, 37, True), (3, 43, False), (1031, 53, True)])
def test_mlp_forward_backward_stay_inside_their_buffers(prec, R, S, dyn):
    net = star_b200.STaR(ref_harness.make_args(num_vehicles=1, N_importance=8, chunk=4096))
    net.load_state_dict(so.init_star_params(1, 8, seed=3, bias_std=0.02))
    net.to(DEV)
    m = net.dynamic_coarse_nerfs[0] if dyn else net.static_coarse_nerf
    precision = _capi.PRECISIONS[prec]
    flat, packed = m._rt.refresh(_capi.PRECISIONS[prec])
    d = _capi.net_desc(m._rt.n_blocks, 10, 4, precision)
    L = _capi.lib()
    ro, rd = so.carla_rays(R, seed=1)
    vd = (rd / rd.norm(dim=-1, keepdim=True)).to(DEV)
    pts, z = so.sample_pts(ro, rd, 0.03, 0.8, S)
    pts = pts.to(DEV).contiguous()
    p12 = F_.pose_to_mat12(so.pose7_to_matrix(so.random_poses7(1, seed=9))[0].to(DEV)).contiguous() if dyn else None
    ar = Arena()
    n = R * S
    ra, rc = ar.take(4 * n, torch.float32), ar.take(12 * n, torch.float32)
    stash = ar.take(L.star_stash_bytes(C.byref(d), n))
    ws = ar.take(L.star_mlp_backward_workspace_bytes(C.byref(d), n))
    grad = ar.take(4 * flat.numel(), torch.float32)
    pacc = ar.take(128, torch.float32)
    grad.zero_(), pacc.zero_()
    st = torch.cuda.current_stream().cuda_stream
    ptr = lambda t: t.data_ptr() if t is not None else None
    rc_ = L.star_mlp_forward(C.byref(d), ptr(packed), ptr(pts), None, None, None, ptr(vd), ptr(p12), None, None, R, S, ptr(ra),
                             ptr(rc), S, ptr(stash), None, st)
    assert rc_ == 0
    ar.check()
    assert bool(torch.isfinite(ra[:n]).all()) and bool(torch.isfinite(rc[:3 * n]).all())
    ga, gc = torch.randn(n, device=DEV), torch.randn(3 * n, device=DEV)
    rc_ = L.star_mlp_backward(C.byref(d), ptr(packed), ptr(flat), ptr(pts), None, None, None, ptr(vd), ptr(p12), None, None, R,
                              S, ptr(ga), ptr(gc), S, ptr(stash), ptr(ws), ptr(grad), ptr(pacc) if dyn else None, st)
    assert rc_ == 0
    ar.check()
    assert bool(torch.isfinite(grad[:flat.numel()]).all()) and float(grad[:flat.numel()].abs().max()) > 0


@pytest.mark.parametrize("V", [0, 3])
def test_single_call_render_stays_inside_its_buffers(V):
    """star_render_forward with every output and the workspace between guards (R off the tile grid; camera mode for
    V = 0 so that ray generation and depth sampling write too)."""
    Nc, Ni = 12, 20
    net = star_b200.STaR(ref_harness.make_args(num_vehicles=V, N_importance=Ni, chunk=50))
    net.load_state_dict(so.init_star_params(V, Ni, seed=4, bias_std=0.02))
    net.to(DEV).eval()
    net.set_precision("fp16")
    L = _capi.lib()
    H, W = 7, 9
    R = H * W
    prec = _capi.PREC_F16
    cfg = _capi.StarRenderCfg(R, Nc, Ni, V, prec, 4, 2, 10, 4, 0, 0, 1, 50, 0.03, 0.8, 1e10)
    ar = Arena()
    f = lambda *s: ar.take(4 * int(torch.tensor(s).prod()), torch.float32)

    def multi(S):
        o = dict(rgb=f(R, 3), disp=f(R), acc=f(R), depth=f(R), weights=f(R, S))
        if V:
            o.update(rgb_static=f(R, 3), depth_static=f(R), rgb_dynamic=f(R, V, 3), depth_dynamic=f(R, V),
                     dynamic_transmittance=f(R, V), rgb_dynamic_all=f(R, 3), regs=f(5))
        return _capi.StarMultiOut(*[o[k].data_ptr() if k in o else None for k in F_.STAR_OUT_KEYS]), o
    mc, oc = multi(Nc)
    mf, of = multi(Nc + Ni)
    extra = dict(dists0=f(R, Nc) if V == 0 else None, dists=f(R, Nc + Ni) if V == 0 else None, z_vals0=f(R, Nc),
                 z_vals=f(R, Nc + Ni), z_samples=f(R, Ni), z_std=f(R), rays_o=f(R, 3), rays_d=f(R, 3), viewdirs=f(R, 3))
    out = _capi.StarRenderOut(mc, mf, *[extra[k].data_ptr() if extra[k] is not None else None for k in
                                        ("dists0", "dists", "z_vals0", "z_vals", "z_samples", "z_std", "rays_o", "rays_d",
                                         "viewdirs")])
    ws_bytes = L.star_render_workspace_bytes(C.byref(cfg))
    ws = ar.take(ws_bytes)
    pk = lambda m: m._rt.refresh(prec)[1]
    keep = [pk(net.static_coarse_nerf), pk(net.static_fine_nerf)] + [pk(m) for m in net.dynamic_coarse_nerfs] + \
        [pk(m) for m in net.dynamic_fine_nerfs]
    arr_c = (C.c_void_p * max(V, 1))(*[pk(m).data_ptr() for m in net.dynamic_coarse_nerfs])
    arr_f = (C.c_void_p * max(V, 1))(*[pk(m).data_ptr() for m in net.dynamic_fine_nerfs])
    c2w = torch.eye(4, device=DEV)[:3].contiguous()
    t_vals = torch.linspace(0, 1, Nc).to(DEV)
    u_det = torch.linspace(0, 1, Ni).to(DEV)
    pose12 = torch.stack([F_.pose_to_mat12(p) for p in so.random_poses7(max(V, 1), seed=5).to(DEV)]).contiguous()
    sin = _capi.StarRenderIn(None, None, None, H, W, 0, H, 10.0, 10.0, 4.5, 3.5, c2w.data_ptr(), None, None, t_vals.data_ptr(),
                             None, None, u_det.data_ptr(), None, pose12.data_ptr() if V else None, None, None,
                             keep[0].data_ptr(), keep[1].data_ptr(), arr_c, arr_f)
    rc = L.star_render_forward(C.byref(cfg), C.byref(sin), C.byref(out), ws.data_ptr(), ws_bytes, None,
                               torch.cuda.current_stream().cuda_stream)
    assert rc == 0, L.star_error_string(rc)
    ar.check()
    for o in (oc, of):
        for k, v in o.items():
            assert bool(torch.isfinite(v[:1]).all()) and not bool((v.view(torch.uint8)[:4] == PATTERN).all()), k
    # a workspace one byte too small is refused, not overrun
    assert L.star_render_forward(C.byref(cfg), C.byref(sin), C.byref(out), ws.data_ptr(), ws_bytes - 1, None,
                                 torch.cuda.current_stream().cuda_stream) == 5

"""CPU-only checks of the host side: the C-ABI library loads and exports every symbol the header
declares, the module tree keeps the reference checkpoint layout, pose / BARF host helpers match the
oracle.  No compute calls (there is no GPU here and no CPU fallback by design)."""
import os
import re

import pytest
import torch

import star_b200
from star_b200 import _capi, functional as F_
from star_b200.models import embedder as emb
from oracle import ref_harness, star_oracle as so
from helpers import assert_close

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "star_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(star_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 20
    lib = _capi.lib()
    for name in declared:
        assert hasattr(lib, name), f"libstar_b200.so does not export {name}"
    assert declared == set(_capi.EXPORTED_SYMBOLS), declared ^ set(_capi.EXPORTED_SYMBOLS)
    assert lib.star_abi_version() == _capi.ABI_VERSION


def test_python_constants_match_the_header():
    """The precision tiers and flag bits of include/star_b200.h and the status codes the binding names must be the same
    numbers in _capi.py (the C ABI is the contract; ctypes passes plain ints)."""
    import re
    hdr = open(os.path.join(ROOT, "include", "star_b200.h")).read()
    defs = {m.group(1): int(m.group(2), 0) for m in re.finditer(r"#define\s+(STAR_\w+)\s+(0x[0-9a-fA-F]+|\d+)\b", hdr)}
    enum = re.search(r"enum\s*\{\s*STAR_PREC_F32\s*=\s*(\d+),\s*STAR_PREC_BF16\s*=\s*(\d+),\s*STAR_PREC_F16\s*=\s*(\d+)", hdr)
    assert enum and tuple(int(x) for x in enum.groups()) == (_capi.PREC_F32, _capi.PREC_BF16, _capi.PREC_F16)
    assert defs["STAR_PREC_FLAG_RETIRED"] == _capi.PREC_FLAG_RETIRED
    assert defs["STAR_PREC_FLAG_DX_PIPELINED"] == _capi.PREC_FLAG_DX_PIPELINED
    assert defs["STAR_PREC_FLAG_NO_WSHARE"] == _capi.PREC_FLAG_NO_WSHARE
    flags = [_capi.PREC_FLAG_RETIRED, _capi.PREC_FLAG_DX_PIPELINED, _capi.PREC_FLAG_NO_WSHARE]
    assert all(f & 0xff == 0 for f in flags) and len(set(flags)) == 3          # the low byte is the tier
    assert (_capi.PREC_FLAG_DX_PIPELINED & _capi.PREC_FLAG_NO_WSHARE) == 0


def test_param_counts_and_packed_sizes():
    import ctypes as C
    lib = _capi.lib()
    for nb, n in ((4, 711300), (2, 448132)):
        d = _capi.net_desc(nb, 10, 4, _capi.PREC_F32)
        assert lib.star_net_param_count(C.byref(d)) == n
        assert lib.star_packed_bytes(C.byref(d)) > 4 * n
        assert lib.star_stash_bytes(C.byref(d), 1000) > 0
    bad = _capi.net_desc(9, 10, 4, _capi.PREC_F32)
    assert lib.star_net_param_count(C.byref(bad)) == 0


def test_no_cpu_fallback():
    x = torch.zeros(4, 3)
    with pytest.raises(_capi.StarError):
        F_.embed(x, 4)


def test_state_dict_layout_matches_reference_keys():
    args = ref_harness.make_args(num_vehicles=2, N_importance=8)
    net = star_b200.STaR(args)
    sd = so.init_star_params(2, 8, seed=0)
    assert set(net.state_dict().keys()) == set(sd.keys())
    net.load_state_dict(sd, strict=True)
    for k, v in net.state_dict().items():
        assert v.shape == sd[k].shape
    n_static = sum(p.numel() for p in net.static_coarse_nerf.parameters())
    n_dyn = sum(p.numel() for p in net.dynamic_coarse_nerfs[0].parameters())
    assert (n_static, n_dyn) == (711300, 448132)
    if ref_harness.reference_available():
        ref = ref_harness.load_reference()
        rnet = ref.star.STaR(args)
        assert list(rnet.state_dict().keys()) == list(net.state_dict().keys())
        assert len(net.get_nerf_params()) == len(rnet.get_nerf_params())


def test_reference_init_statistics():
    """Same initialisers as the reference: fc_1.weight zero, biases zero, kaiming fan-in scale."""
    torch.manual_seed(0)
    net = star_b200.STaR(ref_harness.make_args(num_vehicles=1, N_importance=8))
    n = net.static_coarse_nerf
    assert float(n.pts_net.blocks[0].fc_1.weight.abs().max()) == 0.0
    assert float(n.pts_net.lin_in.bias.abs().max()) == 0.0
    assert abs(float(n.pts_net.blocks[1].fc_0.weight.std()) - (2.0 / 256) ** 0.5) < 2e-3
    assert abs(float(n.pts_net.lin_in.weight.std()) - (2.0 / 63) ** 0.5) < 5e-3


def test_flat_master_order_and_versioning():
    net = star_b200.STaR(ref_harness.make_args(num_vehicles=0, N_importance=0))
    rt = net.static_coarse_nerf._rt
    ps = rt.ordered_params()
    names = {id(p): k for k, p in net.static_coarse_nerf.named_parameters()}
    order = [names[id(p)] for p in ps]
    assert order[:2] == ["pts_net.lin_in.weight", "pts_net.lin_in.bias"]
    assert order[2:6] == ["pts_net.blocks.0.fc_0.weight", "pts_net.blocks.0.fc_0.bias",
                          "pts_net.blocks.0.fc_1.weight", "pts_net.blocks.0.fc_1.bias"]
    assert order[-8:] == ["alpha_linear.weight", "alpha_linear.bias", "feature_linear.weight", "feature_linear.bias",
                          "views_linears.0.weight", "views_linears.0.bias", "rgb_linear.weight", "rgb_linear.bias"]
    assert sum(p.numel() for p in ps) == 711300


def test_pose7_gradient_convention_matches_oracle_pypose_restatement():
    torch.manual_seed(1)
    p7 = so.random_poses7(1, seed=4)[0]
    pts, dirs = torch.randn(30, 3), torch.randn(7, 3)
    wp, wd = torch.randn(30, 3), torch.randn(7, 3)
    a = p7.clone().requires_grad_(True)
    ((so.se3_act(a, pts) * wp).sum() + (so.so3_act(a[3:], dirs) * wd).sum()).backward()
    b = p7.clone().requires_grad_(True)
    m12 = F_.pose_to_mat12(b)
    Rm, t = m12[:9].view(3, 3), m12[9:]
    ((((Rm @ pts.T).T + t) * wp).sum() + ((Rm @ dirs.T).T * wd).sum()).backward()
    assert_close(b.grad, a.grad, 1e-5, 1e-5)
    assert float(b.grad[6]) == 0.0


def test_pose_matrix_slicing_is_euclidean():
    M = so.pose7_to_matrix(so.random_poses7(1, seed=5))[0].requires_grad_(True)
    m12 = F_.pose_to_mat12(M)
    m12.sum().backward()
    assert float(M.grad[3].abs().sum()) == 0.0 and float(M.grad[:3].sum()) == 12.0


@pytest.mark.parametrize("L,step", [(10, 13), (4, 13), (10, 0), (10, 99)])
def test_barf_scale_vector_matches_oracle(L, step):
    x = (torch.rand(11, 3) * 2 - 1)
    plain = so.embed(x, L)
    want = so.embed(x, L, step=step, end_barf=40)
    s = emb.barf_scale_vector(step, 40, L)
    assert_close(plain * s, want, 1e-7)


def test_ray_chunking_covers_all_rays():
    for R, S in ((1, 64), (1000, 512), (640000, 192), (5, 1 << 20)):
        ch = F_._ray_chunks(R, S)
        assert ch[0][0] == 0 and ch[-1][1] == R
        assert all(a[1] == b[0] for a, b in zip(ch, ch[1:]))
        assert all((b - a) * S <= max(S, F_.MAX_SAMPLES_PER_LAUNCH) for a, b in ch)


def test_install_registers_reference_module_names():
    import sys
    saved = {k: sys.modules.get(k) for k in list(sys.modules) if k == "models" or k.startswith("models.")}
    for k in saved:
        sys.modules.pop(k)
    try:
        star_b200.install()
        from models.star__ import STaR
        from models.rendering__ import sample_pts, render_star_online, render_star_appinit, mse2psnr, img2mse, to8b, get_rays, get_rays_np  # noqa: F401
        assert STaR is star_b200.STaR
    finally:
        for k in [k for k in sys.modules if k == "models" or k.startswith("models.")]:
            sys.modules.pop(k)
        sys.modules.update({k: v for k, v in saved.items() if v is not None})


# ------------------------------------------------------------------------------------------ mip variant (row a12), host side
def test_mip_module_tree_param_count_and_install():
    import argparse
    from oracle import mip_oracle as mo
    from star_b200.models.star_mipnerf import STaR as MipSTaR
    args = argparse.Namespace(num_vehicles=2, chunk=64, far_dist=1e10, N_importance=8, N_samples=4, scale_factor=0.01,
                              near=3.0, far=80.0)
    net = MipSTaR(args)
    sd = mo.init_mip_params(2, seed=0)
    assert set(net.state_dict().keys()) == set(sd.keys())
    net.load_state_dict(sd, strict=True)
    lib = _capi.lib()
    n = sum(p.numel() for p in net.static_nerf.parameters())
    assert n == lib.star_mip_param_count() == 589572
    assert sum(o * i for (o, i) in mo.mip_param_shapes().values()) == F_.MIP_MAC_PER_SAMPLE
    assert net.near_plane == pytest.approx(0.03) and net.far_plane == pytest.approx(0.8)
    assert len(net.get_nerf_params()) == 3 * 2 * len(mo.mip_param_shapes())
    for prec in (_capi.PREC_F32, _capi.PREC_BF16, _capi.PREC_F16):
        assert lib.star_mip_packed_bytes(prec) > 2 * n
    # fp32 tier: fp32 activations; 16-bit tiers: the 16 KB operand blocks of the tensor-core backward (smaller)
    assert lib.star_mip_stash_bytes(_capi.PREC_F32, 128) > lib.star_mip_stash_bytes(_capi.PREC_BF16, 128) > 0
    assert lib.star_mip_stash_bytes(_capi.PREC_F16, 128) == lib.star_mip_stash_bytes(_capi.PREC_BF16, 128)
    star_b200.install("models_under_test")
    import importlib
    assert importlib.import_module("models_under_test.star_mipnerf").STaR is MipSTaR
    assert importlib.import_module("models_under_test.rendering__").sample_pts is star_b200.rendering__.sample_pts


def test_c_abi_argument_checks_need_no_gpu():
    """Every entry point validates pointers / shapes before touching the device: status codes, never a crash."""
    import ctypes as C
    lib = _capi.lib()
    assert lib.star_sample_pts(None, None, None, None, 0.0, 1.0, 4, 8, 0, None, None, None) == 3          # STAR_E_NULL
    assert lib.star_get_rays(4, 4, 1.0, 1.0, 2.0, 2.0, None, 0, 4, None, None, None, None) == 3
    assert lib.star_mip_uniform_bins(None, None, 0.0, 1.0, 4, 8, None, None, None) == 3
    assert lib.star_mip_pdf_sample(None, None, 0, None, None, 0.0, 1.0, 4, 8, 8, None, None, None, None, None) == 3
    assert lib.star_mip_field_forward(0, None, None, None, None, None, None, 0.5, 4, 8, None, None, 8, None, None) == 3
    assert lib.star_mip_composite_single_forward(None, None, None, 4, 8, None, None, None, None, None) == 3
    assert lib.star_mip_pack_weights(7, C.c_void_p(16), None, C.c_void_p(16), None) == 2                    # STAR_E_UNSUPPORTED
    assert lib.star_error_string(2).decode().startswith("unsupported")
    d = _capi.net_desc(4, 10, 4, _capi.PREC_F32)
    assert lib.star_mlp_forward(C.byref(d), None, None, None, None, None, None, None, None, None, 4, 8, None, None, 8, None, None, None) == 3


def test_sass_of_the_mlp_kernels_is_tcgen05_tmem_and_bulk_copies():
    """What the built library executes on the hot path, read from its SASS (cuobjdump, no GPU): the MLP kernels issue
    tcgen05.mma (UTCHMMA) with accumulators in TMEM (LDTM), fetch weights with bulk async copies (UBLKCP) -- the inference
    kernel of the weight-sharing launch with the MULTICAST form and multicast commits (UTCBAR.MULTICAST) inside a cluster
    (UCGABAR) -- and contain no warp-level HMMA (mma.sync / wmma), the recompiled-baseline path."""
    import shutil
    import subprocess
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    lib = os.path.join(ROOT, "3d-mot-using-neural-radiance-fields_b200", "libstar_b200.so")
    names = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
    funs = re.findall(r"Function (\S+):", names)

    def sass(pred):
        f = [n for n in funs if pred(n)]
        assert f, "kernel not found in the library"
        out = subprocess.run(["cuobjdump", "-sass", "-fun", f[0], lib], capture_output=True, text=True).stdout
        return [m.group(1) for m in re.finditer(r"^\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\w+\s+)?([A-Z][\w.]*)", out, re.M)]

    fwd_shared = sass(lambda n: n.startswith("_Z17mlp_fwd_tc_kernelILb1ELb0ELb1E"))       # fp16, inference, weight sharing
    fwd_plain = sass(lambda n: n.startswith("_Z17mlp_fwd_tc_kernelILb1ELb0ELb0E"))
    dx = sass(lambda n: n.startswith("_Z17mlp_bwd_tc_kernelILb1ELb0ELb0E"))
    dw = sass(lambda n: n.startswith("_Z12dw_tc_kernelILb1E"))
    for name, ops in (("forward (shared weights)", fwd_shared), ("forward", fwd_plain), ("dX chain", dx), ("dW", dw)):
        assert any(o.startswith("UTCHMMA") for o in ops), name + ": no tcgen05.mma"
        assert any(o.startswith("LDTM") for o in ops), name + ": no TMEM loads"
        assert any(o.startswith("UBLKCP") for o in ops), name + ": no bulk async copies"
        assert not any(o.startswith("HMMA") for o in ops), name + ": warp-level HMMA found"
    assert any(o.startswith("UBLKCP") and "MULTICAST" in o for o in fwd_shared)
    assert any(o.startswith("UTCBAR") and "MULTICAST" in o for o in fwd_shared)
    assert any(o.startswith("UCGABAR") for o in fwd_shared)
    assert not any("MULTICAST" in o for o in fwd_plain)


def test_register_budget_of_the_576_thread_kernels():
    """The persistent tensor-core kernels run 576 (608 with the optional stash copy warp) threads per CTA: more than 96
    registers per thread and the launch fails at run time (warps are allocated in fours: 20 x 32 x 96 = 61 440 of 65 536).
    Read from the built library, no GPU."""
    import shutil
    import subprocess
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    lib = os.path.join(ROOT, "3d-mot-using-neural-radiance-fields_b200", "libstar_b200.so")
    out = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
    usage = dict(re.findall(r"Function (\S+):\s*\n\s*REG:(\d+)", out))
    big = {k: int(v) for k, v in usage.items()
           if re.match(r"_Z1[78]m(lp|ip)_(fwd|bwd|bwd2)_tc_kernel", k)}
    assert len(big) >= 20, sorted(big)
    for k, r in big.items():
        assert r <= 96, (k, r)
    hg = [int(v) for k, v in usage.items() if k.startswith("_Z19head_grad_tc_kernel")]
    assert hg and hg[0] <= 128          # two CTAs of 256 threads per SM


def test_c_abi_empty_batches_return_ok_without_touching_pointers():
    """R = 0: an empty batch carries no pointers (torch's data_ptr() of an empty tensor is 0), so every per-ray entry returns
    STAR_OK before its NULL checks -- and before any CUDA call, which is why this runs without a GPU."""
    import ctypes as C
    lib = _capi.lib()
    assert lib.star_sample_pts(None, None, None, None, 0.0, 1.0, 0, 8, 0, None, None, None) == 0
    assert lib.star_sample_pts(None, None, None, None, 0.0, 1.0, -1, 8, 0, None, None, None) != 0
    assert lib.star_mip_uniform_bins(None, None, 0.0, 1.0, 0, 8, None, None, None) == 0
    assert lib.star_mip_pdf_sample(None, None, 0, None, None, 0.0, 1.0, 0, 8, 8, None, None, None, None, None) == 0
    assert lib.star_mip_field_forward(0, None, None, None, None, None, None, 0.5, 0, 8, None, None, 8, None, None) == 0
    assert lib.star_mip_composite_single_forward(None, None, None, 0, 8, None, None, None, None, None) == 0
    d = _capi.net_desc(4, 10, 4, _capi.PREC_F32)
    assert lib.star_mlp_forward(C.byref(d), None, None, None, None, None, None, None, None, None, 0, 8, None, None, 8, None,
                                None, None) == 0
    for prec in (_capi.PREC_F32, _capi.PREC_F16, _capi.PREC_F16 | _capi.PREC_FLAG_NO_WSHARE):
        cfg = _capi.StarRenderCfg(0, 64, 128, 0, prec, 4, 2, 10, 4, 1, 0, 1, 8192, 2.0, 6.0, 1e10)
        assert lib.star_render_forward(C.byref(cfg), None, None, None, 0, None, None) == 0


def test_train_step_abi_argument_checks_and_host_logic():
    """status codes of the loss / optimiser entry points without a GPU, the run-merging and
    parameter re-homing logic of optim.py (pure host code), and the no-fallback rule."""
    import ctypes as C
    from star_b200 import optim as O_
    from star_b200.models import loss as L_
    lib = _capi.lib()
    assert lib.star_train_ws_bytes() >= 16
    assert lib.star_photometric_loss(None, None, None, 12, None, None, None, None, None) == 3               # STAR_E_NULL
    assert lib.star_photometric_loss(None, C.c_void_p(16), C.c_void_p(16), 0, C.c_void_p(16), None, None,
                                     C.c_void_p(16), None) == 1                                             # STAR_E_BAD_SHAPE
    assert lib.star_photometric_loss(None, C.c_void_p(16), C.c_void_p(16), 3, C.c_void_p(16), None, None,
                                     C.c_void_p(8), None) == 4                                              # STAR_E_ALIGN
    assert lib.star_depth_loss_forward(None, None, 4, 0.0, 1.0, None, None, None) == 3
    assert lib.star_sigma_loss_forward(None, None, None, None, 4, 8, 0.0, 1.0, 1.0, None, None, None, None) == 3
    assert lib.star_sigma_loss_backward(None, None, None, None, 4, 8, 0.0, 1.0, 1.0, None, None, 0, None, None) == 3
    segs = (_capi.StarAdamSeg * 1)()
    assert lib.star_adam_step(segs, 0, 0.9, 0.999, 1e-8, None, 0.0, 0, None) == 0          # nothing to do
    assert lib.star_adam_step(segs, 1, 1.5, 0.999, 1e-8, None, 0.0, 0, None) == 1          # beta1 out of range
    segs[0].n = 8
    assert lib.star_adam_step(segs, 1, 0.9, 0.999, 1e-8, None, 0.0, 0, None) == 3          # NULL run pointers
    segs[0].n = -1
    assert lib.star_adam_step(segs, 1, 0.9, 0.999, 1e-8, None, 0.0, 0, None) == 1
    assert lib.star_grad_scale(segs, 0, None, 1.0, None) == 3
    assert lib.star_grad_sqnorm(segs, 0, None, None) == 3
    # run merging: records whose every address continues the previous one's collapse into one run
    recs = [((1000, 5000), 10, "a"), ((1040, 5040), 6, "a"), ((1064, 5064), 2, "b"), ((2000, 5072), 4, "b")]
    assert O_._runs(recs) == [((1000, 5000), 16, "a"), ((1064, 5064), 2, "b"), ((2000, 5072), 4, "b")]
    # flatten_parameters: values, shapes and state_dict keys survive; each net becomes one zero-copy master vector
    net = star_b200.STaR(ref_harness.make_args(num_vehicles=1, N_importance=8))
    before = {k: v.clone() for k, v in net.state_dict().items()}
    flat = O_.flatten_parameters(net)
    assert flat.numel() == sum(p.numel() for p in net.parameters())
    after = net.state_dict()
    assert list(after.keys()) == list(before.keys())
    assert all(torch.equal(after[k], before[k]) for k in before)
    for m in (net.static_coarse_nerf, net.static_fine_nerf, net.dynamic_coarse_nerfs[0], net.dynamic_fine_nerfs[0]):
        ps = m._rt.ordered_params()
        master = F_.flat_master(ps)
        assert master.data_ptr() == ps[0].data_ptr() and master.numel() == sum(p.numel() for p in ps)
        assert master.untyped_storage().data_ptr() == flat.untyped_storage().data_ptr()
    flat.mul_(2.0)                                      # the modules see writes to the flat buffer
    assert torch.equal(net.static_coarse_nerf.rgb_linear.bias, 2.0 * before["static_coarse_nerf.rgb_linear.bias"])
    # a net whose parameters are separate tensors still gets a (copied) master vector in the same order
    other = star_b200.STaR(ref_harness.make_args(num_vehicles=0, N_importance=8))
    ps = other.static_coarse_nerf._rt.ordered_params()
    assert torch.equal(F_.flat_master(ps), torch.cat([p.detach().reshape(-1) for p in ps]))
    # no CPU fallback anywhere on this side either
    with pytest.raises(ValueError):
        O_.FusedAdam(list(other.parameters()), weight_decay=1e-2)
    opt = O_.FusedAdam(list(other.parameters()), lr=5e-4, max_grad_norm=1.0)
    for p in other.parameters():
        p.grad = torch.zeros_like(p)
    with pytest.raises(_capi.StarError):
        opt.step()
    with pytest.raises(_capi.StarError):
        L_.photometric_loss(torch.rand(4, 3), torch.rand(4, 3), torch.rand(4, 3))
    with pytest.raises(ValueError):
        L_.photometric_loss(torch.rand(4, 3), torch.rand(5, 3), torch.rand(4, 3))


def test_online_frame_scheduler_follows_the_reference_callback():
    """callbacks/online_training_callback.py:91-162: first admission when the loss reaches online_thres (threshold then
    95e-5), later ones need > 70 epochs since the last admission AND the loss under the threshold; stop beyond num_frames."""
    from star_b200.evaluation import OnlineFrameScheduler, frame_poses
    s = OnlineFrameScheduler(online_thres=1e-3, initial_num_frames=5, num_frames=7, precrop_iters=2)
    assert not s.epoch_end(0, 1e-9) and not s.epoch_end(1, 1e-9)          # precrop epochs are skipped
    assert not s.epoch_end(2, 2e-3) and s.current_frame_num == 5
    assert s.epoch_end(3, 1e-3) and s.current_frame_num == 6 and s.online_thres == 95e-5
    for e in range(70):
        assert not s.epoch_end(4 + e, 1e-9)                               # count 1..70: not yet
    assert not s.epoch_end(74, 96e-5)                                     # count 71 but loss above the new threshold
    assert s.epoch_end(75, 95e-5) and s.current_frame_num == 7 and not s.should_stop
    for e in range(70):
        s.epoch_end(76 + e, 1e-9)
    assert s.epoch_end(146, 1e-9) and s.current_frame_num == 8 and s.should_stop
    poses = torch.arange(2 * 3 * 7, dtype=torch.float32).view(2, 3, 7)
    p0 = frame_poses(poses, 0, 3, "cpu")
    assert torch.equal(p0, torch.tensor([[0, 0, 0, 0, 0, 0, 1.0]] * 3)) and torch.equal(frame_poses(poses, 2, 3, "cpu"), poses[1])

"""Times star_mlp_forward (tensor-core tier) alone on a fixed sample count; STAR_TC_DEBUG_MODE selects the
bottleneck experiments documented in csrc/mlp_tc.cu."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import star_b200
from star_b200 import _capi
from oracle import ref_harness, star_oracle as so
R, S = int(os.environ.get("R", 4096)), 128
net = star_b200.STaR(ref_harness.make_args(num_vehicles=0, N_importance=8))
net.load_state_dict(so.init_star_params(0, 8, seed=0)); net.cuda().eval(); net.set_precision(os.environ.get("PREC", "bf16"))
pts = torch.rand(R, S, 3, device="cuda") * 2 - 1
vd = torch.nn.functional.normalize(torch.randn(R, 3, device="cuda"), dim=-1)
with torch.no_grad():
    for _ in range(3): net.static_coarse_nerf.raw(pts, vd, None)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): net.static_coarse_nerf.raw(pts, vd, None)
    e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print("mode %s: %.3f ms per %d samples -> %.1f TFLOP/s" % (os.environ.get("STAR_TC_DEBUG_MODE", "0"), ms, R * S, R * S * 1.416704e6 / ms / 1e9))

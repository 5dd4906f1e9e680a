import sys, torch, time
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import star_b200
from star_b200 import functional as F_, _capi
from star_b200.models import rendering__ as R_
from oracle import ref_harness, star_oracle as so
import test_gpu_parity as T
cu=T.cu
net, params = T.make_star(1, 8, 4096, False, seed=21, training=False)
for (R,S,dyn) in [(4,32,False),(70,33,False),(16,12,True),(512,192,False)]:
    module = net.dynamic_coarse_nerfs[0] if dyn else net.static_coarse_nerf
    ro, rd = so.carla_rays(R, seed=7); vd = rd/rd.norm(dim=-1,keepdim=True)
    pts,_ = so.sample_pts(ro, rd, 0.03, 0.8, S)
    with torch.no_grad():
        module.precision=_capi.PREC_F32
        a32,c32 = module.raw(cu(pts),cu(vd),None)
        module.precision=_capi.PREC_BF16
        a16,c16 = module.raw(cu(pts),cu(vd),None)
        torch.cuda.synchronize()
    print(R,S,dyn,"alpha: max|f32| %.3f err max %.3e mean %.3e ; rgb err max %.3e mean %.3e"%(float(a32.abs().max()), float((a16-a32).abs().max()), float((a16-a32).abs().mean()), float((c16-c32).abs().max()), float((c16-c32).abs().mean())), flush=True)

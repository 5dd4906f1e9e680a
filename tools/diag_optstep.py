import sys, time, cProfile, pstats, argparse, torch
sys.path.insert(0, "/root/repo")
import bench
sys.argv = ["bench.py", "--mode", "train", "--opt-step", "--no-cpu-baseline"]
args = bench.parse()
torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
net, ro_h, rd_h, step_device, params = bench.build_workload(args, dev, 0, 1)
from star_b200 import optim as O_
O_.flatten_parameters(net)
opt = O_.FusedAdam(params, lr=5e-4, max_grad_norm=1.0)
ro, rd = ro_h.to(dev), rd_h.to(dev)
def step():
    l = step_device(ro, rd); opt.step(); return l
for _ in range(3): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5): step()
torch.cuda.synchronize()
print("ms/step", (time.perf_counter() - t0) / 5 * 1e3)
pr = cProfile.Profile(); pr.enable()
for _ in range(5): step()
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(25)

import sys, os, torch, time
sys.path.insert(0, os.getcwd())
import bench, argparse
args = bench.parse.__wrapped__() if hasattr(bench.parse, "__wrapped__") else None
sys.argv = ["bench.py", "--mode", "train", "--no-cpu-baseline"]
args = bench.parse()
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
net, ro_h, rd_h, step, params = bench.build_workload(args, dev, 0, 1)
ro, rd = ro_h.to(dev), rd_h.to(dev)
for _ in range(5): step(ro, rd)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10): step(ro, rd)
t_cpu = time.perf_counter() - t0
torch.cuda.synchronize()
t_all = time.perf_counter() - t0
print("host time per step (no sync) %.2f ms, wall per step %.2f ms" % (t_cpu * 100, t_all * 100))
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3): step(ro, rd)
    torch.cuda.synchronize()
ka = prof.key_averages()
rows = sorted([(k.key[:70], k.count, getattr(k, "device_time_total", getattr(k, "cuda_time_total", 0))) for k in ka if getattr(k, "device_time_total", getattr(k, "cuda_time_total", 0)) > 0 and k.device_type.name == "CUDA"], key=lambda r: -r[2])
tot = sum(r[2] for r in rows)
print("total device time per step %.2f ms over %d kernels/memops per step" % (tot / 3e3, sum(r[1] for r in rows) / 3))
for r in rows[:22]: print("%-72s %4d  %8.3f ms/step" % (r[0], r[1] / 3, r[2] / 3e3))

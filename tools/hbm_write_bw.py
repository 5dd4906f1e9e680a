"""Write-only / read-only / copy bandwidth of this GPU with plain torch ops (CUDA events, best of 10): the HBM figure in
MEASURED_PEAKS.json is a COPY (read + write bytes); the stash-writing kernels are write-only streams."""
import torch
dev = "cuda"
n = 2 << 30   # 2 Gi floats = 8 GiB
a = torch.empty(n, device=dev, dtype=torch.float32)
b = torch.empty(n, device=dev, dtype=torch.float32)
def best(fn, k=10):
    fn(); torch.cuda.synchronize()
    t = 1e9
    for _ in range(k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        t = min(t, e0.elapsed_time(e1))
    return t
t = best(lambda: a.fill_(1.0)); print("fill_ (write only)      %8.3f ms  %7.1f GB/s" % (t, 4 * n / t / 1e6))
t = best(lambda: a.zero_());    print("zero_ (memset)          %8.3f ms  %7.1f GB/s" % (t, 4 * n / t / 1e6))
t = best(lambda: a.sum());      print("sum (read only)         %8.3f ms  %7.1f GB/s" % (t, 4 * n / t / 1e6))
t = best(lambda: b.copy_(a));   print("copy_ (read + write)    %8.3f ms  %7.1f GB/s (both directions counted)" % (t, 8 * n / t / 1e6))
t = best(lambda: a.mul_(1.5));  print("mul_ in place (r + w)   %8.3f ms  %7.1f GB/s (both directions counted)" % (t, 8 * n / t / 1e6))

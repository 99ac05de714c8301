"""HBM bandwidth by access mix on this B200 (torch kernels): copy (read+write), write-only, read-only."""
import torch
dev = torch.device("cuda:0")
n = 1 << 28  # 2 GiB of f64
a = torch.empty(n, dtype=torch.float64, device=dev)
b = torch.empty(n, dtype=torch.float64, device=dev)
def timed(fn, iters=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3
t = timed(lambda: b.copy_(a)); print(f"copy  (read+write): {2 * n * 8 / t / 1e9:8.1f} GB/s")
t = timed(lambda: a.fill_(1.5)); print(f"write-only (fill) : {n * 8 / t / 1e9:8.1f} GB/s")
t = timed(lambda: a.sum()); print(f"read-only (sum)   : {n * 8 / t / 1e9:8.1f} GB/s")
# 80 % writes / 20 % reads, the mix of the LinearStatic path: out[5 chunks] <- f(in[1 chunk])
c = torch.empty(5, n // 8, dtype=torch.float64, device=dev); d = torch.empty(n // 8, dtype=torch.float64, device=dev)
t = timed(lambda: c.copy_(d.expand(5, -1))); print(f"1 read : 5 writes : {6 * (n // 8) * 8 / t / 1e9:8.1f} GB/s")

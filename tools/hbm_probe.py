"""HBM ceilings relevant to the kernels: copy (read+write), write-only (fill), read-only (sum)."""
import torch
n = 1 << 30
a = torch.empty(n, dtype=torch.uint8, device="cuda")
b = torch.empty(n, dtype=torch.uint8, device="cuda")
af = a.view(torch.float32)
def t(fn, bytes_moved, reps=10):
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return bytes_moved / best / 1e6
print("copy  (1 GiB read + 1 GiB write): %.0f GB/s" % t(lambda: b.copy_(a), 2 * n))
print("fill  (1 GiB write only)        : %.0f GB/s" % t(lambda: a.fill_(1), n))
print("zero  (cudaMemset, write only)  : %.0f GB/s" % t(lambda: a.zero_(), n))
print("sum   (1 GiB read only)         : %.0f GB/s" % t(lambda: af.sum(), n))

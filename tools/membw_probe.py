"""HBM write-only / read-only / copy bandwidth on this GPU (context for the store-bound first layer: is 3.0 TB/s of pure writes 46 % of
anything attainable?).  torch fill_ / sum / copy_ on 4 GiB fp16 tensors, CUDA events, best of 5."""
import torch
n = 2 * 1024 ** 3
a = torch.empty(n, dtype=torch.float16, device='cuda'); b = torch.empty_like(a)
def best(fn, bytes_moved):
    ts = []
    for _ in range(6):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    t = min(ts[1:])
    return bytes_moved / t / 1e6
print('write-only (fill_)  GB/s', round(best(lambda: a.fill_(1.0), 2 * n)))
print('write-only (memset) GB/s', round(best(lambda: a.zero_(), 2 * n)))
print('read-only (sum)     GB/s', round(best(lambda: a.view(torch.float32).sum(), 2 * n)))
print('copy (read+write)   GB/s', round(best(lambda: b.copy_(a), 4 * n)))

#!/usr/bin/env python
"""HBM bandwidth of the three access mixes the GEMM epilogues produce: pure write, pure read, copy."""
import torch
x = torch.empty(1 << 29, dtype=torch.float32, device='cuda')     # 2 GiB
y = torch.empty_like(x)


def timed(fn, reps=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


gb = x.numel() * 4 / 1e9
print('write  (fill_)  : %.0f GB/s' % (gb / timed(lambda: x.fill_(1.0)) * 1e3))
print('read   (sum)    : %.0f GB/s' % (gb / timed(lambda: x.sum()) * 1e3))
print('copy   (copy_)  : %.0f GB/s (read + write)' % (2 * gb / timed(lambda: y.copy_(x)) * 1e3))
print('1r:3w  (3 fills + sum): %.0f GB/s' % (4 * gb / timed(lambda: (x.sum(), x.fill_(1.0), y.fill_(2.0), x.fill_(3.0))) * 1e3))

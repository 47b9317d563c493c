#!/usr/bin/env python
"""Read-stream bandwidth of the device (torch.sum over a 16 GB fp64 tensor) beside the copy figure of MEASURED_PEAKS.json:
the SpMV is a pure read stream, so this is the roof it can reach."""
import torch
x = torch.ones(2 * 1024 ** 3, dtype=torch.float64, device='cuda')
for _ in range(2): x.sum()
torch.cuda.synchronize()
best = 0
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); s = x.sum(); e1.record(); e1.synchronize()
    best = max(best, x.numel() * 8 / (e0.elapsed_time(e1) * 1e-3) / 1e9)
y = torch.empty_like(x)
bc = 0
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); y.copy_(x); e1.record(); e1.synchronize()
    bc = max(bc, 2 * x.numel() * 8 / (e0.elapsed_time(e1) * 1e-3) / 1e9)
print('read stream (torch.sum, 16 GiB): %.0f GB/s ; copy (read+write, 2 x 16 GiB): %.0f GB/s' % (best, bc))

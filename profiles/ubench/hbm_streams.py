#!/usr/bin/env python
"""Ceilings for write-only, read-only and copy streams on this GPU (torch library kernels, 1.5 GB buffers):
the denominators that the HBM-bound legs (rfgr2beff: write-only; its adjoint: read-only) can actually reach."""
import json
import torch

dev = torch.device('cuda:0')
n = 1572864000 // 4
a = torch.empty(n, device=dev)
b = torch.empty(n, device=dev)
out = {}


def timed(fn, byt, reps=5):
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return {'ms': best, 'GB/s': byt / (best * 1e-3) / 1e9}


out['fill (write only)'] = timed(lambda: a.fill_(1.0), n * 4)
out['sum (read only)'] = timed(lambda: a.sum(), n * 4)
out['copy (read + write)'] = timed(lambda: b.copy_(a), 2 * n * 4)
out['cudaMemsetAsync'] = timed(lambda: a.zero_(), n * 4)
print(json.dumps(out))

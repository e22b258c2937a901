#!/usr/bin/env python
"""Throughput of the multi-coil (pTx) code paths: nCoils in {1,2,4,8,16} with a b1Map, 64^3 spins x 1000 steps
(python profiles/bench_multicoil.py [n] [nT]: n^3 spins x nT steps)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'mrphy.py_b200'))
import torch
from mrphy import _ops, _cabi
dev = torch.device('cuda:0'); dt = torch.float32
g = torch.Generator(device='cuda').manual_seed(0)
U = lambda *s: torch.rand(s, generator=g, device=dev, dtype=dt) * 2 - 1
nM, nT = (int(sys.argv[1]) if len(sys.argv) > 1 else 64) ** 3, (int(sys.argv[2]) if len(sys.argv) > 2 else 1000)
L = _cabi.lib(); L.mrphy_kernel_timing(1)
out = {}
for nC in (1, 2, 4, 8, 16):
    rf = (U(1, 2, nT, nC) * 0.1 / nC).requires_grad_(True); gr = (U(1, 3, nT) * 2).requires_grad_(True)
    b1 = U(1, nM, 2, nC); loc = U(1, nM, 3) * 12; df = U(1, nM) * 200
    M0 = torch.tensor([0., 0., 1.], device=dev).expand(1, nM, 3).contiguous()
    f, b = [], []
    for _ in range(4):
        rf.grad = gr.grad = None
        Mo = _ops.fused_applypulse(M0, rf, gr, loc, Δf_=df, b1Map_=b1, T1_=torch.tensor(1.47, device=dev),
                                   T2_=torch.tensor(0.07, device=dev), γ_=torch.tensor(4257.6, device=dev),
                                   dt=torch.tensor(4e-6, device=dev))
        f.append(L.mrphy_last_kernel_ms())
        Mo.sum().backward()
        b.append(L.mrphy_last_kernel_ms())
    out[nC] = {'fwd_ms': min(f), 'bwd_ms': min(b), 'spin_steps_per_s': nM * nT / ((min(f) + min(b)) * 1e-3)}
print(json.dumps(out))

#!/usr/bin/env python
"""Backward kernel time at C2 when only some gradients are wanted (rows of the spin reduction on demand)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'mrphy.py_b200'))
import numpy as np, torch, bench
from mrphy import mobjs, _cabi
dev = torch.device('cuda:0'); kw = {'dtype': torch.float32, 'device': dev}
N, n, nT = bench.WORKLOADS['c2']
d = {k: v.to(dev) for k, v in bench.synth(N, n, n, nT, torch.float32).items()}
sp = mobjs.SpinArray((N, d['loc'].shape[1]), M_=d['M0'], **kw)
tgt = torch.tensor([0., 1., 0.], **kw)
L = _cabi.lib(); L.mrphy_kernel_timing(1)
for which in ('rf+gr', 'rf only', 'gr only'):
    pulse = mobjs.Pulse(rf=d['rf'].clone().requires_grad_('rf' in which), gr=d['gr'].clone().requires_grad_('gr' in which), **kw)
    tb = []
    for i in range(8):
        pulse.rf.grad = pulse.gr.grad = None
        M = sp.applypulse(pulse, loc_=d['loc'], Δf_=d['df'], b1Map_=d['b1'])
        ((M - tgt) ** 2).sum().backward()
        if i >= 3:
            tb.append(L.mrphy_last_kernel_ms())
    print(f'{which:8s} backward kernel {np.median(tb):.4f} ms')

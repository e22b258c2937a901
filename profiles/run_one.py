#!/usr/bin/env python
"""Short driver for ncu: a few fwd+bwd steps of the bench workload through the public API.

    python profiles/run_one.py [--workload c2] [--steps 2] [--dtype f32] [--coils 8]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'mrphy.py_b200'))
import torch  # noqa: E402
import bench  # noqa: E402
from mrphy import mobjs  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--workload', default='c2')
ap.add_argument('--steps', type=int, default=2)
ap.add_argument('--dtype', default='f32')
ap.add_argument('--coils', type=int, default=1, help='> 1: parallel transmit, rf (N,2,nT,nC) and a b1Map of nC coils')
a = ap.parse_args()
N, n, nT = bench.WORKLOADS[a.workload]
dtype = torch.float32 if a.dtype == 'f32' else torch.float64
dev = torch.device('cuda:0')
kw = {'dtype': dtype, 'device': dev}
d = {k: v.to(dev) for k, v in bench.synth(N, n, n, nT, dtype).items()}
if a.coils > 1:
    g = torch.Generator(device=dev).manual_seed(1)
    d['rf'] = ((torch.rand((N, 2, nT, a.coils), generator=g, **kw) * 2 - 1) * 0.1 / a.coils)
    d['b1'] = torch.rand((N, d['loc'].shape[1], 2, a.coils), generator=g, **kw) * 2 - 1
sp = mobjs.SpinArray((N, d['loc'].shape[1]), M_=d['M0'], **kw)
pulse = mobjs.Pulse(rf=d['rf'].requires_grad_(True), gr=d['gr'].requires_grad_(True), **kw)
tgt = torch.tensor([0., 1., 0.], **kw)
for _ in range(a.steps):
    pulse.rf.grad = pulse.gr.grad = None
    M = sp.applypulse(pulse, loc_=d['loc'], Δf_=d['df'], b1Map_=d['b1'])
    ((M - tgt) ** 2).sum().backward()
torch.cuda.synchronize()
print('ok', float(pulse.rf.grad.abs().sum()))

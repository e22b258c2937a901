#!/usr/bin/env python
"""Where the HOST spends a small design iteration (16^3 x 256, fused re-parametrisation): cProfile, top functions."""
import cProfile
import os
import pstats
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'mrphy.py_b200'))
import torch
import bench
from mrphy import mobjs, utils, rfmax0, smax0, dt0

dev = torch.device('cuda:0'); kw = {'dtype': torch.float32, 'device': dev}
rfmax0, smax0, dt0 = rfmax0.to(**kw), smax0.to(**kw), dt0.to(**kw)
n, nT = 16, 256
d = {k: t.to(dev) for k, t in bench.synth(1, n, n, nT, torch.float32).items()}
sp = mobjs.SpinArray((1, d['loc'].shape[1]), M_=d['M0'], **kw)
tgt = torch.tensor([0., 1., 0.], **kw)
v = [torch.randn((1, c, nT), **kw).mul_(0.3).requires_grad_(True) for c in (1, 1, 3)]


def it():
    rf, gr = utils.tρθts2rfgr(v[0], v[1], v[2], rfmax0, smax0, dt0)
    pulse = mobjs.Pulse(rf=rf, gr=gr, **kw)
    M = sp.applypulse(pulse, loc_=d['loc'], Δf_=d['df'], b1Map_=d['b1'])
    loss = ((M - tgt) ** 2).sum()
    for x in v:
        x.grad = None
    loss.backward()


for _ in range(20):
    it()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(200):
    it()
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats('cumulative').print_stats(28)

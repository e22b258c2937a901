#!/usr/bin/env python
"""Where the end-to-end step (pinned host inputs -> device -> fwd+bwd -> loss/grads back) spends its time."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'mrphy.py_b200'))
import torch, bench
from mrphy import mobjs
N, n, nT = bench.WORKLOADS['c2']
dev = torch.device('cuda:0'); dtype = torch.float32; kw = {'dtype': dtype, 'device': dev}
host = bench.synth(N, n, n, nT, dtype); pinned = {k: v.pin_memory() for k, v in host.items()}
nM = host['loc'].shape[1]; tgt = torch.tensor([0., 1., 0.], **kw)
def sync(): torch.cuda.synchronize(); return time.perf_counter()
acc = {}
for it in range(12):
    t0 = sync()
    d = {k: v.to(dev, non_blocking=True) for k, v in pinned.items()}
    t1 = sync()
    sp = mobjs.SpinArray((N, nM), M_=d['M0'], **kw)
    pulse = mobjs.Pulse(rf=d['rf'].requires_grad_(True), gr=d['gr'].requires_grad_(True), **kw)
    t2 = sync()
    M = sp.applypulse(pulse, loc_=d['loc'], Δf_=d['df'], b1Map_=d['b1'])
    t3 = sync()
    loss = ((M - tgt) ** 2).sum(); loss.backward()
    t4 = sync()
    out = (loss.item(), pulse.rf.grad.cpu(), pulse.gr.grad.cpu())
    t5 = sync()
    if it >= 2:
        for k, v in (('h2d', t1 - t0), ('objects', t2 - t1), ('fwd', t3 - t2), ('loss+bwd', t4 - t3), ('d2h', t5 - t4), ('total', t5 - t0)):
            acc.setdefault(k, []).append(v * 1e3)
print({k: round(sum(v) / len(v), 3) for k, v in acc.items()})
# host-side cost of the calls alone (no sync inside): how far ahead of the GPU does the CPU run?
t0 = sync()
for _ in range(10):
    pulse.rf.grad = pulse.gr.grad = None
    M = sp.applypulse(pulse, loc_=d['loc'], Δf_=d['df'], b1Map_=d['b1'])
    ((M - tgt) ** 2).sum().backward()
t1 = time.perf_counter(); t2 = sync()
print('host issue time per step %.3f ms, device-complete per step %.3f ms' % ((t1 - t0) * 100, (t2 - t0) * 100))

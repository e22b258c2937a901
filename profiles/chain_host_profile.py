#!/usr/bin/env python
"""Host issue time vs device time of the explicit chain rf,gr -> rfgr2beff -> sims.blochsim -> backward."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'mrphy.py_b200'))
import torch
from mrphy import sims, beffective
dev = torch.device('cuda:0'); dtype = torch.float32
g = torch.Generator(device='cuda').manual_seed(0)
U = lambda *s: torch.rand(s, generator=g, device=dev, dtype=dtype) * 2 - 1
nM, nT = 131072, 1000
rf, gr, loc, df = U(1, 2, nT) * .1, U(1, 3, nT) * 2, U(1, nM, 3) * 12, U(1, nM) * 200
b1 = U(1, nM, 2) * .1
Mi = torch.nn.functional.normalize(U(1, nM, 3), dim=-1)
rfg, grg = rf.clone().requires_grad_(True), gr.clone().requires_grad_(True)
T1, T2 = torch.tensor(1.47, device=dev), torch.tensor(0.07, device=dev)


def once():
    rfg.grad = grg.grad = None
    beff = beffective.rfgr2beff(rfg, grg, loc, Δf=df, b1Map=b1)
    Mo = sims.blochsim(Mi, beff, T1=T1, T2=T2)
    Mo.sum().backward()


for _ in range(3):
    once()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    once()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print('host issue %.3f ms/iter, device-complete %.3f ms/iter' % ((t1 - t0) * 100, (t2 - t0) * 100))
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        once()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by='self_cpu_time_total', row_limit=22, max_name_column_width=48))

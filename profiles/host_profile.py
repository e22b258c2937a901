import os, sys, time, cProfile, pstats
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'mrphy.py_b200'))
import torch, bench
from mrphy import mobjs
N, n, nT = bench.WORKLOADS['small']
dev = torch.device('cuda:0'); dtype = torch.float32; kw = {'dtype': dtype, 'device': dev}
d = {k: v.to(dev) for k, v in bench.synth(N, n, n, nT, dtype).items()}
sp = mobjs.SpinArray((N, d['loc'].shape[1]), M_=d['M0'], **kw)
pulse = mobjs.Pulse(rf=d['rf'].requires_grad_(True), gr=d['gr'].requires_grad_(True), **kw)
tgt = torch.tensor([0., 1., 0.], **kw)
def step():
    pulse.rf.grad = pulse.gr.grad = None
    M = sp.applypulse(pulse, loc_=d['loc'], Δf_=d['df'], b1Map_=d['b1'])
    ((M - tgt) ** 2).sum().backward()
for _ in range(20): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(200): M = sp.applypulse(pulse, loc_=d['loc'], Δf_=d['df'], b1Map_=d['b1'])
torch.cuda.synchronize(); t1 = time.perf_counter()
print('applypulse host+device (tiny problem): %.1f us/call' % ((t1 - t0) / 200 * 1e6))
pr = cProfile.Profile(); pr.enable()
for _ in range(200): step()
torch.cuda.synchronize(); pr.disable()
st = pstats.Stats(pr); st.sort_stats('cumulative').print_stats(28)

#!/usr/bin/env python
"""Per-CTA timeline of the backward kernel (measurement build: python mrphy.py_b200/build.py -DMRPHY_CTA_TRACE
--out=profiles/variants/trace.so; run with MRPHY_B200_LIB pointing at it).  Prints how the CTAs were spread over the
SMs, when they started and how long they ran -- the explanation for achieved vs theoretical occupancy."""
import ctypes, os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'mrphy.py_b200'))
import numpy as np
import torch
import bench
from mrphy import mobjs, _cabi

dev = torch.device('cuda:0'); kw = {'dtype': torch.float32, 'device': dev}
N, n, nT = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else 'c2']
d = {k: v.to(dev) for k, v in bench.synth(N, n, n, nT, torch.float32).items()}
sp = mobjs.SpinArray((N, d['loc'].shape[1]), M_=d['M0'], **kw)
pulse = mobjs.Pulse(rf=d['rf'].requires_grad_(True), gr=d['gr'].requires_grad_(True), **kw)
tgt = torch.tensor([0., 1., 0.], **kw)
for _ in range(3):
    pulse.rf.grad = pulse.gr.grad = None
    M = sp.applypulse(pulse, loc_=d['loc'], Δf_=d['df'], b1Map_=d['b1'])
    ((M - tgt) ** 2).sum().backward()
torch.cuda.synchronize()
L = ctypes.CDLL(_cabi.LIB_PATH)
P = 8192
buf = (ctypes.c_ulonglong * (4 * P))()
assert L.mrphy_debug_cta_trace(buf, P) == 0
a = np.frombuffer(buf, dtype=np.uint64).reshape(P, 4).astype(np.int64)
a = a[a[:, 2] > 0]
t0 = a[:, 1].min()
start, end = (a[:, 1] - t0) / 1e3, (a[:, 2] - t0) / 1e3
per_sm = collections.Counter(a[:, 0].tolist())
print('CTAs traced', len(a), 'SMs used', len(per_sm), 'CTAs per SM histogram', sorted(collections.Counter(per_sm.values()).items()))
print('kernel span %.1f us; CTA start: p50 %.1f p90 %.1f max %.1f us' % (end.max(), np.percentile(start, 50), np.percentile(start, 90), start.max()))
dur = end - start
print('CTA duration: min %.1f p10 %.1f p50 %.1f p90 %.1f max %.1f us' % (dur.min(), np.percentile(dur, 10), np.percentile(dur, 50), np.percentile(dur, 90), dur.max()))
print('CTA end: p10 %.1f p50 %.1f p90 %.1f max %.1f us' % tuple(np.percentile(end, [10, 50, 90, 100])))
print('mean resident fraction of the kernel span: %.3f' % (dur.sum() / (len(a) * end.max())))
by = collections.defaultdict(list)
for smi, du in zip(a[:, 0].tolist(), dur.tolist()):
    by[per_sm[smi]].append(du)
print('mean CTA duration by CTAs-per-SM:', {k: round(float(np.mean(v)), 1) for k, v in sorted(by.items())})
# tiles per SM under the static assignment (CTA b takes tiles b, b+P, ...): the per-SM load the kernel time follows
blkt = int(os.environ.get('TRACE_BLKT', '64'))
tiles = (d['loc'].shape[1] + 2 * blkt - 1) // (2 * blkt)
P_ = len(a)
ids = np.nonzero(np.frombuffer(buf, dtype=np.uint64).reshape(P, 4)[:, 2] > 0)[0]
vids = a[:, 3]   # virtual id served first (== blockIdx without the SM-aware ownership)
per_cta = np.array([(tiles - b + P_ - 1) // P_ if 0 <= b < tiles else 0 for b in vids])
load = collections.Counter()
for smi, t in zip(a[:, 0].tolist(), per_cta.tolist()):
    load[smi] += t
print('tiles', tiles, 'CTAs', P_, 'tiles per SM histogram', sorted(collections.Counter(load.values()).items()))
order = np.argsort(ids)
print('SM of the first 16 CTAs', a[order[:16], 0].tolist(), '... CTAs 148..156', a[order[148:156], 0].tolist())

#!/usr/bin/env python
"""HBM-bound legs of the path: rfgr2beff (12 B/spin.step written), sims.blochsim with an explicit Beff forward
(12 B read) and backward (12 B read + 12 B written), fp32.  Prints achieved GB/s against MEASURED_PEAKS.json.

    python profiles/bench_explicit.py [--nM 131072] [--nT 1000] [--dtype f32]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'mrphy.py_b200'))
import torch  # noqa: E402
from mrphy import sims, beffective, _cabi  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--nM', type=int, default=131072)
ap.add_argument('--nT', type=int, default=1000)
ap.add_argument('--dtype', default='f32')
ap.add_argument('--reps', type=int, default=5)
a = ap.parse_args()
dtype = torch.float32 if a.dtype == 'f32' else torch.float64
esz = 4 if a.dtype == 'f32' else 8
dev = torch.device('cuda:0')
g = torch.Generator(device='cuda').manual_seed(0)
U = lambda *s: torch.rand(s, generator=g, device=dev, dtype=dtype) * 2 - 1
rf, gr, loc, df = U(1, 2, a.nT) * .1, U(1, 3, a.nT) * 2, U(1, a.nM, 3) * 12, U(1, a.nM) * 200
b1 = U(1, a.nM, 2) * .1
Mi = torch.nn.functional.normalize(U(1, a.nM, 3), dim=-1).requires_grad_(True)
L = _cabi.lib()
L.mrphy_kernel_timing(1)
try:
    hbm = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs']
except Exception:
    hbm = 6650.0
res = {'rfgr2beff': [], 'beff_fwd': [], 'beff_bwd': []}
for _ in range(a.reps):
    beff = beffective.rfgr2beff(rf, gr, loc, Δf=df, b1Map=b1)
    res['rfgr2beff'].append(L.mrphy_last_kernel_ms())
    beff.requires_grad_(True)
    Mo = sims.blochsim(Mi, beff, T1=torch.tensor(1.47, device=dev), T2=torch.tensor(0.07, device=dev))
    res['beff_fwd'].append(L.mrphy_last_kernel_ms())
    Mo.sum().backward()
    res['beff_bwd'].append(L.mrphy_last_kernel_ms())
    Mi.grad = None
    del beff, Mo
# the whole API-faithful chain  rf,gr -> rfgr2beff -> sims.blochsim -> loss -> rf.grad, gr.grad  (4 kernels + finalize)
rfg, grg = rf.clone().requires_grad_(True), gr.clone().requires_grad_(True)
T1c, T2c, Mic = torch.tensor(1.47, device=dev), torch.tensor(0.07, device=dev), Mi.detach()
chain = []
for _ in range(a.reps):
    rfg.grad = grg.grad = None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    beff = beffective.rfgr2beff(rfg, grg, loc, Δf=df, b1Map=b1)
    Mo = sims.blochsim(Mic, beff, T1=T1c, T2=T2c)
    Mo.sum().backward()
    e1.record()
    torch.cuda.synchronize()
    res.setdefault('rfgr2beff_bwd', []).append(L.mrphy_last_kernel_ms())
    chain.append(e0.elapsed_time(e1))
    del beff, Mo
units = a.nM * a.nT
byt = {'rfgr2beff': 3 * esz, 'beff_fwd': 3 * esz, 'beff_bwd': 6 * esz, 'rfgr2beff_bwd': 3 * esz}
out = {}
for k, v in res.items():
    ms = min(v)
    out[k] = {'ms': ms, 'GB/s': units * byt[k] / (ms * 1e-3) / 1e9, 'frac_of_measured_hbm': units * byt[k] / (ms * 1e-3) / 1e9 / hbm,
              'spin_steps_per_s': units / (ms * 1e-3)}
out['chain_fwd_bwd'] = {'ms': min(chain), 'spin_steps_per_s': units / (min(chain) * 1e-3)}
print(json.dumps({'nM': a.nM, 'nT': a.nT, 'dtype': a.dtype, 'hbm_peak_gbs': hbm, **out}))

"""Small run of every kernel family for compute-sanitizer memcheck (ragged sizes, all code paths)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'mrphy.py_b200'))
import torch
from mrphy import _ops, sims, beffective, mobjs
dev = torch.device('cuda:0')
g = torch.Generator().manual_seed(0)
U = lambda *s, dt=torch.float32: (torch.rand(s, generator=g, dtype=torch.float64) * 2 - 1).to(dt).to(dev)
for dt in (torch.float32, torch.float64):
    for pack in ('1', '2', '3'):
        os.environ['MRPHY_B200_PACK'] = pack   # only honoured by -DMRPHY_FP32_SCALAR builds
        for (N, nM, nT, nC) in ((2, 131, 70, 1), (1, 65, 33, 3), (1, 200, 129, 0)):
            rf = U(N, 2, nT, nC, dt=dt) if nC else U(N, 2, nT, dt=dt)
            rf.requires_grad_(True)
            gr = U(N, 3, nT, dt=dt).requires_grad_(True)
            M0 = U(N, nM, 3, dt=dt).requires_grad_(True)
            b1 = U(N, nM, 2, nC, dt=dt) if nC else None
            Mo = _ops.fused_applypulse(M0, rf, gr, U(N, nM, 3, dt=dt) * 5, Δf_=U(N, nM, dt=dt) * 100, b1Map_=b1,
                                       T1_=U(N, nM, dt=dt).abs() + 1, T2_=U(N, nM, dt=dt).abs() * 0.05 + 0.02,
                                       γ_=torch.tensor(4257.6, device=dev), dt=torch.tensor(4e-6, device=dev), ckpt=16)
            Mo.sum().backward()
    for nT in (40, 37):
        beff = beffective.rfgr2beff(U(1, 2, nT, dt=dt), U(1, 3, nT, dt=dt), U(1, 77, 3, dt=dt), Δf=U(1, 77, dt=dt),
                                    b1Map=U(1, 77, 2, dt=dt)).requires_grad_(True)
        Mi = U(1, 77, 3, dt=dt).requires_grad_(True)
        sims.blochsim(Mi, beff, T1=torch.tensor(1., device=dev), T2=torch.tensor(.05, device=dev)).sum().backward()
        A, B = beffective.beff2ab(beff.detach(), E1=torch.tensor(.99, device=dev), E2=torch.tensor(.9, device=dev))
    sims.freeprec(U(2, 50, 3, dt=dt).requires_grad_(True), torch.tensor(0.01, device=dev), T1=torch.tensor(1., device=dev),
                  T2=torch.tensor(.05, device=dev), Δf=U(2, 50, dt=dt)).sum().backward()
torch.cuda.synchronize()
print('sanitize_small ok')

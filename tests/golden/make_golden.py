#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ from the UNMODIFIED reference.

Run in the authoring container only (the reference does not travel to the GPU box):

    python tests/golden/make_golden.py

It imports MRphy.py from /root/reference (read-only), runs the reference's own
``sims.blochsim`` / ``slowsims.blochsim`` / ``beffective.*`` / ``mobjs.*`` on seeded
inputs and stores inputs + outputs as small ``.npz`` files.  The tests never import
the reference; they only read these files.

Cases (see SURVEY.md section 8c for the reference tests they mirror):
  kat3      tests/test_slowsims.py:27-98   3 spins, nT=512 (hard-coded Mo0 constants too)
  sims512   tests/test_sims.py:24-143      nM=512 random M0, grads wrt M0 and Beff
  cube27    tests/test_mobjs.py:98-131     masked SpinCube.applypulse, relax / no relax
  interp    tests/test_mobjs.py:160-195    Pulse.interpT (+ the float // length quirk)
  rand_mc   random N=2, nM=7, nT=50, nCoils=2, per-spin gamma/T1/T2 (SURVEY App. A probe)
  rand_nob1 random N=2, nM=5, nT=40, 3 coils summed, no b1Map, no df, per-batch dt
  bench8    8^3 cube, BASELINE.md workload distributions, nT=1000, fp32 and fp64
  freeprec  tests/test_slowsims.py:100-122 + random case
  reparam   utils.tρθ2rf / lρθ2rf / ts2s / s2g with autograd gradients; SpinArray.embed / extract on a random mask
  sequence  SpinCube applypulse -> freeprec -> applypulse chained through doUpdate, gradients w.r.t. both pulses
  standalone beffective.beff2uϕ, rfgr2beff (every gradient: rf, gr, loc, Δf, b1Map, γ; broadcast b1Map cases), beff2ab with
            gradients, utils.rfclamp / sclamp with gradients -- what tests/torch_ref.py is pinned to
"""
import os
import sys

import numpy as np

REF = '/root/reference'
sys.path.insert(0, REF)
import torch  # noqa: E402
import mrphy  # noqa: E402
from mrphy import beffective, mobjs, sims, slowsims  # noqa: E402

assert mrphy.__file__.startswith(REF), mrphy.__file__
HERE = os.path.dirname(os.path.abspath(__file__))
PI = np.pi
f64, f32 = torch.float64, torch.float32
gH = mrphy.γH


def npy(x):
    return None if x is None else x.detach().cpu().numpy()


def save(name, **kw):
    kw = {k: v for k, v in kw.items() if v is not None}
    path = os.path.join(HERE, name + '.npz')
    np.savez_compressed(path, **kw)
    print(f'{name}: {os.path.getsize(path)/1024:.1f} KiB', {k: np.shape(v) for k, v in kw.items()})


def ref_chain(M0, rf, gr, loc, df, b1, T1, T2, gam, dt, w):
    """Reference chain rfgr2beff -> sims.blochsim; returns Mo, beff and grads of sum(w*Mo)."""
    M0 = M0.clone().requires_grad_(True)
    rf = rf.clone().requires_grad_(True)
    gr = gr.clone().requires_grad_(True)
    beff = beffective.rfgr2beff(rf, gr, loc, Δf=df, b1Map=b1, γ=gam)
    Mo = sims.blochsim(M0, beff, T1=T1, T2=T2, γ=gam, dt=dt)
    (Mo * w).sum().backward()
    return Mo.detach(), beff.detach(), rf.grad, gr.grad, M0.grad


def slow_chain(M0, rf, gr, loc, df, b1, T1, T2, gam, dt, w):
    """Same through slowsims (plain autograd): gives grad_M0 where sims.py:267 breaks."""
    M0 = M0.clone().requires_grad_(True)
    rf = rf.clone().requires_grad_(True)
    gr = gr.clone().requires_grad_(True)
    beff = beffective.rfgr2beff(rf, gr, loc, Δf=df, b1Map=b1, γ=gam)
    # slowsims.py:87-88 broadcasts dt against T1 before reshaping, so a per-batch dt must be (N,1)
    dts = dt.reshape(-1, 1) if dt.numel() > 1 else dt
    Mo = slowsims.blochsim(M0, beff, T1=T1, T2=T2, γ=gam, dt=dts)
    (Mo * w).sum().backward()
    return Mo.detach(), rf.grad, gr.grad, M0.grad


# --------------------------------------------------------------------------- kat3
def test_pulse(nT, dkw, ones_y=False):
    t = torch.arange(0, nT, **dkw).reshape((1, 1, nT))
    rf = 10 * torch.cat([torch.cos(t / nT * 2 * PI), torch.sin(t / nT * 2 * PI)], 1)
    gy = torch.ones((1, 1, nT), **dkw) if ones_y else torch.zeros((1, 1, nT), **dkw)
    gr = torch.cat([torch.ones((1, 1, nT), **dkw), gy,
                    10 * torch.atan(t - round(nT / 2)) / PI], 1)
    return rf, gr


def case_kat3():
    dkw = {'dtype': f64}
    gam, dt = gH.to(**dkw), mrphy.dt0.to(**dkw)
    M0 = torch.eye(3, **dkw)[None]
    nM, nT = 3, 512
    T1, T2 = torch.tensor([[1.]], **dkw), torch.tensor([[4e-2]], **dkw)
    lx = torch.linspace(-1., 1., nM, **dkw).reshape(1, nM)
    loc = torch.stack([lx, lx, torch.ones_like(lx)], 2)
    df = -lx * gam
    b1 = torch.tensor([1., 0.], **dkw).reshape(1, 1, 2, 1)
    rf, gr = test_pulse(nT, dkw)
    rf = rf[..., None]
    w = torch.ones(1, nM, 3, **dkw)
    Mo, beff, grf, ggr, gM0 = slow_chain(M0, rf, gr, loc, df, b1, T1, T2, gam, dt, w)[0], None, None, None, None
    Mo_s, grf_s, ggr_s, gM0_s = slow_chain(M0, rf, gr, loc, df, b1, T1, T2, gam, dt, w)
    Mo_e, beff, grf_e, ggr_e, gM0_e = ref_chain(M0, rf, gr, loc, df, b1, T1, T2, gam, dt, w)
    E1, E2 = torch.exp(-dt / T1), torch.exp(-dt / T2)
    A, B = beffective.beff2ab(beff, E1=E1, E2=E2, γ=gam, dt=dt)
    Mo0 = np.array([[[0.559535641648385, 0.663342640621335, 0.416341441715101],
                     [0.391994737048090, 0.210182892388552, -0.860954821972489],
                     [-0.677062008711222, 0.673391604920576, -0.143262993311057]]])
    assert np.abs(npy(Mo_s) - Mo0).max() < 1e-9 and np.abs(npy(Mo_e) - Mo0).max() < 1e-9
    save('kat3', M0=npy(M0), rf=npy(rf), gr=npy(gr), loc=npy(loc), df=npy(df), b1=npy(b1),
         T1=npy(T1), T2=npy(T2), gamma=npy(gam), dt=npy(dt), Mo_const=Mo0,
         Mo_sims=npy(Mo_e), Mo_slow=npy(Mo_s), grf=npy(grf_e), ggr=npy(ggr_e), gM0=npy(gM0_e),
         grf_slow=npy(grf_s), ggr_slow=npy(ggr_s), gM0_slow=npy(gM0_s),
         A=npy(A), B=npy(B), beff_first=npy(beff[:, :, :4]), beff_last=npy(beff[:, :, -4:]))


# ------------------------------------------------------------------------ sims512
def case_sims512():
    dkw = {'dtype': f64}
    gam, dt = gH.to(**dkw), mrphy.dt0.to(**dkw)
    g = torch.Generator().manual_seed(1234)
    nM, nT = 512, 512
    M0 = torch.rand((1, nM, 3), generator=g, **dkw)
    T1, T2 = torch.tensor([[1.]], **dkw), torch.tensor([[4e-2]], **dkw)
    lx = torch.linspace(-1., 1., nM, **dkw).reshape(1, nM)
    loc = torch.stack([lx, lx, torch.ones_like(lx)], 2)
    df = -lx * gam
    b1 = torch.tensor([1., 0.], **dkw).reshape(1, 1, 2, 1)
    rf, gr = test_pulse(nT, dkw)
    rf = rf[..., None]
    beff = beffective.rfgr2beff(rf, gr, loc, Δf=df, b1Map=b1, γ=gam)
    beff_nodim = beffective.rfgr2beff(rf[..., 0], gr, loc, Δf=df, b1Map=b1[..., 0], γ=gam)
    assert torch.equal(beff, beff_nodim)
    out = {}
    sub = slice(None, None, 32)   # keep grad_Beff for 16 of the 512 spins
    for tag, (t1, t2) in {'relax': (T1, T2), 'norelax': (None, None)}.items():
        M0r = M0.clone().requires_grad_(True)
        br = beff.clone().requires_grad_(True)
        Mo = sims.blochsim(M0r, br, T1=t1, T2=t2, γ=gam, dt=dt)
        Mo.sum().backward()
        M0s = M0.clone().requires_grad_(True)
        bs = beff.clone().requires_grad_(True)
        Mos = slowsims.blochsim(M0s, bs, T1=t1, T2=t2, γ=gam, dt=dt)
        Mos.sum().backward()
        assert (M0r.grad - M0s.grad).abs().max() < 1e-9
        assert (br.grad - bs.grad).abs().max() < 1e-9
        out[f'Mo_{tag}'] = npy(Mo)
        out[f'gM0_{tag}'] = npy(M0r.grad)
        out[f'gBeff_sub_{tag}'] = npy(br.grad[:, sub])
        # gradient wrt the waveform through the reference chain (what the fused path returns)
        w = torch.ones_like(M0)
        _, _, grf, ggr, _ = ref_chain(M0, rf, gr, loc, df, b1, t1, t2, gam, dt, w)
        out[f'grf_{tag}'] = npy(grf)
        out[f'ggr_{tag}'] = npy(ggr)
    save('sims512', M0=npy(M0), rf=npy(rf), gr=npy(gr), loc=npy(loc), df=npy(df), b1=npy(b1),
         T1=npy(T1), T2=npy(T2), gamma=npy(gam), dt=npy(dt), sub_step=np.array(32), **out)


# ------------------------------------------------------------------------- cube27
def case_cube27():
    from copy import deepcopy
    dkw = {'dtype': f64, 'device': torch.device('cpu')}
    gam = gH.to(**dkw)
    N, Nd, nT = 1, (3, 3, 3), 512
    rf, gr = test_pulse(nT, dkw, ones_y=True)
    p = mobjs.Pulse(rf=rf, gr=gr, dt=mrphy.dt0, **dkw)
    mask = torch.zeros((1,) + Nd, dtype=torch.bool)
    mask[0, :, 1, :], mask[0, 1, :, :] = True, True
    fov, ofst = torch.tensor([[3., 3., 3.]], **dkw), torch.tensor([[0., 0., 1.]], **dkw)
    cube = mobjs.SpinCube((N,) + Nd, fov, mask=mask, T1_=torch.tensor([[1.]], **dkw), γ=gam, **dkw)
    cube.ofst = ofst
    cube.M_ = torch.tensor([0., 1., 0.])
    cube.T2 = torch.tensor([[4e-2]], **dkw).expand(cube.shape)
    M001, M100 = torch.tensor([0., 0., 1.], **dkw), torch.tensor([1., 0., 0.], **dkw)
    sl = mrphy._slice
    cube.M_[cube.crds_([sl, [0, 1], [1, 0], sl, sl])] = M100
    cube.M_[cube.crds_([sl, [2, 1], [1, 2], sl, sl])] = M001
    cube.Δf = torch.sum(-cube.loc[0:1, :, :, :, 0:2], dim=-1) * cube.γ
    Mi_ = cube.M_.clone()
    Ma = cube.applypulse(p, doEmbed=True)
    Ma_ = cube.applypulse(p, doEmbed=False)
    cube2 = deepcopy(cube)
    cube2.applypulse(p, doEmbed=True, doRelax=False, doUpdate=True)
    Mb = cube2.M
    Mo0a = np.array([[[0.559535641648385, 0.663342640621335, 0.416341441715101],
                      [0.391994737048090, 0.210182892388552, -0.860954821972489],
                      [-0.677062008711222, 0.673391604920576, -0.143262993311057]]])
    Mo0b = np.array([[[0.584337330324116, 0.686096989146395, 0.433382978292808],
                      [0.404188676945936, 0.217027890590635, -0.888555236400348],
                      [-0.703691265981316, 0.694384487290747, -0.150495136106067]]])
    assert np.abs(npy(Ma[0:1, 1, :, 1, :]) - Mo0a).max() < 1e-9
    assert np.abs(npy(Mb[0:1, :, 1, 1, :]) - Mo0b).max() < 1e-9
    save('cube27', rf=npy(p.rf), gr=npy(p.gr), dt=npy(p.dt), mask=npy(mask), fov=npy(fov), ofst=npy(ofst),
         loc_=npy(cube.loc_), df_=npy(cube.Δf_), Mi_=npy(Mi_), T1_=npy(cube.T1_.contiguous()),
         T2_=npy(cube.T2_.contiguous()), gamma_=npy(cube.γ_.contiguous()),
         M_relax=npy(Ma), M_relax_=npy(Ma_), M_norelax=npy(Mb), Mo0a=Mo0a, Mo0b=Mo0b)


# ------------------------------------------------------------------------- interp
def case_interp():
    dkw = {'dtype': f64, 'device': torch.device('cpu')}
    out = {}
    # (1) the reference's own golden: nT=11 ramps, 4us -> 20us
    nT = 11
    kw = {'num': nT, 'axis': 2}
    rf = 0.1 * np.concatenate([np.linspace([[0.]], 1., **kw), np.linspace([[1]], 0., **kw)], 1)
    gr = 0.1 * np.concatenate([np.linspace([[0.]], 1., **kw), np.linspace([[1.]], 0., **kw),
                               np.ones((1, 1, nT))], 1)
    p = mobjs.Pulse(rf=torch.tensor(rf, **dkw), gr=torch.tensor(gr, **dkw), dt=mrphy.dt0, **dkw)
    q = p.interpT(dt=mrphy.dt0 * 5, kind='linear')
    out.update(a_rf=rf, a_gr=gr, a_rf_new=npy(q.rf), a_gr_new=npy(q.gr))
    # (2) refinement with the float floor-division quirk (mobjs.py:212): nT=10, 4us -> 2us => 19
    g = torch.Generator().manual_seed(7)
    rf2 = torch.rand((2, 2, 10), generator=g, **dkw) - .5
    gr2 = torch.rand((2, 3, 10), generator=g, **dkw) - .5
    p2 = mobjs.Pulse(rf=rf2, gr=gr2, dt=mrphy.dt0, **dkw)
    q2 = p2.interpT(dt=torch.tensor(2e-6, dtype=f64))
    out.update(b_rf=npy(rf2), b_gr=npy(gr2), b_rf_new=npy(q2.rf), b_gr_new=npy(q2.gr))
    # (3) multi-coil, cubic kind, 20us -> 4us, fp64 (nT 40 -> 200) and fp32 (-> 199)
    rf3 = torch.rand((1, 2, 40, 2), generator=g, **dkw) - .5
    gr3 = torch.rand((1, 3, 40), generator=g, **dkw) - .5
    p3 = mobjs.Pulse(rf=rf3, gr=gr3, dt=torch.tensor(20e-6, dtype=f64), **dkw)
    q3 = p3.interpT(dt=torch.tensor(4e-6, dtype=f64), kind='cubic')
    p3f = mobjs.Pulse(rf=rf3, gr=gr3, dt=torch.tensor(20e-6, dtype=f64), dtype=f32)
    q3f = p3f.interpT(dt=torch.tensor(4e-6, dtype=f64))
    out.update(c_rf=npy(rf3), c_gr=npy(gr3), c_rf_new=npy(q3.rf), c_gr_new=npy(q3.gr),
               c32_rf_new=npy(q3f.rf), c32_gr_new=npy(q3f.gr),
               c_nT=np.array([q3.rf.shape[2], q3f.rf.shape[2]]))
    save('interp', **out)


# ----------------------------------------------------------------- random chains
def rand_case(seed, N, nM, nT, nC, has_b1, has_df, relax, per_spin, per_batch_dt, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    U = lambda *s: torch.rand(s, generator=g, dtype=f64) * 2 - 1
    rf = U(N, 2, nT, nC) * 0.1 * scale if nC > 0 else U(N, 2, nT) * 0.1 * scale
    gr = U(N, 3, nT) * 2 * scale
    loc = U(N, nM, 3) * 12
    df = U(N, nM) * 200 if has_df else None
    b1 = None
    if has_b1:
        b1 = U(N, nM, 2, max(nC, 1)) * 0.1
        b1[:, :, 0] += 1
    M0 = U(N, nM, 3)
    M0 = M0 / M0.norm(dim=-1, keepdim=True)
    if per_spin:
        gam = gH * (1 + 0.05 * U(N, nM))
        T1 = 1.0 + 0.5 * U(N, nM) if relax else None
        T2 = 0.06 + 0.05 * U(N, nM) if relax else None
    else:
        gam = gH.clone()
        T1 = torch.tensor([[1.47]], dtype=f64) if relax else None
        T2 = torch.tensor([[0.07]], dtype=f64) if relax else None
    dt = (torch.tensor([4e-6, 1e-5][:N], dtype=f64) if per_batch_dt else torch.tensor([4e-6], dtype=f64))
    w = U(N, nM, 3)
    return dict(M0=M0, rf=rf, gr=gr, loc=loc, df=df, b1=b1, T1=T1, T2=T2, gam=gam, dt=dt, w=w)


def run_both_precisions(name, c, also_slow_gM0=True):
    """Round the fp64 inputs to fp32, then run the reference in fp32 and in fp64 on the SAME
    (upcast) values: gives the reference's own fp32 noise floor next to the fp64 truth."""
    c32 = {k: (None if v is None else v.to(f32)) for k, v in c.items()}
    c64 = {k: (None if v is None else v.to(f64)) for k, v in c32.items()}
    out = {}
    for tag, cc in (('f64', c64), ('f32', c32)):
        order = [cc[k] for k in ('M0', 'rf', 'gr', 'loc', 'df', 'b1', 'T1', 'T2', 'gam', 'dt', 'w')]
        Mo, grf, ggr, gM0 = slow_chain(*order)
        out[f'Mo_slow_{tag}'], out[f'grf_slow_{tag}'], out[f'ggr_slow_{tag}'], out[f'gM0_slow_{tag}'] = \
            npy(Mo), npy(grf), npy(ggr), npy(gM0)
        try:
            Mo, beff, grf, ggr, gM0 = ref_chain(*order)
            out[f'gM0_{tag}'] = npy(gM0)
        except RuntimeError as e:  # sims.py:267 indexing bug with per-spin gamma
            print(f'  [{name}/{tag}] reference grad_Mi bug hit: {str(e)[:60]}...')
            M0, rf, gr = cc['M0'], cc['rf'].clone().requires_grad_(True), cc['gr'].clone().requires_grad_(True)
            beff = beffective.rfgr2beff(rf, gr, cc['loc'], Δf=cc['df'], b1Map=cc['b1'], γ=cc['gam'])
            Mo = sims.blochsim(M0, beff, T1=cc['T1'], T2=cc['T2'], γ=cc['gam'], dt=cc['dt'])
            (Mo * cc['w']).sum().backward()
            Mo, beff, grf, ggr = Mo.detach(), beff.detach(), rf.grad, gr.grad
        out[f'Mo_{tag}'], out[f'grf_{tag}'], out[f'ggr_{tag}'] = npy(Mo), npy(grf), npy(ggr)
        if tag == 'f64':
            out['beff_f64'] = npy(beff)
            # explicit-Beff API: grads wrt Mi (slowsims) and Beff (sims)
            b = beff.clone().requires_grad_(True)
            Mo2 = sims.blochsim(cc['M0'], b, T1=cc['T1'], T2=cc['T2'], γ=cc['gam'], dt=cc['dt'])
            (Mo2 * cc['w']).sum().backward()
            out['gBeff_f64'] = npy(b.grad)
            if cc['T1'] is not None:
                dtb = cc['dt'].reshape(-1, 1)
                E1, E2 = torch.exp(-dtb / cc['T1']), torch.exp(-dtb / cc['T2'])
                A, B = beffective.beff2ab(beff, E1=E1, E2=E2, γ=cc['gam'], dt=cc['dt'])
                out['A_f64'], out['B_f64'] = npy(A), npy(B)
    print(f'  [{name}] ref fp32 vs fp64: max|dM|={np.abs(out["Mo_f32"]-out["Mo_f64"]).max():.2e} '
          f'grf rel={np.linalg.norm(out["grf_f32"]-out["grf_f64"])/np.linalg.norm(out["grf_f64"]):.2e} '
          f'ggr rel={np.linalg.norm(out["ggr_f32"]-out["ggr_f64"])/np.linalg.norm(out["ggr_f64"]):.2e}')
    inp = {('in_' + k): npy(v) for k, v in c32.items()}
    save(name, **inp, **out)


def case_bench8():
    """8^3 SpinCube with the BASELINE.md section 2 distributions, nT=1000 (the bench workload in small)."""
    g = torch.Generator().manual_seed(0)
    n, nT = 8, 1000
    U = lambda *s: torch.rand(s, generator=g, dtype=f64) * 2 - 1
    cube = mobjs.SpinCube((1, n, n, n), torch.tensor([[24., 24., 24.]]), dtype=f64)
    loc = cube.loc_.clone()
    nM = n ** 3
    df = U(1, nM) * 200
    b1 = U(1, nM, 2, 1) * 0.1
    b1[:, :, 0] += 1
    c = dict(M0=torch.tensor([0., 0., 1.], dtype=f64).expand(1, nM, 3).clone(),
             rf=U(1, 2, nT, 1) * 0.1, gr=U(1, 3, nT) * 2, loc=loc, df=df, b1=b1,
             T1=torch.tensor([[1.47]], dtype=f64), T2=torch.tensor([[0.07]], dtype=f64),
             gam=gH.clone(), dt=torch.tensor([4e-6], dtype=f64), w=None)
    # loss = sum((M - target)^2), target = [0,1,0]  => dL/dM = 2 (M - target); linearise around fp64 M
    c32 = {k: (None if v is None else v.to(f32)) for k, v in c.items()}
    c64 = {k: (None if v is None else v.to(f64)) for k, v in c32.items()}
    out = {}
    for tag, cc in (('f64', c64), ('f32', c32)):
        rf = cc['rf'].clone().requires_grad_(True)
        gr = cc['gr'].clone().requires_grad_(True)
        beff = beffective.rfgr2beff(rf, gr, cc['loc'], Δf=cc['df'], b1Map=cc['b1'], γ=cc['gam'])
        Mo = sims.blochsim(cc['M0'], beff, T1=cc['T1'], T2=cc['T2'], γ=cc['gam'], dt=cc['dt'])
        tgt = torch.tensor([0., 1., 0.], dtype=Mo.dtype)
        ((Mo - tgt) ** 2).sum().backward()
        out[f'Mo_{tag}'], out[f'grf_{tag}'], out[f'ggr_{tag}'] = npy(Mo), npy(rf.grad), npy(gr.grad)
    print(f'  [bench8] ref fp32 vs fp64: max|dM|={np.abs(out["Mo_f32"]-out["Mo_f64"]).max():.2e}')
    inp = {('in_' + k): npy(v) for k, v in c32.items()}
    save('bench8', **inp, **out)


def case_freeprec():
    dkw = {'dtype': f64}
    Mi = torch.eye(3, **dkw)[None]
    E = torch.tensor([[0.5]], **dkw)
    dur = torch.tensor(0.5, **dkw)
    T1 = T2 = -dur / torch.log(E)
    df = torch.tensor([[1 / 4 / dur, -1 / 4 / dur, 1]], **dkw)
    Mo = sims.freeprec(Mi, dur, T1=T1, T2=T2, Δf=df)
    assert np.abs(npy(Mo) - np.array([[[0., -0.5, 0.5], [-0.5, 0, 0.5], [0., 0., 1.]]])).max() < 1e-9
    g = torch.Generator().manual_seed(11)
    Mr = (torch.rand((2, 9, 3), generator=g, **dkw) - .5).requires_grad_(True)
    dur2 = torch.tensor([0.01, 0.02], **dkw)
    T1r = 1 + torch.rand((2, 9), generator=g, **dkw)
    T2r = 0.05 + 0.05 * torch.rand((2, 9), generator=g, **dkw)
    dfr = 100 * (torch.rand((2, 9), generator=g, **dkw) - .5)
    w = torch.rand((2, 9, 3), generator=g, **dkw)
    Mor = sims.freeprec(Mr, dur2, T1=T1r, T2=T2r, Δf=dfr)
    (Mor * w).sum().backward()
    save('freeprec', a_Mi=npy(Mi), a_dur=npy(dur), a_T1=npy(T1), a_T2=npy(T2), a_df=npy(df), a_Mo=npy(Mo),
         b_Mi=npy(Mr), b_dur=npy(dur2), b_T1=npy(T1r), b_T2=npy(T2r), b_df=npy(dfr), b_w=npy(w),
         b_Mo=npy(Mor), b_gMi=npy(Mr.grad))


def case_reparam():
    from mrphy import utils
    g = torch.Generator().manual_seed(77)
    R = lambda *s: torch.randn(s, generator=g, dtype=f64)
    out = {}
    for tag, nC in (('sc', 0), ('mc', 3)):
        N, nT = 2, 300
        shp = (N, 1, nT) + ((nC,) if nC else ())
        rho, theta, ts = (R(*shp) * 2).requires_grad_(True), (R(*shp) * 3).requires_grad_(True), (R(N, 3, nT) * 1.5).requires_grad_(True)
        rfmax = torch.rand((N, nC) if nC else (N,), generator=g, dtype=f64) * 0.2 + 0.05
        smax = torch.rand((N, 3), generator=g, dtype=f64) * 1e4 + 5e3
        dt = torch.tensor([4e-6, 1e-5], dtype=f64)
        wrf, wg = R(N, 2, *shp[2:]), R(N, 3, nT)
        rf_t = utils.tρθ2rf(rho, theta, rfmax)
        rf_l = utils.lρθ2rf(rho, theta, rfmax)
        s = utils.ts2s(ts, smax)
        gr = utils.s2g(s, dt)
        grho_t, gtheta_t = torch.autograd.grad((rf_t * wrf).sum(), (rho, theta))
        grho_l, gtheta_l = torch.autograd.grad((rf_l * wrf).sum(), (rho, theta))
        gts, = torch.autograd.grad((gr * wg).sum(), (ts,))
        gts_s, = torch.autograd.grad((utils.ts2s(ts, smax) * wg).sum(), (ts,))
        loc = dict(rho=rho, theta=theta, ts=ts, rfmax=rfmax, smax=smax, dt=dt, wrf=wrf, wg=wg, rf_t=rf_t, rf_l=rf_l, s=s,
                   gr=gr, grho_t=grho_t, gtheta_t=gtheta_t, grho_l=grho_l, gtheta_l=gtheta_l, gts=gts, gts_s=gts_s)
        out.update({f'{tag}_{k}': npy(v) for k, v in loc.items()})
    # mask plumbing: SpinArray.embed / extract (mobjs.py:512-553)
    mask = torch.rand((1, 5, 4, 3), generator=g) > 0.4
    sa = mobjs.SpinArray((2, 5, 4, 3), mask=mask, dtype=f64)
    v_ = R(2, sa.nM, 3)
    emb = sa.embed(v_)
    full = R(2, 5, 4, 3, 2)
    out.update(mask=npy(mask), mask_v_=npy(v_), mask_embedded=npy(emb), mask_full=npy(full), mask_extracted=npy(sa.extract(full)))
    save('reparam', **out)


def case_sequence():
    """pulse -> free precession -> pulse on a masked SpinCube, chained through doUpdate exactly as a user of the
    reference writes it (mobjs.py:841-896): forward values from that chain.  Upstream's BlochSim.backward cannot return
    grad_Mi for the (N, nM)-expanded constants mobjs passes (sims.py:267 indexes them wrongly), so the gradients
    w.r.t. both pulses' waveforms come from the same chain written with the reference's plain-autograd simulator
    (rfgr2beff -> slowsims.blochsim, sims.freeprec), which reproduces the forward values."""
    g = torch.Generator().manual_seed(31)
    U = lambda *s: torch.rand(s, generator=g, dtype=f64) * 2 - 1
    dkw = {'dtype': f64}
    N, shape = 2, (2, 4, 3, 3)
    mask = torch.rand((1,) + shape[1:], generator=g) > 0.3
    fov, ofst = torch.tensor([[6., 5., 4.]], **dkw), torch.tensor([[0.5, 0., -0.5]], **dkw)
    cube = mobjs.SpinCube(shape, fov, mask=mask, ofst=ofst, **dkw)
    nM = cube.nM
    cube.Δf_ = U(N, nM) * 150
    cube.T1_ = 1 + 0.3 * U(N, nM)
    cube.T2_ = 0.06 + 0.02 * U(N, nM)
    b1 = torch.stack((1 + 0.1 * U(N, nM), 0.1 * U(N, nM)), dim=-1)
    rf1, gr1 = (U(N, 2, 30) * 0.2).requires_grad_(True), (U(N, 3, 30) * 2).requires_grad_(True)
    rf2, gr2 = (U(N, 2, 20) * 0.2).requires_grad_(True), (U(N, 3, 20) * 2).requires_grad_(True)
    dt1, dt2 = torch.tensor(4e-6, **dkw), torch.tensor(8e-6, **dkw)
    dur = torch.tensor([2e-3, 5e-3], **dkw)
    M0 = cube.M_.clone()
    with torch.no_grad():
        p1 = mobjs.Pulse(rf=rf1.detach(), gr=gr1.detach(), dt=dt1, **dkw)
        p2 = mobjs.Pulse(rf=rf2.detach(), gr=gr2.detach(), dt=dt2, **dkw)
        cube.applypulse(p1, b1Map_=b1, doUpdate=True)
        Ma = npy(cube.M_)
        cube.freeprec(dur, doUpdate=True)
        Mb = npy(cube.M_)
        Mc = npy(cube.applypulse(p2, b1Map_=b1))
    kw = dict(T1=cube.T1_, T2=cube.T2_, γ=cube.γ_)
    m = slowsims.blochsim(M0, beffective.rfgr2beff(rf1, gr1, cube.loc_, Δf=cube.Δf_, b1Map=b1, γ=cube.γ_), dt=dt1, **kw)
    m = sims.freeprec(m, dur, T1=cube.T1_, T2=cube.T2_, Δf=cube.Δf_)
    m = slowsims.blochsim(m, beffective.rfgr2beff(rf2, gr2, cube.loc_, Δf=cube.Δf_, b1Map=b1, γ=cube.γ_), dt=dt2, **kw)
    assert np.abs(npy(m) - Mc).max() < 1e-12
    w = U(N, nM, 3)
    grads = torch.autograd.grad((m * w).sum(), (rf1, gr1, rf2, gr2))
    save('sequence', mask=npy(mask), fov=npy(fov), ofst=npy(ofst), df=npy(cube.Δf_), T1=npy(cube.T1_), T2=npy(cube.T2_),
         b1=npy(b1), rf1=npy(rf1), gr1=npy(gr1), rf2=npy(rf2), gr2=npy(gr2), dur=npy(dur), M0=npy(M0), Ma=Ma, Mb=Mb, Mc=Mc,
         w=npy(w), grf1=npy(grads[0]), ggr1=npy(grads[1]), grf2=npy(grads[2]), ggr2=npy(grads[3]))


def case_standalone():
    """Outputs AND autograd gradients of the reference's stand-alone operators on seeded inputs (fp64)."""
    from mrphy import utils
    g = torch.Generator().manual_seed(2024)
    U = lambda *s: torch.rand(s, generator=g, dtype=f64) * 2 - 1
    out = {}
    # ---- beff2uϕ (beffective.py:18-37), including a zero-field spin
    N, Nd = 2, (3, 4)
    beff = (U(N, *Nd, 3) * 3)
    beff[0, 0, 0] = 0
    beff = beff.requires_grad_(True)
    g2 = (2 * PI * 4257.6 * 4e-6 * (1 + 0.1 * U(N, *Nd))).requires_grad_(True)
    wU, wP = U(N, *Nd, 3), U(N, *Nd)
    Uo, Po = beffective.beff2uϕ(beff, g2)
    gb, gg = torch.autograd.grad((Uo * wU).sum() + (Po * wP).sum(), (beff, g2))
    out.update(uphi_beff=npy(beff), uphi_g=npy(g2), uphi_wU=npy(wU), uphi_wP=npy(wP), uphi_U=npy(Uo), uphi_Phi=npy(Po),
               uphi_gbeff=npy(gb), uphi_gg=npy(gg))
    # ---- rfgr2beff (beffective.py:107-168) with every gradient; coil-broadcast variants
    nT = 9
    for tag, Nd, rf_c, b1_c in (('mc', (7,), 2, 2), ('b1bc', (5,), 3, 1), ('rfbc', (6,), 0, 3), ('nob1', (3, 2, 2), 3, None),
                                ('sc', (4,), 0, 0)):
        rf = (U(N, 2, nT, rf_c) if rf_c else U(N, 2, nT)).requires_grad_(True)
        gr = U(N, 3, nT).requires_grad_(True)
        loc = (U(N, *Nd, 3) * 5).requires_grad_(True)
        df = (U(N, *Nd) * 100).requires_grad_(True)
        gam = (4257.6 * (1 + 0.1 * U(N, *Nd))).requires_grad_(True)
        b1 = None if b1_c is None else (U(N, *Nd, 2, b1_c) if b1_c else U(N, *Nd, 2)).requires_grad_(True)
        w = U(N, *Nd, nT, 3)
        be = beffective.rfgr2beff(rf, gr, loc, Δf=df, b1Map=b1, γ=gam)
        ins = [rf, gr, loc, df, gam] + ([b1] if b1 is not None else [])
        gs = torch.autograd.grad((be * w).sum(), ins)
        names = ['rf', 'gr', 'loc', 'df', 'gam'] + (['b1'] if b1 is not None else [])
        out.update({f'b_{tag}_{k}': npy(v) for k, v in zip(names, ins)})
        out.update({f'b_{tag}_g{k}': npy(v) for k, v in zip(names, gs)})
        out.update({f'b_{tag}_w': npy(w), f'b_{tag}_beff': npy(be)})
    # ---- beff2ab (beffective.py:40-104) with gradients wrt beff, E1, E2
    nT = 17
    bf = (U(N, 6, nT, 3) * 2).requires_grad_(True)
    E1 = (0.95 + 0.04 * U(N, 6)).requires_grad_(True)
    E2 = (0.9 + 0.05 * U(N, 6)).requires_grad_(True)
    wA, wB = U(N, 6, 3, 3), U(N, 6, 3)
    A, B = beffective.beff2ab(bf, E1=E1, E2=E2, γ=gH.to(f64), dt=torch.tensor(4e-6, dtype=f64))
    gbf, gE1, gE2 = torch.autograd.grad((A * wA).sum() + (B * wB).sum(), (bf, E1, E2))
    out.update(ab_beff=npy(bf), ab_E1=npy(E1), ab_E2=npy(E2), ab_wA=npy(wA), ab_wB=npy(wB), ab_A=npy(A), ab_B=npy(B),
               ab_gbeff=npy(gbf), ab_gE1=npy(gE1), ab_gE2=npy(gE2))
    # ---- utils.rfclamp / sclamp (utils.py:217-236, 278-293) with gradients
    for tag, nC in (('sc', 0), ('mc', 3)):
        rf = (U(N, 2, 40, nC) if nC else U(N, 2, 40)) * 0.3
        rf = rf.requires_grad_(True)
        rfmax = torch.rand((N, nC) if nC else (N,), generator=g, dtype=f64) * 0.2 + 0.1
        wr = U(*rf.shape)
        rc = utils.rfclamp(rf, rfmax)
        grc, = torch.autograd.grad((rc * wr).sum(), (rf,))
        out.update({f'cl_{tag}_rf': npy(rf), f'cl_{tag}_rfmax': npy(rfmax), f'cl_{tag}_w': npy(wr), f'cl_{tag}_out': npy(rc),
                    f'cl_{tag}_grf': npy(grc)})
    sl = (U(N, 3, 40) * 2e4).requires_grad_(True)
    smax = torch.rand((N, 3), generator=g, dtype=f64) * 1e4 + 5e3
    ws = U(N, 3, 40)
    sc = utils.sclamp(sl, smax)
    gsl, = torch.autograd.grad((sc * ws).sum(), (sl,))
    out.update(cl_s=npy(sl), cl_smax=npy(smax), cl_ws=npy(ws), cl_sout=npy(sc), cl_gs=npy(gsl))
    save('standalone', **out)


if __name__ == '__main__':
    torch.set_num_threads(8)
    if len(sys.argv) > 1:                      # regenerate selected cases only: python make_golden.py standalone ...
        for name in sys.argv[1:]:
            globals()['case_' + name]()
        sys.exit(0)
    case_kat3()
    case_sims512()
    case_cube27()
    case_interp()
    run_both_precisions('rand_mc', rand_case(3, N=2, nM=7, nT=50, nC=2, has_b1=True, has_df=True,
                                             relax=True, per_spin=True, per_batch_dt=False, scale=10))
    run_both_precisions('rand_nob1', rand_case(4, N=2, nM=5, nT=40, nC=3, has_b1=False, has_df=False,
                                               relax=True, per_spin=False, per_batch_dt=True, scale=10))
    run_both_precisions('rand_norelax', rand_case(5, N=1, nM=33, nT=130, nC=0, has_b1=True, has_df=True,
                                                  relax=False, per_spin=False, per_batch_dt=False, scale=5))
    case_bench8()
    case_freeprec()
    case_reparam()
    case_standalone()
    case_sequence()

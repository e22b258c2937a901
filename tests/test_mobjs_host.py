"""Host-side object layer (mobjs / utils / slowsims) on CPU: coercion rules, mask embed/extract,
asdict/deepcopy round trips, interpT goldens -- mirrors /root/reference/tests/test_mobjs.py:14-59,
160-195 and tests/test_utils.py."""
from copy import deepcopy

import numpy as np
import pytest
import torch
from torch import tensor

import mrphy
from mrphy import γH, dt0, π, _slice, rfmax0, smax0
from mrphy import mobjs, utils, slowsims, beffective

f64 = torch.float64


def _setup(T1_, T2, γ, device, dtype):
    kw = {'dtype': dtype, 'device': device}
    N, Nd, nT = 1, (3, 3, 3), 512
    t = torch.arange(0, nT, **kw).reshape((N, 1, nT))
    rf = 10 * torch.cat([torch.cos(t / nT * 2 * π), torch.sin(t / nT * 2 * π)], 1)
    gr = torch.cat([torch.ones((N, 1, nT), **kw), torch.ones((N, 1, nT), **kw),
                    10 * torch.atan(t - round(nT / 2)) / π], 1)
    p = mobjs.Pulse(rf=rf, gr=gr, dt=dt0, **kw)
    p = deepcopy(p)
    p = mobjs.Pulse(**(p.asdict(toNumpy=False)))
    shape = (N, *Nd)
    mask = torch.zeros((1,) + Nd, device=device, dtype=torch.bool)
    mask[0, :, 1, :], mask[0, 1, :, :] = True, True
    fov, ofst = tensor([[3., 3., 3.]], **kw), tensor([[0., 0., 1.]], **kw)
    cube = mobjs.SpinCube(shape, fov, mask=mask, T1_=T1_.to(**kw), γ=γ, **kw)
    cube = deepcopy(cube)
    d = cube.asdict(toNumpy=False)
    cube = mobjs.SpinCube(**{k: d[k] for k in ('shape', 'fov', 'mask', 'T1', 'γ') + tuple(kw.keys())})
    cube.ofst = ofst
    cube.M_ = tensor([0., 1., 0.])
    cube.T2 = T2.to(**kw).expand(cube.shape)
    M001, M100 = tensor([0., 0., 1.], **kw), tensor([1., 0., 0.], **kw)
    cube.M_[cube.crds_([_slice, [0, 1], [1, 0], _slice, _slice])] = M100
    cube.M_[cube.crds_([_slice, [2, 1], [1, 2], _slice, _slice])] = M001
    return cube, p


def test_constants_are_float64_scalars():
    for c in (mrphy.γH, mrphy.T1G, mrphy.T2G, mrphy.dt0, mrphy.gmax0, mrphy.smax0, mrphy.rfmax0):
        assert c.dtype == f64 and c.ndim == 0
    assert float(mrphy.γH) == 4257.6 and mrphy.__all__ == ['γH', 'utils', 'beffective', 'sims', 'slowsims', 'mobjs']
    assert hasattr(beffective, 'beff2uφ') and hasattr(utils, 'uφrot') and mrphy.sims.__all__ == ['blochsim']


def test_examples_and_inheritance():
    assert isinstance(mobjs.Examples.pulse(), mobjs.Pulse)
    assert isinstance(mobjs.Examples.spincube(), mobjs.SpinCube)
    assert isinstance(mobjs.Examples.spincube(), mobjs.SpinArray)
    assert isinstance(mobjs.Examples.spinarray(), mobjs.SpinArray)


def test_setup_matches_reference_fixture(golden):
    g = golden('cube27')
    cube, p = _setup(tensor([[1.]]), tensor([[4e-2]]), γH.to(f64), torch.device('cpu'), f64)
    cube.Δf = torch.sum(-cube.loc[0:1, :, :, :, 0:2], dim=-1) * cube.γ
    assert p.is_cuda is False and cube.is_cuda is False and cube.dim() == 4 and cube.nM == 15
    for k, ref in (('loc_', 'loc_'), ('Δf_', 'df_'), ('M_', 'Mi_'), ('T1_', 'T1_'), ('T2_', 'T2_'), ('γ_', 'gamma_')):
        assert np.abs(getattr(cube, k).numpy() - g[ref]).max() < 1e-8, k     # T2 is given in fp32 upstream
    assert np.array_equal(cube.mask.numpy(), g['mask'])
    assert np.abs(p.rf.numpy() - g['rf']).max() < 1e-12 and np.abs(p.gr.numpy() - g['gr']).max() < 1e-12
    # embed pads with NaN outside the mask, extract is its inverse on the mask
    M = cube.M
    assert M.shape == (1, 3, 3, 3, 3) and torch.isnan(M[0, 0, 0, 0]).all()
    assert torch.equal(cube.extract(M), cube.M_)


def test_pulse_coercion_rules():
    p = mobjs.Pulse(rf=torch.zeros(2, 2, 7), dt=dt0)
    assert p.shape == (2, 1, 7) and p.gr.shape == (2, 3, 7) and p.dt.shape == (1,) and p.dt.dtype == torch.float32
    assert p.gmax.shape == (1, 3) and p.smax.shape == (1, 3) and p.rfmax.shape == (1,)
    p.rfmax = torch.full((2, 1), 0.2)
    assert p.rfmax.shape == (2,)
    with pytest.raises(AttributeError):
        p.shape = (1, 1, 1)
    with pytest.raises(AssertionError):
        mobjs.Pulse(rf=torch.zeros(1, 2, 3), device='cpu')
    with pytest.raises(AssertionError):
        mobjs.Pulse()
    rf = torch.zeros(1, 2, 5, requires_grad=True)
    assert mobjs.Pulse(rf=rf).rf is rf                      # no cast needed => user's leaf is kept
    q = p.to(dtype=f64)
    assert q.dtype == f64 and q.rfmax.shape == (1,) and p.to(dtype=torch.float32) is p
    assert set(p.asdict()) == {'rf', 'gr', 'dt', 'gmax', 'smax', 'rfmax', 'desc', 'device', 'dtype'}


def test_spinarray_coercion_rules():
    sp = mobjs.SpinArray((2, 4, 5))
    assert sp.nM == 20 and sp.T2_.shape == (2, 20) and sp.T2_.stride() == (0, 0) and sp.M_.is_contiguous()
    assert float(sp.T1_[0, 0]) == pytest.approx(1.47) and sp.M.shape == (2, 4, 5, 3)
    sp.T1 = torch.arange(20.).reshape(1, 4, 5) + 1
    assert sp.T1_.shape == (2, 20) and float(sp.T1_[1, 7]) == 8.
    with pytest.raises(AssertionError):
        mobjs.SpinArray((1, 2), T1=tensor(1.), T1_=tensor(1.))
    with pytest.raises(AttributeError):
        sp.nM = 3
    cube = mobjs.SpinCube((1, 4, 4, 2), tensor([[8., 8., 4.]]))
    assert torch.allclose(cube.loc_[0, 0], tensor([-4., -4., -2.])) and torch.allclose(cube.loc_[0, -1], tensor([2., 2., 0.]))
    assert cube.Δf_.shape == (1, 32) and set(cube.asdict()) >= {'loc', 'Δf', 'fov', 'ofst', 'T1', 'T2', 'γ', 'M', 'mask', 'shape'}
    with pytest.raises(AttributeError):
        cube.loc_ = torch.zeros(1, 32, 3)
    c2 = deepcopy(cube)
    c2.fov = tensor([[4., 4., 4.]])
    assert torch.allclose(c2.loc_[0, 0], tensor([-2., -2., -2.])) and torch.allclose(cube.loc_[0, 0], tensor([-4., -4., -2.]))
    m = torch.zeros(1, 4, 4, 2, dtype=torch.bool)
    m[0, 0] = True
    assert cube.spinarray.mask_(mask=m).sum() == 8


def test_interpT_goldens(golden):
    g = golden('interp')
    kw = {'dtype': f64, 'device': torch.device('cpu')}
    p = mobjs.Pulse(rf=tensor(g['a_rf'], **kw), gr=tensor(g['a_gr'], **kw), dt=dt0, **kw)
    q = p.interpT(dt=dt0 * 5, kind='linear')
    assert q.rf.numpy() == pytest.approx(np.array([[[0.04, 0.09], [0.06, 0.01]]]), abs=1e-9)   # test_mobjs.py:185
    assert q.gr.numpy() == pytest.approx(np.array([[[0.04, 0.09], [0.06, 0.01], [0.1, 0.1]]]), abs=1e-9)
    assert p.interpT(dt=dt0).rf is not p.rf and torch.equal(p.interpT(dt=dt0).rf, p.rf)
    p2 = mobjs.Pulse(rf=tensor(g['b_rf'], **kw), gr=tensor(g['b_gr'], **kw), dt=dt0, **kw)
    q2 = p2.interpT(dt=tensor(2e-6, dtype=f64))
    assert q2.rf.shape[2] == 19                                                              # float // quirk
    assert np.abs(q2.rf.numpy() - g['b_rf_new']).max() < 1e-12 and np.abs(q2.gr.numpy() - g['b_gr_new']).max() < 1e-12
    p3 = mobjs.Pulse(rf=tensor(g['c_rf'], **kw), gr=tensor(g['c_gr'], **kw), dt=tensor(20e-6, dtype=f64), **kw)
    q3 = p3.interpT(dt=tensor(4e-6, dtype=f64), kind='cubic')
    assert np.abs(q3.rf.numpy() - g['c_rf_new']).max() < 1e-12 and q3.rf.shape == (1, 2, 200, 2)
    p3f = mobjs.Pulse(rf=tensor(g['c_rf']), gr=tensor(g['c_gr']), dt=tensor(20e-6, dtype=f64), dtype=torch.float32)
    q3f = p3f.interpT(dt=tensor(4e-6, dtype=f64))
    assert q3f.rf.shape[2] == 199 and np.abs(q3f.rf.numpy() - g['c32_rf_new']).max() < 1e-6
    assert q3f.gmax.shape == (1, 3)


def test_slowsims_match_reference_and_cuda_operators_refuse_cpu(golden):
    """slowsims (plain torch, the reference's own cross-check) reproduces the reference on CPU; the operators of
    beffective / sims are CUDA kernels only and say so instead of computing anything on the host."""
    g = golden('rand_mc')
    T = lambda k: tensor(g[k], dtype=f64)
    beff, dt = T('beff_f64'), T('in_dt')
    Mo = slowsims.blochsim_ab(T('in_M0'), T('A_f64'), T('B_f64'))
    assert np.abs(Mo.numpy() - g['Mo_f64']).max() < 1e-12
    Ms = slowsims.blochsim(T('in_M0'), beff, T1=T('in_T1'), T2=T('in_T2'), γ=T('in_gam'), dt=dt)
    assert np.abs(Ms.numpy() - g['Mo_slow_f64']).max() < 1e-12
    with pytest.raises(RuntimeError, match='CUDA-only'):
        beffective.rfgr2beff(T('in_rf'), T('in_gr'), T('in_loc'), Δf=T('in_df'), b1Map=T('in_b1'), γ=T('in_gam'))
    with pytest.raises(RuntimeError, match='CUDA-only'):
        beffective.beff2ab(beff, E1=torch.exp(-dt / T('in_T1')), E2=torch.exp(-dt / T('in_T2')), γ=T('in_gam'), dt=dt)
    with pytest.raises(RuntimeError, match='CUDA-only'):
        beffective.beff2uϕ(beff[..., 0, :], tensor(1., dtype=f64))


def test_freeprec_goldens(golden):
    """tests/test_slowsims.py:100-122 (slowsims.freeprec on CPU; sims.freeprec is a CUDA kernel -> tests/test_gpu_parity)."""
    g = golden('freeprec')
    T = lambda k: tensor(g[k], dtype=f64)
    Mo = slowsims.freeprec(T('a_Mi'), T('a_dur'), T1=T('a_T1'), T2=T('a_T2'), Δf=T('a_df'))
    assert np.abs(Mo.numpy() - np.array([[[0., -0.5, 0.5], [-0.5, 0, 0.5], [0., 0., 1.]]])).max() < 1e-12
    Mi = T('b_Mi').requires_grad_(True)
    Mo = slowsims.freeprec(Mi, T('b_dur'), T1=T('b_T1'), T2=T('b_T2'), Δf=T('b_df'))
    (Mo * T('b_w')).sum().backward()
    assert np.abs(Mo.detach().numpy() - g['b_Mo']).max() < 1e-13
    assert np.abs(Mi.grad.numpy() - g['b_gMi']).max() < 1e-13
    with pytest.raises(RuntimeError, match='CUDA-only'):
        mrphy.sims.freeprec(T('a_Mi'), T('a_dur'), T1=T('a_T1'), T2=T('a_T2'), Δf=T('a_df'))


class TestUtils:
    """/root/reference/tests/test_utils.py:31-94, fp32, atol 1e-4."""
    kw = {'dtype': torch.float32, 'device': torch.device('cpu')}
    atol = 1e-4

    def test_ctrsub(self):
        assert np.all(utils.ctrsub(torch.arange(7, **self.kw)).numpy() == np.array([0, 0, 1, 1, 2, 2, 3]))

    def test_kgs(self):
        γ, dt = γH.to(**self.kw), dt0.to(**self.kw)
        k = tensor([[[1., 2., 3., 4., 0.]]], **self.kw)
        gTx, gRx = utils.k2g(k, True, γ=γ, dt=dt), utils.k2g(k, False, γ=γ, dt=dt)
        assert utils.g2k(gTx, True, γ=γ, dt=dt).numpy() == pytest.approx(k.numpy(), abs=self.atol)
        assert utils.g2k(gRx, False, γ=γ, dt=dt).numpy() == pytest.approx(k.numpy(), abs=self.atol)
        assert gTx.numpy() == pytest.approx(utils.s2g(utils.g2s(gTx, dt), dt).numpy(), abs=self.atol)

    def test_rc_rf(self):
        x = np.random.rand(1, 2, 5)
        assert x == pytest.approx(utils.rf_c2r(utils.rf_r2c(x)), abs=self.atol)

    def test_rfclamp_tan_logit(self):
        rf0 = utils.rfclamp(rfmax0 * ((torch.rand((1, 2, 10)) - 0.5) * 4), rfmax0)
        assert torch.all(rf0.norm(dim=1) <= rfmax0)
        assert rf0.numpy() == pytest.approx(utils.tρθ2rf(*utils.rf2tρθ(rf0, rfmax0), rfmax0).numpy(), abs=self.atol)
        assert rf0.numpy() == pytest.approx(utils.lρθ2rf(*utils.rf2lρθ(rf0, rfmax0), rfmax0).numpy(), abs=self.atol)

    def test_sclamptan(self):
        s0 = utils.sclamp(smax0 * ((torch.rand((1, 3, 10)) - 0.5) * 4), smax0)
        assert torch.all(s0.abs() <= smax0)
        assert s0.numpy() == pytest.approx(utils.ts2s(utils.s2ts(s0, smax0), smax0).numpy(), abs=self.atol)


@pytest.mark.parametrize('tag', ['sc', 'mc'])
def test_reparam_utils_match_reference_on_cpu(golden, tag):
    """tρθ2rf / lρθ2rf / ts2s / s2g (+ the fused conveniences) on CPU tensors against outputs and autograd gradients
    of the unmodified reference (tests/golden/reparam.npz)."""
    g = {k[len(tag) + 1:]: tensor(v) for k, v in golden('reparam').items() if k.startswith(tag + '_')}
    rho, theta, ts = (g[k].clone().requires_grad_(True) for k in ('rho', 'theta', 'ts'))
    for kind, fn in (('t', utils.tρθ2rf), ('l', utils.lρθ2rf)):
        rf = fn(rho, theta, g['rfmax'])
        assert torch.allclose(rf, g['rf_' + kind], rtol=0, atol=1e-14)
        grho, gtheta = torch.autograd.grad((rf * g['wrf']).sum(), (rho, theta))
        assert torch.allclose(grho, g['grho_' + kind], rtol=1e-12, atol=1e-15)
        assert torch.allclose(gtheta, g['gtheta_' + kind], rtol=1e-12, atol=1e-15)
    s = utils.ts2s(ts, g['smax'])
    gr = utils.s2g(s, g['dt'])
    assert torch.allclose(s, g['s'], rtol=1e-14, atol=0) and torch.allclose(gr, g['gr'], rtol=1e-13, atol=1e-15)
    assert torch.allclose(utils.ts2g(ts, g['smax'], g['dt']), g['gr'], rtol=1e-13, atol=1e-15)
    rf2, gr2 = utils.tρθts2rfgr(rho, theta, ts, g['rfmax'], g['smax'], g['dt'])
    assert torch.allclose(rf2, g['rf_t'], rtol=0, atol=1e-14) and torch.allclose(gr2, g['gr'], rtol=1e-13, atol=1e-15)
    gts, = torch.autograd.grad((gr2 * g['wg']).sum(), (ts,))
    assert torch.allclose(gts, g['gts'], rtol=1e-12, atol=1e-18)


def test_embed_extract_match_reference_on_cpu(golden):
    g = golden('reparam')
    sa = mobjs.SpinArray((2, 5, 4, 3), mask=tensor(g['mask']), dtype=f64)
    emb = sa.embed(tensor(g['mask_v_'])).numpy()
    assert np.array_equal(np.isnan(emb), np.isnan(g['mask_embedded']))
    assert np.array_equal(np.nan_to_num(emb), np.nan_to_num(g['mask_embedded']))
    assert np.array_equal(sa.extract(tensor(g['mask_full'])).numpy(), g['mask_extracted'])


def test_checkpoint_policy_and_its_caches():
    """K = floor(0.4 / (max(dt) / min(T1, T2))) capped at 64 and rounded down to a multiple of 16 when >= 16; min(T1, T2)
    and max(dt) are cached per tensor object AND in-place version, separately (a design loop keeps its spins but builds
    a new Pulse, hence a new dt, every iteration)."""
    from mrphy import _ops
    T1, T2 = torch.full((1, 5), 1.0, dtype=f64), torch.full((1, 5), 0.05, dtype=f64)
    dt = tensor([4e-6], dtype=f64)
    assert _ops.pick_ckpt_interval(dt, None, None) == 64                       # no relaxation: nothing to amplify
    assert _ops.pick_ckpt_interval(dt, T1, T2) == 64                           # 0.4 / (4e-6 / 0.05) = 5000 -> cap
    T2s = torch.full((1, 5), 1e-4, dtype=f64)
    assert _ops.pick_ckpt_interval(dt, T1, T2s) == 10                          # 0.4 / 0.04
    assert _ops.pick_ckpt_interval(tensor([1e-5], dtype=f64), T1, T2s) == 4    # a NEW dt object: its own entry
    T2s[0, 3] = 2e-5                                                           # in-place edit bumps the version
    assert _ops.pick_ckpt_interval(dt, T1, T2s) == 2                           # 0.4 / (4e-6 / 2e-5)
    T2e = tensor(1e-3, dtype=f64).expand(1, 5)                                 # stride-0 storage as mobjs keeps it
    assert _ops.pick_ckpt_interval(dt, T1.expand(1, 5), T2e) == 64             # 100 -> cap 64
    _ops.note_host_max(dt, 8e-6)                                               # a registered host value wins over a read
    assert _ops.pick_ckpt_interval(dt, T1, T2s) == 1
    dt.mul_(1.0)                                                               # ... until the tensor changes
    assert _ops.pick_ckpt_interval(dt, T1, T2s) == 2


def test_reparam_falls_back_to_torch_when_constants_need_grad():
    """Gradients w.r.t. rfmax / smax / dt are not the kernel's business: those calls evaluate the reference's expression."""
    tr, th, ts = torch.randn(2, 1, 9, dtype=f64), torch.randn(2, 1, 9, dtype=f64), torch.randn(2, 3, 9, dtype=f64)
    rfmax = tensor([0.1, 0.2], dtype=f64, requires_grad=True)
    smax = tensor([[1., 2., 3.], [4., 5., 6.]], dtype=f64, requires_grad=True)
    rf, gr = utils.tρθts2rfgr(tr, th, ts, rfmax, smax, tensor(4e-6, dtype=f64))
    (rf.sum() + gr.sum()).backward()
    assert rfmax.grad is not None and smax.grad is not None
    want = tr.atan() / π * 2 * rfmax.detach()[:, None, None] * torch.cat((th.cos(), th.sin()), dim=1)
    assert torch.allclose(rf, want, rtol=1e-14, atol=0)

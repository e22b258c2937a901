"""Pin the CPU oracle (oracle/bloch_oracle.py) against the reference's golden vectors and against
outputs of the unmodified reference (tests/golden/*.npz, made by tests/golden/make_golden.py)."""
import numpy as np
import pytest
import torch

from oracle import bloch_oracle as orc

f64 = torch.float64


def mx(a, b):
    return float(np.abs(np.asarray(a) - np.asarray(b)).max())


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def test_kat3_constants(golden):
    """tests/test_slowsims.py:77-80 hard-coded Mo0 (atol 1e-9 upstream)."""
    g = golden('kat3')
    r = orc.applypulse_fwd_bwd(g['M0'], g['rf'], g['gr'], g['loc'], np.ones((1, 3, 3)), df=g['df'], b1=g['b1'],
                               T1=g['T1'], T2=g['T2'], gamma=g['gamma'], dt=g['dt'])
    assert mx(r['Mo'], g['Mo_const']) < 1e-12
    assert mx(r['Mo'], g['Mo_sims']) < 1e-13
    assert rel(r['grf'], g['grf']) < 1e-12 and rel(r['ggr'], g['ggr']) < 1e-12
    assert rel(r['grf'], g['grf_slow']) < 1e-10 and rel(r['ggr'], g['ggr_slow']) < 1e-10
    assert mx(r['gM0'], g['gM0']) < 1e-12
    assert mx(r['beff'][:, :, :4], g['beff_first']) < 1e-13 and mx(r['beff'][:, :, -4:], g['beff_last']) < 1e-13
    E1, E2 = np.exp(-g['dt'] / g['T1']), np.exp(-g['dt'] / g['T2'])
    A, B = orc.beff2ab(r['beff'], E1, E2, gamma=g['gamma'], dt=g['dt'])
    assert mx(A, g['A']) < 1e-12 and mx(B, g['B']) < 1e-12
    Mo_ab = (A @ torch.as_tensor(g['M0'])[..., None])[..., 0] + B    # slowsims.py:130
    assert mx(Mo_ab, g['Mo_const']) < 1e-12


@pytest.mark.parametrize('tag', ['relax', 'norelax'])
def test_sims512(golden, tag):
    """tests/test_sims.py:104-105,142-143: grad wrt M0 and Beff (atol 1e-9 upstream)."""
    g = golden('sims512')
    T1, T2 = (g['T1'], g['T2']) if tag == 'relax' else (None, None)
    r = orc.applypulse_fwd_bwd(g['M0'], g['rf'], g['gr'], g['loc'], np.ones_like(g['M0']), df=g['df'], b1=g['b1'],
                               T1=T1, T2=T2, gamma=g['gamma'], dt=g['dt'])
    assert mx(r['Mo'], g[f'Mo_{tag}']) < 1e-12
    assert mx(r['gM0'], g[f'gM0_{tag}']) < 1e-12
    assert mx(r['gBeff'][:, ::int(g['sub_step'])], g[f'gBeff_sub_{tag}']) < 1e-12
    assert rel(r['grf'], g[f'grf_{tag}']) < 1e-12 and rel(r['ggr'], g[f'ggr_{tag}']) < 1e-12


def test_cube27(golden):
    """tests/test_mobjs.py:112-120: SpinCube.applypulse goldens Mo0a / Mo0b on the compact spins."""
    g = golden('cube27')
    for T1, T2, key in ((g['T1_'], g['T2_'], 'M_relax'), (None, None, 'M_norelax')):
        beff = orc.rfgr2beff(g['rf'], g['gr'], g['loc_'], df=g['df_'], gamma=g['gamma_'])
        Mo = orc.blochsim_fwd(g['Mi_'], beff, T1, T2, g['gamma_'], g['dt'])
        full = np.full(g['mask'].shape + (3,), np.nan)
        full[g['mask']] = Mo[0].numpy()
        assert np.allclose(full, g[key], atol=1e-12, equal_nan=True)
    assert mx(g['M_relax'][0:1, 1, :, 1, :], g['Mo0a']) < 1e-9


@pytest.mark.parametrize('name', ['rand_mc', 'rand_nob1', 'rand_norelax'])
def test_random_chain(golden, name):
    g = golden(name)
    kw = dict(df=g.get('in_df'), b1=g.get('in_b1'), T1=g.get('in_T1'), T2=g.get('in_T2'), gamma=g['in_gam'],
              dt=g['in_dt'])
    r = orc.applypulse_fwd_bwd(g['in_M0'], g['in_rf'], g['in_gr'], g['in_loc'], g['in_w'], **kw)
    assert mx(r['Mo'], g['Mo_f64']) < 1e-12
    assert mx(r['beff'], g['beff_f64']) < 1e-11
    assert rel(r['gBeff'], g['gBeff_f64']) < 1e-12
    assert rel(r['grf'], g['grf_f64']) < 1e-12 and rel(r['ggr'], g['ggr_f64']) < 1e-12
    assert rel(r['gM0'], g['gM0_slow_f64']) < 1e-11          # sims.py:267 is broken for per-spin gamma
    if 'A_f64' in g:
        dtb = np.asarray(g['in_dt'], dtype=np.float64).reshape(-1, 1)
        A, B = orc.beff2ab(r['beff'], np.exp(-dtb / g['in_T1']), np.exp(-dtb / g['in_T2']),
                           gamma=g['in_gam'], dt=g['in_dt'])
        assert mx(A, g['A_f64']) < 1e-12 and mx(B, g['B_f64']) < 1e-12
    # the oracle evaluated in fp32 sits inside the reference's own fp32 noise band
    r32 = orc.applypulse_fwd_bwd(g['in_M0'], g['in_rf'], g['in_gr'], g['in_loc'], g['in_w'], dtype=torch.float32, **kw)
    floor = mx(g['Mo_f32'], g['Mo_f64'])
    assert mx(r32['Mo'], g['Mo_f64']) < 3 * floor + 1e-5


def test_bench8(golden):
    g = golden('bench8')
    beff = orc.rfgr2beff(g['in_rf'], g['in_gr'], g['in_loc'], df=g['in_df'], b1=g['in_b1'], gamma=g['in_gam'])
    Mo = orc.blochsim_fwd(g['in_M0'], beff, g['in_T1'], g['in_T2'], g['in_gam'], g['in_dt'])
    assert mx(Mo, g['Mo_f64']) < 1e-12
    gMo = 2 * (Mo - torch.tensor([0., 1., 0.], dtype=f64))
    r = orc.applypulse_fwd_bwd(g['in_M0'], g['in_rf'], g['in_gr'], g['in_loc'], gMo, df=g['in_df'], b1=g['in_b1'],
                               T1=g['in_T1'], T2=g['in_T2'], gamma=g['in_gam'], dt=g['in_dt'])
    assert rel(r['grf'], g['grf_f64']) < 1e-11 and rel(r['ggr'], g['ggr_f64']) < 1e-11


def test_freeprec(golden):
    g = golden('freeprec')
    Mo = orc.freeprec(g['a_Mi'], g['a_dur'], g['a_T1'], g['a_T2'], g['a_df'])
    assert mx(Mo, [[[0., -0.5, 0.5], [-0.5, 0, 0.5], [0., 0., 1.]]]) < 1e-12     # test_slowsims.py:117
    Mo = orc.freeprec(g['b_Mi'], g['b_dur'], g['b_T1'], g['b_T2'], g['b_df'])
    assert mx(Mo, g['b_Mo']) < 1e-13


def test_interp(golden):
    g = golden('interp')
    a_rf = orc.interp_linear(g['a_rf'], 4e-6, 4e-6 * 5)
    assert mx(a_rf, [[[0.04, 0.09], [0.06, 0.01]]]) < 1e-12                        # test_mobjs.py:185-187
    assert mx(orc.interp_linear(g['a_gr'], 4e-6, 2e-5), g['a_gr_new']) < 1e-12
    b = orc.interp_linear(g['b_rf'], 4e-6, 2e-6)
    assert b.shape[2] == 19 and mx(b, g['b_rf_new']) < 1e-12                     # float // quirk
    assert mx(orc.interp_linear(g['b_gr'], 4e-6, 2e-6), g['b_gr_new']) < 1e-12
    assert orc.interp_grid(40, 20e-6, 4e-6)[1].shape[0] == int(g['c_nT'][0]) == 200
    assert orc.interp_grid(40, float(np.float32(20e-6)), 4e-6)[1].shape[0] == int(g['c_nT'][1]) == 199


@pytest.mark.parametrize('tag', ['sc', 'mc'])
def test_reparam(golden, tag):
    """utils.tρθ2rf / lρθ2rf / ts2s / s2g and their autograd, as the unmodified reference computes them."""
    g = {k[len(tag) + 1:]: v for k, v in golden('reparam').items() if k.startswith(tag + '_')}
    for kind, logit in (('t', False), ('l', True)):
        rf, s, gr = orc.reparam_fwd(g['rho'], g['theta'], g['rfmax'], g['ts'], g['smax'], g['dt'], logit=logit)
        assert rel(rf, g['rf_' + kind]) < 1e-14 and rel(s, g['s']) < 1e-14 and rel(gr, g['gr']) < 1e-13
        grho, gtheta, gts = orc.reparam_adj(g['wrf'], g['wg'], g['rho'], g['theta'], g['rfmax'], g['ts'], g['smax'], g['dt'],
                                            logit=logit)
        assert rel(grho, g['grho_' + kind]) < 1e-13 and rel(gtheta, g['gtheta_' + kind]) < 1e-13
        assert rel(gts, g['gts']) < 1e-13


def test_mask_plumbing(golden):
    g = golden('reparam')
    emb = orc.mask_embed(g['mask_v_'], g['mask'])
    assert np.array_equal(np.isnan(emb), np.isnan(g['mask_embedded']))
    assert np.array_equal(np.nan_to_num(emb), np.nan_to_num(g['mask_embedded']))
    assert np.array_equal(orc.mask_extract(g['mask_full'], g['mask']), g['mask_extracted'])

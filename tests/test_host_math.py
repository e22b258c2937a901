"""The per-step formulation the CUDA kernels inline (csrc/bloch_math.cuh), compiled for the host and
checked against the oracle and the reference fixtures.  No GPU needed; the product never runs this."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

from oracle import bloch_oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope='module')
def lib(tmp_path_factory):
    out = str(tmp_path_factory.mktemp('hm') / 'libhostmath.so')
    subprocess.check_call(['g++', '-O2', '-std=c++17', '-shared', '-fPIC', '-ffp-contract=off',
                           os.path.join(HERE, 'host_math_harness.cpp'), '-o', out])
    return ctypes.CDLL(out)


def ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def run(lib, dtype, pol, g, K, relax=True):
    """g: fixture dict with in_* keys for a single-coil N=1 case."""
    T = np.float64 if dtype == 'f64' else np.float32
    M0 = np.ascontiguousarray(g['in_M0'][0], dtype=T)
    nM = M0.shape[0]
    rf = np.ascontiguousarray(g['in_rf'][0].reshape(2, -1), dtype=T)
    nT = rf.shape[1]
    gr = np.ascontiguousarray(g['in_gr'][0], dtype=T)
    loc = np.ascontiguousarray(g['in_loc'][0], dtype=T)
    b1 = np.ascontiguousarray(g['in_b1'][0].reshape(nM, 2), dtype=T) if 'in_b1' in g else None
    bc = lambda k, d: (np.ascontiguousarray(np.broadcast_to(np.asarray(g[k], dtype=np.float64).reshape(-1)[:nM] if
                       np.asarray(g[k]).size >= nM else np.asarray(g[k], dtype=np.float64).reshape(-1)[:1], (nM,)))
                       if k in g else d)
    df = bc('in_df', None)
    T1, T2 = bc('in_T1', None), bc('in_T2', None)
    gam = bc('in_gam', None)
    gMo = np.ascontiguousarray(g['in_w'][0], dtype=T)
    Mo, gM0 = np.zeros((nM, 3), T), np.zeros((nM, 3), T)
    grf, ggr = np.zeros((2, nT)), np.zeros((3, nT))
    err = np.zeros(1, T)
    fn = {'f32': lib.host_sim_f32, 'f32x2': lib.host_sim_f32x2, 'f64': lib.host_sim_f64}[dtype]
    fn(ctypes.c_int(pol), ctypes.c_int(int(relax and T1 is not None)), ctypes.c_int(nM), ctypes.c_int(nT),
       ctypes.c_int(K), ptr(M0), ptr(rf), ptr(gr), ptr(loc), ptr(b1), ptr(df), ptr(T1), ptr(T2), ptr(gam),
       ctypes.c_double(float(np.asarray(g['in_dt']).reshape(-1)[0])), ptr(gMo), ptr(Mo), ptr(gM0), ptr(grf), ptr(ggr),
       ptr(err))
    return Mo, gM0, grf, ggr, float(err[0])


def rel(a, b):
    return float(np.linalg.norm(np.asarray(a, np.float64).ravel() - np.asarray(b, np.float64).ravel())
                 / np.linalg.norm(np.asarray(b, np.float64).ravel()))


def test_precise_sincos_rsqrt(lib):
    x = np.concatenate([np.linspace(0, 20, 200001), np.logspace(-12, 3, 20001)]).astype(np.float32)
    s, c = np.zeros_like(x), np.zeros_like(x)
    lib.host_sincos_f32(ctypes.c_int(x.size), ptr(x), ptr(s), ptr(c))
    xd = x.astype(np.float64)
    assert np.abs(s - np.sin(xd)).max() < 1.5e-7 and np.abs(c - np.cos(xd)).max() < 1.5e-7
    xd = np.concatenate([np.linspace(0, 50, 400001), np.logspace(-12, 5, 40001)])
    sd, cd = np.zeros_like(xd), np.zeros_like(xd)
    lib.host_sincos_f64(ctypes.c_int(xd.size), ptr(xd), ptr(sd), ptr(cd))
    xl = xd.astype(np.longdouble)
    assert np.abs(sd - np.sin(xl)).max() < 3e-16 and np.abs(cd - np.cos(xl)).max() < 3e-16
    y = np.logspace(-24, 6, 100001).astype(np.float32)
    r = np.zeros_like(y)
    lib.host_rsq_f32(ctypes.c_int(y.size), ptr(y), ptr(r))
    assert np.abs(r * np.sqrt(y.astype(np.float64)) - 1).max() < 1.2e-7


@pytest.mark.parametrize('K', [1, 16, 64, 4096])
def test_f64_formulation_matches_reference(lib, golden, K):
    """fp64: forward <=1e-12 on M, gradients ~1e-11 relative, with and without resync."""
    for name in ('rand_norelax', 'bench8'):
        g = golden(name)
        if name == 'bench8':
            Mo64 = g['Mo_f64']
            g = dict(g, in_w=2 * (Mo64 - np.array([0., 1., 0.])))
        Mo, gM0, grf, ggr, err = run(lib, 'f64', 0, g, K)
        assert np.abs(Mo - g['Mo_f64'][0]).max() < 1e-12
        assert rel(grf, g['grf_f64'][0].reshape(2, -1)) < 1e-10 and rel(ggr, g['ggr_f64'][0]) < 1e-10
        assert err < 1e-11
        if 'gM0_slow_f64' in g:
            assert rel(gM0, g['gM0_slow_f64'][0]) < 1e-10


@pytest.mark.parametrize('pol', [0, 1])
def test_f32_formulation_noise_floor(lib, golden, pol):
    """fp32: M error vs the fp64 reference on the same inputs stays inside the reference's own
    fp32-vs-fp64 band; gradients well inside 1e-4 relative (north_star tolerance)."""
    g = golden('bench8')
    g = dict(g, in_w=2 * (g['Mo_f64'] - np.array([0., 1., 0.])))
    floor = np.abs(g['Mo_f32'] - g['Mo_f64']).max()
    for K in (16, 64):
        Mo, gM0, grf, ggr, err = run(lib, 'f32', pol, g, K)
        dM = np.abs(Mo - g['Mo_f64'][0]).max()
        print(f'pol={pol} K={K}: max|dM|={dM:.2e} (ref32 floor {floor:.2e}) resync drift={err:.2e} '
              f'grf rel={rel(grf, g["grf_f64"][0].reshape(2, -1)):.2e} ggr rel={rel(ggr, g["ggr_f64"][0]):.2e}')
        assert dM < max(1e-5, 1.5 * floor)
        assert rel(grf, g['grf_f64'][0].reshape(2, -1)) < 1e-4 and rel(ggr, g['ggr_f64'][0]) < 1e-4


@pytest.mark.parametrize('pol', [0, 1])
def test_packed_f2_path_equals_scalar_path(lib, golden, pol):
    """The two-spins-per-thread formulation (FFMA2 on the device) is the same arithmetic as the scalar one.  The only
    difference: a step with |b| > 2 pi sends BOTH spins of a pair through the reduce-by-pi coefficients (one branch per
    f2), so pairs that contain such a step differ from the scalar path by rounding."""
    g = golden('bench8')
    g = dict(g, in_w=2 * (g['Mo_f64'] - np.array([0., 1., 0.])))
    a = run(lib, 'f32', pol, g, 64)
    b = run(lib, 'f32x2', pol, g, 64)
    same = np.all(a[0] == b[0], axis=1) & np.all(a[1] == b[1], axis=1)
    assert same.mean() > (0.9 if pol == 1 else 0.9999), same.mean()
    assert np.abs(a[0] - b[0]).max() < 5e-6 and rel(b[1], a[1]) < 2e-5
    assert rel(b[2], a[2]) < 2e-6 and rel(b[3], a[3]) < 2e-6
    g2 = golden('rand_norelax')
    a, b = run(lib, 'f32', pol, g2, 16), run(lib, 'f32x2', pol, g2, 16)
    assert np.abs(a[0][:32] - b[0][:32]).max() < 5e-6 and np.abs(a[1][:32] - b[1][:32]).max() < 5e-6 * np.abs(a[1]).max()


def test_kat3_through_formulation(lib, golden):
    g = golden('kat3')
    gg = dict(in_M0=g['M0'], in_rf=g['rf'], in_gr=g['gr'], in_loc=g['loc'], in_df=g['df'],
              in_b1=np.broadcast_to(g['b1'], (1, 3, 2, 1)), in_T1=g['T1'], in_T2=g['T2'], in_gam=g['gamma'],
              in_dt=g['dt'], in_w=np.ones((1, 3, 3)))
    Mo, gM0, grf, ggr, err = run(lib, 'f64', 0, gg, 32)
    assert np.abs(Mo - g['Mo_const'][0]).max() < 1e-12     # tests/test_slowsims.py:77-80
    assert rel(grf, g['grf'][0].reshape(2, -1)) < 1e-10 and rel(ggr, g['ggr'][0]) < 1e-10
    assert np.abs(gM0 - g['gM0'][0]).max() < 1e-11

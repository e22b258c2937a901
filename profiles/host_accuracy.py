#!/usr/bin/env python
"""fp32 accuracy of a formulation of csrc/bloch_math.cuh, without a GPU: compiles tests/host_math_harness.cpp (the very
step functions the kernels inline) with the given -D flags and runs bench-distribution problems of several lengths,
comparing the magnetisation and the waveform gradients with the harness' own fp64 run on the same (fp32-rounded) inputs.

    python profiles/host_accuracy.py [-DMRPHY_HALF_ANGLE=0] ...
"""
import ctypes
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from test_host_math import ptr  # noqa: E402


def build(defs):
    out = os.path.join(tempfile.mkdtemp(), 'libhm.so')
    subprocess.check_call(['g++', '-O2', '-std=c++17', '-shared', '-fPIC', '-ffp-contract=off'] + defs +
                          [os.path.join(ROOT, 'tests', 'host_math_harness.cpp'), '-o', out])
    return ctypes.CDLL(out)


def problem(n, nT, seed=0, gscale=2.0):
    rng = np.random.default_rng(seed)
    U = lambda *s: rng.uniform(-1, 1, s)
    ax = (np.arange(n) - n // 2) / n * 24.0
    loc = np.stack(np.meshgrid(ax, ax, ax, indexing='ij'), -1).reshape(-1, 3)
    nM = loc.shape[0]
    return dict(M0=np.tile([0., 0., 1.], (nM, 1)), rf=U(2, nT) * 0.1, gr=U(3, nT) * gscale, loc=loc, b1=np.stack([1 + U(nM) * .1, U(nM) * .1], -1),
                df=U(nM) * 200, T1=np.full(nM, 1.47), T2=np.full(nM, 0.07), gam=np.full(nM, 4257.6), dt=4e-6)


def run(lib, kind, pol, p, K=64):
    T = np.float64 if kind == 'f64' else np.float32
    c = lambda a: np.ascontiguousarray(a, dtype=T)
    d = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    M0, rf, gr, loc, b1 = (c(np.asarray(p[k], np.float32)) for k in ('M0', 'rf', 'gr', 'loc', 'b1'))   # fp32-rounded inputs
    nM, nT = M0.shape[0], rf.shape[1]
    gMo = c(2 * (p['Mo64'] - np.array([0., 1., 0.]))) if 'Mo64' in p else c(np.ones((nM, 3)))
    Mo, gM0, grf, ggr, err = np.zeros((nM, 3), T), np.zeros((nM, 3), T), np.zeros((2, nT)), np.zeros((3, nT)), np.zeros(1, T)
    fn = {'f32': lib.host_sim_f32, 'f32x2': lib.host_sim_f32x2, 'f64': lib.host_sim_f64}[kind]
    df, T1, T2, gam = (d(np.asarray(p[k], np.float32)) for k in ('df', 'T1', 'T2', 'gam'))
    fn(ctypes.c_int(pol), ctypes.c_int(1), ctypes.c_int(nM), ctypes.c_int(nT), ctypes.c_int(K), ptr(M0), ptr(rf), ptr(gr),
       ptr(loc), ptr(b1), ptr(df), ptr(T1), ptr(T2), ptr(gam), ctypes.c_double(float(np.float32(p['dt']))), ptr(gMo),
       ptr(Mo), ptr(gM0), ptr(grf), ptr(ggr), ptr(err))
    return Mo.astype(np.float64), grf, ggr


if __name__ == '__main__':
    defs = [a for a in sys.argv[1:] if a.startswith('-D')]
    lib = build(defs)
    rel = lambda a, b: float(np.linalg.norm(a - b) / np.linalg.norm(b))
    print('flags:', ' '.join(defs) or '(default)')
    for n, nT, gs in ((8, 1000, 2.0), (8, 4000, 2.0), (10, 2000, 2.0), (8, 1000, 0.5), (8, 4000, 4.0)):
        p = problem(n, nT, gscale=gs)
        Mo64, _, _ = run(lib, 'f64', 0, p)
        p['Mo64'] = Mo64
        Mo64, grf64, ggr64 = run(lib, 'f64', 0, p)
        for kind in ('f32x2',):
            Mo, grf, ggr = run(lib, kind, 1, p)
            e = np.abs(Mo - Mo64)
            print(f'  {n}^3 x {nT}, |gr| <= {gs}: {kind} max|dM| {e.max():.2e} rms {np.sqrt((e ** 2).mean()):.2e} '
                  f'| |M| drift mean {np.mean(np.linalg.norm(Mo, axis=1) - np.linalg.norm(Mo64, axis=1)):+.2e} '
                  f'| grf rel {rel(grf, grf64):.2e} ggr rel {rel(ggr, ggr64):.2e}')

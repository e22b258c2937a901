#!/usr/bin/env python
"""Coefficients for the rsqrt-free rotation coefficients of csrc/bloch_math.cuh (policy TRIG_HALF).

The Rodrigues step only needs  a = sin(phi)/phi  and  d = (1 - cos(phi))/phi^2  as functions of p2 = phi^2 = |b|^2.
With the half angle h = phi/2:   a = sinc(h) * cos(h),   d = sinc(h)^2 / 2,   and sinc(h), cos(h) are EVEN entire
functions of h, i.e. short polynomials in p2 on phi <= 2 pi -- no rsqrt, no Newton step, no range reduction, no sign
fix-up.  With A(p2) = sinc(h) and C(p2) = cos(h):   a = A*C,  d = (A/2)*A   (three multiplies; folding the 1/2 into the
polynomials as sqrt(2) factors makes their leading coefficients inexact -- a radius bias of 7e-8 per step).

    python profiles/fit_halfangle.py

Fits both polynomials (weighted least squares iterated towards minimax, in float64), rounds to fp32, emulates the fp32
FMA evaluation exactly and prints the angle / radius error statistics of the resulting rotation against the same
statistics of the reduce-by-pi polynomial path (profiles/fit_sincos.py).  A bias adds up linearly over the time steps,
noise as a random walk, so means matter more than maxima.
"""
import numpy as np

f32 = np.float32
PHIMAX = 2 * np.pi * 1.02          # fit a little beyond 2 pi so the switch-over point is inside the fit


def fma(a, b, c):
    return (np.asarray(a, np.float64) * np.asarray(b, np.float64) + np.asarray(c, np.float64)).astype(f32)


def fit(f, zmax, n):
    """minimax-ish fit of f(z) = 1 + z * poly_{n-1}(z) on [0, zmax]: the leading coefficient is EXACTLY 1 (a rounded
    leading coefficient is a systematic scale error of the rotation, i.e. a radius bias); fitted in t = z / zmax.
    Coefficients are rounded to fp32 ONE AT A TIME, lowest order first, and the remaining ones refitted each time, so
    the higher orders absorb the rounding error of the lower ones (at p2 = 39 the rounding of -1/24 alone is 5e-8)."""
    t = np.cos(np.pi * (np.arange(6000) + 0.5) / 6000) * 0.5 + 0.5
    z = t * zmax
    scale = zmax ** np.arange(n - 1)
    Vfull, y = np.vander(t, n - 1, increasing=True) * z[:, None], f(z) - 1.0
    fixed = []
    for k in range(n - 1):
        V = Vfull[:, k:]
        yk = y - (Vfull[:, :k] @ (np.array(fixed) * scale[:k]) if k else 0.0)
        w = np.ones(6000)
        c = np.linalg.lstsq(V, yk, rcond=None)[0]
        for _ in range(60):
            e = V @ c - yk
            w *= 1 + 4 * np.abs(e) / np.abs(e).max()
            w /= w.mean()
            c = np.linalg.lstsq(V * w[:, None], yk * w, rcond=None)[0]
        fixed.append(float(f32(c[0] / scale[k])))
    e = Vfull @ (np.array(fixed) * scale) - y
    print('    fit error with fp32 coefficients: max %.2e mean %+.2e' % (np.abs(e).max(), e.mean()))
    return np.concatenate([[1.0], fixed])


def horner(p2, c):
    r = np.full_like(p2, f32(c[-1]))
    for k in c[-2::-1]:
        r = fma(r, p2, f32(k))
    return r


def half_coef(p2, ca, cc):
    """fp32 emulation of the TRIG_HALF coefficients -> (a, d) as fp32"""
    A, C = horner(p2, ca), horner(p2, cc)
    return (A * C).astype(f32), ((A * f32(0.5)).astype(f32) * A).astype(f32)


def modpi_coef(p2):
    """fp32 emulation of the shipped path: rsqrt (correctly rounded stand-in for MUFU + Newton), reduce by pi, polynomials"""
    cs = [f32(v) for v in (-1.666666716e-01, 8.333338425e-03, -1.984213741e-04, 2.760374173e-06, -2.543108479e-08)]
    cc = [f32(v) for v in (-0.5, 4.166666046e-02, -1.388868783e-03, 2.477660746e-05, -2.633455551e-07)]
    rs = (1.0 / np.sqrt(p2.astype(np.float64))).astype(f32)
    x = (p2 * rs).astype(f32)
    MAGIC = f32(12582912.0)
    t = fma(x, f32(0.31830988618379067), MAGIC)
    jf = (t - MAGIC).astype(f32)
    r = fma(jf, f32(-3.1415920257568359375), x)
    r = fma(jf, f32(-6.2783295107151866e-07), r)
    r2 = (r * r).astype(f32)
    sp = horner(r2, cs)
    sr = fma((sp * r2).astype(f32), r, r)
    cp = horner(r2, cc)
    cr = fma(cp, r2, f32(1.0))
    sg = np.where(t.view(np.int32) & 1, f32(-1), f32(1))
    s, c = sr * sg, cr * sg
    rs2 = (rs * rs).astype(f32)
    return (s * rs).astype(f32), fma(-c, rs2, rs2)


def stats(p2, a, d):
    """errors of the rotation R = I - d (p2 I - b b^T) - a [b]x the kernels apply: sin = phi a, cos = 1 - p2 d"""
    p = p2.astype(np.float64)
    phi = np.sqrt(p)
    s, c = phi * a.astype(np.float64), 1 - p * d.astype(np.float64)
    ang, rad = s * np.cos(phi) - c * np.sin(phi), s * s + c * c - 1
    return ang.mean(), np.sqrt((ang ** 2).mean()), np.abs(ang).max(), rad.mean(), np.sqrt((rad ** 2).mean()), np.abs(rad).max()


FMT = 'angle mean %+.2e rms %.2e max %.2e | radius mean %+.2e rms %.2e max %.2e'

if __name__ == '__main__':
    zmax = PHIMAX ** 2
    sinc_h = lambda z: np.sinc(np.sqrt(z) / 2 / np.pi)
    cos_h = lambda z: np.cos(np.sqrt(z) / 2)
    rng = np.random.default_rng(0)
    sets = {'phi ~ U(0, 2pi)': rng.uniform(0, 2 * np.pi, 2000000),
            'phi ~ |N(0, 1.5)| (bench)': np.abs(rng.normal(0, 1.5, 2000000)),
            'phi ~ U(0, 0.3)': rng.uniform(0, 0.3, 2000000)}
    sets = {k: (v[v < 2 * np.pi] ** 2).astype(f32) for k, v in sets.items()}
    for na, nc in ((7, 7), (7, 8), (8, 8)):
        ca, cc = fit(sinc_h, zmax, na).astype(f32), fit(cos_h, zmax, nc).astype(f32)
        print(f'--- TRIG_HALF, {na} + {nc} coefficients')
        for name, p2 in sets.items():
            print(f'  {name:28s}: ' + FMT % stats(p2, *half_coef(p2, ca, cc)))
        print('  A:', ', '.join('%.9ef' % v for v in ca))
        print('  C:', ', '.join('%.9ef' % v for v in cc))
    print('--- shipped reduce-by-pi path (rsqrt correctly rounded)')
    for name, p2 in sets.items():
        print(f'  {name:28s}: ' + FMT % stats(p2, *modpi_coef(p2)))

#!/usr/bin/env python
"""Coefficients of the fp32 sincos polynomials on [-pi/2, pi/2] used by csrc/bloch_math.cuh (MRPHY_SC_MODPI): weighted
least-squares fit iterated towards minimax, then a search over neighbouring fp32 values of two coefficients for the pair
with zero MEAN radius and angle error of the whole routine (reduction included) -- a bias adds up linearly over the time
steps of a simulation, noise only as a random walk.  Emulates fp32 FMA exactly (products and sums in float64, one rounding).

    python profiles/fit_sincos.py
"""
import numpy as np

f32 = np.float32
MAGIC = f32(12582912.0)


def fma(a, b, c):
    return (np.asarray(a, np.float64) * np.asarray(b, np.float64) + np.asarray(c, np.float64)).astype(f32)


def fit(f, a, n):
    z = np.cos(np.pi * (np.arange(4000) + 0.5) / 4000) * 0.5 * a * a + 0.5 * a * a
    V, y, w = np.vander(z, n, increasing=True), f(z), np.ones(4000)
    c = np.linalg.lstsq(V, y, rcond=None)[0]
    for _ in range(60):
        e = V @ c - y
        w *= 1 + 4 * np.abs(e) / np.abs(e).max()
        w /= w.mean()
        c = np.linalg.lstsq(V * w[:, None], y * w, rcond=None)[0]
    return c


def sincos(x, cs, cc):
    x = x.astype(f32)
    t = fma(x, f32(0.31830988618379067), MAGIC)
    jf = (t - MAGIC).astype(f32)
    r = fma(jf, f32(-3.1415920257568359375), x)
    r = fma(jf, f32(-6.2783295107151866e-07), r)
    r2 = (r * r).astype(f32)
    sp = np.full_like(r, f32(cs[-1]))
    for k in cs[-2::-1]:
        sp = fma(sp, r2, f32(k))
    sr = fma((sp * r2).astype(f32), r, r)
    cp = np.full_like(r, f32(cc[-1]))
    for k in cc[-2::-1]:
        cp = fma(cp, r2, f32(k))
    cr = fma(cp, r2, f32(1.0))
    sg = np.where(t.view(np.int32) & 1, f32(-1), f32(1))
    return sr * sg, cr * sg


def stats(x, cs, cc):
    s, c = sincos(x, cs, cc)
    xd, s, c = x.astype(f32).astype(np.float64), s.astype(np.float64), c.astype(np.float64)
    ang, rad = s * np.cos(xd) - c * np.sin(xd), s * s + c * c - 1
    return ang.mean(), np.sqrt((ang ** 2).mean()), np.abs(ang).max(), rad.mean(), np.sqrt((rad ** 2).mean())


if __name__ == '__main__':
    a = np.pi / 2
    cs = fit(lambda z: (np.sin(np.sqrt(z)) / np.sqrt(z) - 1) / z, a, 5).astype(f32)
    cc = fit(lambda z: (np.cos(np.sqrt(z)) - 1) / z, a, 5).astype(f32)
    x = np.linspace(0, 4 * np.pi, 1200001)[1:]
    print('fit            : angle mean %+.2e rms %.2e max %.2e | radius mean %+.2e rms %.2e' % stats(x, cs, cc))
    best = None
    for ds in range(-6, 7):
        for dc in range(-8, 9):
            s2, c2 = cs.copy(), cc.copy()
            s2[0] = np.nextafter(s2[0], f32(np.inf if ds > 0 else -np.inf)) if ds else s2[0]
            for _ in range(abs(ds) - 1):
                s2[0] = np.nextafter(s2[0], f32(np.inf if ds > 0 else -np.inf))
            for _ in range(abs(dc)):
                c2[1] = np.nextafter(c2[1], f32(np.inf if dc > 0 else -np.inf))
            st = stats(x[::4], s2, c2)
            score = abs(st[3]) + abs(st[0]) + 0.05 * st[2]
            if best is None or score < best[0]:
                best = (score, ds, dc, s2, c2)
    _, ds, dc, cs, cc = best
    print('tuned (%+d, %+d) : angle mean %+.2e rms %.2e max %.2e | radius mean %+.2e rms %.2e' % ((ds, dc) + stats(x, cs, cc)))
    print('sin:', ', '.join('%.9ef' % v for v in cs))
    print('cos:', ', '.join('%.9ef' % v for v in cc))

// Host-side check of csrc/bloch_math.cuh (TEST INFRASTRUCTURE).  Compiles the very same step
// functions the CUDA kernels inline, runs them spin by spin on the CPU with the kernels' loop
// structure (checkpoint every K steps, time-reversed state reconstruction in the backward
// sweep), so tests can compare the formulation with the oracle without a GPU.
#include <vector>
#include <cstring>
#include "../mrphy.py_b200/csrc/bloch_math.cuh"

using namespace mrphy;

template <typename T, int POL, bool RELAX>
static void run(int nM, int nT, int K, const T* M0, const T* rf /*[2][nT]*/, const T* gr /*[3][nT]*/,
                const T* loc, const T* b1 /*[nM][2] or null*/, const double* df, const double* T1, const double* T2,
                const double* gamma, double dt, const T* gMo, T* Mo, T* gM0, double* grf, double* ggr, T* resync_err) {
  for (int t = 0; t < nT; ++t) { grf[t] = grf[nT + t] = 0; ggr[t] = ggr[nT + t] = ggr[2 * nT + t] = 0; }
  T maxerr = 0;
  for (int i = 0; i < nM; ++i) {
    SpinConst<T, 1> k;
    T br = b1 ? b1[2 * i] : (T)1, bi = b1 ? b1[2 * i + 1] : (T)0;
    make_consts<T, 1>(k, gamma[i], dt, RELAX, RELAX ? T1[i] : 1.0, RELAX ? T2[i] : 1.0, df ? df[i] : 0.0,
                      loc[3 * i], loc[3 * i + 1], loc[3 * i + 2], &br, &bi);
    T mx = M0[3 * i], my = M0[3 * i + 1], mz = M0[3 * i + 2];
    to_frame(k, mx, my);
    std::vector<T> ck;
    for (int t = 0; t < nT; ++t) {
      if (t > 0 && t % K == 0) { ck.push_back(mx); ck.push_back(my); ck.push_back(mz); }
      T bx, by, bz;
      field<T, 1>(k, &rf[t], &rf[nT + t], gr[t], gr[nT + t], gr[2 * nT + t], bx, by, bz);
      step_fwd<T, POL, RELAX>(bx, by, bz, k.e1, k.e2, mx, my, mz);
    }
    { T ox = mx, oy = my; from_frame(k, ox, oy); Mo[3 * i] = ox; Mo[3 * i + 1] = oy; Mo[3 * i + 2] = mz; }
    T hx = gMo[3 * i], hy = gMo[3 * i + 1], hz = gMo[3 * i + 2];
    to_frame(k, hx, hy);
    for (int t = nT - 1; t >= 0; --t) {
      T bx, by, bz, Fx, Fy, Fz;
      field<T, 1>(k, &rf[t], &rf[nT + t], gr[t], gr[nT + t], gr[2 * nT + t], bx, by, bz);
      step_bwd<T, POL, RELAX, 1>(k, bx, by, bz, mx, my, mz, hx, hy, hz, Fx, Fy, Fz);
      T rgx, rgy;
      rf_chain(k, 0, Fx, Fy, rgx, rgy);
      grf[t] -= (double)rgx;
      grf[nT + t] -= (double)rgy;
      ggr[t] -= (double)(k.glx * Fz);
      ggr[nT + t] -= (double)(k.gly * Fz);
      ggr[2 * nT + t] -= (double)(k.glz * Fz);
      if (t > 0 && t % K == 0) {
        int c = t / K - 1;
        T ex = fabs(mx - ck[3 * c]), ey = fabs(my - ck[3 * c + 1]), ez = fabs(mz - ck[3 * c + 2]);
        T e = ex > ey ? (ex > ez ? ex : ez) : (ey > ez ? ey : ez);
        if (e > maxerr) maxerr = e;
        mx = ck[3 * c]; my = ck[3 * c + 1]; mz = ck[3 * c + 2];
      }
    }
    from_frame(k, hx, hy);
    gM0[3 * i] = hx; gM0[3 * i + 1] = hy; gM0[3 * i + 2] = hz;
  }
  *resync_err = maxerr;
}

// packed variant: spins (i, i+1) ride in the two halves of an f2 (nM must be even)
template <int POL, bool RELAX>
static void run2(int nM, int nT, int K, const float* M0, const float* rf, const float* gr, const float* loc,
                 const float* b1, const double* df, const double* T1, const double* T2, const double* gamma, double dt,
                 const float* gMo, float* Mo, float* gM0, double* grf, double* ggr, float* resync_err) {
  for (int t = 0; t < nT; ++t) { grf[t] = grf[nT + t] = 0; ggr[t] = ggr[nT + t] = ggr[2 * nT + t] = 0; }
  float maxerr = 0;
  for (int i = 0; i + 1 < nM; i += 2) {
    SpinConst<float, 1> ks[2];
    for (int q = 0; q < 2; ++q) {
      const int ii = i + q;
      float br = b1 ? b1[2 * ii] : 1.f, bi = b1 ? b1[2 * ii + 1] : 0.f;
      make_consts<float, 1>(ks[q], gamma[ii], dt, RELAX, RELAX ? T1[ii] : 1.0, RELAX ? T2[ii] : 1.0, df ? df[ii] : 0.0,
                            loc[3 * ii], loc[3 * ii + 1], loc[3 * ii + 2], &br, &bi);
    }
    const SpinConst<f2, 1> k = pack2<1>(ks[0], ks[1]);
    f2 mx(M0[3 * i], M0[3 * i + 3]), my(M0[3 * i + 1], M0[3 * i + 4]), mz(M0[3 * i + 2], M0[3 * i + 5]);
    to_frame(k, mx, my);
    std::vector<f2> ck;
    for (int t = 0; t < nT; ++t) {
      if (t > 0 && t % K == 0) { ck.push_back(mx); ck.push_back(my); ck.push_back(mz); }
      f2 bx, by, bz, rx(rf[t]), ry(rf[nT + t]);
      field<f2, 1>(k, &rx, &ry, f2(gr[t]), f2(gr[nT + t]), f2(gr[2 * nT + t]), bx, by, bz);
      step_fwd<f2, POL, RELAX>(bx, by, bz, k.e1, k.e2, mx, my, mz);
    }
    {
      f2 ox = mx, oy = my;
      from_frame(k, ox, oy);
      Mo[3 * i] = ox.v.x; Mo[3 * i + 1] = oy.v.x; Mo[3 * i + 2] = mz.v.x;
      Mo[3 * i + 3] = ox.v.y; Mo[3 * i + 4] = oy.v.y; Mo[3 * i + 5] = mz.v.y;
    }
    f2 hx(gMo[3 * i], gMo[3 * i + 3]), hy(gMo[3 * i + 1], gMo[3 * i + 4]), hz(gMo[3 * i + 2], gMo[3 * i + 5]);
    to_frame(k, hx, hy);
    for (int t = nT - 1; t >= 0; --t) {
      f2 bx, by, bz, Fx, Fy, Fz, rx(rf[t]), ry(rf[nT + t]);
      field<f2, 1>(k, &rx, &ry, f2(gr[t]), f2(gr[nT + t]), f2(gr[2 * nT + t]), bx, by, bz);
      step_bwd<f2, POL, RELAX, 1>(k, bx, by, bz, mx, my, mz, hx, hy, hz, Fx, Fy, Fz);
      f2 rgx, rgy;
      rf_chain(k, 0, Fx, Fy, rgx, rgy);
      grf[t] -= (double)hsum(rgx);
      grf[nT + t] -= (double)hsum(rgy);
      ggr[t] -= (double)hsum(k.glx * Fz);
      ggr[nT + t] -= (double)hsum(k.gly * Fz);
      ggr[2 * nT + t] -= (double)hsum(k.glz * Fz);
      if (t > 0 && t % K == 0) {
        int c = t / K - 1;
        float e = fmaxf(fmaxf(fabsf(mx.v.x - ck[3 * c].v.x), fabsf(my.v.y - ck[3 * c + 1].v.y)), fabsf(mz.v.x - ck[3 * c + 2].v.x));
        if (e > maxerr) maxerr = e;
        mx = ck[3 * c]; my = ck[3 * c + 1]; mz = ck[3 * c + 2];
      }
    }
    from_frame(k, hx, hy);
    gM0[3 * i] = hx.v.x; gM0[3 * i + 1] = hy.v.x; gM0[3 * i + 2] = hz.v.x;
    gM0[3 * i + 3] = hx.v.y; gM0[3 * i + 4] = hy.v.y; gM0[3 * i + 5] = hz.v.y;
  }
  *resync_err = maxerr;
}

#define ARGS(T) int nM, int nT, int K, const T* M0, const T* rf, const T* gr, const T* loc, const T* b1, \
  const double* df, const double* T1, const double* T2, const double* gamma, double dt, const T* gMo, T* Mo, T* gM0, \
  double* grf, double* ggr, T* resync_err
#define PASS nM, nT, K, M0, rf, gr, loc, b1, df, T1, T2, gamma, dt, gMo, Mo, gM0, grf, ggr, resync_err

extern "C" void host_sim_f32(int pol, int relax, ARGS(float)) {
  if (pol == 0) { if (relax) run<float, 0, true>(PASS); else run<float, 0, false>(PASS); }
  else          { if (relax) run<float, 1, true>(PASS); else run<float, 1, false>(PASS); }
}
extern "C" void host_sim_f32x2(int pol, int relax, ARGS(float)) {
  if (pol == 0) { if (relax) run2<0, true>(PASS); else run2<0, false>(PASS); }
  else          { if (relax) run2<1, true>(PASS); else run2<1, false>(PASS); }
}
extern "C" void host_sim_f64(int pol, int relax, ARGS(double)) {
  if (relax) run<double, 0, true>(PASS); else run<double, 0, false>(PASS);
}
extern "C" void host_sincos_f32(int n, const float* x, float* s, float* c) {
  for (int i = 0; i < n; ++i) Fn<float, TRIG_PRECISE>::sc(x[i], s[i], c[i]);
}
extern "C" void host_sincos_f64(int n, const double* x, double* s, double* c) {
  for (int i = 0; i < n; ++i) sincos_f64(x[i], s[i], c[i]);
}
extern "C" void host_rsq_f32(int n, const float* x, float* r) {
  for (int i = 0; i < n; ++i) r[i] = Fn<float, TRIG_PRECISE>::rsq(x[i]);
}

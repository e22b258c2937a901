// One Bloch step and its adjoint, per spin, in registers.
//
// Restates (does not copy) the mathematics of MRphy.py:
//   forward  mrphy/sims.py:100-126   m' = E (.) R(u,-phi) m + (1-E1) z
//   adjoint  mrphy/sims.py:204-261   h0 = R(u,+phi)(E (.) h1),  dL/dBeff per step
//   field    mrphy/beffective.py:137-167
//
// Formulation used here (b = 2*pi*gamma*dt*Beff, all per-spin constants pre-multiplied so the
// per-step work is FMAs only):
//   p2 = max(|b|^2, 1e-24)   rs = 1/sqrt(p2)   phi = p2*rs        (sims.py:100-101: phi>=1e-12)
//   a  = sin(phi)*rs         c1 = cos(phi)-1   d = -c1*rs^2
//   R(u,-phi) v = cos*v + d*(b.v)*b - a*(b x v)       R(u,+phi) v = cos*v + d*(b.v)*b + a*(b x v)
// The backward pass does not store states: the step is inverted exactly
//   m~ = E^-1 (m' + (E1-1) z),  m = R(u,+phi) m~
// and re-synchronised with forward checkpoints every K steps.
//
// Everything is MRPHY_HD so the same code compiles for the device kernels and for the
// host-side math check in tests/ (tests/host_math_harness.cpp).
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define MRPHY_HD __host__ __device__ __forceinline__
#else
#define MRPHY_HD inline
#endif

namespace mrphy {

enum TrigPolicy { TRIG_FAST = 0, TRIG_PRECISE = 1 };

// ---- f2: two spins side by side -----------------------------------------------------------
// On sm_100 the arithmetic below is ONE packed instruction per pair (FFMA2 / FMUL2 / FADD2): half the
// issue slots of scalar code for the same FP32-pipe work, which is what lets the non-FMA
// instructions (MUFU, LDS, STS, integer) issue in the shadow of the FMA pipe.
#if !defined(__CUDACC__)
struct float2 { float x, y; };
static inline float2 make_float2(float x, float y) { float2 r; r.x = x; r.y = y; return r; }
#endif
struct f2 {
  float2 v;
  MRPHY_HD f2() {}
  MRPHY_HD explicit f2(float s) { v = make_float2(s, s); }
  MRPHY_HD f2(float a, float b) { v = make_float2(a, b); }
};
MRPHY_HD f2 operator-(f2 a) { return f2(-a.v.x, -a.v.y); }
MRPHY_HD f2 operator*(f2 a, f2 b) {
#if defined(__CUDA_ARCH__) && __CUDA_ARCH__ >= 1000
  f2 r; r.v = __fmul2_rn(a.v, b.v); return r;
#else
  return f2(a.v.x * b.v.x, a.v.y * b.v.y);
#endif
}
MRPHY_HD f2 operator+(f2 a, f2 b) {
#if defined(__CUDA_ARCH__) && __CUDA_ARCH__ >= 1000
  f2 r; r.v = __fadd2_rn(a.v, b.v); return r;
#else
  return f2(a.v.x + b.v.x, a.v.y + b.v.y);
#endif
}
MRPHY_HD f2 operator-(f2 a, f2 b) { return a + (-b); }

template <typename T> MRPHY_HD T fma_(T a, T b, T c);
template <> MRPHY_HD float fma_<float>(float a, float b, float c) { return fmaf(a, b, c); }
template <> MRPHY_HD double fma_<double>(double a, double b, double c) { return fma(a, b, c); }
template <> MRPHY_HD f2 fma_<f2>(f2 a, f2 b, f2 c) {
#if defined(__CUDA_ARCH__) && __CUDA_ARCH__ >= 1000
  f2 r; r.v = __ffma2_rn(a.v, b.v, c.v); return r;
#else
  return f2(fmaf(a.v.x, b.v.x, c.v.x), fmaf(a.v.y, b.v.y, c.v.y));
#endif
}
// Negated forms.  FFMA2 has no operand-negate modifier (ptxas inserts two FADDs per negated pair), so
// for f2 these are written as two scalar FFMAs, whose negate modifiers are free: same FMA-pipe time as
// one FFMA2, one extra issue slot.  fnma_ = c - a*b ; fms_ = a*b - c.
template <typename T> MRPHY_HD T fnma_(T a, T b, T c) { return fma_(-a, b, c); }
template <> MRPHY_HD f2 fnma_<f2>(f2 a, f2 b, f2 c) { return f2(fmaf(-a.v.x, b.v.x, c.v.x), fmaf(-a.v.y, b.v.y, c.v.y)); }
template <typename T> MRPHY_HD T fms_(T a, T b, T c) { return fma_(a, b, -c); }
template <> MRPHY_HD f2 fms_<f2>(f2 a, f2 b, f2 c) { return f2(fmaf(a.v.x, b.v.x, -c.v.x), fmaf(a.v.y, b.v.y, -c.v.y)); }

MRPHY_HD float max_(float a, float b) { return fmaxf(a, b); }
MRPHY_HD double max_(double a, double b) { return fmax(a, b); }
MRPHY_HD f2 max_(f2 a, f2 b) { return f2(fmaxf(a.v.x, b.v.x), fmaxf(a.v.y, b.v.y)); }

// scalar type behind a (possibly packed) value type, and lane access
template <typename T> struct Scalar { typedef T type; };
template <> struct Scalar<f2> { typedef float type; };
MRPHY_HD float lane0(float a) { return a; }
MRPHY_HD double lane0(double a) { return a; }
MRPHY_HD float lane0(f2 a) { return a.v.x; }
MRPHY_HD float hsum(float a) { return a; }
MRPHY_HD double hsum(double a) { return a; }
MRPHY_HD float hsum(f2 a) { return a.v.x + a.v.y; }

// ---- rsqrt / sincos policies ------------------------------------------------------------
// float FAST   : MUFU.RSQ, MUFU.SIN, MUFU.COS (abs err ~2^-21.4 on sin/cos, 2 ulp on rsqrt)
// float PRECISE: MUFU.RSQ + one Newton step; Cody-Waite reduction + minimax polynomials on the
//                FMA pipe (~1 ulp), no local memory, valid for 0 <= phi < ~1e5 rad
// double       : rsqrt() and sincos() of the CUDA math library (FP64 pipe), both policies.
template <typename T, int POL> struct Fn;

// double: rsqrt() of the CUDA math library; sincos by a branch-free Cody-Waite reduction (three-part pi/2, exact for
// |x| < 2^20 * pi/2) and the fdlibm kernel polynomials (< 1 ulp on [-pi/4, pi/4]).  The library sincos() costs about
// twice as many FP64 instructions and carries a divergent slow path with a local-memory frame.
MRPHY_HD void sincos_f64(double x, double& s, double& c) {
  const double t = fma(x, 0.63661977236758138, 6755399441055744.0);   // x*2/pi + 1.5*2^52: low bits = quadrant
  const double jf = t - 6755399441055744.0;
  double r = fma(jf, -1.57079632673412561417e+00, x);
  r = fma(jf, -6.07710050630396597660e-11, r);
  r = fma(jf, -2.02226624879595063154e-21, r);
  const double z = r * r;
  double sp = fma(z, 1.58969099521155010221e-10, -2.50507602534068634195e-08);
  sp = fma(sp, z, 2.75573137070700676789e-06);
  sp = fma(sp, z, -1.98412698298579493134e-04);
  sp = fma(sp, z, 8.33333333332248946124e-03);
  sp = fma(sp, z, -1.66666666666666324348e-01);
  const double sr = fma(sp * z, r, r);
  double cp = fma(z, -1.13596475577881948265e-11, 2.08757232129817482790e-09);
  cp = fma(cp, z, -2.75573143513906633035e-07);
  cp = fma(cp, z, 2.48015872894767294178e-05);
  cp = fma(cp, z, -1.38888888888741095749e-03);
  cp = fma(cp, z, 4.16666666666666019037e-02);
  const double cr = fma(z * z, cp, fma(z, -0.5, 1.0));
#if defined(__CUDA_ARCH__)
  const int j = __double2loint(t);
#else
  union { double d; long long i; } u; u.d = t; const int j = (int)(u.i & 0xffffffffLL);
#endif
  const double ss = (j & 1) ? cr : sr;
  const double cc = (j & 1) ? sr : cr;
  s = (j & 2) ? -ss : ss;
  c = ((j + 1) & 2) ? -cc : cc;
}

template <int POL> struct Fn<double, POL> {
  static MRPHY_HD double rsq(double x) {
#if defined(__CUDA_ARCH__)
    return rsqrt(x);
#else
    return 1.0 / sqrt(x);
#endif
  }
  static MRPHY_HD void sc(double x, double& s, double& c) { sincos_f64(x, s, c); }
};

template <> struct Fn<float, TRIG_FAST> {
  static MRPHY_HD float rsq(float x) {
#if defined(__CUDA_ARCH__)
    return rsqrtf(x);
#else
    return 1.0f / sqrtf(x);
#endif
  }
  static MRPHY_HD void sc(float x, float& s, float& c) {
#if defined(__CUDA_ARCH__)
    __sincosf(x, &s, &c);
#else
    s = sinf(x); c = cosf(x);
#endif
  }
};

// Cody-Waite reduction with the round-to-nearest-via-magic-constant trick (two-term constant: exact to < 1e-8 rad
// for x < ~1e4 rad), then minimax sin/cos polynomials and a bit-level fix-up; no XU, no branches.  Shipped form
// (MRPHY_SC_MODPI, below): reduce by pi, polynomials on [-pi/2, pi/2], both results flip sign with the parity of the
// multiple.  -DMRPHY_SC_MODPI=0 is the earlier form: reduce by pi/2, polynomials on [-pi/4, pi/4], swap + signs.
#define MRPHY_SC_MAGIC 12582912.0f   /* 1.5 * 2^23 */
#ifndef MRPHY_PRECISE_NEWTON
#define MRPHY_PRECISE_NEWTON 1
#endif
MRPHY_HD int f_as_i(float x) {
#if defined(__CUDA_ARCH__)
  return __float_as_int(x);
#else
  union { float f; int i; } u; u.f = x; return u.i;
#endif
}
MRPHY_HD float i_as_f(int x) {
#if defined(__CUDA_ARCH__)
  return __int_as_float(x);
#else
  union { float f; int i; } u; u.i = x; return u.f;
#endif
}
// quadrant fix-up: (sr, cr) = sin/cos of the reduced argument, j = quadrant index bits
MRPHY_HD void sc_quadrant(int j, float sr, float cr, float& s, float& c) {
  const float ss = (j & 1) ? cr : sr;
  const float cc = (j & 1) ? sr : cr;
  s = i_as_f(f_as_i(ss) ^ ((j & 2) << 30));
  c = i_as_f(f_as_i(cc) ^ (((j + 1) & 2) << 30));
}
// Reduction by pi instead of pi/2 (MRPHY_SC_MODPI, default on): sin and cos of the reduced argument only change SIGN with
// the parity of the multiple, so the per-lane quadrant fix-up shrinks from a swap + two sign computations (8 scalar
// instructions per spin, which cannot be packed) to one shift + two XORs, for one more (packed) term in each polynomial:
// minimax on [-pi/2, pi/2]: max error 1.1e-7 / 9.0e-8, rms 1.9e-8 / 2.1e-8, mean angle error < 7e-9, mean of sin^2 + cos^2 - 1
// -7.7e-9 (profiles/fit_sincos.py).  Moving two coefficients by 3 and 8 ulp brings that radius mean to 2e-11, but the simulated
// magnetisation got WORSE with it (bench8: max|dM| 1.53e-5 -> 2.02e-5): the fp32 state update has a drift of its own that the
// small shrink happens to offset, so the plain minimax coefficients stay.
#ifndef MRPHY_SC_MODPI
#define MRPHY_SC_MODPI 1
#endif
template <typename V> MRPHY_HD void sc_poly_pi(V r, V& sr, V& cr) {
  const V r2 = r * r;
  V sp = fma_(r2, V(-2.543108479e-08f), V(2.760374173e-06f));
  sp = fma_(sp, r2, V(-1.984213741e-04f));
  sp = fma_(sp, r2, V(8.333338425e-03f));
  sp = fma_(sp, r2, V(-1.666666716e-01f));
  sr = fma_(sp * r2, r, r);
  V cp = fma_(r2, V(-2.633455551e-07f), V(2.477660746e-05f));
  cp = fma_(cp, r2, V(-1.388868783e-03f));
  cp = fma_(cp, r2, V(4.166666046e-02f));
  cp = fma_(cp, r2, V(-0.5f));
  cr = fma_(cp, r2, V(1.0f));
}
// both change sign with the parity of j
MRPHY_HD void sc_sign(int j, float sr, float cr, float& s, float& c) {
  const int sg = j << 31;
  s = i_as_f(f_as_i(sr) ^ sg);
  c = i_as_f(f_as_i(cr) ^ sg);
}
template <typename V> MRPHY_HD void sc_poly(V r, V& sr, V& cr) {
  const V r2 = r * r;
  V sp = fma_(r2, V(2.86567956e-6f), V(-1.98559923e-4f));
  sp = fma_(sp, r2, V(8.33338592e-3f));
  sp = fma_(sp, r2, V(-1.66666672e-1f));
  sr = fma_(sp * r2, r, r);
  V cp = fma_(r2, V(2.44677067e-5f), V(-1.38877297e-3f));
  cp = fma_(cp, r2, V(4.16666567e-2f));
  cp = fma_(cp, r2, V(-0.5f));
  cr = fma_(cp, r2, V(1.0f));
}

template <> struct Fn<float, TRIG_PRECISE> {
  static MRPHY_HD float rsq(float x) {
#if defined(__CUDA_ARCH__)
    float r = rsqrtf(x);
#else
    float r = (float)(1.0 / sqrt((double)x)) * (1.0f + 1.2e-7f);   // host: perturb so Newton does work
#endif
#if MRPHY_PRECISE_NEWTON
    // one Newton-Raphson step: r <- r + r*(0.5 - 0.5*x*r*r)
    float h = 0.5f * x * r;
    return fmaf(r, fmaf(-h, r, 0.5f), r);
#else
    return r;
#endif
  }
  static MRPHY_HD void sc(float x, float& s, float& c) {
#if MRPHY_SC_MODPI
    const float t = fmaf(x, 0.31830988618379067f, MRPHY_SC_MAGIC);
    const float jf = t - MRPHY_SC_MAGIC;
    float r = fmaf(jf, -3.1415920257568359375f, x);          // pi to 18 bits: jf * hi is exact for jf < 64
    r = fmaf(jf, -6.2783295107151866e-07f, r);
    float sr, cr;
    sc_poly_pi<float>(r, sr, cr);
    sc_sign(f_as_i(t), sr, cr, s, c);
#else
    const float t = fmaf(x, 0.63661977236758134f, MRPHY_SC_MAGIC);
    const float jf = t - MRPHY_SC_MAGIC;
    float r = fmaf(jf, -1.57079601287841796875f, x);
    r = fmaf(jf, -3.1391647326017846353e-07f, r);
    float sr, cr;
    sc_poly<float>(r, sr, cr);
    sc_quadrant(f_as_i(t), sr, cr, s, c);
#endif
  }
};

template <> struct Fn<f2, TRIG_FAST> {
  static MRPHY_HD f2 rsq(f2 x) { return f2(Fn<float, TRIG_FAST>::rsq(x.v.x), Fn<float, TRIG_FAST>::rsq(x.v.y)); }
  static MRPHY_HD void sc(f2 x, f2& s, f2& c) {
    Fn<float, TRIG_FAST>::sc(x.v.x, s.v.x, c.v.x);
    Fn<float, TRIG_FAST>::sc(x.v.y, s.v.y, c.v.y);
  }
};
template <> struct Fn<f2, TRIG_PRECISE> {
  static MRPHY_HD f2 rsq(f2 x) {
#if defined(__CUDA_ARCH__)
    f2 r(rsqrtf(x.v.x), rsqrtf(x.v.y));
#else
    f2 r((float)(1.0 / sqrt((double)x.v.x)) * (1.0f + 1.2e-7f), (float)(1.0 / sqrt((double)x.v.y)) * (1.0f + 1.2e-7f));
#endif
#if MRPHY_PRECISE_NEWTON
    const f2 nh = (f2(-0.5f) * x) * r;
    return fma_(r, fma_(nh, r, f2(0.5f)), r);
#else
    return r;
#endif
  }
  static MRPHY_HD void sc(f2 x, f2& s, f2& c) {
#if MRPHY_SC_MODPI
    const f2 t = fma_(x, f2(0.31830988618379067f), f2(MRPHY_SC_MAGIC));
    const f2 jf = t + f2(-MRPHY_SC_MAGIC);
    f2 r = fma_(jf, f2(-3.1415920257568359375f), x);
    r = fma_(jf, f2(-6.2783295107151866e-07f), r);
    f2 sr, cr;
    sc_poly_pi<f2>(r, sr, cr);
    sc_sign(f_as_i(t.v.x), sr.v.x, cr.v.x, s.v.x, c.v.x);
    sc_sign(f_as_i(t.v.y), sr.v.y, cr.v.y, s.v.y, c.v.y);
#else
    const f2 t = fma_(x, f2(0.63661977236758134f), f2(MRPHY_SC_MAGIC));
    const f2 jf = t + f2(-MRPHY_SC_MAGIC);
    f2 r = fma_(jf, f2(-1.57079601287841796875f), x);
    r = fma_(jf, f2(-3.1391647326017846353e-07f), r);
    f2 sr, cr;
    sc_poly<f2>(r, sr, cr);
    sc_quadrant(f_as_i(t.v.x), sr.v.x, cr.v.x, s.v.x, c.v.x);
    sc_quadrant(f_as_i(t.v.y), sr.v.y, cr.v.y, s.v.y, c.v.y);
#endif
  }
};

// ---- per-spin constants -----------------------------------------------------------------
// NC is the number of transmit coils held in registers (template); NC==1 covers "no b1Map"
// (coils pre-summed by the pack kernel, cbr=g, cbi=0).
template <typename T, int NC> struct SpinConst {
  T cbr[NC], cbi[NC];   // g*Re(b1), g*Im(b1)          g = 2*pi*gamma*dt  (sims.py:62);  NC == 1: g*|b1|, 0 (spin frame)
  T fc, fs;             // NC == 1: cos, sin of arg(b1): the spin's own transverse frame (make_consts); else 1, 0
  T glx, gly, glz;      // g*loc                        (beffective.py:137)
  T gbz0;               // g*df/gamma = 2*pi*dt*df      (beffective.py:142)
  T e1, e2;             // E1-1, E2-1 (expm1, full relative precision); 0 when no relaxation
  T iE1, iE2;           // 1/E1, 1/E2  (backward only)
  T e1i;                // (E1-1)/E1 = -expm1(dt/T1): m~z = (m'z + E1 - 1)/E1 = m'z*iE1 + e1i in ONE fma (backward only)
};

// Built once per spin, in double, from the inputs as given (gamma, dt, T1, T2, df may be fp32 or
// fp64 tensors; loc and b1 are in the working type): one rounding per constant.
template <typename T, int NC>
MRPHY_HD void make_consts(SpinConst<T, NC>& k, double gamma, double dt, bool relax, double T1, double T2, double df,
                          T lx, T ly, T lz, const T* b1r, const T* b1i) {
  const double g = 6.283185307179586476925286766559 * gamma * dt;
  k.fc = (T)1;
  k.fs = (T)0;
  if (NC == 1) {
    // One transmit channel: simulate in the spin's OWN transverse frame, the lab frame turned about z by arg(b1).  There
    // the sensitivity is real, (Bx + i By) = |b1| (rx + i ry): 2 multiplies per step instead of 4 multiply-adds for the
    // field, and 2 instead of 4 for the chain rule to rf in the adjoint.  A fixed rotation about z commutes with Bz and
    // with the relaxation, so only Mi / dL/dMo enter the frame (to_frame) and Mo / dL/dMi leave it (from_frame).
    const double br = b1r ? (double)b1r[0] : 1.0, bi = b1i ? (double)b1i[0] : 0.0;
    const double mag = sqrt(br * br + bi * bi);
    k.cbr[0] = (T)(g * mag);
    k.cbi[0] = (T)0;
    if (mag > 0.0) {
      k.fc = (T)(br / mag);
      k.fs = (T)(bi / mag);
    }
  } else {
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      k.cbr[c] = (T)(g * (b1r ? (double)b1r[c] : 1.0));
      k.cbi[c] = (T)(g * (b1i ? (double)b1i[c] : 0.0));
    }
  }
  k.glx = (T)(g * (double)lx);
  k.gly = (T)(g * (double)ly);
  k.glz = (T)(g * (double)lz);
  k.gbz0 = (T)(6.283185307179586476925286766559 * dt * df);
  if (relax) {
    const double x1 = -dt / T1, x2 = -dt / T2;
    k.e1 = (T)expm1(x1);
    k.e2 = (T)expm1(x2);
    k.iE1 = (T)exp(-x1);
    k.e1i = (T)(-expm1(-x1));
    k.iE2 = (T)exp(-x2);
  } else {
    k.e1 = k.e2 = (T)0;
    k.iE1 = k.iE2 = (T)1;
    k.e1i = (T)0;
  }
}

// two spins' constants side by side
template <int NC>
MRPHY_HD SpinConst<f2, NC> pack2(const SpinConst<float, NC>& a, const SpinConst<float, NC>& b) {
  SpinConst<f2, NC> k;
#pragma unroll
  for (int c = 0; c < NC; ++c) { k.cbr[c] = f2(a.cbr[c], b.cbr[c]); k.cbi[c] = f2(a.cbi[c], b.cbi[c]); }
  k.fc = f2(a.fc, b.fc); k.fs = f2(a.fs, b.fs);
  k.glx = f2(a.glx, b.glx); k.gly = f2(a.gly, b.gly); k.glz = f2(a.glz, b.glz); k.gbz0 = f2(a.gbz0, b.gbz0);
  k.e1 = f2(a.e1, b.e1); k.e2 = f2(a.e2, b.e2); k.iE1 = f2(a.iE1, b.iE1); k.iE2 = f2(a.iE2, b.iE2); k.e1i = f2(a.e1i, b.e1i);
  return k;
}

// Rotation coefficients shared by forward and backward.
template <typename T> struct RotCoef { T c, a, d, ne; };   // cos, sin/phi, (1 - cos)/phi^2, -(1 - sin/phi)/phi^2 (adjoint only)

// fp32 "precise", common case |b| <= 2 pi: the rotation only needs a = sin(phi)/phi and d = (1 - cos(phi))/phi^2, both
// EVEN entire functions of phi, i.e. functions of p2 = |b|^2 -- so neither the angle nor 1/phi is ever formed.  With the
// half angle h = phi/2:  a = sinc(h) cos(h),  d = sinc(h)^2 / 2, and sinc(h), cos(h) are 7-term polynomials in p2 on
// phi <= 2 pi (h <= pi: terms <= 1.7, no cancellation trouble).  Against MUFU.RSQ + Newton + Cody-Waite reduction + two
// polynomials + sign fix-ups this is 12 FMAs with immediate coefficients and 3 multiplies: 21 FP32-pipe operations and 2
// MUFU fewer per spin-step, at the same error (profiles/fit_halfangle.py: angle rms 4e-8, radius rms 1e-7, means < 1e-8
// on the bench distribution; the reduce-by-pi path: 8e-8, 1e-7, 8e-9; measured on the simulation, profiles/
// host_accuracy.py: rms |dM| 0.9x, max 0.9x those of the reduce-by-pi path).  Leading coefficients are exactly 1, so
// there is no systematic scale (radius) error.  A warp with a step of |b|^2 > MRPHY_HALF_P2MAX (2.7e-4 of the warp-steps
// of the bench distributions, 1.3e-5 of the spin-steps) takes the reduce-by-pi path: one vote and a predicated-off call.
#ifndef MRPHY_HALF_ANGLE
#define MRPHY_HALF_ANGLE 1
#endif
#ifndef MRPHY_VOTE_MASK
#define MRPHY_VOTE_MASK __activemask()   /* kernels whose loops run with all 32 lanes define it as 0xffffffffu before including this file */
#endif
#define MRPHY_HALF_P2MAX 40.0f   /* (2 pi * 1.0066)^2; the polynomials are fitted on [0, (2 pi * 1.02)^2] */
MRPHY_HD bool any_gt(float a, float lim) { return a > lim; }
MRPHY_HD bool any_gt(double a, double lim) { return a > lim; }
MRPHY_HD bool any_gt(f2 a, float lim) { return fmaxf(a.v.x, a.v.y) > lim; }
template <typename V> MRPHY_HD RotCoef<V> rot_coef_half(V p2) {
  // A = sinc(phi/2) = 1 + p2 PA(p2),  C = cos(phi/2) = 1 + p2 PC(p2)   (profiles/fit_halfangle.py, coefficients rounded
  // to fp32 one at a time with the remaining ones refitted, so the means of the angle and radius errors stay < 1e-8)
  V PA = fma_(p2, V(3.280267225e-14f), V(-2.409831189e-11f));
  PA = fma_(PA, p2, V(1.075399236e-08f));
  PA = fma_(PA, p2, V(-3.100041567e-06f));
  PA = fma_(PA, p2, V(5.208322546e-04f));
  PA = fma_(PA, p2, V(-4.166666418e-02f));
  const V A = fma_(PA, p2, V(1.0f));
  V PC = fma_(p2, V(4.191810202e-13f), V(-2.642699393e-10f));
  PC = fma_(PC, p2, V(9.675193269e-08f));
  PC = fma_(PC, p2, V(-2.169963955e-05f));
  PC = fma_(PC, p2, V(2.604155801e-03f));
  PC = fma_(PC, p2, V(-1.249999776e-01f));
  const V C = fma_(PC, p2, V(1.0f));
  RotCoef<V> r;
  r.a = A * C;
  r.d = (A * V(0.5f)) * A;
  r.c = fnma_(p2, r.d, V(1.0f));
  // -(1 - a)/p2 without a division: 1 - A C = -p2 (PA + PC A)                     (adjoint only; dead code in the forward)
  r.ne = fma_(PC, A, PA);
  return r;
}

// any angle: rsqrt, phi, sincos by the policy
template <typename T, int POL>
MRPHY_HD RotCoef<T> rot_coef_reduced(T p2) {
  T rs = Fn<T, POL>::rsq(p2);
  T phi = p2 * rs;
  T s, c;
  Fn<T, POL>::sc(phi, s, c);
  RotCoef<T> r;
  r.c = c;
  r.a = s * rs;
  const T rs2 = rs * rs;
  r.d = fnma_(c, rs2, rs2);     // (1 - cos) / phi^2
  r.ne = fms_(r.a, rs2, rs2);   // -(1 - sin/phi) / phi^2
  return r;
}

#ifndef MRPHY_HALF_BWD
#define MRPHY_HALF_BWD 1
#endif
// the rare |b| > 2 pi path as a real call: the common path then carries one predicated-off CALL instead of a taken
// branch around ~40 inlined instructions per step (-DMRPHY_HALF_INLINE_SLOW=1 inlines it again)
#ifndef MRPHY_HALF_INLINE_SLOW
#define MRPHY_HALF_INLINE_SLOW 0
#endif
#if defined(__CUDACC__) && !MRPHY_HALF_INLINE_SLOW
template <typename T, int POL>
__device__ __noinline__ RotCoef<T> rot_coef_reduced_call(T p2) { return rot_coef_reduced<T, POL>(p2); }
#endif
template <typename T, int POL>
MRPHY_HD void rot_coef_slow(T p2, RotCoef<T>& r) {
#if defined(__CUDA_ARCH__) && !MRPHY_HALF_INLINE_SLOW
  r = rot_coef_reduced_call<T, POL>(p2);
#else
  r = rot_coef_reduced<T, POL>(p2);
#endif
}

template <typename T, int POL, bool HALF_OK = true>
MRPHY_HD RotCoef<T> rot_coef(T bx, T by, T bz) {
  T p2;
  if (sizeof(typename Scalar<T>::type) == 4) {
    // fp32: the reference's phi >= 1e-12 clamp as +1e-24 under the root -- invisible next to any |b|^2 >= 1e-17 (fp32
    // resolution), the same 1e-24 for a zero field (exact identity step), and two FMNMX per thread-step cheaper
    p2 = fma_(bx, bx, fma_(by, by, fma_(bz, bz, (T)1e-24f)));
#if MRPHY_HALF_ANGLE
    if (POL == TRIG_PRECISE && HALF_OK) {
      RotCoef<T> r = rot_coef_half<T>(p2);
#ifndef MRPHY_HALF_NOFALLBACK   /* modelling builds only (profiles/model_all.py): the common path without the branch */
      // the vote makes the branch warp-uniform: no divergence bookkeeping (BSSY/BSYNC) on the common path, and the
      // reduce-by-pi coefficients are valid for every lane, so the whole warp may take them
#if defined(__CUDA_ARCH__) && !defined(MRPHY_HALF_LANE_BRANCH)
      if (__builtin_expect(__any_sync(MRPHY_VOTE_MASK, any_gt(p2, (typename Scalar<T>::type)MRPHY_HALF_P2MAX)), 0)) rot_coef_slow<T, POL>(p2, r);
#else
      if (any_gt(p2, (typename Scalar<T>::type)MRPHY_HALF_P2MAX)) r = rot_coef_reduced<T, POL>(p2);
#endif
#endif
      return r;
    }
#endif
  } else {
    p2 = fma_(bx, bx, fma_(by, by, bz * bz));
    p2 = max_(p2, (T)1e-24f);
  }
  return rot_coef_reduced<T, POL>(p2);
}

// field of one step from the staged waveform sample (rx[c], ry[c], gx, gy, gz)
template <typename T, int NC>
MRPHY_HD void field(const SpinConst<T, NC>& k, const T* rx, const T* ry, T gx, T gy, T gz, T& bx, T& by, T& bz) {
  bx = k.cbr[0] * rx[0];
  by = k.cbr[0] * ry[0];
  if (NC > 1) {
    bx = fnma_(k.cbi[0], ry[0], bx);
    by = fma_(k.cbi[0], rx[0], by);
  }
#pragma unroll
  for (int c = 1; c < NC; ++c) {
    bx = fma_(k.cbr[c], rx[c], bx);
    by = fma_(k.cbr[c], ry[c], by);
    bx = fnma_(k.cbi[c], ry[c], bx);
    by = fma_(k.cbi[c], rx[c], by);
  }
  bz = fma_(k.glx, gx, fma_(k.gly, gy, fma_(k.glz, gz, k.gbz0)));
}

// lab frame <-> the spin's own transverse frame (NC == 1, see make_consts): v' = Rz(-arg b1) v, v = Rz(+arg b1) v'
template <typename T, int NC> MRPHY_HD void to_frame(const SpinConst<T, NC>& k, T& x, T& y) {
  if (NC == 1) {
    const T x1 = fma_(k.fs, y, k.fc * x), y1 = fnma_(k.fs, x, k.fc * y);
    x = x1; y = y1;
  }
}
template <typename T, int NC> MRPHY_HD void from_frame(const SpinConst<T, NC>& k, T& x, T& y) {
  if (NC == 1) {
    const T x1 = fnma_(k.fs, y, k.fc * x), y1 = fma_(k.fs, x, k.fc * y);
    x = x1; y = y1;
  }
}
// chain rule of the field to the rf sample of coil q: (d/drx, d/dry) contributions of one spin, F = -dL/db
template <typename T, int NC> MRPHY_HD void rf_chain(const SpinConst<T, NC>& k, int q, T Fx, T Fy, T& gx, T& gy) {
  if (NC == 1) {
    gx = k.cbr[0] * Fx;
    gy = k.cbr[0] * Fy;
  } else {
    gx = fma_(k.cbr[q], Fx, k.cbi[q] * Fy);
    gy = fnma_(k.cbi[q], Fx, k.cbr[q] * Fy);
  }
}

// ---- forward step -------------------------------------------------------------------------
// apply_fwd: the recurrent part (needs the previous state); rot_coef above is the part that only needs the waveform.
template <typename T, bool RELAX>
MRPHY_HD void apply_fwd(const RotCoef<T>& r, T bx, T by, T bz, T e1, T e2, T& mx, T& my, T& mz) {
  T kk = r.d * fma_(bx, mx, fma_(by, my, bz * mz));
  T abx = r.a * bx, aby = r.a * by, abz = r.a * bz;
  // m~ = c*m + kk*b - (ab x m)
  T nx = fma_(abz, my, fnma_(aby, mz, fma_(kk, bx, r.c * mx)));
  T ny = fma_(abx, mz, fnma_(abz, mx, fma_(kk, by, r.c * my)));
  T nz = fma_(aby, mx, fnma_(abx, my, fma_(kk, bz, r.c * mz)));
  if (RELAX) {   // E2*m~xy ; E1*m~z - (E1-1)  written with e = E-1 so no cancellation
    nx = fma_(e2, nx, nx);
    ny = fma_(e2, ny, ny);
    nz = fma_(e1, nz + (T)(-1.0f), nz);
  }
  mx = nx; my = ny; mz = nz;
}
template <typename T, int POL, bool RELAX>
MRPHY_HD void step_fwd(T bx, T by, T bz, T e1, T e2, T& mx, T& my, T& mz) {
  const RotCoef<T> r = rot_coef<T, POL>(bx, by, bz);
  apply_fwd<T, RELAX>(r, bx, by, bz, e1, e2, mx, my, mz);
}

// ---- backward step --------------------------------------------------------------------------
// in : (mx,my,mz) state AFTER the step, (hx,hy,hz) = dL/d(state after the step)
// out: state BEFORE the step, dL/d(state before), and F = -(1/g) dL/dBeff (sign/scale folded
//      into the per-spin constants and the finalize kernel)
template <typename T, bool RELAX, int NC>
MRPHY_HD void apply_bwd(const SpinConst<T, NC>& k, const RotCoef<T>& r, T bx, T by, T bz, T& mx, T& my, T& mz,
                        T& hx, T& hy, T& hz, T& Fx, T& Fy, T& Fz) {
  T tx = mx, ty = my, tz = mz;   // m~ (pre-relaxation state)
  T gx = hx, gy = hy, gz = hz;   // h~
  if (RELAX) {
    tx = mx * k.iE2;
    ty = my * k.iE2;
    tz = fma_(mz, k.iE1, k.e1i);
    gx = fma_(k.e2, hx, hx);
    gy = fma_(k.e2, hy, hy);
    gz = fma_(k.e1, hz, hz);
  }
  T Q = fma_(bx, tx, fma_(by, ty, bz * tz));   // b.m~ == b.m0
  T P = fma_(bx, gx, fma_(by, gy, bz * gz));   // b.h~
  T kq = r.d * Q, kp = r.d * P;
  T abx = r.a * bx, aby = r.a * by, abz = r.a * bz;
  // m0 = c*m~ + kq*b + (ab x m~)
  T px = fnma_(abz, ty, fma_(aby, tz, fma_(kq, bx, r.c * tx)));
  T py = fnma_(abx, tz, fma_(abz, tx, fma_(kq, by, r.c * ty)));
  T pz = fnma_(aby, tx, fma_(abx, ty, fma_(kq, bz, r.c * tz)));
  // wb = b x h~
  T wx = fms_(by, gz, bz * gy);
  T wy = fms_(bz, gx, bx * gz);
  T wz = fms_(bx, gy, by * gx);
  // F = -dL/db.  With R = exp(-[b]x) the differential of the rotation is dR = -[J db]x R, J = a I + e b b^T + d [b]x the
  // left Jacobian of SO(3) in b-form (e = (1 - a)/|b|^2), so -dL/db = J (m~ x h~); and with b x (m~ x h~) = P m~ - Q h~,
  // b.(m~ x h~) = -m~.wb:
  //     F = a (m~ x h~) + d (P m~ - Q h~) - e (m~.wb) b
  // (same value as the reference's closed form, sims.py:204-261, which is written in the pre-rotation state)
  T E = r.ne * fma_(tx, wx, fma_(ty, wy, tz * wz));   // -e (m~.wb)
  T cx = fms_(ty, gz, tz * gy);
  T cy = fms_(tz, gx, tx * gz);
  T cz = fms_(tx, gy, ty * gx);
  Fx = fma_(E, bx, fnma_(kq, gx, fma_(kp, tx, r.a * cx)));
  Fy = fma_(E, by, fnma_(kq, gy, fma_(kp, ty, r.a * cy)));
  Fz = fma_(E, bz, fnma_(kq, gz, fma_(kp, tz, r.a * cz)));
  // h0 = c*h~ + kp*b + a*wb
  hx = fma_(r.a, wx, fma_(kp, bx, r.c * gx));
  hy = fma_(r.a, wy, fma_(kp, by, r.c * gy));
  hz = fma_(r.a, wz, fma_(kp, bz, r.c * gz));
  mx = px; my = py; mz = pz;
}
template <typename T, int POL, bool RELAX, int NC>
MRPHY_HD void step_bwd(const SpinConst<T, NC>& k, T bx, T by, T bz, T& mx, T& my, T& mz, T& hx, T& hy, T& hz,
                       T& Fx, T& Fy, T& Fz) {
  const RotCoef<T> r = rot_coef<T, POL, MRPHY_HALF_BWD != 0>(bx, by, bz);
  apply_bwd<T, RELAX, NC>(k, r, bx, by, bz, mx, my, mz, hx, hy, hz, Fx, Fy, Fz);
}

}  // namespace mrphy

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

template <typename T> MRPHY_HD T fma_(T a, T b, T c);
template <> MRPHY_HD float fma_<float>(float a, float b, float c) { return fmaf(a, b, c); }
template <> MRPHY_HD double fma_<double>(double a, double b, double c) { return fma(a, b, c); }

// ---- rsqrt / sincos policies ------------------------------------------------------------
// float FAST   : MUFU.RSQ, MUFU.SIN, MUFU.COS (abs err ~2^-21.4 on sin/cos, 2 ulp on rsqrt)
// float PRECISE: MUFU.RSQ + one Newton step; Cody-Waite reduction + minimax polynomials on the
//                FMA pipe (~1 ulp), no local memory, valid for 0 <= phi < ~1e5 rad
// double       : rsqrt() and sincos() of the CUDA math library (FP64 pipe), both policies.
template <typename T, int POL> struct Fn;

template <int POL> struct Fn<double, POL> {
  static MRPHY_HD double rsq(double x) {
#if defined(__CUDA_ARCH__)
    return rsqrt(x);
#else
    return 1.0 / sqrt(x);
#endif
  }
  static MRPHY_HD void sc(double x, double& s, double& c) {
#if defined(__CUDA_ARCH__)
    sincos(x, &s, &c);
#else
    s = sin(x); c = cos(x);
#endif
  }
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

template <> struct Fn<float, TRIG_PRECISE> {
  static MRPHY_HD float rsq(float x) {
#if defined(__CUDA_ARCH__)
    float r = rsqrtf(x);
#else
    float r = (float)(1.0 / sqrt((double)x)) * (1.0f + 1.2e-7f);   // host: perturb so Newton does work
#endif
    // one Newton-Raphson step: r <- r * (1.5 - 0.5*x*r*r)
    float h = 0.5f * x * r;
    return fmaf(r, fmaf(-h, r, 0.5f), r);
  }
  static MRPHY_HD void sc(float x, float& s, float& c) {
    // j = nearest integer to x*2/pi ; r = x - j*pi/2 in three Cody-Waite pieces
    float jf = rintf(x * 0.63661977236758134f);
    float r = fmaf(jf, -1.57079601287841796875f, x);
    r = fmaf(jf, -3.1391647326017846353e-07f, r);
    r = fmaf(jf, -5.3903029534742383927e-15f, r);
    int j = (int)jf;
    float r2 = r * r;
    // sin(r), cos(r) on [-pi/4, pi/4]  (minimax, ~1 ulp)
    float sp = fmaf(r2, 2.86567956e-6f, -1.98559923e-4f);
    sp = fmaf(sp, r2, 8.33338592e-3f);
    sp = fmaf(sp, r2, -1.66666672e-1f);
    float sr = fmaf(sp * r2, r, r);
    float cp = fmaf(r2, 2.44677067e-5f, -1.38877297e-3f);
    cp = fmaf(cp, r2, 4.16666567e-2f);
    cp = fmaf(cp, r2, -0.5f);
    float cr = fmaf(cp, r2, 1.0f);
    float ss = (j & 1) ? cr : sr;
    float cc = (j & 1) ? sr : cr;
    s = (j & 2) ? -ss : ss;
    c = ((j + 1) & 2) ? -cc : cc;
  }
};

// ---- per-spin constants -----------------------------------------------------------------
// NC is the number of transmit coils held in registers (template); NC==1 covers "no b1Map"
// (coils pre-summed by the pack kernel, cbr=g, cbi=0).
template <typename T, int NC> struct SpinConst {
  T cbr[NC], cbi[NC];   // g*Re(b1), g*Im(b1)          g = 2*pi*gamma*dt  (sims.py:62)
  T glx, gly, glz;      // g*loc                        (beffective.py:137)
  T gbz0;               // g*df/gamma = 2*pi*dt*df      (beffective.py:142)
  T e1, e2;             // E1-1, E2-1 (expm1, full relative precision); 0 when no relaxation
  T iE1, iE2;           // 1/E1, 1/E2  (backward only)
};

// Built once per spin, in double, from the inputs as given (gamma, dt, T1, T2, df may be fp32 or
// fp64 tensors; loc and b1 are in the working type): one rounding per constant.
template <typename T, int NC>
MRPHY_HD void make_consts(SpinConst<T, NC>& k, double gamma, double dt, bool relax, double T1, double T2, double df,
                          T lx, T ly, T lz, const T* b1r, const T* b1i) {
  const double g = 6.283185307179586476925286766559 * gamma * dt;
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    k.cbr[c] = (T)(g * (b1r ? (double)b1r[c] : 1.0));
    k.cbi[c] = (T)(g * (b1i ? (double)b1i[c] : 0.0));
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
    k.iE2 = (T)exp(-x2);
  } else {
    k.e1 = k.e2 = (T)0;
    k.iE1 = k.iE2 = (T)1;
  }
}

// Rotation coefficients shared by forward and backward.
template <typename T> struct RotCoef { T c, a, d, rs2; };

template <typename T, int POL>
MRPHY_HD RotCoef<T> rot_coef(T bx, T by, T bz) {
  T p2 = fma_(bx, bx, fma_(by, by, bz * bz));
  p2 = p2 > (T)1e-24 ? p2 : (T)1e-24;
  T rs = Fn<T, POL>::rsq(p2);
  T phi = p2 * rs;
  T s, c;
  Fn<T, POL>::sc(phi, s, c);
  RotCoef<T> r;
  r.c = c;
  r.a = s * rs;
  r.rs2 = rs * rs;
  r.d = ((T)1 - c) * r.rs2;
  return r;
}

// field of one step from the staged waveform sample (rx[c], ry[c], gx, gy, gz)
template <typename T, int NC>
MRPHY_HD void field(const SpinConst<T, NC>& k, const T* rx, const T* ry, T gx, T gy, T gz, T& bx, T& by, T& bz) {
  bx = k.cbr[0] * rx[0];
  by = k.cbr[0] * ry[0];
  bx = fma_(-k.cbi[0], ry[0], bx);
  by = fma_(k.cbi[0], rx[0], by);
#pragma unroll
  for (int c = 1; c < NC; ++c) {
    bx = fma_(k.cbr[c], rx[c], bx);
    by = fma_(k.cbr[c], ry[c], by);
    bx = fma_(-k.cbi[c], ry[c], bx);
    by = fma_(k.cbi[c], rx[c], by);
  }
  bz = fma_(k.glx, gx, fma_(k.gly, gy, fma_(k.glz, gz, k.gbz0)));
}

// ---- forward step -------------------------------------------------------------------------
template <typename T, int POL, bool RELAX>
MRPHY_HD void step_fwd(T bx, T by, T bz, T e1, T e2, T& mx, T& my, T& mz) {
  RotCoef<T> r = rot_coef<T, POL>(bx, by, bz);
  T kk = r.d * fma_(bx, mx, fma_(by, my, bz * mz));
  T abx = r.a * bx, aby = r.a * by, abz = r.a * bz;
  // m~ = c*m + kk*b - (ab x m)
  T nx = fma_(abz, my, fma_(-aby, mz, fma_(kk, bx, r.c * mx)));
  T ny = fma_(abx, mz, fma_(-abz, mx, fma_(kk, by, r.c * my)));
  T nz = fma_(aby, mx, fma_(-abx, my, fma_(kk, bz, r.c * mz)));
  if (RELAX) {   // E2*m~xy ; E1*m~z - (E1-1)  written with e = E-1 so no cancellation
    nx = fma_(e2, nx, nx);
    ny = fma_(e2, ny, ny);
    nz = fma_(e1, nz - (T)1, nz);
  }
  mx = nx; my = ny; mz = nz;
}

// ---- backward step --------------------------------------------------------------------------
// in : (mx,my,mz) state AFTER the step, (hx,hy,hz) = dL/d(state after the step)
// out: state BEFORE the step, dL/d(state before), and F = -(1/g) dL/dBeff (sign/scale folded
//      into the per-spin constants and the finalize kernel)
template <typename T, int POL, bool RELAX, int NC>
MRPHY_HD void step_bwd(const SpinConst<T, NC>& k, T bx, T by, T bz, T& mx, T& my, T& mz, T& hx, T& hy, T& hz,
                       T& Fx, T& Fy, T& Fz) {
  RotCoef<T> r = rot_coef<T, POL>(bx, by, bz);
  T tx = mx, ty = my, tz = mz;   // m~ (pre-relaxation state)
  T gx = hx, gy = hy, gz = hz;   // h~
  if (RELAX) {
    tx = mx * k.iE2;
    ty = my * k.iE2;
    tz = (mz + k.e1) * k.iE1;
    gx = fma_(k.e2, hx, hx);
    gy = fma_(k.e2, hy, hy);
    gz = fma_(k.e1, hz, hz);
  }
  T Q = fma_(bx, tx, fma_(by, ty, bz * tz));   // b.m~ == b.m0
  T P = fma_(bx, gx, fma_(by, gy, bz * gz));   // b.h~
  T kq = r.d * Q, kp = r.d * P;
  T abx = r.a * bx, aby = r.a * by, abz = r.a * bz;
  // m0 = c*m~ + kq*b + (ab x m~)
  T px = fma_(-abz, ty, fma_(aby, tz, fma_(kq, bx, r.c * tx)));
  T py = fma_(-abx, tz, fma_(abz, tx, fma_(kq, by, r.c * ty)));
  T pz = fma_(-aby, tx, fma_(abx, ty, fma_(kq, bz, r.c * tz)));
  // wb = b x h~
  T wx = fma_(by, gz, -bz * gy);
  T wy = fma_(bz, gx, -bx * gz);
  T wz = fma_(bx, gy, -by * gx);
  // F = a (m0 x h~) - d (Q h~ + P m0) - [ (m~ - a m0).wb - 2 d P Q ] rs^2 b
  T nx = fma_(-r.a, px, tx), ny = fma_(-r.a, py, ty), nz = fma_(-r.a, pz, tz);
  T C = fma_(nx, wx, fma_(ny, wy, nz * wz));
  C = fma_((T)-2 * kp, Q, C) * r.rs2;
  T cx = fma_(py, gz, -pz * gy);
  T cy = fma_(pz, gx, -px * gz);
  T cz = fma_(px, gy, -py * gx);
  Fx = fma_(-C, bx, fma_(r.a, cx, -fma_(kq, gx, kp * px)));
  Fy = fma_(-C, by, fma_(r.a, cy, -fma_(kq, gy, kp * py)));
  Fz = fma_(-C, bz, fma_(r.a, cz, -fma_(kq, gz, kp * pz)));
  // h0 = c*h~ + kp*b + a*wb
  hx = fma_(r.a, wx, fma_(kp, bx, r.c * gx));
  hy = fma_(r.a, wy, fma_(kp, by, r.c * gy));
  hz = fma_(r.a, wz, fma_(kp, bz, r.c * gz));
  mx = px; my = py; mz = pz;
}

}  // namespace mrphy

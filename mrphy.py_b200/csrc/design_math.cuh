// Device side of the optimiser's re-parametrisation chain (utils.py:114-131, 239-256, 293-330), shared by
// design_waveform_kernel (design_ops.cu: the chain and its adjoint as launches of their own) and by the design tail of
// grad_finalize_design_kernel (grad_finalize.cuh: the adjoint evaluated by the last CTA of the gradient epilogue).
// All arithmetic in double, one final rounding to T.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/mrphy_b200.h"
#include "abi_common.cuh"

namespace mrphy {

constexpr double TWO_OVER_PI = 0.63661977236758134307553505349006;

// inclusive scan of one value per thread across a CTA of NT threads (tid = linear thread index, warps = tid / 32);
// returns the CTA total through `total`.  warp_tot: NT / 32 doubles of shared memory.
template <int NT>
__device__ __forceinline__ double block_scan_incl(double v, double* warp_tot, double& total, int tid) {
  const int lane = tid & 31, w = tid >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double u = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += u;
  }
  if (lane == 31) warp_tot[w] = v;
  __syncthreads();
  double pre = 0.0, tot = 0.0;
#pragma unroll
  for (int q = 0; q < NT / 32; ++q) {
    const double x = warp_tot[q];
    if (q < w) pre += x;
    tot += x;
  }
  __syncthreads();   // warp_tot is reused by the next chunk
  total = tot;
  return v + pre;
}

// gradient half, one (n, xyz) row handled by a whole CTA: chunks of NT samples, running carry between chunks.
//   forward:  ts | s -> g | s        adjoint:  gg = dL/dg row -> dL/dts row (reversed running sum, then through atan)
template <typename T, int NT>
__device__ __forceinline__ void design_gr_row(const mrphy_reparam_args& a, int row, const T* gg, double* warp_tot, int tid) {
  const int n = row / 3, x = row % 3, nT = a.nT;
  const bool use_atan = a.gr_kind != 2, scan = a.gr_kind != 3;
  const double smax = use_atan ? (double)((const T*)a.smax)[(int64_t)n * a.smax_sn + (int64_t)x * a.smax_sx] : 1.0;
  const double dt = scan ? ld_param(a.dt, n, 0) : 1.0;
  const T* in = (const T*)a.ts + (int64_t)row * nT;
  double carry = 0.0;
  if (!a.adjoint) {
    T* out = (T*)a.gr + (int64_t)row * nT;
    for (int base = 0; base < nT; base += NT) {
      const int t = base + tid;
      double s = 0.0;
      if (t < nT) s = use_atan ? atan((double)in[t]) * TWO_OVER_PI * smax : (double)in[t];
      if (!scan) {
        if (t < nT) out[t] = (T)s;
        continue;
      }
      double tot;
      const double inc = block_scan_incl<NT>(s, warp_tot, tot, tid);
      if (t < nT) out[t] = (T)(dt * (carry + inc));
      carry += tot;
    }
  } else {
    // dL/ds[t] = dt * sum_{t' >= t} dL/dg[t']  (reversed running sum), then through atan
    T* out = (T*)a.gts + (int64_t)row * nT;
    for (int base = 0; base < nT; base += NT) {
      const int t = nT - 1 - (base + tid);
      double g = t >= 0 ? (double)gg[t] : 0.0;
      if (scan) {
        double tot;
        const double inc = block_scan_incl<NT>(g, warp_tot, tot, tid);
        g = dt * (carry + inc);
        carry += tot;
      }
      if (t >= 0) {
        if (use_atan) {
          const double v = (double)in[t];
          g *= TWO_OVER_PI * smax / (1.0 + v * v);
        }
        out[t] = (T)g;
      }
    }
  }
}

// rf half, element e = (n, t, c) of rho / theta: forward writes rf, adjoint reads grf = dL/drf and writes dL/drho, dL/dtheta
template <typename T>
__device__ __forceinline__ void design_rf_elem(const mrphy_reparam_args& a, int64_t e, const T* grf) {
  const int64_t per = (int64_t)a.nT * a.nC;
  const int n = (int)(e / per);
  const int64_t r = e - (int64_t)n * per;
  const int c = (int)(r % a.nC);
  const double rho = (double)((const T*)a.rho)[e], th = (double)((const T*)a.theta)[e];
  const double rfmax = (double)((const T*)a.rfmax)[(int64_t)n * a.rfmax_sn + (int64_t)c * a.rfmax_sc];
  double A, dA;
  if (a.rf_kind == 1) {
    A = atan(rho) * TWO_OVER_PI;
    dA = TWO_OVER_PI / (1.0 + rho * rho);
  } else {
    A = 1.0 / (1.0 + exp(-rho));
    dA = A * (1.0 - A);
  }
  double sn, cs;
  sincos(th, &sn, &cs);
  const int64_t ox = (int64_t)n * 2 * per + r, oy = ox + per;
  if (!a.adjoint) {
    T* rf = (T*)a.rf;
    rf[ox] = (T)(A * rfmax * cs);
    rf[oy] = (T)(A * rfmax * sn);
  } else {
    const double gx = (double)grf[ox], gy = (double)grf[oy];
    ((T*)a.grho)[e] = (T)(dA * rfmax * (cs * gx + sn * gy));
    ((T*)a.gtheta)[e] = (T)(A * rfmax * (cs * gy - sn * gx));
  }
}

// The whole adjoint of batch entry n by ONE CTA of NT threads (the design tail of the gradient epilogue):
// grf (N,2,nT,nC) / ggr (N,3,nT) are the finished waveform gradients.
template <typename T, int NT>
__device__ __forceinline__ void design_adjoint_entry(const mrphy_reparam_args& a, int n, const T* grf, const T* ggr,
                                                     double* warp_tot, int tid) {
  if (a.gr_kind) {
    for (int x = 0; x < 3; ++x) design_gr_row<T, NT>(a, n * 3 + x, ggr + ((int64_t)n * 3 + x) * a.nT, warp_tot, tid);
  }
  if (a.rf_kind) {
    const int64_t per = (int64_t)a.nT * a.nC;
    for (int64_t r = tid; r < per; r += NT) design_rf_elem<T>(a, (int64_t)n * per + r, grf);
  }
}

}  // namespace mrphy

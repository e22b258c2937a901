// Waveform-sized and mask-sized companions of the simulation path, sm_100a.
//
//   design_waveform_kernel<T>   the optimiser's re-parametrisation chain in ONE launch (and its adjoint in one):
//                               rf = A(rho)*rfmax*(cos theta, sin theta)   utils.py:114-131 (l-rho), 311-330 (t-rho)
//                               s  = atan(ts)*2/pi*smax                     utils.py:293-308 (ts2s)
//                               g  = dt*cumsum(s)                           utils.py:239-256 (s2g)
//                               upstream: ~10 elementwise launches + a scan forward, ~25 launches backward.
//   clamp_waveform_kernel<T>    utils.rfclamp / utils.sclamp (utils.py:217-236, 278-293) and their adjoints, one launch each
//   mask_copy_kernel<T>         SpinArray.extract / .embed (mobjs.py:512-553) as one pass over the output with the NaN
//                               padding fused (upstream: full() + masked assignment).
//
// These are O(N*nT) / O(N*prod(Nd)) and HBM- or latency-bound; all arithmetic is done in double and rounded once, so
// the fp32 results are correctly rounded images of the reference's formula.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/mrphy_b200.h"
#include "abi_common.cuh"
#include "design_math.cuh"

namespace mrphy {

constexpr int RP_THREADS = 256;

template <typename T>
__global__ void __launch_bounds__(RP_THREADS) design_waveform_kernel(const mrphy_reparam_args a, const int gr_rows) {
  __shared__ double warp_tot[RP_THREADS / 32];
  const int tid = threadIdx.x;
  if ((int)blockIdx.x < gr_rows) {   // gradient half: one CTA per (n, xyz) row
    const int row = blockIdx.x;
    design_gr_row<T, RP_THREADS>(a, row, a.adjoint ? (const T*)a.ggr + (int64_t)row * a.nT : nullptr, warp_tot, tid);
    return;
  }
  // rf half: one thread per (n, t, c)
  const int64_t e = (int64_t)(blockIdx.x - gr_rows) * RP_THREADS + tid;
  if (e < (int64_t)a.N * a.nT * a.nC) design_rf_elem<T>(a, e, (const T*)a.grf);
}

// utils.rfclamp / utils.sclamp and their adjoints: one thread per (n, t, c) rf sample pair, or per slew sample
template <typename T>
__global__ void __launch_bounds__(256) clamp_waveform_kernel(const mrphy_clamp_args a, const int64_t total) {
  const int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (e >= total) return;
  const T* x = (const T*)a.x;
  const T* g = (const T*)a.g;
  T* out = (T*)a.out;
  if (a.kind == 1) {
    const int64_t per = (int64_t)a.nT * a.nC;
    const int n = (int)(e / per);
    const int64_t r = e - (int64_t)n * per;
    const int c = (int)(r % a.nC);
    const int64_t ix = (int64_t)n * 2 * per + r, iy = ix + per;
    const double vx = (double)x[ix], vy = (double)x[iy];
    const double lim = (double)((const T*)a.lim)[(int64_t)n * a.lim_sn + (int64_t)c * a.lim_sc] - a.eps;
    const double rr = sqrt(vx * vx + vy * vy);
    const bool inside = !(lim / rr < 1.0);          // lim / 0 = inf: inside, like clamp_(max=1) upstream
    if (!a.adjoint) {
      const double sc = inside ? 1.0 : lim / rr;
      out[ix] = (T)(vx * sc);
      out[iy] = (T)(vy * sc);
    } else {
      const double gx = (double)g[ix], gy = (double)g[iy];
      if (inside) {
        out[ix] = (T)gx;
        out[iy] = (T)gy;
      } else {
        const double sc = lim / rr, rad = (gx * vx + gy * vy) / (rr * rr);
        out[ix] = (T)(sc * (gx - rad * vx));
        out[iy] = (T)(sc * (gy - rad * vy));
      }
    }
  } else {
    const int64_t row = e / a.nT;                    // (n, xyz) row
    const int n = (int)(row / 3), ax = (int)(row % 3);
    const T lim = ((const T*)a.lim)[(int64_t)n * a.lim_sn + (int64_t)ax * a.lim_sc];
    const T v = x[e];
    if (!a.adjoint) {
      out[e] = v > lim ? lim : (v < -lim ? -lim : v);
    } else {
      // s.max(-smax).min(smax): torch splits the gradient evenly on a tie
      const T w1 = v > -lim ? (T)1 : (v == -lim ? (T)0.5 : (T)0);
      const T m = v > -lim ? v : -lim;
      const T w2 = m < lim ? (T)1 : (m == lim ? (T)0.5 : (T)0);
      out[e] = g[e] * w1 * w2;
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256) mask_copy_kernel(const mrphy_mask_args a, const int64_t total) {
  const int64_t per = a.nOut * a.inner;
  for (int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x; e < total; e += (int64_t)gridDim.x * 256) {
    const int64_t n = e / per, r = e - n * per;
    const int64_t j = r / a.inner, k = r - j * a.inner;
    const int64_t src = a.idx[j];
    T v;
    if (src >= 0) v = ((const T*)a.in)[(n * a.nIn + src) * a.inner + k];
    else v = a.fill_zero ? (T)0 : (T)NAN;
    ((T*)a.out)[e] = v;
  }
}

}  // namespace mrphy

using namespace mrphy;

extern "C" int mrphy_design_waveform(const mrphy_reparam_args* a, void* cuda_stream) {
  launch_count() = 0;
  err_buf()[0] = 0;
  if (!a) return fail(MRPHY_ERR_ARG, "null args%s");
  if ((a->dtype != MRPHY_F32 && a->dtype != MRPHY_F64) || a->N < 1 || a->nT < 1 || a->nC < 1)
    return fail(MRPHY_ERR_ARG, "bad sizes or dtype%s");
  if (a->rf_kind < 0 || a->rf_kind > 2 || a->gr_kind < 0 || a->gr_kind > 3 || (a->rf_kind == 0 && a->gr_kind == 0))
    return fail(MRPHY_ERR_ARG, "rf_kind must be 0..2 and gr_kind 0..3, not both 0%s");
  if (a->rf_kind) {
    if (!a->rho || !a->theta || !a->rfmax) return fail(MRPHY_ERR_ARG, "rho, theta, rfmax are required%s");
    if (a->adjoint ? (!a->grf || !a->grho || !a->gtheta) : !a->rf) return fail(MRPHY_ERR_ARG, "rf half: output (or gradient) buffers missing%s");
  }
  if (a->gr_kind) {
    if (!a->ts) return fail(MRPHY_ERR_ARG, "ts is required%s");
    if (a->gr_kind != 2 && !a->smax) return fail(MRPHY_ERR_ARG, "smax is required%s");
    if (a->gr_kind != 3 && !a->dt.ptr) return fail(MRPHY_ERR_ARG, "dt is required%s");
    if (a->adjoint ? (!a->ggr || !a->gts) : !a->gr) return fail(MRPHY_ERR_ARG, "gradient half: output (or gradient) buffers missing%s");
  }
  const int64_t rows = a->gr_kind ? (int64_t)a->N * 3 : 0;
  const int64_t blocks = a->rf_kind ? ((int64_t)a->N * a->nT * a->nC + RP_THREADS - 1) / RP_THREADS : 0;
  if (rows + blocks > 2147483647LL) return fail(MRPHY_ERR_ARG, "too many elements for one launch%s");
  cudaStream_t st = (cudaStream_t)cuda_stream;
  if (a->dtype == MRPHY_F64) design_waveform_kernel<double><<<(unsigned)(rows + blocks), RP_THREADS, 0, st>>>(*a, (int)rows);
  else design_waveform_kernel<float><<<(unsigned)(rows + blocks), RP_THREADS, 0, st>>>(*a, (int)rows);
  ++launch_count();
  CK(cudaGetLastError());
  return MRPHY_OK;
}

extern "C" int mrphy_clamp_waveform(const mrphy_clamp_args* a, void* cuda_stream) {
  launch_count() = 0;
  err_buf()[0] = 0;
  if (!a) return fail(MRPHY_ERR_ARG, "null args%s");
  if ((a->dtype != MRPHY_F32 && a->dtype != MRPHY_F64) || a->N < 1 || a->nT < 1 || a->nC < 1 || (a->kind != 1 && a->kind != 2))
    return fail(MRPHY_ERR_ARG, "bad sizes, dtype or kind%s");
  if (!a->x || !a->lim || !a->out || (a->adjoint && !a->g)) return fail(MRPHY_ERR_ARG, "x, lim, out (and g for the adjoint) are required%s");
  const int64_t total = a->kind == 1 ? (int64_t)a->N * a->nT * a->nC : (int64_t)a->N * 3 * a->nT;
  const int64_t blocks = (total + 255) / 256;
  if (blocks > 2147483647LL) return fail(MRPHY_ERR_ARG, "too many elements for one launch%s");
  cudaStream_t st = (cudaStream_t)cuda_stream;
  if (a->dtype == MRPHY_F64) clamp_waveform_kernel<double><<<(unsigned)blocks, 256, 0, st>>>(*a, total);
  else clamp_waveform_kernel<float><<<(unsigned)blocks, 256, 0, st>>>(*a, total);
  ++launch_count();
  CK(cudaGetLastError());
  return MRPHY_OK;
}

extern "C" int mrphy_mask_copy(const mrphy_mask_args* a, void* cuda_stream) {
  launch_count() = 0;
  err_buf()[0] = 0;
  if (!a) return fail(MRPHY_ERR_ARG, "null args%s");
  if ((a->dtype != MRPHY_F32 && a->dtype != MRPHY_F64) || a->N < 1 || a->nOut < 0 || a->nIn < 0 || a->inner < 1)
    return fail(MRPHY_ERR_ARG, "bad sizes or dtype%s");
  const int64_t total = (int64_t)a->N * a->nOut * a->inner;
  if (total == 0) return MRPHY_OK;
  if (!a->idx || !a->in || !a->out) return fail(MRPHY_ERR_ARG, "idx, in, out are required%s");
  cudaStream_t st = (cudaStream_t)cuda_stream;
  const int64_t want = (total + 255) / 256;
  const unsigned grid = (unsigned)(want < 148 * 32 ? want : 148 * 32);   // grid-stride above 32 CTAs per SM
  if (a->dtype == MRPHY_F64) mask_copy_kernel<double><<<grid, 256, 0, st>>>(*a, total);
  else mask_copy_kernel<float><<<grid, 256, 0, st>>>(*a, total);
  ++launch_count();
  CK(cudaGetLastError());
  return MRPHY_OK;
}

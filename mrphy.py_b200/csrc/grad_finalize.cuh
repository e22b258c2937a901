// Deterministic second stage of the spin reductions, shared by the fused backward (blochsim_fused.cu) and the
// adjoint of rfgr2beff (aux_ops.cu): out[n][w][t] = sign * sum_p partials[n][p][w][t], summed in a fixed order
// (bitwise reproducible), written in the reference layout grf (N,2,nT[,nC]) / ggr (N,3,nT).
//
// grad_finalize_design_kernel is the same epilogue with the optimiser's re-parametrisation fused in: the
// CTA that finishes LAST for a batch entry (one atomic ticket per CTA) finds that entry's dL/drf, dL/dgr complete and runs
// the adjoint of rf = A(rho) rfmax (cos theta, sin theta), g = dt cumsum(atan(ts) 2/pi smax) on them (design_math.cuh),
// writing dL/drho, dL/dtheta, dL/dts -- the waveform gradients never make a round trip through another launch.
#pragma once
#include <cuda_runtime.h>

#include "design_math.cuh"

namespace mrphy {

template <typename T>
__device__ __forceinline__ void grad_finalize_tile(const T* __restrict__ partials, int P, int W, int NC, int nC, int nT,
                                                   int coil_dim, int bcast_coils, T sign, T* __restrict__ grf,
                                                   T* __restrict__ ggr, T (*sm)[33]) {
  constexpr int NY = 32;   // slices of the partial index summed in parallel, then combined in fixed order
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int t = blockIdx.x * 32 + tx, w = blockIdx.y, n = blockIdx.z;
  T sum = (T)0;
  if (t < nT) {
    const T* p = partials + ((size_t)n * P * W + w) * (size_t)nT + t;
    const size_t stride = (size_t)W * nT;
    int q = ty;
    // four independent loads in flight per thread (the grid is small: latency, not bandwidth, is what this kernel waits for);
    // the order of the additions is fixed by (P, ty) alone
    for (; q + 3 * NY < P; q += 4 * NY) {
      const T v0 = p[(size_t)q * stride], v1 = p[(size_t)(q + NY) * stride], v2 = p[(size_t)(q + 2 * NY) * stride],
              v3 = p[(size_t)(q + 3 * NY) * stride];
      sum += v0; sum += v1; sum += v2; sum += v3;
    }
    for (; q < P; q += NY) sum += p[(size_t)q * stride];
  }
  sm[ty][tx] = sum;
  __syncthreads();
  if (ty == 0 && t < nT) {
    T tot = sm[0][tx];
#pragma unroll
    for (int q = 1; q < NY; ++q) tot += sm[q][tx];
    tot *= sign;
    if (w >= 2 * NC) {
      ggr[((size_t)n * 3 + (w - 2 * NC)) * nT + t] = tot;
    } else {
      const int x = w / NC, coil = w % NC;
      const int nCo = coil_dim ? nC : 1;   // trailing dim of grf
      if (bcast_coils) {                   // no b1Map: every coil sees the same gradient
        for (int q = 0; q < nCo; ++q) grf[(((size_t)n * 2 + x) * nT + t) * nCo + q] = tot;
      } else if (coil < nC) {
        grf[(((size_t)n * 2 + x) * nT + t) * nCo + coil] = tot;
      }
    }
  }
}

// grid (ceil(nT / 32), W, N), block (32, 32).  Rows outside [w_lo, w_hi) were not asked for (and not produced).
template <typename T>
__global__ void __launch_bounds__(1024) grad_finalize_kernel(const T* __restrict__ partials, int P, int W, int NC, int nC,
                                                             int nT, int coil_dim, int bcast_coils, T sign,
                                                             T* __restrict__ grf, T* __restrict__ ggr, int w_lo = 0,
                                                             int w_hi = 1 << 30, T* __restrict__ tail = nullptr) {
  __shared__ T sm[32][33];
  if (tail && blockIdx.x + blockIdx.y + blockIdx.z == 0 && threadIdx.y == 0 && threadIdx.x < 4) tail[threadIdx.x] = (T)0;
  if ((int)blockIdx.y < w_lo || (int)blockIdx.y >= w_hi) return;
  grad_finalize_tile<T>(partials, P, W, NC, nC, nT, coil_dim, bcast_coils, sign, grf, ggr, sm);
}

// As above plus the design tail.  done[n] counts the finished CTAs of batch entry n: zero at launch, left zero again.
// d.adjoint != 0; d.grf / d.ggr are ignored (the tail reads this kernel's own outputs).
template <typename T>
__global__ void __launch_bounds__(1024) grad_finalize_design_kernel(const T* __restrict__ partials, int P, int W, int NC,
                                                                    int nC, int nT, int coil_dim, int bcast_coils, T sign,
                                                                    T* grf, T* ggr, int w_lo, int w_hi,
                                                                    const mrphy_reparam_args d, int* __restrict__ done,
                                                                    T* tail) {
  __shared__ T sm[32][33];
  __shared__ double warp_tot[32];
  __shared__ int s_last;
  if (tail && blockIdx.x + blockIdx.y + blockIdx.z == 0 && threadIdx.y == 0 && threadIdx.x < 4) tail[threadIdx.x] = (T)0;
  if ((int)blockIdx.y >= w_lo && (int)blockIdx.y < w_hi)
    grad_finalize_tile<T>(partials, P, W, NC, nC, nT, coil_dim, bcast_coils, sign, grf, ggr, sm);
  const int tid = threadIdx.y * 32 + threadIdx.x, n = blockIdx.z;
  __threadfence();     // this CTA's gradient samples are visible device-wide before the CTA is counted
  __syncthreads();
  if (tid == 0) {
    const int total = gridDim.x * gridDim.y;
    const int last = atomicAdd(&done[n], 1) == total - 1;
    if (last) done[n] = 0;   // ready for the next launch (CUDA-graph replays included)
    s_last = last;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();     // acquire side: every other CTA's samples are ordered before the ticket that made this one last
  design_adjoint_entry<T, 1024>(d, n, grf, ggr, warp_tot, tid);
}

}  // namespace mrphy

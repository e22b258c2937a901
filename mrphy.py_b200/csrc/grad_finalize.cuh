// Deterministic second stage of the spin reductions, shared by the fused backward (blochsim_fused.cu) and the
// adjoint of rfgr2beff (aux_ops.cu): out[n][w][t] = sign * sum_p partials[n][p][w][t], summed in a fixed order
// (bitwise reproducible), written in the reference layout grf (N,2,nT[,nC]) / ggr (N,3,nT).
#pragma once
#include <cuda_runtime.h>

namespace mrphy {

template <typename T>
__global__ void grad_finalize_kernel(const T* __restrict__ partials, int P, int W, int NC, int nC, int nT,
                                     int coil_dim, int bcast_coils, T sign, T* __restrict__ grf, T* __restrict__ ggr,
                                     int w_lo = 0, int w_hi = 1 << 30) {
  constexpr int NY = 32;   // slices of the partial index summed in parallel, then combined in fixed order
  __shared__ T sm[NY][33];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int t = blockIdx.x * 32 + tx, w = blockIdx.y, n = blockIdx.z;
  if (w < w_lo || w >= w_hi) return;   // rows the caller did not ask for (and the backward did not produce)
  T sum = (T)0;
  if (t < nT) {
    const T* p = partials + ((size_t)n * P * W + w) * (size_t)nT + t;
    for (int q = ty; q < P; q += NY) sum += p[(size_t)q * W * nT];
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

}  // namespace mrphy

// Validation + timing of the tensor-core transmit-field product (csrc/tc_helpers.cuh) before it goes into the kernels:
//   D[128 spins][N] = A[128][K] * B[N][K]^T   (K = 2 nC, N = 2 steps per chunk), TF32 tcgen05.mma with 3-way operand split,
// D read back per thread (= per spin = per TMEM lane) with tcgen05.ld.  Compares with an fp64 host product.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/tc_field profiles/ubench/tc_field.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../../mrphy.py_b200/csrc/tc_helpers.cuh"

using namespace mrphy;

template <int K, int N>
__global__ void __launch_bounds__(128) tc_field_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D,
                                                        int reps, int nprod) {
  constexpr int KC = K / 4;                        // 16-byte chunks along K
  extern __shared__ __align__(128) unsigned char dyn[];
  float (*sa)[KC][128][4] = reinterpret_cast<float (*)[KC][128][4]>(dyn);
  float (*sb)[KC][N][4] = reinterpret_cast<float (*)[KC][N][4]>(dyn + sizeof(float) * 3 * KC * 128 * 4);
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) tc::tmem_alloc<N>(&tmem_slot);
  // operands: thread s owns row s of A; rows of B are spread over the threads
  for (int e = 0; e < K; ++e) {
    float h, m, l;
    tc::split3(A[tid * K + e], h, m, l);
    sa[0][e / 4][tid][e % 4] = h; sa[1][e / 4][tid][e % 4] = m; sa[2][e / 4][tid][e % 4] = l;
  }
  for (int i = tid; i < N * K; i += 128) {
    const int n = i / K, e = i % K;
    float h, m, l;
    tc::split3(B[i], h, m, l);
    sb[0][e / 4][n][e % 4] = h; sb[1][e / 4][n][e % 4] = m; sb[2][e / 4][n][e % 4] = l;
  }
  fence_proxy_async_smem();      // generic-proxy writes -> visible to the tensor core (async proxy)
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tmem_slot;
  constexpr uint32_t idesc = tc::idesc_tf32(128, N);
  float acc[16];
  float sum = 0.f;
  for (int r = 0; r < reps; ++r) {
    if (tid == 0) {
      // small terms first: (lo,hi) (hi,lo) (mid,mid) (mid,hi) (hi,mid) (hi,hi)
      const int pa[6] = {2, 0, 1, 1, 0, 0}, pb[6] = {0, 2, 1, 0, 1, 0};
      bool first = true;
      for (int q = 6 - nprod; q < 6; ++q)
        for (int ks = 0; ks < K / 8; ++ks) {
          tc::mma_tf32(tmem, tc::kmajor_desc(&sa[pa[q]][2 * ks][0][0], 128), tc::kmajor_desc(&sb[pb[q]][2 * ks][0][0], N), idesc,
                       !first);
          first = false;
        }
      tc::commit(&bar);
    }
    mbar_wait(&bar, r & 1);
    tc::fence_after_sync();
    for (int c0 = 0; c0 < N; c0 += 16) {
      tc::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, acc);
      if (r == reps - 1) {
#pragma unroll
        for (int i = 0; i < 16; ++i) D[(size_t)(blockIdx.x * 128 + tid) * N + c0 + i] = acc[i];
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) sum += acc[i];
      }
    }
    tc::fence_before_sync();
    __syncthreads();           // every thread has read this round's D before the next MMA overwrites it
    tc::fence_after_sync();
  }
  if (sum == 123.456f) D[0] = sum;
  if (warp == 0) tc::tmem_free<N>(tmem);
}

template <int K, int N>
int run(int nprod) {
  std::vector<float> A(128 * K), B(N * K), D(128 * N);
  srand(1);
  for (auto& x : A) x = (float)rand() / RAND_MAX * 2 - 1;
  for (auto& x : B) x = ((float)rand() / RAND_MAX * 2 - 1) * 0.1f;
  float *dA, *dB, *dD;
  const int grid = 148 * 3;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, (size_t)grid * D.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  const size_t smem = sizeof(float) * 3 * (K / 4) * (128 + N) * 4;
  cudaFuncSetAttribute(tc_field_kernel<K, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  tc_field_kernel<K, N><<<1, 128, smem>>>(dA, dB, dD, 1, nprod);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("K=%d N=%d: CUDA error %s\n", K, N, cudaGetErrorString(e)); return 1; }
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0, maxref = 0, err32 = 0;
  for (int s = 0; s < 128; ++s)
    for (int n = 0; n < N; ++n) {
      double ref = 0; float f = 0.f;
      for (int k = 0; k < K; ++k) { ref += (double)A[s * K + k] * (double)B[n * K + k]; f = fmaf(A[s * K + k], B[n * K + k], f); }
      maxerr = fmax(maxerr, fabs(ref - (double)D[s * N + n]));
      err32 = fmax(err32, fabs(ref - (double)f));
      maxref = fmax(maxref, fabs(ref));
    }
  // timing: many chunks per CTA, 3 CTAs per SM
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int reps = 2000;
  tc_field_kernel<K, N><<<grid, 128, smem>>>(dA, dB, dD, reps, nprod);
  cudaEventRecord(e0);
  tc_field_kernel<K, N><<<grid, 128, smem>>>(dA, dB, dD, reps, nprod);
  cudaEventRecord(e1);
  e = cudaDeviceSynchronize();
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
  const double fields = (double)grid * reps * 128.0 * (N / 2);       // (spin, step) pairs
  printf("K=%2d (nC=%2d) N=%3d products=%d: max|D - fp64| = %.3e (fp32 FMA chain: %.3e, max|D| = %.2f)   %.3f ms for %d chunks/CTA x %d CTAs"
         " = %.3e spin-steps/s field-only (%s)\n", K, K / 2, N, nprod, maxerr, err32, maxref, ms, reps, grid, fields / (ms * 1e-3),
         e == cudaSuccess ? "ok" : cudaGetErrorString(e));
  cudaFree(dA); cudaFree(dB); cudaFree(dD);
  return 0;
}

// The field product with A in TMEM (thread t stores its own row with tcgen05.st: no shared memory for A), two parts, 3 products.
template <int K, int N>
__global__ void __launch_bounds__(128) tc_field_ts_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D) {
  constexpr int KC = K / 4;
  extern __shared__ __align__(128) unsigned char dyn[];
  float (*sb)[KC][N][4] = reinterpret_cast<float (*)[KC][N][4]>(dyn);     // [2 parts][KC][N][4]
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  constexpr int COLS = (N + 2 * K) <= 128 ? 128 : 256;                    // D: N columns, then A hi (K), A mid (K)
  if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) tc::tmem_alloc<COLS>(&tmem_slot);
  for (int i = tid; i < N * K; i += 128) {
    const int n = i / K, e = i % K;
    const float h = tc::tf32_rn(B[i]), m = tc::tf32_rn(B[i] - h);
    sb[0][e / 4][n][e % 4] = h; sb[1][e / 4][n][e % 4] = m;
  }
  fence_proxy_async_smem();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tmem_slot, tlane = tmem + ((uint32_t)(warp * 32) << 16);
  for (int e0 = 0; e0 < K; e0 += 8) {          // this thread's row of A, both parts, 8 columns at a time
    float h[8], m[8];
    for (int e = 0; e < 8; ++e) { const float x = A[tid * K + e0 + e]; h[e] = tc::tf32_rn(x); m[e] = tc::tf32_rn(x - h[e]); }
    tc::tmem_st8(tlane + N + e0, h);
    tc::tmem_st8(tlane + N + K + e0, m);
  }
  tc::tmem_st_wait();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  if (tid == 0) {
    constexpr uint32_t idesc = tc::idesc_tf32(128, N);
    const int pa[3] = {1, 0, 0}, pb[3] = {0, 1, 0};
    bool acc = false;
    for (int q = 0; q < 3; ++q)
      for (int ks = 0; ks < K / 8; ++ks) {
        tc::mma_tf32_ts(tmem, tmem + N + pa[q] * K + 8 * ks, tc::kmajor_desc(&sb[pb[q]][2 * ks][0][0], N), idesc, acc);
        acc = true;
      }
    tc::commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc::fence_after_sync();
  float acc16[16];
  for (int c0 = 0; c0 < N; c0 += 16) {
    tc::tmem_ld16(tlane + c0, acc16);
#pragma unroll
    for (int i = 0; i < 16; ++i) D[(size_t)tid * N + c0 + i] = acc16[i];
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_free<COLS>(tmem);
}

template <int K, int N>
int run_ts() {
  std::vector<float> A(128 * K), B(N * K), D(128 * N);
  srand(1);
  for (auto& x : A) x = (float)rand() / RAND_MAX * 2 - 1;
  for (auto& x : B) x = ((float)rand() / RAND_MAX * 2 - 1) * 0.1f;
  float *dA, *dB, *dD;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  const size_t smem = sizeof(float) * 2 * (K / 4) * N * 4;
  cudaFuncSetAttribute(tc_field_ts_kernel<K, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  tc_field_ts_kernel<K, N><<<1, 128, smem>>>(dA, dB, dD);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("A in TMEM K=%d N=%d: CUDA error %s\n", K, N, cudaGetErrorString(e)); return 1; }
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0, err32 = 0;
  for (int s = 0; s < 128; ++s)
    for (int n = 0; n < N; ++n) {
      double ref = 0; float f = 0.f;
      for (int k = 0; k < K; ++k) { ref += (double)A[s * K + k] * (double)B[n * K + k]; f = fmaf(A[s * K + k], B[n * K + k], f); }
      maxerr = fmax(maxerr, fabs(ref - (double)D[s * N + n]));
      err32 = fmax(err32, fabs(ref - (double)f));
    }
  printf("field product with A in TMEM (tcgen05.st rows, 2 parts, 3 products), K=%d N=%d: max|D - fp64| = %.3e (fp32 FMA chain: %.3e) -> %s\n",
         K, N, maxerr, err32, maxerr < 4 * err32 + 1e-7 ? "OK" : "WRONG");
  cudaFree(dA); cudaFree(dB); cudaFree(dD);
  return maxerr < 4 * err32 + 1e-7 ? 0 : 1;
}

// The spin reduction of the adjoint as a product (not shipped): D2[e][n] = sum_s Wt[s][e] * F[n][s], K = 128 spins.
// B2 = F^T K-major (tile[chunk s/4][row n][4]).  A2 = the weights, either re-laid out spin-minor (K-major, amn = 0: works) or as the
// MN-MAJOR view of the tile the field product reads (tile[chunk e/4][spin][4], amn = 1): an unswizzled MN-major TF32 descriptor
// is accepted but produces zeros -- the reduce product would need its own copy of the weights.
template <int N>
__global__ void __launch_bounds__(128) tc_reduce_kernel(const float* __restrict__ Wt /*[128 spins][128]*/, const float* __restrict__ F /*[N][128 spins]*/,
                                                        float* __restrict__ D /*[128][N]*/, uint32_t lbo, uint32_t sbo, uint32_t kstep, int amn) {
  extern __shared__ __align__(128) unsigned char dyn[];
  float (*sa)[128][4] = reinterpret_cast<float (*)[128][4]>(dyn);                            // [32 chunks of e][128 spins][4]
  float (*sb)[N][4] = reinterpret_cast<float (*)[N][4]>(dyn + sizeof(float) * 32 * 128 * 4);  // [32 chunks of s][N rows][4]
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  constexpr int COLS = N <= 32 ? 32 : N <= 64 ? 64 : 128;
  if (warp == 0) tc::tmem_alloc<COLS>(&tmem_slot);
  for (int e = 0; e < 128; ++e) {       // thread = spin
    if (amn) sa[e / 4][tid][e % 4] = tc::tf32_rn(Wt[tid * 128 + e]);
    else sa[tid / 4][e][tid % 4] = tc::tf32_rn(Wt[tid * 128 + e]);   // K-major control: tile[chunk s/4][row e][4 spins]
  }
  for (int i = tid; i < N * 128; i += 128) { const int n = i / 128, sp = i % 128; sb[sp / 4][n][sp % 4] = tc::tf32_rn(F[i]); }
  fence_proxy_async_smem();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tmem_slot;
  if (tid == 0) {
    const uint32_t idesc = tc::idesc_tf32(128, N, /*a_mn=*/amn != 0, /*b_mn=*/false);
    for (int ks = 0; ks < 16; ++ks)     // 8 spins per instruction: 8 rows of A2's tile, 2 chunks of B2's
    {
      uint64_t d = (uint64_t)(((smem_u32(&sa[0][0][0]) + ks * kstep) >> 4) & 0x3fff) | (uint64_t)((lbo >> 4) & 0x3fff) << 16 |
                   (uint64_t)((sbo >> 4) & 0x3fff) << 32 | (uint64_t)1 << 46;
      if (!amn) d = tc::kmajor_desc(&sa[2 * ks][0][0], 128);
      tc::mma_tf32(tmem, d, tc::kmajor_desc(&sb[2 * ks][0][0], N), idesc, ks > 0);
    }
    tc::commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc::fence_after_sync();
  float acc[16];
  for (int c0 = 0; c0 < N; c0 += 16) {
    tc::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, acc);
#pragma unroll
    for (int i = 0; i < 16; ++i) D[(size_t)tid * N + c0 + i] = acc[i];
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_free<COLS>(tmem);
}

template <int N>
int run_reduce(uint32_t lbo, uint32_t sbo, uint32_t kstep, int amn = 1) {
  std::vector<float> W(128 * 128), F(N * 128), D(128 * N);
  srand(2);
  for (auto& x : W) x = (float)rand() / RAND_MAX * 2 - 1;
  for (auto& x : F) x = (float)rand() / RAND_MAX * 2 - 1;
  float *dW, *dF, *dD;
  cudaMalloc(&dW, W.size() * 4); cudaMalloc(&dF, F.size() * 4); cudaMalloc(&dD, D.size() * 4);
  cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dF, F.data(), F.size() * 4, cudaMemcpyHostToDevice);
  const size_t smem = sizeof(float) * 32 * (128 + N) * 4;
  cudaFuncSetAttribute(tc_reduce_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  tc_reduce_kernel<N><<<1, 128, smem>>>(dW, dF, dD, lbo, sbo, kstep, amn);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("reduce N=%d: CUDA error %s\n", N, cudaGetErrorString(e)); return 1; }
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0, maxref = 0;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < N; ++n) {
      double ref = 0;
      for (int sp = 0; sp < 128; ++sp) ref += (double)W[sp * 128 + m] * (double)F[n * 128 + sp];
      maxerr = fmax(maxerr, fabs(ref - (double)D[m * N + n]));
      maxref = fmax(maxref, fabs(ref));
    }
  printf("reduce product D2 = Wt^T F, N=%d, K=128 spins, LBO=%u SBO=%u kstep=%u amn=%d, single TF32 part: max|D - fp64| = %.3e (max|D| = %.2f)"
         " -> descriptor %s   D[0][0..2] = %.3f %.3f %.3f\n", N, lbo, sbo, kstep, amn, maxerr, maxref, maxerr < 2e-2 ? "OK (error = TF32 rounding of the operands)" : "NO RESULT", D[0], D[1], D[2]);
  cudaFree(dW); cudaFree(dF); cudaFree(dD);
  return maxerr < 2e-2 ? 0 : 1;
}

int main() {
  int rc = 0;
  rc |= run<8, 128>(6);
  rc |= run<16, 128>(6);
  rc |= run<16, 128>(3);
  rc |= run<16, 128>(1);
  rc |= run<32, 128>(6);
  rc |= run<16, 64>(6);
  rc |= run<32, 64>(6);
  rc |= run_ts<16, 64>();
  rc |= run_ts<32, 64>();
  rc |= run_ts<8, 16>();
  rc |= run_reduce<64>(0, 0, 0, 0);          // control: weights re-laid out spin-minor (K-major A)
  run_reduce<64>(128, 2048, 128, 1);         // the field product's tile read MN-major, unswizzled: all zeros on B200 (TF32 MN-major
  run_reduce<64>(2048, 128, 128, 1);         // operands exist only in the 128B_BASE32B swizzled layout); either LBO/SBO assignment
  return rc;
}

// Fused Bloch simulation for sm_100a: waveform -> Beff -> (u,phi) -> Rodrigues -> relaxation for
// all nT steps with the magnetisation in registers, and its adjoint with on-chip reduction of
// dL/drf, dL/dgr over spins.  Replaces beffective.rfgr2beff (beffective.py:107-168) +
// sims.BlochSim.forward/backward (sims.py:31-269) + the autograd of rfgr2beff.
//
// Kernels in this file
//   pack_waveform_kernel      rf (N,2,nT[,nC]), gr (N,3,nT) -> wave[N][chunk][W][TCP]  (W = 2*NC+3)
//   fused_fwd_kernel<T,..>    one spin (or S spins) per thread; checkpoint every K steps
//   fused_bwd_kernel<T,..>    time-reversed state reconstruction + adjoint + spin reduction
//   grad_finalize_kernel<T>   deterministic sum of the per-CTA partials, reference layout out
//
// Data layout in HBM (T = float | double):
//   wave      [N][nChunks][W][TCP]   one chunk = K steps (TCP = K rounded up to 4): a chunk is
//                                    ONE contiguous 16-byte-aligned block -> one 1-D TMA bulk copy
//   ckpt      [N][nChunks-1][3][nM]  state after (c+1)*K steps, SoA so warps store 128-B lines
//   partials  [N][P][W][nT]          per-CTA gradient partial sums (P CTAs per batch entry)
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/mrphy_b200.h"
#include "abi_common.cuh"
#include "bloch_math.cuh"
#include "ptx_helpers.cuh"

namespace mrphy {

constexpr int TCMAX = 64;   // max steps per staged waveform chunk (== max checkpoint interval)
constexpr int BLK = 128;    // threads per CTA
constexpr int NWARP = BLK / 32;

// steps per gradient-reduction tile: the per-warp transposition tile [W][TR][32] is kept <= 10 KB
constexpr int pick_tr(int W, int elem) {
  int tr = 16;
  while (tr > 1 && W * tr * 32 * elem > 10240) tr >>= 1;
  return tr;
}

// dynamic shared memory layout of the backward kernel
template <typename T, int NC> struct BwdSmem {
  static constexpr int W = 2 * NC + 3;
  static constexpr int TR = pick_tr(W, (int)sizeof(T));
  static constexpr size_t wbuf = 0;                                                // T[2][W*TCMAX]
  static constexpr size_t red = (2 * W * TCMAX * sizeof(T) + 127) / 128 * 128;     // T[NWARP][W][TR][32]
  static constexpr size_t cta = red + (size_t)NWARP * W * TR * 32 * sizeof(T);     // T[NWARP][W][TR]
  static constexpr size_t bar = (cta + (size_t)NWARP * W * TR * sizeof(T) + 15) / 16 * 16;   // uint64_t[2]
  static constexpr size_t bytes = bar + 16;
};

template <typename T> struct KArgs {
  int N, nM, nT, K, TCP, nChunks, P;
  const T* Mi; int64_t Mi_sn, Mi_sm;
  const T* loc; int64_t loc_sn, loc_sm;
  const T* b1; int64_t b1_sn, b1_sm; int nC;
  mrphy_param df, T1, T2, gamma, dt;
  T* Mo;
  T* ckpt;
  const T* wave;
  const T* gMo; int64_t gMo_sn, gMo_sm;
  T* gMi;
  T* partials;
};

__device__ __forceinline__ void load4(const float* p, float (&v)[4]) {
  const float4 q = *reinterpret_cast<const float4*>(p);
  v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
}
__device__ __forceinline__ void load4(const double* p, double (&v)[4]) {
  const double2 a = *reinterpret_cast<const double2*>(p);
  const double2 b = *reinterpret_cast<const double2*>(p + 2);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}

// per-spin prologue shared by forward and backward
template <typename T, int NC, bool RELAX>
__device__ __forceinline__ void load_spin(const KArgs<T>& a, int n, int i, SpinConst<T, NC>& k) {
  const T* lp = a.loc + (int64_t)n * a.loc_sn + (int64_t)i * a.loc_sm;
  T br[NC], bi[NC];
  if (a.b1) {
    const T* bp = a.b1 + (int64_t)n * a.b1_sn + (int64_t)i * a.b1_sm;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      br[c] = c < a.nC ? bp[c] : (T)0;
      bi[c] = c < a.nC ? bp[a.nC + c] : (T)0;
    }
  }
  const double gam = ld_param(a.gamma, n, i);
  const double dt = ld_param(a.dt, n, 0);
  const double df = a.df.ptr ? ld_param(a.df, n, i) : 0.0;
  const double t1 = RELAX ? ld_param(a.T1, n, i) : 1.0;
  const double t2 = RELAX ? ld_param(a.T2, n, i) : 1.0;
  make_consts<T, NC>(k, gam, dt, RELAX, t1, t2, df, lp[0], lp[1], lp[2], a.b1 ? br : nullptr, a.b1 ? bi : nullptr);
}

// ------------------------------------------------------------------------------------------
// pack: one thread per element of wave[N][nChunks][W][TCP]
template <typename T>
__global__ void pack_waveform_kernel(const T* __restrict__ rf, int64_t rf_sn, int64_t rf_sx, int64_t rf_st, int64_t rf_sc,
                                     const T* __restrict__ gr, int64_t gr_sn, int64_t gr_sx, int64_t gr_st, int nC,
                                     int NC, int sum_coils, int nT, int K, int TCP, int nChunks, int64_t total,
                                     T* __restrict__ wave) {
  const int W = 2 * NC + 3;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int j = (int)(e % TCP);
    const int w = (int)((e / TCP) % W);
    const int c = (int)((e / ((int64_t)TCP * W)) % nChunks);
    const int n = (int)(e / ((int64_t)TCP * W * nChunks));
    const int t = c * K + j;
    T v = (T)0;
    if (j < K && t < nT) {
      if (w < 2 * NC) {
        const int x = w / NC, coil = w % NC;
        const T* p = rf + n * rf_sn + x * rf_sx + t * rf_st;
        if (sum_coils) {   // no b1Map: beffective.py:147-151 sums rf over coils
          for (int q = 0; q < nC; ++q) v += p[q * rf_sc];
        } else if (coil < nC) {
          v = p[coil * rf_sc];
        }
      } else {
        v = gr[n * gr_sn + (w - 2 * NC) * gr_sx + t * gr_st];
      }
    }
    wave[e] = v;
  }
}

// ------------------------------------------------------------------------------------------
// forward
template <typename T, int POL, bool RELAX, int NC, int S>
__global__ void __launch_bounds__(BLK) fused_fwd_kernel(const KArgs<T> a) {
  constexpr int W = 2 * NC + 3;
  __shared__ __align__(128) T wbuf[2][W * TCMAX];
  __shared__ __align__(8) uint64_t full[2];
  const int tid = threadIdx.x, n = blockIdx.y;
  const int TCP = a.TCP, K = a.K, nT = a.nT, nChunks = a.nChunks, nM = a.nM;
  const uint32_t chunk_bytes = (uint32_t)(W * TCP * sizeof(T));
  const T* wave_n = a.wave + (size_t)n * nChunks * W * TCP;
  const int tiles = (nM + BLK * S - 1) / (BLK * S);
  const int my_tiles = ((int)blockIdx.x < tiles) ? (tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const uint32_t total = (uint32_t)my_tiles * (uint32_t)nChunks;
  if (tid == 0) {
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (tid == 0 && total > 0) {
    mbar_arrive_expect_tx(&full[0], chunk_bytes);
    bulk_g2s(wbuf[0], wave_n, chunk_bytes, &full[0]);
  }
  uint32_t it = 0;
  for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    SpinConst<T, NC> k[S];
    T mx[S], my[S], mz[S];
    int idx[S];
    bool ok[S];
#pragma unroll
    for (int s = 0; s < S; ++s) {
      const int i = (tile * S + s) * BLK + tid;
      ok[s] = i < nM;
      idx[s] = ok[s] ? i : nM - 1;
      load_spin<T, NC, RELAX>(a, n, idx[s], k[s]);
      const T* mp = a.Mi + (int64_t)n * a.Mi_sn + (int64_t)idx[s] * a.Mi_sm;
      mx[s] = mp[0]; my[s] = mp[1]; mz[s] = mp[2];
    }
    for (int c = 0; c < nChunks; ++c, ++it) {
      if (tid == 0 && it + 1 < total) {   // prefetch the next chunk (possibly chunk 0 of the next tile)
        const int cn = (c + 1 == nChunks) ? 0 : c + 1;
        const uint32_t sn = (it + 1) & 1;
        mbar_arrive_expect_tx(&full[sn], chunk_bytes);
        bulk_g2s(wbuf[sn], wave_n + (size_t)cn * W * TCP, chunk_bytes, &full[sn]);
      }
      mbar_wait(&full[it & 1], (it >> 1) & 1);
      const T* wb = wbuf[it & 1];
      const int ns = min(K, nT - c * K);
      int j = 0;
      // 4 steps per iteration with 128-bit broadcast loads, while the staged samples fit ~40 registers
      constexpr bool VEC4 = W * sizeof(T) <= 44;
      for (; VEC4 && j + 4 <= ns; j += 4) {
        T wv[W][4];
#pragma unroll
        for (int w = 0; w < W; ++w) load4(wb + w * TCP + j, wv[w]);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          T rx[NC], ry[NC];
#pragma unroll
          for (int q = 0; q < NC; ++q) { rx[q] = wv[q][u]; ry[q] = wv[NC + q][u]; }
#pragma unroll
          for (int s = 0; s < S; ++s) {
            T bx, by, bz;
            field<T, NC>(k[s], rx, ry, wv[2 * NC][u], wv[2 * NC + 1][u], wv[2 * NC + 2][u], bx, by, bz);
            step_fwd<T, POL, RELAX>(bx, by, bz, k[s].e1, k[s].e2, mx[s], my[s], mz[s]);
          }
        }
      }
      for (; j < ns; ++j) {
        T rx[NC], ry[NC];
#pragma unroll
        for (int q = 0; q < NC; ++q) { rx[q] = wb[q * TCP + j]; ry[q] = wb[(NC + q) * TCP + j]; }
        const T gx = wb[2 * NC * TCP + j], gy = wb[(2 * NC + 1) * TCP + j], gz = wb[(2 * NC + 2) * TCP + j];
#pragma unroll
        for (int s = 0; s < S; ++s) {
          T bx, by, bz;
          field<T, NC>(k[s], rx, ry, gx, gy, gz, bx, by, bz);
          step_fwd<T, POL, RELAX>(bx, by, bz, k[s].e1, k[s].e2, mx[s], my[s], mz[s]);
        }
      }
      if (c + 1 < nChunks) {   // checkpoint: state after (c+1)*K steps
        T* cp = a.ckpt + ((size_t)n * (nChunks - 1) + c) * 3 * (size_t)nM;
#pragma unroll
        for (int s = 0; s < S; ++s)
          if (ok[s]) {
            cp[idx[s]] = mx[s];
            cp[(size_t)nM + idx[s]] = my[s];
            cp[2 * (size_t)nM + idx[s]] = mz[s];
          }
      }
      __syncthreads();   // everyone is done with wbuf[it&1] before it is refilled
    }
#pragma unroll
    for (int s = 0; s < S; ++s)
      if (ok[s]) {
        T* op = a.Mo + ((size_t)n * nM + idx[s]) * 3;
        op[0] = mx[s]; op[1] = my[s]; op[2] = mz[s];
      }
  }
}

// ------------------------------------------------------------------------------------------
// backward
template <typename T, int POL, bool RELAX, int NC, int S>
__global__ void __launch_bounds__(BLK) fused_bwd_kernel(const KArgs<T> a, const int need_gmi) {
  using L = BwdSmem<T, NC>;
  constexpr int W = L::W, TR = L::TR;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  T(*wbuf)[W * TCMAX] = reinterpret_cast<T(*)[W * TCMAX]>(smem_raw + L::wbuf);
  T(*red)[W][TR][32] = reinterpret_cast<T(*)[W][TR][32]>(smem_raw + L::red);   // per-warp transposition tile
  T(*cta)[W][TR] = reinterpret_cast<T(*)[W][TR]>(smem_raw + L::cta);          // per-warp tile sums
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + L::bar);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, n = blockIdx.y;
  const int TCP = a.TCP, K = a.K, nT = a.nT, nChunks = a.nChunks, nM = a.nM;
  const uint32_t chunk_bytes = (uint32_t)(W * TCP * sizeof(T));
  const T* wave_n = a.wave + (size_t)n * nChunks * W * TCP;
  const int tiles = (nM + BLK * S - 1) / (BLK * S);
  const int my_tiles = ((int)blockIdx.x < tiles) ? (tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const uint32_t total = (uint32_t)my_tiles * (uint32_t)nChunks;
  T* part = a.partials + ((size_t)n * a.P + blockIdx.x) * W * (size_t)nT;
  if (tid == 0) {
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (tid == 0 && total > 0) {
    mbar_arrive_expect_tx(&full[0], chunk_bytes);
    bulk_g2s(wbuf[0], wave_n + (size_t)(nChunks - 1) * W * TCP, chunk_bytes, &full[0]);
  }
  if (my_tiles == 0) {   // a CTA without work still owns a partial slot: zero it
    for (int e = tid; e < W * nT; e += BLK) part[e] = (T)0;
    return;
  }
  uint32_t it = 0;
  bool first = true;
  for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, first = false) {
    SpinConst<T, NC> k[S];
    T mx[S], my[S], mz[S], hx[S], hy[S], hz[S];
    int idx[S];
    bool ok[S];
#pragma unroll
    for (int s = 0; s < S; ++s) {
      const int i = (tile * S + s) * BLK + tid;
      ok[s] = i < nM;
      idx[s] = ok[s] ? i : nM - 1;
      load_spin<T, NC, RELAX>(a, n, idx[s], k[s]);
      const T* mp = a.Mo + ((size_t)n * nM + idx[s]) * 3;
      mx[s] = mp[0]; my[s] = mp[1]; mz[s] = mp[2];
      const T* gp = a.gMo + (int64_t)n * a.gMo_sn + (int64_t)idx[s] * a.gMo_sm;
      hx[s] = ok[s] ? gp[0] : (T)0;   // padding lanes carry a zero adjoint: they add nothing
      hy[s] = ok[s] ? gp[1] : (T)0;
      hz[s] = ok[s] ? gp[2] : (T)0;
    }
    for (int c = nChunks - 1; c >= 0; --c, ++it) {
      if (tid == 0 && it + 1 < total) {
        const int cn = (c == 0) ? nChunks - 1 : c - 1;
        const uint32_t sn = (it + 1) & 1;
        mbar_arrive_expect_tx(&full[sn], chunk_bytes);
        bulk_g2s(wbuf[sn], wave_n + (size_t)cn * W * TCP, chunk_bytes, &full[sn]);
      }
      // prefetch the checkpoint this chunk ends on (state after c*K steps) while we compute
      T kx[S], ky[S], kz[S];
      if (c > 0) {
        const T* cp = a.ckpt + ((size_t)n * (nChunks - 1) + (c - 1)) * 3 * (size_t)nM;
#pragma unroll
        for (int s = 0; s < S; ++s) {
          kx[s] = cp[idx[s]];
          ky[s] = cp[(size_t)nM + idx[s]];
          kz[s] = cp[2 * (size_t)nM + idx[s]];
        }
      }
      mbar_wait(&full[it & 1], (it >> 1) & 1);
      const T* wb = wbuf[it & 1];
      const int ns = min(K, nT - c * K);
      for (int j1 = ns; j1 > 0;) {
        const int j0 = ((j1 - 1) / TR) * TR;   // tile [j0, j1), at most TR steps
        for (int j = j1 - 1; j >= j0; --j) {
          T rx[NC], ry[NC];
#pragma unroll
          for (int q = 0; q < NC; ++q) { rx[q] = wb[q * TCP + j]; ry[q] = wb[(NC + q) * TCP + j]; }
          const T gx = wb[2 * NC * TCP + j], gy = wb[(2 * NC + 1) * TCP + j], gz = wb[(2 * NC + 2) * TCP + j];
          T acc[W];
#pragma unroll
          for (int w = 0; w < W; ++w) acc[w] = (T)0;
#pragma unroll
          for (int s = 0; s < S; ++s) {
            T bx, by, bz, Fx, Fy, Fz;
            field<T, NC>(k[s], rx, ry, gx, gy, gz, bx, by, bz);
            step_bwd<T, POL, RELAX, NC>(k[s], bx, by, bz, mx[s], my[s], mz[s], hx[s], hy[s], hz[s], Fx, Fy, Fz);
#pragma unroll
            for (int q = 0; q < NC; ++q) {
              acc[q] = fma_(k[s].cbr[q], Fx, fma_(k[s].cbi[q], Fy, acc[q]));
              acc[NC + q] = fma_(k[s].cbr[q], Fy, fma_(-k[s].cbi[q], Fx, acc[NC + q]));
            }
            acc[2 * NC] = fma_(k[s].glx, Fz, acc[2 * NC]);
            acc[2 * NC + 1] = fma_(k[s].gly, Fz, acc[2 * NC + 1]);
            acc[2 * NC + 2] = fma_(k[s].glz, Fz, acc[2 * NC + 2]);
          }
#pragma unroll
          for (int w = 0; w < W; ++w) red[warp][w][j - j0][lane] = acc[w];
        }
        __syncwarp();
        {   // lane l owns row r = l%TR and sums the TR source lanes of its group (rotated start:
            // conflict-free), then the 32/TR groups meet through shuffles
          const int r = lane & (TR - 1), grp = lane & ~(TR - 1);
#pragma unroll
          for (int w = 0; w < W; ++w) {
            T sum = (T)0;
#pragma unroll
            for (int q = 0; q < TR; ++q) sum += red[warp][w][r][grp + ((q + lane) & (TR - 1))];
#pragma unroll
            for (int o = TR; o < 32; o <<= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            if (lane < TR) cta[warp][w][r] = sum;
          }
        }
        __syncthreads();
        for (int e = tid; e < W * TR; e += BLK) {   // fixed-order combine of the warps
          const int w = e / TR, r = e % TR;
          if (j0 + r < j1) {
            T sum = cta[0][w][r];
#pragma unroll
            for (int q = 1; q < NWARP; ++q) sum += cta[q][w][r];
            T* dst = part + (size_t)w * nT + (c * K + j0 + r);
            *dst = first ? sum : *dst + sum;
          }
        }
        __syncthreads();
        j1 = j0;
      }
      if (c > 0) {   // resynchronise the reconstructed state with the forward checkpoint
#pragma unroll
        for (int s = 0; s < S; ++s) { mx[s] = kx[s]; my[s] = ky[s]; mz[s] = kz[s]; }
      }
    }
    if (need_gmi) {
#pragma unroll
      for (int s = 0; s < S; ++s)
        if (ok[s]) {
          T* op = a.gMi + ((size_t)n * nM + idx[s]) * 3;
          op[0] = hx[s]; op[1] = hy[s]; op[2] = hz[s];
        }
    }
  }
}

// ------------------------------------------------------------------------------------------
// finalize: out[n][w][t] = - sum_p partials[n][p][w][t]   (fixed order => bitwise reproducible)
template <typename T>
__global__ void grad_finalize_kernel(const T* __restrict__ partials, int P, int W, int NC, int nC, int nT,
                                     int coil_dim, int bcast_coils, T* __restrict__ grf, T* __restrict__ ggr) {
  __shared__ T sm[8][33];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int t = blockIdx.x * 32 + tx, w = blockIdx.y, n = blockIdx.z;
  T sum = (T)0;
  if (t < nT) {
    const T* p = partials + ((size_t)n * P * W + w) * (size_t)nT + t;
    for (int q = ty; q < P; q += 8) sum += p[(size_t)q * W * nT];
  }
  sm[ty][tx] = sum;
  __syncthreads();
  if (ty == 0 && t < nT) {
    T tot = sm[0][tx];
#pragma unroll
    for (int q = 1; q < 8; ++q) tot += sm[q][tx];
    tot = -tot;
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

// =============================================================================================
// C ABI
using namespace mrphy;

static thread_local char g_err[512] = "";
static thread_local int g_launches = 0;
// bench-only, process-wide on purpose: backward runs on the autograd engine's thread
static int g_timing = 0;
static cudaEvent_t g_ev0 = nullptr, g_ev1 = nullptr;
static int g_ev_valid = 0;
namespace mrphy {
char* err_buf() { return g_err; }
int& launch_count() { return g_launches; }
void timing_begin(cudaStream_t st) {
  if (!g_timing) return;
  if (!g_ev0 && (cudaEventCreate(&g_ev0) != cudaSuccess || cudaEventCreate(&g_ev1) != cudaSuccess)) {
    cudaGetLastError();
    g_ev0 = g_ev1 = nullptr;
    return;
  }
  g_ev_valid = 0;
  cudaEventRecord(g_ev0, st);
}
void timing_end(cudaStream_t st) {
  if (!g_timing || !g_ev0) return;
  cudaEventRecord(g_ev1, st);
  g_ev_valid = 1;
}
}  // namespace mrphy

extern "C" int mrphy_kernel_timing(int enable) {
  g_timing = enable ? 1 : 0;
  g_ev_valid = 0;
  return MRPHY_OK;
}
extern "C" float mrphy_last_kernel_ms(void) {
  if (!g_ev_valid) return -1.0f;
  float ms = -1.0f;
  if (cudaEventSynchronize(g_ev1) != cudaSuccess || cudaEventElapsedTime(&ms, g_ev0, g_ev1) != cudaSuccess) {
    cudaGetLastError();
    return -1.0f;
  }
  return ms;
}

extern "C" int mrphy_abi_version(void) { return MRPHY_ABI_VERSION; }
extern "C" const char* mrphy_last_error(void) { return g_err; }
extern "C" int mrphy_last_launch_count(void) { return g_launches; }

extern "C" int mrphy_device_sm_count(int device) {
  int v = 0;
  if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) {
    fail(MRPHY_ERR_CUDA, "cudaDeviceGetAttribute failed%s");
    cudaGetLastError();
    return MRPHY_ERR_CUDA;
  }
  return v;
}

namespace {

struct Plan {
  int NC;        // coils held in registers (1,2,4,8)
  int S;         // spins per thread
  int K, TCP, nChunks, W;
  int sum_coils; // no b1Map
  int tiles, P;
};

int sm_count_cached() {
  static int sms = -1;
  if (sms < 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) {
      cudaGetLastError();
      sms = 148;   // B200; only used for grid sizing
    }
  }
  return sms;
}

int make_plan(const mrphy_fused_args* a, Plan* p, bool need_device) {
  if (!a) return fail(MRPHY_ERR_ARG, "null args%s");
  if (a->dtype != MRPHY_F32 && a->dtype != MRPHY_F64) return fail(MRPHY_ERR_ARG, "dtype must be MRPHY_F32 or MRPHY_F64%s");
  if (a->N < 1 || a->nM < 1 || a->nT < 1 || a->nC < 1) return fail(MRPHY_ERR_ARG, "N, nM, nT, nC must be >= 1%s");
  if (a->K < 1 || a->K > TCMAX) return fail(MRPHY_ERR_ARG, "checkpoint interval K must be in [1, 64]%s");
  p->sum_coils = a->b1 == nullptr;
  const int nc = p->sum_coils ? 1 : a->nC;
  if (nc > 8) return fail(MRPHY_ERR_ARG, "more than 8 transmit coils with a b1Map are not supported yet%s");
  p->NC = nc <= 1 ? 1 : nc <= 2 ? 2 : nc <= 4 ? 4 : 8;
  p->W = 2 * p->NC + 3;
  p->K = a->K;
  p->TCP = (a->K + 3) & ~3;
  p->nChunks = (a->nT + a->K - 1) / a->K;
  p->S = 1;
  p->tiles = (a->nM + BLK * p->S - 1) / (BLK * p->S);
  // grid: one tile per CTA while that stays within 8 resident waves, else a grid-stride loop
  int cap = (need_device ? sm_count_cached() : 148) * 8 * 8;
  if (const char* e = getenv("MRPHY_B200_MAX_CTAS")) {   // tests use it to force the multi-tile path
    const int v = atoi(e);
    if (v > 0) cap = v;
  }
  int P = p->tiles;
  if ((int64_t)P * a->N > cap) P = cap / a->N > 0 ? cap / a->N : 1;
  p->P = P;
  return MRPHY_OK;
}

}  // namespace

extern "C" size_t mrphy_fused_ckpt_elems(const mrphy_fused_args* a) {
  Plan p;
  if (make_plan(a, &p, false) != MRPHY_OK) return 0;
  const size_t n = (size_t)a->N * (size_t)(p.nChunks - 1) * 3 * (size_t)a->nM;
  return n ? n : 1;
}
extern "C" size_t mrphy_fused_wave_elems(const mrphy_fused_args* a) {
  Plan p;
  if (make_plan(a, &p, false) != MRPHY_OK) return 0;
  return (size_t)a->N * p.nChunks * p.W * p.TCP;
}
extern "C" size_t mrphy_fused_partial_elems(const mrphy_fused_args* a) {
  Plan p;
  if (make_plan(a, &p, true) != MRPHY_OK) return 0;
  return (size_t)a->N * p.P * p.W * (size_t)a->nT;
}

namespace {

template <typename T>
int check_common(const mrphy_fused_args* a, bool bwd) {
  if (!a->Mi && !bwd) return fail(MRPHY_ERR_ARG, "Mi is null%s");
  if (!a->rf || !a->gr || !a->loc) return fail(MRPHY_ERR_ARG, "rf, gr and loc are required%s");
  if (!a->gamma.ptr || !a->dt.ptr) return fail(MRPHY_ERR_ARG, "gamma and dt are required%s");
  if ((a->T1.ptr == nullptr) != (a->T2.ptr == nullptr)) return fail(MRPHY_ERR_ARG, "T1 and T2: both or neither (sims.py:68)%s");
  if (!a->Mo || !a->ckpt || !a->wave) return fail(MRPHY_ERR_ARG, "Mo, ckpt and wave buffers are required%s");
  if (bwd && (!a->gMo || !a->grf || !a->ggr || !a->partials)) return fail(MRPHY_ERR_ARG, "gMo, grf, ggr, partials are required%s");
  if (bwd && (a->flags & MRPHY_NEED_GMI) && !a->gMi) return fail(MRPHY_ERR_ARG, "gMi is null but MRPHY_NEED_GMI is set%s");
  return MRPHY_OK;
}

template <typename T>
KArgs<T> make_kargs(const mrphy_fused_args* a, const Plan& p) {
  KArgs<T> k;
  memset(&k, 0, sizeof(k));
  k.N = a->N; k.nM = a->nM; k.nT = a->nT; k.K = p.K; k.TCP = p.TCP; k.nChunks = p.nChunks; k.P = p.P;
  k.Mi = (const T*)a->Mi; k.Mi_sn = a->Mi_sn; k.Mi_sm = a->Mi_sm;
  k.loc = (const T*)a->loc; k.loc_sn = a->loc_sn; k.loc_sm = a->loc_sm;
  k.b1 = (const T*)a->b1; k.b1_sn = a->b1_sn; k.b1_sm = a->b1_sm; k.nC = a->nC;
  k.df = a->df; k.T1 = a->T1; k.T2 = a->T2; k.gamma = a->gamma; k.dt = a->dt;
  k.Mo = (T*)a->Mo; k.ckpt = (T*)a->ckpt; k.wave = (const T*)a->wave;
  k.gMo = (const T*)a->gMo; k.gMo_sn = a->gMo_sn; k.gMo_sm = a->gMo_sm;
  k.gMi = (T*)a->gMi; k.partials = (T*)a->partials;
  return k;
}

template <typename T>
int launch_pack(const mrphy_fused_args* a, const Plan& p, cudaStream_t st) {
  const int64_t total = (int64_t)a->N * p.nChunks * p.W * p.TCP;
  const int grid = (int)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
  const int64_t rf_sc = (a->flags & MRPHY_RF_COIL_DIM) ? a->rf_sc : 0;
  pack_waveform_kernel<T><<<grid, 256, 0, st>>>((const T*)a->rf, a->rf_sn, a->rf_sx, a->rf_st, rf_sc, (const T*)a->gr,
                                                a->gr_sn, a->gr_sx, a->gr_st, a->nC, p.NC, p.sum_coils, a->nT, p.K,
                                                p.TCP, p.nChunks, total, (T*)a->wave);
  ++g_launches;
  CK(cudaGetLastError());
  return MRPHY_OK;
}

template <typename T, int POL, bool RELAX, int NC>
int launch_fwd_s(const KArgs<T>& k, const Plan& p, cudaStream_t st) {
  dim3 grid(p.P, k.N);
  timing_begin(st);
  fused_fwd_kernel<T, POL, RELAX, NC, 1><<<grid, BLK, 0, st>>>(k);
  timing_end(st);
  ++g_launches;
  CK(cudaGetLastError());
  return MRPHY_OK;
}
template <typename T, int POL, bool RELAX, int NC>
int launch_bwd_s(const KArgs<T>& k, const Plan& p, int need_gmi, cudaStream_t st) {
  dim3 grid(p.P, k.N);
  constexpr size_t smem = BwdSmem<T, NC>::bytes;
  auto kern = fused_bwd_kernel<T, POL, RELAX, NC, 1>;
  if (smem > 48 * 1024) CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  timing_begin(st);
  kern<<<grid, BLK, smem, st>>>(k, need_gmi);
  timing_end(st);
  ++g_launches;
  CK(cudaGetLastError());
  return MRPHY_OK;
}

template <typename T, int POL, bool RELAX>
int dispatch_nc(bool bwd, const KArgs<T>& k, const Plan& p, int need_gmi, cudaStream_t st) {
  switch (p.NC) {
    case 1: return bwd ? launch_bwd_s<T, POL, RELAX, 1>(k, p, need_gmi, st) : launch_fwd_s<T, POL, RELAX, 1>(k, p, st);
    case 2: return bwd ? launch_bwd_s<T, POL, RELAX, 2>(k, p, need_gmi, st) : launch_fwd_s<T, POL, RELAX, 2>(k, p, st);
    case 4: return bwd ? launch_bwd_s<T, POL, RELAX, 4>(k, p, need_gmi, st) : launch_fwd_s<T, POL, RELAX, 4>(k, p, st);
    case 8: return bwd ? launch_bwd_s<T, POL, RELAX, 8>(k, p, need_gmi, st) : launch_fwd_s<T, POL, RELAX, 8>(k, p, st);
  }
  return fail(MRPHY_ERR_ARG, "internal: bad NC%s");
}

template <typename T>
int dispatch(bool bwd, const mrphy_fused_args* a, const Plan& p, cudaStream_t st) {
  const KArgs<T> k = make_kargs<T>(a, p);
  const bool relax = a->T1.ptr != nullptr;
  const bool precise = (a->flags & MRPHY_TRIG_PRECISE) != 0 && sizeof(T) == 4;
  const int need_gmi = (a->flags & MRPHY_NEED_GMI) ? 1 : 0;
  if (precise) {
    return relax ? dispatch_nc<T, TRIG_PRECISE, true>(bwd, k, p, need_gmi, st)
                 : dispatch_nc<T, TRIG_PRECISE, false>(bwd, k, p, need_gmi, st);
  }
  return relax ? dispatch_nc<T, TRIG_FAST, true>(bwd, k, p, need_gmi, st)
               : dispatch_nc<T, TRIG_FAST, false>(bwd, k, p, need_gmi, st);
}

template <typename T>
int run_fwd(const mrphy_fused_args* a, cudaStream_t st) {
  Plan p;
  int rc = make_plan(a, &p, true);
  if (rc) return rc;
  if ((rc = check_common<T>(a, false))) return rc;
  if ((rc = launch_pack<T>(a, p, st))) return rc;
  return dispatch<T>(false, a, p, st);
}

template <typename T>
int run_bwd(const mrphy_fused_args* a, int wave_is_packed, cudaStream_t st) {
  Plan p;
  int rc = make_plan(a, &p, true);
  if (rc) return rc;
  if ((rc = check_common<T>(a, true))) return rc;
  if (!wave_is_packed && (rc = launch_pack<T>(a, p, st))) return rc;
  if ((rc = dispatch<T>(true, a, p, st))) return rc;
  dim3 grid((a->nT + 31) / 32, p.W, a->N), block(32, 8);
  grad_finalize_kernel<T><<<grid, block, 0, st>>>((const T*)a->partials, p.P, p.W, p.NC, a->nC, a->nT,
                                                  (a->flags & MRPHY_RF_COIL_DIM) ? 1 : 0, p.sum_coils, (T*)a->grf,
                                                  (T*)a->ggr);
  ++g_launches;
  CK(cudaGetLastError());
  return MRPHY_OK;
}

}  // namespace

extern "C" int mrphy_blochsim_fused_fwd(const mrphy_fused_args* a, void* cuda_stream) {
  g_launches = 0;
  g_err[0] = 0;
  if (!a) return fail(MRPHY_ERR_ARG, "null args%s");
  cudaStream_t st = (cudaStream_t)cuda_stream;
  return a->dtype == MRPHY_F64 ? run_fwd<double>(a, st) : run_fwd<float>(a, st);
}

extern "C" int mrphy_blochsim_fused_bwd(const mrphy_fused_args* a, int wave_is_packed, void* cuda_stream) {
  g_launches = 0;
  g_err[0] = 0;
  if (!a) return fail(MRPHY_ERR_ARG, "null args%s");
  cudaStream_t st = (cudaStream_t)cuda_stream;
  return a->dtype == MRPHY_F64 ? run_bwd<double>(a, wave_is_packed, st) : run_bwd<float>(a, wave_is_packed, st);
}

// tcgen05 / TMEM wrappers (sm_100a) for the one contraction on this path: the multi-coil transmit field
//   Bx + i By [spin][step] = sum_c b1[spin][c] * rf[c][step]
// as a TF32 tensor-core product with fp32 accuracy (operands split into exactly representable TF32 parts).
//
// Shared-memory operand tiles use the canonical K-major layout without swizzle: 16-byte K-chunks (4 tf32), a core matrix is
// 8 rows x 16 B = 128 contiguous bytes; tile[chunk][row][4]: the next 8 rows are SBO = 128 B further, the next K-chunk
// LBO = rows * 16 B further.  One tcgen05.mma (kind::tf32) consumes K = 8 = two chunks.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "ptx_helpers.cuh"

namespace mrphy {
namespace tc {

// matrix descriptor of a K-major, unswizzled tile whose rows are 16 B apart (see above); `rows` = rows of the whole tile
__device__ __forceinline__ uint64_t kmajor_desc(const void* smem_tile, int rows) {
  const uint32_t addr = smem_u32(smem_tile);
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3fff);                    // start address            bits [0,14)
  d |= (uint64_t)((((uint32_t)rows * 16u) >> 4) & 0x3fff) << 16;   // leading byte offset (next K-chunk)  bits [16,30)
  d |= (uint64_t)((128u >> 4) & 0x3fff) << 32;              // stride byte offset (next 8 rows)    bits [32,46)
  d |= (uint64_t)1 << 46;                                   // descriptor version (Blackwell)      bits [46,48)
  return d;                                                 // base offset 0, layout type 0 = no swizzle
}
// instruction descriptor: D fp32, A and B tf32; a_mn / b_mn: operand is MN-major (TF32: only in the 128B_BASE32B swizzled
// layout -- an unswizzled MN-major descriptor is accepted and yields zeros, profiles/ubench/tc_field.cu)
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N, bool a_mn = false, bool b_mn = false) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]^T, issued by ONE thread
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}
// arrive on `bar` when every tcgen05.mma issued so far by this thread has completed (implies fence::before_thread_sync)
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// TMEM allocation by ONE whole warp; COLS a power of two >= 32.  The base address lands in *slot (shared memory).
template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* slot) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "n"(COLS) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_free(uint32_t base) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "n"(COLS) : "memory");
}

// 16 consecutive columns of this thread's TMEM lane (warp w of the CTA owns lanes 32 (w % 4) ...): taddr = lane << 16 | column.
// tmem_ld16_issue starts the read, tmem_ld_wait makes the registers of EVERY read issued so far valid; the empty asm ties the
// values to the wait so that no use can be scheduled above it.
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, float (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]), "=f"(v[8]), "=f"(v[9]),
        "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait(float (&v)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  asm volatile("" : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]), "+f"(v[4]), "+f"(v[5]), "+f"(v[6]), "+f"(v[7]), "+f"(v[8]),
               "+f"(v[9]), "+f"(v[10]), "+f"(v[11]), "+f"(v[12]), "+f"(v[13]), "+f"(v[14]), "+f"(v[15]));
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  tmem_ld16_issue(taddr, v);
  tmem_ld_wait(v);
}

// D[tmem] (+)= A[tmem] * B[smem]^T: A read from TMEM -- row i of A = lane i, its K values in consecutive 32-bit columns
__device__ __forceinline__ void mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}
// 8 consecutive columns of this thread's TMEM lane <- registers (warp-collective); tmem_st_wait before anything reads them
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "f"(v[0]), "f"(v[1]),
               "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// x = hi + mid + lo exactly, every part representable in TF32 (11-bit significand): products of parts are exact in the
// tensor core, so the six products hi*hi, hi*mid, mid*hi, mid*mid, hi*lo, lo*hi reproduce x*y to 2^-33.
__device__ __forceinline__ float tf32_rn(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
__device__ __forceinline__ void split3(float x, float& hi, float& mid, float& lo) {
  hi = tf32_rn(x);
  const float r = x - hi;      // exact
  mid = tf32_rn(r);
  lo = r - mid;                // exact, <= 2 significant bits left
}

}  // namespace tc
}  // namespace mrphy

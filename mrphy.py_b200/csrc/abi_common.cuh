// Error/launch-count plumbing shared by the translation units of libmrphy_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdio.h>

#include "../../include/mrphy_b200.h"

namespace mrphy {
char* err_buf();           // thread-local, 512 bytes
int& launch_count();       // thread-local
// bench-only event bracket around the main kernel of an entry point (see mrphy_kernel_timing)
void timing_begin(cudaStream_t st);
void timing_end(cudaStream_t st);
inline int fail(int code, const char* fmt, const char* detail = "") {
  snprintf(err_buf(), 512, fmt, detail);
  return code;
}
__device__ __forceinline__ double ld_param(const mrphy_param& p, int n, int i) {
  const int64_t o = (int64_t)n * p.sn + (int64_t)i * p.sm;
  return p.f64 ? reinterpret_cast<const double*>(p.ptr)[o] : (double)reinterpret_cast<const float*>(p.ptr)[o];
}
}  // namespace mrphy

#define CK(call)                                                                                      \
  do {                                                                                                \
    cudaError_t e_ = (call);                                                                          \
    if (e_ != cudaSuccess) return mrphy::fail(MRPHY_ERR_CUDA, #call ": %s", cudaGetErrorString(e_)); \
  } while (0)

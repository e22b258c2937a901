// Does FFMA2 throughput depend on how many distinct register-pair operands it reads?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_operands ffma2_operands.cu
#include <cstdio>
#include <cuda_runtime.h>
constexpr int ITER = 2048, U = 12;

template <int MODE>
__global__ void __launch_bounds__(256) k(float2* out, float a, float b) {
  float2 y[U];
  float x[U];
#pragma unroll
  for (int i = 0; i < U; ++i) { y[i] = make_float2(threadIdx.x * 1e-3f + i, 0.5f * i + a); x[i] = y[i].x * 1.1f; }
  const float2 a2 = make_float2(a, a * 1.0001f), b2 = make_float2(b, b * 0.999f);
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < U; ++i) {
      const int j = (i + 4) % U, l = (i + 8) % U;
      if (MODE == 0) y[i] = __ffma2_rn(y[i], a2, b2);                               // 1 varying pair + 2 loop-invariant pairs
      if (MODE == 1) y[i] = __ffma2_rn(y[i], y[j], b2);                             // 2 varying pairs
      if (MODE == 2) y[i] = __ffma2_rn(y[j], y[l], y[i]);                           // 3 varying pairs
      if (MODE == 3) y[i] = __ffma2_rn(y[j], make_float2(x[l], x[l]), y[i]);        // 2 pairs + 1 broadcast scalar
      if (MODE == 4) y[i] = __fmul2_rn(y[j], y[l]);                                 // FMUL2, 2 varying pairs
      if (MODE == 5) x[i] = fmaf(x[j], x[l], x[i]);                                 // scalar FFMA, 3 varying
      if (MODE == 6) { y[i].x = fmaf(y[j].x, y[l].x, y[i].x); y[i].y = fmaf(y[j].y, y[l].y, y[i].y); }   // 2 scalar FFMAs on pairs
      if (MODE == 8) x[i] = fmaf(x[(i + 1) % U], x[(i + 2) % U], x[i]);                      // scalar FFMA, 3 varying, registers of both parities
      if (MODE == 9) x[i] = fmaf(x[(i + 1) % U], x[(i + 3) % U], x[i]);                      // scalar FFMA, 3 varying, another mix
      if (MODE == 7) y[i] = __ffma2_rn(y[j], y[l], make_float2(0.5f, 0.5f));        // 2 pairs + immediate
    }
  }
  float2 s = make_float2(0, 0);
#pragma unroll
  for (int i = 0; i < U; ++i) { s.x += y[i].x + x[i]; s.y += y[i].y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, double instr_per_iter, float2* out, int sms) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int grid = sms * 8;
  k<MODE><<<grid, 256>>>(out, 1.0001f, 0.5f);
  cudaEventRecord(e0);
  k<MODE><<<grid, 256>>>(out, 1.0001f, 0.5f);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double wi = (double)grid * 8 * ITER * U * instr_per_iter;
  printf("%-46s %7.3f ms  %6.3f cyc/warp-instr/SMSP (1.92 GHz)\n", name, ms, (ms * 1e-3 * 1.92e9) / (wi / sms / 4));
}

int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float2* out; cudaMalloc(&out, sizeof(float2) * sms * 8 * 256);
  run<0>("FFMA2 1 varying pair + 2 invariant", 1, out, sms);
  run<1>("FFMA2 2 varying pairs + 1 invariant", 1, out, sms);
  run<2>("FFMA2 3 varying pairs", 1, out, sms);
  run<3>("FFMA2 2 pairs + broadcast scalar", 1, out, sms);
  run<4>("FMUL2 2 varying pairs", 1, out, sms);
  run<5>("FFMA scalar 3 varying", 1, out, sms);
  run<6>("2x FFMA scalar on pair halves (2 instr)", 2, out, sms);
  run<7>("FFMA2 2 pairs + immediate", 1, out, sms);
  run<8>("FFMA scalar 3 varying (i,i+1,i+2)", 1, out, sms);
  run<9>("FFMA scalar 3 varying (i,i+1,i+3)", 1, out, sms);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}

// Micro-benchmark of the issue/pipe rates that bound the Bloch kernels on B200:
// FFMA vs FFMA2 (packed f32x2), MUFU (sin/rsq), FRND/F2I, broadcast LDS.  Prints instr/clk/SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu && ./pipes
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITER = 4096, UNR = 16;

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, float a, float b) {
  float x[UNR];
  float2 y[UNR];
#pragma unroll
  for (int i = 0; i < UNR; ++i) { x[i] = threadIdx.x * 1e-3f + i; y[i] = make_float2(x[i], x[i] + 1.f); }
  const float2 a2 = make_float2(a, a * 1.0001f), b2 = make_float2(b, b * 0.999f);
  __shared__ float4 sm[64];
  if (threadIdx.x < 64) sm[threadIdx.x] = make_float4(a, b, a, b);
  __syncthreads();
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < UNR; ++i) {
      if (MODE == 0) x[i] = fmaf(x[i], a, b);                                   // FFMA
      if (MODE == 1) y[i] = __ffma2_rn(y[i], a2, b2);                           // FFMA2
      if (MODE == 2) x[i] = __sinf(x[i]);                                       // FMUL + MUFU.SIN
      if (MODE == 3) x[i] = rsqrtf(x[i]);                                       // MUFU.RSQ
      if (MODE == 4) x[i] = rintf(x[i] * a);                                    // FMUL + FRND
      if (MODE == 5) x[i] = (float)__float2int_rn(x[i]) * a;                    // F2I + I2F + FMUL
      if (MODE == 6) { float4 v = sm[(it + i) & 63]; x[i] = fmaf(x[i], v.x, v.y); }   // broadcast LDS.128 + FFMA
      if (MODE == 7) { x[i] = fmaf(x[i], a, b); if (i % 4 == 0) x[i] = __sinf(x[i]); }   // 4 FFMA : 1 MUFU mix
      if (MODE == 8) { y[i] = __ffma2_rn(y[i], a2, b2); if (i % 4 == 0) y[i].x = __sinf(y[i].x); }   // 4 FFMA2 : 1 MUFU
      if (MODE == 9) { y[i] = __ffma2_rn(y[i], a2, b2); x[i] = fmaf(x[i], a, b); }   // FFMA2 + FFMA interleaved
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < UNR; ++i) s += x[i] + y[i].x + y[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, double ops_per_iter, float* out, int sms, double ghz_hint) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int grid = sms * 8;
  k<MODE><<<grid, 256>>>(out, 1.0001f, 0.5f);
  cudaEventRecord(e0);
  k<MODE><<<grid, 256>>>(out, 1.0001f, 0.5f);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double warp_instr = (double)grid * 8 * ITER * UNR * ops_per_iter;
  printf("%-34s %8.3f ms  %7.3f warp-instr/clk/SM (at %.3f GHz)\n", name, ms, warp_instr / (ms * 1e-3) / sms / (ghz_hint * 1e9),
         ghz_hint);
}

int main() {
  int sms, khz;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double ghz = khz * 1e-6;
  float* out;
  cudaMalloc(&out, sizeof(float) * sms * 8 * 256);
  printf("SMs %d, clock %.3f GHz\n", sms, ghz);
  run<0>("FFMA", 1, out, sms, ghz);
  run<1>("FFMA2 (counted as 1 instr)", 1, out, sms, ghz);
  run<2>("FMUL.RZ + MUFU.SIN (2 instr)", 2, out, sms, ghz);
  run<3>("MUFU.RSQ", 1, out, sms, ghz);
  run<4>("FMUL + FRND (2 instr)", 2, out, sms, ghz);
  run<5>("F2I + I2F + FMUL (3 instr)", 3, out, sms, ghz);
  run<6>("LDS.128 bcast + FFMA (2 instr)", 2, out, sms, ghz);
  run<7>("4 FFMA : 1 (FMUL+MUFU) (1.5/iter)", 1.5, out, sms, ghz);
  run<8>("4 FFMA2 : 1 (FMUL+MUFU) (1.5/iter)", 1.5, out, sms, ghz);
  run<9>("FFMA2 + FFMA (2 instr)", 2, out, sms, ghz);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}

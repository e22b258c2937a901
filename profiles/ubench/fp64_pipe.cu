// FP64 pipe rate of B200 (sm_100a): DFMA / DMUL / DADD warp-instructions per clock per SM, the denominator of the fp64
// kernels' roofline.  Same shape as pipes.cu.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_pipe fp64_pipe.cu && ./fp64_pipe
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITER = 2048, UNR = 16;

template <int MODE>
__global__ void __launch_bounds__(256) k(double* out, double a, double b) {
  double x[UNR];
#pragma unroll
  for (int i = 0; i < UNR; ++i) x[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < UNR; ++i) {
      if (MODE == 0) x[i] = fma(x[i], a, b);                     // DFMA, 1 varying + 2 invariant operands
      if (MODE == 1) x[i] = x[i] * a;                            // DMUL
      if (MODE == 2) x[i] = x[i] + b;                            // DADD
      if (MODE == 3) x[i] = fma(x[i], x[(i + 1) % UNR], x[(i + 2) % UNR]);   // DFMA, three distinct register pairs
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < UNR; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, double* out, int sms, double ghz) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int grid = sms * 8;
  k<MODE><<<grid, 256>>>(out, 1.0000001, 0.5);
  cudaEventRecord(e0);
  k<MODE><<<grid, 256>>>(out, 1.0000001, 0.5);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double warp_instr = (double)grid * 8 * ITER * UNR;
  printf("%-40s %8.3f ms  %7.4f warp-instr/clk/SM (at %.3f GHz)  = %6.2f TFLOP/s (2 flop, 32 lanes)\n", name, ms,
         warp_instr / (ms * 1e-3) / sms / (ghz * 1e9), ghz, warp_instr * 64 / (ms * 1e-3) / 1e12);
}

int main() {
  int sms, khz;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double ghz = khz * 1e-6;
  double* out;
  cudaMalloc(&out, sizeof(double) * sms * 8 * 256);
  printf("SMs %d, clock %.3f GHz\n", sms, ghz);
  run<0>("DFMA (1 varying operand)", out, sms, ghz);
  run<1>("DMUL", out, sms, ghz);
  run<2>("DADD", out, sms, ghz);
  run<3>("DFMA (three distinct register pairs)", out, sms, ghz);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}

// Accuracy of the fp32 sincos options on B200: raw MUFU, MUFU + first-order correction of the range-reduction
// rounding, and the polynomial path of csrc/bloch_math.cuh.  Max/RMS abs error vs double over several ranges.
#include <cstdio>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include "../../mrphy.py_b200/csrc/bloch_math.cuh"
using namespace mrphy;

__device__ __forceinline__ void sc_corrected(float x, float& s, float& c) {
  float s0, c0;
  __sincosf(x, &s0, &c0);                                   // FMUL.RZ by 1/2pi, MUFU.SIN, MUFU.COS
  const float r = __fmul_rz(x, 0.15915493667125701904f);    // the very argument MUFU saw (in revolutions)
  float d = fmaf(-r, 6.2831854820251465f, x);               // x - 2*pi*r, exactly (two-part 2*pi)
  d = fmaf(r, 1.7484555e-7f, d);
  s = fmaf(d, c0, s0);
  c = fmaf(-d, s0, c0);
}

__global__ void k(const float* x, int n, float* out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s, c;
  __sincosf(x[i], &s, &c); out[i] = s; out[n + i] = c;
  sc_corrected(x[i], s, c); out[2 * n + i] = s; out[3 * n + i] = c;
  Fn<float, TRIG_PRECISE>::sc(x[i], s, c); out[4 * n + i] = s; out[5 * n + i] = c;
  out[6 * n + i] = rsqrtf(x[i] + 1e-3f); out[7 * n + i] = Fn<float, TRIG_PRECISE>::rsq(x[i] + 1e-3f);
}

int main() {
  const int n = 1 << 22;
  const double ranges[4][2] = {{0, 0.5}, {0, 3.2}, {0, 30}, {0, 1000}};
  float *dx, *dout; cudaMalloc(&dx, n * 4); cudaMalloc(&dout, 8 * n * 4);
  std::vector<float> x(n), out(8 * (size_t)n);
  for (auto& rg : ranges) {
    for (int i = 0; i < n; ++i) x[i] = (float)(rg[0] + (rg[1] - rg[0]) * (i + 0.37) / n);
    cudaMemcpy(dx, x.data(), n * 4, cudaMemcpyHostToDevice);
    k<<<n / 256, 256>>>(dx, n, dout);
    cudaMemcpy(out.data(), dout, 8 * (size_t)n * 4, cudaMemcpyDeviceToHost);
    double mx[4] = {0, 0, 0, 0}, ss[4] = {0, 0, 0, 0};
    for (int i = 0; i < n; ++i) {
      const double xs = x[i], sd = sin(xs), cd = cos(xs);
      for (int v = 0; v < 3; ++v) {
        const double e = fmax(fabs(out[(2 * v) * (size_t)n + i] - sd), fabs(out[(2 * v + 1) * (size_t)n + i] - cd));
        mx[v] = fmax(mx[v], e); ss[v] += e * e;
      }
    }
    printf("x in [%g,%g]: max|err| (rms)  mufu %.2e (%.1e)   mufu+corr %.2e (%.1e)   poly %.2e (%.1e)\n", rg[0], rg[1], mx[0],
           sqrt(ss[0] / n), mx[1], sqrt(ss[1] / n), mx[2], sqrt(ss[2] / n));
  }
  double m0 = 0, m1 = 0;
  for (int i = 0; i < n; ++i) { const double t = 1.0 / sqrt((double)(x[i] + 1e-3f)); m0 = fmax(m0, fabs(out[6 * (size_t)n + i] / t - 1)); m1 = fmax(m1, fabs(out[7 * (size_t)n + i] / t - 1)); }
  printf("rsqrt max rel err: MUFU.RSQ %.2e   + Newton %.2e\n%s\n", m0, m1, cudaGetErrorString(cudaGetLastError()));
}

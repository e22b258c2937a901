// Is the error of MUFU.SIN/COS a BIAS (accumulates linearly over the time steps of a Bloch simulation) or noise (random
// walk)?  Prints mean / rms / max of the ANGLE error eps = S*cos(x) - C*sin(x) of the (S, C) pair each variant returns,
// and of the radius error S^2 + C^2 - 1, over x in [0, 6.3] (the per-step rotation angles of the bench workloads).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -ftz=true -o trig_bias trig_bias.cu
#include <cstdio>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include "../../mrphy.py_b200/csrc/bloch_math.cuh"
using namespace mrphy;

#define TWO_PI_HI 6.2831854820251465f
#define TWO_PI_LO (-1.7484555e-7f)
#define INV_2PI 0.15915493667125701904f

// first-order correction with the product MUFU's range reduction sees (FMUL.RZ)
__device__ __forceinline__ void sc_corr_rz(float x, float& s, float& c) {
  float s0, c0;
  __sincosf(x, &s0, &c0);
  const float r = __fmul_rz(x, INV_2PI);
  float d = fmaf(-r, TWO_PI_HI, x);
  d = fmaf(-r, TWO_PI_LO, d);
  s = fmaf(d, c0, s0);
  c = fmaf(-d, s0, c0);
}
__global__ void k(const float* x, int n, float* out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s, c;
  __sincosf(x[i], &s, &c); out[i] = s; out[n + i] = c;
  sc_corr_rz(x[i], s, c); out[2 * n + i] = s; out[3 * n + i] = c;
  Fn<float, TRIG_PRECISE>::sc(x[i], s, c); out[4 * n + i] = s; out[5 * n + i] = c;
}

int main() {
  const int n = 1 << 22, NV = 3;
  const char* names[NV] = {"mufu raw", "mufu + corr (rz product)", "polynomial (shipped)"};
  float *dx, *dout; cudaMalloc(&dx, n * 4); cudaMalloc(&dout, 2 * NV * (size_t)n * 4);
  std::vector<float> x(n), out(2 * NV * (size_t)n);
  const double ranges[3][2] = {{0, 0.8}, {0, 3.2}, {0, 6.3}};
  for (auto& rg : ranges) {
    for (int i = 0; i < n; ++i) x[i] = (float)(rg[0] + (rg[1] - rg[0]) * (i + 0.37) / n);
    cudaMemcpy(dx, x.data(), n * 4, cudaMemcpyHostToDevice);
    k<<<n / 256, 256>>>(dx, n, dout);
    cudaMemcpy(out.data(), dout, 2 * NV * (size_t)n * 4, cudaMemcpyDeviceToHost);
    printf("x in [%g, %g]\n", rg[0], rg[1]);
    for (int v = 0; v < NV; ++v) {
      double m = 0, ss = 0, mx = 0, mr = 0, sr = 0;
      for (int i = 0; i < n; ++i) {
        const double xs = x[i], S = out[(2 * v) * (size_t)n + i], C = out[(2 * v + 1) * (size_t)n + i];
        const double e = S * cos(xs) - C * sin(xs), r = S * S + C * C - 1;
        m += e; ss += e * e; mx = fmax(mx, fabs(e)); mr += r; sr += r * r;
      }
      printf("  %-32s angle err: mean %+.2e  rms %.2e  max %.2e   radius err: mean %+.2e rms %.2e\n", names[v], m / n, sqrt(ss / n), mx,
             mr / n, sqrt(sr / n));
    }
  }
  // per binade of t = x/2pi: mean angle error of the raw and of the rz-corrected pair, mean radius error
  {
    for (int i = 0; i < n; ++i) x[i] = (float)(0.02 + (12.6 - 0.02) * (i + 0.37) / n);
    cudaMemcpy(dx, x.data(), n * 4, cudaMemcpyHostToDevice);
    k<<<n / 256, 256>>>(dx, n, dout);
    cudaMemcpy(out.data(), dout, 2 * NV * (size_t)n * 4, cudaMemcpyDeviceToHost);
    double m[2][12] = {{0}}, r2[12] = {0}, ssq[12] = {0}; long cnt[12] = {0};
    for (int i = 0; i < n; ++i) {
      const double xs = x[i], t = xs / 6.283185307179586;
      int b = (int)floor(log2(t)) + 9; if (b < 0 || b > 11) continue;
      for (int v = 0; v < 2; ++v) {
        const double S = out[(2 * v) * (size_t)n + i], C = out[(2 * v + 1) * (size_t)n + i];
        m[v][b] += S * cos(xs) - C * sin(xs);
        if (v == 1) { r2[b] += S * S + C * C - 1; const double e = S * cos(xs) - C * sin(xs); ssq[b] += e * e; }
      }
      ++cnt[b];
    }
    for (int b = 0; b < 12; ++b) if (cnt[b])
      printf("t in [2^%d, 2^%d) rev (x from %.3f rad): raw mean %+.2e   rz-corrected mean %+.2e rms %.2e   radius mean %+.2e   (n=%ld)\n", b - 9, b - 8,
             6.283185307179586 * pow(2.0, b - 9), m[0][b] / cnt[b], m[1][b] / cnt[b], sqrt(ssq[b] / cnt[b]), r2[b] / cnt[b], cnt[b]);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}

// Error statistics (mean = bias, rms, max of the relative error against exp2() in double) of three
// float32 exp2 evaluations over the argument ranges the P3D cell evaluation uses:
//   raw   ex2.approx.ftz.f32 (MUFU.EX2)
//   rr    range-reduced: 2^round(x) * MUFU.EX2(x - round(x))
//   poly  range-reduced degree-7 polynomial on the FMA pipe (what k_p3d_hankel uses)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mufu_error tools/mufu_error.cu && /tmp/mufu_error
#include <cstdio>
#include <cmath>
#include <vector>
__device__ __forceinline__ float ex2_raw(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float ex2_rr(float x) {
  x = fminf(fmaxf(x, -125.0f), 126.0f);
  const float t = x + 12582912.0f;
  const float f = x - (t - 12582912.0f);
  return __int_as_float(__float_as_int(ex2_raw(f)) + (__float_as_int(t) << 23));
}
__device__ __forceinline__ float ex2_poly(float x) {
  x = fminf(fmaxf(x, -125.0f), 126.0f);
  const float t = x + 12582912.0f;
  const float f = x - (t - 12582912.0f);
  float p = 1.529732435301412e-05f;
  p = fmaf(p, f, 0.00015461444854736328f);
  p = fmaf(p, f, 0.0013333501992747188f);
  p = fmaf(p, f, 0.009618056938052177f);
  p = fmaf(p, f, 0.05550410971045494f);
  p = fmaf(p, f, 0.24022650718688965f);
  p = fmaf(p, f, 0.6931471824645996f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}
__global__ void k(const float* x, float* y, int n, int which) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = which == 0 ? ex2_raw(x[i]) : (which == 1 ? ex2_rr(x[i]) : ex2_poly(x[i]));
}
int main() {
  const int n = 1 << 22;
  std::vector<float> hx(n), hy(n);
  float *dx, *dy;
  cudaMalloc(&dx, n * 4); cudaMalloc(&dy, n * 4);
  const double lo[] = {-0.5, 0, -1, -8, -30, 0, -0.25}, hi[] = {0, 0.5, 0, 0, -8, 4, 0.25};
  const char* names[] = {"raw ", "rr  ", "poly"};
  for (int r = 0; r < 7; ++r) {
    for (int i = 0; i < n; ++i) hx[i] = (float)(lo[r] + (hi[r] - lo[r]) * ((i * 2654435761u) % 1000003) / 1000003.0);
    cudaMemcpy(dx, hx.data(), n * 4, cudaMemcpyHostToDevice);
    for (int w = 0; w < 3; ++w) {
      k<<<(n + 255) / 256, 256>>>(dx, dy, n, w);
      cudaMemcpy(hy.data(), dy, n * 4, cudaMemcpyDeviceToHost);
      double s = 0, s2 = 0, mx = 0;
      for (int i = 0; i < n; ++i) {
        double t = exp2((double)hx[i]);
        double e = (hy[i] - t) / t;
        s += e; s2 += e * e; mx = fmax(mx, fabs(e));
      }
      printf("%s x in [%5g,%4g]: mean %+.3e  rms %.3e  max %.3e\n", names[w], lo[r], hi[r], s / n, sqrt(s2 / n), mx);
    }
  }
  return 0;
}

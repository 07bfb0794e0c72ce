// Mean relative error of ex2.approx.ftz.f32 in bins of the argument (is the bias a smooth function of x?)
#include <cstdio>
#include <cmath>
#include <vector>
__global__ void k(const float* x, float* y, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y[i]) : "f"(x[i]));
}
int main() {
  const int n = 1 << 21;
  std::vector<float> hx(n), hy(n);
  float *dx, *dy;
  cudaMalloc(&dx, n * 4); cudaMalloc(&dy, n * 4);
  for (double lo = -2.0; lo < 2.0; lo += 0.125) {
    const double hi = lo + 0.125;
    for (int i = 0; i < n; ++i) hx[i] = (float)(lo + (hi - lo) * ((i * 2654435761u) % 1000003) / 1000003.0);
    cudaMemcpy(dx, hx.data(), n * 4, cudaMemcpyHostToDevice);
    k<<<(n + 255) / 256, 256>>>(dx, dy, n);
    cudaMemcpy(hy.data(), dy, n * 4, cudaMemcpyDeviceToHost);
    double s = 0, s2 = 0;
    for (int i = 0; i < n; ++i) { double t = exp2((double)hx[i]); double e = (hy[i] - t) / t; s += e; s2 += e * e; }
    printf("x in [%+.3f,%+.3f): mean %+.3e rms %.3e\n", lo, hi, s / n, sqrt(s2 / n));
  }
  return 0;
}

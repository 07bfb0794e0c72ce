// Same question for plain DFMA: does FP32 work on the same sub-partition slow down DFMA?
// Warps 0..3 run 16 independent DFMA chains; warps 4..15 run FFMA / FFMA2 / nothing.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* cyc, int iters, int mode, int other_iters, double a, double b, int hi_first) {
  const int warp = threadIdx.x >> 5;
  if (warp >= hi_first * 12 && warp < hi_first * 12 + 4) {
    double acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x + i;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = fma(acc[i], a, b);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i];
    if (s == 1.2345) out[0] = s;
    if (blockIdx.x == 0 && (threadIdx.x & 31) == 0 && (warp & 3) == 0) cyc[0] = t1 - t0;
  } else if (mode == 1) {
    float x0 = threadIdx.x, x1 = 1, x2 = 2, x3 = 3, x4 = 4, x5 = 5, x6 = 6, x7 = 7;
    const float y = (float)a, z = (float)b;
    for (int it = 0; it < other_iters; ++it) {
      x0 = fmaf(x0, y, z); x1 = fmaf(x1, y, z); x2 = fmaf(x2, y, z); x3 = fmaf(x3, y, z);
      x4 = fmaf(x4, y, z); x5 = fmaf(x5, y, z); x6 = fmaf(x6, y, z); x7 = fmaf(x7, y, z);
    }
    if (x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7 == 1.2345f) out[1] = x0;
  } else if (mode == 2) {
    float2 x0 = make_float2(threadIdx.x, 1), x1 = make_float2(1, 2), x2 = make_float2(2, 3), x3 = make_float2(3, 4);
    float2 x4 = make_float2(4, 1), x5 = make_float2(5, 2), x6 = make_float2(6, 3), x7 = make_float2(7, 4);
    const float2 y = make_float2((float)a, (float)a), z = make_float2((float)b, (float)b);
    for (int it = 0; it < other_iters; ++it) {
      x0 = __ffma2_rn(x0, y, z); x1 = __ffma2_rn(x1, y, z); x2 = __ffma2_rn(x2, y, z); x3 = __ffma2_rn(x3, y, z);
      x4 = __ffma2_rn(x4, y, z); x5 = __ffma2_rn(x5, y, z); x6 = __ffma2_rn(x6, y, z); x7 = __ffma2_rn(x7, y, z);
    }
    if (x0.x + x1.x + x2.x + x3.x + x4.y + x5.y + x6.y + x7.y == 1.2345f) out[1] = x0.x;
  }
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  double* d; long long* c; cudaMalloc(&d, 64); cudaMalloc(&c, 8);
  const char* names[] = {"idle", "FFMA", "FFMA2"};
  const int iters = 20000;
  for (int hi_first = 0; hi_first < 2; ++hi_first)
  for (int mode = 0; mode < 3; ++mode) {
    k<<<p.multiProcessorCount, 512>>>(d, c, iters, mode, iters * 16 * 2, 1.0000001, 0.5, hi_first);
    cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    printf("FP64 warps = %s  other warps: %-6s  %.2f cycles per DFMA on the DFMA-issuing warp (1 warp per sub-partition)\n", hi_first ? "12..15" : "0..3  ", names[mode], (double)h / (iters * 16.0));
  }
  return 0;
}

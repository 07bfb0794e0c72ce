// Can the FP64 tensor pipe (DMMA) and the FP32 FMA pipe (FFMA2) run concurrently on one SM
// sub-partition when they come from DIFFERENT warps?  Measures both rates alone and together.
//   warps 0..3  : 12 independent DMMA chains each (one warp per sub-partition)
//   warps 4..15 : packed FFMA2, `ch` independent chains each (three warps per sub-partition)
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__global__ void k(double* out, long long* cyc, int it_d, int it_f, double a, double b) {
  const int warp = threadIdx.x >> 5;
  __shared__ volatile int go;
  if (threadIdx.x == 0) go = 1;
  __syncthreads();
  if (warp < 4) {
    if (it_d == 0) return;
    double acc[12][2];
#pragma unroll
    for (int i = 0; i < 12; ++i) acc[i][0] = acc[i][1] = threadIdx.x + i;
    long long t0 = clock64();
    for (int it = 0; it < it_d; ++it) {
#pragma unroll
      for (int i = 0; i < 12; ++i) dmma(acc[i][0], acc[i][1], a, b);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < 12; ++i) s += acc[i][0] + acc[i][1];
    if (s == 1.2345) out[0] = s;
    if (blockIdx.x == 0 && threadIdx.x == 0) cyc[0] = t1 - t0;
  } else {
    if (it_f == 0) return;
    float2 x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = make_float2(threadIdx.x + i, i);
    const float2 y = make_float2((float)a, (float)a), z = make_float2((float)b, (float)b);
    long long t0 = clock64();
    for (int it = 0; it < it_f; ++it) {
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = __ffma2_rn(x[i], y, z);
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i].x + x[i].y;
    if (s == 1.2345f) out[1] = s;
    if (blockIdx.x == 0 && threadIdx.x == 128) cyc[1] = t1 - t0;
  }
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  double* d; long long* c; cudaMalloc(&d, 64); cudaMalloc(&c, 16);
  const int ND = 20000, NF = 20000 * 12 * 16 / (8 * 2 * 3) * 1;   // comparable durations
  struct { int d, f; const char* name; } cases[] = {{ND, 0, "DMMA alone"}, {0, NF, "FFMA2 alone"}, {ND, NF, "both"}, {ND, NF * 2, "both (FFMA2 runs longer)"}};
  for (auto& cs : cases) {
    cudaMemset(c, 0, 16);
    k<<<p.multiProcessorCount, 512>>>(d, c, cs.d, cs.f, 1.0000001, 0.5);
    cudaDeviceSynchronize();
    long long h[2]; cudaMemcpy(h, c, 16, cudaMemcpyDeviceToHost);
    printf("%-28s", cs.name);
    if (cs.d) printf("  DMMA: %.1f cycles each (1 warp/SMSP)", (double)h[0] / (cs.d * 12.0));
    if (cs.f) printf("  FFMA2: %.2f cycles each per warp = %.2f per SMSP (3 warps/SMSP)", (double)h[1] / (cs.f * 8.0), (double)h[1] / (cs.f * 8.0) / 3.0);
    printf("\n");
  }
  return 0;
}

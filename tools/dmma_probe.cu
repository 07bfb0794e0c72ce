// DMMA (mma.sync.m8n8k4.f64) throughput on one SM sub-partition as a function of the number of
// resident warps and of independent accumulator chains per warp.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/dmma_probe tools/dmma_probe.cu && /tmp/dmma_probe
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
template <int CH>
__global__ void k(double* out, int iters, double a, double b) {
  double acc[CH][2];
#pragma unroll
  for (int i = 0; i < CH; ++i) acc[i][0] = acc[i][1] = threadIdx.x + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i) dmma(acc[i][0], acc[i][1], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < CH; ++i) s += acc[i][0] + acc[i][1];
  if (s == 1.2345) out[0] = s;
}
template <int CH>
void run(int warps_per_sm, int nsm, double* d) {
  const int iters = 20000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int r = 0; r < 3; ++r) {
    cudaEventRecord(e0);
    k<CH><<<nsm, 32 * warps_per_sm>>>(d, iters, 1.0000001, 0.5);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  int dev; cudaGetDevice(&dev); int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev);
  const double cycles = best * 1e-3 * clk * 1e3;
  const double dm_per_smsp = (double)iters * CH * warps_per_sm / 4.0;
  printf("chains/warp %2d  warps/SM %2d (%.1f/SMSP): %.1f cycles per DMMA per SMSP  (%.1f cycles between DMMAs of one warp)  %.1f TFLOP/s\n",
         CH, warps_per_sm, warps_per_sm / 4.0, cycles / dm_per_smsp, cycles / ((double)iters * CH),
         2.0 * 256 * iters * CH * warps_per_sm * (double)nsm / (best * 1e-3) / 1e12);
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  double* d; cudaMalloc(&d, 64);
  const int nsm = p.multiProcessorCount;
  for (int w : {4, 8, 16, 32}) { run<1>(w, nsm, d); run<4>(w, nsm, d); run<12>(w, nsm, d); }
  return 0;
}

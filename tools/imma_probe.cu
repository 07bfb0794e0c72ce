// Legacy (mma.sync) INT8 tensor path on B200: rate of mma.sync.m16n8k32.s32.{u8,s8}.s8.s32 on one SM sub-partition,
// alone and next to packed FP32 (FFMA2) issued from other warps of the same sub-partition -- the question being whether
// an exact-integer Hankel contraction issued with mma.sync could overlap the P3D cell evaluation, which the FP64 DMMA
// does not (profiles/pipe_overlap_r01.txt).  Also: TF32 m16n8k8 and BF16 m16n8k16 for scale.
//   warps 0..W-1 (W = 4 * wps): `NCH` independent MMA chains each;   warps 4*wps .. 15: FFMA2, 8 chains each
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int KIND>
__device__ __forceinline__ void mma(int (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  if (KIND == 0)
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  else if (KIND == 1)
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  else if (KIND == 2)
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  else
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
template <int KIND, int NCH>
__global__ void k(int* out, long long* cyc, int it_m, int it_f, int mma_warps, uint32_t seed) {
  const int warp = threadIdx.x >> 5;
  if (warp < mma_warps) {
    if (it_m == 0) return;
    int acc[NCH][4];
#pragma unroll
    for (int i = 0; i < NCH; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0;
    uint32_t a[4] = {seed, seed * 3u, seed * 5u, seed * 7u}, b[2] = {seed * 11u, seed * 13u};
    long long t0 = clock64();
    for (int it = 0; it < it_m; ++it) {
#pragma unroll
      for (int i = 0; i < NCH; ++i) mma<KIND>(acc[i], a, b);
    }
    long long t1 = clock64();
    int s = 0;
#pragma unroll
    for (int i = 0; i < NCH; ++i) s += acc[i][0] + acc[i][1] + acc[i][2] + acc[i][3];
    if (s == 123457) out[0] = s;
    if (blockIdx.x == 0 && threadIdx.x == 0) cyc[0] = t1 - t0;
  } else {
    if (it_f == 0) return;
    float2 x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = make_float2(threadIdx.x + i, i);
    const float2 y = make_float2(1.0000001f, 1.0000001f), z = make_float2(0.5f, 0.5f);
    long long t0 = clock64();
    for (int it = 0; it < it_f; ++it) {
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = __ffma2_rn(x[i], y, z);
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i].x + x[i].y;
    if (s == 1.2345f) out[1] = (int)s;
    if (blockIdx.x == 0 && threadIdx.x == 32 * mma_warps) cyc[1] = t1 - t0;
  }
}
template <int KIND, int NCH>
void run(const char* name, int sms, int* d, long long* c, double ops_per_mma) {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int NM = 4000;
  for (int wps = 1; wps <= 2; ++wps) {              // MMA warps per sub-partition
    const int mw = 4 * wps, fw = 16 - mw;
    const int NF = NM * NCH * 4;                     // FFMA2 iterations of comparable duration
    struct { int m, f; const char* what; } cases[] = {{NM, 0, "alone"}, {NM, NF, "beside FFMA2"}, {0, NF, "FFMA2 alone"}};
    for (auto& cs : cases) {
      cudaMemset(c, 0, 16);
      k<KIND, NCH><<<sms, 512>>>(d, c, cs.m, cs.f, mw, 0x01020304u);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
      long long h[2]; cudaMemcpy(h, c, 16, cudaMemcpyDeviceToHost);
      printf("%-34s %d MMA warp(s)/SMSP, %d chains  %-14s", name, wps, NCH, cs.what);
      if (cs.m) {
        const double cyc_per = (double)h[0] / (cs.m * (double)NCH) / wps;      // cycles per MMA per sub-partition
        printf("  MMA: %6.2f cycles each per SMSP = %7.1f Tops/s at %d SMs x %.2f GHz", cyc_per,
               ops_per_mma / cyc_per * 4 * sms * p.clockRate * 1e3 / 1e12, sms, p.clockRate / 1e6);
      }
      if (cs.f) printf("  FFMA2: %.2f cycles each per SMSP (%d warps/SMSP)", (double)h[1] / (cs.f * 8.0) / (fw / 4.0), fw / 4);
      printf("\n");
    }
  }
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int* d; long long* c; cudaMalloc(&d, 64); cudaMalloc(&c, 16);
  const int sms = p.multiProcessorCount;
  run<0, 8>("m16n8k32 s8.s8 -> s32", sms, d, c, 2.0 * 16 * 8 * 32);
  run<1, 8>("m16n8k32 u8.s8 -> s32", sms, d, c, 2.0 * 16 * 8 * 32);
  run<0, 2>("m16n8k32 s8.s8 -> s32", sms, d, c, 2.0 * 16 * 8 * 32);
  run<2, 8>("m16n8k8 tf32 -> f32", sms, d, c, 2.0 * 16 * 8 * 8);
  run<3, 8>("m16n8k16 bf16 -> f32", sms, d, c, 2.0 * 16 * 8 * 16);
  return 0;
}

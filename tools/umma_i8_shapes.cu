// Rate of back-to-back tcgen05.mma kind::i8 (M = 128, K = 32) as a function of N, operands resident in
// shared memory (K-major, no swizzle), one issuing thread per SM, all SMs busy.  Complements
// tools/umma_i8_probe.cu: k_window_tc issues N = 256 / 192 / 128 / 64 per K step.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ uint64_t desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return static_cast<uint64_t>((addr & 0x3FFFFu) >> 4) | (static_cast<uint64_t>(lbo >> 4) << 16) | (static_cast<uint64_t>(sbo >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ uint32_t idesc(int m, int n) { return (2u << 4) | (1u << 10) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24); }
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(id), "r"(acc) : "memory");
}
__global__ void __launch_bounds__(128, 1) k(int n, int rounds, int lds_warps, long long* cyc, float* sink) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < (4096 + 320 * 32) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = i * 2654435761u;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot;
  if (threadIdx.x == 0) {
    const uint64_t ad = desc(smem_u32(smem), 2048, 128), bd = desc(smem_u32(smem + 4096), (320 / 8) * 128, 128);
    const long long t0 = clock64();
    for (int r = 0; r < rounds; ++r) mma(tmem, ad, bd, idesc(128, n), r > 0);
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(smem_u32(&bar)) : "memory");
    if (blockIdx.x == 0) cyc[0] = clock64() - t0;
  } else if (threadIdx.x >= 32 && threadIdx.x < 32 + 32 * lds_warps) {
    // shared-memory load traffic beside the MMAs (what the worker warps of the real kernels generate)
    float4 acc = make_float4(0, 0, 0, 0);
    const float4* p = reinterpret_cast<const float4*>(smem + 16384) + threadIdx.x;
    for (int r = 0; r < rounds * 8; ++r) {
      const float4 v = p[(r & 63) * 128];
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    if (acc.x + acc.y + acc.z + acc.w == 1.2345f) sink[0] = acc.x;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  long long* c; float* s; cudaMalloc(&c, 8); cudaMalloc(&s, 4);
  const int smem = 16384 + 64 * 128 * 16 + 4096;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int R = 4000;
  for (int lds = 0; lds <= 3; lds += 3)
    for (int n : {32, 64, 128, 192, 256}) {
      k<<<p.multiProcessorCount, 128, smem>>>(n, R, lds, c, s);
      cudaError_t e = cudaDeviceSynchronize();
      long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
      printf("M 128 N %3d K 32, %d LDS.128 warps beside: %s  %.1f cycles per MMA  %.0f INT8 MAC/clk/SM\n", n, lds, cudaGetErrorString(e), (double)h / R, 128.0 * n * 32 * R / h);
    }
  return 0;
}

// k_window_tc: window convolution + theta rebinning + whitening + chi2 on the 5th-generation tensor cores.
//
// Same mathematics as k_window<MT, CHI2 = true> (likelihood.py:53-88, 170-194):
//   chi2_A[pt] = || yd[A] - Y ||^2,   Y[pt][M] = sum_{pairs p in A} sum_m T_chi[p][M][m] * Px_Zam[pt][a_p][m]
// with the FP64 contraction done by an error-free integer split on tcgen05.mma kind::i8 (INT32
// accumulators in tensor memory), as in cpx_hankel_tc.cuh.  Here both operands are plain matrices, so
// the split costs a few instructions per *input* element and nothing per multiply-add:
//
//   * T_chi is a constant of the batch:  T[p][M][m] = 2^c(p, m/16) * rs[A][M] * sum_j B_j 256^-(j+1)  with five
//     balanced base-256 digits (s8) stacked along N (5 x NB columns, NB = coarse-k count rounded up to 16).
//     The column scale 2^c ~ max_M |T[p][M][m]| over blocks of 16 fine-k columns is moved onto Px, so that
//     every Px element enters the fixed-point conversion with the size of its largest contribution to any
//     output -- Px falls by many decades along k_par while T rises with 1/sigma;
//   * the Px row of a (point, theta_A) -- all pairs, all m -- becomes 40-bit two's-complement fixed point
//     relative to its exact maximum exponent (every tile is streamed twice: the first read, from HBM,
//     only takes the maximum; the second, from L2, converts with one magic-number DFMA per element);
//     its five bytes are the operand slices (top byte s8, others u8);
//   * slice pairs with i + j = s share TMEM columns NB*s .. NB*s + NB, pairs with i + j > 4 (< 2^-40) are
//     dropped; the epilogue combines the five INT32 per output by Horner in float64.
//
// CTA = 16 warps, one CTA per SM (5 x 64 accumulator columns round up to the whole tensor memory).  A tile is
// 128 points x one theta_A x one z; a CTA walks the theta_A bins blockIdx.y, blockIdx.y + gridDim.y, ... of its
// 128 points.  Roles (they synchronise through mbarriers and two shared-memory counters only):
//   warps 0-7   workers: convert the chunks of the second Px stream (32 KB = 128 points x 32 fine-k values, the
//               same bytes the scouts saw, now from L2), write one 16-byte core-matrix row per slice and chunk
//               into a 3-stage operand ring, and run the epilogue of the tile (tcgen05.ld, Horner, chi2);
//   warp 8      one thread issues the six tcgen05.mma of a chunk and commits them;
//   warp 9      one thread streams the digit tiles of T (1-D TMA bulk copies);
//   warps 10-11 one thread each streams Px_Zam with 3-D tiled TMA loads from a tensor map of the scratch buffer
//               (128-byte swizzle: the per-point reads are bank-conflict free) -- once for the scouts (this is
//               the read that comes from HBM), once for the workers;
//   warps 12-15 scouts: run one tile ahead and only take the largest exponent of every row (sE).
// HBM traffic is one read of Px_Zam; the kernel is bound by that read (profiles/k_window_tc_ncu_r01.txt).
#pragma once
#include "cpx_hankel_tc.cuh"

namespace cpx {

constexpr int kWtPts = 128;          // points per tile = UMMA M
constexpr int kWtNA = 5;             // Px bytes
constexpr int kWtNW = 5;             // T digits
#ifndef CPX_WT_STAGES
#define CPX_WT_STAGES 3
#endif
#ifndef CPX_WT_RING
#define CPX_WT_RING 2
#endif
#ifndef CPX_WT_SRING
#define CPX_WT_SRING 2
#endif
constexpr int kWtStages = CPX_WT_STAGES;   // operand stages (u8 slices of Px + digit tile of T)
constexpr int kWtRing = CPX_WT_RING;       // ring slots (32 KB chunks of Px) of the converting workers (second read: L2)
constexpr int kWtSRing = CPX_WT_SRING;     // ring slots of the scout warps (first read: HBM)
constexpr int kWtScouts = 4;            // scout warps
constexpr int kWtKC = 32;            // K per stage (one UMMA K step)
constexpr int kWtWorkers = 8;        // worker warps
constexpr int kWtThreads = 32 * (kWtWorkers + 4 + kWtScouts);   // workers, MMA issuer, 3 TMA threads (digit tiles, Px for the scouts, Px for the workers), scouts
constexpr int kWtASlice = kWtPts * kWtKC;               // 4 KB
constexpr int kWtAStage = kWtNA * kWtASlice;            // 20 KB
constexpr int kWtRingBytes = kWtWorkers * 32 * 16 * 8;  // 32 KB: 128 bytes per worker thread, laid out [8][256 threads][16 B]
__host__ __device__ constexpr int wt_b_stage(int nb) { return kWtNW * nb * kWtKC; }
__host__ __device__ constexpr int wt_smem_bytes(int nb) {
  return kWtStages * (kWtAStage + wt_b_stage(nb)) + (kWtRing + kWtSRing) * kWtRingBytes + 2 * kWtPts * 4 + 2 * kWtPts * 8 + 1024;   // + slack to align the ring to 1 KB
}

__device__ __forceinline__ uint32_t umma_idesc_i8(int m, int n, bool a_signed) {
  return (2u << 4) | ((a_signed ? 1u : 0u) << 7) | (1u << 10) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, int32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr));
}
// 3-D tiled TMA load (coordinates fastest dimension first) into shared memory, completion on an mbarrier
__device__ __forceinline__ void tma_load_3d(uint32_t dst_smem, const void* tmap, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst_smem),
      "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
      : "memory");
}
// the same with an L2 cache policy: the scouts' read (from HBM) is marked evict_last so that the line is still there when the
// workers read it again a tile later, the workers' read evict_first because nobody reads it a third time (CPX_WT_L2HINT)
__device__ __forceinline__ void tma_load_3d_hint(uint32_t dst_smem, const void* tmap, int c0, int c1, int c2, uint64_t* bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4}], [%5], %6;" ::"r"(dst_smem),
      "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
// CPX_WT_REVERSE: the workers walk the chunks of a tile in the opposite order to the scouts, so that their (second) read
// starts with the lines the scouts touched last -- the ones most likely still in L2 (integer accumulation: same result)
#ifndef CPX_WT_REVERSE
#define CPX_WT_REVERSE 0
#endif
#ifndef CPX_WT_L2HINT
#define CPX_WT_L2HINT 1
#endif
__device__ __forceinline__ void bar_workers() { asm volatile("bar.sync 2, %0;" ::"n"(kWtWorkers * 32) : "memory"); }

__global__ void __launch_bounds__(kWtThreads, 1) k_window_tc(const ZDev* __restrict__ ztab, int zi, WindowArgs w) {
  if (w.zslot < 0) {
    zi = blockIdx.z;
    w.zslot = blockIdx.z;
  }
  const ZDev& Z = ztab[zi];
  const int n_A = Z.n_A;
  if (static_cast<int>(blockIdx.y) >= n_A) return;
  const int NB = Z.wt_nb;
  const int B_STAGE = wt_b_stage(NB);
  const uint32_t A_CHUNK = (kWtPts / 8) * 128;              // bytes between the two 16-byte K chunks of a Px tile
  const uint32_t B_CHUNK = (kWtNW * NB / 8) * 128;

  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char* const sA = smem_raw;
  unsigned char* const sB = sA + kWtStages * kWtAStage;
  // float64 ring, 1 KB aligned for the 128-byte TMA swizzle: slot = [K half][128 points][128 bytes], the 16-byte unit
  // index XOR-ed with (point & 7) by the TMA engine
  unsigned char* const sR = sB + kWtStages * B_STAGE + ((1024u - (smem_u32(sB + kWtStages * B_STAGE) & 1023u)) & 1023u);
  unsigned char* const sS = sR + kWtRing * kWtRingBytes;                  // ring of the scouts
  int* const sE = reinterpret_cast<int*>(sS + kWtSRing * kWtRingBytes);   // [2][128] biased exponent of the row maximum (tile parity)
  double* const sChi = reinterpret_cast<double*>(sE + 2 * kWtPts);        // [2][128] partial chi2 of the two column halves
  __shared__ __align__(8) uint64_t bar_fullA[kWtStages], bar_fullB[kWtStages], bar_free[kWtStages], bar_done, bar_epi;
  __shared__ __align__(8) uint64_t bar_rfull[kWtRing], bar_rfree[kWtRing], bar_sfull[kWtSRing], bar_sfree[kWtSRing];
  __shared__ int s_scout_done, s_epi_done;   // counters: scout warps finished with a tile / epilogues completed
  __shared__ uint32_t tmem_slot;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int pt_base = blockIdx.x * kWtPts;
  const int cpp = Z.n_m_pad / kWtKC;                        // chunks per pair
  const int a_step = gridDim.y;                             // this CTA takes theta_A = blockIdx.y, + a_step, ...

  if (tid == 0) {
    for (int s = 0; s < kWtStages; ++s) {
      mbar_init(&bar_fullA[s], kWtWorkers);
      mbar_init(&bar_fullB[s], 1);
      mbar_init(&bar_free[s], 1);
    }
    mbar_init(&bar_done, 1);
    mbar_init(&bar_epi, kWtWorkers);
    for (int i = 0; i < kWtRing; ++i) {
      mbar_init(&bar_rfull[i], 1);
      mbar_init(&bar_rfree[i], kWtWorkers);
    }
    for (int i = 0; i < kWtSRing; ++i) {
      mbar_init(&bar_sfull[i], 1);
      mbar_init(&bar_sfree[i], kWtScouts);
    }
    s_scout_done = 0;
    s_epi_done = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kWtWorkers) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;

  if (warp == kWtWorkers) {
    // =============================== MMA issue (one thread) ===============================
    if (lane == 0) {
      const uint32_t a_base = smem_u32(sA), b_base = smem_u32(sB);
      int gc = 0, tile = 0;
      for (int A = blockIdx.y; A < n_A; A += a_step, ++tile) {
        const int n_chunks = (Z.pair_start[A + 1] - Z.pair_start[A]) * cpp;
        // the single accumulator is free once the workers have read the previous tile out of tensor memory
        if (tile > 0) {
          mbar_wait_parked(&bar_epi, static_cast<uint32_t>(tile - 1) & 1u);
          tc_fence_after();
        }
        for (int c = 0; c < n_chunks; ++c, ++gc) {
          const int s = gc % kWtStages;
          const uint32_t ph = static_cast<uint32_t>(gc / kWtStages) & 1u;
          mbar_wait_parked(&bar_fullB[s], ph);
          mbar_wait_parked(&bar_fullA[s], ph);
          tc_fence_after();
          const uint64_t bd = umma_smem_desc(b_base + s * B_STAGE, B_CHUNK, 128);
#pragma unroll
          for (int i = 0; i < kWtNA; ++i) {
            const uint64_t ad = umma_smem_desc(a_base + s * kWtAStage + i * kWtASlice, A_CHUNK, 128);
            const int n = NB * (kWtNW - i);                    // digits 0 .. 4 - i
            const uint32_t acc = (c > 0 || i > 0) ? 1u : 0u;
            if (n <= 256) {
              umma_i8(tmem + NB * i, ad, bd, umma_idesc_i8(kWtPts, n, i == 0), acc);
            } else {                                           // N = 320: digits 0-3, then digit 4
              umma_i8(tmem + NB * i, ad, bd, umma_idesc_i8(kWtPts, 4 * NB, i == 0), acc);
              const uint64_t bd4 = umma_smem_desc(b_base + s * B_STAGE + (4 * NB / 8) * 128, B_CHUNK, 128);
              umma_i8(tmem + NB * (i + 4), ad, bd4, umma_idesc_i8(kWtPts, NB, i == 0), acc);
            }
          }
          umma_commit(&bar_free[s]);
        }
        umma_commit(&bar_done);
      }
    }
    __syncwarp();
  } else if (warp == kWtWorkers + 1) {
    // =============================== TMA: digit tiles of T ===============================
    if (lane == 0) {
      int gc = 0;
      for (int A = blockIdx.y; A < n_A; A += a_step) {
        const int n_chunks = (Z.pair_start[A + 1] - Z.pair_start[A]) * cpp;
        const unsigned char* src = reinterpret_cast<const unsigned char*>(Z.wt_dig) + static_cast<size_t>(Z.wt_off[A]) * B_STAGE;
        for (int c = 0; c < n_chunks; ++c, ++gc) {
          const int s = gc % kWtStages;
          if (gc >= kWtStages) {
            const uint32_t ph = static_cast<uint32_t>(gc / kWtStages - 1) & 1u;
            mbar_wait_parked(&bar_free[s], ph);
          }
          mbar_expect_tx(&bar_fullB[s], static_cast<uint32_t>(B_STAGE));
          const int cw = CPX_WT_REVERSE ? n_chunks - 1 - c : c;      // the workers' chunk order
          bulk_g2s(sB + s * B_STAGE, src + static_cast<size_t>(cw) * B_STAGE, static_cast<uint32_t>(B_STAGE), &bar_fullB[s]);
        }
      }
    }
    __syncwarp();
  } else if (warp == kWtWorkers + 2 || warp == kWtWorkers + 3) {
    // =============================== TMA: the two Px streams ===============================
    // Every tile (theta_A) is read twice, chunk by chunk (c over (pair, m / 32)): once for the scout warps, which only
    // take the largest exponent of each row and run a tile ahead (this is the read that comes from HBM), and once
    // for the converting workers (the same bytes again, from L2).
    if (lane == 0) {
      const bool scout = warp == kWtWorkers + 2;
      const int nslots = scout ? kWtSRing : kWtRing;
      uint64_t* const full = scout ? bar_sfull : bar_rfull;
      uint64_t* const free_ = scout ? bar_sfree : bar_rfree;
      const uint32_t ring = smem_u32(scout ? sS : sR);
#if CPX_WT_L2HINT
      const uint64_t l2pol = scout ? l2_policy_evict_last() : l2_policy_evict_first();
#endif
      int rc = 0;
      const bool rev = CPX_WT_REVERSE && !scout;
      for (int A = blockIdx.y; A < n_A; A += a_step)
        for (int pi = Z.pair_start[A]; pi < Z.pair_start[A + 1]; ++pi) {
          const int p = rev ? Z.pair_start[A] + Z.pair_start[A + 1] - 1 - pi : pi;
          const int a = Z.pair_a[p];
          for (int mi = 0; mi < cpp; ++mi, ++rc) {
            const int mc = rev ? cpp - 1 - mi : mi;
            const int slot = rc % nslots;
            if (rc >= nslots) mbar_wait_parked(&free_[slot], static_cast<uint32_t>(rc / nslots - 1) & 1u);
            mbar_expect_tx(&full[slot], static_cast<uint32_t>(kWtRingBytes));
#if CPX_WT_L2HINT
            tma_load_3d_hint(ring + slot * kWtRingBytes, Z.px_tmap, mc * kWtKC, a, pt_base, &full[slot], l2pol);
            tma_load_3d_hint(ring + slot * kWtRingBytes + kWtRingBytes / 2, Z.px_tmap, mc * kWtKC + 16, a, pt_base, &full[slot], l2pol);
#else
            tma_load_3d(ring + slot * kWtRingBytes, Z.px_tmap, mc * kWtKC, a, pt_base, &full[slot]);
            tma_load_3d(ring + slot * kWtRingBytes + kWtRingBytes / 2, Z.px_tmap, mc * kWtKC + 16, a, pt_base, &full[slot]);
#endif
          }
        }
    }
    __syncwarp();
  } else if (warp >= kWtWorkers + 4) {
    // =============================== scouts: row maxima, one tile ahead of the workers ===============================
    const int r = tid - 32 * (kWtWorkers + 4);               // point of the tile (both K halves)
    uint32_t soff[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) soff[q] = (q >> 3) * (kWtRingBytes / 2) + r * 128 + (((q & 7) ^ (r & 7)) << 4);
    int rc = 0, tile = 0;
    for (int A = blockIdx.y; A < n_A; A += a_step, ++tile) {
      const int p0 = Z.pair_start[A], n_chunks = (Z.pair_start[A + 1] - p0) * cpp;
      const int* const cs = Z.wt_cs + 2 * p0 * cpp;          // exponent of the column scale per (pair, 16 columns)
      // sE[tile parity] was last read by the epilogue of tile - 2
      if (tile >= 2) while (*reinterpret_cast<volatile int*>(&s_epi_done) < tile - 1) __nanosleep(64);
      int emax = 0;
      int2 cs_next = *reinterpret_cast<const int2*>(cs);       // column-scale exponents of the two K halves, one chunk ahead
      for (int c = 0; c < n_chunks; ++c, ++rc) {
        const int2 cs_cur = cs_next;
        if (c + 1 < n_chunks) cs_next = *reinterpret_cast<const int2*>(cs + 2 * (c + 1));
        mbar_wait(&bar_sfull[rc % kWtSRing], static_cast<uint32_t>(rc / kWtSRing) & 1u);
        const unsigned char* vsrc = sS + (rc % kWtSRing) * kWtRingBytes;
        int h0 = 0, h1 = 0;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const uint4 v0 = *reinterpret_cast<const uint4*>(vsrc + soff[q]);
          const uint4 v1 = *reinterpret_cast<const uint4*>(vsrc + soff[8 + q]);
          h0 = max(h0, max(static_cast<int>(v0.y & 0x7fffffffu), static_cast<int>(v0.w & 0x7fffffffu)));
          h1 = max(h1, max(static_cast<int>(v1.y & 0x7fffffffu), static_cast<int>(v1.w & 0x7fffffffu)));
        }
        if ((h0 >> 20) > 0) emax = max(emax, (h0 >> 20) + cs_cur.x);
        if ((h1 >> 20) > 0) emax = max(emax, (h1 >> 20) + cs_cur.y);
        __syncwarp();
        if (lane == 0) mbar_arrive1(&bar_sfree[rc % kWtSRing]);
      }
      sE[(tile & 1) * kWtPts + r] = min(max(emax, 64), 2000);
      __threadfence_block();
      __syncwarp();
      if (lane == 0) atomicAdd(&s_scout_done, 1);
    }
  } else {
    // =============================== workers: slicing of Px, then the epilogue ===============================
    const int h = lane >> 4, r = 16 * warp + (lane & 15);    // half of the 32-wide K chunk, point of the tile
    unsigned char* const arow = sA + (h * 16 + (r >> 3)) * 128 + (r & 7) * 16;
    const int er = 32 * (warp & 3) + lane, hh = warp >> 2;   // epilogue: TMEM lane = point, column half
    const int ept = pt_base + er;
    const int d = (w.data_idx != nullptr && ept < w.count) ? min(max(w.data_idx[ept], 0), Z.n_data - 1) : 0;
    const uint32_t taddr = tmem + (static_cast<uint32_t>(32 * (warp & 3)) << 16);
    uint32_t roff[8];                                        // byte offsets of this thread's 8 swizzled units within a ring slot
#pragma unroll
    for (int q = 0; q < 8; ++q) roff[q] = h * (kWtRingBytes / 2) + r * 128 + ((q ^ (r & 7)) << 4);

    int gc = 0, rc = 0, tile = 0;     // operand-stage counter, ring counter, tile counter

    for (int A = blockIdx.y; A < n_A; A += a_step, ++tile) {
      const int p0 = Z.pair_start[A], n_chunks = (Z.pair_start[A + 1] - p0) * cpp;
      const int* const cs = Z.wt_cs + 2 * p0 * cpp + h;
      // the scouts have the row maxima of this tile ready (they run one tile ahead)
      while (*reinterpret_cast<volatile int*>(&s_scout_done) < kWtScouts * (tile + 1)) __nanosleep(32);
      __threadfence_block();
      const int emax = sE[(tile & 1) * kWtPts + r];
      // ---- convert, slice, store ----
      int cs_next = cs[CPX_WT_REVERSE ? 2 * (n_chunks - 1) : 0];      // column-scale exponent, loaded one chunk ahead
      for (int c = 0; c < n_chunks; ++c, ++gc, ++rc) {
        const int s = gc % kWtStages;
        const int cs_cur = cs_next;
        if (c + 1 < n_chunks) cs_next = cs[2 * (CPX_WT_REVERSE ? n_chunks - 2 - c : c + 1)];
        mbar_wait(&bar_rfull[rc % kWtRing], static_cast<uint32_t>(rc / kWtRing) & 1u);
        // |x 2^c| < 2^(emax - 1022)  =>  |x 2^(c + 1061 - emax)| < 2^39
        const int sh = min(max(cs_cur + 1061 - emax, -1000), 1000);
        const double mult = __hiloint2double((1023 + sh) << 20, 0);
        const unsigned char* vsrc = sR + (rc % kWtRing) * kWtRingBytes;
        uint32_t wd[kWtNA][4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const double2 va = *reinterpret_cast<const double2*>(vsrc + roff[2 * q]);
          const double2 vb = *reinterpret_cast<const double2*>(vsrc + roff[2 * q + 1]);
          uint32_t L[4], H[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const double x = k == 0 ? va.x : (k == 1 ? va.y : (k == 2 ? vb.x : vb.y));
            const double y = fma(x, mult, 6755399441055744.0);      // 1.5 * 2^52: low 40 bits = rni(x mult), two's complement
            L[k] = static_cast<uint32_t>(__double2loint(y));
            H[k] = static_cast<uint32_t>(__double2hiint(y));
          }
          const uint32_t xa = prmt(L[0], L[1], 0x6273), ya = prmt(L[0], L[1], 0x4051);
          const uint32_t xb = prmt(L[2], L[3], 0x6273), yb = prmt(L[2], L[3], 0x4051);
          wd[0][q] = prmt(prmt(H[0], H[1], 0x4040), prmt(H[2], H[3], 0x4040), 0x5410);   // bits 32..39, signed
          wd[1][q] = prmt(xa, xb, 0x5410);                                               // byte 3
          wd[2][q] = prmt(xa, xb, 0x7632);
          wd[3][q] = prmt(ya, yb, 0x5410);
          wd[4][q] = prmt(ya, yb, 0x7632);                                               // byte 0
        }
        if (gc >= kWtStages) mbar_wait(&bar_free[s], static_cast<uint32_t>(gc / kWtStages - 1) & 1u);
#pragma unroll
        for (int i = 0; i < kWtNA; ++i)
          *reinterpret_cast<uint4*>(arow + s * kWtAStage + i * kWtASlice) = make_uint4(wd[i][0], wd[i][1], wd[i][2], wd[i][3]);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive1(&bar_fullA[s]);
          mbar_arrive1(&bar_rfree[rc % kWtRing]);               // ring slot may be refilled
        }
      }

      // ---- epilogue: TMEM lane = point, warp w reads lanes 32 (w % 4) .. and the column half w / 4 ----
      mbar_wait(&bar_done, static_cast<uint32_t>(tile) & 1u);
      tc_fence_after();
      const double* const yd = Z.yd + (static_cast<size_t>(d) * n_A + A) * Z.n_M_pad;
      const double* const rs = Z.wt_rs + static_cast<size_t>(A) * NB;
      // Y = rs[M] 2^(e - 1069) sum_s D_s 256^(4-s)
      const double rowfac = __hiloint2double((sE[(tile & 1) * kWtPts + er] - 1069 + 1023) << 20, 0);
      const int c_end = min((hh + 1) * (NB / 2), Z.n_M_pad);
      double chi = 0.0;
      for (int col = hh * (NB / 2); col < c_end; col += 8) {     // NB / 2 and n_M_pad are multiples of 8
        int32_t v[kWtNW][8];
#pragma unroll
        for (int sg = 0; sg < kWtNW; ++sg) tmem_ld8(taddr + NB * sg + col, v[sg]);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const double hi = fma(fma(i2d(v[0][j]), 256.0, i2d(v[1][j])), 256.0, i2d(v[2][j]));
          const double lo = fma(i2d(v[3][j]), 256.0, i2d(v[4][j]));
          const double y = fma(hi, 65536.0, lo) * (rowfac * rs[col + j]);
          const double res = yd[col + j] - y;
          chi = fma(res, res, chi);
        }
      }
      sChi[hh * kWtPts + er] = chi;
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive1(&bar_epi);                   // tensor memory may be overwritten by the next tile
      bar_workers();
      if (warp < 4 && ept < w.count)
        w.chi2_zA[(static_cast<size_t>(ept) * w.n_zl + w.zslot) * w.nA_max + A] = sChi[er] + sChi[kWtPts + er];
      bar_workers();                                           // sChi is reused by the next tile
      if (tid == 0) atomicAdd(&s_epi_done, 1);                 // sE[tile parity] may be rewritten by the scouts
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kWtWorkers) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
}

}  // namespace cpx

// k_p3d_hankel_tc: the P3D -> Px Hankel contraction on the 5th-generation tensor cores.
//
// Same mathematics as k_p3d_hankel (theory.py:250-344 + the theta average of likelihood.py:99-130):
//   Px_Zam[pt][a][m] = amp(m) * sum_q W[a][q] * P3D'(m, q; pt)
// but the FP64 contraction, which on B200 runs at the plain DFMA rate and shares the issue port with
// the FP32 cell evaluation, is replaced by an *error-free integer split* executed
// by tcgen05.mma kind::i8 with INT32 accumulators in tensor memory:
//
//   * the weights are constants of the batch:  W[a][q] = s_q * t_a * sum_j B_j[a][q] 256^-(j+1) with
//     balanced base-256 digits B_j in [-128, 127] (5 digits = 40 bits), s_q = 2^round(log2 max_a |W[a][q]|)
//     folded into the linear-power constant of the cell, t_a applied in the epilogue;
//   * each FP32 cell is turned into a 32-bit fixed-point number relative to the exact maximum of its
//     (point, k_par row): U = rni(cell * 2^(158 - e)), e = biased exponent of the row maximum.  The
//     four bytes of U are the four u8 operand slices A_0..A_3 -- no further arithmetic;
//   * sum_q A_i[q] B_j[q] is exact in INT32 (<= 4 * 96 * 255 * 128 < 2^24 per accumulator); all slice
//     pairs with i + j = s accumulate into the same 32 TMEM columns (digit blocks are stacked along N,
//     slice i starts at column 32 i), pairs with i + j > 4 (< 2^-40 of the row scale) are not issued;
//   * the epilogue reads the five INT32 partial sums per output, combines them by Horner in float64
//     (exact to 2^-53) and applies 2^(e - 174) t_a amp cont.
//
// Error of the split against the float64 dot product of the same FP32 cells: <= 1e-9 of the Px scale
// per theta bin over the whole ForestFlow parameter box and far outside it (tests/test_tc_split.py
// emulates the integer pipeline in NumPy); the FP32 cell rounding itself is 3e-8 .. 1e-7.
//
// CTA = 17 warps, one CTA per SM, persistent over a contiguous range of work items
// (z, 8-row k_par tile, theta block of 32, group of 16 parameter points):
//   warps 0-15  one parameter point each: lane = (row, quarter) evaluates NPL packed cell pairs of its row
//               into registers (a quarter of the k_perp nodes), the four quarters exchange the row
//               maximum, then the cells are scaled, converted (F2I), byte-transposed (PRMT) and stored
//               as the four K-major u8 operand tiles [128 x KP] (canonical no-swizzle core-matrix
//               order, conflict-free STS.64).  Four compute warps per SM sub-partition hide the FP32
//               and MUFU latencies; 22 live cells per lane keep the kernel under 96 registers;
//   warp 16     one thread issues the NKC x 4 tcgen05.mma (M = 128, N = 160/128/96/64, K = 32) of an item
//               as soon as all 16 compute warps have delivered its operands, and commits them to an
//               mbarrier; the tensor pipe works on item n while the compute warps evaluate item n+1;
//   epilogue    shared by the compute warps: after storing the operands of item n, warp w reads its
//               share (TMEM lanes 32 (w % 4).., theta columns of group w / 4) of the accumulators of
//               item n-1 with tcgen05.ld, applies Horner, scaling and the contaminant terms and stores
//               Px_Zam.  Every warp does the same amount of work, and no role waits for another one
//               except at the two mbarriers (operands ready, accumulators ready).
// Operand tiles and accumulators are double buffered (2 x 48 KB shared memory, 2 x 256 TMEM columns);
// the tensor pipe needs ~900 cycles per item and does not take SM issue slots (tools/umma_i8_probe.cu).
#pragma once
#include "cpx_kernels.cuh"

namespace cpx {

constexpr int kTcRows = 8;                  // k_par rows per tile
constexpr int kTcPts = 16;                  // parameter points per item (= compute warps)
constexpr int kTcQuarters = 32 / kTcRows;   // lanes per row
constexpr int kTcM = kTcRows * kTcPts;      // 128 = UMMA M
constexpr int kTcNA = 4;                    // cell bytes (u8 slices)
constexpr int kTcNW = 5;                    // weight digits (s8 slices)
constexpr int kTcNB = 32;                   // theta columns per block
constexpr int kTcBRows = kTcNW * kTcNB;     // 160 = rows of the stacked digit tile
constexpr int kTcThreads = 32 * (kTcPts + 1);   // 16 compute/epilogue warps + the MMA issuer
constexpr int kTcStageCols = 256;           // TMEM columns per accumulator stage (160 used)
constexpr int kTcConsts = 6;
constexpr int kTcStageBytes = 256;          // per-warp parameter staging: 64 B of PtPar + the 8 RowPar of the tile

// np = packed cell pairs per lane (n_q <= 8 np); K range padded to 64 or 96; quarter stride of the constants table
// padded to an odd number of pairs so that the four 64-byte row segments a warp reads fall into distinct banks
__host__ __device__ constexpr int tc_kp(int np) { return np <= 8 ? 64 : 96; }
__host__ __device__ constexpr int tc_qs(int np) { return np | 1; }
__host__ __device__ constexpr int tc_const_bytes(int np) { return kTcConsts * kTcQuarters * tc_qs(np) * kTcRows * 8; }
__host__ __device__ constexpr int tc_smem_bytes(int np) {
  return tc_const_bytes(np) + kTcBRows * tc_kp(np) + 2 * kTcNA * kTcM * tc_kp(np) + 3 * kTcM * 4 + 3 * kTcM * 8 + 2 * kTcPts * kTcStageBytes;
}
// byte offset of element (row r, K index k) in a K-major no-swizzle operand tile with `groups` 8-row groups
__host__ __device__ constexpr int tc_canon(int r, int k, int groups) {
  return ((k >> 4) * groups + (r >> 3)) * 128 + (r & 7) * 16 + (k & 15);
}

__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  // start >> 4 at [0,14), leading (K-chunk) byte offset >> 4 at [16,30), stride (8-row group) byte offset >> 4 at [32,46),
  // descriptor version 1 at [46,48), no swizzle
  return static_cast<uint64_t>((addr & 0x3FFFFu) >> 4) | (static_cast<uint64_t>(lbo >> 4) << 16) |
         (static_cast<uint64_t>(sbo >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ uint32_t umma_idesc_u8s8(int m, int n) {
  // D = S32 (2) at [4,6); A = U8 (0) at [7,10); B = S8 (1) at [10,13); both K-major; N >> 3 at [17,23); M >> 4 at [24,29)
  return (2u << 4) | (1u << 10) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}
__device__ __forceinline__ void umma_i8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive1(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld1(uint32_t taddr, int32_t& v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, int32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void bar_compute() { asm volatile("bar.sync 1, %0;" ::"n"(kTcPts * 32) : "memory"); }
__device__ __forceinline__ uint32_t f2u_rni(float x) {
  uint32_t u;
  asm("cvt.rni.u32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return u;
}
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
  return d;
}
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// wait that parks the thread in hardware (suspend-time hint, wakes up as soon as the phase completes): for the
// single-thread MMA-issue and TMA loops, which would otherwise burn issue slots polling
__device__ __forceinline__ void mbar_wait_parked(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAITP_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
      "@p bra DONEP_%=;\n\t"
      "bra WAITP_%=;\n\t"
      "DONEP_%=:\n\t}" ::"r"(smem_u32(bar)),
      "r"(parity), "r"(20000u)
      : "memory");
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
// exact int32 -> float64 on the ALU + FP64 pipes (I2F.F64 runs at the MUFU rate)
__device__ __forceinline__ double i2d(int32_t v) {
  return __hiloint2double(0x43300000, v ^ 0x80000000) - 4503601774854144.0;   // 2^52 + 2^31
}

// first work item of a CTA: decode once with divisions (tile / theta-block counts of the tensor path)
__device__ __forceinline__ void tc_decode(const ZDev* __restrict__ ztab, int n_z, int n_pg, long long it0, int& zi, int& tile,
                                          int& ablk, int& pg) {
  zi = 0;
  while (zi + 1 < n_z && ztab[zi + 1].item_start <= it0) ++zi;
  long long r = it0 - ztab[zi].item_start;
  pg = static_cast<int>(r % n_pg);
  r /= n_pg;
  ablk = static_cast<int>(r % ztab[zi].tc_nablk);
  tile = static_cast<int>(r / ztab[zi].tc_nablk);
}
__device__ __forceinline__ void tc_next(int nablk, int nt8, int n_pg, int& zi, int& tile, int& ablk, int& pg) {
  if (++pg == n_pg) {
    pg = 0;
    if (++ablk == nablk) {
      ablk = 0;
      if (++tile == nt8) {
        tile = 0;
        ++zi;
      }
    }
  }
}

template <int NP, bool ADD>
__global__ void __launch_bounds__(kTcThreads, 1)
    k_p3d_hankel_tc(const ZDev* __restrict__ ztab, int n_z, int count, long long n_items) {
  constexpr int KP = tc_kp(NP);
  constexpr int NKC = KP / 32;                      // K = 32 steps per item
  constexpr int KB = KP / kTcQuarters;              // K bytes (cells) per lane
  constexpr int QS = tc_qs(NP);
  constexpr int A_SLICE = kTcM * KP;
  constexpr int A_STAGE = kTcNA * A_SLICE;
  constexpr int B_BYTES = kTcBRows * KP;
  constexpr int CONST_BYTES = tc_const_bytes(NP);
  constexpr uint32_t A_CHUNK = (kTcM / 8) * 128;    // bytes between 16-byte K chunks of a cell tile
  constexpr uint32_t B_CHUNK = (kTcBRows / 8) * 128;
  static_assert(2 * NP <= KB && KB % 8 == 0, "cells of one lane must fit its quarter of the K range");

  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char* const sConst = smem_raw;
  unsigned char* const sB = smem_raw + CONST_BYTES;
  unsigned char* const sA = sB + B_BYTES;
  double* const sAC = reinterpret_cast<double*>(sA + 2 * A_STAGE);   // [3][128] amp * cont of the row (0 for padded rows)
  int* const sX = reinterpret_cast<int*>(sAC + 3 * kTcM);            // [3][128] biased exponent of the row maximum
  unsigned char* const sP = reinterpret_cast<unsigned char*>(sX + 3 * kTcM);   // [2][16 warps][256 B] parameter staging
  __shared__ __align__(8) uint64_t bar_tma, bar_full[2], bar_dfull[2];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) ZDev s_Z;

  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int n_pg = (count + kTcPts - 1) / kTcPts;
  const long long it0 = n_items * blockIdx.x / gridDim.x;
  const long long it1 = n_items * (blockIdx.x + 1) / gridDim.x;
  if (it0 >= it1) return;
  const int n_local = static_cast<int>(it1 - it0);

  if (threadIdx.x == 0) {
    mbar_init(&bar_tma, 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(&bar_full[b], kTcPts);
      mbar_init(&bar_dfull[b], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kTcPts) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "n"(2 * kTcStageCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;

  if (warp == kTcPts) {
    // =============================== MMA issue (one thread) ===============================
    // Item n may overwrite accumulator stage n & 1 once every compute warp has arrived for it: a warp
    // arrives only after its epilogue share of item n - 1, and it did its share of item n - 2 before that.
    if (lane == 0) {
      const uint32_t a_base = smem_u32(sA), b_base = smem_u32(sB);
      for (int n = 0; n < n_local; ++n) {
        const int b = n & 1;
        while (!mbar_test(&bar_full[b], static_cast<uint32_t>(n >> 1) & 1u)) __nanosleep(64);   // do not spin on issue slots
        tc_fence_after();
        const uint32_t d0 = tmem + b * kTcStageCols;
#pragma unroll
        for (int kc = 0; kc < NKC; ++kc)
#pragma unroll
          for (int i = 0; i < kTcNA; ++i) {
            const uint64_t ad = umma_smem_desc(a_base + b * A_STAGE + i * A_SLICE + kc * 2 * A_CHUNK, A_CHUNK, 128);
            const uint64_t bd = umma_smem_desc(b_base + kc * 2 * B_CHUNK, B_CHUNK, 128);
            umma_i8(d0 + kTcNB * i, ad, bd, umma_idesc_u8s8(kTcM, kTcNB * (kTcNW - i)), (kc > 0 || i > 0) ? 1u : 0u);
          }
        umma_commit(&bar_dfull[b]);
      }
    }
    __syncwarp();
  } else {
    // =============================== cell evaluation + slicing + epilogue share ===============================
    const int row = lane & (kTcRows - 1), quarter = lane / kTcRows;
    const int r_mma = warp * kTcRows + row;                 // operand-tile row this lane produces
    // epilogue share: TMEM lanes 32 (warp % 4) + lane = operand rows of compute warps 4 (warp % 4) .. + 3
    const int e_rmma = 32 * (warp & 3) + lane;
    const int e_wpt = e_rmma / kTcRows, e_row = e_rmma % kTcRows, e_grp = warp >> 2;
    uint32_t tma_phase = 0;
    int zi, tile, ablk, pg;
    tc_decode(ztab, n_z, n_pg, it0, zi, tile, ablk, pg);
    int cur_z = -1, cur_tile = -1, cur_ablk = -1;
    int p_zi = 0, p_tile = 0, p_ablk = 0, p_pg = 0;           // previous item (its epilogue is still to do)

    // epilogue share of the item whose coordinates are (ez, etile, eablk, epg), local index m_idx
    auto epilogue = [&](int m_idx, int ez, int etile, int eablk, int epg) {
      const int b = m_idx & 1;
      const ZDev& Zg = (ez == cur_z) ? s_Z : ztab[ez];
      const int n_a = Zg.n_a, n_m_pad = Zg.n_m_pad;
      const int pt = epg * kTcPts + e_wpt;
      const int ptc = min(pt, count - 1);
      const int mrow = etile * kTcRows + e_row;
      const bool pt_ok = pt < count;
      const int a0 = eablk * kTcNB;
      const int na_here = min(kTcNB, n_a - a0);
      const int cpw = (na_here + 3) >> 2;                     // theta columns per warp group
      const int c_beg = e_grp * cpw, c_end = min(c_beg + cpw, na_here);
      const double* ta = Zg.tc_ta + a0;
      double* const outp = Zg.pxzam + (static_cast<size_t>(ptc) * n_a + a0) * n_m_pad + mrow;

      mbar_wait(&bar_dfull[b], static_cast<uint32_t>(m_idx >> 1) & 1u);
      tc_fence_after();
      const int slot = (m_idx % 3) * kTcM + e_rmma;
      // 2^(eb - 174) * amp * cont (the producer stored amp * cont = 0 for padded rows)
      const double rowfac = __hiloint2double((sX[slot] - 174 + 1023) << 20, 0) * sAC[slot];
      const uint32_t taddr = tmem + b * kTcStageCols + (static_cast<uint32_t>(32 * (warp & 3)) << 16);
      for (int c = c_beg; c < c_end; c += 4) {
        int32_t v[kTcNW][4];
#pragma unroll
        for (int sgrp = 0; sgrp < kTcNW; ++sgrp) tmem_ld4(taddr + kTcNB * sgrp + c, v[sgrp]);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (c + j < c_end) {
            // sum_s D_s 256^(4-s): the upper three and the lower two digits are exact on their own
            const double hi = fma(fma(i2d(v[0][j]), 256.0, i2d(v[1][j])), 256.0, i2d(v[2][j]));
            const double lo = fma(i2d(v[3][j]), 256.0, i2d(v[4][j]));
            double px = fma(hi, 65536.0, lo) * (rowfac * ta[c + j]);
            if (ADD) {
              const int a = a0 + c + j, n_m = Zg.n_m, flags = Zg.flags;
              if (mrow < n_m) {
                const PtPar& pr = Zg.ptpar[ptc];
                const double cont = Zg.rowcont != nullptr ? Zg.rowcont[static_cast<size_t>(ptc) * n_m_pad + mrow] : 1.0;
                double add = 0.0;
                if (flags & 2) {   // SiIII auto + cross from precomputed Hankel moments, theory.py:361-438
                  const size_t o = static_cast<size_t>(a) * n_m + mrow, st = static_cast<size_t>(n_a) * n_m;
                  const double bx = pr.b_X, btx = pr.beta_X;
                  add += bx * bx * (Zg.metal_auto[o] + 2.0 * btx * Zg.metal_auto[o + st] + btx * btx * Zg.metal_auto[o + 2 * st]);
                  add += pr.bias * bx * (Zg.metal_cross[o] + (pr.beta + btx) * Zg.metal_cross[o + st] + pr.beta * btx * Zg.metal_cross[o + 2 * st]);
                }
                if (flags & 4) add += pr.b_noise * Zg.sky_smooth[mrow] * Zg.sky_xi[a];   // theory.py:476-506
                px = fma(add, cont, px);
              }
            }
            if (pt_ok) outp[static_cast<size_t>(c + j) * n_m_pad] = px;
          }
        }
      }
      tc_fence_before();
    };

    // asynchronous copy of this warp's parameters of one item into its staging slot
    auto stage_params = [&](int st, const ZDev& Z, int pt_, int tile_) {
      unsigned char* dst = sP + (st * kTcPts + warp) * kTcStageBytes;
      if (lane < 4) {
        cp_async16(dst + 16 * lane, reinterpret_cast<const unsigned char*>(&Z.ptpar[pt_]) + 16 * lane);
      } else if (lane < 12 && (Z.flags & 9)) {      // 8 rows x 16 bytes
        const unsigned char* src = reinterpret_cast<const unsigned char*>(&Z.rowpar[static_cast<size_t>(pt_) * Z.n_m_pad + tile_ * kTcRows]);
        cp_async16(dst + 64 + 16 * (lane - 4), src + 16 * (lane - 4));
      }
      cp_async_commit();
    };
    bool staged = false;     // parameters of the current item already on their way to shared memory

    for (int n = 0; n < n_local; ++n) {
      const int b = n & 1;
      if (zi != cur_z) {
        bar_compute();
        const int* src = reinterpret_cast<const int*>(&ztab[zi]);
        int* dst = reinterpret_cast<int*>(&s_Z);
        for (int i = threadIdx.x; i < static_cast<int>(sizeof(ZDev) / 4); i += kTcPts * 32) dst[i] = src[i];
        bar_compute();
      }
      const ZDev& Z = s_Z;
      const bool new_tile = (zi != cur_z) || (tile != cur_tile);
      const bool new_w = (zi != cur_z) || (ablk != cur_ablk);
      if (new_tile || new_w) {
        bar_compute();   // every compute warp is done reading the previous constants
        if (threadIdx.x == 0) {
          // the MMAs of the previous item are the last readers of the digit tile
          if (new_w && n >= 1) mbar_wait(&bar_dfull[(n - 1) & 1], static_cast<uint32_t>((n - 1) >> 1) & 1u);
          fence_proxy_async();
          mbar_expect_tx(&bar_tma, (new_tile ? static_cast<uint32_t>(CONST_BYTES) : 0u) + (new_w ? static_cast<uint32_t>(B_BYTES) : 0u));
          if (new_tile) {
            const unsigned char* src = reinterpret_cast<const unsigned char*>(Z.tc_cell) + static_cast<size_t>(tile) * CONST_BYTES;
            for (uint32_t off = 0; off < static_cast<uint32_t>(CONST_BYTES); off += 16384u)
              bulk_g2s(sConst + off, src + off, min(16384u, static_cast<uint32_t>(CONST_BYTES) - off), &bar_tma);
          }
          if (new_w) bulk_g2s(sB, reinterpret_cast<const unsigned char*>(Z.tc_wdig) + static_cast<size_t>(ablk) * B_BYTES, B_BYTES, &bar_tma);
        }
        mbar_wait(&bar_tma, tma_phase);
        tma_phase ^= 1u;
        cur_z = zi;
        cur_tile = tile;
        cur_ablk = ablk;
      }

      const int ptc = min(pg * kTcPts + warp, count - 1);
      const int m = tile * kTcRows + row;
      const bool rowwise = (Z.flags & 9) != 0;
      // parameters come through the per-warp staging slot; the copy for the next item is started now and
      // lands while this item's cells are evaluated (the loads would otherwise stall every item start)
      if (!staged) stage_params(b, Z, ptc, tile);
      cp_async_wait<0>();
      __syncwarp();
      {
        int nzi = zi, ntile = tile, nablk = ablk, npg = pg;
        tc_next(Z.tc_nablk, Z.tc_nt, n_pg, nzi, ntile, nablk, npg);
        staged = (n + 1 < n_local) && nzi == zi;
        if (staged) stage_params(b ^ 1, Z, min(npg * kTcPts + warp, count - 1), ntile);
      }
      const unsigned char* const pst = sP + (b * kTcPts + warp) * kTcStageBytes;
      const float* const pp = reinterpret_cast<const float*>(pst);
      const float2 q1l = dup2(pp[0]), q1lo = dup2(pp[1]), q2l = dup2(pp[2]), q2lo = dup2(pp[3]);
      const float2 av = dup2(pp[4]), bv = dup2(pp[6]), nlkv = dup2(pp[7]), nlkvlo = dup2(pp[8]);
      const float2 ikp2l = dup2(pp[9]), ikp2lo = dup2(pp[10]);
      float bhi, blo;
      double ampcont;
      if (rowwise) {
        const RowPar& rp = reinterpret_cast<const RowPar*>(pst + 64)[row];
        bhi = rp.betaT_hi;
        blo = rp.betaT_lo;
        ampcont = rp.ampcont;
      } else {
        const double bias = *reinterpret_cast<const double*>(pst + 48), beta = *reinterpret_cast<const double*>(pst + 56);
        split_hi_lo(beta, bhi, blo);
        ampcont = bias * bias * Z.dAA;
      }
      if (m >= Z.n_m) ampcont = 0.0;
      const float2 betaT = dup2(bhi), betaTlo = dup2(kKaiserLoScale * blo);

      // ---- the lane's 2 NP cells (nodes quarter * 2 NP + 2 p + {0, 1}), kept in registers ----
      float2 x[NP];
      float rmax = 0.0f;
      const float4* cc = reinterpret_cast<const float4*>(sConst) + (quarter * QS) * kTcRows + row;
      constexpr int CS = kTcQuarters * QS * kTcRows;     // float4 stride between the three constant pairs
#pragma unroll
      for (int p = 0; p < NP; ++p) {
        const float4 c0 = cc[0 * CS + p * kTcRows], c1 = cc[1 * CS + p * kTcRows], c2 = cc[2 * CS + p * kTcRows];
        const float2 D = make_float2(c0.x, c0.y), L2K = make_float2(c0.z, c0.w), L2MU = make_float2(c1.x, c1.y);
        const float2 MU2 = make_float2(c1.z, c1.w), NK2 = make_float2(c2.x, c2.y), PL = make_float2(c2.z, c2.w);
        // identical arithmetic to k_p3d_hankel (see the comments there)
        const float2 nl = mul2(D, add2(fma2(q2l, D, q1l), fma2(q2lo, D, q1lo)));
        const float2 nv = exp2_rr2_neg(add2(fma2(av, L2K, fma2(bv, L2MU, nlkv)), nlkvlo));
        const float2 tt = fma2(NK2, ikp2lo, fma2(NK2, ikp2l, nl));
        const float2 e = kExpDnl(fma2(nl, nv, tt));
        float2 y = mul2(PL, e);
#if CPX_KAISER == 0
        y = fma2(mul2(y, betaT), MU2, y);
        y = fma2(mul2(y, betaT), MU2, fma2(mul2(y, betaTlo), MU2, y));
#else
        const float2 bm = fma2(betaT, MU2, mul2(betaTlo, MU2));
        y = fma2(y, bm, y);
        y = fma2(y, bm, y);
#endif
        x[p] = y;
        rmax = fmax3(rmax, y.x, y.y);
      }
#pragma unroll
      for (int o = kTcRows; o < 32; o <<= 1) rmax = fmaxf(rmax, __shfl_xor_sync(0xffffffffu, rmax, o));
      // fixed point relative to the row maximum: rmax < 2^(e - 126)  =>  cell * 2^(158 - e) < 2^32
      const uint32_t eb = min(max(__float_as_uint(rmax) >> 23, 31u), 254u);
      const float2 scale = dup2(__uint_as_float((285u - eb) << 23));

      // Operand stage b was last read by the MMAs of item n - 2, whose completion this warp observed
      // before its epilogue share of that item.  The exponent table has three slots because a slow
      // warp may still be reading slot n - 2 (item n - 2 is finished by everybody only at arrival n - 1).
      if (quarter == 0) {
        sX[(n % 3) * kTcM + r_mma] = static_cast<int>(eb);
        sAC[(n % 3) * kTcM + r_mma] = ampcont;
      }
      unsigned char* const arow = sA + b * A_STAGE + (r_mma >> 3) * 128 + (r_mma & 7) * 16;
#pragma unroll
      for (int g = 0; g < KB / 8; ++g) {          // 8 cells = 4 pairs -> one 8-byte store per slice
        uint32_t w[kTcNA][2];
#pragma unroll
        for (int q4 = 0; q4 < 2; ++q4) {
          const int p0 = 4 * g + 2 * q4;       // pairs p0, p0 + 1 = cells 8 g + 4 q4 .. + 3
          uint32_t u0 = 0, u1 = 0, u2 = 0, u3 = 0;
          if (p0 < NP) {
            const float2 sc = mul2(x[p0 < NP ? p0 : 0], scale);
            u0 = f2u_rni(sc.x);
            u1 = f2u_rni(sc.y);
          }
          if (p0 + 1 < NP) {
            const float2 sc = mul2(x[p0 + 1 < NP ? p0 + 1 : 0], scale);
            u2 = f2u_rni(sc.x);
            u3 = f2u_rni(sc.y);
          }
          // 4 x 4 byte transpose: slice i takes byte 3 - i of the four cells, cell 0 at the lowest address
          const uint32_t xa = prmt(u0, u1, 0x6273), ya = prmt(u0, u1, 0x4051);
          const uint32_t xb = prmt(u2, u3, 0x6273), yb = prmt(u2, u3, 0x4051);
          w[0][q4] = prmt(xa, xb, 0x5410);
          w[1][q4] = prmt(xa, xb, 0x7632);
          w[2][q4] = prmt(ya, yb, 0x5410);
          w[3][q4] = prmt(ya, yb, 0x7632);
        }
        const int k = quarter * KB + 8 * g;        // K index of the first of these 8 cells
        unsigned char* const dst = arow + (k >> 4) * A_CHUNK + (k & 15);
#pragma unroll
        for (int i = 0; i < kTcNA; ++i) *reinterpret_cast<uint2*>(dst + i * A_SLICE) = make_uint2(w[i][0], w[i][1]);
      }
      fence_proxy_async();     // generic-proxy stores -> visible to the tensor core (async proxy)
      // epilogue share of the previous item (its MMAs ran while this warp evaluated the cells above);
      // it must precede the arrival: arriving for item n licenses the issuer to overwrite accumulator
      // stage b, and item n + 1 to overwrite what item n - 1 occupies
      if (n >= 1) epilogue(n - 1, p_zi, p_tile, p_ablk, p_pg);
      __syncwarp();
      if (lane == 0) mbar_arrive1(&bar_full[b]);
      p_zi = zi;
      p_tile = tile;
      p_ablk = ablk;
      p_pg = pg;
      tc_next(Z.tc_nablk, Z.tc_nt, n_pg, zi, tile, ablk, pg);
    }
    epilogue(n_local - 1, p_zi, p_tile, p_ablk, p_pg);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kTcPts)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(2 * kTcStageCols) : "memory");
}

}  // namespace cpx

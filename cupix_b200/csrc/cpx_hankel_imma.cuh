// k_p3d_hankel_imma: the P3D -> Px Hankel contraction as an error-free integer split on the INT8 tensor path issued with
// mma.sync (IMMA.16832), accumulators in registers.
//
// Same mathematics and the same FP32 cell evaluation as k_p3d_hankel (theory.py:250-344 + the theta average of
// likelihood.py:99-130):   Px_Zam[pt][a][m] = amp(m) * sum_q W[a][q] * P3D'(m, q; pt).
// Why: on B200 the FP64 DMMA occupies the datapath the FP32 cell evaluation needs, so the two add
// (T = 16 N_DMMA + T_FP32).  mma.sync on s8/u8 runs on the tensor cores proper: 8 cycles per
// m16n8k32 per sub-partition, 11.5 beside saturating FFMA2 traffic from other warps, which it does not slow down
// (tools/imma_probe.cu, profiles/imma_probe_r02.txt).  Unlike the tcgen05 version of the same split (cpx_hankel_tc.cuh) this
// one keeps the warp-level structure of k_p3d_hankel -- cells generated in the lanes that feed them to the MMA, no
// shared-memory operand tiles, no tensor memory, no inter-warp hand-off -- which is what that version paid for.
//
//   * weights, constants of the batch:  W[a][q] = s_q t_a sum_j B_j[a][q] 256^-(j+1),  kImNW balanced base-256 digits (s8),
//     s_q = 2^round(log2 max_a |W[a][q]|) folded into the P_L constant of the cell (exact), t_a = 2^k_a (a power of two as
//     well) applied in the epilogue;
//   * a warp owns 16 k_par rows (rows g, g + 8 of the lane, as the packed FP32 pair) of one point at a time: lane (g, t)
//     evaluates the cells of nodes 4 ks + t, ks = 0 .. n_q4 - 1, keeps them in registers, takes the exact row maximum
//     (two shuffles), converts to 32-bit fixed point relative to it (U = rni(cell 2^(158 - e))) and transposes bytes:
//     the four bytes of four consecutive cells are the A-fragment registers of the four u8 slices.  K slot (block b,
//     half h, 4 t + j) of the MMA is node 4 (8 b + 4 h + j) + t: any permutation of K is allowed, the digit table uses
//     the same one;
//   * slice pairs (i, j) with i + j <= kImSmax go to accumulator i + j (INT32, exact: < 2^24 per accumulator); the
//     rest (< 2^-32 of the row scale, unbiased: balanced digits, round-to-nearest cells) is not issued;
//   * the epilogue combines the classes exactly in 64-bit integers and applies 2^(e - 166) t_a amp cont with one FP64
//     instruction per output.
// Error against the float64 dot product of the same FP32 cells: <= 4e-9 of the Px scale per theta bin with 4 x 4 slices
// and 10 pairs (tests/test_tc_split.py), a tenth of the float32 rounding of the cells themselves.
//
// CTA = 8 warps, 2 CTAs per SM, persistent over (z, 32-row tile, theta block, point group) items exactly like
// k_p3d_hankel; cell constants of the tile (P_L scaled by s_q) and the digit table of the theta block are staged with
// TMA bulk copies and stay resident.  Instantiated for 52, 56 and 88 nodes (n_q4 = 13, 14, 22), without the COSMO variant.
//
// Status: opt-in (CPX_IMMA=1).  Bit-for-bit the cells of k_p3d_hankel and chi2 within 1e-10 relative of it, but not
// faster: 50.7 against 49.3 ms per 131 072-point S5 step at 56 nodes, 73.1 against 73.7 at 88.  The contraction does
// leave the FP32 datapath (fmaheavy 47 % busy, tensor pipe 17 %), but conversion, byte transposes and the integer
// epilogue make it 16 % more instructions per row than the DMMA kernel, and with 4 warps per sub-partition (128 registers)
// the warp-serial phases leave issue at 54 % (profiles/k_p3d_hankel_imma_ncu_r02.txt, profiles/hankel_imma_r02.txt).
#pragma once
#include "cpx_kernels.cuh"

namespace cpx {

constexpr int kImNA = 4;        // cell bytes (u8 slices)
constexpr int kImNW = 4;        // weight digits (s8 slices)
constexpr int kImSmax = 3;      // highest slice-pair class i + j issued
constexpr int kImCls = kImSmax + 1;

__host__ __device__ constexpr int imma_dig_bytes(int nkb, int nt) { return nkb * kImNW * nt * 32 * 8; }

__device__ __forceinline__ void imma_u8s8(int (&c)[4], const uint32_t (&a)[4], uint2 b) {
  asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b.x), "r"(b.y));
}
__device__ __forceinline__ uint32_t im_f2u(float x) {
  uint32_t u;
  asm("cvt.rni.u32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return u;
}
__device__ __forceinline__ uint32_t im_prmt(uint32_t a, uint32_t b, uint32_t sel) {
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
  return d;
}
// exact int32 -> float64 on the ALU + FP64 pipes (I2F.F64 runs at the MUFU rate)
__device__ __forceinline__ double im_i2d(int v) {
  return __hiloint2double(0x43300000, v ^ 0x80000000) - 4503601774854144.0;   // 2^52 + 2^31
}

// pipe-rate probe: 8 independent m16n8k32 chains per warp
__global__ void k_probe_imma(int* out, int iters, uint32_t seed) {
  int c[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0;
  const uint32_t a[4] = {seed, seed * 3u, seed * 5u, seed * 7u};
  const uint2 b = make_uint2(seed * 11u, seed * 13u);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) imma_u8s8(c[i], a, b);
  }
  int s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  if (s == 123457) out[0] = s;
}

// NKS = n_q4 = ceil(n_q / 4): the lane's cells per row, a compile-time constant so that the cell loop carries neither
// per-slot branches nor address arithmetic (a runtime count cost 18 % of the kernel in branch, move and LEA instructions)
template <int NKS, int NT, bool ADD, bool COSMO>
__global__ void __launch_bounds__(kK1Threads, 2)
    k_p3d_hankel_imma(const ZDev* __restrict__ ztab, int n_z, int count, long long n_items) {
  constexpr int NKB = (NKS + 7) / 8;                   // 32-slot K blocks of the MMA
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ __align__(16) ZDev s_Z;

  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int n_pg = (count + kK1Warps - 1) / kK1Warps;

  const long long it0 = n_items * blockIdx.x / gridDim.x;
  const long long it1 = n_items * (blockIdx.x + 1) / gridDim.x;
  if (it0 >= it1) return;

  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  uint32_t phase = 0;
  int cur_z = -1, cur_tile = -1, cur_ablk = -1;

  int zi, tile, ablk, pg;
  {
    zi = 0;
    while (zi + 1 < n_z && ztab[zi + 1].item_start <= it0) ++zi;
    long long r = it0 - ztab[zi].item_start;
    pg = static_cast<int>(r % n_pg);
    r /= n_pg;
    ablk = static_cast<int>(r % ztab[zi].n_ablk);
    tile = static_cast<int>(r / ztab[zi].n_ablk);
  }

  for (long long it = it0; it < it1; ++it) {
    if (zi != cur_z) {
      __syncthreads();
      const int* src = reinterpret_cast<const int*>(&ztab[zi]);
      int* dst = reinterpret_cast<int*>(&s_Z);
      for (int i = threadIdx.x; i < static_cast<int>(sizeof(ZDev) / 4); i += blockDim.x) dst[i] = src[i];
      __syncthreads();
    }
    const ZDev& Z = s_Z;
    constexpr uint32_t cell_bytes = static_cast<uint32_t>(NKS) * (kCellConsts * kTileRows * 4 * 4u);      // Z.n_q4 == NKS (dispatch)
    constexpr uint32_t dig_bytes = imma_dig_bytes(NKB, NT);
    constexpr uint32_t ta_bytes = 8 * NT * 4u;
    const float4* s_cell = reinterpret_cast<const float4*>(smem_raw);
    const uint2* s_dig = reinterpret_cast<const uint2*>(smem_raw + cell_bytes);
    const int* s_tae = reinterpret_cast<const int*>(smem_raw + cell_bytes + dig_bytes);      // k_a: t_a = 2^k_a

    const bool new_tile = (zi != cur_z) || (tile != cur_tile);
    const bool new_w = (zi != cur_z) || (ablk != cur_ablk);
    if (new_tile || new_w) {
      __syncthreads();   // every warp is done reading the previous tile / digits
      if (threadIdx.x == 0) {
        fence_proxy_async();
        mbar_expect_tx(&bar, (new_tile ? cell_bytes : 0u) + (new_w ? dig_bytes + ta_bytes : 0u));
        if (new_tile) {
          const unsigned char* src = reinterpret_cast<const unsigned char*>(Z.im_cell) + static_cast<size_t>(tile) * cell_bytes;
          for (uint32_t off = 0; off < cell_bytes; off += 32768u)
            bulk_g2s(smem_raw + off, src + off, min(32768u, cell_bytes - off), &bar);
        }
        if (new_w) {
          bulk_g2s(smem_raw + cell_bytes, reinterpret_cast<const unsigned char*>(Z.im_wdig) + static_cast<size_t>(ablk) * dig_bytes, dig_bytes, &bar);
          bulk_g2s(smem_raw + cell_bytes + dig_bytes, reinterpret_cast<const unsigned char*>(Z.im_ta) + static_cast<size_t>(ablk) * ta_bytes,
                   ta_bytes, &bar);
        }
      }
      mbar_wait(&bar, phase);
      phase ^= 1u;
      cur_z = zi;
      cur_tile = tile;
      cur_ablk = ablk;
    }

    const int pt = pg * kK1Warps + warp;           // this warp's point
    if (pt >= count) {
      next_item(Z, n_pg, zi, tile, ablk, pg);
      continue;
    }
    const int m_base = tile * kTileRows;
    const bool rowwise = (Z.flags & 9) != 0;       // HCD or continuum: per-row scalars from k_rowpars
    const PtPar& pr = Z.ptpar[pt];
    const RowPar* const rp = Z.rowpar + static_cast<size_t>(pt) * Z.n_m_pad + m_base + g;

    // warm L1 / L2 with the parameters of this warp's next item (same tile, next point group)
    if (pg + 1 < n_pg) {
      const int ptn = min(pt + kK1Warps, count - 1);
      if (lane == 0) prefetch_l1(&Z.ptpar[ptn]);
      if (rowwise && lane < 8) prefetch_l1(&Z.rowpar[static_cast<size_t>(ptn) * Z.n_m_pad + m_base + 4 * lane]);
    }

    const int a_lane0 = ablk * (8 * NT) + 2 * t;
    const int n_a = Z.n_a, n_m = Z.n_m, n_m_pad = Z.n_m_pad;
    double* const outp = Z.pxzam + (static_cast<size_t>(pt) * n_a + a_lane0) * n_m_pad + m_base + g;

    // per-point scalars, read once per item (packed FP32 instructions take a scalar operand for both halves: one register
    // each).  Read inside the loop over halves they cost 5 % of the kernel in first-use stalls (ncu source view).
    const float p_q1l = pr.q1l, p_q1lo = pr.q1l_lo, p_q2l = pr.q2l, p_q2lo = pr.q2l_lo, p_av = pr.av, p_bv = pr.bv;
    const float p_nlkv = pr.nlkv, p_nlkvlo = pr.nlkv_lo, p_ikp2l = pr.ikp2l, p_ikp2lo = pr.ikp2l_lo;
    float p_dn = 0.0f, p_lr = 0.0f, p_lrlo = 0.0f;
    if (COSMO) {
      p_dn = pr.dn;
      p_lr = pr.lr;
      p_lrlo = pr.lr_lo;
    }
    float b0_hi = 0.0f, b0_lo = 0.0f;
    double amp0 = 0.0;
    if (!rowwise) {
      split_hi_lo(pr.beta, b0_hi, b0_lo);
      amp0 = pr.bias * pr.bias * Z.dAA;
    }

    // the two 16-row halves of the tile, one after the other: rows 16 h2 + g (.x) and 16 h2 + 8 + g (.y)
#pragma unroll 1
    for (int h2 = 0; h2 < 2; ++h2) {
      uint32_t A[NKB][kImNA][4];
      int ebx, eby;
      {
        const float2 q1l = dup2(p_q1l), q1lo = dup2(p_q1lo), q2l = dup2(p_q2l), q2lo = dup2(p_q2lo);
        const float2 av = dup2(p_av), bv = dup2(p_bv), nlkv = dup2(p_nlkv), nlkvlo = dup2(p_nlkvlo);
        const float2 ikp2l = dup2(p_ikp2l), ikp2lo = dup2(p_ikp2lo);
        const float2 cdn = dup2(p_dn), clr = dup2(p_lr), clrlo = dup2(p_lrlo);
        // Kaiser beta of the two rows
        float2 betaT = dup2(b0_hi), betaTlo = dup2(kKaiserLoScale * b0_lo);
        if (rowwise) {
          betaT = make_float2(rp[16 * h2].betaT_hi, rp[16 * h2 + 8].betaT_hi);
          betaTlo = make_float2(kKaiserLoScale * rp[16 * h2].betaT_lo, kKaiserLoScale * rp[16 * h2 + 8].betaT_lo);
        }

        // ---- the lane's cells of this half: nodes 4 ks + t, both rows packed ----
        float2 x[NKS];
        float mx = 0.0f, my = 0.0f;
        const float4* const c = s_cell + h2 * 32 + lane;       // [ks][3 constant pairs][half][lane]: three LDS.128 per cell pair
#pragma unroll
        for (int ks = 0; ks < NKS; ++ks) {
          const float4* const ck = c + ks * (3 * 2 * 32);
          const float4 v0 = ck[0 * 64], v1 = ck[1 * 64], v2 = ck[2 * 64];
          const float2 D0 = make_float2(v0.x, v0.y), L2K = make_float2(v0.z, v0.w), L2MU = make_float2(v1.x, v1.y);
          const float2 MU2 = make_float2(v1.z, v1.w), NK2 = make_float2(v2.x, v2.y), PL0 = make_float2(v2.z, v2.w);
          float2 D = D0, PL = PL0;
          if (COSMO) {   // P_L and Delta^2 of the rescaled cosmology (theory.py:54-56)
            const float2 f = exp2_rr2(add2(fma2(cdn, L2K, clr), clrlo));
            D = mul2(D0, f);
            PL = mul2(PL0, f);
          }
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
          x[ks] = y;
          mx = fmaxf(mx, y.x);
          my = fmaxf(my, y.y);
        }
        // exact row maxima over the four lanes that share the rows
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
        my = fmaxf(my, __shfl_xor_sync(0xffffffffu, my, 1));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
        my = fmaxf(my, __shfl_xor_sync(0xffffffffu, my, 2));
        // fixed point relative to the row maximum: rmax < 2^(e - 126)  =>  cell * 2^(158 - e) < 2^32
        ebx = static_cast<int>(min(max(__float_as_uint(mx) >> 23, 31u), 254u));
        eby = static_cast<int>(min(max(__float_as_uint(my) >> 23, 31u), 254u));
        const float2 scale = make_float2(__uint_as_float(static_cast<uint32_t>(285 - ebx) << 23),
                                         __uint_as_float(static_cast<uint32_t>(285 - eby) << 23));
        // ---- 32-bit fixed point, 4 x 4 byte transposes: slice i takes byte 3 - i of four consecutive cells ----
#pragma unroll
        for (int q4 = 0; q4 < 2 * NKB; ++q4) {
          uint32_t ux[4], uy[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            ux[j] = uy[j] = 0u;
            if (4 * q4 + j < NKS) {             // compile-time: the slots past the last node hold zeros
              const float2 sc = mul2(x[4 * q4 + j], scale);
              ux[j] = im_f2u(sc.x);
              uy[j] = im_f2u(sc.y);
            }
          }
          const int b = q4 >> 1, hh = q4 & 1;      // MMA K block and half: A registers 2 hh (row g) and 2 hh + 1 (row g + 8)
          {
            const uint32_t xa = im_prmt(ux[0], ux[1], 0x6273), ya = im_prmt(ux[0], ux[1], 0x4051);
            const uint32_t xb = im_prmt(ux[2], ux[3], 0x6273), yb = im_prmt(ux[2], ux[3], 0x4051);
            A[b][0][2 * hh] = im_prmt(xa, xb, 0x5410);
            A[b][1][2 * hh] = im_prmt(xa, xb, 0x7632);
            A[b][2][2 * hh] = im_prmt(ya, yb, 0x5410);
            A[b][3][2 * hh] = im_prmt(ya, yb, 0x7632);
          }
          {
            const uint32_t xa = im_prmt(uy[0], uy[1], 0x6273), ya = im_prmt(uy[0], uy[1], 0x4051);
            const uint32_t xb = im_prmt(uy[2], uy[3], 0x6273), yb = im_prmt(uy[2], uy[3], 0x4051);
            A[b][0][2 * hh + 1] = im_prmt(xa, xb, 0x5410);
            A[b][1][2 * hh + 1] = im_prmt(xa, xb, 0x7632);
            A[b][2][2 * hh + 1] = im_prmt(ya, yb, 0x5410);
            A[b][3][2 * hh + 1] = im_prmt(ya, yb, 0x7632);
          }
        }
      }

      int acc[NT][kImCls][4];
#pragma unroll
      for (int jn = 0; jn < NT; ++jn)
#pragma unroll
        for (int s = 0; s < kImCls; ++s) acc[jn][s][0] = acc[jn][s][1] = acc[jn][s][2] = acc[jn][s][3] = 0;
      // consecutive MMAs go to different theta tiles: an accumulator is touched again NT MMAs later at the earliest
#pragma unroll
      for (int b = 0; b < NKB; ++b)
#pragma unroll
        for (int d = 0; d < kImNW; ++d) {
          uint2 B[NT];
#pragma unroll
          for (int jn = 0; jn < NT; ++jn) B[jn] = s_dig[((b * kImNW + d) * NT + jn) * 32 + lane];
#pragma unroll
          for (int i = 0; i < kImNA; ++i)
            if (i + d <= kImSmax) {
#pragma unroll
              for (int jn = 0; jn < NT; ++jn) imma_u8s8(acc[jn][i + d], A[b][i], B[jn]);
            }
        }
      // ---- row factors of the epilogue: s = 2^(e - 166) amp cont; Px = (s 2^k_a H + additive) * cont, t_a = 2^k_a ----
      // FP64 instructions are starved by the FP32 traffic of the other warps (a DFMA takes 55 cycles
      // beside three FFMA warps), so the epilogue spends one per output: the classes are combined in 64-bit integers, the
      // integer becomes a double by the 2^52 trick (d = M + H, M = 2^52 + 2^51), and  Px = fma(d, s', -M s')  with the
      // power-of-two column scale added to the exponent fields of s and -M s (integer adds).
      const int m0 = m_base + 16 * h2 + g, m1 = m0 + 8;
      double acx = amp0, acy = amp0, cont_x = 1.0, cont_y = 1.0;
      if (rowwise) {
        acx = rp[16 * h2].ampcont;
        acy = rp[16 * h2 + 8].ampcont;
        if (ADD && Z.rowcont != nullptr) {
          cont_x = Z.rowcont[static_cast<size_t>(pt) * n_m_pad + m0];
          cont_y = Z.rowcont[static_cast<size_t>(pt) * n_m_pad + m1];
        }
      }
      constexpr int E0 = 158 - 8 * (kImNA - 2 - kImSmax);       // 166 for 4 bytes and classes 0..3
      constexpr double kMagic = 6755399441055744.0;             // 2^52 + 2^51
      // s = 2^(e - 166) amp cont: the power of two goes into the exponent field (amp cont is far from the ends of the range)
      const double sx = (m0 < n_m && acx != 0.0) ? __hiloint2double(__double2hiint(acx) + ((ebx - E0) << 20), __double2loint(acx)) : 0.0;
      const double sy = (m1 < n_m && acy != 0.0) ? __hiloint2double(__double2hiint(acy) + ((eby - E0) << 20), __double2loint(acy)) : 0.0;
      const double cx = -kMagic * sx, cy = -kMagic * sy;

      // a zero row factor (padded row, zero amplitude) must stay zero under the exponent add
      const int mask_x = (sx != 0.0) ? -1 : 0, mask_y = (sy != 0.0) ? -1 : 0;
      const int sxh = __double2hiint(sx), sxl = __double2loint(sx), syh = __double2hiint(sy), syl = __double2loint(sy);
      const int cxh = __double2hiint(cx), cxl = __double2loint(cx), cyh = __double2hiint(cy), cyl = __double2loint(cy);
      double* const prow = outp + 16 * h2;
      // C fragment: (row g, cols 2t, 2t + 1), (row g + 8, cols 2t, 2t + 1)
#pragma unroll
      for (int jn = 0; jn < NT; ++jn)
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int cc = r & 1, yrow = r >> 1;
          const int a = a_lane0 + 8 * jn + cc;
          long long H;
          if (kImCls == 4) {       // two 32-bit halves, then one wide multiply-add (exact: |H| < 2^48)
            const int hi32 = acc[jn][0][r] * 256 + acc[jn][1][r];
            const long long lo = (NKB <= 2) ? static_cast<long long>(acc[jn][2][r] * 256 + acc[jn][3][r])
                                            : static_cast<long long>(acc[jn][2][r]) * 256 + acc[jn][3][r];
            H = static_cast<long long>(hi32) * 65536 + lo;
          } else {
            H = acc[jn][0][r];
#pragma unroll
            for (int s = 1; s < kImCls; ++s) H = H * 256 + acc[jn][s][r];
          }
          const double dH = __hiloint2double(static_cast<int>(H >> 32) + 0x43380000, static_cast<int>(H & 0xffffffffLL));   // M + H
          const int kadd = (s_tae[8 * jn + 2 * t + cc] << 20) & (yrow ? mask_y : mask_x);
          double px = fma(dH, __hiloint2double((yrow ? syh : sxh) + kadd, yrow ? syl : sxl),
                          __hiloint2double((yrow ? cyh : cxh) + kadd, yrow ? cyl : cxl));
          const int m = yrow ? m1 : m0;
          if (ADD && m < n_m && a < n_a) {
            const double cont = yrow ? cont_y : cont_x;
            double add = 0.0;
            if (Z.flags & 2) {   // SiIII auto + cross from precomputed Hankel moments, theory.py:361-438
              const size_t o = static_cast<size_t>(a) * n_m + m, st = static_cast<size_t>(n_a) * n_m;
              const double bx = pr.b_X, btx = pr.beta_X;
              if (COSMO && pr.dn_64 != 0.0) {
                const double dn = pr.dn_64;
                const int ord = Z.metal_dn_order;
                add += bx * bx * (metal_moment(Z.metal_auto, Z.metal_auto_dn, ord, o, st, 0, dn) +
                                  2.0 * btx * metal_moment(Z.metal_auto, Z.metal_auto_dn, ord, o, st, 1, dn) +
                                  btx * btx * metal_moment(Z.metal_auto, Z.metal_auto_dn, ord, o, st, 2, dn));
                add += pr.bias * bx * (metal_moment(Z.metal_cross, Z.metal_cross_dn, ord, o, st, 0, dn) +
                                       (pr.beta + btx) * metal_moment(Z.metal_cross, Z.metal_cross_dn, ord, o, st, 1, dn) +
                                       pr.beta * btx * metal_moment(Z.metal_cross, Z.metal_cross_dn, ord, o, st, 2, dn));
              } else {
                add += bx * bx * (Z.metal_auto[o] + 2.0 * btx * Z.metal_auto[o + st] + btx * btx * Z.metal_auto[o + 2 * st]);
                add += pr.bias * bx * (Z.metal_cross[o] + (pr.beta + btx) * Z.metal_cross[o + st] + pr.beta * btx * Z.metal_cross[o + 2 * st]);
              }
              if (COSMO) add *= pr.lin_amp;
            }
            if (Z.flags & 4) add += pr.b_noise * Z.sky_smooth[m] * Z.sky_xi[a];    // theory.py:476-506
            px = fma(add, cont, px);
          }
          if (a < n_a) prow[static_cast<size_t>(8 * jn + cc) * n_m_pad + 8 * yrow] = px;
        }
    }
    next_item(Z, n_pg, zi, tile, ablk, pg);
  }
}

}  // namespace cpx

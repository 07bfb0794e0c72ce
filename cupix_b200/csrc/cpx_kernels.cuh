// cupix_b200 device kernels (sm_100a).  See DESIGN.md section 3 for the data layout and section 4
// for the roofline of each kernel.
//
//   k_params      per-(point, z) parameter expansion: defaults (+) overrides, derived scalars as
//                 float32 hi + lo pairs (model_lya.py:190-196, model_contaminants.py:103-144; grid
//                 generation of scripts/chi2_scan_mock.py:107-111, bit-identical to numpy.linspace)
//   k_rowpars     per-(point, k_par row) float64 scalars of the HCD bias mixing and the continuum
//                 distortion (theory.py:267-282, 525-535)
//   k_p3d_hankel  fused Arinyo P3D evaluation on the shared-memory-staged (k_par, k_perp) grid
//                 (packed FP32) + J0 Hankel quadrature as an FP64 tensor-core (DMMA) contraction +
//                 contaminant epilogues (theory.py:95-123, 250-344; the theta average of
//                 likelihood.py:99-130 is inside the weights)
//   k_p3d_hankel_f64  the same with float64 cells (CPX_PRECISE_CELLS, reference-precision mode)
//   k_window      window convolution + theta rebinning (+ Cholesky whitening and chi2) as a batched
//                 FP64 tensor-core (DMMA) contraction (likelihood.py:53-88, 170-194)
//   k_window_tc   (cpx_window_tc.cuh) the chi2 flavour of k_window on tcgen05: exact INT8 split of Px_Zam and of
//                 the whitened window operator, INT32 accumulators in tensor memory, HBM-bound; the default
//   k_p3d_hankel_imma  (cpx_hankel_imma.cuh) the same contraction as an exact integer split on mma.sync INT8 (opt-in,
//                 CPX_IMMA=1: level with the DMMA kernel)
//   k_p3d_hankel_tc  (cpx_hankel_tc.cuh) the Hankel contraction of k_p3d_hankel on tcgen05 with the same split;
//                 opt-in (CPX_TC=1), currently slower than the DMMA kernel
//   k_reduce      chi2[pt] = sum_{z,A} chi2_zA in fixed order (likelihood.py:191)
//   k_mock_noise  many noisy realisations of the data vector drawn on the device (likelihood.py:37-50), Philox4x32-10
//   k_jbar        set-up: J0 sums of the Hankel weights (cupix_b200/quadrature.py)
//   k_hankel_apply  the Hankel contraction alone, for caller-evaluated P3D cells (theory.py:250-264)
//   (cpx_factor.cuh) k_basis_f64, k_window_gram, k_gram, k_quad: the factorised path for scans and amplitude fits;
//                 k_p3d_hankel_f64s: the staged float64-cell Hankel kernel
//   k_probe_*     pipe-rate probes for the roofline denominators
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace cpx {

constexpr int kTileRows = 32;        // k_par rows per shared-memory tile (one per lane)
constexpr int kCellConsts = 6;       // D, L2K, L2MU, MU2, -K2, PL
constexpr int kK1Warps = 8;
constexpr int kK1Threads = kK1Warps * 32;
constexpr int kMaxZ = 64;

// per-(z, point) parameters written by k_params and read by k_p3d_hankel
struct PtPar {
  // cell evaluation (FP32 pipe).  Scalars whose float32 rounding would shift every cell the same
  // way (and so survive the averaging over cells) are carried as hi + lo pairs.
  float q1l, q1l_lo;   // q1 * log2(e)
  float q2l, q2l_lo;   // q2 * log2(e)
  float av, av_lo;
  float bv;
  float nlkv, nlkv_lo; // -av * log2(kv_Mpc)
  float ikp2l, ikp2l_lo;   // log2(e) / kp_Mpc^2
  float pad;
  // epilogues (FP64)
  double bias, beta;
  double b_H, beta_H, L_H;
  double b_X, beta_X;
  double b_noise;
  double kC, pC;
  // float64 copies for the reference-precision cell evaluation (CPX_PRECISE_CELLS)
  double q1_64, q2_64, av_64, bv_64, l2kv_64, ikp2_64;
  // linear-power rescaling f(k) = 2^(dn log2 k + lr) for CPX_AS / CPX_NS columns (appended: k_p3d_hankel_tc
  // stages the first 64 bytes of this struct); dn = ns - ns0, lr = log2(As / As0) - dn log2(k_pivot)
  float dn, lr, lr_lo, pad2;
  double lin_amp;             // As / As0: factor of the SiIII tables (linear in P_L; the tilt is refused with metals)
  double dn_64, lr_64;
  double pad3;                // keeps the size a multiple of 16: k_p3d_hankel_tc fetches the head with 16-byte cp.async
};
static_assert(sizeof(PtPar) == 224 && sizeof(PtPar) % 16 == 0, "PtPar layout");

// per-z device data for k_p3d_hankel / k_params
struct RowPar;
struct ZDev {
  int n_a, n_m, n_q, n_tiles;          // n_tiles = ceil(n_m / 32)
  int n_m_pad;                          // 32 * n_tiles
  int na_blk, n_ablk;                   // theta-block width (8 * NT) and count
  int flags;
  int n_q4;                             // ceil(n_q / 4): k-steps of the DMMA contraction
  int n_mblk;                           // coarse-k column blocks of k_window (n_M_pad = n_mblk * 8 * MT)
  double dAA;
  const float* cellc;                   // [n_tiles][n_q4][6][2][32][2] A-fragment order: lane = 4*g + t holds rows 8*(2*pr+e)+g, e = 0,1 (packed pair), node 4*ks+t
  const double* Wd;                     // [n_ablk][n_q4][NT][32]     B-fragment order: lane = 4*g + t holds W[theta 8*j+g][node 4*ks+t]
  const double* cell64;                 // [n_tiles][n_q4][6][4][32]  float64 cell constants (+K2, not -K2) for CPX_PRECISE_CELLS
  const double* kpar;                   // [n_m_pad]  k_par of the P3D models (cells, HCD)
  const double* kpar_row;               // [n_m_pad]  k_par of the continuum distortion (differs only in kpar0 = 'reference' mode)
  const double* metal_auto;             // [3][n_a][n_m] or null
  const double* metal_cross;            // [3][n_a][n_m] or null
  const double* metal_auto_dn;          // [metal_dn_order + 1][3][n_a][n_m] Taylor terms of metal_auto in the tilt dn, or null
  const double* metal_cross_dn;         // same for metal_cross
  int metal_dn_order, pad_dn;
  const double* sky_xi;                 // [n_a] or null
  const double* sky_smooth;             // [n_m] or null
  double defaults[18];
  double k_pivot;                       // pivot of the primordial power law (CPX_AS / CPX_NS), 0 = not available
  // window part
  int n_A, n_M, n_M_pad, n_pairs, n_data;
  int pad0;
  const int* pair_start;                // [n_A + 1] pairs of coarse bin A are pair_start[A] .. pair_start[A+1]
  const int* pair_a;                    // [n_pairs] fine bin of each pair
  const double* T_chi;                  // [n_pairs][n_M_pad][n_m_pad]  L^-1 diag(V/Vsum) U
  const double* T_model;                // [n_pairs][n_M_pad][n_m_pad]  diag(V/Vsum) U
  const double* yd;                     // [n_data][n_A][n_M_pad]       L^-1 d
  const double* ydd;                    // [n_data]                     |L^-1 d|^2 summed over (A, M): constant term of the factorised chi2
  // per-chunk buffers
  PtPar* ptpar;                         // [chunk]
  struct RowPar* rowpar;                // [chunk][n_m_pad], valid when flags & (HCD | CONTINUUM)
  double* rowcont;                      // [chunk][n_m_pad] continuum factor alone, or null (only with CONTINUUM and additive terms)
  double* pxzam;                        // [chunk][n_a][n_m_pad]
  long long item_start;                 // first k_p3d_hankel work item of this z in the launch
  // tensor-core (tcgen05) Hankel path, cpx_hankel_tc.cuh; tc_np == 0: not available for this z (n_q > 96)
  int tc_np;                            // packed cell pairs per lane (n_q <= 8 * tc_np)
  int tc_nt;                            // 8-row k_par tiles (n_m_pad / 8)
  int tc_nablk;                         // theta blocks of 32
  int pad1;
  const float* tc_cell;                 // [tc_nt][6][quarter 4][tc_np | 1][row 8][2] cell constants, P_L scaled by s_q
  const signed char* tc_wdig;           // [tc_nablk][160 x KP] balanced base-256 digits of W / (s_q t_a), K-major core-matrix order
  const double* tc_ta;                  // [tc_nablk * 32] column scales t_a
  // tensor-core window path, cpx_window_tc.cuh; wt_nb == 0: not available (more than 64 coarse-k bins)
  int wt_nb;                            // coarse-k columns rounded up to 16
  int pad2;
  const signed char* wt_dig;            // [chunk][5 * wt_nb x 32] digits of T_chi / (2^c rs), chunks ordered (A, pair in A, m / 32)
  const int* wt_off;                    // [n_A] first chunk of coarse bin A
  const int* wt_cs;                     // [n_pairs][n_m_pad / 16] exponent c of the column scale
  const double* wt_rs;                  // [n_A][wt_nb] row scales
  const void* px_tmap;                  // CUtensorMap of pxzam viewed as [chunk][n_a][n_m_pad] float64, box 128 x 1 x 16, 128-byte swizzle
  // INT8 mma.sync Hankel path, cpx_hankel_imma.cuh; im_cell == null: not available (n_q > 96)
  const float* im_cell;                 // [n_tiles][n_q4][3][2][32][4] cellc with P_L scaled by s_q, constants interleaved in pairs (LDS.128)
  const unsigned char* im_wdig;         // [n_ablk][NKB][4 digits][NT][32 lanes][8] balanced base-256 digits of W / (s_q t_a), B-fragment order
  const int* im_ta;                     // [n_ablk * na_blk] exponents k_a of the column scales t_a = 2^k_a
};

// ------------------------------------------------------------------------------------------
// small PTX helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// TMA 1-D bulk copy global -> shared, completion on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
// FP64 tensor-core MMA, D(8x8) = A(8x4) * B(4x8) + C
__device__ __forceinline__ void dmma_884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

// ------------------------------------------------------------------------------------------
// k_params
// ------------------------------------------------------------------------------------------
struct ParamsArgs {
  const double* points;   // [B][P] or null (grid mode)
  int P;
  int n_z;                // number of z bins evaluated
  long long first;        // chunk offset into points / grid
  int count;              // points in this chunk
  int grid_mode;
  double grid_lo[16], grid_hi[16], grid_step[16];
  int grid_n[16];
  int slot[16], zsel[16];
  long long grid_first;
};

__device__ __forceinline__ void split_hi_lo(double x, float& hi, float& lo) {
  hi = static_cast<float>(x);
  lo = static_cast<float>(x - static_cast<double>(hi));
}

__global__ void k_params(ParamsArgs a, const ZDev* __restrict__ ztab, const int* __restrict__ zlist) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int zi = blockIdx.y;
  if (i >= a.count) return;
  const ZDev& Z = ztab[zi];
  const int zglobal = zlist[zi];
  double v[18];
#pragma unroll
  for (int s = 0; s < 18; ++s) v[s] = Z.defaults[s];
  if (a.grid_mode) {
    long long idx = a.grid_first + a.first + i;
    for (int j = a.P - 1; j >= 0; --j) {
      const int n = a.grid_n[j];
      const int ij = static_cast<int>(idx % n);
      idx /= n;
      // bit-identical to numpy.linspace: arange * step + start without FMA contraction, last = stop
      const double x = (ij == n - 1 && n > 1) ? a.grid_hi[j]
                                              : __dadd_rn(__dmul_rn(static_cast<double>(ij), a.grid_step[j]), a.grid_lo[j]);
      if (a.zsel[j] < 0 || a.zsel[j] == zglobal) v[a.slot[j]] = x;
    }
  } else {
    const double* p = a.points + (a.first + i) * a.P;
    for (int j = 0; j < a.P; ++j)
      if (a.zsel[j] < 0 || a.zsel[j] == zglobal) v[a.slot[j]] = p[j];
  }
  PtPar o;
  const double l2e = 1.4426950408889634074;
  split_hi_lo(v[2] * l2e, o.q1l, o.q1l_lo);
  split_hi_lo(v[3] * l2e, o.q2l, o.q2l_lo);
  split_hi_lo(v[5], o.av, o.av_lo);
  o.bv = static_cast<float>(v[6]);
  // av itself stays a single float: its rounding moves chi2 by < 2e-5; the
  // offset -av*log2(kv) is formed from the rounded av so that only that residual is left
  split_hi_lo(-v[5] * log2(v[4]), o.nlkv, o.nlkv_lo);
  split_hi_lo(l2e / (v[7] * v[7]), o.ikp2l, o.ikp2l_lo);
  o.pad = 0.f;
  o.q1_64 = v[2];
  o.q2_64 = v[3];
  o.av_64 = v[5];
  o.bv_64 = v[6];
  o.l2kv_64 = log2(v[4]);
  o.ikp2_64 = 1.0 / (v[7] * v[7]);
  o.bias = v[0];
  o.beta = v[1];
  o.b_H = v[8];
  o.beta_H = v[9];
  o.L_H = v[10];
  o.b_X = v[11];
  o.beta_X = v[12];
  o.b_noise = v[13];
  o.kC = v[14];
  o.pC = v[15];
  // theory.py:54-56: same background, primordial power rescaled relative to the plan's cosmology
  double dn = 0.0, ratio = 1.0;
  if (Z.k_pivot > 0.0) {
    dn = v[17] - Z.defaults[17];
    ratio = v[16] / Z.defaults[16];
  }
  const double lr = log2(ratio) - dn * log2(Z.k_pivot > 0.0 ? Z.k_pivot : 1.0);
  o.dn = static_cast<float>(dn);
  split_hi_lo(lr, o.lr, o.lr_lo);
  o.pad2 = 0.f;
  o.lin_amp = ratio;
  o.dn_64 = dn;
  o.lr_64 = lr;
  Z.ptpar[i] = o;
}

// ------------------------------------------------------------------------------------------
// k_rowpars: per-(point, k_par row) scalars that need float64 transcendentals -- the HCD
// scale-dependent bias/beta (theory.py:267-282) and the continuum distortion (theory.py:525-535).
// Computed once here so that the fused kernel's prologue/epilogue stay free of FP64 exp/pow/tanh.
// Only launched when include_hcd or include_continuum is set; otherwise the values are uniform.
// ------------------------------------------------------------------------------------------
struct RowPar {
  float betaT_hi, betaT_lo;   // Kaiser beta of the row (hi + lo float pair)
  double ampcont;             // b_T(k_par)^2 * dAA_dMpc * tanh((k_par / kC)^pC)  (the last factor is 1 without continuum distortion)
};
static_assert(sizeof(RowPar) == 16, "RowPar layout");
// The continuum factor on its own (ZDev::rowcont) is needed only by the additive terms (SiIII, sky), which it multiplies too:
// it is written when the z bin has both.  16 instead of 24 bytes per (point, z, row): k_rowpars is bound by its writes.

__global__ void k_rowpars(const ZDev* __restrict__ ztab, int count) {
  const ZDev& Z = ztab[blockIdx.y];
  if ((Z.flags & 9) == 0) return;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;   // flat (point, row)
  if (i >= static_cast<long long>(count) * Z.n_m_pad) return;
  const int pt = static_cast<int>(i / Z.n_m_pad);
  const int m = static_cast<int>(i % Z.n_m_pad);
  const PtPar& pr = Z.ptpar[pt];
  const double kpar = Z.kpar[m];
  double bT = pr.bias, betaT = pr.beta;
  if (Z.flags & 1) {
    const double F = exp(-kpar * pr.L_H);
    bT = pr.bias + pr.b_H * F;
    betaT = (pr.bias * pr.beta + pr.b_H * pr.beta_H * F) / bT;
  }
  RowPar o;
  split_hi_lo(betaT, o.betaT_hi, o.betaT_lo);
  const double cont = (Z.flags & 8) ? tanh(pow(Z.kpar_row[m] / pr.kC, pr.pC)) : 1.0;
  o.ampcont = bT * bT * Z.dAA * cont;
  Z.rowpar[static_cast<size_t>(pt) * Z.n_m_pad + m] = o;
  if (Z.rowcont != nullptr) Z.rowcont[static_cast<size_t>(pt) * Z.n_m_pad + m] = cont;
}

// ------------------------------------------------------------------------------------------
// k_p3d_hankel
// ------------------------------------------------------------------------------------------
// Px_Zam[pt][a][m] = amp(m) * sum_q W[a][q] * P3D'(m, q; pt)   as an FP64 tensor-core contraction.
//
// The Hankel sum is a GEMM with M = (point, k_par row), K = k_perp node, N = theta bin whose A
// operand is *generated* in registers: lane (g, t) of a warp evaluates the Arinyo P3D cell
// (row 8*mt + g, node 4*ks + t), which is exactly the A-fragment element the m8n8k4 DMMA wants
// from that lane, so no shuffle or shared-memory round trip is needed.  B fragments are the Hankel
// weights (conflict-free LDS.64), accumulators stay in registers for the whole k_perp loop.
//
// One CTA = 8 warps.  A warp owns a 32-row k_par tile (4 m-tiles) and PP parameter points at a
// time; a CTA pass covers 8*PP points.  The cell constants of the tile ([n_q/4][6][4][32] floats,
// fragment order) and the weights of the theta block ([n_q/4][NT][32] doubles) are staged in shared
// memory by TMA bulk copies and stay resident while the CTA walks its contiguous range of work
// items (z, tile, theta-block, point-group).
//
// Per pair of cells and point: 27 packed FMA-pipe instructions (hi+lo scalars, degree-6 polynomial exp2 for
// D_NL, range-reduced MUFU.EX2 for the velocity term, Kaiser factor), ~14 ALU, 2 MUFU, 2 NT DMMAs.
// Bound: FP64 tensor pipe + FP32 FMA pipe, which share issue.

// Packed FP32 (FFMA2 / FADD2 / FMUL2, new on sm_100): two P3D cells per instruction.
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 add2(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 dup2(float a) { return make_float2(a, a); }

// Range reduction shared by the two exp2 flavours: x = n + f, n = round(x) (returned in the low
// mantissa bits of t), |f| <= 0.5 exactly.
// CPX_EXP_REDUCE 0 (default): three packed FMA-pipe instructions.  1: the rounded value comes from FRND on the
// conversion pipe and one packed instruction goes away (same t, same f, bit for bit) -- measured slower
// (37.25 against 36.91 ms per 65 536-point S5 step, profiles/hankel_variants_r01.txt): FRND is not free here.
#ifndef CPX_EXP_REDUCE
#define CPX_EXP_REDUCE 0
#endif
template <int MAGIC_EXTRA = 0>
__device__ __forceinline__ void exp2_reduce(float2 x, float2& t, float2& f) {
  constexpr float magic = 12582912.0f + MAGIC_EXTRA;      // 1.5 * 2^23 (+ 256: see exp2_rr2_neg)
  t = add2(x, dup2(magic));
#if CPX_EXP_REDUCE == 0
  f = fma2(add2(t, dup2(-magic)), dup2(-1.0f), x);
#else
  f = fma2(make_float2(rintf(x.x), rintf(x.y)), dup2(-1.0f), x);
#endif
}
__device__ __forceinline__ float exp2_scale(float p, float t) {
  // funnel shift + add keep this on the ALU pipe; a plain (t << 23) + p becomes an IMAD on the FMA
  // pipe, which is the one the FP64 tensor pipe competes with (profiles/dmma_contention_r01.txt)
  return __int_as_float(__float_as_int(p) + static_cast<int>(__funnelshift_l(0u, static_cast<unsigned>(__float_as_int(t)), 23)));
}
// 2^x, unbiased: 2^f is a degree-6 polynomial on the FMA pipe.  MUFU.EX2 alone is biased by
// -2e-8 .. -5e-8 on this range (tools/mufu_error.cu, profiles/mufu_error_r01.txt), which is coherent
// over all cells and would shift chi2 by ~5e-7 relative.  Degree 6 (minimax fit of the relative error on
// [-0.5, 0.5], coefficients rounded to float32 one at a time with the others refitted) approximates 2^f to
// 3.0e-9 with a mean of -2e-12; evaluated in float32 its error statistics (bias < 3e-10, rms 2.6e-8) are those
// of the degree-7 polynomial used before: the rounding of the last two Horner steps dominates either way.
__device__ __forceinline__ float2 exp2_poly2(float2 x) {
  x.x = fminf(fmaxf(x.x, -125.0f), 126.0f);
  x.y = fminf(fmaxf(x.y, -125.0f), 126.0f);
  float2 t, f;
  exp2_reduce<0>(x, t, f);
  float2 p = dup2(0.00015326094580814242f);
  p = fma2(p, f, dup2(0.001339080510661006f));
  p = fma2(p, f, dup2(0.0096185067668557167f));
  p = fma2(p, f, dup2(0.055503599345684052f));
  p = fma2(p, f, dup2(0.24022647738456726f));
  p = fma2(p, f, dup2(0.69314718246459961f));
  p = fma2(p, f, dup2(1.0f));
  return make_float2(exp2_scale(p.x, t.x), exp2_scale(p.y, t.y));
}
// 2^x with MUFU.EX2 on the reduced argument only (round-to-nearest reduction, f in [-0.5, 0.5]:
// bias ~ -8e-9, enough for the velocity term, which enters the exponent of D_NL multiplied by
// Delta^2 (q1 + q2 Delta^2)).  x <= ~30 here.
__device__ __forceinline__ float2 exp2_rr2(float2 x) {
  x.x = fmaxf(x.x, -125.0f);
  x.y = fmaxf(x.y, -125.0f);
  float2 t, f;
  exp2_reduce<0>(x, t, f);
  return make_float2(exp2_scale(ex2_approx(f.x), t.x), exp2_scale(ex2_approx(f.y), t.y));
}
// -2^x for free: with 256 added to the rounding constant, bit 8 of the integer part is set, and the shift by 23
// that moves the integer part into the exponent field moves that bit into the sign.  The caller folds the
// negation into an FFMA2 instead of spending a packed multiply on it.
__device__ __forceinline__ float2 exp2_rr2_neg(float2 x) {
  x.x = fmaxf(x.x, -125.0f);
  x.y = fmaxf(x.y, -125.0f);
  float2 t, f;
  exp2_reduce<256>(x, t, f);
  return make_float2(exp2_scale(ex2_approx(f.x), t.x), exp2_scale(ex2_approx(f.y), t.y));
}
// 2^x with MUFU.EX2 on f = x - floor(x) in [0, 1): the bias of MUFU.EX2 depends on the sign and
// the integer part of its argument (profiles/mufu_bias_curve_r01.txt: -2e-8..-6e-8 for x < 0,
// +2e-8..+5e-8 for x > 1, but +5.6e-9 averaged over [0, 1)).  Six FMA-pipe instructions cheaper
// per cell pair than the polynomial, at the price of that 5.6e-9 bias.
__device__ __forceinline__ float2 exp2_fl2(float2 x) {
  x.x = fminf(fmaxf(x.x, -125.0f), 126.0f);
  x.y = fminf(fmaxf(x.y, -125.0f), 126.0f);
  const float2 t = add2(add2(x, dup2(-0.5f)), dup2(12582912.0f));        // integer part = floor(x) (ties: round)
  const float2 f = fma2(add2(t, dup2(-12582912.0f)), dup2(-1.0f), x);    // x - floor(x), exact
  return make_float2(exp2_scale(ex2_approx(f.x), t.x), exp2_scale(ex2_approx(f.y), t.y));
}

// SiIII Hankel moment j of theta bin / row `o` (stride `st` between moments) for a point whose primordial tilt differs by dn
// from the plan's cosmology: Horner in dn over the Taylor tables (plan.METAL_TILT_ORDER), the plain table when dn == 0
__device__ __forceinline__ double metal_moment(const double* __restrict__ tab0, const double* __restrict__ tab_dn, int order,
                                               size_t o, size_t st, int j, double dn) {
  if (dn == 0.0 || tab_dn == nullptr) return tab0[o + j * st];
  const double* t = tab_dn + j * st + o;
  double s = t[static_cast<size_t>(order) * 3 * st];
  for (int n = order - 1; n >= 0; --n) s = fma(s, dn, t[static_cast<size_t>(n) * 3 * st]);
  return s;
}

// advance (zi, tile, ablk, pg) to the next work item without divisions
__device__ __forceinline__ void next_item(const ZDev& Z, int n_pg, int& zi, int& tile, int& ablk, int& pg) {
  if (++pg == n_pg) {
    pg = 0;
    if (++ablk == Z.n_ablk) {
      ablk = 0;
      if (++tile == Z.n_tiles) {
        tile = 0;
        ++zi;
      }
    }
  }
}

// Kaiser factor (1 + beta mu^2)^2 of a cell: CPX_KAISER 0 = six packed instructions, 1 (default) = four
#ifndef CPX_KAISER
#define CPX_KAISER 1
#endif
constexpr float kKaiserLoScale = CPX_KAISER == 0 ? 2.0f : 1.0f;

#ifdef CPX_DNL_MUFU
#define kExpDnl exp2_fl2
#else
#define kExpDnl exp2_poly2
#endif

template <int PP, int NT, int MINB, bool ADD, bool COSMO = false>
__global__ void __launch_bounds__(kK1Threads, MINB)
    k_p3d_hankel(const ZDev* __restrict__ ztab, int n_z, int count, long long n_items) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ __align__(16) ZDev s_Z;

  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int n_pg = (count + kK1Warps * PP - 1) / (kK1Warps * PP);

  // contiguous slice of the flattened item space for this CTA
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

  // first item: decode once with divisions, then advance incrementally
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
      // keep this z bin's descriptor in shared memory: every later field access is an LDS, and the
      // per-item parameter loads no longer wait on a pointer fetched from global memory
      __syncthreads();
      const int* src = reinterpret_cast<const int*>(&ztab[zi]);
      int* dst = reinterpret_cast<int*>(&s_Z);
      for (int i = threadIdx.x; i < static_cast<int>(sizeof(ZDev) / 4); i += blockDim.x) dst[i] = src[i];
      __syncthreads();
    }
    const ZDev& Z = s_Z;
    const int n_ks = Z.n_q4;
    const uint32_t cell_bytes = static_cast<uint32_t>(n_ks) * (kCellConsts * kTileRows * 4 * 4u);
    const uint32_t w_bytes = static_cast<uint32_t>(n_ks) * (NT * 32 * 8u);
    const float2* s_cell = reinterpret_cast<const float2*>(smem_raw);
    const double* s_w = reinterpret_cast<const double*>(smem_raw + cell_bytes);

    const bool new_tile = (zi != cur_z) || (tile != cur_tile);
    const bool new_w = (zi != cur_z) || (ablk != cur_ablk);
    if (new_tile || new_w) {
      __syncthreads();   // every warp is done reading the previous tile / weights
      if (threadIdx.x == 0) {
        fence_proxy_async();
        mbar_expect_tx(&bar, (new_tile ? cell_bytes : 0u) + (new_w ? w_bytes : 0u));
        if (new_tile) {
          const unsigned char* src = reinterpret_cast<const unsigned char*>(Z.cellc) + static_cast<size_t>(tile) * cell_bytes;
          for (uint32_t off = 0; off < cell_bytes; off += 32768u)
            bulk_g2s(smem_raw + off, src + off, min(32768u, cell_bytes - off), &bar);
        }
        if (new_w) {
          const unsigned char* src = reinterpret_cast<const unsigned char*>(Z.Wd) + static_cast<size_t>(ablk) * w_bytes;
          for (uint32_t off = 0; off < w_bytes; off += 32768u)
            bulk_g2s(smem_raw + cell_bytes + off, src + off, min(32768u, w_bytes - off), &bar);
        }
      }
      mbar_wait(&bar, phase);
      phase ^= 1u;
      cur_z = zi;
      cur_tile = tile;
      cur_ablk = ablk;
    }

    const int pt0 = (pg * kK1Warps + warp) * PP;       // first point of this warp
    if (pt0 >= count) {
      next_item(Z, n_pg, zi, tile, ablk, pg);
      continue;
    }
    const int m_base = tile * kTileRows;
    const bool rowwise = (Z.flags & 9) != 0;   // HCD or continuum: per-row scalars from k_rowpars

    // ---- per-point scalars (warp-uniform, duplicated into both halves of a packed register) and
    //      the Kaiser beta of this lane's rows; m-tiles are processed in pairs (2*pr, 2*pr + 1) ----
    float2 q1l[PP], q1lo[PP], q2l[PP], q2lo[PP], av[PP], bv[PP], nlkv[PP], nlkvlo[PP], ikp2l[PP], ikp2lo[PP];
    float2 betaT[2][PP], betaTlo[2][PP];
    float2 cdn[PP], clr[PP], clrlo[PP];                 // COSMO: linear-power rescaling 2^(dn log2 k + lr)
#pragma unroll
    for (int p = 0; p < PP; ++p) {
      const PtPar& pr = Z.ptpar[min(pt0 + p, count - 1)];
      if (COSMO) {
        cdn[p] = dup2(pr.dn);
        clr[p] = dup2(pr.lr);
        clrlo[p] = dup2(pr.lr_lo);
      }
      q1l[p] = dup2(pr.q1l);
      q1lo[p] = dup2(pr.q1l_lo);
      q2l[p] = dup2(pr.q2l);
      q2lo[p] = dup2(pr.q2l_lo);
      av[p] = dup2(pr.av);
      bv[p] = dup2(pr.bv);
      nlkv[p] = dup2(pr.nlkv);
      nlkvlo[p] = dup2(pr.nlkv_lo);
      ikp2l[p] = dup2(pr.ikp2l);
      ikp2lo[p] = dup2(pr.ikp2l_lo);
#pragma unroll
      for (int mt = 0; mt < 4; ++mt) {
        float hi, lo;
        if (rowwise) {
          const RowPar& rp = Z.rowpar[static_cast<size_t>(min(pt0 + p, count - 1)) * Z.n_m_pad + m_base + 8 * mt + g];
          hi = rp.betaT_hi;
          lo = rp.betaT_lo;
        } else {
          split_hi_lo(pr.beta, hi, lo);
        }
        if (mt & 1) {
          betaT[mt >> 1][p].y = hi;
          betaTlo[mt >> 1][p].y = kKaiserLoScale * lo;
        } else {
          betaT[mt >> 1][p].x = hi;
          betaTlo[mt >> 1][p].x = kKaiserLoScale * lo;
        }
      }
    }

    // warm L1/L2 with the parameters of this warp's next item (same tile, next point group)
    if (pg + 1 < n_pg) {
      const int ptn = min(pt0 + kK1Warps * PP, count - 1);
      if (lane == 0) prefetch_l1(&Z.ptpar[ptn]);
      if (rowwise && lane < 8) prefetch_l1(&Z.rowpar[static_cast<size_t>(ptn) * Z.n_m_pad + m_base + 4 * lane]);
    }

    double acc[4][PP][NT][2];
#pragma unroll
    for (int mt = 0; mt < 4; ++mt)
#pragma unroll
      for (int p = 0; p < PP; ++p)
#pragma unroll
        for (int j = 0; j < NT; ++j) acc[mt][p][j][0] = acc[mt][p][j][1] = 0.0;

    // ---- k_perp loop: P3D cells (A fragments) -> DMMA against the Hankel weights (B fragments) ----
    const float2* c = s_cell + lane;
    const double* wb = s_w + lane;
#pragma unroll 2
    for (int ks = 0; ks < n_ks; ++ks) {
      double bfrag[NT];
#pragma unroll
      for (int j = 0; j < NT; ++j) bfrag[j] = wb[j * 32];
#pragma unroll
      for (int pr2 = 0; pr2 < 2; ++pr2) {
        const float2 D0 = c[(0 * 2 + pr2) * 32], L2K = c[(1 * 2 + pr2) * 32], L2MU = c[(2 * 2 + pr2) * 32];
        const float2 MU2 = c[(3 * 2 + pr2) * 32], NK2 = c[(4 * 2 + pr2) * 32], PL0 = c[(5 * 2 + pr2) * 32];
#pragma unroll
        for (int p = 0; p < PP; ++p) {
          float2 D = D0, PL = PL0;
          if (COSMO) {   // P_L and Delta^2 of the rescaled cosmology (theory.py:54-56)
            const float2 f = exp2_rr2(add2(fma2(cdn[p], L2K, clr[p]), clrlo[p]));
            D = mul2(D0, f);
            PL = mul2(PL0, f);
          }
          // Delta^2 (q1 + q2 Delta^2) log2e
          const float2 nl = mul2(D, add2(fma2(q2l[p], D, q1l[p]), fma2(q2lo[p], D, q1lo[p])));
          // (k/kv)^av mu^bv
          const float2 nv = exp2_rr2_neg(add2(fma2(av[p], L2K, fma2(bv[p], L2MU, nlkv[p])), nlkvlo[p]));   // -(k/kv)^av mu^bv
          // D_NL = 2^(nl (1 - v) - (k/kp)^2 log2e)
          const float2 tt = fma2(NK2, ikp2lo[p], fma2(NK2, ikp2l[p], nl));
          const float2 e = kExpDnl(fma2(nl, nv, tt));
          // Kaiser factor applied to the running product (x -> x + x (beta mu^2), twice) so that its float32
          // rounding decorrelates from cell to cell; (1 + beta mu^2) rounded on its own takes the same error
          // for every cell with mu ~ 1 and would bias Px coherently by ~2e-8.
          float2 x = mul2(PL, e);
#if CPX_KAISER == 0
          x = fma2(mul2(x, betaT[pr2][p]), MU2, x);
          x = fma2(mul2(x, betaT[pr2][p]), MU2, fma2(mul2(x, betaTlo[pr2][p]), MU2, x));   // betaTlo holds 2 * lo
#else
          // beta mu^2 of the cell, correctly rounded from the hi + lo pair (its rounding differs from cell to cell),
          // then x -> x + x (beta mu^2) twice: four packed instructions instead of six
          const float2 bm = fma2(betaT[pr2][p], MU2, mul2(betaTlo[pr2][p], MU2));         // betaTlo holds lo
          x = fma2(x, bm, x);
          x = fma2(x, bm, x);
#endif
          const double pd0 = static_cast<double>(x.x), pd1 = static_cast<double>(x.y);
#pragma unroll
          for (int j = 0; j < NT; ++j) dmma_884(acc[2 * pr2][p][j][0], acc[2 * pr2][p][j][1], pd0, bfrag[j]);
#pragma unroll
          for (int j = 0; j < NT; ++j) dmma_884(acc[2 * pr2 + 1][p][j][0], acc[2 * pr2 + 1][p][j][1], pd1, bfrag[j]);
        }
      }
      c += kCellConsts * 2 * 32;
      wb += NT * 32;
    }

    // ---- epilogue: amplitude, continuum (row scalars), metals + sky (ADD); store Px_Zam[pt][a][m] ----
    // C fragment: row 8*mt + g, theta columns 8*j + 2*t + {0, 1}.  Px = (amp * acc + additive) * cont.
    const int a_lane0 = ablk * (8 * NT) + 2 * t;
    const int n_a = Z.n_a, n_m = Z.n_m, n_m_pad = Z.n_m_pad;
#pragma unroll
    for (int p = 0; p < PP; ++p) {
      const int pt = pt0 + p;
      if (pt >= count) break;
      const PtPar& pr = Z.ptpar[pt];
      const double amp0 = pr.bias * pr.bias * Z.dAA;
      double* outp = Z.pxzam + (static_cast<size_t>(pt) * n_a + a_lane0) * n_m_pad + m_base + g;
      const RowPar* rp = Z.rowpar + static_cast<size_t>(pt) * n_m_pad + m_base + g;
#pragma unroll
      for (int mt = 0; mt < 4; ++mt) {
        const int m = m_base + 8 * mt + g;
        const bool row_ok = m < n_m;
        double ampcont = amp0, cont = 1.0;
        if (rowwise) {
          ampcont = rp[8 * mt].ampcont;
          if (ADD && Z.rowcont != nullptr) cont = Z.rowcont[static_cast<size_t>(pt) * n_m_pad + m];
        }
        const double scale = row_ok ? ampcont : 0.0;
        double sky = 0.0;
        if (ADD && (Z.flags & 4) && row_ok) sky = pr.b_noise * Z.sky_smooth[m];    // theory.py:476-506
#pragma unroll
        for (int j = 0; j < NT; ++j)
#pragma unroll
          for (int cc = 0; cc < 2; ++cc) {
            const int a = a_lane0 + 8 * j + cc;
            if (a >= n_a) continue;
            double px = scale * acc[mt][p][j][cc];
            if (ADD && row_ok) {
              double add = 0.0;
              if (Z.flags & 2) {   // SiIII auto + cross from precomputed Hankel moments, theory.py:361-438
                const size_t o = static_cast<size_t>(a) * n_m + m, st = static_cast<size_t>(n_a) * n_m;
                const double bx = pr.b_X, btx = pr.beta_X;
                if (COSMO && pr.dn_64 != 0.0) {      // tilted primordial power: metal moments from their Taylor tables in dn
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
                  add += pr.bias * bx * (Z.metal_cross[o] + (pr.beta + btx) * Z.metal_cross[o + st] +
                                         pr.beta * btx * Z.metal_cross[o + 2 * st]);
                }
                if (COSMO) add *= pr.lin_amp;
              }
              if (Z.flags & 4) add += sky * Z.sky_xi[a];
              px = fma(add, cont, px);
            }
            outp[static_cast<size_t>(8 * j + cc) * n_m_pad + 8 * mt] = px;
          }
      }
    }
    next_item(Z, n_pg, zi, tile, ablk, pg);
  }
}

// ------------------------------------------------------------------------------------------
// k_p3d_hankel_f64: reference-precision variant (CPX_PRECISE_CELLS).  Same DMMA contraction and
// epilogue, but the P3D cell is evaluated in float64 straight from the formulas of
// theory.py:285-308 with float64 cell constants read from global memory.  One warp per
// (z, tile, theta-block, point); no shared-memory staging -- this is the parity mode, not the fast one.
// ------------------------------------------------------------------------------------------
template <int NT>
__global__ void __launch_bounds__(256) k_p3d_hankel_f64(const ZDev* __restrict__ ztab, int n_z, int count, long long n_items) {
  const int lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const long long wid = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  for (long long it = wid; it < n_items; it += nwarps) {
    int zi = 0;
    while (zi + 1 < n_z && ztab[zi + 1].item_start <= it) ++zi;
    const ZDev& Z = ztab[zi];
    long long r = it - Z.item_start;
    const int pt = static_cast<int>(r % count);
    r /= count;
    const int ablk = static_cast<int>(r % Z.n_ablk);
    const int tile = static_cast<int>(r / Z.n_ablk);
    const int m_base = tile * kTileRows;
    const bool rowwise = (Z.flags & 9) != 0;
    const PtPar& pr = Z.ptpar[pt];
    const double q1 = pr.q1_64, q2 = pr.q2_64, av = pr.av_64, bv = pr.bv_64, l2kv = pr.l2kv_64, ikp2 = pr.ikp2_64;
    const bool cosmo = pr.dn_64 != 0.0 || pr.lr_64 != 0.0;
    double betaT[4];
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) {
      betaT[mt] = pr.beta;
      if (Z.flags & 1) {
        const double F = exp(-Z.kpar[m_base + 8 * mt + g] * pr.L_H);
        betaT[mt] = (pr.bias * pr.beta + pr.b_H * pr.beta_H * F) / (pr.bias + pr.b_H * F);
      }
    }
    double acc[4][NT][2];
#pragma unroll
    for (int mt = 0; mt < 4; ++mt)
#pragma unroll
      for (int j = 0; j < NT; ++j) acc[mt][j][0] = acc[mt][j][1] = 0.0;
    const double* c = Z.cell64 + static_cast<size_t>(tile) * Z.n_q4 * (kCellConsts * 4 * 32) + lane;
    const double* wb = Z.Wd + static_cast<size_t>(ablk) * Z.n_q4 * (NT * 32) + lane;
    for (int ks = 0; ks < Z.n_q4; ++ks) {
      double bfrag[NT];
#pragma unroll
      for (int j = 0; j < NT; ++j) bfrag[j] = wb[j * 32];
#pragma unroll
      for (int mt = 0; mt < 4; ++mt) {
        const double L2K = c[(1 * 4 + mt) * 32], L2MU = c[(2 * 4 + mt) * 32];
        const double MU2 = c[(3 * 4 + mt) * 32], K2 = c[(4 * 4 + mt) * 32];
        double D = c[(0 * 4 + mt) * 32], PL = c[(5 * 4 + mt) * 32];
        if (cosmo) {                                                           // rescaled linear power, theory.py:54-56
          const double f = exp2(fma(pr.dn_64, L2K, pr.lr_64));
          D *= f;
          PL *= f;
        }
        const double nl = D * (q1 + q2 * D);                                   // natural-log units here
        const double v = exp2(av * (L2K - l2kv) + bv * L2MU);                  // (k/kv)^av mu^bv
        const double e = exp(nl * (1.0 - v) - K2 * ikp2);
        const double ka = 1.0 + betaT[mt] * MU2;
        const double pd = PL * ka * ka * e;
#pragma unroll
        for (int j = 0; j < NT; ++j) dmma_884(acc[mt][j][0], acc[mt][j][1], pd, bfrag[j]);
      }
      c += kCellConsts * 4 * 32;
      wb += NT * 32;
    }
    // epilogue (same composition as k_p3d_hankel)
    const int a_lane0 = ablk * (8 * NT) + 2 * t;
    const int n_a = Z.n_a, n_m = Z.n_m, n_m_pad = Z.n_m_pad;
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) {
      const int m = m_base + 8 * mt + g;
      const bool row_ok = m < n_m;
      const double kpar = Z.kpar[m];
      double bT = pr.bias;
      if (Z.flags & 1) bT = pr.bias + pr.b_H * exp(-kpar * pr.L_H);
      const double amp = bT * bT * Z.dAA;
      const double cont = (Z.flags & 8) ? tanh(pow(Z.kpar_row[m] / pr.kC, pr.pC)) : 1.0;
      double sky = 0.0;
      if ((Z.flags & 4) && row_ok) sky = pr.b_noise * Z.sky_smooth[m];
#pragma unroll
      for (int j = 0; j < NT; ++j)
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          const int a = a_lane0 + 8 * j + cc;
          if (a >= n_a) continue;
          double px = 0.0;
          if (row_ok) {
            px = amp * acc[mt][j][cc];
            if (Z.flags & 2) {
              const size_t o = static_cast<size_t>(a) * n_m + m, st = static_cast<size_t>(n_a) * n_m;
              const double bx = pr.b_X, btx = pr.beta_X;
              const double dn = pr.dn_64;
              const int ord = Z.metal_dn_order;
              double ma = bx * bx * (metal_moment(Z.metal_auto, Z.metal_auto_dn, ord, o, st, 0, dn) +
                                     2.0 * btx * metal_moment(Z.metal_auto, Z.metal_auto_dn, ord, o, st, 1, dn) +
                                     btx * btx * metal_moment(Z.metal_auto, Z.metal_auto_dn, ord, o, st, 2, dn));
              double mc = pr.bias * bx * (metal_moment(Z.metal_cross, Z.metal_cross_dn, ord, o, st, 0, dn) +
                                          (pr.beta + btx) * metal_moment(Z.metal_cross, Z.metal_cross_dn, ord, o, st, 1, dn) +
                                          pr.beta * btx * metal_moment(Z.metal_cross, Z.metal_cross_dn, ord, o, st, 2, dn));
              if (cosmo) {
                ma *= pr.lin_amp;
                mc *= pr.lin_amp;
              }
              px += ma;
              px += mc;
            }
            if (Z.flags & 4) px += sky * Z.sky_xi[a];
            px *= cont;
          }
          Z.pxzam[(static_cast<size_t>(pt) * n_a + a) * n_m_pad + m] = px;
        }
    }
    (void)rowwise;
  }
}

// ------------------------------------------------------------------------------------------
// k_window: Y[pt][M] = sum_{pairs (A,a)} sum_m T[pair][M][m] * Px_Zam[pt][a][m]   (FP64 DMMA)
//   MODE_CHI2 : T = L^-1 diag(V/Vsum) U, out chi2_zA[pt] = || yd[data_idx[pt]][A] - Y ||^2
//   MODE_MODEL: T =      diag(V/Vsum) U, out model[pt][A][M] = Y
// CTA = 4 warps, 64 points x (8*MT) coarse-k columns for one coarse theta bin A of one z.
// ------------------------------------------------------------------------------------------
constexpr int kWinPts = 64;
constexpr int kWinKC = 16;
constexpr int kWinLd = kWinKC + 4;   // padded row stride (doubles): conflict-free LDS.64 fragments

struct WindowArgs {
  int count;          // points in chunk
  int zslot;          // index of this z in the evaluated list, or -1: one launch covers all z (blockIdx.z)
  int n_zl;           // number of evaluated z
  int nA_max, nM_max; // output strides
  const int* data_idx;   // [count] or null
  double* chi2_zA;    // [count][n_zl][nA_max]
  double* model;      // [count][n_zl][nA_max][nM_max] or null
  int whiten;         // model mode with the whitened operator T_chi (basis vectors of the factorised path, cpx_factor.cuh)
};

template <int MT, bool CHI2>
__global__ void __launch_bounds__(128) k_window(const ZDev* __restrict__ ztab, int zi, WindowArgs w) {
  if (w.zslot < 0) {   // all evaluated z bins in one grid: fewer launches, no per-z wave quantisation
    zi = blockIdx.z;
    w.zslot = blockIdx.z;
  }
  const ZDev& Z = ztab[zi];
  const int A = blockIdx.y;
  const int pt_base = blockIdx.x * kWinPts;
  if (A >= Z.n_A) return;
  __shared__ __align__(16) double sA[2][kWinPts][kWinLd];
  __shared__ __align__(16) double sB[2][8 * MT][kWinLd];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int p0 = Z.pair_start[A], p1 = Z.pair_start[A + 1];
  const int kchunks = Z.n_m_pad / kWinKC;
  const int n_steps = (p1 - p0) * kchunks;
  const double* Tbase = (CHI2 || w.whiten) ? Z.T_chi : Z.T_model;
  double chi2_part[2] = {0.0, 0.0};

  // coarse-k columns in blocks of 8*MT (one block unless n_M > 64)
  for (int mb = 0; mb < Z.n_mblk; ++mb) {
    const int col0 = mb * 8 * MT;
    double acc[2][MT][2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < MT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    auto load_stage = [&](int step, int buf) {
      const int pair = p0 + step / kchunks;
      const int k0 = (step % kchunks) * kWinKC;
      const int a = Z.pair_a[pair];
      // A tile: 64 points x 16 doubles = 512 x 16B
      for (int i = tid; i < kWinPts * (kWinKC / 2); i += 128) {
        const int row = i / (kWinKC / 2), c2 = i % (kWinKC / 2);
        const int pt = min(pt_base + row, w.count - 1);
        const double* src = Z.pxzam + (static_cast<size_t>(pt) * Z.n_a + a) * Z.n_m_pad + k0 + 2 * c2;
        cp_async16(&sA[buf][row][2 * c2], src);
      }
      // B tile: (8*MT) coarse-k rows x 16 doubles
      for (int i = tid; i < 8 * MT * (kWinKC / 2); i += 128) {
        const int row = i / (kWinKC / 2), c2 = i % (kWinKC / 2);
        const double* src = Tbase + (static_cast<size_t>(pair) * Z.n_M_pad + col0 + row) * Z.n_m_pad + k0 + 2 * c2;
        cp_async16(&sB[buf][row][2 * c2], src);
      }
      cp_async_commit();
    };

    load_stage(0, 0);
    for (int step = 0; step < n_steps; ++step) {
      const int buf = step & 1;
      if (step + 1 < n_steps) {
        load_stage(step + 1, buf ^ 1);
        cp_async_wait<1>();
      } else {
        cp_async_wait<0>();
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < kWinKC; kk += 4) {
        double af[2];
        af[0] = sA[buf][warp * 16 + g][kk + t];
        af[1] = sA[buf][warp * 16 + 8 + g][kk + t];
#pragma unroll
        for (int j = 0; j < MT; ++j) {
          const double bf = sB[buf][8 * j + g][kk + t];
          dmma_884(acc[0][j][0], acc[0][j][1], af[0], bf);
          dmma_884(acc[1][j][0], acc[1][j][1], af[1], bf);
        }
      }
      __syncthreads();
    }

    // epilogue of this column block.  C fragment: rows g (+8), columns 8j + 2t + {0,1}
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int pt = pt_base + warp * 16 + 8 * i + g;
      if (CHI2) {
        // a data index outside the handle's data vectors (device-resident indices cannot be validated on the host) is clamped
        const int d = (w.data_idx != nullptr && pt < w.count) ? min(max(w.data_idx[pt], 0), Z.n_data - 1) : 0;
        const double* yd = Z.yd + (static_cast<size_t>(d) * Z.n_A + A) * Z.n_M_pad + col0;
        double s = chi2_part[i];
#pragma unroll
        for (int j = 0; j < MT; ++j) {
          const int col = 8 * j + 2 * t;
          const double y0 = yd[col] - acc[i][j][0];
          const double y1 = yd[col + 1] - acc[i][j][1];
          s = fma(y0, y0, s);
          s = fma(y1, y1, s);
        }
        chi2_part[i] = s;
      } else if (pt < w.count) {
        double* out = w.model + ((static_cast<size_t>(pt) * w.n_zl + w.zslot) * w.nA_max + A) * w.nM_max + col0;
#pragma unroll
        for (int j = 0; j < MT; ++j) {
          const int col = 8 * j + 2 * t;
          if (col0 + col < Z.n_M) out[col] = acc[i][j][0];
          if (col0 + col + 1 < Z.n_M) out[col + 1] = acc[i][j][1];
        }
      }
    }
  }
  if (CHI2) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int pt = pt_base + warp * 16 + 8 * i + g;
      double s = chi2_part[i];
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      if (t == 0 && pt < w.count)
        w.chi2_zA[(static_cast<size_t>(pt) * w.n_zl + w.zslot) * w.nA_max + A] = s;
    }
  }
}

// chi2 from an already computed model (used when both model and chi2 are requested is not needed:
// the two window modes are launched separately).

// ------------------------------------------------------------------------------------------
// k_reduce: chi2[pt] = sum over (z, A) in fixed order
// ------------------------------------------------------------------------------------------
__global__ void k_reduce(const double* __restrict__ chi2_zA, double* __restrict__ chi2, int count, int n_zl,
                         int nA_max, const int* __restrict__ nA_of_z) {
  const int pt = blockIdx.x * blockDim.x + threadIdx.x;
  if (pt >= count) return;
  double s = 0.0;
  for (int z = 0; z < n_zl; ++z) {
    const double* r = chi2_zA + (static_cast<size_t>(pt) * n_zl + z) * nA_max;
    for (int A = 0; A < nA_of_z[z]; ++A) s += r[A];
  }
  chi2[pt] = s;
}

// copy Px_Zam out of the padded per-z buffer: dst[pt][zslot][a][m]
__global__ void k_unpad_pxzam(const ZDev* __restrict__ ztab, int zi, int zslot, int n_zl, int na_max, int nm_max,
                              int count, double* __restrict__ dst) {
  const ZDev& Z = ztab[zi];
  const size_t n = static_cast<size_t>(count) * Z.n_a * Z.n_m;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int m = static_cast<int>(i % Z.n_m);
    const size_t r = i / Z.n_m;
    const int a = static_cast<int>(r % Z.n_a);
    const size_t pt = r / Z.n_a;
    dst[((pt * n_zl + zslot) * na_max + a) * nm_max + m] = Z.pxzam[(pt * Z.n_a + a) * Z.n_m_pad + m];
  }
}

// ------------------------------------------------------------------------------------------
// k_mock_noise: many noisy realisations of a data vector, generated where they are scored (likelihood.py:37-50 draws
// d = m + L n per mock on the host).  The device keeps data vectors whitened, yd = L^-1 d, so a mock is simply
//   yd[d][A][M] = (L^-1 m)[A][M] + n,   n ~ N(0, 1)
// with (L^-1 m) the whitened model from k_window (WindowArgs::whiten).  Counter-based generator (Philox4x32-10, key =
// seed, counter = (pair index, z, 0): reproducible and independent of the launch geometry), Box-Muller on two 53-bit
// uniforms per pair of elements.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t k0, uint32_t k1, uint32_t c[4]) {
#pragma unroll
  for (int round = 0; round < 10; ++round) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
    c[0] = n0;
    c[1] = lo1;
    c[2] = n2;
    c[3] = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}

__global__ void __launch_bounds__(256) k_mock_noise(const double* __restrict__ ymodel /*[n_A][n_M]*/, int n_A, int n_M, int n_M_pad,
                                                    int n_mocks, unsigned long long seed, int zglobal, double* __restrict__ yd) {
  const long long n_el = static_cast<long long>(n_mocks) * n_A * n_M;
  const long long pair = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (2 * pair >= n_el) return;
  uint32_t c[4] = {static_cast<uint32_t>(pair), static_cast<uint32_t>(pair >> 32), static_cast<uint32_t>(zglobal), 0u};
  philox4x32_10(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32), c);
  // two uniforms in (0, 1): 53 bits each, never exactly 0
  const double u1 = (static_cast<double>((static_cast<unsigned long long>(c[0] >> 5) << 26) | (c[1] >> 6)) + 0.5) * 0x1.0p-53;
  const double u2 = (static_cast<double>((static_cast<unsigned long long>(c[2] >> 5) << 26) | (c[3] >> 6)) + 0.5) * 0x1.0p-53;
  const double rad = sqrt(-2.0 * log(u1));
  double sn, cs;
  sincospi(2.0 * u2, &sn, &cs);
  const double z01[2] = {rad * cs, rad * sn};
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const long long el = 2 * pair + e;
    if (el >= n_el) break;
    const int M = static_cast<int>(el % n_M);
    const long long dA = el / n_M;
    const int A = static_cast<int>(dA % n_A);
    yd[dA * n_M_pad + M] = ymodel[A * n_M + M] + z01[e];
  }
}

// ydd[d] = sum_{A, M} yd^2 (the constant term of the factorised chi2); one warp per data vector
__global__ void k_ydd(const double* __restrict__ yd, int per_data, int n_data, double* __restrict__ ydd) {
  const int d = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (d >= n_data) return;
  double s = 0.0;
  for (int i = lane; i < per_data; i += 32) {
    const double v = yd[static_cast<size_t>(d) * per_data + i];
    s = fma(v, v, s);
  }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) ydd[d] = s;
}

// ------------------------------------------------------------------------------------------
// k_jbar: set-up helper for the Hankel weights (cupix_b200/quadrature.py): the theta-weighted average of J0 over the
// separations of each fine theta bin (likelihood.py:99-130 folded into the weights) on the fine k_perp grid,
//   jbar[g][f] = meas[f] / sum_{i in g} w_i  *  sum_{i in g} w_i J0(kf[f] r_i)
// grid = (n_f / 256, n_groups).  2000 separations x 75 000 grid points per z bin at the S5 shape: seconds of NumPy, ms here.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_jbar(const int* __restrict__ gstart, const double* __restrict__ r, const double* __restrict__ w,
                                              const double* __restrict__ kf, const double* __restrict__ meas, long long n_f,
                                              double* __restrict__ jbar) {
  const long long f = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (f >= n_f) return;
  const int g = blockIdx.y;
  const double k = kf[f];
  double acc = 0.0, norm = 0.0;
  for (int i = gstart[g]; i < gstart[g + 1]; ++i) {
    acc = fma(w[i], j0(k * r[i]), acc);
    norm += w[i];
  }
  jbar[static_cast<size_t>(g) * n_f + f] = acc * (meas[f] / norm);
}

// ------------------------------------------------------------------------------------------
// k_hankel_apply: the Hankel contraction alone, for P3D cells the caller evaluated (Theory._compute_px_from_p3d with an
// arbitrary P3D callable, theory.py:250-264):  px[r][k] = sum_q W[r][q] cells[k][q]  as an FP64 DMMA GEMM.
// A warp owns 8 separations x 64 k_par values; both operands are row-major with q contiguous, which is the m8n8k4
// fragment order of A (row.) and B (.col) as they lie in memory.  grid = (ceil(n_k / 64), ceil(n_r / 8 / 4)), 4 warps.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_hankel_apply(const double* __restrict__ W, const double* __restrict__ cells, int n_r, int n_q,
                                                      int n_k, double* __restrict__ px) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int r0 = (blockIdx.y * 4 + warp) * 8, k0 = blockIdx.x * 64;
  if (r0 >= n_r) return;
  const int ra = min(r0 + g, n_r - 1);
  double acc[8][2];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j][0] = acc[j][1] = 0.0;
  for (int q0 = 0; q0 < n_q; q0 += 4) {
    const bool in_q = q0 + t < n_q;
    const double a = in_q ? W[static_cast<size_t>(ra) * n_q + q0 + t] : 0.0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int kb = min(k0 + 8 * j + g, n_k - 1);
      const double b = in_q ? cells[static_cast<size_t>(kb) * n_q + q0 + t] : 0.0;
      dmma_884(acc[j][0], acc[j][1], a, b);
    }
  }
  if (r0 + g < n_r) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int kc = k0 + 8 * j + 2 * t;
      if (kc < n_k) px[static_cast<size_t>(r0 + g) * n_k + kc] = acc[j][0];
      if (kc + 1 < n_k) px[static_cast<size_t>(r0 + g) * n_k + kc + 1] = acc[j][1];
    }
  }
}

// ------------------------------------------------------------------------------------------
// pipe-rate probes (roofline denominators)
// ------------------------------------------------------------------------------------------
__global__ void k_probe_dfma(double* out, int iters, double x) {
  double a0 = threadIdx.x, a1 = 1, a2 = 2, a3 = 3, a4 = 4, a5 = 5, a6 = 6, a7 = 7;
  const double y = x * 0.5;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, x, y); a1 = fma(a1, x, y); a2 = fma(a2, x, y); a3 = fma(a3, x, y);
    a4 = fma(a4, x, y); a5 = fma(a5, x, y); a6 = fma(a6, x, y); a7 = fma(a7, x, y);
  }
  if (a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 == 12345.678) out[0] = a0;
}
__global__ void k_probe_dmma(double* out, int iters, double x) {
  double c[8][2];
#pragma unroll
  for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = threadIdx.x + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) dmma_884(c[i][0], c[i][1], x, 0.5);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  if (s == 12345.678) out[0] = s;
}
__global__ void k_probe_ffma(float* out, int iters, float x) {
  float a0 = threadIdx.x, a1 = 1, a2 = 2, a3 = 3, a4 = 4, a5 = 5, a6 = 6, a7 = 7;
  const float y = x * 0.5f;
  for (int i = 0; i < iters; ++i) {
    a0 = fmaf(a0, x, y); a1 = fmaf(a1, x, y); a2 = fmaf(a2, x, y); a3 = fmaf(a3, x, y);
    a4 = fmaf(a4, x, y); a5 = fmaf(a5, x, y); a6 = fmaf(a6, x, y); a7 = fmaf(a7, x, y);
  }
  if (a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 == 12345.678f) out[0] = a0;
}
__global__ void k_probe_mufu(float* out, int iters, float x) {
  float a0 = x, a1 = x + 1, a2 = x + 2, a3 = x + 3, a4 = x + 4, a5 = x + 5, a6 = x + 6, a7 = x + 7;
  for (int i = 0; i < iters; ++i) {
    a0 = ex2_approx(a0); a1 = ex2_approx(a1); a2 = ex2_approx(a2); a3 = ex2_approx(a3);
    a4 = ex2_approx(a4); a5 = ex2_approx(a5); a6 = ex2_approx(a6); a7 = ex2_approx(a7);
  }
  if (a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 == 12345.678f) out[0] = a0;
}

}  // namespace cpx

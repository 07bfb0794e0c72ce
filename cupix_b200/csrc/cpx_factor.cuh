// Factorised chi2 path: the amplitude parameters taken out of the Hankel stage.
//
// The Px model of theory.py:95-123 is a *polynomial* in (bias, beta, b_H, beta_H, b_X, beta_X, b_noise) whose
// coefficients are tables that depend on the point only through the remaining ("tuple") parameters
// (q1, q2, kv, av, bv, kp; L_H, kC, pC; As, ns):
//
//   Lya (+HCD, theory.py:267-344):   P3D = b_T^2 (1 + beta_T mu^2)^2 P_L D_NL,  b_T = b + h F(k_par),
//                                    b_T beta_T = b beta + h eta F            (F = exp(-k_par L_H))
//        => Px_Zam = dAA cont(m) sum_{i=0..2} sum_{j=0..2} c_ij F^i H_j,   H_j[a][m] = sum_q W[a][q] mu^2j P_L D_NL
//           c_0j = (b^2, 2 b^2 beta, b^2 beta^2), c_1j = (2bh, 2bh(beta + eta), 2bh beta eta), c_2j = (h^2, 2h^2 eta, h^2 eta^2)
//   SiIII (theory.py:361-438):       x^2 (Ma_0 + 2 xi Ma_1 + xi^2 Ma_2) + b x (Mc_0 + (beta + xi) Mc_1 + beta xi Mc_2)
//   sky (theory.py:476-506):         n xi_noise(theta) sinc^2
//   continuum (theory.py:525-535):   multiplies every table (cont(m), a tuple parameter)
//
// Window convolution, theta rebinning and whitening are linear (likelihood.py:53-88), so with
// Y_k = L^-1 diag(V/sum V) U B_k the chi2 of likelihood.py:170-194 is a quadratic form per point,
//
//   chi2 = sum_z  dd - 2 c.g + c^T G c,      G_kl = Y_k . Y_l,  g_k = (L^-1 d) . Y_k,  dd = |L^-1 d|^2,
//
// and the Hankel stage runs once per distinct tuple instead of once per point -- the structure of the
// reference's scan scripts (scripts/chi2_scan_mock.py:107-140: a bias x beta grid is ONE tuple;
// scripts/chi2_scan_forecast.py:116-125: 64^4 points are 4096 tuples).  Because so few cells are left, they
// are evaluated in float64 (the formulas of theory.py:285-308 as written): chi2 from this path agrees with
// the float64 oracle to ~1e-9 everywhere, not only near the best fit.
//
//   k_basis_f64    per (z, tuple): float64 P3D cells, three mu-moment Hankel sums on the FP64 tensor pipe (DMMA), epilogue
//                  writes the n_b basis tables B_k[a][m] as pseudo-points of the Px_Zam scratch
//   k_window_gram  per (z, 64 / n_b tuples): the whitened window of the basis tables (FP64 DMMA) and G, g in one kernel;
//                  the Y_k stay in shared memory.  For n_M > 64 or many data vectors: k_window<MT, false> with
//                  WindowArgs::whiten (Y_k to HBM) followed by k_gram
//   k_quad         per point: coefficients + quadratic form, summed over z
//   k_p3d_hankel_f64s  (general points, CPX_PRECISE_CELLS) the float64-cell Hankel kernel on the same staged skeleton
#pragma once
#include "cpx_kernels.cuh"

namespace cpx {

constexpr int kMaxBasis = 16;

// number of basis tables for a flag set (CPX_INCLUDE_*)
__host__ __device__ constexpr int basis_count(int flags) {
  return ((flags & 1) ? 9 : 3) + ((flags & 2) ? 6 : 0) + ((flags & 4) ? 1 : 0);
}

// ------------------------------------------------------------------------------------------
// k_basis_f64: one warp per (z, 32-row k_par tile, theta block, tuple).  Same DMMA fragment layout as
// k_p3d_hankel_f64; the four m-tiles are walked one after the other so that only 3 x NT accumulators are live.
// Output: Z.pxzam viewed as [tuple * nb + k][a][m_pad].
// STAGE: the float64 cell constants of the (z, tile) and the weights of the theta block are staged in shared memory and
// shared by the 16 warps of the CTA (16 consecutive tuples); a CTA walks a contiguous range of (z, tile, theta block,
// tuple group) items, so the 86 KB tile (n_q = 56) is reloaded only at tile boundaries.  Without STAGE (node sets whose
// tile does not fit) every warp reads its constants through L1/L2.
// WARPS = 16 (default): 4 warps per scheduler -- the cell evaluation is a long dependent chain (exp2 -> exp -> DMMA) -- at
// 128 registers; WARPS = 8 keeps 160 registers and no spills (CPX_BASIS_WARPS selects).  Measured per S5 scan: 12.0 ms
// unstaged, 7.2 ms staged with 8 warps, 5.0 ms staged with 16.
// ------------------------------------------------------------------------------------------
__host__ __device__ constexpr size_t basis_smem_bytes(int n_q4, int nt) {
  return static_cast<size_t>(n_q4) * (kCellConsts * 4 * 32 + nt * 32) * sizeof(double);
}

template <int NT, bool STAGE, int WARPS>
__global__ void __launch_bounds__(32 * WARPS) k_basis_f64(const ZDev* __restrict__ ztab, int n_z, int count, long long n_items, int nb) {
  extern __shared__ __align__(16) double sm_basis[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  constexpr int kBasisWarps = WARPS;
  const int n_grp = (count + kBasisWarps - 1) / kBasisWarps;           // tuple groups per (z, tile, theta block)
  const long long it0 = n_items * blockIdx.x / gridDim.x, it1 = n_items * (blockIdx.x + 1) / gridDim.x;
  int cur_z = -1, cur_tile = -1, cur_ablk = -1;
  for (long long it = it0; it < it1; ++it) {
    int zi = 0;
    while (zi + 1 < n_z && ztab[zi + 1].item_start <= it) ++zi;
    const ZDev& Z = ztab[zi];
    long long r = it - Z.item_start;
    const int grp = static_cast<int>(r % n_grp);
    r /= n_grp;
    const int ablk = static_cast<int>(r % Z.n_ablk);
    const int tile = static_cast<int>(r / Z.n_ablk);
    const int n_ks = Z.n_q4;
    const double* cg = Z.cell64 + static_cast<size_t>(tile) * n_ks * (kCellConsts * 4 * 32);
    const double* wg = Z.Wd + static_cast<size_t>(ablk) * n_ks * (NT * 32);
    if (STAGE) {
      double* const s_w = sm_basis + static_cast<size_t>(n_ks) * (kCellConsts * 4 * 32);
      const bool new_tile = zi != cur_z || tile != cur_tile, new_w = zi != cur_z || ablk != cur_ablk;
      if (new_tile || new_w) {
        __syncthreads();                                                   // every warp is done with the previous tile
        if (new_tile) {
          const double2* src = reinterpret_cast<const double2*>(cg);
          double2* dst = reinterpret_cast<double2*>(sm_basis);
          for (int i = threadIdx.x; i < n_ks * (kCellConsts * 4 * 32) / 2; i += blockDim.x) dst[i] = src[i];
        }
        if (new_w) {
          const double2* src = reinterpret_cast<const double2*>(wg);
          double2* dst = reinterpret_cast<double2*>(s_w);
          for (int i = threadIdx.x; i < n_ks * (NT * 32) / 2; i += blockDim.x) dst[i] = src[i];
        }
        __syncthreads();
        cur_z = zi;
        cur_tile = tile;
        cur_ablk = ablk;
      }
      cg = sm_basis;
      wg = s_w;
    }
    const int tp = grp * kBasisWarps + warp;
    if (tp >= count) continue;
    const int m_base = tile * kTileRows;
    const PtPar& pr = Z.ptpar[tp];
    const double q1 = pr.q1_64, q2 = pr.q2_64, av = pr.av_64, bv = pr.bv_64, l2kv = pr.l2kv_64, ikp2 = pr.ikp2_64;
    const bool cosmo = pr.dn_64 != 0.0 || pr.lr_64 != 0.0;
    const int n_a = Z.n_a, n_m = Z.n_m, n_m_pad = Z.n_m_pad;
    const int a_lane0 = ablk * (8 * NT) + 2 * t;
    const int n_i = (Z.flags & 1) ? 3 : 1;
    const size_t tab = static_cast<size_t>(n_a) * n_m_pad;                      // doubles per basis table
    double* const out0 = Z.pxzam + static_cast<size_t>(tp) * nb * tab;
#pragma unroll 1
    for (int mt = 0; mt < 4; ++mt) {
      double acc[3][NT][2];
#pragma unroll
      for (int jm = 0; jm < 3; ++jm)
#pragma unroll
        for (int j = 0; j < NT; ++j) acc[jm][j][0] = acc[jm][j][1] = 0.0;
      const double* c = cg + lane;
      const double* wb = wg + lane;
#pragma unroll 1
      for (int ks = 0; ks < n_ks; ++ks) {
        double bfrag[NT];
#pragma unroll
        for (int j = 0; j < NT; ++j) bfrag[j] = wb[j * 32];
        const double L2K = c[(1 * 4 + mt) * 32], L2MU = c[(2 * 4 + mt) * 32];
        const double MU2 = c[(3 * 4 + mt) * 32], K2 = c[(4 * 4 + mt) * 32];
        double D = c[(0 * 4 + mt) * 32], PL = c[(5 * 4 + mt) * 32];
        if (cosmo) {                                                             // rescaled linear power, theory.py:54-56
          const double f = exp2(fma(pr.dn_64, L2K, pr.lr_64));
          D *= f;
          PL *= f;
        }
        const double nl = D * (q1 + q2 * D);                                     // theory.py:285-295
        const double v = exp2(av * (L2K - l2kv) + bv * L2MU);
        const double p0 = PL * exp(nl * (1.0 - v) - K2 * ikp2);
        const double p1 = p0 * MU2, p2 = p1 * MU2;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
          dmma_884(acc[0][j][0], acc[0][j][1], p0, bfrag[j]);
          dmma_884(acc[1][j][0], acc[1][j][1], p1, bfrag[j]);
          dmma_884(acc[2][j][0], acc[2][j][1], p2, bfrag[j]);
        }
        c += kCellConsts * 4 * 32;
        wb += NT * 32;
      }
      // ---- epilogue of this m-tile: row m = m_base + 8 mt + g, theta columns a_lane0 + 8 j + {0, 1} ----
      const int m = m_base + 8 * mt + g;
      const bool row_ok = m < n_m;
      const double kpar = Z.kpar[m];
      const double F = (Z.flags & 1) ? exp(-kpar * pr.L_H) : 1.0;                // theory.py:267-282
      double cont = (Z.flags & 8) ? tanh(pow(Z.kpar_row[m] / pr.kC, pr.pC)) : 1.0;   // theory.py:525-535
      if (!row_ok) cont = 0.0;
      const double lya = cont * Z.dAA;
      const double sky = ((Z.flags & 4) && row_ok) ? cont * Z.sky_smooth[m] : 0.0;
#pragma unroll
      for (int j = 0; j < NT; ++j)
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          const int a = a_lane0 + 8 * j + cc;
          if (a >= n_a) continue;
          double* o = out0 + static_cast<size_t>(a) * n_m_pad + m;
          double Fi = lya;
          for (int i = 0; i < n_i; ++i) {
#pragma unroll
            for (int jm = 0; jm < 3; ++jm) o[static_cast<size_t>(3 * i + jm) * tab] = Fi * acc[jm][j][cc];
            Fi *= F;
          }
          int k = 3 * n_i;
          if (Z.flags & 2) {                                                     // SiIII moments, theory.py:361-438
            const size_t om = static_cast<size_t>(a) * n_m + m, st = static_cast<size_t>(n_a) * n_m;
            const double f = cont * pr.lin_amp;
#pragma unroll
            for (int jm = 0; jm < 3; ++jm) {
              o[static_cast<size_t>(k + jm) * tab] =
                  row_ok ? f * metal_moment(Z.metal_auto, Z.metal_auto_dn, Z.metal_dn_order, om, st, jm, pr.dn_64) : 0.0;
              o[static_cast<size_t>(k + 3 + jm) * tab] =
                  row_ok ? f * metal_moment(Z.metal_cross, Z.metal_cross_dn, Z.metal_dn_order, om, st, jm, pr.dn_64) : 0.0;
            }
            k += 6;
          }
          if (Z.flags & 4) o[static_cast<size_t>(k) * tab] = sky * Z.sky_xi[a];  // theory.py:476-506
        }
    }
  }
}

// ------------------------------------------------------------------------------------------
// k_p3d_hankel_f64s: the reference-precision cell evaluation of k_p3d_hankel_f64 (CPX_PRECISE_CELLS) on the skeleton of
// k_basis_f64 -- float64 cell constants and weights staged in shared memory, WARPS points per CTA, contiguous item ranges.
// Same arithmetic as k_p3d_hankel_f64 (theory.py:285-344 as written), one accumulator set; epilogue as in k_p3d_hankel.
// ------------------------------------------------------------------------------------------
template <int NT, int WARPS>
__global__ void __launch_bounds__(32 * WARPS) k_p3d_hankel_f64s(const ZDev* __restrict__ ztab, int n_z, int count, long long n_items) {
  extern __shared__ __align__(16) double sm_basis[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int n_grp = (count + WARPS - 1) / WARPS;
  const long long it0 = n_items * blockIdx.x / gridDim.x, it1 = n_items * (blockIdx.x + 1) / gridDim.x;
  int cur_z = -1, cur_tile = -1, cur_ablk = -1;
  for (long long it = it0; it < it1; ++it) {
    int zi = 0;
    while (zi + 1 < n_z && ztab[zi + 1].item_start <= it) ++zi;
    const ZDev& Z = ztab[zi];
    long long r = it - Z.item_start;
    const int grp = static_cast<int>(r % n_grp);
    r /= n_grp;
    const int ablk = static_cast<int>(r % Z.n_ablk);
    const int tile = static_cast<int>(r / Z.n_ablk);
    const int n_ks = Z.n_q4;
    double* const s_w = sm_basis + static_cast<size_t>(n_ks) * (kCellConsts * 4 * 32);
    const bool new_tile = zi != cur_z || tile != cur_tile, new_w = zi != cur_z || ablk != cur_ablk;
    if (new_tile || new_w) {
      __syncthreads();
      if (new_tile) {
        const double2* src = reinterpret_cast<const double2*>(Z.cell64 + static_cast<size_t>(tile) * n_ks * (kCellConsts * 4 * 32));
        double2* dst = reinterpret_cast<double2*>(sm_basis);
        for (int i = threadIdx.x; i < n_ks * (kCellConsts * 4 * 32) / 2; i += blockDim.x) dst[i] = src[i];
      }
      if (new_w) {
        const double2* src = reinterpret_cast<const double2*>(Z.Wd + static_cast<size_t>(ablk) * n_ks * (NT * 32));
        double2* dst = reinterpret_cast<double2*>(s_w);
        for (int i = threadIdx.x; i < n_ks * (NT * 32) / 2; i += blockDim.x) dst[i] = src[i];
      }
      __syncthreads();
      cur_z = zi;
      cur_tile = tile;
      cur_ablk = ablk;
    }
    const int pt = grp * WARPS + warp;
    if (pt >= count) continue;
    const int m_base = tile * kTileRows;
    const PtPar& pr = Z.ptpar[pt];
    const double q1 = pr.q1_64, q2 = pr.q2_64, av = pr.av_64, bv = pr.bv_64, l2kv = pr.l2kv_64, ikp2 = pr.ikp2_64;
    const bool cosmo = pr.dn_64 != 0.0 || pr.lr_64 != 0.0;
    const int n_a = Z.n_a, n_m = Z.n_m, n_m_pad = Z.n_m_pad;
    const int a_lane0 = ablk * (8 * NT) + 2 * t;
#pragma unroll 1
    for (int mt = 0; mt < 4; ++mt) {
      const int m = m_base + 8 * mt + g;
      const bool row_ok = m < n_m;
      const double kpar = Z.kpar[m];
      double bT = pr.bias, betaT = pr.beta;
      if (Z.flags & 1) {                                                         // theory.py:267-282
        const double F = exp(-kpar * pr.L_H);
        bT = pr.bias + pr.b_H * F;
        betaT = (pr.bias * pr.beta + pr.b_H * pr.beta_H * F) / bT;
      }
      double acc[NT][2];
#pragma unroll
      for (int j = 0; j < NT; ++j) acc[j][0] = acc[j][1] = 0.0;
      const double* c = sm_basis + lane;
      const double* wb = s_w + lane;
#pragma unroll 1
      for (int ks = 0; ks < n_ks; ++ks) {
        const double L2K = c[(1 * 4 + mt) * 32], L2MU = c[(2 * 4 + mt) * 32];
        const double MU2 = c[(3 * 4 + mt) * 32], K2 = c[(4 * 4 + mt) * 32];
        double D = c[(0 * 4 + mt) * 32], PL = c[(5 * 4 + mt) * 32];
        if (cosmo) {
          const double f = exp2(fma(pr.dn_64, L2K, pr.lr_64));
          D *= f;
          PL *= f;
        }
        const double nl = D * (q1 + q2 * D);
        const double v = exp2(av * (L2K - l2kv) + bv * L2MU);
        const double e = exp(nl * (1.0 - v) - K2 * ikp2);
        const double ka = 1.0 + betaT * MU2;
        const double pd = PL * ka * ka * e;
#pragma unroll
        for (int j = 0; j < NT; ++j) dmma_884(acc[j][0], acc[j][1], pd, wb[j * 32]);
        c += kCellConsts * 4 * 32;
        wb += NT * 32;
      }
      const double amp = bT * bT * Z.dAA;
      const double cont = (Z.flags & 8) ? tanh(pow(Z.kpar_row[m] / pr.kC, pr.pC)) : 1.0;
      const double sky = ((Z.flags & 4) && row_ok) ? pr.b_noise * Z.sky_smooth[m] : 0.0;
#pragma unroll
      for (int j = 0; j < NT; ++j)
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          const int a = a_lane0 + 8 * j + cc;
          if (a >= n_a) continue;
          double px = 0.0;
          if (row_ok) {
            px = amp * acc[j][cc];
            if (Z.flags & 2) {
              const size_t o = static_cast<size_t>(a) * n_m + m, st = static_cast<size_t>(n_a) * n_m;
              const double bx = pr.b_X, btx = pr.beta_X, dn = pr.dn_64;
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
  }
}

// ------------------------------------------------------------------------------------------
// k_gram: block = (tuple, z).  Y[(tuple * nb + k)][z][A][M] (the model-output layout of k_window) ->
//   G[tuple][z][k][l] = sum_{A, M} Y_k Y_l,   g[tuple][z][d][k] = sum_{A, M} yd[d][A][M] Y_k
// rows are padded to an odd stride so that the threads of a warp,
// which hold different (k, l), read distinct banks.
// ------------------------------------------------------------------------------------------
// `at` coarse theta bins are staged per barrier pair (host: ~24 KB of shared memory, so that 8 CTAs per SM overlap each
// other's loads; staging everything at once -- 93 KB, 2 CTAs per SM -- measured slower)
__host__ __device__ constexpr int gram_row_stride(int nM_max, int at) { return (at * (nM_max | 1)) | 1; }

__global__ void __launch_bounds__(256) k_gram(const ZDev* __restrict__ ztab, const double* __restrict__ Y, int n_zl, int nA_max,
                                              int nM_max, int nb, int nd_max, int at, double* __restrict__ G, double* __restrict__ gv) {
  extern __shared__ double sY[];                       // [nb][row stride]: row k holds `at` x (nM_max | 1) values; then the partial sums
  const int tp = blockIdx.x, zi = blockIdx.y;
  const ZDev& Z = ztab[zi];
  const int tid = threadIdx.x;
  const int nM = Z.n_M, nd = Z.n_data, nA = Z.n_A;
  const int LD = nM_max | 1, RS = gram_row_stride(nM_max, at);
  // every (k, l) pair is shared by n_seg threads, each taking every n_seg-th coarse-k column with four independent
  // accumulators: 243 of 256 threads busy for 9 tables, and no 1280-long dependent chain of FMAs behind two loads each
  const int n_pairs = nb * nb, n_seg = max(1, static_cast<int>(blockDim.x) / n_pairs);
  const int pair = tid % n_pairs, seg = tid / n_pairs;
  const bool dot = seg < n_seg;
  const int k = pair / nb, l = pair % nb;
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  double* const gz = gv + (static_cast<size_t>(tp) * n_zl + zi) * nd_max * nb;
  const size_t ystride = static_cast<size_t>(n_zl) * nA_max * nM_max;          // between basis tables of a tuple
  const double* const Y0 = Y + (static_cast<size_t>(tp) * nb * n_zl + zi) * nA_max * nM_max;
  for (int A0 = 0; A0 < nA; A0 += at) {
    const int na = min(at, nA - A0);
    __syncthreads();
    for (int i = tid; i < nb * na * nM; i += blockDim.x) {
      const int M = i % nM, rest = i / nM;
      const int a = rest % na, kk = rest / na;
      sY[kk * RS + a * LD + M] = Y0[kk * ystride + static_cast<size_t>(A0 + a) * nM_max + M];
    }
    __syncthreads();
    if (dot) {
      for (int a = 0; a < na; ++a) {
        const double* yk = sY + k * RS + a * LD;
        const double* yl = sY + l * RS + a * LD;
        int M = seg;
        for (; M + 3 * n_seg < nM; M += 4 * n_seg) {
          acc[0] = fma(yk[M], yl[M], acc[0]);
          acc[1] = fma(yk[M + n_seg], yl[M + n_seg], acc[1]);
          acc[2] = fma(yk[M + 2 * n_seg], yl[M + 2 * n_seg], acc[2]);
          acc[3] = fma(yk[M + 3 * n_seg], yl[M + 3 * n_seg], acc[3]);
        }
        for (; M < nM; M += n_seg) acc[0] = fma(yk[M], yl[M], acc[0]);
      }
    }
    for (int item = tid; item < nd * nb; item += blockDim.x) {
      const int d = item / nb, kk = item % nb;
      double s = (A0 == 0) ? 0.0 : gz[item];
      for (int a = 0; a < na; ++a) {
        const double* yd = Z.yd + (static_cast<size_t>(d) * nA + A0 + a) * Z.n_M_pad;
        const double* yk = sY + kk * RS + a * LD;
        for (int M = 0; M < nM; ++M) s = fma(yd[M], yk[M], s);
      }
      gz[item] = s;
    }
  }
  __syncthreads();                                     // the staging area becomes [n_seg][n_pairs] partial sums
  if (dot) sY[seg * n_pairs + pair] = (acc[0] + acc[1]) + (acc[2] + acc[3]);
  __syncthreads();
  if (tid < n_pairs) {
    double s = 0.0;
    for (int sg = 0; sg < n_seg; ++sg) s += sY[sg * n_pairs + tid];
    G[(static_cast<size_t>(tp) * n_zl + zi) * n_pairs + tid] = s;
  }
}

// ------------------------------------------------------------------------------------------
// k_window_gram: the whitened window of the basis tables and their Gram data in one kernel (n_M <= 64, few data vectors).
// CTA = 4 warps = 64 pseudo-points = the basis tables of 64 / nb whole tuples, one z bin, ALL coarse theta bins in turn:
// the FP64 DMMA contraction of k_window per theta bin, the 64 x n_M tile of Y written to shared memory (over the operand
// stages, which are idle then), and every (tuple, k, l) dot product accumulated in the registers of one fixed thread --
// deterministic, and the Y_k never travel to HBM (912 MB written and read back per tuple batch by k_window + k_gram).
// ------------------------------------------------------------------------------------------
constexpr int kWgMaxAcc = 8;     // (tuple, k, l) triples per thread: 64 / nb tuples x nb^2 pairs over 128 threads (nb = 16: 8)
constexpr int kWgMaxG = 4;       // (tuple, data vector, k) items per thread: 64 / nb x nd x nb <= 512

// blockIdx.z splits the coarse theta bins between kWgSplit CTAs (partial sums, added by k_gsum in fixed order): the
// 1416 CTAs of an S5 tuple batch are 2.4 waves of 4 CTAs per SM, 2832 half-length ones 4.8
constexpr int kWgSplit = 2;

template <int MT>
__global__ void __launch_bounds__(128) k_window_gram(const ZDev* __restrict__ ztab, int npseudo, int n_zl, int nb, int nd_max,
                                                     double* __restrict__ G, double* __restrict__ gv, size_t g_stride, size_t gv_stride) {
  const int zi = blockIdx.y;
  G += blockIdx.z * g_stride;
  gv += blockIdx.z * gv_stride;
  const ZDev& Z = ztab[zi];
  const int tp_cta = kWinPts / nb;                       // whole tuples per CTA
  const int tup0 = blockIdx.x * tp_cta;
  const int pt_base = tup0 * nb;
  constexpr int LDY = 8 * MT + 1;
  __shared__ __align__(16) double smem[2 * kWinPts * kWinLd + 2 * 8 * MT * kWinLd];
  double (*sA)[kWinPts][kWinLd] = reinterpret_cast<double (*)[kWinPts][kWinLd]>(smem);
  double (*sB)[8 * MT][kWinLd] = reinterpret_cast<double (*)[8 * MT][kWinLd]>(smem + 2 * kWinPts * kWinLd);
  double* const sY = smem;                               // [64][LDY], aliases the operand stages between theta bins
  static_assert(kWinPts * LDY <= 2 * kWinPts * kWinLd + 2 * 8 * MT * kWinLd, "Y tile must fit the operand stages");

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int kchunks = Z.n_m_pad / kWinKC;
  const int nM = Z.n_M, nd = Z.n_data, nA = Z.n_A;
  const int n_trip = tp_cta * nb * nb, n_gi = tp_cta * nd * nb;
  double gacc[kWgMaxAcc], hacc[kWgMaxG];
#pragma unroll
  for (int i = 0; i < kWgMaxAcc; ++i) gacc[i] = 0.0;
#pragma unroll
  for (int i = 0; i < kWgMaxG; ++i) hacc[i] = 0.0;

  const int a_per = (nA + gridDim.z - 1) / gridDim.z;
  const int A_end = min(nA, static_cast<int>(blockIdx.z + 1) * a_per);
  for (int A = blockIdx.z * a_per; A < A_end; ++A) {
    const int p0 = Z.pair_start[A], p1 = Z.pair_start[A + 1];
    const int n_steps = (p1 - p0) * kchunks;
    double acc[2][MT][2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < MT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    auto load_stage = [&](int step, int buf) {
      const int pair = p0 + step / kchunks;
      const int k0 = (step % kchunks) * kWinKC;
      const int a = Z.pair_a[pair];
      for (int i = tid; i < kWinPts * (kWinKC / 2); i += 128) {
        const int row = i / (kWinKC / 2), c2 = i % (kWinKC / 2);
        const int pt = min(pt_base + row, npseudo - 1);
        cp_async16(&sA[buf][row][2 * c2], Z.pxzam + (static_cast<size_t>(pt) * Z.n_a + a) * Z.n_m_pad + k0 + 2 * c2);
      }
      for (int i = tid; i < 8 * MT * (kWinKC / 2); i += 128) {
        const int row = i / (kWinKC / 2), c2 = i % (kWinKC / 2);
        cp_async16(&sB[buf][row][2 * c2], Z.T_chi + (static_cast<size_t>(pair) * Z.n_M_pad + row) * Z.n_m_pad + k0 + 2 * c2);
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
        const double af0 = sA[buf][warp * 16 + g][kk + t], af1 = sA[buf][warp * 16 + 8 + g][kk + t];
#pragma unroll
        for (int j = 0; j < MT; ++j) {
          const double bf = sB[buf][8 * j + g][kk + t];
          dmma_884(acc[0][j][0], acc[0][j][1], af0, bf);
          dmma_884(acc[1][j][0], acc[1][j][1], af1, bf);
        }
      }
      __syncthreads();
    }
    // ---- the Y tile of this theta bin into shared memory (C fragment: rows g (+8), columns 8j + 2t + {0, 1}) ----
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      double* yr = sY + (warp * 16 + 8 * i + g) * LDY;
#pragma unroll
      for (int j = 0; j < MT; ++j) {
        yr[8 * j + 2 * t] = acc[i][j][0];
        yr[8 * j + 2 * t + 1] = acc[i][j][1];
      }
    }
    __syncthreads();
    // ---- Gram contributions of this theta bin ----
#pragma unroll
    for (int i = 0; i < kWgMaxAcc; ++i) {
      const int tr = tid + 128 * i;
      if (tr < n_trip) {
        const int tl = tr / (nb * nb), kl = tr % (nb * nb);
        const double* yk = sY + (tl * nb + kl / nb) * LDY;
        const double* yl = sY + (tl * nb + kl % nb) * LDY;
        double s0 = 0.0, s1 = 0.0;
        int M = 0;
        for (; M + 1 < nM; M += 2) {
          s0 = fma(yk[M], yl[M], s0);
          s1 = fma(yk[M + 1], yl[M + 1], s1);
        }
        if (M < nM) s0 = fma(yk[M], yl[M], s0);
        gacc[i] += s0 + s1;
      }
    }
#pragma unroll
    for (int i = 0; i < kWgMaxG; ++i) {
      const int it = tid + 128 * i;
      if (it < n_gi) {
        const int kk = it % nb, d = (it / nb) % nd, tl = it / (nb * nd);
        const double* yd = Z.yd + (static_cast<size_t>(d) * nA + A) * Z.n_M_pad;
        const double* yk = sY + (tl * nb + kk) * LDY;
        double s = 0.0;
        for (int M = 0; M < nM; ++M) s = fma(yd[M], yk[M], s);
        hacc[i] += s;
      }
    }
    __syncthreads();                                     // the next theta bin's operand loads overwrite the Y tile
  }
  const int nt = (npseudo + nb - 1) / nb;
#pragma unroll
  for (int i = 0; i < kWgMaxAcc; ++i) {
    const int tr = tid + 128 * i;
    if (tr < n_trip) {
      const int tl = tr / (nb * nb), kl = tr % (nb * nb);
      if (tup0 + tl < nt) G[(static_cast<size_t>(tup0 + tl) * n_zl + zi) * nb * nb + kl] = gacc[i];
    }
  }
#pragma unroll
  for (int i = 0; i < kWgMaxG; ++i) {
    const int it = tid + 128 * i;
    if (it < n_gi) {
      const int kk = it % nb, d = (it / nb) % nd, tl = it / (nb * nd);
      if (tup0 + tl < nt) gv[((static_cast<size_t>(tup0 + tl) * n_zl + zi) * nd_max + d) * nb + kk] = hacc[i];
    }
  }
}

// partial Gram sums of the theta-bin splits of k_window_gram -> split 0, in fixed order
__global__ void k_gsum(double* __restrict__ x, size_t n, size_t stride, int n_split) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double s = x[i];
  for (int k = 1; k < n_split; ++k) s += x[i + k * stride];
  x[i] = s;
}

// ------------------------------------------------------------------------------------------
// k_quad: one thread per (tuple of the batch, amplitude point).  Decodes the amplitude values (grid box or explicit
// points), forms the coefficient vector per z and evaluates the quadratic form.
// ------------------------------------------------------------------------------------------
struct QuadArgs {
  int P, n_z, nb, grid_mode;
  int slot[16], zsel[16], is_amp[16];
  // grid box: axis j covers indices ax_start[j] .. ax_start[j] + ax_len[j] - 1
  int ax_start[16], ax_len[16], grid_n[16];
  long long ax_stride[16];
  double grid_lo[16], grid_hi[16], grid_step[16];
  long long out_first;          // flat grid index of chi2[0]
  long long n_amp;              // amplitude points per tuple (grid: product of the amplitude-axis lengths; points mode: B)
  long long t0;                 // first tuple of this batch within the box's tuple enumeration
  int nt;                       // tuples in this batch
  int nd_max;
  const double* points;         // points mode: [B][P]
  // explicit points grouped by tuple on the host (a meshgrid handed in as a point list): tuple t of the call owns the
  // points perm[t_start[t] .. t_start[t + 1]) of the caller's array
  const int* perm;              // [B] or null
  const long long* t_start;     // [tuples of the call + 1]
  const int* data_idx;          // [B] or null (indexed like chi2)
  double* chi2;
  const double* G;              // [nt][n_z][nb][nb]
  const double* g;              // [nt][n_z][nd_max][nb]
};

// FLAGS = CPX_INCLUDE_{HCD, METAL, SKY} of the evaluated z bins (one layout for all of them): the basis count is a
// compile-time constant, the coefficient vector lives in registers.  grid = (amplitude points / 512, tuples of the batch):
// a block's threads share their tuple, whose Gram matrix (and g, when there is one data vector) is staged in shared
// memory z by z and read as broadcasts; every thread scores two points, and only the upper triangle of G is read -- the
// kernel is bound by those shared-memory broadcasts (one per multiply-add), not by the FP64 pipe.
constexpr int kQuadPts = 2;

__device__ __forceinline__ long long quad_decode(const QuadArgs& q, int tl, long long r, double* x) {
  if (!q.grid_mode) {
    if (q.perm) r = q.perm[q.t_start[q.t0 + tl] + r];
    for (int j = 0; j < q.P; ++j) x[j] = q.points[r * q.P + j];
    return r;
  }
  long long tr = q.t0 + tl, flat = 0;
  for (int j = q.P - 1; j >= 0; --j) {
    int ij;
    if (q.is_amp[j]) {
      ij = q.ax_start[j] + static_cast<int>(r % q.ax_len[j]);
      r /= q.ax_len[j];
    } else {
      ij = q.ax_start[j] + static_cast<int>(tr % q.ax_len[j]);
      tr /= q.ax_len[j];
    }
    flat += ij * q.ax_stride[j];
    const int n = q.grid_n[j];
    // bit-identical to numpy.linspace (k_params)
    x[j] = (ij == n - 1 && n > 1) ? q.grid_hi[j] : __dadd_rn(__dmul_rn(static_cast<double>(ij), q.grid_step[j]), q.grid_lo[j]);
  }
  return flat - q.out_first;
}

template <int FLAGS>
__device__ __forceinline__ void quad_coefs(const QuadArgs& q, const ZDev& Z, int zglobal, const double* x, double* c) {
  double b = Z.defaults[0], be = Z.defaults[1], h = Z.defaults[8], eta = Z.defaults[9];
  double bx = Z.defaults[11], xi = Z.defaults[12], bn = Z.defaults[13];
  for (int j = 0; j < q.P; ++j) {
    if (!q.is_amp[j] || !(q.zsel[j] < 0 || q.zsel[j] == zglobal)) continue;
    const double v = x[j];
    switch (q.slot[j]) {
      case 0: b = v; break;
      case 1: be = v; break;
      case 8: h = v; break;
      case 9: eta = v; break;
      case 11: bx = v; break;
      case 12: xi = v; break;
      case 13: bn = v; break;
    }
  }
  int k = 0;
  c[k++] = b * b;
  c[k++] = 2.0 * b * b * be;
  c[k++] = b * b * be * be;
  if (FLAGS & 1) {
    c[k++] = 2.0 * b * h;
    c[k++] = 2.0 * b * h * (be + eta);
    c[k++] = 2.0 * b * h * be * eta;
    c[k++] = h * h;
    c[k++] = 2.0 * h * h * eta;
    c[k++] = h * h * eta * eta;
  }
  if (FLAGS & 2) {
    c[k++] = bx * bx;
    c[k++] = 2.0 * bx * bx * xi;
    c[k++] = bx * bx * xi * xi;
    c[k++] = b * bx;
    c[k++] = b * bx * (be + xi);
    c[k++] = b * bx * be * xi;
  }
  if (FLAGS & 4) c[k++] = bn;
  (void)bn;
}

template <int FLAGS>
__global__ void __launch_bounds__(256) k_quad(QuadArgs q, const ZDev* __restrict__ ztab, const int* __restrict__ zlist) {
  constexpr int NB = basis_count(FLAGS);
  __shared__ double sG[NB * NB + NB];
  const int tl = blockIdx.y;
  double x[kQuadPts][16];
  long long out[kQuadPts];
  bool live[kQuadPts];
  int d[kQuadPts];
  double chi2[kQuadPts];
#pragma unroll
  for (int p = 0; p < kQuadPts; ++p) {
    const long long r = (static_cast<long long>(blockIdx.x) * kQuadPts + p) * blockDim.x + threadIdx.x;
    live[p] = r < (q.perm ? q.t_start[q.t0 + tl + 1] - q.t_start[q.t0 + tl] : q.n_amp);
    out[p] = live[p] ? quad_decode(q, tl, r, x[p]) : 0;
    d[p] = (live[p] && q.data_idx) ? q.data_idx[out[p]] : 0;
    chi2[p] = 0.0;
  }
  const bool one_data = q.nd_max == 1;
  for (int zi = 0; zi < q.n_z; ++zi) {
    const ZDev& Z = ztab[zi];
    const double* Gz = q.G + (static_cast<size_t>(tl) * q.n_z + zi) * NB * NB;
    const double* gz = q.g + (static_cast<size_t>(tl) * q.n_z + zi) * q.nd_max * NB;
    __syncthreads();
    for (int i = threadIdx.x; i < NB * NB + (one_data ? NB : 0); i += blockDim.x) sG[i] = i < NB * NB ? Gz[i] : gz[i - NB * NB];
    __syncthreads();
    const int zglobal = zlist[zi];
    double c[kQuadPts][NB], row[kQuadPts];
#pragma unroll
    for (int p = 0; p < kQuadPts; ++p) quad_coefs<FLAGS>(q, Z, zglobal, x[p], c[p]);
    double s[kQuadPts];
#pragma unroll
    for (int p = 0; p < kQuadPts; ++p) s[p] = Z.ydd[d[p]];
    // chi2 = dd + sum_k c_k (-2 g_k + G_kk c_k + 2 sum_{l > k} G_kl c_l)
#pragma unroll
    for (int kk = 0; kk < NB; ++kk) {
      const double gkk = sG[kk * NB + kk];
#pragma unroll
      for (int p = 0; p < kQuadPts; ++p) {
        const double gk = one_data ? sG[NB * NB + kk] : gz[static_cast<size_t>(d[p]) * NB + kk];
        row[p] = fma(gkk, c[p][kk], -2.0 * gk);
      }
#pragma unroll
      for (int l = kk + 1; l < NB; ++l) {
        const double g2 = 2.0 * sG[kk * NB + l];
#pragma unroll
        for (int p = 0; p < kQuadPts; ++p) row[p] = fma(g2, c[p][l], row[p]);
      }
#pragma unroll
      for (int p = 0; p < kQuadPts; ++p) s[p] = fma(c[p][kk], row[p], s[p]);
    }
#pragma unroll
    for (int p = 0; p < kQuadPts; ++p) chi2[p] += s[p];
  }
#pragma unroll
  for (int p = 0; p < kQuadPts; ++p)
    if (live[p]) q.chi2[out[p]] = chi2[p];
}

}  // namespace cpx

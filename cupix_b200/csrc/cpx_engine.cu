// cupix_b200 host engine + C ABI (include/cupix_b200.h).  Compiled with
//   nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC
// There is deliberately no CPU fallback: every entry point that computes needs a CUDA device.
#include "../../include/cupix_b200.h"

#include <cuda.h>   // CUtensorMap types only; cuTensorMapEncodeTiled is fetched through cudaGetDriverEntryPoint

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

#include "cpx_kernels.cuh"
#include "cpx_hankel_tc.cuh"
#include "cpx_hankel_imma.cuh"
#include "cpx_window_tc.cuh"
#include "cpx_factor.cuh"

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}

#define CPX_CUDA(expr)                                                                      \
  do {                                                                                      \
    cudaError_t e_ = (expr);                                                                \
    if (e_ != cudaSuccess)                                                                  \
      return fail(CPX_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_));        \
  } while (0)

template <typename T>
int upload(const std::vector<T>& h, T** d, size_t* bytes_total) {
  *d = nullptr;
  if (h.empty()) return CPX_OK;
  CPX_CUDA(cudaMalloc(reinterpret_cast<void**>(d), h.size() * sizeof(T)));
  CPX_CUDA(cudaMemcpy(*d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  *bytes_total += h.size() * sizeof(T);
  return CPX_OK;
}

int round_up(int x, int m) { return (x + m - 1) / m * m; }

struct ZHost {
  cpx::ZDev dev;                 // device pointers + dims (mirrored into the device table)
  std::vector<double> Linv;      // [n_A][n_M][n_M] inverse Cholesky factors (kept for cpx_set_data)
  std::vector<double> Lfac;      // [n_A][n_M][n_M] Cholesky factors (cpx_generate_mocks returns mocks in data space)
  std::vector<void*> owned;      // device allocations of this z
};

}  // namespace

struct cpx_handle {
  int device = 0;
  int sm_count = 0;
  cudaStream_t stream = nullptr;
  std::vector<ZHost> z;
  cpx::ZDev* d_ztab = nullptr;       // table of the z bins selected for the current launch
  int* d_zlist = nullptr;
  int* d_nA_of_z = nullptr;
  int na_blk = 0;                    // a-block width shared by all z (template NA)
  int max_na = 0, max_nm = 0, max_nA = 0, max_nM = 0, max_nq = 0;
  int chunk = 0;                     // points the per-chunk scratch is currently sized for (grows on demand)
  int chunk_max = 0;                 // upper limit: as many points as fit CPX_SCRATCH_MB of Px_Zam, at most 65536
  size_t device_bytes = 0;
  long long launches = 0;
  int wt_nb = 0;                     // > 0: k_window_tc (tcgen05 window contraction) for the chi2 stage
  int tc_np = 0;                     // > 0: k_p3d_hankel_tc (tcgen05 Hankel contraction) with this many cell pairs per lane
  int im_nks = 0;                    // > 0: k_p3d_hankel_imma (mma.sync INT8 Hankel contraction) instantiated for this n_q4 (all z alike)
  // per-chunk scratch
  double* d_points = nullptr;        // [chunk][16]
  int* d_data_idx = nullptr;         // [chunk]
  double* d_chi2 = nullptr;          // [chunk]
  double* d_chi2_zA = nullptr;       // [chunk][n_z][max_nA]
  double* d_model = nullptr;         // [chunk][n_z][max_nA][max_nM]  (lazy)
  double* d_pxout = nullptr;         // [chunk][n_z][max_na][max_nm]  (lazy)
  std::vector<void*> owned;
  std::vector<void*> chunk_owned;    // the per-chunk scratch buffers (freed and re-allocated when the chunk grows)
  size_t chunk_bytes = 0;
  void* encode_tmap = nullptr;       // cuTensorMapEncodeTiled, when k_window_tc is available
  void* d_tmaps = nullptr;           // CUtensorMap[n_z] of the Px_Zam scratch
  // factorised path (cpx_factor.cuh): Gram matrices, tuple points, whole-batch staging in host-pointer mode
  double* d_G = nullptr;
  double* d_g = nullptr;
  double* d_tpts = nullptr;
  size_t cap_G = 0, cap_g = 0, cap_tpts = 0;
  double* d_big_pts = nullptr;
  double* d_big_chi2 = nullptr;
  int* d_big_didx = nullptr;
  int* d_perm = nullptr;
  long long* d_tstart = nullptr;
  size_t cap_big_pts = 0, cap_big_chi2 = 0, cap_big_didx = 0, cap_perm = 0, cap_tstart = 0;
  long long factored_calls = 0, factored_tuples = 0, factored_cache_hits = 0;
  // single-tuple calls (amplitude-only batches of samplers and minimisers, bias x beta scans) repeat the same tuple call
  // after call: the Gram data of the last one is kept and re-used while its signature (tuple values, columns, z list,
  // data epoch) matches -- such a call is then one k_quad launch
  std::vector<double> basis_sig;
  long long data_epoch = 0;
  // calls on different streams share the scratch: each call records ev_done, the next one on another stream waits for it
  cudaEvent_t ev_done = nullptr;
  cudaStream_t last_stream = nullptr;
  bool have_last = false;
  // CPX_GUARD=1: every writable scratch buffer sits between two bands of kGuardBytes filled with kGuardByte;
  // cpx_get_info(CPX_INFO_GUARD_VIOLATIONS) counts the band bytes that changed (the GPU pool has no compute-sanitizer)
  bool guard = false;
  struct Guarded { unsigned char* base; size_t bytes; bool per_chunk; };
  std::vector<Guarded> guarded;      // (start of the front band, payload bytes)
};

namespace {
constexpr size_t kGuardBytes = 65536;
constexpr int kGuardByte = 0xA5;

// writable device buffer owned by the handle; with h->guard the payload is fenced by two guard bands
int scratch_alloc(cpx_handle* h, void** p, size_t bytes, bool per_chunk = false) {
  std::vector<void*>& owned = per_chunk ? h->chunk_owned : h->owned;
  if (!h->guard) {
    CPX_CUDA(cudaMalloc(p, bytes));
    owned.push_back(*p);
    h->device_bytes += bytes;
    if (per_chunk) h->chunk_bytes += bytes;
    return CPX_OK;
  }
  const size_t payload = (bytes + 255) / 256 * 256;         // keeps the rear band (and the payload) 256-byte aligned
  unsigned char* base = nullptr;
  CPX_CUDA(cudaMalloc(reinterpret_cast<void**>(&base), payload + 2 * kGuardBytes));
  owned.push_back(base);
  h->device_bytes += payload + 2 * kGuardBytes;
  if (per_chunk) h->chunk_bytes += payload + 2 * kGuardBytes;
  CPX_CUDA(cudaMemset(base, kGuardByte, kGuardBytes));
  CPX_CUDA(cudaMemset(base + kGuardBytes + bytes, kGuardByte, payload - bytes + kGuardBytes));
  h->guarded.push_back({base, bytes, per_chunk});
  *p = base + kGuardBytes;
  return CPX_OK;
}

// buffer of the factorised path that only ever grows (contents need not survive)
template <typename T>
int grow(cpx_handle* h, T** p, size_t* cap, size_t n, bool holds_cache = false) {
  if (n <= *cap && *p) return CPX_OK;
  if (*p) {
    CPX_CUDA(cudaDeviceSynchronize());
    CPX_CUDA(cudaFree(*p));
    h->owned.erase(std::remove(h->owned.begin(), h->owned.end(), static_cast<void*>(*p)), h->owned.end());
    h->device_bytes -= *cap * sizeof(T);
    *p = nullptr;
  }
  if (holds_cache) h->basis_sig.clear();   // the cached Gram data lives in the buffer being replaced
  const size_t want = std::max(n, *cap + *cap / 2);
  CPX_CUDA(cudaMalloc(reinterpret_cast<void**>(p), want * sizeof(T)));
  h->owned.push_back(*p);
  h->device_bytes += want * sizeof(T);
  *cap = want;
  return CPX_OK;
}


// ---- dense helpers (float64, tiny sizes) ---------------------------------------------------
bool cholesky_lower(const double* C, int n, std::vector<double>& L) {
  L.assign(static_cast<size_t>(n) * n, 0.0);
  for (int j = 0; j < n; ++j) {
    double s = C[j * n + j];
    for (int k = 0; k < j; ++k) s -= L[j * n + k] * L[j * n + k];
    if (!(s > 0.0)) return false;
    const double d = std::sqrt(s);
    L[j * n + j] = d;
    for (int i = j + 1; i < n; ++i) {
      double t = C[i * n + j];
      for (int k = 0; k < j; ++k) t -= L[i * n + k] * L[j * n + k];
      L[i * n + j] = t / d;
    }
  }
  return true;
}

void invert_lower(const std::vector<double>& L, int n, double* Li) {
  std::fill(Li, Li + static_cast<size_t>(n) * n, 0.0);
  for (int c = 0; c < n; ++c) {
    Li[c * n + c] = 1.0 / L[c * n + c];
    for (int r = c + 1; r < n; ++r) {
      double s = 0.0;
      for (int k = c; k < r; ++k) s += L[r * n + k] * Li[k * n + c];
      Li[r * n + c] = -s / L[r * n + r];
    }
  }
}

int build_zbin(cpx_handle* h, const cpx_zbin_desc& d, int na_blk, ZHost& zh) {
  using namespace cpx;
  ZDev& Z = zh.dev;
  std::memset(&Z, 0, sizeof(Z));
  Z.n_a = d.n_a;
  Z.n_m = d.n_m;
  Z.n_q = d.n_q;
  Z.n_tiles = (d.n_m + kTileRows - 1) / kTileRows;
  Z.n_m_pad = Z.n_tiles * kTileRows;
  Z.na_blk = na_blk;
  Z.n_ablk = (d.n_a + na_blk - 1) / na_blk;
  Z.n_q4 = (d.n_q + 3) / 4;
  Z.flags = d.flags;
  Z.dAA = d.dAA_dMpc;
  std::memcpy(Z.defaults, d.defaults, sizeof(Z.defaults));
  Z.k_pivot = (d.k_pivot_Mpc > 0.0 && d.defaults[CPX_AS] > 0.0) ? d.k_pivot_Mpc : 0.0;
  size_t& bytes = h->device_bytes;
  auto keep = [&](void* p) { if (p) zh.owned.push_back(p); };

  // cell constants in DMMA A-fragment order, m-tile pairs packed: [tile][ks][6][pair][lane = 4*g + t][2]
  {
    std::vector<float> cell(static_cast<size_t>(Z.n_tiles) * Z.n_q4 * kCellConsts * 4 * 32, 0.0f);
    std::vector<double> cell64(cell.size(), 0.0);   // [tile][ks][6][mt][lane], unpaired, for CPX_PRECISE_CELLS
    const double two_pi2 = 2.0 * M_PI * M_PI;
    for (int m = 0; m < d.n_m; ++m) {
      const int tile = m / kTileRows, r = m % kTileRows, mt = r / 8, g = r % 8;
      const double kp = d.kpar_Mpc[m];
      for (int q = 0; q < d.n_q; ++q) {
        const int ks = q / 4, t = q % 4;
        const double kt = d.kperp_Mpc[q];
        const double k2 = kp * kp + kt * kt;
        float* c = &cell[((((static_cast<size_t>(tile) * Z.n_q4 + ks) * kCellConsts) * 2 + mt / 2) * 32 + 4 * g + t) * 2 + (mt & 1)];
        if (!(k2 > 0.0)) continue;   // k == 0: P_L(0) = 0, cell contributes nothing
        const double k = std::sqrt(k2);
        const double mu = kp / k;
        const double PL = d.linP_cell[static_cast<size_t>(m) * d.n_q + q];
        const size_t cs = 4 * 32;    // stride between the six constants
        c[0 * cs] = static_cast<float>(k * k2 * PL / two_pi2);            // Delta^2(k)
        c[1 * cs] = static_cast<float>(std::log2(k));
        c[2 * cs] = (mu > 0.0) ? static_cast<float>(std::log2(mu)) : -1.0e30f;   // mu = 0: mu^bv -> 0 (bv>0) or 1 (bv=0)
        c[3 * cs] = static_cast<float>(mu * mu);
        c[4 * cs] = static_cast<float>(-k2);
        c[5 * cs] = static_cast<float>(PL);
        double* c64 = &cell64[(((static_cast<size_t>(tile) * Z.n_q4 + ks) * kCellConsts) * 4 + mt) * 32 + 4 * g + t];
        c64[0 * cs] = k * k2 * PL / two_pi2;
        c64[1 * cs] = std::log2(k);
        c64[2 * cs] = (mu > 0.0) ? std::log2(mu) : -1.0e300;
        c64[3 * cs] = mu * mu;
        c64[4 * cs] = k2;
        c64[5 * cs] = PL;
      }
    }
    float* dptr;
    int rc = upload(cell, &dptr, &bytes);
    if (rc) return rc;
    Z.cellc = dptr;
    keep(dptr);
    double* dptr64;
    rc = upload(cell64, &dptr64, &bytes);
    if (rc) return rc;
    Z.cell64 = dptr64;
    keep(dptr64);
  }
  // Hankel weights in DMMA B-fragment order [ablk][ks][j][lane = 4*g + t] = W[a = ablk*na_blk + 8j + g][q = 4ks + t]
  {
    const int nt = na_blk / 8;
    std::vector<double> W(static_cast<size_t>(Z.n_ablk) * Z.n_q4 * nt * 32, 0.0);
    for (int a = 0; a < d.n_a; ++a)
      for (int q = 0; q < d.n_q; ++q) {
        const int ablk = a / na_blk, j = (a % na_blk) / 8, g = a % 8, ks = q / 4, t = q % 4;
        W[((static_cast<size_t>(ablk) * Z.n_q4 + ks) * nt + j) * 32 + 4 * g + t] = d.hankel_w[static_cast<size_t>(a) * d.n_q + q];
      }
    double* dptr;
    int rc = upload(W, &dptr, &bytes);
    if (rc) return rc;
    Z.Wd = dptr;
    keep(dptr);
  }
  // ---- tensor-core Hankel path (cpx_hankel_tc.cuh): scaled cell constants, weight digits, column scales ----
  Z.tc_np = 0;
  if (d.n_q <= 96) {
    const int np = d.n_q <= 64 ? 8 : (d.n_q <= 88 ? 11 : 12);     // packed cell pairs per lane (4 lanes per k_par row)
    const int KP = tc_kp(np), QS = tc_qs(np);
    Z.tc_np = np;
    Z.tc_nt = Z.n_m_pad / kTcRows;
    Z.tc_nablk = (d.n_a + kTcNB - 1) / kTcNB;
    // node q -> (lane quarter, pair, component) and K index of the operand tiles
    auto k_index = [&](int q) { return (q / (2 * np)) * (KP / kTcQuarters) + q % (2 * np); };
    // s_q: power of two nearest to max_a |W[a][q]|, folded into P_L so that every node enters the fixed-point
    // conversion with the magnitude of its largest contribution to any theta bin
    std::vector<double> sq(d.n_q, 1.0);
    for (int q = 0; q < d.n_q; ++q) {
      double mx = 0.0;
      for (int a = 0; a < d.n_a; ++a) mx = std::max(mx, std::fabs(d.hankel_w[static_cast<size_t>(a) * d.n_q + q]));
      if (mx > 0.0) sq[q] = std::ldexp(1.0, std::max(-100, std::min(100, static_cast<int>(std::lround(std::log2(mx))))));
    }
    // [tile][3 constant pairs (D, log2 k), (log2 mu, mu^2), (-k^2, P_L s_q)][quarter][QS][row][4]: one LDS.128 per pair
    const size_t cs = static_cast<size_t>(kTcQuarters) * QS * kTcRows * 4;     // float stride between the constant pairs
    std::vector<float> cell(static_cast<size_t>(Z.tc_nt) * 3 * cs, 0.0f);
    const double two_pi2 = 2.0 * M_PI * M_PI;
    for (int m = 0; m < d.n_m; ++m) {
      const int t8 = m / kTcRows, row = m % kTcRows;
      const double kp = d.kpar_Mpc[m];
      for (int q = 0; q < d.n_q; ++q) {
        const double kt = d.kperp_Mpc[q];
        const double k2 = kp * kp + kt * kt;
        if (!(k2 > 0.0)) continue;
        const int quarter = q / (2 * np), pr = (q % (2 * np)) / 2, comp = q % 2;
        const double k = std::sqrt(k2), mu = kp / k;
        const double PL = d.linP_cell[static_cast<size_t>(m) * d.n_q + q];
        float* c = &cell[static_cast<size_t>(t8) * 3 * cs + ((static_cast<size_t>(quarter) * QS + pr) * kTcRows + row) * 4 + comp];
        c[0 * cs + 0] = static_cast<float>(k * k2 * PL / two_pi2);
        c[0 * cs + 2] = static_cast<float>(std::log2(k));
        c[1 * cs + 0] = (mu > 0.0) ? static_cast<float>(std::log2(mu)) : -1.0e30f;
        c[1 * cs + 2] = static_cast<float>(mu * mu);
        c[2 * cs + 0] = static_cast<float>(-k2);
        c[2 * cs + 2] = static_cast<float>(PL) * static_cast<float>(sq[q]);      // power-of-two scale: exact
      }
    }
    float* dcell;
    int rc = upload(cell, &dcell, &bytes);
    if (rc) return rc;
    Z.tc_cell = dcell;
    keep(dcell);
    std::vector<double> ta(static_cast<size_t>(Z.tc_nablk) * kTcNB, 1.0);
    std::vector<signed char> dig(static_cast<size_t>(Z.tc_nablk) * kTcBRows * KP, 0);
    for (int a = 0; a < d.n_a; ++a) {
      double mx = 0.0;
      for (int q = 0; q < d.n_q; ++q) mx = std::max(mx, std::fabs(d.hankel_w[static_cast<size_t>(a) * d.n_q + q] / sq[q]));
      const double t = mx > 0.0 ? 2.05 * mx : 1.0;
      ta[a] = t;
      const int ablk = a / kTcNB, al = a % kTcNB;
      for (int q = 0; q < d.n_q; ++q) {
        // balanced base-256 digits, least significant first: X = sum_j dgt_j 256^(NW-1-j)
        long long X = std::llrint(d.hankel_w[static_cast<size_t>(a) * d.n_q + q] / sq[q] / t * 1099511627776.0);   // 256^5
        for (int j = kTcNW - 1; j >= 0; --j) {
          long long dg = ((X + 128) % 256 + 256) % 256 - 128;
          X = (X - dg) / 256;
          dig[static_cast<size_t>(ablk) * kTcBRows * KP + tc_canon(kTcNB * j + al, k_index(q), kTcBRows / 8)] = static_cast<signed char>(dg);
        }
      }
    }
    signed char* ddig;
    rc = upload(dig, &ddig, &bytes);
    if (rc) return rc;
    Z.tc_wdig = ddig;
    keep(ddig);
    double* dta;
    rc = upload(ta, &dta, &bytes);
    if (rc) return rc;
    Z.tc_ta = dta;
    keep(dta);
  }
  // ---- INT8 mma.sync Hankel path (cpx_hankel_imma.cuh): cell constants with P_L s_q, weight digits in B-fragment order ----
  Z.im_cell = nullptr;
  Z.im_wdig = nullptr;
  Z.im_ta = nullptr;
  if (h->im_nks > 0 && Z.n_q4 == h->im_nks) {
    const int NKB = (h->im_nks + 7) / 8, nt = na_blk / 8;
    std::vector<double> sq(d.n_q, 1.0);
    for (int q = 0; q < d.n_q; ++q) {
      double mx = 0.0;
      for (int a = 0; a < d.n_a; ++a) mx = std::max(mx, std::fabs(d.hankel_w[static_cast<size_t>(a) * d.n_q + q]));
      if (mx > 0.0) sq[q] = std::ldexp(1.0, std::max(-100, std::min(100, static_cast<int>(std::lround(std::log2(mx))))));
    }
    // cellc is still on the host side of this function only as `d`: rebuild the P_L entries scaled by s_q (exact)
    // the P_L entries scaled by s_q (exact), constants interleaved in pairs: [tile][ks][3][half][lane][(c even).xy, (c odd).xy]
    std::vector<float> src(static_cast<size_t>(Z.n_tiles) * Z.n_q4 * kCellConsts * 4 * 32, 0.0f), cell(src.size(), 0.0f);
    CPX_CUDA(cudaMemcpy(src.data(), Z.cellc, src.size() * sizeof(float), cudaMemcpyDeviceToHost));
    for (int tile = 0; tile < Z.n_tiles; ++tile)
      for (int ks = 0; ks < Z.n_q4; ++ks)
        for (int cst = 0; cst < kCellConsts; ++cst)
          for (int pr = 0; pr < 2; ++pr)
            for (int lane = 0; lane < 32; ++lane)
              for (int e = 0; e < 2; ++e) {
                const int q = 4 * ks + (lane & 3);
                float v = src[(((((static_cast<size_t>(tile) * Z.n_q4 + ks) * kCellConsts + cst) * 2 + pr) * 32 + lane) * 2) + e];
                if (cst == 5 && q < d.n_q) v *= static_cast<float>(sq[q]);
                cell[(((((static_cast<size_t>(tile) * Z.n_q4 + ks) * 3 + cst / 2) * 2 + pr) * 32 + lane) * 4) + (cst & 1) * 2 + e] = v;
              }
    float* dcell;
    int rc = upload(cell, &dcell, &bytes);
    if (rc) return rc;
    Z.im_cell = dcell;
    keep(dcell);
    std::vector<int> ta(static_cast<size_t>(Z.n_ablk) * na_blk, 0);          // k_a: column scale t_a = 2^k_a
    const size_t dig_bytes = static_cast<size_t>(imma_dig_bytes(NKB, nt));
    std::vector<unsigned char> dig(static_cast<size_t>(Z.n_ablk) * dig_bytes, 0);
    const double full = std::ldexp(1.0, 8 * kImNW);
    for (int a = 0; a < d.n_a; ++a) {
      double mx = 0.0;
      for (int q = 0; q < d.n_q; ++q) mx = std::max(mx, std::fabs(d.hankel_w[static_cast<size_t>(a) * d.n_q + q] / sq[q]));
      const int ka = mx > 0.0 ? std::max(-200, std::min(200, static_cast<int>(std::ceil(std::log2(2.05 * mx))))) : 0;
      const double t = std::ldexp(1.0, ka);
      ta[a] = ka;
      const int ablk = a / na_blk, jn = (a % na_blk) / 8, g = a % 8;
      for (int q = 0; q < d.n_q; ++q) {
        // K slot of node q: lane t = q % 4, ks = q / 4 = 8 b + 4 half + byte
        const int tq = q % 4, ks = q / 4, b = ks / 8, half = (ks % 8) / 4, byte = ks % 4;
        long long X = std::llrint(d.hankel_w[static_cast<size_t>(a) * d.n_q + q] / sq[q] / t * full);
        for (int j = kImNW - 1; j >= 0; --j) {          // balanced base-256 digits, least significant first
          long long dg = ((X + 128) % 256 + 256) % 256 - 128;
          X = (X - dg) / 256;
          dig[ablk * dig_bytes + ((((static_cast<size_t>(b) * kImNW + j) * nt + jn) * 32 + 4 * g + tq) * 2 + half) * 4 + byte] =
              static_cast<unsigned char>(static_cast<signed char>(dg));
        }
      }
    }
    unsigned char* ddig;
    rc = upload(dig, &ddig, &bytes);
    if (rc) return rc;
    Z.im_wdig = ddig;
    keep(ddig);
    int* dta;
    rc = upload(ta, &dta, &bytes);
    if (rc) return rc;
    Z.im_ta = dta;
    keep(dta);
  }
  auto up_copy = [&](const double* src, size_t n, const double** dst) -> int {
    *dst = nullptr;
    if (!src) return CPX_OK;
    std::vector<double> v(src, src + n);
    double* dptr;
    int rc = upload(v, &dptr, &bytes);
    if (rc) return rc;
    *dst = dptr;
    keep(dptr);
    return CPX_OK;
  };
  {
    std::vector<double> kp(Z.n_m_pad, 0.0);
    std::copy(d.kpar_Mpc, d.kpar_Mpc + d.n_m, kp.begin());
    double* dptr;
    int rc = upload(kp, &dptr, &bytes);
    if (rc) return rc;
    Z.kpar = dptr;
    Z.kpar_row = dptr;
    keep(dptr);
    if (d.kpar_row_Mpc) {
      std::copy(d.kpar_row_Mpc, d.kpar_row_Mpc + d.n_m, kp.begin());
      double* drow;
      rc = upload(kp, &drow, &bytes);
      if (rc) return rc;
      Z.kpar_row = drow;
      keep(drow);
    }
  }
  int rc;
  if (d.flags & CPX_INCLUDE_METAL) {
    if (!d.metal_auto || !d.metal_cross) return fail(CPX_ERR_ARG, "include_metal set but metal tables are NULL");
    if ((rc = up_copy(d.metal_auto, 3ull * d.n_a * d.n_m, &Z.metal_auto))) return rc;
    if ((rc = up_copy(d.metal_cross, 3ull * d.n_a * d.n_m, &Z.metal_cross))) return rc;
    if (d.metal_auto_dn && d.metal_cross_dn && d.metal_dn_order > 0) {
      const size_t n = static_cast<size_t>(d.metal_dn_order + 1) * 3ull * d.n_a * d.n_m;
      if ((rc = up_copy(d.metal_auto_dn, n, &Z.metal_auto_dn))) return rc;
      if ((rc = up_copy(d.metal_cross_dn, n, &Z.metal_cross_dn))) return rc;
      Z.metal_dn_order = d.metal_dn_order;
    }
  }
  if (d.flags & CPX_INCLUDE_SKY) {
    if (!d.sky_xi_a || !d.sky_smooth_m) return fail(CPX_ERR_ARG, "include_sky set but sky tables are NULL");
    if ((rc = up_copy(d.sky_xi_a, d.n_a, &Z.sky_xi))) return rc;
    if ((rc = up_copy(d.sky_smooth_m, d.n_m, &Z.sky_smooth))) return rc;
  }

  // ---- window / covariance operators ----
  Z.n_A = d.n_A;
  if (d.n_A > 0) {
    if (!d.B_A_a || !d.U_aMn || !d.V_aM || !d.Px_AM || !d.cov_AMM || d.n_M <= 0 || d.n_data <= 0)
      return fail(CPX_ERR_ARG, "measurement arrays missing for a z bin with n_A > 0");
    const int nA = d.n_A, nM = d.n_M, na = d.n_a, nm = d.n_m;
    Z.n_M = nM;
    {   // coarse-k columns in n_mblk blocks of 8*MT (MT <= 8 n-tiles per k_window pass)
      const int tiles = (nM + 7) / 8;
      Z.n_mblk = (tiles + 7) / 8;
      const int mt = (tiles + Z.n_mblk - 1) / Z.n_mblk;
      Z.n_M_pad = Z.n_mblk * mt * 8;
    }
    Z.n_data = d.n_data;
    std::vector<int> pair_start(nA + 1, 0), pair_a;
    for (int A = 0; A < nA; ++A) {
      for (int a = 0; a < na; ++a)
        if (d.B_A_a[static_cast<size_t>(A) * na + a] != 0.0) pair_a.push_back(a);
      pair_start[A + 1] = static_cast<int>(pair_a.size());
      if (pair_start[A + 1] == pair_start[A])
        return fail(CPX_ERR_ARG, "coarse theta bin without fine bins in B_A_a");
    }
    Z.n_pairs = static_cast<int>(pair_a.size());
    const size_t tstride = static_cast<size_t>(Z.n_M_pad) * Z.n_m_pad;
    std::vector<double> Tm(Z.n_pairs * tstride, 0.0), Tc(Z.n_pairs * tstride, 0.0);
    zh.Linv.assign(static_cast<size_t>(nA) * nM * nM, 0.0);
    zh.Lfac.assign(static_cast<size_t>(nA) * nM * nM, 0.0);
    std::vector<double> L, vsum(nM);
    for (int A = 0; A < nA; ++A) {
      if (!cholesky_lower(d.cov_AMM + static_cast<size_t>(A) * nM * nM, nM, L)) {
        char buf[128];
        std::snprintf(buf, sizeof buf, "covariance block of coarse theta bin %d is not positive definite", A);
        return fail(CPX_ERR_COV, buf);
      }
      double* Li = &zh.Linv[static_cast<size_t>(A) * nM * nM];
      invert_lower(L, nM, Li);
      std::copy(L.begin(), L.end(), zh.Lfac.begin() + static_cast<size_t>(A) * nM * nM);
      std::fill(vsum.begin(), vsum.end(), 0.0);
      for (int p = pair_start[A]; p < pair_start[A + 1]; ++p)
        for (int M = 0; M < nM; ++M) vsum[M] += d.V_aM[static_cast<size_t>(pair_a[p]) * nM + M];
      for (int p = pair_start[A]; p < pair_start[A + 1]; ++p) {
        const int a = pair_a[p];
        double* tm = &Tm[p * tstride];
        double* tc = &Tc[p * tstride];
        for (int M = 0; M < nM; ++M) {
          const double wgt = d.V_aM[static_cast<size_t>(a) * nM + M] / vsum[M];
          const double* u = d.U_aMn + (static_cast<size_t>(a) * nM + M) * nm;
          for (int m = 0; m < nm; ++m) tm[static_cast<size_t>(M) * Z.n_m_pad + m] = wgt * u[m];
        }
        for (int M = 0; M < nM; ++M)
          for (int Mp = 0; Mp <= M; ++Mp) {
            const double l = Li[M * nM + Mp];
            const double* src = &tm[static_cast<size_t>(Mp) * Z.n_m_pad];
            double* dst = &tc[static_cast<size_t>(M) * Z.n_m_pad];
            for (int m = 0; m < nm; ++m) dst[m] += l * src[m];
          }
      }
    }
    int* dps;
    int* dpa;
    double *dTm, *dTc;
    if ((rc = upload(pair_start, &dps, &bytes))) return rc;
    if ((rc = upload(pair_a, &dpa, &bytes))) return rc;
    if ((rc = upload(Tm, &dTm, &bytes))) return rc;
    if ((rc = upload(Tc, &dTc, &bytes))) return rc;
    Z.pair_start = dps;
    Z.pair_a = dpa;
    Z.T_model = dTm;
    Z.T_chi = dTc;
    keep(dps);
    keep(dpa);
    keep(dTm);
    keep(dTc);

    // ---- tensor-core window path (cpx_window_tc.cuh): column scales, row scales and base-256 digits of T_chi ----
    Z.wt_nb = 0;
    if (Z.n_mblk == 1 && Z.n_M_pad <= 64) {
      const int NB = round_up(nM, 16);
      const int cpp = Z.n_m_pad / kWtKC, bpp = Z.n_m_pad / 16;
      std::vector<int> cs(static_cast<size_t>(Z.n_pairs) * bpp, 0), off(nA, 0);
      std::vector<double> rs(static_cast<size_t>(nA) * NB, 1.0);
      for (int p = 0; p < Z.n_pairs; ++p)
        for (int b = 0; b < bpp; ++b) {
          double mx = 0.0;
          for (int M = 0; M < nM; ++M)
            for (int m = 16 * b; m < 16 * b + 16; ++m) mx = std::max(mx, std::fabs(Tc[p * tstride + static_cast<size_t>(M) * Z.n_m_pad + m]));
          if (mx > 0.0) cs[static_cast<size_t>(p) * bpp + b] = std::max(-300, std::min(300, static_cast<int>(std::lround(std::log2(mx)))));
        }
      int chunks = 0;
      for (int A = 0; A < nA; ++A) {
        off[A] = chunks;
        chunks += (pair_start[A + 1] - pair_start[A]) * cpp;
      }
      const size_t tile = static_cast<size_t>(wt_b_stage(NB));
      std::vector<signed char> dig(static_cast<size_t>(chunks) * tile, 0);
      for (int A = 0; A < nA; ++A) {
        for (int M = 0; M < nM; ++M) {
          double mx = 0.0;
          for (int p = pair_start[A]; p < pair_start[A + 1]; ++p)
            for (int m = 0; m < nm; ++m)
              mx = std::max(mx, std::fabs(std::ldexp(Tc[p * tstride + static_cast<size_t>(M) * Z.n_m_pad + m], -cs[static_cast<size_t>(p) * bpp + m / 16])));
          const double r = mx > 0.0 ? 2.05 * mx : 1.0;
          rs[static_cast<size_t>(A) * NB + M] = r;
          for (int p = pair_start[A]; p < pair_start[A + 1]; ++p)
            for (int m = 0; m < nm; ++m) {
              const double v = std::ldexp(Tc[p * tstride + static_cast<size_t>(M) * Z.n_m_pad + m], -cs[static_cast<size_t>(p) * bpp + m / 16]) / r;
              long long X = std::llrint(v * 1099511627776.0);   // 256^5
              const size_t chunk = static_cast<size_t>(off[A]) + static_cast<size_t>(p - pair_start[A]) * cpp + m / kWtKC;
              for (int j = kWtNW - 1; j >= 0; --j) {
                const long long dg = ((X + 128) % 256 + 256) % 256 - 128;
                X = (X - dg) / 256;
                dig[chunk * tile + tc_canon(NB * j + M, m % kWtKC, kWtNW * NB / 8)] = static_cast<signed char>(dg);
              }
            }
        }
      }
      signed char* ddig;
      int *doff, *dcs;
      double* drs;
      if ((rc = upload(dig, &ddig, &bytes))) return rc;
      if ((rc = upload(off, &doff, &bytes))) return rc;
      if ((rc = upload(cs, &dcs, &bytes))) return rc;
      if ((rc = upload(rs, &drs, &bytes))) return rc;
      Z.wt_nb = NB;
      Z.wt_dig = ddig;
      Z.wt_off = doff;
      Z.wt_cs = dcs;
      Z.wt_rs = drs;
      keep(ddig);
      keep(doff);
      keep(dcs);
      keep(drs);
    }
  }
  return CPX_OK;
}

// whitened data vectors yd[d][A][M_pad] = L_A^-1 d_A
int upload_yd(cpx_handle* h, ZHost& zh, const double* Px_AM, int n_data) {
  cpx::ZDev& Z = zh.dev;
  const int nA = Z.n_A, nM = Z.n_M;
  std::vector<double> yd(static_cast<size_t>(n_data) * nA * Z.n_M_pad, 0.0), ydd(n_data, 0.0);
  for (int dd = 0; dd < n_data; ++dd)
    for (int A = 0; A < nA; ++A) {
      const double* Li = &zh.Linv[static_cast<size_t>(A) * nM * nM];
      const double* x = Px_AM + (static_cast<size_t>(dd) * nA + A) * nM;
      double* y = &yd[(static_cast<size_t>(dd) * nA + A) * Z.n_M_pad];
      for (int M = 0; M < nM; ++M) {
        double s = 0.0;
        for (int Mp = 0; Mp <= M; ++Mp) s += Li[M * nM + Mp] * x[Mp];
        y[M] = s;
        ydd[dd] += s * s;
      }
    }
  for (const double* old : {Z.yd, Z.ydd})
    if (old) {
      cudaFree(const_cast<double*>(old));
      zh.owned.erase(std::remove(zh.owned.begin(), zh.owned.end(), (void*)old), zh.owned.end());
    }
  double *dptr, *dptr2;
  int rc = upload(yd, &dptr, &h->device_bytes);
  if (rc) return rc;
  zh.owned.push_back(dptr);
  if ((rc = upload(ydd, &dptr2, &h->device_bytes))) return rc;
  zh.owned.push_back(dptr2);
  Z.yd = dptr;
  Z.ydd = dptr2;
  Z.n_data = n_data;
  ++h->data_epoch;
  return CPX_OK;
}

// ---- kernel dispatch -----------------------------------------------------------------------
typedef void (*K1Fn)(const cpx::ZDev*, int, int, long long);

// PP = 1 fits two CTAs per SM (<= 128 registers); PP = 2 runs one CTA per SM.
// ADD: some evaluated z bin has additive contaminants (SiIII metals or sky) in its epilogue.
// COSMO: some column of the call is CPX_AS / CPX_NS (linear power rescaled per point inside the cell loop).
template <bool ADD, bool COSMO>
K1Fn k1_select_add(int pp, int nt) {
  if (pp == 1) {
    switch (nt) {
      case 1: return cpx::k_p3d_hankel<1, 1, 2, ADD, COSMO>;
      case 2: return cpx::k_p3d_hankel<1, 2, 2, ADD, COSMO>;
      case 3: return cpx::k_p3d_hankel<1, 3, 2, ADD, COSMO>;
    }
  } else if (pp == 2) {
    switch (nt) {
      case 1: return cpx::k_p3d_hankel<2, 1, 1, ADD, COSMO>;
      case 2: return cpx::k_p3d_hankel<2, 2, 1, ADD, COSMO>;
      case 3: return cpx::k_p3d_hankel<2, 3, 1, ADD, COSMO>;
    }
  }
  return nullptr;
}
K1Fn k1_select(int pp, int nt, bool add, bool cosmo = false) {
  if (cosmo) return add ? k1_select_add<true, true>(pp, nt) : k1_select_add<false, true>(pp, nt);
  return add ? k1_select_add<true, false>(pp, nt) : k1_select_add<false, false>(pp, nt);
}
int k1_ctas_per_sm(int pp) { return pp == 1 ? 2 : 1; }

K1Fn k1_select_tc(int np, bool add) {
  switch (np) {
    case 8: return add ? cpx::k_p3d_hankel_tc<8, true> : cpx::k_p3d_hankel_tc<8, false>;
    case 11: return add ? cpx::k_p3d_hankel_tc<11, true> : cpx::k_p3d_hankel_tc<11, false>;
    case 12: return add ? cpx::k_p3d_hankel_tc<12, true> : cpx::k_p3d_hankel_tc<12, false>;
  }
  return nullptr;
}

template <int NKS, bool ADD>
K1Fn k1_select_imma_nt(int nt) {
  switch (nt) {
    case 1: return cpx::k_p3d_hankel_imma<NKS, 1, ADD, false>;
    case 2: return cpx::k_p3d_hankel_imma<NKS, 2, ADD, false>;
    case 3: return cpx::k_p3d_hankel_imma<NKS, 3, ADD, false>;
  }
  return nullptr;
}
// n_q4 of the node sets the kernel is instantiated for: 52 (the default of the 'smooth' family), 56 and 88 nodes; calls with
// cosmology columns (CPX_AS / CPX_NS) keep the DMMA kernel (the COSMO instantiation exists in the template, it is left out
// of the build to keep compile time down)
bool imma_supported(int nks) { return nks == 13 || nks == 14 || nks == 22; }
K1Fn k1_select_imma(int nks, int nt, bool add) {
  switch (nks) {
    case 13: return add ? k1_select_imma_nt<13, true>(nt) : k1_select_imma_nt<13, false>(nt);
    case 14: return add ? k1_select_imma_nt<14, true>(nt) : k1_select_imma_nt<14, false>(nt);
    case 22: return add ? k1_select_imma_nt<22, true>(nt) : k1_select_imma_nt<22, false>(nt);
  }
  return nullptr;
}
size_t imma_smem_bytes(int nks, int nt) {
  return static_cast<size_t>(nks) * (cpx::kCellConsts * 4 * 32 * 4) + cpx::imma_dig_bytes((nks + 7) / 8, nt) + 8 * nt * 4;
}

K1Fn k1_select_f64s(int nt) {      // shared-memory staged variant, 16 points per CTA (cpx_factor.cuh)
  switch (nt) {
    case 1: return cpx::k_p3d_hankel_f64s<1, 16>;
    case 2: return cpx::k_p3d_hankel_f64s<2, 16>;
    case 3: return cpx::k_p3d_hankel_f64s<3, 16>;
  }
  return nullptr;
}

K1Fn k1_select_f64(int nt) {
  switch (nt) {
    case 1: return cpx::k_p3d_hankel_f64<1>;
    case 2: return cpx::k_p3d_hankel_f64<2>;
    case 3: return cpx::k_p3d_hankel_f64<3>;
  }
  return nullptr;
}

typedef void (*K2Fn)(const cpx::ZDev*, int, cpx::WindowArgs);
template <bool CHI2>
K2Fn k2_select(int mt) {
  switch (mt) {
    case 1: return cpx::k_window<1, CHI2>;
    case 2: return cpx::k_window<2, CHI2>;
    case 3: return cpx::k_window<3, CHI2>;
    case 4: return cpx::k_window<4, CHI2>;
    case 5: return cpx::k_window<5, CHI2>;
    case 6: return cpx::k_window<6, CHI2>;
    case 7: return cpx::k_window<7, CHI2>;
    case 8: return cpx::k_window<8, CHI2>;
  }
  return nullptr;
}

int choose_na_blk(int max_na) {
  if (max_na <= 24) return round_up(max_na, 8);
  const int nblk = (max_na + 23) / 24;
  return round_up((max_na + nblk - 1) / nblk, 8);
}

int env_int(const char* name, int dflt) {
  const char* s = std::getenv(name);
  return s ? std::atoi(s) : dflt;
}

// Per-chunk scratch (PtPar, RowPar, Px_Zam of every z; points, data indices, chi2, chi2_zA) for at least `need` points
// (clamped to chunk_max).  Grows geometrically; growing synchronises the device, frees the old buffers and re-encodes the
// tensor maps k_window_tc reads Px_Zam through.  The lazily allocated model / Px output buffers are dropped with it.
int ensure_scratch(cpx_handle* h, long long need) {
  need = std::max(1LL, std::min<long long>(need, h->chunk_max));
  if (need <= h->chunk) return CPX_OK;
  long long c = 256;
  while (c < need) c *= 2;
  c = std::min<long long>(c, h->chunk_max);
  CPX_CUDA(cudaDeviceSynchronize());
  for (void* p : h->chunk_owned) cudaFree(p);
  h->chunk_owned.clear();
  h->device_bytes -= h->chunk_bytes;
  h->chunk_bytes = 0;
  h->guarded.erase(std::remove_if(h->guarded.begin(), h->guarded.end(), [](const cpx_handle::Guarded& g) { return g.per_chunk; }),
                   h->guarded.end());
  h->d_model = h->d_pxout = nullptr;
  h->chunk = 0;
  const int n_z = static_cast<int>(h->z.size());
  auto dmalloc = [&](void** p, size_t bytes) -> int { return scratch_alloc(h, p, bytes, true); };
  int rc;
  for (int i = 0; i < n_z; ++i) {
    cpx::ZDev& Z = h->z[i].dev;
    Z.ptpar = nullptr;
    Z.rowpar = nullptr;
    Z.rowcont = nullptr;
    Z.pxzam = nullptr;
    if ((rc = dmalloc(reinterpret_cast<void**>(&Z.ptpar), sizeof(cpx::PtPar) * c))) return rc;
    if (Z.flags & (CPX_INCLUDE_HCD | CPX_INCLUDE_CONTINUUM))
      if ((rc = dmalloc(reinterpret_cast<void**>(&Z.rowpar), sizeof(cpx::RowPar) * static_cast<size_t>(c) * Z.n_m_pad))) return rc;
    // the continuum factor on its own: only the additive terms (SiIII, sky) need it beside RowPar::ampcont
    if ((Z.flags & CPX_INCLUDE_CONTINUUM) && (Z.flags & (CPX_INCLUDE_METAL | CPX_INCLUDE_SKY)))
      if ((rc = dmalloc(reinterpret_cast<void**>(&Z.rowcont), sizeof(double) * static_cast<size_t>(c) * Z.n_m_pad))) return rc;
    if ((rc = dmalloc(reinterpret_cast<void**>(&Z.pxzam), static_cast<size_t>(c) * Z.n_a * Z.n_m_pad * 8 + 256))) return rc;
  }
  if ((rc = dmalloc(reinterpret_cast<void**>(&h->d_points), sizeof(double) * 16 * c))) return rc;
  if ((rc = dmalloc(reinterpret_cast<void**>(&h->d_data_idx), sizeof(int) * c))) return rc;
  if ((rc = dmalloc(reinterpret_cast<void**>(&h->d_chi2), sizeof(double) * c))) return rc;
  if (h->max_nA > 0)
    if ((rc = dmalloc(reinterpret_cast<void**>(&h->d_chi2_zA), sizeof(double) * static_cast<size_t>(c) * n_z * h->max_nA))) return rc;
  if (h->wt_nb > 0) {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    std::vector<CUtensorMap> maps(n_z);
    for (int i = 0; i < n_z; ++i) {
      const cpx::ZDev& Z = h->z[i].dev;
      if (Z.n_A <= 0) continue;
      const cuuint64_t dims[3] = {static_cast<cuuint64_t>(Z.n_m_pad), static_cast<cuuint64_t>(Z.n_a), static_cast<cuuint64_t>(c)};
      const cuuint64_t strides[2] = {static_cast<cuuint64_t>(Z.n_m_pad) * 8, static_cast<cuuint64_t>(Z.n_a) * Z.n_m_pad * 8};
      const cuuint32_t box[3] = {16, 1, static_cast<cuuint32_t>(cpx::kWtPts)};
      const cuuint32_t estr[3] = {1, 1, 1};
      const CUresult r = reinterpret_cast<EncodeFn>(h->encode_tmap)(&maps[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, Z.pxzam, dims, strides, box, estr,
                                                                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                                                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) return fail(CPX_ERR_CUDA, "cuTensorMapEncodeTiled failed for the Px_Zam scratch");
    }
    CPX_CUDA(cudaMemcpy(h->d_tmaps, maps.data(), sizeof(CUtensorMap) * n_z, cudaMemcpyHostToDevice));
    for (int i = 0; i < n_z; ++i) h->z[i].dev.px_tmap = static_cast<CUtensorMap*>(h->d_tmaps) + i;
  }
  h->chunk = static_cast<int>(c);
  return CPX_OK;
}


// ---- factorised path ---------------------------------------------------------------------------
// slots whose parameter enters the model polynomially (cpx_factor.cuh); everything else defines the tuple
bool is_amp_slot(int s) {
  return s == CPX_BIAS || s == CPX_BETA || s == CPX_B_H || s == CPX_BETA_H || s == CPX_B_X || s == CPX_BETA_X || s == CPX_B_NOISE_MPC;
}

struct GridBox {
  int start[16], len[16];
  long long n_tup, n_amp;
};

// the flat index range [first, first + B) of a mixed-radix grid (last axis fastest) as at most 2 P hyper-rectangles
void decompose_range(long long first, long long B, int P, const int32_t* n, const bool* amp, std::vector<GridBox>& boxes) {
  long long stride[16];
  stride[P - 1] = 1;
  for (int j = P - 2; j >= 0; --j) stride[j] = stride[j + 1] * n[j + 1];
  long long lo = first;
  const long long hi = first + B;
  while (lo < hi) {
    int j = P - 1;
    for (int jj = 0; jj < P; ++jj)
      if (lo % stride[jj] == 0 && stride[jj] <= hi - lo) {
        j = jj;
        break;
      }
    const long long idx = (lo / stride[j]) % n[j];
    const long long cnt = std::min<long long>((hi - lo) / stride[j], n[j] - idx);
    GridBox b;
    b.n_tup = b.n_amp = 1;
    for (int i = 0; i < P; ++i) {
      if (i < j) {
        b.start[i] = static_cast<int>((lo / stride[i]) % n[i]);
        b.len[i] = 1;
      } else if (i == j) {
        b.start[i] = static_cast<int>(idx);
        b.len[i] = static_cast<int>(cnt);
      } else {
        b.start[i] = 0;
        b.len[i] = n[i];
      }
      (amp[i] ? b.n_amp : b.n_tup) *= b.len[i];
    }
    boxes.push_back(b);
    lo += cnt * stride[j];
  }
}

typedef void (*KWgFn)(const cpx::ZDev*, int, int, int, int, double*, double*, size_t, size_t);
KWgFn kwg_select(int mt) {
  switch (mt) {
    case 1: return cpx::k_window_gram<1>;
    case 2: return cpx::k_window_gram<2>;
    case 3: return cpx::k_window_gram<3>;
    case 4: return cpx::k_window_gram<4>;
    case 5: return cpx::k_window_gram<5>;
    case 6: return cpx::k_window_gram<6>;
    case 7: return cpx::k_window_gram<7>;
    case 8: return cpx::k_window_gram<8>;
  }
  return nullptr;
}

typedef void (*KBasisFn)(const cpx::ZDev*, int, int, long long, int);
template <int W>
KBasisFn kbasis_select_w(int nt, bool stage) {
  switch (nt) {
    case 1: return stage ? cpx::k_basis_f64<1, true, W> : cpx::k_basis_f64<1, false, W>;
    case 2: return stage ? cpx::k_basis_f64<2, true, W> : cpx::k_basis_f64<2, false, W>;
    case 3: return stage ? cpx::k_basis_f64<3, true, W> : cpx::k_basis_f64<3, false, W>;
  }
  return nullptr;
}
KBasisFn kbasis_select(int nt, bool stage, int warps) {
  return warps == 8 ? kbasis_select_w<8>(nt, stage) : kbasis_select_w<16>(nt, stage);
}

// Explicit host points whose tuple columns take few distinct values -- a meshgrid handed in as a point list, the way the
// reference's scan scripts build theirs (scripts/chi2_scan_mock.py:107-111) -- grouped by tuple: tuple values, the points of
// each tuple (perm, start) and the largest group.  Returns false when the points do not share tuples (decided on a strided
// sample first, so that a batch of unrelated points costs a few milliseconds to rule out).
struct TupleKey {
  double v[11];
  int n;
  bool operator==(const TupleKey& o) const { return n == o.n && std::memcmp(v, o.v, sizeof(double) * n) == 0; }
};
struct TupleKeyHash {
  size_t operator()(const TupleKey& k) const {
    uint64_t h = 1469598103934665603ull;
    for (int i = 0; i < k.n; ++i) {
      uint64_t b;
      std::memcpy(&b, &k.v[i], 8);
      h = (h ^ b) * 1099511628211ull;
      h ^= h >> 29;
    }
    return static_cast<size_t>(h);
  }
};
struct PointGroups {
  std::vector<double> tuple_vals;      // [T][n_tcol]
  std::vector<long long> start;        // [T + 1]
  std::vector<int> perm;               // [B]
  long long max_count = 0;
};

bool group_points(const double* pts, int64_t B, int P, const bool* amp, int n_tcol, long long min_ratio, PointGroups& g) {
  if (n_tcol > 11 || B >= (1LL << 31)) return false;
  int tcol[16], nt = 0;
  for (int j = 0; j < P; ++j)
    if (!amp[j]) tcol[nt++] = j;
  auto key_of = [&](int64_t i) {
    TupleKey k;
    k.n = n_tcol;
    for (int j = 0; j < n_tcol; ++j) k.v[j] = pts[i * P + tcol[j]];
    return k;
  };
  // Birthday test on a pseudo-random sample first: points that share tuples in groups of >= min_ratio (at most B / min_ratio
  // tuples) repeat keys among S = 4 sqrt(B) samples about S^2 min_ratio / 2 B >= 190 times, unrelated points never.  Random
  // positions, not strided ones: a stride can fall in step with the axes of a meshgrid.  Costs ~0.1 ms for 131 072 points
  // (the 65 536-sample version of this test took 7 ms per call, 8 % of a general-points step of the bench).
  const int64_t S = std::min<int64_t>(B, std::max<int64_t>(2048, static_cast<int64_t>(4.0 * std::sqrt(static_cast<double>(B)))));
  if (B > S) {
    std::unordered_map<TupleKey, int64_t, TupleKeyHash> seen;      // key -> position it was first seen at
    seen.reserve(static_cast<size_t>(S) * 2);
    uint64_t x = 0x9E3779B97F4A7C15ull;
    int64_t repeats = 0;
    for (int64_t s = 0; s < S; ++s) {
      x += 0x9E3779B97F4A7C15ull;                      // splitmix64
      uint64_t z = x;
      z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
      z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
      z ^= z >> 31;
      const int64_t pos = static_cast<int64_t>(z % static_cast<uint64_t>(B));
      const auto r = seen.emplace(key_of(pos), pos);
      if (!r.second && r.first->second != pos) ++repeats;        // the same key at another position (not the same sample twice)
    }
    if (repeats < 8) return false;
  }
  std::unordered_map<TupleKey, int, TupleKeyHash> ids;
  ids.reserve(1 << 16);
  std::vector<int> id_of(static_cast<size_t>(B));
  std::vector<long long> count;
  const long long max_tuples = B / min_ratio;
  for (int64_t i = 0; i < B; ++i) {
    auto it = ids.find(key_of(i));
    int id;
    if (it == ids.end()) {
      id = static_cast<int>(count.size());
      if (id >= max_tuples) return false;                 // too many tuples for the points there are
      const TupleKey k = key_of(i);
      ids.emplace(k, id);
      g.tuple_vals.insert(g.tuple_vals.end(), k.v, k.v + n_tcol);
      count.push_back(0);
    } else {
      id = it->second;
    }
    id_of[i] = id;
    ++count[id];
  }
  const size_t T = count.size();
  g.start.assign(T + 1, 0);
  for (size_t t = 0; t < T; ++t) {
    g.start[t + 1] = g.start[t] + count[t];
    g.max_count = std::max(g.max_count, count[t]);
  }
  g.perm.resize(static_cast<size_t>(B));
  std::vector<long long> fill(g.start.begin(), g.start.end() - 1);
  for (int64_t i = 0; i < B; ++i) g.perm[fill[id_of[i]]++] = static_cast<int>(i);
  return true;
}

// Evaluates the call through the factorised path if it qualifies (*taken = true), else leaves it to the caller.
int eval_factored(cpx_handle* h, const cpx_eval_args* a, cudaStream_t st, const std::vector<int>& zl, bool* taken) {
  using namespace cpx;
  *taken = false;
  if ((a->flags & CPX_NO_FACTOR) || env_int("CPX_FACTOR", 1) == 0) return CPX_OK;
  if (!a->chi2 || a->chi2_zA || a->model_zAM || a->px_zam || a->P < 1) return CPX_OK;
  const long long min_ratio = std::max(1, env_int("CPX_FACTOR_MIN", 24));
  const int n_zl = static_cast<int>(zl.size());
  const int P = a->P;
  const bool dev_io = (a->flags & CPX_DEVICE_IO) != 0;
  const int flags0 = h->z[zl[0]].dev.flags;
  for (int zi : zl)
    if (h->z[zi].dev.flags != flags0 || h->z[zi].dev.n_A <= 0) return CPX_OK;     // one basis layout for all evaluated z
  bool amp[16];
  int n_tcol = 0;
  for (int j = 0; j < P; ++j) {
    amp[j] = is_amp_slot(a->slot[j]);
    n_tcol += !amp[j];
  }
  std::vector<GridBox> boxes;
  PointGroups groups;
  bool grouped = false;
  if (a->grid_n) {
    long long total = 1;
    for (int j = 0; j < P; ++j) total *= a->grid_n[j];
    if (a->grid_first + a->B > total) return fail(CPX_ERR_ARG, "cpx_eval: grid slice runs past the end of the grid");
    decompose_range(a->grid_first, a->B, P, a->grid_n, amp, boxes);
  } else {
    GridBox b;
    std::memset(&b, 0, sizeof b);
    b.n_tup = 1;
    b.n_amp = a->B;
    if (n_tcol > 0) {
      // explicit points: grouped by tuple on the host when they share tuples (host pointers only)
      if (dev_io || a->B < min_ratio || env_int("CPX_FACTOR_GROUP", 1) == 0) return CPX_OK;
      if (!group_points(a->points, a->B, P, amp, n_tcol, min_ratio, groups)) return CPX_OK;
      grouped = true;
      b.n_tup = static_cast<long long>(groups.start.size()) - 1;
      b.n_amp = groups.max_count;
    }
    boxes.push_back(b);
  }
  long long tuples = 0;
  for (const GridBox& b : boxes) tuples += b.n_tup;
  if (a->B < min_ratio * tuples) return CPX_OK;      // too few points per tuple for the (float64, 3-moment) tuple stage to pay

  *taken = true;
  if (!(boxes.size() == 1 && boxes[0].n_tup == 1)) h->basis_sig.clear();     // this call overwrites the cached Gram data

  const int nb = basis_count(flags0);
  int nA_max = 0, nM_max = 0, nd_max = 0;
  for (int zi : zl) {
    nA_max = std::max(nA_max, h->z[zi].dev.n_A);
    nM_max = std::max(nM_max, h->z[zi].dev.n_M);
    nd_max = std::max(nd_max, h->z[zi].dev.n_data);
  }
  if (a->data_idx && !dev_io)
    for (int64_t i = 0; i < a->B; ++i)
      for (int zi : zl)
        if (a->data_idx[i] < 0 || a->data_idx[i] >= h->z[zi].dev.n_data) return fail(CPX_ERR_ARG, "cpx_eval: data_idx out of range");

  // tuples per batch: nb pseudo-points of the Px_Zam scratch each, Gram data within 256 MB
  long long max_tup = 0;
  for (const GridBox& b : boxes) max_tup = std::max(max_tup, b.n_tup);
  int rc;
  if ((rc = ensure_scratch(h, max_tup * nb))) return rc;
  long long nt_cap = std::max(1, h->chunk / nb);
  nt_cap = std::min<long long>(nt_cap, std::max<long long>(1, (256LL << 20) / (static_cast<long long>(n_zl) * nd_max * nb * 8)));
  const size_t nz_all = h->z.size();
  // fused window + Gram kernel (k_window_gram): one coarse-k block (n_M <= 64) and the same window shape for all z, few
  // enough data vectors for its per-thread g accumulators; otherwise k_window writes the Y_k (model buffer) and k_gram
  // reduces them
  bool fused = env_int("CPX_FUSED_GRAM", 1) != 0 && (kWinPts / nb) * nd_max * nb <= 128 * kWgMaxG;
  for (int zi : zl)
    fused = fused && h->z[zi].dev.n_mblk == 1 && h->z[zi].dev.n_M_pad == h->z[zl[0]].dev.n_M_pad;
  if (!fused && !h->d_model)
    if ((rc = scratch_alloc(h, reinterpret_cast<void**>(&h->d_model), static_cast<size_t>(h->chunk) * nz_all * h->max_nA * h->max_nM * sizeof(double), true)))
      return rc;
  const long long nt_alloc = std::min(nt_cap, max_tup);
  const int n_split = fused ? kWgSplit : 1;           // partial sums of k_window_gram's theta-bin splits
  const size_t G_stride = static_cast<size_t>(nt_alloc) * n_zl * nb * nb, g_stride = static_cast<size_t>(nt_alloc) * n_zl * nd_max * nb;
  if ((rc = grow(h, &h->d_G, &h->cap_G, G_stride * n_split, true))) return rc;
  if ((rc = grow(h, &h->d_g, &h->cap_g, g_stride * n_split, true))) return rc;
  if ((rc = grow(h, &h->d_tpts, &h->cap_tpts, static_cast<size_t>(nt_alloc) * std::max(n_tcol, 1)))) return rc;

  // whole-batch staging in host-pointer mode
  const double* d_points = a->points;
  const int* d_didx = a->data_idx;
  double* d_chi2 = a->chi2;
  if (!dev_io) {
    if ((rc = grow(h, &h->d_big_chi2, &h->cap_big_chi2, static_cast<size_t>(a->B)))) return rc;
    d_chi2 = h->d_big_chi2;
    if (!a->grid_n) {
      if ((rc = grow(h, &h->d_big_pts, &h->cap_big_pts, static_cast<size_t>(a->B) * P))) return rc;
      CPX_CUDA(cudaMemcpyAsync(h->d_big_pts, a->points, sizeof(double) * a->B * P, cudaMemcpyHostToDevice, st));
      d_points = h->d_big_pts;
    }
    if (a->data_idx) {
      if ((rc = grow(h, &h->d_big_didx, &h->cap_big_didx, static_cast<size_t>(a->B)))) return rc;
      CPX_CUDA(cudaMemcpyAsync(h->d_big_didx, a->data_idx, sizeof(int) * a->B, cudaMemcpyHostToDevice, st));
      d_didx = h->d_big_didx;
    }
  }

  if (grouped) {
    if ((rc = grow(h, &h->d_perm, &h->cap_perm, groups.perm.size()))) return rc;
    if ((rc = grow(h, &h->d_tstart, &h->cap_tstart, groups.start.size()))) return rc;
    CPX_CUDA(cudaMemcpyAsync(h->d_perm, groups.perm.data(), sizeof(int) * groups.perm.size(), cudaMemcpyHostToDevice, st));
    CPX_CUDA(cudaMemcpyAsync(h->d_tstart, groups.start.data(), sizeof(long long) * groups.start.size(), cudaMemcpyHostToDevice, st));
  }

  std::vector<cudaEvent_t> evs;
  auto mark = [&]() -> int {
    if (!a->stage_ms) return CPX_OK;
    cudaEvent_t ev;
    CPX_CUDA(cudaEventCreate(&ev));
    CPX_CUDA(cudaEventRecord(ev, st));
    evs.push_back(ev);
    return CPX_OK;
  };

  CPX_CUDA(cudaMemcpyAsync(h->d_zlist, zl.data(), sizeof(int) * n_zl, cudaMemcpyHostToDevice, st));

  // tuple columns only, for k_params
  ParamsArgs pa;
  std::memset(&pa, 0, sizeof(pa));
  pa.P = n_tcol;
  pa.n_z = n_zl;
  pa.points = h->d_tpts;
  QuadArgs q;
  std::memset(&q, 0, sizeof(q));
  q.P = P;
  q.n_z = n_zl;
  q.nb = nb;
  q.grid_mode = a->grid_n ? 1 : 0;
  q.nd_max = nd_max;
  q.points = d_points;
  q.perm = grouped ? h->d_perm : nullptr;
  q.t_start = grouped ? h->d_tstart : nullptr;
  q.data_idx = d_didx;
  q.chi2 = d_chi2;
  q.G = h->d_G;
  q.g = h->d_g;
  q.out_first = a->grid_first;
  {
    int jt = 0;
    long long stride = 1;
    for (int j = P - 1; j >= 0; --j) {
      q.slot[j] = a->slot[j];
      q.zsel[j] = a->zsel ? a->zsel[j] : -1;
      q.is_amp[j] = amp[j] ? 1 : 0;
      if (a->grid_n) {
        q.grid_n[j] = a->grid_n[j];
        q.grid_lo[j] = a->grid_lo[j];
        q.grid_hi[j] = a->grid_hi[j];
        q.grid_step[j] = a->grid_n[j] > 1 ? (a->grid_hi[j] - a->grid_lo[j]) / (a->grid_n[j] - 1) : 0.0;
        q.ax_stride[j] = stride;
        stride *= a->grid_n[j];
      }
    }
    for (int j = 0; j < P; ++j)
      if (!amp[j]) {
        pa.slot[jt] = a->slot[j];
        pa.zsel[jt] = a->zsel ? a->zsel[j] : -1;
        ++jt;
      }
  }

  const int bwarps = env_int("CPX_BASIS_WARPS", 16) == 8 ? 8 : 16;       // tuples per CTA of k_basis_f64
  std::vector<ZDev> tab(n_zl);
  std::vector<double> tpts;
  // signature of a single-tuple call (see cpx_handle::basis_sig); filled below once the tuple values are known
  const bool single_tuple = boxes.size() == 1 && boxes[0].n_tup == 1 && env_int("CPX_BASIS_CACHE", 1) != 0;
  std::vector<double> sig;
  bool same_window_shape = true;
  for (int i = 1; i < n_zl; ++i)
    same_window_shape &= h->z[zl[i]].dev.n_M_pad == h->z[zl[0]].dev.n_M_pad && h->z[zl[i]].dev.n_mblk == h->z[zl[0]].dev.n_mblk;
  int last_nt = -1;
  for (const GridBox& b : boxes) {
    for (int j = 0; j < P; ++j) {
      q.ax_start[j] = b.start[j];
      q.ax_len[j] = a->grid_n ? b.len[j] : 1;
    }
    q.n_amp = b.n_amp;
    for (long long t0 = 0; t0 < b.n_tup; t0 += nt_cap) {
      const int nt = static_cast<int>(std::min<long long>(nt_cap, b.n_tup - t0));
      // ---- stage 0: tuple parameters ----
      if ((rc = mark())) return rc;
      if (grouped) {
        tpts.assign(groups.tuple_vals.begin() + static_cast<size_t>(t0) * n_tcol, groups.tuple_vals.begin() + static_cast<size_t>(t0 + nt) * n_tcol);
        CPX_CUDA(cudaMemcpyAsync(h->d_tpts, tpts.data(), sizeof(double) * tpts.size(), cudaMemcpyHostToDevice, st));
        long long mx = 0;                                  // largest group of this batch: the k_quad grid
        for (int tl = 0; tl < nt; ++tl) mx = std::max(mx, groups.start[t0 + tl + 1] - groups.start[t0 + tl]);
        q.n_amp = mx;
      } else if (n_tcol > 0) {
        tpts.resize(static_cast<size_t>(nt) * n_tcol);
        for (int tl = 0; tl < nt; ++tl) {
          long long tr = t0 + tl;
          int jt = n_tcol;
          for (int j = P - 1; j >= 0; --j) {
            if (amp[j]) continue;
            const int ij = b.start[j] + static_cast<int>(tr % b.len[j]);
            tr /= b.len[j];
            const int n = a->grid_n[j];
            // numpy.linspace, as k_params / k_quad form it (no FMA contraction on the host either)
            volatile double prod = static_cast<double>(ij) * q.grid_step[j];
            tpts[static_cast<size_t>(tl) * n_tcol + --jt] = (ij == n - 1 && n > 1) ? a->grid_hi[j] : prod + a->grid_lo[j];
          }
        }
        CPX_CUDA(cudaMemcpyAsync(h->d_tpts, tpts.data(), sizeof(double) * tpts.size(), cudaMemcpyHostToDevice, st));
      }
      bool cached = false;
      if (single_tuple) {
        sig.assign(tpts.begin(), tpts.begin() + n_tcol);
        for (int j = 0; j < n_tcol; ++j) {
          sig.push_back(pa.slot[j]);
          sig.push_back(pa.zsel[j]);
        }
        for (int zi : zl) sig.push_back(zi);
        sig.push_back(static_cast<double>(h->data_epoch));
        sig.push_back(flags0);
        sig.push_back(nd_max);
        cached = sig == h->basis_sig;
        h->basis_sig.clear();                // re-armed after the call has been enqueued completely
      }
      if (nt != last_nt) {
        long long items = 0;
        for (int i = 0; i < n_zl; ++i) {
          tab[i] = h->z[zl[i]].dev;
          tab[i].item_start = items;
          items += static_cast<long long>(tab[i].n_tiles) * tab[i].n_ablk * ((nt + bwarps - 1) / bwarps);
        }
        CPX_CUDA(cudaMemcpyAsync(h->d_ztab, tab.data(), sizeof(ZDev) * n_zl, cudaMemcpyHostToDevice, st));
        last_nt = nt;
      }
      long long n_items = 0;
      for (int i = 0; i < n_zl; ++i) n_items += static_cast<long long>(tab[i].n_tiles) * tab[i].n_ablk * ((nt + bwarps - 1) / bwarps);
      pa.first = 0;
      pa.count = nt;
      if (!cached) {
        k_params<<<dim3((nt + 127) / 128, n_zl), 128, 0, st>>>(pa, h->d_ztab, h->d_zlist);
        ++h->launches;
      }
      // ---- stage 1: float64 cells, three mu-moment Hankel sums, basis tables ----
      if ((rc = mark())) return rc;
      if (!cached) {
        // the float64 tile of cell constants staged in shared memory when it fits (n_q <= 132 at 24 theta columns)
        const size_t bsmem = basis_smem_bytes((h->max_nq + 3) / 4, h->na_blk / 8);
        const bool stage = bsmem + 1024 <= 227u * 1024u && env_int("CPX_BASIS_STAGE", 1) != 0;
        KBasisFn fn = kbasis_select(h->na_blk / 8, stage, bwarps);
        if (stage) CPX_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bsmem)));
        const int grid = static_cast<int>(std::min<long long>(n_items, static_cast<long long>(h->sm_count) * (stage ? 1 : 4)));
        fn<<<grid, 32 * bwarps, stage ? bsmem : 0, st>>>(h->d_ztab, n_zl, nt, n_items, nb);
        ++h->launches;
      }
      // ---- stage 2: whitened window of every basis table, Gram matrices ----
      if ((rc = mark())) return rc;
      if (!cached && fused) {
        const int npseudo = nt * nb, tp_cta = kWinPts / nb;
        const int mt = tab[0].n_M_pad / 8;
        kwg_select(mt)<<<dim3((nt + tp_cta - 1) / tp_cta, n_zl, n_split), 128, 0, st>>>(h->d_ztab, npseudo, n_zl, nb, nd_max, h->d_G,
                                                                                          h->d_g, G_stride, g_stride);
        const size_t nG = static_cast<size_t>(nt) * n_zl * nb * nb, ng = static_cast<size_t>(nt) * n_zl * nd_max * nb;
        k_gsum<<<static_cast<unsigned>((nG + 255) / 256), 256, 0, st>>>(h->d_G, nG, G_stride, n_split);
        k_gsum<<<static_cast<unsigned>((ng + 255) / 256), 256, 0, st>>>(h->d_g, ng, g_stride, n_split);
        h->launches += 3;
      } else if (!cached) {
        const int npseudo = nt * nb;
        auto window_launch = [&](int i, int zslot, int gz, int nA_grid) {
          const ZDev& Z = tab[i];
          WindowArgs w;
          w.count = npseudo;
          w.zslot = zslot;
          w.n_zl = n_zl;
          w.nA_max = nA_max;
          w.nM_max = nM_max;
          w.data_idx = nullptr;
          w.chi2_zA = nullptr;
          w.model = h->d_model;
          w.whiten = 1;
          dim3 grid((npseudo + kWinPts - 1) / kWinPts, std::max(nA_grid, 1), gz);
          const int mt = Z.n_M_pad / 8 / std::max(Z.n_mblk, 1);
          k2_select<false>(mt)<<<grid, 128, 0, st>>>(h->d_ztab, i, w);
          ++h->launches;
        };
        if (same_window_shape) {
          window_launch(0, -1, n_zl, nA_max);
        } else {
          for (int i = 0; i < n_zl; ++i) window_launch(i, i, 1, tab[i].n_A);
        }
        // coarse theta bins staged per barrier pair: ~24 KB of shared memory (8 CTAs per SM)
        const int at = std::max(1, std::min(nA_max, static_cast<int>(24576 / (sizeof(double) * nb * (nM_max | 1)))));
        const size_t gsmem = std::max(sizeof(double) * nb * gram_row_stride(nM_max, at), sizeof(double) * 256);
        if (gsmem > 48u * 1024u) CPX_CUDA(cudaFuncSetAttribute(k_gram, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(gsmem)));
        k_gram<<<dim3(nt, n_zl), 256, gsmem, st>>>(h->d_ztab, h->d_model, n_zl, nA_max, nM_max, nb, nd_max, at, h->d_G, h->d_g);
        ++h->launches;
      }
      // ---- stage 3: quadratic form per point ----
      if ((rc = mark())) return rc;
      q.t0 = t0;
      q.nt = nt;
      {
        const dim3 qgrid(static_cast<unsigned>((q.n_amp + 256 * kQuadPts - 1) / (256 * kQuadPts)), nt);
        switch (flags0 & 7) {
          case 0: k_quad<0><<<qgrid, 256, 0, st>>>(q, h->d_ztab, h->d_zlist); break;
          case 1: k_quad<1><<<qgrid, 256, 0, st>>>(q, h->d_ztab, h->d_zlist); break;
          case 2: k_quad<2><<<qgrid, 256, 0, st>>>(q, h->d_ztab, h->d_zlist); break;
          case 3: k_quad<3><<<qgrid, 256, 0, st>>>(q, h->d_ztab, h->d_zlist); break;
          case 4: k_quad<4><<<qgrid, 256, 0, st>>>(q, h->d_ztab, h->d_zlist); break;
          case 5: k_quad<5><<<qgrid, 256, 0, st>>>(q, h->d_ztab, h->d_zlist); break;
          case 6: k_quad<6><<<qgrid, 256, 0, st>>>(q, h->d_ztab, h->d_zlist); break;
          case 7: k_quad<7><<<qgrid, 256, 0, st>>>(q, h->d_ztab, h->d_zlist); break;
        }
        ++h->launches;
      }
      if ((rc = mark())) return rc;
      if (cached) ++h->factored_cache_hits;
      else h->factored_tuples += nt;
    }
  }
  ++h->factored_calls;
  CPX_CUDA(cudaGetLastError());
  if (single_tuple) h->basis_sig = sig;
  if (!dev_io) CPX_CUDA(cudaMemcpyAsync(a->chi2, d_chi2, sizeof(double) * a->B, cudaMemcpyDeviceToHost, st));
  if (!dev_io || a->stage_ms) {
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return fail(CPX_ERR_CUDA, std::string("cpx_eval (factorised): ") + cudaGetErrorString(e));
  }
  if (a->stage_ms) {
    for (int s = 0; s < 4; ++s) a->stage_ms[s] = 0.f;
    for (size_t i = 0; i + 4 < evs.size(); i += 5)
      for (int s = 0; s < 4; ++s) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, evs[i + s], evs[i + s + 1]);
        a->stage_ms[s] += ms;
      }
    for (cudaEvent_t ev : evs) cudaEventDestroy(ev);
  }
  return CPX_OK;
}

}  // namespace

// ============================================================================================
// C ABI
// ============================================================================================
extern "C" {

const char* cpx_last_error(void) { return g_err.c_str(); }
int cpx_version(void) { return 200; }
int cpx_device_count(void) {
  int n = 0;
  return cudaGetDeviceCount(&n) == cudaSuccess ? n : 0;
}

void cpx_destroy(cpx_handle* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  for (auto& zh : h->z)
    for (void* p : zh.owned) cudaFree(p);
  for (void* p : h->owned) cudaFree(p);
  for (void* p : h->chunk_owned) cudaFree(p);
  if (h->ev_done) cudaEventDestroy(h->ev_done);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

int cpx_create(const cpx_zbin_desc* zbins, int32_t n_zbins, int32_t device, cpx_handle** out) {
  if (!zbins || n_zbins <= 0 || n_zbins > cpx::kMaxZ || !out) return fail(CPX_ERR_ARG, "cpx_create: bad arguments");
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(CPX_ERR_CUDA, std::string("cpx_create: no CUDA device (") + cudaGetErrorString(e) +
                                  "); cupix_b200 has no CPU fallback");
  if (device < 0 || device >= ndev) return fail(CPX_ERR_ARG, "cpx_create: device index out of range");
  CPX_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  CPX_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10)
    return fail(CPX_ERR_UNSUPPORTED, "cpx_create: kernels are built for sm_100a (Blackwell) only");

  cpx_handle* h = new cpx_handle;
  h->device = device;
  h->sm_count = prop.multiProcessorCount;
  auto bail = [&](int rc) { cpx_destroy(h); return rc; };
  if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess)
    return bail(fail(CPX_ERR_CUDA, "cudaStreamCreate failed"));

  for (int i = 0; i < n_zbins; ++i) {
    const cpx_zbin_desc& d = zbins[i];
    if (d.n_a <= 0 || d.n_m <= 0 || d.n_q <= 0 || !d.kpar_Mpc || !d.kperp_Mpc || !d.hankel_w || !d.linP_cell)
      return bail(fail(CPX_ERR_ARG, "cpx_create: z bin with missing theory arrays"));
    h->max_na = std::max(h->max_na, d.n_a);
    h->max_nm = std::max(h->max_nm, d.n_m);
    h->max_nA = std::max(h->max_nA, d.n_A);
    h->max_nM = std::max(h->max_nM, d.n_M);
    h->max_nq = std::max(h->max_nq, d.n_q);
  }
  h->na_blk = choose_na_blk(h->max_na);
  const size_t smem_need = static_cast<size_t>((h->max_nq + 3) / 4) * (cpx::kCellConsts * 4 * 32 * 4 + h->na_blk * 32);
  if (smem_need > static_cast<size_t>(prop.sharedMemPerBlockOptin) - 1024)
    return bail(fail(CPX_ERR_UNSUPPORTED, "cpx_create: n_q too large for the shared-memory-resident grid"));

  // INT8 mma.sync Hankel path, opt-in (CPX_IMMA=1): parity-green, measured level with the FP64 DMMA kernel (50.7 against
  // 49.3 ms per 131072-point S5 step at 56 nodes, 73.1 against 73.7 at 88; profiles/hankel_imma_r02.txt)
  h->im_nks = 0;
  if (env_int("CPX_IMMA", 0) != 0 && imma_supported((h->max_nq + 3) / 4)) h->im_nks = (h->max_nq + 3) / 4;

  h->z.resize(n_zbins);
  size_t per_point_bytes = 0;
  for (int i = 0; i < n_zbins; ++i) {
    int rc = build_zbin(h, zbins[i], h->na_blk, h->z[i]);
    if (rc) return bail(rc);
    if (zbins[i].n_A > 0) {
      rc = upload_yd(h, h->z[i], zbins[i].Px_AM, zbins[i].n_data);
      if (rc) return bail(rc);
    }
    per_point_bytes += static_cast<size_t>(h->z[i].dev.n_a) * h->z[i].dev.n_m_pad * 8 + sizeof(cpx::PtPar) + sizeof(cpx::RowPar) * h->z[i].dev.n_m_pad;
  }
  // chunk_max: Px_Zam scratch of all z within the budget (default 6 GiB), multiple of 256 points.  The scratch itself is
  // allocated by ensure_scratch() for the largest batch seen so far, so that engines which only serve scalar calls
  // (Theory.get_px_obs, Likelihood.get_chi2) do not pin gigabytes of HBM each.
  const size_t budget = static_cast<size_t>(env_int("CPX_SCRATCH_MB", 6144)) << 20;
  long long chunk = static_cast<long long>(budget / std::max<size_t>(per_point_bytes, 1));
  chunk = std::max(256LL, std::min(chunk, 65536LL)) / 256 * 256;
  // The window kernels run one CTA per (128-point tile, z) and SM: within the upper half of what the budget allows, take the
  // chunk whose grid fills its last wave best (S5: 9472 points = 74 tiles x 12 z = 6.00 waves of 148 SMs instead of 4.70 with
  // 7424 points: 18.7 against 19.4 ms per 131 072-point step, profiles/window_tc_depths_r02.txt)
  if (chunk >= 2048) {
    long long best = chunk;
    double best_eff = 0.0;
    for (long long c = chunk; c >= chunk / 2; c -= 256) {
      const double T = static_cast<double>(c / 128) * n_zbins;
      const double eff = T / (std::ceil(T / h->sm_count) * h->sm_count);
      if (eff > best_eff + 1e-9) {
        best_eff = eff;
        best = c;
      }
    }
    chunk = best;
  }
  h->chunk_max = static_cast<int>(chunk);
  h->chunk = 0;

  h->guard = env_int("CPX_GUARD", 0) != 0;
  auto dmalloc = [&](void** p, size_t bytes) -> int { return scratch_alloc(h, p, bytes); };
  int rc;
  if ((rc = dmalloc(reinterpret_cast<void**>(&h->d_ztab), sizeof(cpx::ZDev) * n_zbins))) return bail(rc);
  if ((rc = dmalloc(reinterpret_cast<void**>(&h->d_zlist), sizeof(int) * n_zbins))) return bail(rc);
  if ((rc = dmalloc(reinterpret_cast<void**>(&h->d_nA_of_z), sizeof(int) * n_zbins))) return bail(rc);
  if (cudaEventCreateWithFlags(&h->ev_done, cudaEventDisableTiming) != cudaSuccess)
    return bail(fail(CPX_ERR_CUDA, "cudaEventCreate failed"));

  // opt in to the large dynamic shared memory for every k_p3d_hankel instantiation we may launch
  for (int pp = 1; pp <= 2; ++pp)
    for (int add = 0; add < 4; ++add) {
      K1Fn fn = k1_select(pp, h->na_blk / 8, (add & 1) != 0, (add & 2) != 0);
      if (!fn) return bail(fail(CPX_ERR_UNSUPPORTED, "no k_p3d_hankel instantiation for this theta-bin count"));
      if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_need)) != cudaSuccess)
        return bail(fail(CPX_ERR_CUDA, "cudaFuncSetAttribute(MaxDynamicSharedMemorySize) failed"));
    }
  if (h->im_nks > 0) {
    bool ok = true;
    for (auto& zh : h->z) ok &= zh.dev.im_cell != nullptr;        // every z on the same node count
    if (!ok) h->im_nks = 0;
    for (int add = 0; add < 2 && h->im_nks > 0; ++add) {
      K1Fn fn = k1_select_imma(h->im_nks, h->na_blk / 8, add != 0);
      if (!fn || cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(imma_smem_bytes(h->im_nks, h->na_blk / 8))) != cudaSuccess)
        return bail(fail(CPX_ERR_CUDA, "cudaFuncSetAttribute(MaxDynamicSharedMemorySize) failed for k_p3d_hankel_imma"));
    }
  }
  // tensor-core Hankel path: every z must have its tables (n_q <= 96) and share the pairs-per-lane template value
  h->tc_np = 0;
  if (env_int("CPX_TC", 0) != 0) {   // opt-in: measured 43.7 ms against 38.7 ms of the FP64 DMMA kernel per 65536-point S5 step
    int np = 0;
    bool ok = true;
    for (auto& zh : h->z) {
      ok &= zh.dev.tc_np > 0;
      np = std::max(np, zh.dev.tc_np);
    }
    for (auto& zh : h->z) ok &= zh.dev.tc_np == np;
    if (ok && static_cast<size_t>(cpx::tc_smem_bytes(np)) + 4096 <= static_cast<size_t>(prop.sharedMemPerBlockOptin)) {
      for (int add = 0; add < 2; ++add)
        if (cudaFuncSetAttribute(k1_select_tc(np, add != 0), cudaFuncAttributeMaxDynamicSharedMemorySize, cpx::tc_smem_bytes(np)) != cudaSuccess)
          return bail(fail(CPX_ERR_CUDA, "cudaFuncSetAttribute(MaxDynamicSharedMemorySize) failed for k_p3d_hankel_tc"));
      h->tc_np = np;
    }
  }
  // tensor-core window path: every z with a measurement needs its digit tables
  h->wt_nb = 0;
  if (env_int("CPX_WTC", 1) != 0) {
    int nb = 0;
    bool ok = true;
    for (auto& zh : h->z)
      if (zh.dev.n_A > 0) {
        ok &= zh.dev.wt_nb > 0;
        nb = std::max(nb, zh.dev.wt_nb);
      }
    if (ok && nb > 0 && static_cast<size_t>(cpx::wt_smem_bytes(nb)) + 2048 <= static_cast<size_t>(prop.sharedMemPerBlockOptin)) {
      if (cudaFuncSetAttribute(cpx::k_window_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, cpx::wt_smem_bytes(nb)) != cudaSuccess)
        return bail(fail(CPX_ERR_CUDA, "cudaFuncSetAttribute(MaxDynamicSharedMemorySize) failed for k_window_tc"));
      // tensor maps of the Px_Zam scratch of every z (k_window_tc streams it with tiled TMA loads) are encoded by
      // ensure_scratch() whenever the scratch is (re-)allocated
      void* fnp = nullptr;
      cudaDriverEntryPointQueryResult qres;
      if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &qres) != cudaSuccess || !fnp ||
          qres != cudaDriverEntryPointSuccess)
        return bail(fail(CPX_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver"));
      h->encode_tmap = fnp;
      if ((rc = dmalloc(&h->d_tmaps, sizeof(CUtensorMap) * n_zbins))) return bail(rc);
      h->wt_nb = nb;
    }
  }
  *out = h;
  return CPX_OK;
}

int cpx_set_data(cpx_handle* h, int32_t iz, const double* Px_AM, int32_t n_data) {
  if (!h || iz < 0 || iz >= static_cast<int>(h->z.size()) || !Px_AM || n_data <= 0)
    return fail(CPX_ERR_ARG, "cpx_set_data: bad arguments");
  if (h->z[iz].dev.n_A <= 0) return fail(CPX_ERR_ARG, "cpx_set_data: z bin has no measurement");
  CPX_CUDA(cudaSetDevice(h->device));
  CPX_CUDA(cudaDeviceSynchronize());     // evaluations may be in flight on any caller stream
  return upload_yd(h, h->z[iz], Px_AM, n_data);
}

int cpx_generate_mocks(cpx_handle* h, int32_t P, const int32_t* slot, const int32_t* zsel, const double* point,
                       int32_t n_mocks, uint64_t seed, double* mocks_zdAM) {
  if (!h || n_mocks <= 0 || P < 0 || P > 16 || (P > 0 && (!slot || !point)))
    return fail(CPX_ERR_ARG, "cpx_generate_mocks: bad arguments");
  const int n_z = static_cast<int>(h->z.size());
  int nA_max = 0, nM_max = 0;
  for (auto& zh : h->z) {
    if (zh.dev.n_A <= 0) return fail(CPX_ERR_ARG, "cpx_generate_mocks: a z bin of the handle has no measurement");
    nA_max = std::max(nA_max, zh.dev.n_A);
    nM_max = std::max(nM_max, zh.dev.n_M);
  }
  // whitened model L^-1 m of every z at the truth: one point through the full pipeline
  std::vector<double> ym(static_cast<size_t>(n_z) * nA_max * nM_max, 0.0);
  cpx_eval_args ea;
  std::memset(&ea, 0, sizeof ea);
  ea.B = 1;
  ea.P = P;
  ea.flags = CPX_MODEL_WHITENED | CPX_NO_FACTOR;
  ea.points = point;
  ea.slot = slot;
  ea.zsel = zsel;
  ea.model_zAM = ym.data();
  int rc = cpx_eval(h, &ea);
  if (rc) return rc;
  CPX_CUDA(cudaSetDevice(h->device));
  CPX_CUDA(cudaDeviceSynchronize());         // evaluations may be in flight on any caller stream
  double* d_ym = nullptr;
  CPX_CUDA(cudaMalloc(reinterpret_cast<void**>(&d_ym), sizeof(double) * std::max(1, nA_max * nM_max)));
  std::vector<double> yd_host, tmp;
  for (int iz = 0; iz < n_z; ++iz) {
    ZHost& zh = h->z[iz];
    cpx::ZDev& Z = zh.dev;
    const int nA = Z.n_A, nM = Z.n_M;
    tmp.assign(static_cast<size_t>(nA) * nM, 0.0);
    for (int A = 0; A < nA; ++A)
      std::copy(&ym[(static_cast<size_t>(iz) * nA_max + A) * nM_max], &ym[(static_cast<size_t>(iz) * nA_max + A) * nM_max] + nM, &tmp[static_cast<size_t>(A) * nM]);
    cudaError_t e = cudaMemcpy(d_ym, tmp.data(), sizeof(double) * tmp.size(), cudaMemcpyHostToDevice);
    double *d_yd = nullptr, *d_ydd = nullptr;
    const size_t per_data = static_cast<size_t>(nA) * Z.n_M_pad;
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&d_yd), sizeof(double) * per_data * n_mocks);
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&d_ydd), sizeof(double) * n_mocks);
    if (e == cudaSuccess) e = cudaMemset(d_yd, 0, sizeof(double) * per_data * n_mocks);
    if (e == cudaSuccess) {
      const long long pairs = (static_cast<long long>(n_mocks) * nA * nM + 1) / 2;
      cpx::k_mock_noise<<<static_cast<unsigned>((pairs + 255) / 256), 256>>>(d_ym, nA, nM, Z.n_M_pad, n_mocks, seed, iz, d_yd);
      cpx::k_ydd<<<(n_mocks * 32 + 255) / 256, 256>>>(d_yd, static_cast<int>(per_data), n_mocks, d_ydd);
      h->launches += 2;
      e = cudaDeviceSynchronize();
    }
    if (e != cudaSuccess) {
      cudaFree(d_yd);
      cudaFree(d_ydd);
      cudaFree(d_ym);
      return fail(CPX_ERR_CUDA, std::string("cpx_generate_mocks: ") + cudaGetErrorString(e));
    }
    for (const double* old : {Z.yd, Z.ydd})
      if (old) {
        cudaFree(const_cast<double*>(old));
        zh.owned.erase(std::remove(zh.owned.begin(), zh.owned.end(), (void*)old), zh.owned.end());
      }
    zh.owned.push_back(d_yd);
    zh.owned.push_back(d_ydd);
    Z.yd = d_yd;
    Z.ydd = d_ydd;
    Z.n_data = n_mocks;
    ++h->data_epoch;
    if (mocks_zdAM) {     // data space: d = L yd
      yd_host.resize(per_data * n_mocks);
      CPX_CUDA(cudaMemcpy(yd_host.data(), d_yd, sizeof(double) * yd_host.size(), cudaMemcpyDeviceToHost));
      for (int dd = 0; dd < n_mocks; ++dd)
        for (int A = 0; A < nA; ++A) {
          const double* L = &zh.Lfac[static_cast<size_t>(A) * nM * nM];
          const double* y = &yd_host[(static_cast<size_t>(dd) * nA + A) * Z.n_M_pad];
          double* out = mocks_zdAM + ((static_cast<size_t>(iz) * n_mocks + dd) * nA_max + A) * nM_max;
          for (int M = 0; M < nM; ++M) {
            double sacc = 0.0;
            for (int Mp = 0; Mp <= M; ++Mp) sacc += L[M * nM + Mp] * y[Mp];
            out[M] = sacc;
          }
        }
    }
  }
  cudaFree(d_ym);
  return CPX_OK;
}

int cpx_get_info(const cpx_handle* h, int32_t what, int64_t* value) {
  if (!h || !value) return fail(CPX_ERR_ARG, "cpx_get_info: bad arguments");
  switch (what) {
    case CPX_INFO_N_ZBINS: *value = static_cast<int64_t>(h->z.size()); break;
    case CPX_INFO_MAX_NA: *value = h->max_na; break;
    case CPX_INFO_MAX_NM: *value = h->max_nm; break;
    case CPX_INFO_MAX_N_A: *value = h->max_nA; break;
    case CPX_INFO_MAX_N_M: *value = h->max_nM; break;
    case CPX_INFO_CHUNK: *value = h->chunk_max; break;
    case CPX_INFO_CHUNK_ALLOCATED: *value = h->chunk; break;
    case CPX_INFO_FACTORED_CALLS: *value = h->factored_calls; break;
    case CPX_INFO_FACTORED_TUPLES: *value = h->factored_tuples; break;
    case CPX_INFO_FACTORED_CACHE_HITS: *value = h->factored_cache_hits; break;
    case CPX_INFO_HANKEL_KERNEL: *value = h->tc_np > 0 ? 1 : (h->im_nks > 0 ? 2 : 0); break;
    case CPX_INFO_DEVICE_BYTES: *value = static_cast<int64_t>(h->device_bytes); break;
    case CPX_INFO_KERNEL_LAUNCHES: *value = h->launches; break;
    case CPX_INFO_SM_COUNT: *value = h->sm_count; break;
    case CPX_INFO_GUARD_BUFFERS: *value = static_cast<int64_t>(h->guarded.size()); break;
    case CPX_INFO_GUARD_VIOLATIONS: {
      // synchronises the device; counts guard-band bytes that no longer hold the fill pattern
      if (cudaSetDevice(h->device) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess)
        return fail(CPX_ERR_CUDA, "cpx_get_info: device synchronisation failed");
      std::vector<unsigned char> band(kGuardBytes + 256);
      int64_t bad = 0;
      for (const auto& g : h->guarded) {
        const size_t payload = (g.bytes + 255) / 256 * 256;
        const size_t rear = payload - g.bytes + kGuardBytes;
        CPX_CUDA(cudaMemcpy(band.data(), g.base, kGuardBytes, cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < kGuardBytes; ++i) bad += band[i] != kGuardByte;
        CPX_CUDA(cudaMemcpy(band.data(), g.base + kGuardBytes + g.bytes, rear, cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < rear; ++i) bad += band[i] != kGuardByte;
      }
      *value = bad;
      break;
    }
    default: return fail(CPX_ERR_ARG, "cpx_get_info: unknown item");
  }
  return CPX_OK;
}

int cpx_eval(cpx_handle* h, const cpx_eval_args* a) {
  using namespace cpx;
  if (!h || !a) return fail(CPX_ERR_ARG, "cpx_eval: null handle or args");
  if (a->B < 0 || a->P < 0 || a->P > 16) return fail(CPX_ERR_ARG, "cpx_eval: B < 0 or P outside 0..16");
  if (a->P > 0 && !a->slot) return fail(CPX_ERR_ARG, "cpx_eval: slot is NULL");
  if (a->P > 0 && !a->points && !a->grid_n) return fail(CPX_ERR_ARG, "cpx_eval: neither points nor a grid given");
  if (a->grid_n && (!a->grid_lo || !a->grid_hi)) return fail(CPX_ERR_ARG, "cpx_eval: grid_n given without grid_lo / grid_hi");
  if (a->grid_n && a->grid_first < 0) return fail(CPX_ERR_ARG, "cpx_eval: grid_first < 0");
  for (int j = 0; j < a->P; ++j) {
    if (a->slot[j] < 0 || a->slot[j] >= CPX_NSLOTS) return fail(CPX_ERR_ARG, "cpx_eval: slot out of range");
    if (a->grid_n && a->grid_n[j] < 1) return fail(CPX_ERR_ARG, "cpx_eval: grid_n < 1");
  }
  // CPX_AS / CPX_NS columns: the linear power is rescaled per point inside the cell loop
  bool cosmo = false, tilt = false;
  for (int j = 0; j < a->P; ++j) {
    cosmo |= a->slot[j] == CPX_AS || a->slot[j] == CPX_NS;
    tilt |= a->slot[j] == CPX_NS;
  }
  if (cosmo)
    for (auto& zh : h->z) {
      if (!(zh.dev.k_pivot > 0.0))
        return fail(CPX_ERR_UNSUPPORTED, "cpx_eval: As / ns columns need k_pivot_Mpc and the default As of the plan's cosmology");
      // the SiIII tables are Hankel moments of P_L at other redshifts: an amplitude factors out of them, a tilt does not --
      // it needs their Taylor tables in the tilt (cpx_zbin_desc::metal_auto_dn / metal_cross_dn)
      if (tilt && (zh.dev.flags & CPX_INCLUDE_METAL) && !(zh.dev.metal_auto_dn && zh.dev.metal_cross_dn))
        return fail(CPX_ERR_UNSUPPORTED, "cpx_eval: ns as a batch column with include_metal needs the tilt tables metal_auto_dn / metal_cross_dn in the descriptor");
    }
  if (a->B == 0) return CPX_OK;
  CPX_CUDA(cudaSetDevice(h->device));
  const bool dev_io = (a->flags & CPX_DEVICE_IO) != 0;
  const bool precise = (a->flags & CPX_PRECISE_CELLS) != 0;
  // Device I/O: the caller's buffers are ordered on a->stream, NULL meaning the CUDA default (legacy) stream -- what
  // torch.cuda.current_stream().cuda_stream is for torch's default stream.  Host I/O: the handle's own stream unless given.
  cudaStream_t st = (dev_io || a->stream) ? static_cast<cudaStream_t>(a->stream) : h->stream;
  // all calls share the handle's scratch: order this call after the previous one if that ran on another stream
  if (h->have_last && h->last_stream != st) CPX_CUDA(cudaStreamWaitEvent(st, h->ev_done, 0));
  struct Done {        // records the ordering event on every exit path once work may have been enqueued
    cpx_handle* h;
    cudaStream_t st;
    ~Done() {
      if (cudaEventRecord(h->ev_done, st) == cudaSuccess) {
        h->last_stream = st;
        h->have_last = true;
      }
    }
  } done_guard{h, st};

  // ---- z selection ----
  std::vector<int> zl;
  if (a->z_list) {
    if (a->n_z_list <= 0) return fail(CPX_ERR_ARG, "cpx_eval: empty z_list");
    for (int i = 0; i < a->n_z_list; ++i) {
      if (a->z_list[i] < 0 || a->z_list[i] >= static_cast<int>(h->z.size()))
        return fail(CPX_ERR_ARG, "cpx_eval: z_list entry out of range");
      zl.push_back(a->z_list[i]);
    }
  } else {
    for (size_t i = 0; i < h->z.size(); ++i) zl.push_back(static_cast<int>(i));
  }
  const int n_zl = static_cast<int>(zl.size());
  const bool want_chi2 = a->chi2 || a->chi2_zA;
  const bool want_model = a->model_zAM != nullptr;
  const bool want_px = a->px_zam != nullptr;
  if (want_chi2 || want_model)
    for (int zi : zl)
      if (h->z[zi].dev.n_A <= 0) return fail(CPX_ERR_ARG, "cpx_eval: chi2/model requested for a theory-only z bin");
  if (a->data_idx && dev_io == false) {
    // validated on the host when available
    for (int64_t i = 0; i < a->B; ++i)
      for (int zi : zl)
        if (a->data_idx[i] < 0 || a->data_idx[i] >= h->z[zi].dev.n_data)
          return fail(CPX_ERR_ARG, "cpx_eval: data_idx out of range");
  }

  int rc;
  // ---- factorised path (cpx_factor.cuh): Hankel + window once per distinct tuple, chi2 per point a quadratic form ----
  {
    bool taken = false;
    rc = eval_factored(h, a, st, zl, &taken);
    if (taken || rc) return rc;
  }
  if ((rc = ensure_scratch(h, a->B))) return rc;

  // lazy scratch for the bulky optional outputs (sized by the chunk, dropped when the chunk grows)
  auto lazy = [&](double** p, size_t n) -> int {
    if (*p) return CPX_OK;
    return scratch_alloc(h, reinterpret_cast<void**>(p), n * sizeof(double), true);
  };
  const size_t nz_all = h->z.size();
  if (want_model && !dev_io)
    if ((rc = lazy(&h->d_model, static_cast<size_t>(h->chunk) * nz_all * h->max_nA * h->max_nM))) return rc;
  if (want_px && !dev_io)
    if ((rc = lazy(&h->d_pxout, static_cast<size_t>(h->chunk) * nz_all * h->max_na * h->max_nm))) return rc;

  // ---- device table for this launch ----
  std::vector<ZDev> tab(n_zl);
  std::vector<int> nA_of_z(n_zl);
  const int pp_env = env_int("CPX_PP", 0);
  // output strides follow the evaluated list
  int na_max = 0, nm_max = 0, nA_max = 0, nM_max = 0;
  for (int zi : zl) {
    na_max = std::max(na_max, h->z[zi].dev.n_a);
    nm_max = std::max(nm_max, h->z[zi].dev.n_m);
    nA_max = std::max(nA_max, h->z[zi].dev.n_A);
    nM_max = std::max(nM_max, h->z[zi].dev.n_M);
  }

  ParamsArgs pa;
  std::memset(&pa, 0, sizeof(pa));
  pa.P = a->P;
  pa.n_z = n_zl;
  pa.grid_mode = a->grid_n ? 1 : 0;
  pa.grid_first = a->grid_first;
  for (int j = 0; j < a->P; ++j) {
    pa.slot[j] = a->slot[j];
    pa.zsel[j] = a->zsel ? a->zsel[j] : -1;
    if (a->grid_n) {
      pa.grid_n[j] = a->grid_n[j];
      pa.grid_lo[j] = a->grid_lo[j];
      pa.grid_hi[j] = a->grid_hi[j];
      pa.grid_step[j] = a->grid_n[j] > 1 ? (a->grid_hi[j] - a->grid_lo[j]) / (a->grid_n[j] - 1) : 0.0;
    }
  }

  std::vector<cudaEvent_t> evs;
  auto mark = [&]() -> int {
    if (!a->stage_ms) return CPX_OK;
    cudaEvent_t ev;
    CPX_CUDA(cudaEventCreate(&ev));
    CPX_CUDA(cudaEventRecord(ev, st));
    evs.push_back(ev);
    return CPX_OK;
  };

  CPX_CUDA(cudaMemcpyAsync(h->d_zlist, zl.data(), sizeof(int) * n_zl, cudaMemcpyHostToDevice, st));
  for (int i = 0; i < n_zl; ++i) nA_of_z[i] = h->z[zl[i]].dev.n_A;
  CPX_CUDA(cudaMemcpyAsync(h->d_nA_of_z, nA_of_z.data(), sizeof(int) * n_zl, cudaMemcpyHostToDevice, st));

  // chi2 stage on the tensor cores (k_window_tc) when the call carries enough (point, z) pairs to amortise its per-tile
  // fixed costs (tensor-memory allocation, pipeline fill); below that the FP64 DMMA GEMM has the lower latency
  // (profiles/latency_report_r01.txt).  One decision per call, so that all chunks of a batch take the same kernel.
  const bool wtc = h->wt_nb > 0 && want_chi2 && static_cast<long long>(a->B) * n_zl >= env_int("CPX_WTC_MIN", 49152);
  int last_table_count = -1, last_pp = -1;
  for (int64_t first = 0; first < a->B; first += h->chunk) {
    const int count = static_cast<int>(std::min<int64_t>(h->chunk, a->B - first));
    // one point per warp and two CTAs per SM measured fastest; node sets whose staged tile leaves room for one CTA only
    // (n_q > 116: the 'bao' family from 128 nodes) take two points per warp instead
    const size_t k1_smem = static_cast<size_t>((h->max_nq + 3) / 4) * (kCellConsts * 4 * 32 * 4 + h->na_blk * 32);
    int pp = pp_env > 0 ? std::min(pp_env, 2) : (2 * (k1_smem + 2048) <= 227u * 1024u ? 1 : 2);
    const bool tc = h->tc_np > 0 && !precise && !cosmo;   // k_p3d_hankel_tc has no rescaling variant
    const bool imma = h->im_nks > 0 && !precise && !tc && !cosmo;   // INT8 mma.sync Hankel kernel: one point per warp (the default)
    if (imma) pp = 1;
    // float64 cells: the shared-memory staged kernel (16 points per CTA) when the float64 tile fits, else one warp per point
    const size_t f64_smem = basis_smem_bytes((h->max_nq + 3) / 4, h->na_blk / 8);
    const bool f64_staged = precise && f64_smem + 1024 <= 227u * 1024u && env_int("CPX_F64_STAGE", 1) != 0;
    const int n_pg = precise ? (f64_staged ? (count + 15) / 16 : count)
                             : (tc ? (count + kTcPts - 1) / kTcPts : (count + kK1Warps * pp - 1) / (kK1Warps * pp));
    auto items_of = [&](const ZDev& Z) {
      return tc ? static_cast<long long>(Z.tc_nt) * Z.tc_nablk * n_pg : static_cast<long long>(Z.n_tiles) * Z.n_ablk * n_pg;
    };
    if (count != last_table_count || pp != last_pp) {
      long long items = 0;
      for (int i = 0; i < n_zl; ++i) {
        tab[i] = h->z[zl[i]].dev;
        tab[i].item_start = items;
        items += items_of(tab[i]);
      }
      // stream-ordered: the previous chunk's kernels have been enqueued before this copy
      CPX_CUDA(cudaMemcpyAsync(h->d_ztab, tab.data(), sizeof(ZDev) * n_zl, cudaMemcpyHostToDevice, st));
      last_table_count = count;
      last_pp = pp;
    }
    long long n_items = 0;
    for (int i = 0; i < n_zl; ++i) n_items += items_of(tab[i]);

    // ---- stage 0: parameters ----
    if ((rc = mark())) return rc;
    pa.first = 0;
    pa.count = count;
    if (!pa.grid_mode) {
      if (dev_io) {
        pa.points = a->points + first * a->P;
      } else if (a->P > 0) {
        CPX_CUDA(cudaMemcpyAsync(h->d_points, a->points + first * a->P, sizeof(double) * count * a->P,
                                 cudaMemcpyHostToDevice, st));
        pa.points = h->d_points;
      }
    } else {
      pa.first = first;
    }
    const int* d_didx = nullptr;
    if (a->data_idx) {
      if (dev_io) {
        d_didx = a->data_idx + first;
      } else {
        CPX_CUDA(cudaMemcpyAsync(h->d_data_idx, a->data_idx + first, sizeof(int) * count, cudaMemcpyHostToDevice, st));
        d_didx = h->d_data_idx;
      }
    }
    {
      dim3 grid((count + 127) / 128, n_zl);
      k_params<<<grid, 128, 0, st>>>(pa, h->d_ztab, h->d_zlist);
      ++h->launches;
      bool any_rowwise = false;
      int nm_pad_max = 0;
      for (int i = 0; i < n_zl; ++i) {
        any_rowwise |= (tab[i].flags & (CPX_INCLUDE_HCD | CPX_INCLUDE_CONTINUUM)) != 0;
        nm_pad_max = std::max(nm_pad_max, tab[i].n_m_pad);
      }
      if (any_rowwise) {
        dim3 rgrid(static_cast<unsigned>((static_cast<long long>(count) * nm_pad_max + 255) / 256), n_zl);
        k_rowpars<<<rgrid, 256, 0, st>>>(h->d_ztab, count);
        ++h->launches;
      }
    }
    // ---- stage 1: P3D + Hankel ----
    if ((rc = mark())) return rc;
    {
      bool additive = false;
      for (int i = 0; i < n_zl; ++i) additive |= (tab[i].flags & (CPX_INCLUDE_METAL | CPX_INCLUDE_SKY)) != 0;
      if (precise && f64_staged) {
        K1Fn fn = k1_select_f64s(h->na_blk / 8);
        CPX_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(f64_smem)));
        const int grid = static_cast<int>(std::min<long long>(n_items, h->sm_count));
        fn<<<grid, 512, f64_smem, st>>>(h->d_ztab, n_zl, count, n_items);
      } else if (precise) {
        K1Fn fn = k1_select_f64(h->na_blk / 8);
        const int grid = static_cast<int>(std::min<long long>((n_items + 7) / 8, 8LL * h->sm_count));
        fn<<<grid, 256, 0, st>>>(h->d_ztab, n_zl, count, n_items);
      } else if (imma) {
        K1Fn fn = k1_select_imma(h->im_nks, h->na_blk / 8, additive);
        const int grid = static_cast<int>(std::min<long long>(n_items, 2LL * h->sm_count));
        fn<<<grid, kK1Threads, imma_smem_bytes(h->im_nks, h->na_blk / 8), st>>>(h->d_ztab, n_zl, count, n_items);
      } else if (tc) {
        K1Fn fn = k1_select_tc(h->tc_np, additive);
        const int grid = static_cast<int>(std::min<long long>(n_items, h->sm_count));
        fn<<<grid, kTcThreads, tc_smem_bytes(h->tc_np), st>>>(h->d_ztab, n_zl, count, n_items);
      } else {
        K1Fn fn = k1_select(pp, h->na_blk / 8, additive, cosmo);
        const size_t smem = static_cast<size_t>((h->max_nq + 3) / 4) * (kCellConsts * 4 * 32 * 4 + h->na_blk * 32);
        const int grid = static_cast<int>(std::min<long long>(n_items, static_cast<long long>(h->sm_count) * k1_ctas_per_sm(pp)));
        fn<<<grid, kK1Threads, smem, st>>>(h->d_ztab, n_zl, count, n_items);
      }
      ++h->launches;
    }
    // ---- stage 2: window (+ whitening, chi2) ----
    if ((rc = mark())) return rc;
    double* chi2_zA_dst = nullptr;
    if (want_chi2) chi2_zA_dst = (dev_io && a->chi2_zA) ? a->chi2_zA + first * n_zl * nA_max : h->d_chi2_zA;
    if (a->chi2_zA) {   // slots A >= n_A of a z bin with fewer coarse theta bins than the widest one read as zero
      bool ragged = false;
      for (int i = 0; i < n_zl; ++i) ragged |= tab[i].n_A < nA_max;
      if (ragged) CPX_CUDA(cudaMemsetAsync(chi2_zA_dst, 0, sizeof(double) * count * n_zl * nA_max, st));
    }
    double* model_dst = nullptr;
    if (want_model) model_dst = dev_io ? a->model_zAM + first * n_zl * nA_max * nM_max : h->d_model;
    double* px_dst = nullptr;
    if (want_px) px_dst = dev_io ? a->px_zam + first * n_zl * na_max * nm_max : h->d_pxout;
    bool same_window_shape = true;
    for (int i = 1; i < n_zl; ++i)
      same_window_shape &= tab[i].n_M_pad == tab[0].n_M_pad && tab[i].n_mblk == tab[0].n_mblk;
    auto window_launch = [&](int i, int zslot, int gz, int nA_grid) {
      const ZDev& Z = tab[i];
      WindowArgs w;
      w.count = count;
      w.zslot = zslot;
      w.n_zl = n_zl;
      w.nA_max = nA_max;
      w.nM_max = nM_max;
      w.data_idx = d_didx;
      w.chi2_zA = chi2_zA_dst;
      w.model = model_dst;
      w.whiten = (a->flags & CPX_MODEL_WHITENED) ? 1 : 0;
      dim3 grid((count + kWinPts - 1) / kWinPts, std::max(nA_grid, 1), gz);
      const int mt = Z.n_M_pad / 8 / std::max(Z.n_mblk, 1);
      if (want_chi2) {
        if (wtc) {
          // theta_A groups per 128-point tile: one CTA walking all theta_A re-uses its pipeline fill best (19.4 against 20.2 ms
          // per 131 072-point S5 step, profiles/window_tc_depths_r02.txt); more groups only while the grid is below four waves
          const long long tiles = static_cast<long long>((count + kWtPts - 1) / kWtPts) * gz;
          const int auto_groups = static_cast<int>(std::min<long long>(4, std::max<long long>(1, (4LL * h->sm_count + tiles - 1) / tiles)));
          const int agroups = env_int("CPX_WTC_AGROUPS", 0) > 0 ? env_int("CPX_WTC_AGROUPS", 0) : auto_groups;
          dim3 tgrid((count + kWtPts - 1) / kWtPts, std::min(std::max(nA_grid, 1), agroups), gz);
          k_window_tc<<<tgrid, kWtThreads, wt_smem_bytes(h->wt_nb), st>>>(h->d_ztab, i, w);
        } else {
          k2_select<true>(mt)<<<grid, 128, 0, st>>>(h->d_ztab, i, w);
        }
        ++h->launches;
      }
      if (want_model) {
        k2_select<false>(mt)<<<grid, 128, 0, st>>>(h->d_ztab, i, w);
        ++h->launches;
      }
    };
    if (want_chi2 || want_model) {
      if (same_window_shape) {
        window_launch(0, -1, n_zl, nA_max);          // one grid over (point tiles, theta_A, z)
      } else {
        for (int i = 0; i < n_zl; ++i) window_launch(i, i, 1, tab[i].n_A);
      }
    }
    if (want_px)
      for (int i = 0; i < n_zl; ++i) {
        k_unpad_pxzam<<<std::min(4 * h->sm_count, 1024), 256, 0, st>>>(h->d_ztab, i, i, n_zl, na_max, nm_max, count, px_dst);
        ++h->launches;
      }
    // ---- stage 3: reduce ----
    if ((rc = mark())) return rc;
    if (a->chi2) {
      double* dst = dev_io ? a->chi2 + first : h->d_chi2;
      k_reduce<<<(count + 255) / 256, 256, 0, st>>>(chi2_zA_dst, dst, count, n_zl, nA_max, h->d_nA_of_z);
      ++h->launches;
    }
    if ((rc = mark())) return rc;
    CPX_CUDA(cudaGetLastError());
    if (!dev_io) {
      if (a->chi2)
        CPX_CUDA(cudaMemcpyAsync(a->chi2 + first, h->d_chi2, sizeof(double) * count, cudaMemcpyDeviceToHost, st));
      if (a->chi2_zA)
        CPX_CUDA(cudaMemcpyAsync(a->chi2_zA + first * n_zl * nA_max, h->d_chi2_zA,
                                 sizeof(double) * count * n_zl * nA_max, cudaMemcpyDeviceToHost, st));
      if (a->model_zAM)
        CPX_CUDA(cudaMemcpyAsync(a->model_zAM + first * n_zl * nA_max * nM_max, h->d_model,
                                 sizeof(double) * count * n_zl * nA_max * nM_max, cudaMemcpyDeviceToHost, st));
      if (a->px_zam)
        CPX_CUDA(cudaMemcpyAsync(a->px_zam + first * n_zl * na_max * nm_max, h->d_pxout,
                                 sizeof(double) * count * n_zl * na_max * nm_max, cudaMemcpyDeviceToHost, st));
    }
  }
  if (!dev_io || a->stage_ms) {
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return fail(CPX_ERR_CUDA, std::string("cpx_eval: ") + cudaGetErrorString(e));
  }
  if (a->stage_ms) {
    for (int s = 0; s < 4; ++s) a->stage_ms[s] = 0.f;
    for (size_t i = 0; i + 4 < evs.size(); i += 5)   // five marks per chunk
      for (int s = 0; s < 4; ++s) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, evs[i + s], evs[i + s + 1]);
        a->stage_ms[s] += ms;
      }
    for (cudaEvent_t ev : evs) cudaEventDestroy(ev);
  }
  return CPX_OK;
}

int cpx_hankel_jbar(int32_t device, int32_t n_groups, const int32_t* group_start, const double* r_Mpc, const double* weight,
                    int64_t n_f, const double* kf, const double* meas, double* jbar) {
  if (n_groups <= 0 || !group_start || !r_Mpc || !weight || n_f <= 0 || !kf || !meas || !jbar)
    return fail(CPX_ERR_ARG, "cpx_hankel_jbar: bad arguments");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev)
    return fail(CPX_ERR_CUDA, "cpx_hankel_jbar: no such CUDA device");
  const int n_r = group_start[n_groups];
  for (int g = 0; g < n_groups; ++g)
    if (group_start[g + 1] <= group_start[g]) return fail(CPX_ERR_ARG, "cpx_hankel_jbar: empty or unsorted group");
  CPX_CUDA(cudaSetDevice(device));
  int* d_gs = nullptr;
  double *d_r = nullptr, *d_w = nullptr, *d_kf = nullptr, *d_meas = nullptr, *d_out = nullptr;
  auto cleanup = [&]() {
    cudaFree(d_gs); cudaFree(d_r); cudaFree(d_w); cudaFree(d_kf); cudaFree(d_meas); cudaFree(d_out);
  };
  cudaError_t e = cudaSuccess;
  auto up = [&](void** dst, const void* src, size_t bytes) {
    if (e != cudaSuccess) return;
    e = cudaMalloc(dst, bytes);
    if (e == cudaSuccess) e = cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice);
  };
  up(reinterpret_cast<void**>(&d_gs), group_start, sizeof(int) * (n_groups + 1));
  up(reinterpret_cast<void**>(&d_r), r_Mpc, sizeof(double) * n_r);
  up(reinterpret_cast<void**>(&d_w), weight, sizeof(double) * n_r);
  up(reinterpret_cast<void**>(&d_kf), kf, sizeof(double) * n_f);
  up(reinterpret_cast<void**>(&d_meas), meas, sizeof(double) * n_f);
  if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&d_out), sizeof(double) * n_groups * n_f);
  if (e == cudaSuccess) {
    cpx::k_jbar<<<dim3(static_cast<unsigned>((n_f + 255) / 256), n_groups), 256>>>(d_gs, d_r, d_w, d_kf, d_meas, n_f, d_out);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpy(jbar, d_out, sizeof(double) * n_groups * n_f, cudaMemcpyDeviceToHost);
  cleanup();
  if (e != cudaSuccess) return fail(CPX_ERR_CUDA, std::string("cpx_hankel_jbar: ") + cudaGetErrorString(e));
  return CPX_OK;
}

int cpx_hankel_apply(int32_t device, int32_t n_r, int32_t n_q, int32_t n_k, const double* W, const double* cells, double* px) {
  if (n_r <= 0 || n_q <= 0 || n_k <= 0 || !W || !cells || !px) return fail(CPX_ERR_ARG, "cpx_hankel_apply: bad arguments");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev)
    return fail(CPX_ERR_CUDA, "cpx_hankel_apply: no such CUDA device");
  CPX_CUDA(cudaSetDevice(device));
  double *d_W = nullptr, *d_c = nullptr, *d_px = nullptr;
  const size_t bW = sizeof(double) * n_r * static_cast<size_t>(n_q), bc = sizeof(double) * n_k * static_cast<size_t>(n_q),
               bp = sizeof(double) * n_r * static_cast<size_t>(n_k);
  cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&d_W), bW);
  if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&d_c), bc);
  if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&d_px), bp);
  if (e == cudaSuccess) e = cudaMemcpy(d_W, W, bW, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(d_c, cells, bc, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    cpx::k_hankel_apply<<<dim3((n_k + 63) / 64, (n_r + 31) / 32), 128>>>(d_W, d_c, n_r, n_q, n_k, d_px);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpy(px, d_px, bp, cudaMemcpyDeviceToHost);
  cudaFree(d_W); cudaFree(d_c); cudaFree(d_px);
  if (e != cudaSuccess) return fail(CPX_ERR_CUDA, std::string("cpx_hankel_apply: ") + cudaGetErrorString(e));
  return CPX_OK;
}

int cpx_probe_peaks(int32_t device, double* dfma_per_s, double* ffma_per_s, double* mufu_per_s, double* dmma_fma_per_s) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev)
    return fail(CPX_ERR_CUDA, "cpx_probe_peaks: no such CUDA device");
  CPX_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  CPX_CUDA(cudaGetDeviceProperties(&prop, device));
  const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 1 << 15;
  double* dbuf;
  CPX_CUDA(cudaMalloc(reinterpret_cast<void**>(&dbuf), 64));
  cudaEvent_t e0, e1;
  CPX_CUDA(cudaEventCreate(&e0));
  CPX_CUDA(cudaEventCreate(&e1));
  double* outs[4] = {dfma_per_s, ffma_per_s, mufu_per_s, dmma_fma_per_s};
  for (int which = 0; which < 4; ++which) {
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
      CPX_CUDA(cudaEventRecord(e0));
      if (which == 0) cpx::k_probe_dfma<<<blocks, threads>>>(dbuf, iters, 1.0000001);
      if (which == 1) cpx::k_probe_ffma<<<blocks, threads>>>(reinterpret_cast<float*>(dbuf), iters, 1.0000001f);
      if (which == 2) cpx::k_probe_mufu<<<blocks, threads>>>(reinterpret_cast<float*>(dbuf), iters, -0.5f);
      if (which == 3) cpx::k_probe_dmma<<<blocks, threads>>>(dbuf, iters / 8, 1.0000001);
      CPX_CUDA(cudaEventRecord(e1));
      CPX_CUDA(cudaEventSynchronize(e1));
      float ms;
      CPX_CUDA(cudaEventElapsedTime(&ms, e0, e1));
      if (rep > 0) best = std::min(best, ms);
    }
    // ops per launch: 8 chains x iters per thread; a DMMA is 256 FMAs per warp = 8 per thread
    const double per_thread = which == 3 ? 8.0 * (iters / 8) * 8.0 : 8.0 * iters;
    if (outs[which]) *outs[which] = per_thread * static_cast<double>(blocks) * threads / (best * 1e-3);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(dbuf);
  return CPX_OK;
}

int cpx_probe_imma(int32_t device, double* imma_ops_per_s) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev)
    return fail(CPX_ERR_CUDA, "cpx_probe_imma: no such CUDA device");
  CPX_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  CPX_CUDA(cudaGetDeviceProperties(&prop, device));
  const int blocks = prop.multiProcessorCount * 4, threads = 256, iters = 1 << 13;
  int* dbuf;
  CPX_CUDA(cudaMalloc(reinterpret_cast<void**>(&dbuf), 64));
  cudaEvent_t e0, e1;
  CPX_CUDA(cudaEventCreate(&e0));
  CPX_CUDA(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    CPX_CUDA(cudaEventRecord(e0));
    cpx::k_probe_imma<<<blocks, threads>>>(dbuf, iters, 0x01020304u);
    CPX_CUDA(cudaEventRecord(e1));
    CPX_CUDA(cudaEventSynchronize(e1));
    float ms;
    CPX_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0) best = std::min(best, ms);
  }
  // one m16n8k32 = 16 x 8 x 32 multiply-adds = 8192 integer operations per warp
  if (imma_ops_per_s) *imma_ops_per_s = 8192.0 * 8.0 * iters * (static_cast<double>(blocks) * threads / 32.0) / (best * 1e-3);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(dbuf);
  return CPX_OK;
}

}  // extern "C"

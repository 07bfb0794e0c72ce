/*
 * cupix_b200 -- C ABI of the B200 (sm_100a) Px-likelihood engine.
 *
 * Drop-in boundary for the hot path of igmhub/cupix: evaluate the Px(k_par, theta) theory for a
 * batch of parameter points and score each point against a measurement.  Plain C, plain pointers
 * and sizes; no torch / numpy types.  Every entry point returns 0 on success or a negative
 * CPX_ERR_* code and never throws; cpx_last_error() gives the message for the calling thread.
 *
 * What each entry point replaces in the reference (paths relative to the cupix repository):
 *
 *   cpx_create      the one-off work of Theory.__init__ (cupix/likelihood/theory.py:14-46),
 *                   Likelihood.__init__ (cupix/likelihood/likelihood.py:10-34) and of reading the
 *                   measurement arrays of px_data (cupix/px_data/data_DESI_DR2.py:12-62):
 *                   everything that does not depend on the parameter point is uploaded to HBM
 *                   once -- window matrices, weights, covariance (Cholesky-factored here instead
 *                   of inv() per call, likelihood.py:180-185), data vectors, the (k_par, k_perp)
 *                   grid with the linear power sampled on it (theory.py:300) and the Hankel
 *                   weights that stand for forestflow.pcross.Px_Mpc_detailed (theory.py:250-264).
 *   cpx_eval        Likelihood.get_chi2 / get_log_like / get_convolved_px / _compute_Px_Zam
 *                   (likelihood.py:53-149,170-194) and Theory.get_px_obs (theory.py:95-123) for B
 *                   parameter points at once; with grid_n != NULL also the meshgrid construction
 *                   and scatter of the scan scripts (scripts/chi2_scan_mock.py:107-134).
 *   cpx_set_data    swapping the data vector between mocks (scripts/fit_many_mocks.py:151-170
 *                   builds a new DESI_DR2 + Likelihood per mock).
 *   cpx_destroy     object teardown.
 *
 * Kernel selection (environment, read by cpx_create / cpx_eval; results agree to 2e-9 relative either way):
 *   CPX_WTC=0       chi2 stage as the FP64 DMMA GEMM instead of the tcgen05 kernel (k_window_tc, the default for
 *                   calls with at least CPX_WTC_MIN = 49152 (point, z) pairs)
 *   CPX_TC=1        Hankel contraction on tcgen05 (k_p3d_hankel_tc) instead of FP64 DMMA
 *   CPX_SCRATCH_MB  upper limit of the per-chunk Px_Zam scratch buffer (default 6144); the buffer is sized by the
 *                   largest batch seen so far, so an engine that only serves scalar calls stays small
 *   CPX_FACTOR=0    switch the factorised path (see CPX_NO_FACTOR below) off for the whole process
 *   CPX_GUARD=1     debugging: fence every writable scratch buffer with 64 KiB guard bands, checked by
 *                   cpx_get_info(CPX_INFO_GUARD_VIOLATIONS)
 *
 * Parameter slots (a point is a row of `P` doubles; slot[j] says which parameter column j sets,
 * zsel[j] restricts it to one z bin of the handle or -1 for all; parameters not named keep the
 * per-z defaults -- the "defaults overridden by matching keys" rule of
 * cupix/likelihood/model_lya.py:190-196 and model_contaminants.py:103-144):
 */
#ifndef CUPIX_B200_H
#define CUPIX_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
  CPX_BIAS = 0, CPX_BETA = 1, CPX_Q1 = 2, CPX_Q2 = 3, CPX_KV_MPC = 4, CPX_AV = 5, CPX_BV = 6,
  CPX_KP_MPC = 7,                                    /* Arinyo: theory.py:285-308            */
  CPX_B_H = 8, CPX_BETA_H = 9, CPX_L_H_MPC = 10,     /* HCD: theory.py:267-282               */
  CPX_B_X = 11, CPX_BETA_X = 12,                     /* SiIII: theory.py:361-438             */
  CPX_B_NOISE_MPC = 13,                              /* sky: theory.py:476-506               */
  CPX_KC_MPC = 14, CPX_PC = 15,                      /* continuum: theory.py:525-535         */
  CPX_AS = 16, CPX_NS = 17,                          /* primordial amplitude and tilt at fixed background: the linear  */
                                                     /* power of every cell is rescaled by (As/As0) (k/k_pivot)^(ns-ns0), */
                                                     /* As0, ns0 = defaults[] = the cosmology linP_cell was sampled from  */
                                                     /* (theory.py:54-56, RescaledCosmology); needs k_pivot_Mpc > 0       */
  CPX_NSLOTS = 18
};

enum {
  CPX_OK = 0,
  CPX_ERR_ARG = -1,        /* bad argument / inconsistent sizes                      */
  CPX_ERR_CUDA = -2,       /* CUDA runtime error (message has the CUDA string)       */
  CPX_ERR_NOMEM = -3,      /* device or host allocation failed                       */
  CPX_ERR_COV = -4,        /* a covariance block is not positive definite            */
  CPX_ERR_UNSUPPORTED = -5 /* shape outside what the kernels were instantiated for   */
};

enum {
  CPX_INCLUDE_HCD = 1, CPX_INCLUDE_METAL = 2, CPX_INCLUDE_SKY = 4, CPX_INCLUDE_CONTINUUM = 8
};

/* One redshift bin.  All pointers are HOST pointers, borrowed for the duration of cpx_create
 * (contents are copied to the device); all arrays are C-contiguous float64 unless noted. */
typedef struct cpx_zbin_desc {
  int32_t n_a;        /* fine theta bins (or plain theta values for a theory-only bin)         */
  int32_t n_m;        /* fine k_par values                                                      */
  int32_t n_q;        /* k_perp quadrature nodes                                                */
  int32_t n_A;        /* coarse theta bins; 0 => theory-only bin (no window / chi2)             */
  int32_t n_M;        /* coarse k_par bins                                                      */
  int32_t n_data;     /* number of data vectors sharing windows and covariance (>= 1 if n_A)    */
  int32_t flags;      /* CPX_INCLUDE_* (Theory config include_hcd/metal/sky/continuum)          */
  int32_t reserved;
  double dAA_dMpc;    /* Px_AA = Px_Mpc * dAA_dMpc (theory.py:142)                              */
  const double* kpar_Mpc;     /* [n_m]        k_m * dAA_dMpc(z, lambda_Lya)   (theory.py:136)   */
  const double* kperp_Mpc;    /* [n_q]        quadrature nodes                                  */
  const double* hankel_w;     /* [n_a][n_q]   J0 weights incl. k dk/2pi and the theta average   */
  const double* linP_cell;    /* [n_m][n_q]   P_L(z, sqrt(kpar^2+kperp^2))    (theory.py:300)   */
  const double* metal_auto;   /* [3][n_a][n_m] Hankel moments (mu^0,mu^2,mu^4) of                */
                              /*              P_L(z_X) exp(-(k/10)^2) * dAA(z_X)  (theory.py:361-381), or NULL */
  const double* metal_cross;  /* [3][n_a][n_m] same at z_cross, times 2cos(kpar dr) (theory.py:398-438), or NULL */
  const double* sky_xi_a;     /* [n_a]        theta-averaged xi_noise            (theory.py:490) or NULL */
  const double* sky_smooth_m; /* [n_m]        sinc^2(x k/2) * dAA_dMpc           (theory.py:499-504) or NULL */
  double defaults[CPX_NSLOTS];/* per-z default parameter values                                  */
  /* measurement (ignored when n_A == 0) -- the px_data arrays for this z (data_DESI_DR2.py:23-37) */
  const double* B_A_a;        /* [n_A][n_a]   0/1 membership of fine bins in coarse bins         */
  const double* U_aMn;        /* [n_a][n_M][n_m] window matrices                                 */
  const double* V_aM;         /* [n_a][n_M]   theta-rebinning weights                            */
  const double* Px_AM;        /* [n_data][n_A][n_M] data vectors                                 */
  const double* cov_AMM;      /* [n_A][n_M][n_M] covariance blocks                               */
  double k_pivot_Mpc;         /* pivot of the primordial power law for CPX_AS / CPX_NS columns;  */
                              /* 0: the cosmology object does not expose one, such columns are refused */
  const double* kpar_row_Mpc; /* [n_m] or NULL (= kpar_Mpc): k_par seen by the continuum distortion    */
                              /* (theory.py:525-535) when kpar_Mpc carries the reference's k_par = 0 -> 1e-4 */
                              /* substitution for the P3D models (lyaP3D.py:35-39; Theory config kpar0)      */
  const double* metal_auto_dn;  /* [metal_dn_order + 1][3][n_a][n_m] or NULL: Taylor terms of metal_auto in the tilt, */
  const double* metal_cross_dn; /* moments of P_L (ln k / k_pivot)^n / n! (term 0 = metal_auto / metal_cross); with    */
  int32_t metal_dn_order;       /* them CPX_NS may be a batch column together with CPX_INCLUDE_METAL (theory.py:54-56, */
  int32_t reserved2;            /* 361-438), without them that combination returns CPX_ERR_UNSUPPORTED                 */
} cpx_zbin_desc;

typedef struct cpx_handle cpx_handle;

/* Arguments of one batched evaluation.  Input/outputs may live on the host (default; the call
 * copies and synchronises) or on the device (CPX_DEVICE_IO; the call is asynchronous on `stream`).
 * Any output pointer may be NULL. */
enum {
  CPX_DEVICE_IO = 1,      /* points / outputs are device pointers, call is asynchronous on `stream`          */
  CPX_PRECISE_CELLS = 2,  /* evaluate the P3D cells in float64 too (reference-precision mode, ~4x slower):   */
                          /* chi2 then agrees with the float64 oracle to ~1e-9 everywhere */
  CPX_NO_FACTOR = 4,      /* never take the factorised path (below): every point through the full pipeline   */
  CPX_MODEL_WHITENED = 8  /* model_zAM receives L_A^-1 model (the whitened model the chi2 is formed from) instead */
                          /* of the model itself                                                               */
};

/* Factorised path (cpx_factor.cuh; taken automatically, CPX_FACTOR=0 or CPX_NO_FACTOR switch it off).  The model
 * is a polynomial in the AMPLITUDE parameters bias, beta, b_H, beta_H, b_X, beta_X, b_noise_Mpc whose coefficient
 * tables depend on the point only through the other ("tuple") parameters.  When a call asks for `chi2` alone and
 * its points share tuples -- a grid whose points-per-distinct-tuple ratio is >= CPX_FACTOR_MIN (24), explicit
 * points whose columns are all amplitude parameters (one tuple) with B >= CPX_FACTOR_MIN, or explicit HOST points
 * that share tuples at that ratio (a meshgrid handed in as a point list: grouped by tuple on the host) -- the P3D / Hankel /
 * window stages run once per distinct tuple (float64 cells) and chi2 per point is a quadratic form.  This is the
 * shape of the reference's scans (scripts/chi2_scan_mock.py:107-140, chi2_scan_forecast.py:116-125).  Results agree
 * with the float64 oracle to ~1e-9 relative; with the float32-cell pipeline to its own ~1e-7. */

typedef struct cpx_eval_args {
  int64_t B;                 /* number of parameter points                                      */
  int32_t P;                 /* varied parameters per point                                     */
  int32_t flags;             /* CPX_DEVICE_IO | CPX_PRECISE_CELLS                               */
  const double* points;      /* [B][P] row-major; NULL when grid_n is given                     */
  const int32_t* slot;       /* [P] host: CPX_* slot of each column                             */
  const int32_t* zsel;       /* [P] host: z bin index the column applies to, -1 = all; NULL = all -1 */
  /* on-device grid generation (meshgrid indexing='ij', last axis fastest): point i of the full */
  /* grid has coordinate lo[j] + idx_j (hi[j]-lo[j])/(n[j]-1); this call evaluates the B points  */
  /* starting at flat index grid_first (how one rank takes its slice of a sharded scan).          */
  const double* grid_lo;     /* [P] host, or NULL                                               */
  const double* grid_hi;     /* [P] host                                                        */
  const int32_t* grid_n;     /* [P] host                                                        */
  int64_t grid_first;
  const int32_t* data_idx;   /* [B] which data vector each point is scored against; NULL = 0    */
  const int32_t* z_list;     /* [n_z_list] host: z bins to evaluate; NULL = all bins of the handle */
  int32_t n_z_list;
  int32_t reserved;
  double* chi2;              /* [B]                 sum over evaluated z and theta_A             */
  double* chi2_zA;           /* [B][n_z_list][n_A_max]  per-(z, theta_A) chi2 (= -2 log_like_all) */
  double* model_zAM;         /* [B][n_z_list][n_A_max][n_M_max] convolved + rebinned model Px    */
  double* px_zam;            /* [B][n_z_list][n_a_max][n_m_max] unconvolved theory Px_Zam        */
  void* stream;              /* cudaStream_t.  Host I/O: NULL = the handle's own stream (the call synchronises).   */
                             /* CPX_DEVICE_IO: the stream the caller's buffers are ordered on; NULL = the CUDA     */
                             /* default (legacy) stream, as everywhere in CUDA.  Calls on different streams are    */
                             /* ordered against each other by the library (they share the handle's scratch).       */
  float* stage_ms;           /* host [4] or NULL: device time of {params, p3d+hankel, window, reduce} */
                             /* summed over chunks, measured with CUDA events on the launching stream */
} cpx_eval_args;

int cpx_create(const cpx_zbin_desc* zbins, int32_t n_zbins, int32_t device, cpx_handle** out);
int cpx_eval(cpx_handle* h, const cpx_eval_args* args);
int cpx_set_data(cpx_handle* h, int32_t iz, const double* Px_AM, int32_t n_data);
/* n_mocks noisy realisations d = m(point) + L n of the data vector of EVERY z bin of the handle, generated on the device
 * and installed as its data vectors 0 .. n_mocks-1 (what data_idx selects) -- likelihood.py:37-50
 * (generate_px_forecast(add_noise=True)) for many mocks at once; scripts/fit_many_mocks.py:151-170 draws and loads one mock
 * at a time on the host.  `point` / `slot` / `zsel` name the truth like one row of cpx_eval (P may be 0: the defaults).
 * Counter-based Philox generator: the same seed gives the same mocks on any GPU.  mocks_zdAM: host
 * [n_z][n_mocks][n_A_max][n_M_max] to receive the mocks in data space, or NULL. */
int cpx_generate_mocks(cpx_handle* h, int32_t P, const int32_t* slot, const int32_t* zsel, const double* point,
                       int32_t n_mocks, uint64_t seed, double* mocks_zdAM);
void cpx_destroy(cpx_handle* h);
const char* cpx_last_error(void);

/* introspection */
int cpx_version(void);
int cpx_device_count(void);          /* CUDA devices visible to the library, 0 if none (never fails) */
int cpx_get_info(const cpx_handle* h, int32_t what, int64_t* value);
enum { CPX_INFO_N_ZBINS = 0, CPX_INFO_MAX_NA = 1, CPX_INFO_MAX_NM = 2, CPX_INFO_MAX_N_A = 3,
       CPX_INFO_MAX_N_M = 4, CPX_INFO_CHUNK = 5, CPX_INFO_DEVICE_BYTES = 6, CPX_INFO_KERNEL_LAUNCHES = 7,
       CPX_INFO_SM_COUNT = 8,
       CPX_INFO_GUARD_BUFFERS = 9,      /* scratch buffers fenced by guard bands (handles created with CPX_GUARD=1) */
       CPX_INFO_GUARD_VIOLATIONS = 10,  /* guard-band bytes overwritten so far; synchronises the device            */
       CPX_INFO_FACTORED_CALLS = 11,    /* cpx_eval calls that took the factorised path                            */
       CPX_INFO_FACTORED_TUPLES = 12,   /* distinct tuples those calls evaluated (Hankel stage runs once per tuple)*/
       CPX_INFO_CHUNK_ALLOCATED = 13,   /* points the per-chunk scratch is currently sized for (grows on demand    */
                                        /* up to CPX_INFO_CHUNK)                                                    */
       CPX_INFO_FACTORED_CACHE_HITS = 14,/* single-tuple factorised calls that re-used the previous call's Gram data */
                                        /* (same tuple values, columns, z list and data vectors): one kernel launch */
       CPX_INFO_HANKEL_KERNEL = 15      /* Hankel kernel of the general (float32-cell) pipeline: 0 = k_p3d_hankel (FP64 */
                                        /* DMMA, the default), 1 = k_p3d_hankel_tc (tcgen05 INT8, CPX_TC=1), 2 =         */
                                        /* k_p3d_hankel_imma (mma.sync INT8, CPX_IMMA=1; 52, 56 or 88 k_perp nodes)      */ };

/* Set-up helper (no handle needed): the J0 part of the Hankel weights that stand for forestflow.pcross.Px_Mpc_detailed
 * (theory.py:250-264) with the theta sub-sampling average of likelihood.py:99-130 folded in.  For n_groups fine theta
 * bins whose separations r_Mpc[group_start[g] .. group_start[g+1]) carry the weights weight[] (theta itself), and a fine
 * k_perp grid kf[n_f] with measure meas[n_f]:
 *     jbar[g][f] = meas[f] * sum_i weight_i J0(kf[f] r_i) / sum_i weight_i               (host arrays in, host array out)
 * cupix_b200/quadrature.py projects jbar onto the cardinal splines of the node set; this call replaces its NumPy loop
 * (2000 separations x 75 000 grid points per z bin at the stress shape: seconds per z on the host, milliseconds here). */
int cpx_hankel_jbar(int32_t device, int32_t n_groups, const int32_t* group_start, const double* r_Mpc, const double* weight,
                    int64_t n_f, const double* kf, const double* meas, double* jbar);

/* The Hankel contraction alone, for P3D cells the caller evaluated: px[r][k] = sum_q W[r][q] cells[k][q] in FP64 on the
 * device (DMMA).  Replaces the transform inside Theory._compute_px_from_p3d (cupix/likelihood/theory.py:250-264 ->
 * forestflow.pcross.Px_Mpc_detailed) when the P3D is an arbitrary host callable: the caller samples it on the rule's
 * (k_par, k_perp node) grid, W are the rule's weights for the requested separations (cpx_hankel_jbar + spline
 * projection).  Host pointers: W [n_r][n_q], cells [n_k][n_q], px [n_r][n_k]; synchronous. */
int cpx_hankel_apply(int32_t device, int32_t n_r, int32_t n_q, int32_t n_k, const double* W, const double* cells, double* px);

/* Pipe-rate probes for the roofline denominators MEASURED_PEAKS.json does not hold: sustained
 * FP64 FMA (DFMA), FP32 FMA, MUFU.EX2 and FP64 tensor (DMMA m8n8k4, counted in FMAs) rates of this
 * GPU (results in ops/s; an FMA counts as one op).  Any pointer may be NULL. */
int cpx_probe_peaks(int32_t device, double* dfma_per_s, double* ffma_per_s, double* mufu_per_s,
                    double* dmma_fma_per_s);

/* Rate of the INT8 tensor path as mma.sync issues it (m16n8k32, u8 x s8 -> s32; integer operations per second, a
 * multiply-add counts as two): the pipe k_p3d_hankel_imma contracts on. */
int cpx_probe_imma(int32_t device, double* imma_ops_per_s);

#ifdef __cplusplus
}
#endif
#endif /* CUPIX_B200_H */

"""CPU ORACLE (test infrastructure, NOT product code) -- float64 NumPy restatement of the cupix
Px-likelihood hot path.  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline /
reference arm may import this module; nothing under cupix_b200/ does.

Every function cites the reference lines it restates (paths relative to /root/reference).

Parity status
-------------
* theory.py orchestration / P3D / contaminant formulas, likelihood.py theta-average, window,
  rebin and chi2: PINNED -- tests/golden/make_golden.py imports the unmodified reference modules
  (with stand-ins for the absent third-party packages lace / forestflow / astropy) and stores
  their outputs; tests/test_oracle_golden.py checks this restatement against them.
* the P3D -> Px Hankel transform (``forestflow.pcross.Px_Mpc_detailed``, third-party, un-pinned,
  absent): PINNED TO THE REFERENCE'S IN-TREE COPY of that algorithm.  tests/golden/make_golden_hankel.py
  runs cupix/likelihood/lyaP3D.py (Px_Mpc_withSiIII :127-362, P1D_Mpc :66-106, "copied from ForestFlow")
  unmodified, with hankl.FFTLog supplied by scipy.fft.fht, and stores its Px for four Arinyo parameter
  sets; tests/test_oracle_hankel_golden.py checks the FFTLog restatement in oracle/hankel.py against it
  (1e-10) and the fixed-node rule the device evaluates (<= 1e-5 of the Px scale for r >= 0.3 Mpc).  What
  stays unpinned is only the difference between that in-tree copy and today's ForestFlow HEAD (the
  ``max_k_for_p3d`` argument of theory.py:264, which the in-tree copy does not have).  Round 2 added fixtures with a
  BAO-like wiggle in the linear power and the reference's k_par = 0 -> 1e-4 row (lyaP3D.py:35-39, executed by the
  generator), and profiles/transform_choice_r02.txt states what the choice of transform costs in chi2.
* the posterior layer (posterior.py:39-84, free_parameter.py:34-50): PINNED -- tests/golden/make_golden_posterior.py runs
  the unmodified reference classes on the reference Likelihood; tests/test_posterior_golden.py checks this oracle's
  log-likelihood under cupix_b200's Posterior against them.
"""
import numpy as np

LR_LYA = 1215.67      # cupix/likelihood/model_lya.py:30
LR_METAL = 1206.52    # cupix/likelihood/model_contaminants.py:12
KT_MPC_MAX = 200.0    # cupix/likelihood/theory.py:320,356,393,450


class OracleTheory(object):
    """Restates ``cupix.likelihood.theory.Theory`` for a fixed cosmology object.

    cosmo      object with get_linP_Mpc / get_darc_dMpc / get_dAA_dMpc (theory.py:132-136,300)
    lya_defaults   dict bias, beta, q1, q2, kv_Mpc, av, bv, kp_Mpc        (model_lya.py:175-198)
    cont_defaults  dict b_H, beta_H, L_H_Mpc, b_X, beta_X, b_noise_Mpc, kC_Mpc, pC
                                                                  (model_contaminants.py:103-144)
    xi_noise   (theta_arc[], xi[]) table, linearly interpolated (model_contaminants.py:147-165)
    hankel     callable(rt_Mpc[Nr], kp_Mpc[Nk], p3d_func(k, mu)) -> [Nr, Nk]
               (stands where theory.py:250-264 calls forestflow.pcross.Px_Mpc_detailed)
    """

    def __init__(self, z, cosmo, lya_defaults, cont_defaults, xi_noise, hankel,
                 include_hcd=False, include_metal=False, include_sky=False, include_continuum=False):
        self.z = z
        self.cosmo = cosmo
        self.lya_defaults = dict(lya_defaults)
        self.cont_defaults = dict(cont_defaults)
        self.xi_theta, self.xi_val = xi_noise
        self.hankel = hankel
        self.include_hcd = include_hcd
        self.include_metal = include_metal
        self.include_sky = include_sky
        self.include_continuum = include_continuum

    # -- parameter merging: defaults overridden by matching keys (model_lya.py:190-196) -------
    def _merge(self, defaults, params, keys):
        out = {k: defaults[k] for k in keys}
        for k in keys:
            if k in params:
                out[k] = params[k]
        return out

    def lya_params(self, params):
        return self._merge(self.lya_defaults, params,
                           ("bias", "beta", "q1", "q2", "kv_Mpc", "av", "bv", "kp_Mpc"))

    def cont_params(self, params):
        return self._merge(self.cont_defaults, params,
                           ("b_H", "beta_H", "L_H_Mpc", "b_X", "beta_X", "b_noise_Mpc", "kC_Mpc", "pC"))

    # -- redshifts of the metal terms (theory.py:168-197) --------------------------------------
    def z_metal_auto(self):
        return (1 + self.z) * LR_LYA / LR_METAL - 1

    def z_metal_cross(self):
        return np.sqrt((1 + self.z) * (1 + self.z_metal_auto())) - 1

    # -- P3D models -------------------------------------------------------------------------
    def linP(self, z, k):
        return self.cosmo.get_linP_Mpc(z, k)

    @staticmethod
    def dnl_arinyo(k, mu, linP, lp):
        """theory.py:285-295"""
        delta2 = (1 / (2 * np.pi ** 2)) * k ** 3 * linP
        nonlin = delta2 * (lp["q1"] + lp["q2"] * delta2)
        with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
            vel = (k / lp["kv_Mpc"]) ** lp["av"] * mu ** lp["bv"]
        press = (k / lp["kp_Mpc"]) ** 2
        with np.errstate(over="ignore", invalid="ignore"):
            return np.exp(nonlin * (1 - vel) - press)

    def p3d_lya(self, k, mu, params):
        """theory.py:298-308"""
        linP = self.linP(self.z, k)
        lp = self.lya_params(params)
        p3d = lp["bias"] ** 2 * (1 + lp["beta"] * mu ** 2) ** 2 * linP
        return p3d * self.dnl_arinyo(k, mu, linP, lp)

    def p3d_lya_hcd(self, k, mu, params):
        """theory.py:325-344 with the scale-dependent biases of theory.py:267-282"""
        linP = self.linP(self.z, k)
        lp = self.lya_params(params)
        cp = self.cont_params(params)
        kpar = k * mu
        F_H = np.exp(-kpar * cp["L_H_Mpc"])
        b_T = lp["bias"] + cp["b_H"] * F_H
        beta_T = (lp["bias"] * lp["beta"] + cp["b_H"] * cp["beta_H"] * F_H) / b_T
        p3d = b_T ** 2 * (1 + beta_T * mu ** 2) ** 2 * linP
        return p3d * self.dnl_arinyo(k, mu, linP, lp)

    def p3d_metal_auto(self, k, mu, params):
        """theory.py:361-381"""
        linP = self.linP(self.z_metal_auto(), k)
        cp = self.cont_params(params)
        p3d = cp["b_X"] ** 2 * (1 + cp["beta_X"] * mu ** 2) ** 2 * linP
        return p3d * np.exp(-(k / 10.0) ** 2)

    def p3d_metal_cross(self, k, mu, params):
        """theory.py:398-438"""
        z_cross = self.z_metal_cross()
        linP = self.linP(z_cross, k)
        lp = self.lya_params(params)
        cp = self.cont_params(params)
        p3d = lp["bias"] * cp["b_X"] * (1 + lp["beta"] * mu ** 2) * (1 + cp["beta_X"] * mu ** 2) * linP
        smooth = np.exp(-(k / 10.0) ** 2)
        lr_cross = np.sqrt(LR_LYA * LR_METAL)
        dX_AA = np.log(LR_LYA / LR_METAL) * (1 + z_cross) * lr_cross
        drp_Mpc = dX_AA / self.cosmo.get_dAA_dMpc(z_cross, lambda_rest_AA=lr_cross)
        return p3d * smooth * 2 * np.cos(k * mu * drp_Mpc)

    # -- Px in observed units (theory.py:126-165, 200-247, 455-473, 509-535) --------------------
    def _px_obs(self, theta_arc, k_AA, z_eval, lr, p3d_func):
        darc_dMpc = self.cosmo.get_darc_dMpc(z_eval)
        dAA_dMpc = self.cosmo.get_dAA_dMpc(z_eval, lambda_rest_AA=lr)
        rt_Mpc = np.asarray(theta_arc) / darc_dMpc
        kp_Mpc = np.asarray(k_AA) * dAA_dMpc
        return self.hankel(rt_Mpc, kp_Mpc, p3d_func) * dAA_dMpc

    def px_sky_obs(self, theta_arc, k_AA, params):
        """theory.py:455-506"""
        dAA_dMpc = self.cosmo.get_dAA_dMpc(self.z, lambda_rest_AA=LR_LYA)
        kp_Mpc = np.asarray(k_AA) * dAA_dMpc
        xi = np.interp(np.asarray(theta_arc), self.xi_theta, self.xi_val)
        x = (0.8 / dAA_dMpc) * kp_Mpc / 2.0
        smooth = np.ones_like(x)
        mask = x > 0
        smooth[mask] = (np.sin(x[mask]) / x[mask]) ** 2
        return self.cont_params(params)["b_noise_Mpc"] * np.outer(xi, smooth) * dAA_dMpc

    def continuum_distortion_obs(self, k_AA, params):
        """theory.py:509-535"""
        dAA_dMpc = self.cosmo.get_dAA_dMpc(self.z, lambda_rest_AA=LR_LYA)
        cp = self.cont_params(params)
        return np.tanh((np.asarray(k_AA) * dAA_dMpc / cp["kC_Mpc"]) ** cp["pC"])

    def get_px_obs(self, theta_arc, k_AA, params={}):
        """theory.py:95-123"""
        if self.include_hcd:
            px = self._px_obs(theta_arc, k_AA, self.z, LR_LYA, lambda k, mu: self.p3d_lya_hcd(k, mu, params))
        else:
            px = self._px_obs(theta_arc, k_AA, self.z, LR_LYA, lambda k, mu: self.p3d_lya(k, mu, params))
        if self.include_metal:
            px = px + self._px_obs(theta_arc, k_AA, self.z_metal_auto(), LR_METAL,
                                   lambda k, mu: self.p3d_metal_auto(k, mu, params))
            px = px + self._px_obs(theta_arc, k_AA, self.z_metal_cross(), np.sqrt(LR_METAL * LR_LYA),
                                   lambda k, mu: self.p3d_metal_cross(k, mu, params))
        if self.include_sky:
            px = px + self.px_sky_obs(theta_arc, k_AA, params)
        if self.include_continuum:
            px = px * self.continuum_distortion_obs(k_AA, params)
        return px


# ---- likelihood.py / window_and_rebin.py -----------------------------------------------------

def fine_thetas(theta_min_a, theta_max_a, N):
    """likelihood.py:105-116: N mid-points per fine bin (N == 1: the bin centres, :99-103)."""
    theta_min_a, theta_max_a = np.asarray(theta_min_a), np.asarray(theta_max_a)
    if N == 1:
        return (theta_min_a + theta_max_a) / 2.0
    out = []
    for lo, hi in zip(theta_min_a, theta_max_a):
        d = (hi - lo) / N
        out.append(np.linspace(lo + 0.5 * d, hi - 0.5 * d, N))
    return np.concatenate(out)


def compute_Px_Zam(get_px_obs, theta_min_a, theta_max_a, k_m, params, N_theta_average):
    """likelihood.py:91-132: theory on the fine theta grid, theta-weighted mean per bin."""
    N = N_theta_average
    th = fine_thetas(theta_min_a, theta_max_a, N)
    px_fine = get_px_obs(th, k_m, params)
    if N == 1:
        return px_fine
    w = th.reshape(-1, N)
    return (px_fine.reshape(-1, N, px_fine.shape[1]) * w[:, :, None]).sum(axis=1) / w.sum(axis=1)[:, None]


def convolve_window(U_Mn, model_n):
    """window_and_rebin.py:59-66"""
    return np.matmul(U_Mn, model_n)


def rebin_theta(V_aM, P_aM):
    """window_and_rebin.py:68-76"""
    return np.sum(V_aM * P_aM, axis=0) / np.sum(V_aM, axis=0)


def convolved_px(Px_Zam, B_A_a, U_aMn, V_aM):
    """likelihood.py:53-88 for one redshift bin: window per fine bin, then V-weighted theta rebin."""
    member = np.asarray(B_A_a).astype(bool)
    out = []
    for A in range(member.shape[0]):
        idx = np.where(member[A])[0]
        P = np.array([convolve_window(U_aMn[a], Px_Zam[a].T) for a in idx])
        out.append(rebin_theta(V_aM[idx], P))
    return np.asarray(out)


def log_like(model_AM, data_AM, cov_AMM):
    """likelihood.py:170-194: -1/2 sum_A d^T C^-1 d with an explicit inverse, no log-det term."""
    ll = np.zeros(model_AM.shape[0])
    for A in range(model_AM.shape[0]):
        icov = np.linalg.inv(cov_AMM[A])
        diff = np.squeeze(data_AM[A] - model_AM[A])
        ll[A] = -0.5 * np.dot(np.dot(icov, diff), diff)
    return np.sum(ll), ll


class OracleLikelihood(object):
    """Restates ``cupix.likelihood.likelihood.Likelihood`` (one z bin) on plain arrays."""

    def __init__(self, theory, iz, theta_min_a, theta_max_a, k_m, B_A_a, U_aMn, V_aM,
                 Px_AM, cov_AMM, N_theta_average=100):
        self.theory, self.iz = theory, iz
        self.theta_min_a, self.theta_max_a, self.k_m = theta_min_a, theta_max_a, k_m
        self.B_A_a, self.U_aMn, self.V_aM = B_A_a, U_aMn, V_aM
        self.Px_AM, self.cov_AMM = Px_AM, cov_AMM
        self.N_theta_average = N_theta_average

    @classmethod
    def from_data(cls, data, theory, iz, N_theta_average=100):
        return cls(theory, iz, data.theta_min_a_arcmin, data.theta_max_a_arcmin, data.k_m[iz],
                   data.B_A_a, data.U_ZaMn[iz], data.V_ZaM[iz], data.Px_ZAM[iz], data.cov_ZAM[iz],
                   N_theta_average)

    def Px_Zam(self, params={}):
        return compute_Px_Zam(self.theory.get_px_obs, self.theta_min_a, self.theta_max_a,
                              self.k_m, params, self.N_theta_average)

    def get_convolved_px(self, params={}):
        return convolved_px(self.Px_Zam(params), self.B_A_a, self.U_aMn, self.V_aM)

    def get_log_like(self, params={}, return_info=False):
        ll, ll_all = log_like(self.get_convolved_px(params), self.Px_AM, self.cov_AMM)
        return (ll, {"log_like_all": ll_all}) if return_info else ll

    def get_chi2(self, params={}, return_info=False):
        ll, ll_all = log_like(self.get_convolved_px(params), self.Px_AM, self.cov_AMM)
        return (-2.0 * ll, {"log_like_all": ll_all}) if return_info else -2.0 * ll

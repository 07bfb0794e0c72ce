"""Host-side (float64) construction of everything that is constant across the parameter batch.

For one redshift bin this turns (Theory config, cosmology, theta binning, k_m, measurement) into
the plain arrays of ``cpx_zbin_desc`` (include/cupix_b200.h):

* the (k_par, k_perp) grid and the linear power sampled on it     (theory.py:300 does this per call)
* the folded Hankel weights W[a, q]                               (theory.py:250-264 + likelihood.py:99-130)
* Hankel moments of the SiIII auto / cross terms, which depend on the parameter point only through
  b_X, beta_X, bias, beta *polynomially*  (theory.py:361-438), so they need no per-point quadrature
* the theta-averaged xi_noise and the pixel smoothing of the sky term (theory.py:476-506)
"""
import numpy as np

from cupix_b200.quadrature import KperpRule

SLOT_NAMES = ("bias", "beta", "q1", "q2", "kv_Mpc", "av", "bv", "kp_Mpc",
              "b_H", "beta_H", "L_H_Mpc", "b_X", "beta_X", "b_noise_Mpc", "kC_Mpc", "pC",
              "As", "ns")
SLOT_INDEX = {n: i for i, n in enumerate(SLOT_NAMES)}
# primordial amplitude and tilt at fixed background (theory.py:54-56): batch columns only -- the device
# rescales the linear power of the plan's cosmology by (As/As0) (k/k_pivot)^(ns-ns0) per point.  The
# dict-based scalar API keeps the reference behaviour (a new cosmology object, hence a new plan).
COSMO_SLOT_NAMES = ("As", "ns")
# parameters the model is a polynomial in: the factorised path of cpx_eval (csrc/cpx_factor.cuh) takes them out of the
# Hankel stage; all others define the "tuple" of a point
AMP_SLOT_NAMES = ("bias", "beta", "b_H", "beta_H", "b_X", "beta_X", "b_noise_Mpc")

INCLUDE_HCD, INCLUDE_METAL, INCLUDE_SKY, INCLUDE_CONTINUUM = 1, 2, 4, 8
METAL_KP_MPC = 10.0      # theory.py:378,421
PIXEL_AA = 0.8           # theory.py:495
KPAR0_REFERENCE_MPC = 1e-4   # lyaP3D.py:35-39: what the reference's in-tree Px path substitutes for k_par = 0
# ns as a batch column with SiIII on: the tilt (k / k_pivot)^dn does not factor out of the metal Hankel moments, so they are
# carried as a Taylor series in dn = ns - ns0: table_n = moments of P_L (ln k/k_pivot)^n / n!.  Order 8 keeps the series
# within 1.4e-5 of the metal term (1e-6 of Px) for |dn ln(k/k_pivot)| <= 1.2, i.e. |ns - ns0| <= 0.2 over the k that matter.
METAL_TILT_ORDER = 8


class ZbinPlan(object):
    """Plain-array description of one redshift bin (fields of cpx_zbin_desc)."""
    n_A = 0
    n_M = 0
    n_data = 0
    metal_auto = metal_cross = sky_xi_a = sky_smooth_m = None
    metal_auto_dn = metal_cross_dn = None
    metal_dn_order = 0
    kpar_row_Mpc = None
    B_A_a = U_aMn = V_aM = Px_AM = cov_AMM = None


def fine_theta_grid(theta_min_a, theta_max_a, N):
    """likelihood.py:99-116: bin centres for N == 1, else N mid-points per fine bin.
    Returns (theta[n], group[n])."""
    lo, hi = np.asarray(theta_min_a, dtype=np.float64), np.asarray(theta_max_a, dtype=np.float64)
    if N == 1:
        return 0.5 * (lo + hi), np.arange(lo.size)
    th, grp = [], []
    for i, (a, b) in enumerate(zip(lo, hi)):
        d = (b - a) / N
        th.append(np.linspace(a + 0.5 * d, b - 0.5 * d, N))
        grp.append(np.full(N, i))
    return np.concatenate(th), np.concatenate(grp)


def _cells(kpar, kperp):
    k2 = kpar[:, None] ** 2 + kperp[None, :] ** 2
    k = np.sqrt(k2)
    with np.errstate(divide="ignore", invalid="ignore"):
        mu = np.where(k > 0, kpar[:, None] / np.where(k > 0, k, 1.0), 0.0)
    return k, mu


def _linP_on_cells(cosmo, z, k):
    P = np.zeros_like(k)
    pos = k > 0
    P[pos] = cosmo.get_linP_Mpc(z, k[pos])
    return P


def default_vector(theory, cosmo):
    """The per-z default parameter values in slot order (16 model parameters, then As and ns of
    ``cosmo``; 0 for a cosmology object that does not expose them, which switches the As / ns
    columns off -- see ``k_pivot_of``)."""
    lya = theory.lya_model.get_lya_params(cosmo, {})
    c = theory.cont_model
    vals = dict(lya)
    vals.update(c.get_hcd_params({}))
    vals.update(c.get_metal_params({}))
    vals.update(c.get_sky_params({}))
    vals.update(c.get_continuum_params({}))
    vals["As"] = getattr(cosmo, "As", 0.0) if k_pivot_of(cosmo) > 0 else 0.0
    vals["ns"] = getattr(cosmo, "ns", 0.0) if k_pivot_of(cosmo) > 0 else 0.0
    return np.array([float(vals[n]) for n in SLOT_NAMES])


def k_pivot_of(cosmo):
    """Pivot [1/Mpc] of the primordial power law of ``cosmo`` if the object exposes ``As``, ``ns`` and
    ``k_pivot`` and its linear power follows P_L ~ As (k/k_pivot)^(ns-1) at fixed background (true for
    cupix_b200.cosmology; a foreign object must opt in by carrying the three attributes), else 0."""
    if all(hasattr(cosmo, a) for a in ("As", "ns", "k_pivot")):
        return float(cosmo.k_pivot)
    return 0.0


def weights_device(device):
    """The device on which the J0 sums of the Hankel weights are evaluated (cpx_hankel_jbar), or None where no CUDA
    device is visible: plans are then built with the NumPy implementation -- host-side set-up logic that the CPU tests
    exercise; evaluating a plan always needs the GPU (cpx_create raises without one)."""
    from cupix_b200 import engine
    try:
        return device if engine.device_count() > 0 else None
    except engine.CpxError:
        return None


def build_theory_plan(theory, cosmo, theta_arc, theta_group, k_AA, rule=None, device=None, include=None):
    """Theory part of a plan for separations ``theta_arc`` (arcmin) grouped into bins by
    ``theta_group`` (theta-weighted average inside a group) and wavenumbers ``k_AA``.  ``device``: evaluate the J0 sums
    of the Hankel weights on that GPU (cupix_b200.engine.hankel_jbar) instead of NumPy threads."""
    rule = rule or KperpRule()
    # include = (hcd, metal, sky, continuum) overrides the theory's flags (component accessors of Theory)
    inc_hcd, inc_metal, inc_sky, inc_cont = include if include is not None else (
        theory.include_hcd, theory.include_metal, theory.include_sky, theory.include_continuum)
    jbar_fn = None
    if device is not None:
        from cupix_b200 import engine

        def jbar_fn(start, r, w, kf, meas, device=device):
            return engine.hankel_jbar(device, start, r, w, kf, meas)
    p = ZbinPlan()
    z = theory.z
    theta_arc = np.asarray(theta_arc, dtype=np.float64)
    theta_group = np.asarray(theta_group)
    k_AA = np.asarray(k_AA, dtype=np.float64)
    n_a = int(theta_group.max()) + 1
    lr_lya, lr_metal = theory.lya_model.lr_lya, theory.cont_model.lr_metal

    darc = cosmo.get_darc_dMpc(z)
    dAA = cosmo.get_dAA_dMpc(z, lambda_rest_AA=lr_lya)
    p.n_a, p.n_m, p.n_q = n_a, k_AA.size, rule.n_q
    p.dAA_dMpc = float(dAA)
    # k_par = 0 (every DESI-like k_m grid starts there).  kpar0 = 'exact' (default): the mu = 0 limit of the integrand.
    # kpar0 = 'reference': every P3D model sees k_par = 1e-4 1/Mpc instead, as the reference's in-tree Px path does
    # (lyaP3D.py:35-39); the continuum distortion and the sky term, which the reference evaluates on k_AA directly
    # (theory.py:476-535), keep the true value.
    sub0 = getattr(theory, "kpar0", "exact") == "reference"

    def p3d_kpar(kpar):
        return np.where(kpar == 0, KPAR0_REFERENCE_MPC, kpar) if sub0 else kpar
    p.kpar_Mpc = p3d_kpar(k_AA * dAA)
    p.kpar_row_Mpc = k_AA * dAA if sub0 else None
    p.kperp_Mpc = rule.k.copy()
    p.hankel_w = rule.weights(theta_arc / darc, theta_group, theta_arc, n_a, jbar_fn=jbar_fn)
    k, _ = _cells(p.kpar_Mpc, p.kperp_Mpc)
    p.linP_cell = _linP_on_cells(cosmo, z, k)
    p.defaults = default_vector(theory, cosmo)
    p.k_pivot_Mpc = k_pivot_of(cosmo)
    p.flags = (INCLUDE_HCD * bool(inc_hcd) | INCLUDE_METAL * bool(inc_metal)
               | INCLUDE_SKY * bool(inc_sky) | INCLUDE_CONTINUUM * bool(inc_cont))

    if inc_metal:
        # auto: z_X, lambda_X (theory.py:200-222, 361-381); cross: z_cross, sqrt(l_a l_X) (:225-247, 398-438)
        z_auto = theory.get_z_metal_auto()
        z_cross = theory.get_z_metal_cross()
        lr_cross = np.sqrt(lr_metal * lr_lya)
        tabs, dn_tabs = [], []
        for zz, lr, wiggle in ((z_auto, lr_metal, False), (z_cross, lr_cross, True)):
            darc_x = cosmo.get_darc_dMpc(zz)
            dAA_x = cosmo.get_dAA_dMpc(zz, lambda_rest_AA=lr)
            kpar_x = p3d_kpar(k_AA * dAA_x)
            W_x = rule.weights(theta_arc / darc_x, theta_group, theta_arc, n_a, jbar_fn=jbar_fn)       # [n_a, n_q]
            kx, mux = _cells(kpar_x, rule.k)
            base = _linP_on_cells(cosmo, zz, kx) * np.exp(-(kx / METAL_KP_MPC) ** 2)  # [n_m, n_q]
            mom = np.stack([W_x @ (base * mux ** (2 * j)).T for j in range(3)]) * dAA_x   # [3, n_a, n_m]
            phase = 1.0
            if wiggle:
                dX_AA = np.log(lr_lya / lr_metal) * (1 + zz) * lr
                drp_Mpc = dX_AA / dAA_x
                phase = (2.0 * np.cos(kpar_x * drp_Mpc))[None, None, :]
                mom = mom * phase
            tabs.append(np.ascontiguousarray(mom))
            if p.k_pivot_Mpc > 0:
                # Taylor terms of the tilt: moments of P_L (ln k / k_pivot)^n / n!, n = 0 .. METAL_TILT_ORDER
                with np.errstate(divide="ignore"):
                    lnk = np.where(kx > 0, np.log(np.where(kx > 0, kx, 1.0) / p.k_pivot_Mpc), 0.0)
                series, term = [], base
                for n in range(METAL_TILT_ORDER + 1):
                    series.append(np.stack([W_x @ (term * mux ** (2 * j)).T for j in range(3)]) * dAA_x * phase)
                    term = term * lnk / (n + 1)
                dn_tabs.append(np.ascontiguousarray(np.stack(series)))                     # [order + 1, 3, n_a, n_m]
        p.metal_auto, p.metal_cross = tabs
        if dn_tabs:
            p.metal_auto_dn, p.metal_cross_dn = dn_tabs
            p.metal_dn_order = METAL_TILT_ORDER

    if inc_sky:
        xi = theory.cont_model.get_xi_noise(theta_arc)
        wsum = np.bincount(theta_group, weights=theta_arc, minlength=n_a)
        p.sky_xi_a = np.bincount(theta_group, weights=theta_arc * xi, minlength=n_a) / wsum
        x = (PIXEL_AA / dAA) * (k_AA * dAA) / 2.0
        smooth = np.ones_like(x)
        pos = x > 0
        smooth[pos] = (np.sin(x[pos]) / x[pos]) ** 2
        p.sky_smooth_m = smooth * dAA
    return p


def attach_measurement(p, data, iz):
    """Measurement part of a plan: the px_data arrays of redshift bin ``iz``."""
    p.n_A = len(data.theta_min_A_arcmin)
    p.n_M = len(data.k_M_edges[iz]) - 1
    p.n_data = 1
    p.B_A_a = np.asarray(data.B_A_a, dtype=np.float64)
    p.U_aMn = np.asarray(data.U_ZaMn[iz], dtype=np.float64)
    p.V_aM = np.asarray(data.V_ZaM[iz], dtype=np.float64)
    p.Px_AM = np.asarray(data.Px_ZAM[iz], dtype=np.float64)[None]
    p.cov_AMM = np.asarray(data.cov_ZAM[iz], dtype=np.float64)
    assert p.U_aMn.shape == (p.n_a, p.n_M, p.n_m), "window matrices do not match the binning"
    return p


def build_likelihood_plan(theory, cosmo, data, iz, N_theta_average, rule=None, device=None):
    th, grp = fine_theta_grid(data.theta_min_a_arcmin, data.theta_max_a_arcmin, N_theta_average)
    p = build_theory_plan(theory, cosmo, th, grp, data.k_m[iz], rule, device=device)
    return attach_measurement(p, data, iz)

"""``Theory``: Px(k_par, theta) predictions, same interface as ``cupix/likelihood/theory.py``.

The reference evaluates one parameter point per call in NumPy and delegates the Hankel transform to
ForestFlow.  Here the object only prepares float64 constants (cupix_b200.plan) and hands batches
of parameter points to the CUDA engine (cupix_b200.engine.PxEngine); ``get_px_obs`` is the
batch-of-one special case.  No CPU evaluation path exists.
"""
import numpy as np

from cupix_b200 import cosmology as _cosmology
from cupix_b200 import plan as _plan
from cupix_b200.engine import PxEngine, slots_from_names
from cupix_b200.likelihood.likelihood_parameter import dict_from_likeparam
from cupix_b200.likelihood.model_contaminants import ContaminantsModel
from cupix_b200.likelihood.model_lya import LyaModel
from cupix_b200 import quadrature as _quad
from cupix_b200.quadrature import DEFAULT_K0, DEFAULT_KMAX, PRESETS, KperpRule

COSMO_KEYS = ("H0", "omch2", "ombh2", "mnu", "omk", "As", "ns", "nrun", "w")


class Theory(object):
    """Likelihood will ask this object for Px predictions"""

    def __init__(self, z, fid_cosmo=None, config={'verbose': False}):
        """Inputs (theory.py:14-46):
            - z: redshift of the Lya bin
            - fid_cosmo: fiducial cosmology object (unit conversions, linear power)
            - config: include_hcd / include_metal / include_sky / include_continuum, the LyaModel and
              ContaminantsModel keys, plus engine settings kperp_rule, n_q, k0_Mpc, kmax_Mpc, device."""
        self.verbose = config.get('verbose', False)
        self.z = z
        self.zs = z
        self.fid_cosmo = _cosmology.Cosmology() if fid_cosmo is None else fid_cosmo
        self.lya_model = LyaModel(z, config)
        self.cont_model = ContaminantsModel(z, config)
        self.include_hcd = config.get('include_hcd', False)
        self.include_metal = config.get('include_metal', False)
        self.include_sky = config.get('include_sky', False)
        self.include_continuum = config.get('include_continuum', False)
        # k_perp quadrature (cupix_b200.quadrature).  'kperp_rule': 'auto' (default) looks at the fiducial linear power --
        # baryon wiggles need the refined 'bao' node family, a no-wiggle power the cheaper log-uniform 'smooth' one --
        # and takes the cheapest node set whose measured error against the reference's algorithm is <= px_rtol / 3
        # ('px_rtol', default 1e-5 = the north star's tolerance on model Px).  'smooth' / 'bao' force the family,
        # 'smooth_fine' / 'bao_fine' name the finest sets; n_q / kperp_refine / kperp_k_refine override single fields.
        kind = config.get('kperp_rule', 'auto')
        self.px_rtol = float(config.get('px_rtol', _quad.DEFAULT_PX_RTOL))
        self.wiggle_amplitude = None
        if kind == 'auto':
            self.wiggle_amplitude = _quad.wiggle_amplitude(self.fid_cosmo, z)
            kind = 'bao' if self.wiggle_amplitude > _quad.WIGGLE_THRESHOLD else 'smooth'
        if kind in _quad.RULES:
            preset, self.rule_error, met = _quad.choose_rule(kind, self.px_rtol)
            if not met and 'n_q' not in config:
                import warnings
                warnings.warn("cupix_b200: no '%s' k_perp node set is measured at px_rtol = %.1e / 3; using the finest "
                              "one (error %.1e)" % (kind, self.px_rtol, self.rule_error))
        else:
            preset, self.rule_error = dict(PRESETS[kind]), None
        self.kperp_family = kind
        self.rule = _quad.get_rule(config.get('n_q', preset['n_q']), config.get('k0_Mpc', DEFAULT_K0),
                                   config.get('kmax_Mpc', DEFAULT_KMAX),
                                   refine=config.get('kperp_refine', preset.get('refine', 0.0)),
                                   k_refine=config.get('kperp_k_refine', preset.get('k_refine', 0.3)))
        if self.rule.n_q != preset['n_q'] or self.rule.refine != preset.get('refine', 0.0):
            # a node set named by hand: its measured error if it is one of the listed sets, else unknown
            listed = [e for fam in _quad.RULES.values() for kw, e in fam
                      if kw['n_q'] == self.rule.n_q and kw.get('refine', 0.0) == self.rule.refine and
                      (self.rule.refine == 0.0 or kw.get('k_refine') == self.rule.k_refine)]
            self.rule_error = listed[0] if listed else None
        self.device = config.get('device', 0)
        self.precise_cells = config.get('precise_cells', False)
        # k_par = 0 row: 'exact' = the mu = 0 limit; 'reference' = evaluated at 1e-4 1/Mpc like the reference's in-tree
        # Px path (lyaP3D.py:35-39) -- for comparisons against outputs of that path (cupix_b200.plan.build_theory_plan)
        self.kpar0 = config.get('kpar0', 'exact')
        assert self.kpar0 in ('exact', 'reference'), "kpar0 must be 'exact' or 'reference'"
        self._engines = {}

    # -- cosmology (theory.py:49-64) ---------------------------------------------------------
    def get_cosmology(self, cosmo=None, params={}):
        if cosmo is not None:
            return cosmo
        if not any(k in params for k in COSMO_KEYS):
            return self.fid_cosmo
        if self.fid_cosmo.same_background(cosmo_params=params):
            if isinstance(self.fid_cosmo, _cosmology.Cosmology):
                return _cosmology.RescaledCosmology(self.fid_cosmo, params)
            from lace.cosmo import rescale_cosmology
            return rescale_cosmology.RescaledCosmology(fid_cosmo=self.fid_cosmo, new_params_dict=params)
        if isinstance(self.fid_cosmo, _cosmology.Cosmology):
            return _cosmology.Cosmology(cosmo_params_dict=params)
        from lace.cosmo import cosmology
        return cosmology.Cosmology(cosmo_params_dict=params)

    def get_z_metal_auto(self):
        """theory.py:168-180"""
        return (1 + self.z) * self.lya_model.lr_lya / self.cont_model.lr_metal - 1

    def get_z_metal_cross(self):
        """theory.py:183-197"""
        return np.sqrt((1 + self.z) * (1 + self.get_z_metal_auto())) - 1

    # -- parameters ---------------------------------------------------------------------------
    @staticmethod
    def split_params(params):
        """dict -> (names, values) of the model keys the device path knows; the rest is dropped silently
        (model_lya.py:193-196).  Cosmology keys are not columns here: get_cosmology turns them into a
        cosmology object, as in the reference."""
        names = [k for k in params if k in _plan.SLOT_INDEX and k not in _plan.COSMO_SLOT_NAMES]
        return names, np.array([[float(params[k]) for k in names]], dtype=np.float64).reshape(1, len(names))

    def resolve_params(self, cosmo, params):
        """(names, values [1, P]) of the device columns for a scalar call.  Arinyo mode: the model keys of ``params``.
        IGM (emulator) mode (model_lya.py:175-235): the emulator is evaluated at the IGM keys of ``params`` (mF, gamma,
        sigT_Mpc, kF_Mpc; Delta2_p, n_p from ``cosmo``) and its eight Arinyo values become explicit columns -- the
        reference re-emulates on every call; Arinyo keys in ``params`` are ignored in that mode, as there."""
        if self.lya_model.default_igm_params is None:
            return self.split_params(params)
        lya = self.lya_model.get_lya_params(cosmo, params)
        merged = {k: float(lya[k]) for k in _plan.SLOT_NAMES[:8]}
        for k, v in params.items():
            if k in _plan.SLOT_INDEX and k not in merged and k not in _plan.COSMO_SLOT_NAMES:
                merged[k] = float(v)
        return self.split_params(merged)

    # -- engine cache ---------------------------------------------------------------------------
    def _theta_engine(self, theta_arc, k_AA, cosmo, include=None):
        theta_arc = np.atleast_1d(np.asarray(theta_arc, dtype=np.float64))
        k_AA = np.atleast_1d(np.asarray(k_AA, dtype=np.float64))
        key = (theta_arc.tobytes(), k_AA.tobytes(), _cosmology.cache_key(cosmo, self.fid_cosmo), include)
        if key not in self._engines:
            if len(self._engines) > 8:
                self._engines.pop(next(iter(self._engines)))[0].close()
            p = _plan.build_theory_plan(self, cosmo, theta_arc, np.arange(theta_arc.size), k_AA, self.rule,
                                        device=_plan.weights_device(self.device), include=include)
            # the cosmology object is kept with its engine (see cosmology.cache_key)
            self._engines[key] = (PxEngine([p], device=self.device, precise=self.precise_cells), cosmo)
        return self._engines[key][0]

    # -- predictions --------------------------------------------------------------------------
    def get_px_obs(self, theta_arc, k_AA, cosmo=None, params={}):
        """Px [N_theta, N_k] in Angstrom (theory.py:95-123)."""
        cosmo = self.get_cosmology(cosmo=cosmo, params=params)
        names, values = self.resolve_params(cosmo, params)
        eng = self._theta_engine(theta_arc, k_AA, cosmo)
        slots, zsel = slots_from_names(names)
        return eng.eval(points=values, slots=slots, want=("px_zam",))["px_zam"][0, 0]

    # -- the components of get_px_obs, one by one (theory.py:126-247, 455-535): same device path with the other terms
    #    switched off -- (hcd, metal, sky, continuum) flags of a separate plan -- or, for additive terms, zero bias
    def _component(self, theta_arc, k_AA, cosmo, params, include, force=None):
        cosmo = self.get_cosmology(cosmo=cosmo, params=params)
        names, values = self.resolve_params(cosmo, params)
        for k, v in (force or {}).items():           # forced columns win over params (and over emulated values)
            if k in names:
                values[0, names.index(k)] = v
            else:
                names, values = names + [k], np.concatenate([values, [[v]]], axis=1)
        slots, _ = slots_from_names(names)
        return self._theta_engine(theta_arc, k_AA, cosmo, include=include).eval(points=values, slots=slots, want=("px_zam",))["px_zam"][0, 0]

    def get_px_lya_obs(self, theta_arc, k_AA, cosmo=None, params={}):
        """Clean Lya term (theory.py:126-144)."""
        return self._component(theta_arc, k_AA, cosmo, params, (False, False, False, False))

    def get_px_lya_hcd_obs(self, theta_arc, k_AA, cosmo=None, params={}):
        """Lya with the HCD scale-dependent biases (theory.py:147-165)."""
        return self._component(theta_arc, k_AA, cosmo, params, (True, False, False, False))

    def get_px_metal_auto_obs(self, theta_arc, k_AA, cosmo=None, params={}):
        """SiIII auto term (theory.py:200-222): the metal plan evaluated at bias = 0, where the Lya and cross terms vanish."""
        return self._component(theta_arc, k_AA, cosmo, params, (False, True, False, False), force={"bias": 0.0})

    def get_px_metal_cross_obs(self, theta_arc, k_AA, cosmo=None, params={}):
        """Lya x SiIII cross term (theory.py:225-247): bilinear in (bias, b_X), so total - Lya-only - auto-only."""
        tot = self._component(theta_arc, k_AA, cosmo, params, (False, True, False, False))
        return tot - self.get_px_lya_obs(theta_arc, k_AA, cosmo, params) - self.get_px_metal_auto_obs(theta_arc, k_AA, cosmo, params)

    def get_px_sky_obs(self, theta_arc, k_AA, cosmo=None, params={}):
        """Correlated sky residuals (theory.py:455-473): the sky plan at bias = 0."""
        return self._component(theta_arc, k_AA, cosmo, params, (False, False, True, False), force={"bias": 0.0})

    def get_continuum_distortion_obs(self, k_AA, cosmo=None, params={}):
        """tanh((k_par / kC)^pC) on the given wavenumbers (theory.py:509-535): a closed form of k alone."""
        cosmo = self.get_cosmology(cosmo=cosmo, params=params)
        cp = self.cont_model.get_continuum_params(params)
        kp_Mpc = np.asarray(k_AA, dtype=np.float64) * cosmo.get_dAA_dMpc(self.z, lambda_rest_AA=self.lya_model.lr_lya)
        return np.tanh((kp_Mpc / cp["kC_Mpc"]) ** cp["pC"])

    # -- the P3D models as plain functions of (k, mu) (theory.py:285-344, 361-438), for plots and diagnostics: the device
    #    evaluates the same expressions on its (k_par, k_perp) cell grid and never hands P3D itself out
    def _lya_and_linP(self, k, cosmo, params, z=None):
        cosmo = self.get_cosmology(cosmo=cosmo, params=params)
        return cosmo, self.lya_model.get_lya_params(cosmo, params), cosmo.get_linP_Mpc(self.z if z is None else z, k)

    @staticmethod
    def _dnl(k, mu, linP, lp):
        d2 = k ** 3 * linP / (2.0 * np.pi ** 2)
        growth = d2 * (lp["q1"] + lp["q2"] * d2)
        damping = (k / lp["kv_Mpc"]) ** lp["av"] * mu ** lp["bv"]
        return np.exp(growth * (1.0 - damping) - (k / lp["kp_Mpc"]) ** 2)

    def get_p3d_lya_Mpc(self, k, mu, cosmo=None, params={}):
        _, lp, linP = self._lya_and_linP(k, cosmo, params)
        return (lp["bias"] * (1.0 + lp["beta"] * mu ** 2)) ** 2 * linP * self._dnl(k, mu, linP, lp)

    def get_p3d_lya_hcd_Mpc(self, k, mu, cosmo=None, params={}):
        _, lp, linP = self._lya_and_linP(k, cosmo, params)
        hp = self.cont_model.get_hcd_params(params)
        F = np.exp(-k * mu * hp["L_H_Mpc"])
        b_T = lp["bias"] + hp["b_H"] * F
        bbeta_T = lp["bias"] * lp["beta"] + hp["b_H"] * hp["beta_H"] * F
        return (b_T + bbeta_T * mu ** 2) ** 2 * linP * self._dnl(k, mu, linP, lp)

    def get_p3d_metal_auto_Mpc(self, k, mu, cosmo=None, params={}):
        cosmo = self.get_cosmology(cosmo=cosmo, params=params)
        mp = self.cont_model.get_metal_params(params)
        linP = cosmo.get_linP_Mpc(self.get_z_metal_auto(), k)
        return (mp["b_X"] * (1.0 + mp["beta_X"] * mu ** 2)) ** 2 * linP * np.exp(-(k / _plan.METAL_KP_MPC) ** 2)

    def get_p3d_metal_cross_Mpc(self, k, mu, cosmo=None, params={}):
        cosmo, lp, _ = self._lya_and_linP(k, cosmo, params)
        mp = self.cont_model.get_metal_params(params)
        zc = self.get_z_metal_cross()
        lr_x = np.sqrt(self.lya_model.lr_lya * self.cont_model.lr_metal)
        dr_Mpc = np.log(self.lya_model.lr_lya / self.cont_model.lr_metal) * (1.0 + zc) * lr_x / cosmo.get_dAA_dMpc(zc, lambda_rest_AA=lr_x)
        kaiser = lp["bias"] * mp["b_X"] * (1.0 + lp["beta"] * mu ** 2) * (1.0 + mp["beta_X"] * mu ** 2)
        return kaiser * cosmo.get_linP_Mpc(zc, k) * np.exp(-(k / _plan.METAL_KP_MPC) ** 2) * 2.0 * np.cos(k * mu * dr_Mpc)

    def _compute_lya_hcd_biases(self, kpar, lya_params, hcd_params):
        """(b_T, beta_T) of Lya + HCD along k_par (theory.py:267-282); on the device: k_rowpars."""
        F_H = np.exp(-kpar * hcd_params['L_H_Mpc'])
        b_T = lya_params['bias'] + hcd_params['b_H'] * F_H
        return b_T, (lya_params['bias'] * lya_params['beta'] + hcd_params['b_H'] * hcd_params['beta_H'] * F_H) / b_T

    def _compute_DNL_Arinyo(self, k, mu, linP, lya_params):
        """D_NL of Arinyo-i-Prats et al. 2015 (theory.py:285-295)."""
        return self._dnl(k, mu, linP, lya_params)

    # -- Px in Mpc units (theory.py:311-322, 347-358, 384-395, 441-452, 476-506, 525-535).  The built-in P3D models go
    #    through the fused device path (cells evaluated inside k_p3d_hankel); an arbitrary P3D callable goes through
    #    _compute_px_from_p3d: cells sampled on the host by the callable, contraction on the device.
    def _compute_px_from_p3d(self, rt_Mpc, kp_Mpc, p3d_func_kmu, kt_Mpc_max=DEFAULT_KMAX):
        """Px(k_par, r_perp) [N_r, N_k] in Mpc of an arbitrary ``p3d_func_kmu(k, mu)`` (theory.py:250-264, where ForestFlow's
        ``Px_Mpc_detailed`` does the transform): the callable is sampled on this object's k_perp node set, the Hankel sum
        runs on the GPU (``cpx_hankel_apply``) against weights whose J0 sums also come from the GPU.  P3D is taken as zero
        at k = 0 and for k_perp > kt_Mpc_max; the k_par = 0 row follows ``kpar0`` like every other plan."""
        rt = np.atleast_1d(np.asarray(rt_Mpc, dtype=np.float64))
        kp = np.atleast_1d(np.asarray(kp_Mpc, dtype=np.float64))
        if self.kpar0 == 'reference':
            kp = np.where(kp == 0.0, _plan.KPAR0_REFERENCE_MPC, kp)
        from cupix_b200 import engine as _engine
        if _engine.device_count() <= 0:
            raise _engine.CpxError("cupix_b200: _compute_px_from_p3d needs a CUDA device (no CPU evaluation path exists)")
        dev = self.device
        W = self.rule.weights(rt, jbar_fn=lambda gs, r, w, kf, meas: _engine.hankel_jbar(dev, gs, r, w, kf, meas))
        k, mu = _plan._cells(kp, self.rule.k)
        with np.errstate(all='ignore'):
            cells = np.asarray(p3d_func_kmu(np.where(k > 0, k, 1.0), mu), dtype=np.float64)
        cells = np.where((k > 0) & (self.rule.k[None, :] <= kt_Mpc_max * (1 + 1e-12)), cells, 0.0)
        return _engine.hankel_apply(dev, W, cells)

    def _px_Mpc(self, accessor, z, lambda_rest, rt_Mpc, kp_Mpc, cosmo, params):
        cosmo = self.get_cosmology(cosmo=cosmo, params=params)
        dAA_dMpc = cosmo.get_dAA_dMpc(z, lambda_rest_AA=lambda_rest)
        theta_arc = np.atleast_1d(np.asarray(rt_Mpc, dtype=np.float64)) * cosmo.get_darc_dMpc(z)
        k_AA = np.atleast_1d(np.asarray(kp_Mpc, dtype=np.float64)) / dAA_dMpc
        return accessor(theta_arc, k_AA, cosmo, params) / dAA_dMpc

    def get_px_lya_Mpc(self, rt_Mpc, kp_Mpc, cosmo=None, params={}):
        return self._px_Mpc(self.get_px_lya_obs, self.z, self.lya_model.lr_lya, rt_Mpc, kp_Mpc, cosmo, params)

    def get_px_lya_hcd_Mpc(self, rt_Mpc, kp_Mpc, cosmo=None, params={}):
        return self._px_Mpc(self.get_px_lya_hcd_obs, self.z, self.lya_model.lr_lya, rt_Mpc, kp_Mpc, cosmo, params)

    def get_px_metal_auto_Mpc(self, rt_Mpc, kp_Mpc, cosmo=None, params={}):
        """In the units of the metal redshift (theory.py:200-222)."""
        return self._px_Mpc(self.get_px_metal_auto_obs, self.get_z_metal_auto(), self.cont_model.lr_metal, rt_Mpc, kp_Mpc, cosmo, params)

    def get_px_metal_cross_Mpc(self, rt_Mpc, kp_Mpc, cosmo=None, params={}):
        """In the units of the cross redshift and wavelength (theory.py:225-247)."""
        lr_cross = np.sqrt(self.cont_model.lr_metal * self.lya_model.lr_lya)
        return self._px_Mpc(self.get_px_metal_cross_obs, self.get_z_metal_cross(), lr_cross, rt_Mpc, kp_Mpc, cosmo, params)

    def get_px_sky_Mpc(self, rt_Mpc, kp_Mpc, cosmo=None, params={}):
        return self._px_Mpc(self.get_px_sky_obs, self.z, self.lya_model.lr_lya, rt_Mpc, kp_Mpc, cosmo, params)

    def get_continuum_distortion_Mpc(self, kp_Mpc, params={}):
        """tanh((k_par / kC)^pC) (theory.py:525-535)."""
        cp = self.cont_model.get_continuum_params(params)
        return np.tanh((np.asarray(kp_Mpc, dtype=np.float64) / cp["kC_Mpc"]) ** cp["pC"])

    def get_px_obs_batch(self, theta_arc, k_AA, points, names, cosmo=None):
        """Px [B, N_theta, N_k] for B parameter points (columns named by ``names``)."""
        cosmo = self.fid_cosmo if cosmo is None else cosmo
        eng = self._theta_engine(theta_arc, k_AA, cosmo)
        slots, zsel = slots_from_names(names)
        return eng.eval(points=points, slots=slots, want=("px_zam",))["px_zam"][:, 0]

    def get_px_AA(self, k_AA, theta_arcmin, zs=None, like_params={}, return_arinyo_coeffs=False, verbose=None):
        """Legacy entry point kept for the old Likelihood class (theory.py:68-92)."""
        params_dict = dict_from_likeparam(like_params)
        assert zs == self.z, "Input redshift does not match one in theory"
        return self.get_px_obs(theta_arc=theta_arcmin, k_AA=k_AA, params=params_dict)

"""``Likelihood``: Gaussian chi2 of a Px measurement, same interface as
``cupix/likelihood/likelihood.py``, evaluated on the GPU for whole batches of parameter points.

Scalar calls (``get_chi2(params)``, ``get_log_like``, ``get_convolved_px``) are the batch-of-one case
of the ``*_batch`` / ``*_grid`` methods; all of them go through the C ABI (cupix_b200.engine).
"""
import numpy as np
from scipy.stats.distributions import chi2 as chi2_scipy

from cupix_b200 import cosmology as _cosmology
from cupix_b200 import plan as _plan
from cupix_b200.engine import PxEngine, slots_from_names
from cupix_b200.likelihood.likelihood_parameter import dict_from_likeparam


class Likelihood(object):
    """Likelihood class, holds data, theory, and knows about parameters"""

    def __init__(self, data, theory, iz, config={}):
        """Setup likelihood from theory and data (likelihood.py:10-34):
        - data (required) is the data to model
        - theory (required) instance of Theory
        - iz (required) is the index of the redshift bin to use"""
        self.data = data
        self.theory = theory
        self.iz = iz
        assert self.data.z[iz] == self.theory.z, "inconsistent redshifts"
        self.verbose = config.get('verbose', False)
        self.N_theta_average = config.get('N_theta_average', 100)
        self.device = config.get('device', getattr(theory, 'device', 0))
        # float64 P3D cells as well (reference-precision mode, slower); default: float32 cells
        self.precise_cells = config.get('precise_cells', getattr(theory, 'precise_cells', False))
        self._engines = {}
        self._mocks = None            # data vectors installed by set_mock_data / generate_mocks, re-applied to every engine
        self._warned_mock0 = False
        if self.verbose:
            print("Likelihood will evaluate redshift bin", self.iz, "corresponds to z =", self.theory.z)

    # ---- engine ------------------------------------------------------------------------------
    def build_plan(self, cosmo=None):
        cosmo = self.theory.fid_cosmo if cosmo is None else cosmo
        return _plan.build_likelihood_plan(self.theory, cosmo, self.data, self.iz,
                                           self.N_theta_average, self.theory.rule, device=_plan.weights_device(self.device))

    def engine(self, cosmo=None):
        """Device engine for the current N_theta_average (callers toggle that attribute, e.g.
        notebooks/examples/average_theory.py:267-271) and cosmology."""
        key = (self.N_theta_average, _cosmology.cache_key(cosmo, self.theory.fid_cosmo))
        if key not in self._engines:
            if len(self._engines) > 4:
                self._engines.pop(next(iter(self._engines)))[0].close()
            # the cosmology object is kept with its engine (see cosmology.cache_key)
            eng = PxEngine([self.build_plan(cosmo)], device=self.device, precise=self.precise_cells)
            if self._mocks is not None:       # installed mocks are data: every engine of this likelihood scores against them
                eng.set_data(0, self._mocks)
            self._engines[key] = (eng, cosmo)
        return self._engines[key][0]

    def _scalar_eval(self, params, want):
        if self._mocks is not None and not self._warned_mock0 and any(w.startswith("chi2") for w in want):
            import warnings
            warnings.warn("cupix_b200: mock data vectors are installed; scalar chi2 / log-likelihood calls score against "
                          "mock 0, not the measurement (restore_data() re-installs it)")
            self._warned_mock0 = True
        cosmo = self.theory.get_cosmology(params=params)
        names, values = self.theory.resolve_params(cosmo, params)
        slots, _ = slots_from_names(names)
        return self.engine(cosmo).eval(points=values, slots=slots, want=want)

    # ---- reference API -----------------------------------------------------------------------
    def generate_px_forecast(self, params={}, add_noise=False):
        """Px data vector from the theory and the data window, optionally with noise drawn from the
        data covariance (likelihood.py:37-50)."""
        px_theory = self.get_convolved_px(params=params)
        if add_noise:
            for A in range(px_theory.shape[0]):
                L = np.linalg.cholesky(self.data.cov_ZAM[self.iz, A, :, :])
                px_theory[A] = px_theory[A] + np.dot(L, np.random.normal(size=px_theory[A].shape))
        return px_theory

    def get_convolved_px(self, params={}):
        """Window-convolved, theta-rebinned model [N_A, N_M] (likelihood.py:53-88)."""
        return self._scalar_eval(params, ("model",))["model"][0, 0]

    def _compute_Px_Zam(self, params):
        """Theory per fine bin (theta_a, k_m), theta-averaged [N_a, N_m] (likelihood.py:91-132)."""
        return self._scalar_eval(params, ("px_zam",))["px_zam"][0, 0]

    def _compute_log_like(self, params={}):
        out = self._scalar_eval(params, ("chi2", "chi2_zA"))
        log_like_all = -0.5 * out["chi2_zA"][0, 0]
        return -0.5 * out["chi2"][0], {'log_like_all': log_like_all}

    def get_chi2(self, params={}, return_info=False, like_params=None):
        if like_params is not None:          # legacy keyword (old_likelihood.py:127-141)
            params = dict_from_likeparam(like_params)
        log_like, info = self._compute_log_like(params)
        chi2 = -2.0 * log_like
        return (chi2, info) if return_info else chi2

    def get_log_like(self, params={}, return_info=False):
        log_like, info = self._compute_log_like(params=params)
        return (log_like, info) if return_info else log_like

    def get_probability(self, params={}, n_free_p=0):
        """Probability given the number of degrees of freedom (likelihood.py:152-161)."""
        chi2 = self.get_chi2(params=params, return_info=False)
        return chi2_scipy.sf(chi2, self.get_ndata() - n_free_p)

    def get_ndata(self):
        return self.data.Px_ZAM[self.iz].size

    # ---- batched API (new) -------------------------------------------------------------------
    def get_chi2_batch(self, points, names, return_info=False, data_idx=None, factor=None, out=None):
        """chi2 [B] for ``points`` [B, P] whose columns are the parameters ``names``.  ``factor=False`` keeps every point
        on the full pipeline (see PxEngine._flags); ``out``: float64 array [B] to receive the result."""
        if self.theory.lya_model.default_igm_params is not None:
            assert not return_info, "return_info is not available in IGM (emulator) mode"
            return self._chi2_batch_igm(points, names, data_idx)
        slots, _ = slots_from_names(names)
        want = ("chi2", "chi2_zA") if return_info else ("chi2",)
        out = self.engine().eval(points=points, slots=slots, want=want, data_idx=data_idx, factor=factor, chi2_out=out)
        if return_info:
            return out["chi2"], {'log_like_all': -0.5 * out["chi2_zA"][:, 0]}
        return out["chi2"]

    def _chi2_batch_igm(self, points, names, data_idx=None):
        """IGM (emulator) mode for a batch (model_lya.py:200-235 once per point in the reference): the IGM columns go
        through the emulator network in PyTorch on this likelihood's GPU, its [B, 8] Arinyo output and the remaining
        columns are scored without leaving the device."""
        import torch
        from cupix_b200.likelihood.model_lya import FF_OUTPUT_NAMES, IGM_PARAM_NAMES, LYA_PARAM_NAMES
        lm = self.theory.lya_model
        bad = [n for n in names if n in LYA_PARAM_NAMES]
        if bad:
            raise KeyError("Arinyo parameters %s cannot be columns in IGM (emulator) mode: the emulator sets them" % bad)
        dev = torch.device("cuda", self.device)
        pts = torch.as_tensor(np.ascontiguousarray(points, dtype=np.float64), device=dev)
        igm_cols = [j for j, n in enumerate(names) if n in IGM_PARAM_NAMES]
        other = [j for j, n in enumerate(names) if n not in IGM_PARAM_NAMES]
        ff = lm.emulate_lya_params_tensor(self.theory.fid_cosmo, [names[j] for j in igm_cols], pts[:, igm_cols]).to(torch.float64)
        col = {n: ff[:, i] for i, n in enumerate(FF_OUTPUT_NAMES)}
        kv = torch.exp(torch.log(col["kvav"]) / col["av"])                                   # model_lya.py:13-18
        cols = [col["bias"], col["beta"], col["q1"], col["q2"], kv, col["av"], col["bv"], col["kp"]] + [pts[:, j] for j in other]
        full = torch.stack(cols, dim=1).contiguous()
        slots, zsel = slots_from_names(list(LYA_PARAM_NAMES) + [names[j] for j in other])
        out = torch.empty(full.shape[0], dtype=torch.float64, device=dev)
        didx = None if data_idx is None else torch.as_tensor(np.ascontiguousarray(data_idx, dtype=np.int32), device=dev)
        self.engine().eval_device(full.shape[0], full.shape[1], slots, points_ptr=full.data_ptr(), zsel=zsel,
                                  data_idx_ptr=None if didx is None else didx.data_ptr(), chi2_ptr=out.data_ptr(),
                                  stream=torch.cuda.current_stream(dev).cuda_stream)
        return out.cpu().numpy()

    def get_log_like_batch(self, points, names, data_idx=None):
        return -0.5 * self.get_chi2_batch(points, names, data_idx=data_idx)

    def get_convolved_px_batch(self, points, names):
        slots, _ = slots_from_names(names)
        return self.engine().eval(points=points, slots=slots, want=("model",))["model"][:, 0]

    def get_chi2_grid(self, names, lo, hi, n, first=0, count=None, factor=None):
        """chi2 on the meshgrid(indexing='ij') of linspace(lo[j], hi[j], n[j]); points are generated
        on the device (scripts/chi2_scan_mock.py:107-111).  ``first``/``count`` select a flat slice."""
        slots, _ = slots_from_names(names)
        return self.engine().eval(slots=slots, grid=(lo, hi, n), grid_first=first, B=count, factor=factor)["chi2"]

    def get_chi2_from_arinyo_tensor(self, arinyo, names=("bias", "beta", "q1", "kvav", "av", "bv", "kp", "q2")):
        """Tensor hand-off from the ForestFlow emulator (model_lya.py:200-235): ``arinyo`` is a
        float64 CUDA tensor [B, 8] in ForestFlow's naming (the output of ``predict_Arinyos`` for B
        IGM-parameter points, evaluated in PyTorch).  The kp -> kp_Mpc, kvav -> kv_Mpc = exp(ln(kvav)/av)
        renaming of model_lya.py:13-18 is done on the device and the points are scored without
        leaving the GPU; returns a CUDA tensor chi2 [B] (asynchronous on the current stream)."""
        import torch
        assert arinyo.is_cuda and arinyo.dtype == torch.float64 and arinyo.dim() == 2 and arinyo.shape[1] == len(names)
        assert arinyo.device.index == self.device, "tensor lives on another GPU than this likelihood's engine"
        col = {n: arinyo[:, i] for i, n in enumerate(names)}
        kv = torch.exp(torch.log(col["kvav"]) / col["av"])
        pts = torch.stack([col["bias"], col["beta"], col["q1"], col["q2"], kv, col["av"], col["bv"], col["kp"]], dim=1).contiguous()
        out = torch.empty(pts.shape[0], dtype=torch.float64, device=arinyo.device)
        self.engine().eval_device(pts.shape[0], 8, np.arange(8, dtype=np.int32), points_ptr=pts.data_ptr(),
                                  chi2_ptr=out.data_ptr(), stream=torch.cuda.current_stream(arinyo.device).cuda_stream)
        out._cpx_keepalive = pts          # the kernels read pts asynchronously
        return out

    def set_mock_data(self, Px_AM):
        """Replace the data vector(s) [n_data, N_A, N_M] scored by ``data_idx`` (many-mock fits).  The vectors are kept
        with the likelihood and installed in every engine it has or builds later (another N_theta_average, another
        cosmology in ``params``); ``restore_data()`` goes back to the measurement."""
        Px_AM = np.array(Px_AM, dtype=np.float64)
        self._mocks = Px_AM[None] if Px_AM.ndim == 2 else Px_AM
        self._warned_mock0 = False
        for eng, _ in self._engines.values():
            eng.set_data(0, self._mocks)
        self.engine()

    def restore_data(self):
        """Score against the measurement ``data.Px_ZAM[iz]`` again."""
        self._mocks = None
        for eng, _ in self._engines.values():
            eng.set_data(0, self.data.Px_ZAM[self.iz])

    def generate_mocks(self, n_mocks, params={}, seed=None, install=True):
        """``n_mocks`` noisy realisations ``model(params) + L n`` of the data vector, L L^T = cov per coarse theta
        bin -- ``generate_px_forecast(add_noise=True)`` (likelihood.py:37-50) for many mocks at once.  With ``install``
        (default) they are drawn ON THE DEVICE (cpx_generate_mocks: the whitened model plus unit normals from a
        counter-based Philox generator, written straight into the engine's data-vector table) and become the data
        vectors that ``data_idx = 0 .. n_mocks-1`` selects (scripts/fit_many_mocks.py:151-170 rebuilds data and
        likelihood per mock instead); without, they are drawn on the host with NumPy and only returned.  Seeded either
        way (the two generators give different, equally distributed draws).  Returns [n_mocks, N_A, N_M]."""
        if not install:
            model = self.get_convolved_px(params=params)
            L = np.linalg.cholesky(self.data.cov_ZAM[self.iz])                       # [N_A, N_M, N_M]
            noise = np.random.default_rng(seed).normal(size=(int(n_mocks),) + model.shape)
            return model[None] + np.einsum("aij,naj->nai", L, noise)
        seed = int(np.random.SeedSequence().entropy % (1 << 63)) if seed is None else int(seed)
        cosmo = self.theory.get_cosmology(params=params)
        names, values = self.theory.split_params(params)
        slots, _ = slots_from_names(names)
        mocks = self.engine(cosmo).generate_mocks(n_mocks, seed, point=values, slots=slots)[0]
        self._mocks = mocks
        self._warned_mock0 = False
        for eng, _ in self._engines.values():                 # engines for another binning / cosmology get the same vectors
            if eng is not self.engine(cosmo):
                eng.set_data(0, mocks)
        return mocks


class JointLikelihood(object):
    """Several redshift bins scored in one device call: chi2 = sum over z (full-shape evaluation).
    Parameter names may carry a ``_<iz>`` suffix to apply to one bin only."""

    def __init__(self, likes, device=0, precise_cells=None):
        self.likes = list(likes)
        self.device = device
        self.precise_cells = all(lk.precise_cells for lk in self.likes) if precise_cells is None else precise_cells
        self._engine = None

    def engine(self):
        if self._engine is None:
            self._engine = PxEngine([lk.build_plan() for lk in self.likes], device=self.device, precise=self.precise_cells)
        return self._engine

    def get_ndata(self):
        return sum(lk.get_ndata() for lk in self.likes)

    def get_chi2_batch(self, points, names, return_info=False, data_idx=None, factor=None, out=None):
        """chi2 [B] summed over all z bins for ``points`` [B, P].  ``out``: float64 array [B] (e.g. pinned memory) to
        receive the result instead of a fresh one."""
        slots, zsel = slots_from_names(names)
        want = ("chi2", "chi2_zA") if return_info else ("chi2",)
        out = self.engine().eval(points=points, slots=slots, zsel=zsel, want=want, data_idx=data_idx, factor=factor, chi2_out=out)
        return (out["chi2"], {'log_like_all': -0.5 * out["chi2_zA"]}) if return_info else out["chi2"]

    def get_chi2(self, params={}, return_info=False):
        names = [k for k in params]
        pts = np.array([[params[k] for k in names]], dtype=np.float64).reshape(1, len(names))
        r = self.get_chi2_batch(pts, names, return_info=return_info)
        return (r[0][0], r[1]) if return_info else r[0]

    def get_chi2_grid(self, names, lo, hi, n, first=0, count=None, factor=None):
        slots, zsel = slots_from_names(names)
        return self.engine().eval(slots=slots, zsel=zsel, grid=(lo, hi, n), grid_first=first, B=count, factor=factor)["chi2"]

    def get_convolved_px_batch(self, points, names):
        slots, zsel = slots_from_names(names)
        return self.engine().eval(points=points, slots=slots, zsel=zsel, want=("model",))["model"]

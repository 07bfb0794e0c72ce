"""CPU: host-side logic that needs no GPU -- parameter naming, sharding, posterior, data I/O."""
import os

import numpy as np
import pytest

from cupix_b200 import plan as _plan
from cupix_b200.engine import slots_from_names
from cupix_b200.likelihood.free_parameter import FreeParameter
from cupix_b200.likelihood.likelihood_parameter import LikelihoodParameter, dict_from_likeparam, likeparam_from_dict
from cupix_b200.likelihood.posterior import Posterior
from cupix_b200.parallel import block_partition, scan_chi2_sharded
from cupix_b200.px_data import synthetic
from cupix_b200.px_data.data_DESI_DR2 import DESI_DR2, write_px_file


def test_slot_names_match_header():
    hdr = open(os.path.join(os.path.dirname(__file__), "..", "include", "cupix_b200.h")).read()
    for i, n in enumerate(_plan.SLOT_NAMES):
        assert "CPX_%s = %d" % (n.upper(), i) in hdr, n


def test_slots_from_names():
    s, z = slots_from_names(["bias", "beta_3", "kp_Mpc", "b_noise_Mpc_0"])
    assert list(s) == [0, 1, 7, 13] and list(z) == [-1, 3, -1, 0]
    with pytest.raises(KeyError):
        slots_from_names(["kCb_Mpc"])


def test_block_partition_matches_reference_rule():
    counts, starts = block_partition(16384, 64)
    assert sum(counts) == 16384 and counts[0] == 256 and starts[-1] == 16384 - 256
    counts, starts = block_partition(10, 4)
    assert counts == [3, 3, 2, 2] and starts == [0, 3, 6, 8]


def test_sharded_scan_single_rank_equals_full():
    lo, hi, n = [0.0, 1.0], [1.0, 2.0], [5, 7]
    grid = np.stack(np.meshgrid(np.linspace(0, 1, 5), np.linspace(1, 2, 7), indexing="ij"), -1).reshape(-1, 2)

    def evaluate(first, count):
        return (grid[first:first + count] ** 2).sum(axis=1)
    full = scan_chi2_sharded(None, ["a", "b"], lo, hi, n, evaluate=evaluate)
    parts = [scan_chi2_sharded(None, ["a", "b"], lo, hi, n, rank=r, world_size=3, gather=False, evaluate=evaluate)
             for r in range(3)]
    np.testing.assert_array_equal(np.concatenate(parts), full)


class _FakeLike(object):
    """quadratic log-likelihood with the batched interface of Likelihood"""
    verbose = False
    theory = type("T", (), {"verbose": False})()

    def get_ndata(self):
        return 10

    def get_log_like(self, params={}):
        return -0.5 * sum((v - 1.0) ** 2 for v in params.values())

    def get_log_like_batch(self, values, names, data_idx=None):
        return -0.5 * ((np.asarray(values) - 1.0) ** 2).sum(axis=1)


def test_posterior_scalar_and_batch_agree():
    fp = [FreeParameter("bias", -2, 2, gauss_prior_mean=0.5, gauss_prior_width=0.3), FreeParameter("beta", 0, 3)]
    post = Posterior(_FakeLike(), fp)
    vals = np.array([[0.3, 1.2], [1.0, 1.0], [5.0, 1.0], [0.0, -1.0]])
    batch = post.get_log_posterior_batch(vals)
    for v, b in zip(vals, batch):
        assert post.get_log_posterior_from_values(v) == pytest.approx(b)
    assert batch[2] == -np.inf and batch[3] == -np.inf      # out of bounds is -inf, not the reference's +inf
    assert post.get_ndf() == 8


def test_likelihood_parameter_helpers():
    ps = likeparam_from_dict({"bias": -0.1, "beta": 1.5})
    assert isinstance(ps[0], LikelihoodParameter)
    assert dict_from_likeparam(ps) == {"bias": -0.1, "beta": 1.5}
    assert dict_from_likeparam({"q1": 0.2}) == {"q1": 0.2}


def test_npz_roundtrip(tmp_path):
    d = synthetic.make_measurement_dict("tiny_rebin")
    path = str(tmp_path / "px.npz")
    write_px_file(path, d)
    a, b = DESI_DR2(d), DESI_DR2(path)
    for key in ("Px_ZAM", "cov_ZAM", "U_ZaMn", "V_ZaM", "k_m", "k_M_edges", "B_A_a"):
        np.testing.assert_array_equal(np.asarray(getattr(a, key)), np.asarray(getattr(b, key)))


def test_synthetic_covariance_is_spd():
    d = synthetic.make_measurement_dict("tiny")
    model = np.abs(np.random.default_rng(0).normal(size=d["Px_ZAM"].shape)) + 0.1
    f = synthetic.fill_from_model(d, model, add_noise=True)
    for blk in f["cov_ZAM"].reshape(-1, model.shape[-1], model.shape[-1]):
        assert np.all(np.linalg.eigvalsh(blk) > 0)


def test_ensemble_sampler_recovers_gaussian():
    from cupix_b200.likelihood.sampler import Sampler
    fp = [FreeParameter("bias", -4, 6, ini_value=0.5), FreeParameter("beta", -4, 6, ini_value=0.5)]
    post = Posterior(_FakeLike(), fp)
    smp = Sampler(post, {"nwalkers": 32, "max_nsteps": 600, "nburnin": 150, "seed": 3})
    flat = smp.run_sampler()
    assert 0.2 < smp.acceptance_fraction < 0.9
    assert np.all(np.abs(flat.mean(axis=0) - 1.0) < 0.1)      # _FakeLike: unit Gaussian centred on 1
    assert np.all(np.abs(flat.std(axis=0) - 1.0) < 0.15)
    # the reference's interface (sampler.py:28-107): burn-in + max_nsteps iterations, emcee-style accessors, start points
    es = smp.emcee_sampler
    assert es.iteration == 750 and es.nwalkers == 32 and es.get_chain().shape == (750, 32, 2)
    np.testing.assert_array_equal(es.get_chain(discard=150, flat=True), flat)
    assert es.get_log_prob(discard=150, flat=True).shape == (600 * 32,)
    fp2 = [FreeParameter("bias", -4, 6, ini_value=5.9, delta=0.5), FreeParameter("beta", -4, 6, ini_value=0.5, delta=0.2)]
    s2 = Sampler(Posterior(_FakeLike(), fp2), {"seed": 1})
    w = s2.get_initial_walkers()
    assert w.shape == (10, 2) and s2.nburnin == 100 and s2.max_nsteps == 1000
    assert np.all((w[:, 1] >= 0.5) & (w[:, 1] <= 0.7)) and np.all(w[:, 0] <= 6.0) and np.any(w[:, 0] == 6.0 - 0.005)
    s2.verbose = True
    s2.silence()
    assert s2.verbose is False


def test_cosmology_cache_key_and_branches():
    """theory.py:49-64 branches, and the identity the engine caches use for a cosmology."""
    from cupix_b200.cosmology import Cosmology, RescaledCosmology, cache_key
    from cupix_b200.likelihood.theory import Theory
    fid = Cosmology()
    th = Theory(2.2, fid_cosmo=fid, config={"default_lya_model": "best_fit_arinyo_from_colore"})
    assert th.get_cosmology(params={"bias": -0.1}) is fid
    assert th.get_cosmology(cosmo=fid, params={"ns": 0.9}) is fid         # explicit cosmology wins
    a = th.get_cosmology(params={"ns": 0.95})
    b = th.get_cosmology(params={"ns": 0.95, "bias": -0.1})
    c = th.get_cosmology(params={"ns": 0.96})
    assert isinstance(a, RescaledCosmology) and a is not b
    assert a.ns == 0.95 and a.As == fid.As and a.params["H0"] == fid.params["H0"]
    new = th.get_cosmology(params={"H0": 70.0})
    assert type(new) is Cosmology and new.params["H0"] == 70.0
    assert cache_key(fid, fid) == 0 and cache_key(None, fid) == 0
    assert cache_key(a, fid) == cache_key(b, fid) != cache_key(c, fid)
    assert cache_key(new, fid) not in (0, cache_key(a, fid))

    class Foreign(object):
        pass
    f = Foreign()
    assert cache_key(f, fid) == ("id", id(f))
    # the rescaled linear power is the fiducial one times (As'/As) (k/k_pivot)^(ns'-ns)
    k = np.array([1e-3, 0.05, 0.7, 10.0])
    r = RescaledCosmology(fid, {"As": 2.3e-9, "ns": 0.95})
    np.testing.assert_allclose(r.get_linP_Mpc(2.2, k) / fid.get_linP_Mpc(2.2, k),
                               (2.3e-9 / fid.As) * (k / fid.k_pivot) ** (0.95 - fid.ns), rtol=1e-12)
    assert r.get_darc_dMpc(2.2) == fid.get_darc_dMpc(2.2)


class _QuadraticLike(object):
    """chi2 = (x - mu)^T C^-1 (x - mu) + 7 with the batched and scalar entry points the fit drivers use."""
    verbose = False

    def __init__(self, names, mu, cov, ndata=50):
        self.names, self.mu, self.icov, self.ndata = list(names), np.asarray(mu, float), np.linalg.inv(cov), ndata
        self.n_batched = 0

    def get_chi2_batch(self, points, names, data_idx=None):
        self.n_batched += 1
        d = np.asarray(points)[:, [list(names).index(n) for n in self.names]] - self.mu
        return np.einsum("ni,ij,nj->n", d, self.icov, d) + 7.0

    def get_log_like_batch(self, points, names, data_idx=None):
        return -0.5 * self.get_chi2_batch(points, names)

    def get_chi2(self, params={}, return_info=False):
        c = self.get_chi2_batch(np.array([[params[n] for n in self.names]]), self.names)[0]
        return (c, {"log_like_all": np.array([-0.5 * c])}) if return_info else c

    def get_log_like(self, params={}, return_info=False):
        return -0.5 * self.get_chi2(params)

    def get_probability(self, params={}, n_free_p=0):
        from scipy.stats.distributions import chi2 as chi2_scipy
        return chi2_scipy.sf(self.get_chi2(params), self.ndata - n_free_p)

    def get_ndata(self):
        return self.ndata


def test_minimizer_mirror_on_a_quadratic(tmp_path):
    """minimize_posterior.py interface: best fit, HESSE-like errors (errordef 0.5 on -log P), priors, limits."""
    from cupix_b200.likelihood.minimize_posterior import IminuitMinimizer, Minimizer
    cov = np.array([[4e-4, 1.2e-4], [1.2e-4, 9e-4]])
    like = _QuadraticLike(["bias", "beta"], [-0.12, 1.6], cov)
    pars = [FreeParameter("bias", -0.5, 0.0, ini_value=-0.1, delta=0.02, true_value=-0.12),
            FreeParameter("beta", 0.5, 3.0, ini_value=1.4, delta=0.05)]
    mini = Minimizer(Posterior(like, pars), {"n_starts": 4})
    best = mini.get_best_fit_params()                      # minimises on demand
    assert mini.valid and abs(best["bias"] + 0.12) < 1e-6 and abs(best["beta"] - 1.6) < 1e-6
    np.testing.assert_allclose(mini.covariance, cov, rtol=1e-5)
    v, e = mini.get_best_fit_value("beta", return_hesse=True)
    np.testing.assert_allclose(e, 0.03, rtol=1e-5)
    assert abs(mini.get_best_fit_chi2() - 7.0) < 1e-8 and abs(mini.fmin - 3.5) < 1e-8
    res = mini.get_results_dict()
    assert set(res) == {"bias", "bias_err", "beta", "beta_err", "cov", "prob", "chi2", "ndata", "ndf"}
    assert res["ndf"] == 48 and 0.0 < res["prob"] <= 1.0
    assert like.n_batched < 40                             # a handful of batched stencils, not hundreds of scalar calls
    path = mini.save_results(outfile="fit.npz", outpath=str(tmp_path))
    with np.load(path, allow_pickle=True) as f:
        assert abs(float(f["beta"]) - 1.6) < 1e-6
    # a Gaussian prior pulls the best fit and tightens the error as for independent measurements
    pars_p = [FreeParameter("bias", -0.5, 0.0, ini_value=-0.1, delta=0.02),
              FreeParameter("beta", 0.5, 3.0, ini_value=1.4, delta=0.05, gauss_prior_mean=1.5, gauss_prior_width=0.03)]
    like1 = _QuadraticLike(["bias", "beta"], [-0.12, 1.6], np.diag([4e-4, 9e-4]))
    mp = IminuitMinimizer(like1, pars_p)
    v, e = mp.get_best_fit_value("beta", return_hesse=True)
    np.testing.assert_allclose([v, e], [1.55, 0.03 / np.sqrt(2)], rtol=1e-5)
    # a limit that excludes the minimum: the fit stops at the bound
    pars_l = [FreeParameter("bias", -0.5, 0.0, ini_value=-0.1, delta=0.02),
              FreeParameter("beta", 0.5, 1.5, ini_value=1.4, delta=0.05)]
    ml = Minimizer(Posterior(like1, pars_l))
    assert abs(ml.get_best_fit_value("beta") - 1.5) < 1e-9
    assert mp.minus_log_prob_interface([-0.12, 1.55]) == pytest.approx(3.5 + 0.5 * (0.05 / 0.03) ** 2 * 2, rel=1e-9)


def test_forecast_writer_roundtrip(tmp_path):
    """generate_forecast_data.py:132-173: data copied, P_Z_AM replaced by window-convolved theory (+ seeded noise),
    model parameters stored; the file reads back as a measurement."""
    from cupix_b200.px_data.data_DESI_DR2 import data_to_dict, write_forecast
    d = synthetic.make_measurement_dict("tiny_rebin", Nz=2)
    data = DESI_DR2(d)

    class _Lya(object):
        default_lya_model = "best_fit_arinyo_from_colore"
        default_lya_params = {"bias": -0.115, "beta": 1.5}

    class _Like(object):                     # stands for a Likelihood: the writer only needs these members
        def __init__(self, iz):
            self.iz, self.data = iz, data
            self.theory = type("T", (), {"lya_model": _Lya()})()

        def get_convolved_px(self, params={}):
            return np.full(data.Px_ZAM[self.iz].shape, 0.1 * (self.iz + 1) + params.get("bias", 0.0))

    likes = [_Like(0), _Like(1)]
    path = str(tmp_path / "forecast.npz")
    px = write_forecast(likes, path, params={"bias": -0.01})
    back = DESI_DR2(path)
    np.testing.assert_array_equal(back.Px_ZAM, px)
    np.testing.assert_allclose(back.Px_ZAM[1], 0.19)
    for key in ("cov_ZAM", "U_ZaMn", "V_ZaM", "k_m", "B_A_a"):
        np.testing.assert_array_equal(np.asarray(getattr(back, key)), np.asarray(getattr(data, key)))
    with np.load(path) as f:
        assert str(f["attr_z0_default_lya_model"]) == "best_fit_arinyo_from_colore"
        assert float(f["attr_z1_lya_beta"]) == 1.5 and float(f["attr_z0_forecast_bias"]) == -0.01
        np.testing.assert_array_equal(f["attr_z_centers_orig"], data.z)
    # seeded noise: reproducible, and white after the data's own Cholesky factor
    n1 = write_forecast(likes, str(tmp_path / "n1.npz"), add_noise=True, seed=3)
    n2 = write_forecast(likes, str(tmp_path / "n2.npz"), add_noise=True, seed=3)
    np.testing.assert_array_equal(n1, n2)
    white = np.linalg.solve(np.linalg.cholesky(data.cov_ZAM[0]), (n1[0] - 0.1)[..., None])
    assert 0.5 < white.std() < 1.5
    assert set(data_to_dict(data)) >= {"Px_ZAM", "cov_ZAM", "U_ZaMn", "V_ZaM", "B_A_a", "z_centers"}


def test_igm_mode_reemulates_per_call_and_keeps_igm_keys():
    """IGM (emulator) mode, model_lya.py:175-235: every call re-emulates the Arinyo parameters from the IGM keys of
    ``params``; they are not dropped.  The ForestFlow network is absent here: a mock with its interface."""
    from cupix_b200.cosmology import Cosmology
    from cupix_b200.likelihood.model_lya import LYA_PARAM_NAMES, lya_params_from_forestflow_params
    from cupix_b200.likelihood.theory import Theory
    from tests import helpers as H
    emu = H.MockEmulator()
    cosmo = Cosmology()
    th = Theory(2.4, fid_cosmo=cosmo, config={"default_lya_model": "best_fit_igm_from_gadget", "emulator": emu, "n_q": 56})
    lm = th.lya_model
    assert lm.default_lya_params is None and set(lm.default_igm_params) >= {"mF", "gamma", "sigT_Mpc", "kF_Mpc"}
    base = lm.get_lya_params(cosmo, {})
    moved = lm.get_lya_params(cosmo, {"mF": lm.default_igm_params["mF"] * 0.9, "bias": -5.0})   # Arinyo keys are ignored here
    assert set(base) == set(LYA_PARAM_NAMES) and moved["bias"] != base["bias"] and moved["bias"] != -5.0
    lin = cosmo.get_linP_Mpc_params(z=2.4, kp_Mpc=0.7)
    emu_in = dict(lm.default_igm_params, Delta2_p=lin["Delta2_p"], n_p=lin["n_p"])
    assert base == lya_params_from_forestflow_params(emu.predict_Arinyos(emu_params=emu_in))
    names, values = th.resolve_params(cosmo, {"gamma": 1.7, "b_H": -0.03, "not_a_key": 1.0})
    assert names[:8] == list(LYA_PARAM_NAMES) and names[8:] == ["b_H"] and values.shape == (1, 9)
    assert values[0, 1] == lm.get_lya_params(cosmo, {"gamma": 1.7})["beta"]
    # the plan's defaults are the emulated values at the default IGM point
    from cupix_b200 import plan as _plan
    np.testing.assert_allclose(_plan.default_vector(th, cosmo)[:8], [base[k] for k in LYA_PARAM_NAMES], rtol=0)
    # the batched emulator call (one network evaluation) equals the per-row calls
    import torch
    x = torch.tensor([[0.70, 1.4], [0.80, 1.6], [0.75, 1.5]], dtype=torch.float64)
    ff = lm.emulate_lya_params_tensor(cosmo, ["mF", "gamma"], x)
    for i in range(3):
        one = emu.predict_Arinyos(emu_params=dict(emu_in, mF=float(x[i, 0]), gamma=float(x[i, 1])))
        np.testing.assert_allclose(ff[i].numpy(), [one[k] for k in ("bias", "beta", "q1", "kvav", "av", "bv", "kp", "q2")], rtol=1e-14)


def test_likelihood_parameter_helpers_mirror_the_reference():
    """The legacy parameter helpers of cupix/likelihood/likelihood_parameter.py:4-141 (unit-cube mapping, list <-> dict,
    `_iz` key formatting, look-ups), value by value as the reference defines them."""
    from cupix_b200.likelihood import likelihood_parameter as LP
    p = LP.LikelihoodParameter("bias", min_value=-0.2, max_value=0.0, value=-0.05, Gauss_priors_width=0.3)
    assert p.value_in_cube() == pytest.approx(0.75) and p.get_value_in_cube(-0.1) == pytest.approx(0.5)
    assert p.value_from_cube(0.25) == pytest.approx(-0.15) and p.err_from_cube(0.1) == pytest.approx(0.02)
    q = p.get_new_parameter(0.5)
    assert (q.name, q.min_value, q.max_value, q.Gauss_priors_width) == ("bias", -0.2, 0.0, 0.3) and q.value == pytest.approx(-0.1)
    p.set_without_cube(-0.12)
    assert p.value == -0.12 and p.info_str() == "bias = -0.12" and p.info_str(all_info=True) == "bias = -0.12 , -0.2 , 0.0"
    with pytest.raises(AssertionError):
        p.set_without_cube(0.1)
    assert LP.dict_from_likeparam(None) == {} and LP.likeparam_from_dict(None) == []
    lst = LP.likeparam_from_dict({"bias": -0.1, "beta": 1.5})
    assert [(x.name, x.min_value, x.max_value, x.value) for x in lst] == [("bias", -1000, 1000, -0.1), ("beta", -1000, 1000, 1.5)]
    assert lst[1].value_in_cube() == pytest.approx((1.5 + 1000) / 2000) and LP.dict_from_likeparam(lst) == {"bias": -0.1, "beta": 1.5}
    assert LP.index_by_name(lst, "beta") == 1 and LP.par_index(lst, "bias") == 0 and LP.like_parameter_by_name(lst, "beta") is lst[1]
    with pytest.raises(ValueError):
        LP.par_index(lst, "q1")
    assert LP.format_like_params_dict(2, {"bias": -0.1, "beta_2": 1.5}) == {"bias_2": -0.1, "beta_2": 1.5}
    assert LP.format_like_params_dict([0, 1], {"bias_0": -0.1, "beta_1": 1.5, "q1": 0.3}) == {"bias_0": -0.1, "beta_1": 1.5}


def test_batch_minimizer_stencil_stays_inside_the_bounds():
    """No finite-difference stencil point is evaluated outside [lower, upper] (kp_Mpc <= 0 would be), and a
    fit whose optimum sits on a bound still converges onto it."""
    from cupix_b200.likelihood.batch_minimizer import BatchMinimizer

    class Like(object):
        seen_lo, seen_hi = np.inf, -np.inf

        def get_chi2_batch(self, pts, names, data_idx=None):
            Like.seen_lo, Like.seen_hi = min(Like.seen_lo, pts[:, 1].min()), max(Like.seen_hi, pts[:, 1].max())
            assert np.all(pts[:, 1] > 0.0), "kp evaluated at or below zero"
            return 50.0 * (pts[:, 0] - 1.0) ** 2 + 3.0 * (np.log(pts[:, 1]) - np.log(0.02)) ** 2 + (pts[:, 0] - 1.0) * (pts[:, 1] - 0.02) * 5

    fit = BatchMinimizer(Like(), ["a", "kp"], steps=[0.01, 0.05], lower=[-5.0, 0.04], upper=[5.0, 30.0])
    res = fit.minimize(np.array([[1.3, 0.06], [0.8, 2.0]]))
    assert Like.seen_lo >= 0.04 and Like.seen_hi <= 30.0
    assert np.all(np.abs(res["x"][:, 1] - 0.04) < 0.02) and np.all(np.abs(res["x"][:, 0] - 1.0) < 0.05)


def test_hdf5_layout_round_trip_and_reference_reader(monkeypatch):
    """The HDF5 branches of the reader / writer / forecast writer (data_DESI_DR2.py:66-99 layout) executed against an
    in-memory h5py stand-in (h5py is not installed here): our writer -> our reader, and -- where /root/reference is
    present -- our writer -> the REFERENCE's own DESI_DR2 reader, array for array."""
    import importlib
    import os
    import sys
    import types
    from tests import fake_h5py
    monkeypatch.setitem(sys.modules, "h5py", fake_h5py)
    from cupix_b200.px_data import synthetic
    from cupix_b200.px_data.data_DESI_DR2 import DESI_DR2, write_px_file
    d = synthetic.make_measurement_dict("tiny_rebin")
    rng = np.random.default_rng(2)
    d["Px_ZAM"] = rng.normal(size=d["Px_ZAM"].shape)
    d["cov_ZAM"] = d["cov_ZAM"] * rng.uniform(1, 2, size=d["cov_ZAM"].shape[:2])[:, :, None, None]
    write_px_file("/fake/px_measurement.hdf5", d, extra_attrs={"note": "unit test"})
    a = DESI_DR2("/fake/px_measurement.hdf5")
    b = DESI_DR2(d)
    for name in ("Px_ZAM", "cov_ZAM", "U_ZaMn", "V_ZaM", "B_A_a", "k_m", "k_M_edges", "z", "theta_min_a_arcmin", "theta_max_A_arcmin"):
        np.testing.assert_array_equal(np.asarray(getattr(a, name)), np.asarray(getattr(b, name)), err_msg=name)
    cut = DESI_DR2("/fake/px_measurement.hdf5", theta_min_cut_arcmin=1.0, kM_max_cut_AA=0.2, km_max_cut_AA=0.22)
    assert cut.U_ZaMn.shape == DESI_DR2(d, theta_min_cut_arcmin=1.0, kM_max_cut_AA=0.2, km_max_cut_AA=0.22).U_ZaMn.shape
    # forecast writer, HDF5 flavour (generate_forecast_data.py:132-173): model parameters stored per z next to P_Z_AM
    from cupix_b200.px_data.data_DESI_DR2 import write_forecast

    class _Like(object):
        def __init__(self, iz):
            self.iz, self.data = iz, a
            lm = type("L", (), {"default_lya_model": "best_fit_arinyo_from_colore", "default_lya_params": {"bias": -0.115, "beta": 1.5}})()
            self.theory = type("T", (), {"lya_model": lm})()

        def get_convolved_px(self, params={}):
            return np.full(a.Px_ZAM[self.iz].shape, 0.25 * (self.iz + 1))
    px = write_forecast([_Like(0), _Like(1)], "/fake/forecast.hdf5", params={"beta": 1.6})
    back = DESI_DR2("/fake/forecast.hdf5")
    np.testing.assert_array_equal(back.Px_ZAM, px)
    np.testing.assert_array_equal(back.U_ZaMn, a.U_ZaMn)
    with fake_h5py.File("/fake/forecast.hdf5", "r") as f:
        assert f["P_Z_AM/z_1"].attrs["default_lya_model"] == "best_fit_arinyo_from_colore"
        assert f["P_Z_AM/z_0/lya_params"].attrs["beta"] == 1.5 and bool(f["metadata"].attrs["forecast_add_noise"]) is False
    ref_file = "/root/reference/cupix/px_data/data_DESI_DR2.py"
    if not os.path.exists(ref_file):
        return                                      # the GPU box has no /root/reference; the first half ran
    pkg = types.ModuleType("cupix")
    pkg.__path__ = ["/root/reference/cupix"]
    mpl = types.ModuleType("matplotlib")
    mpl.__path__ = []
    mpl.pyplot = types.ModuleType("matplotlib.pyplot")
    mpl.colors = types.ModuleType("matplotlib.colors")
    mpl.colors.ListedColormap = object
    for name, mod in (("cupix", pkg), ("matplotlib", mpl), ("matplotlib.pyplot", mpl.pyplot), ("matplotlib.colors", mpl.colors)):
        if name == "cupix" or name not in sys.modules:
            monkeypatch.setitem(sys.modules, name, mod)
    for name in ("cupix.px_data", "cupix.px_data.base_px_data", "cupix.px_data.data_DESI_DR2"):
        monkeypatch.delitem(sys.modules, name, raising=False)
    ref = importlib.import_module("cupix.px_data.data_DESI_DR2")
    r = ref.DESI_DR2("/fake/px_measurement.hdf5")
    for ours, theirs in (("Px_ZAM", "Px_ZAM"), ("cov_ZAM", "cov_ZAM"), ("U_ZaMn", "U_ZaMn"), ("V_ZaM", "V_ZaM"), ("B_A_a", "B_A_a"),
                         ("k_m", "k_m"), ("k_M_edges", "k_M_edges"), ("z", "z")):
        np.testing.assert_array_equal(np.asarray(getattr(a, ours)), np.asarray(getattr(r, theirs)), err_msg="reference reader: " + ours)
    for name in list(sys.modules):
        if name == "cupix" or name.startswith("cupix."):
            monkeypatch.delitem(sys.modules, name, raising=False)


def test_p3d_model_functions_equal_the_oracle():
    """Theory.get_p3d_*_Mpc (theory.py:298-308, 325-344, 361-381, 398-438), the P3D models as functions of (k, mu) that
    notebooks plot, against the oracle's restatement of the same reference lines."""
    from tests import helpers as H
    from tests.golden.ref_theory_param_sets import FLAG_SETS
    th = H.make_theory(2.4, "best_fit_arinyo_from_gadget", FLAG_SETS["all"])
    orc = H.oracle_for(th)
    rng = np.random.default_rng(0)
    k, mu = rng.uniform(0.01, 5.0, 64), rng.uniform(0.0, 1.0, 64)
    params = {"bias": -0.13, "beta": 1.4, "q1": 0.5, "b_H": -0.03, "L_H_Mpc": 6.0, "b_X": -0.007, "beta_X": 0.6}
    for ours, theirs in ((th.get_p3d_lya_Mpc, orc.p3d_lya), (th.get_p3d_lya_hcd_Mpc, orc.p3d_lya_hcd),
                         (th.get_p3d_metal_auto_Mpc, orc.p3d_metal_auto), (th.get_p3d_metal_cross_Mpc, orc.p3d_metal_cross)):
        np.testing.assert_allclose(ours(k, mu, params=params), theirs(k, mu, params), rtol=1e-13)


def test_reference_named_helpers_of_the_host_mirrors():
    """Names the reference's scripts and notebooks reach for (data_DESI_DR2.py:66-99 per-dataset getters,
    model_contaminants.py:17-100 default getters, window_and_rebin.py:59-76), on the host side of the boundary."""
    from cupix_b200.likelihood.model_contaminants import ContaminantsModel
    from cupix_b200.likelihood.window_and_rebin import convolve_window, rebin_theta
    from oracle import px_oracle
    d = synthetic.make_measurement_dict("S1", N_A=4, fine_per_coarse=2, N_m=30, rebin_k=3, Nz=2)
    data = DESI_DR2(d, kM_max_cut_AA=0.12)         # cuts change the object, not what the file getters return
    k_m, k_M_edges, tmin_a, tmax_a, tmin_A, tmax_A, zc, N_fft, L_fft, B_A_a = data.read_from_file()
    np.testing.assert_array_equal(k_M_edges, d["k_M_edges"])
    np.testing.assert_array_equal(B_A_a, d["B_A_a"])
    np.testing.assert_array_equal(data.get_Px_z_T(1, 2), d["Px_ZAM"][1][2])
    np.testing.assert_array_equal(data.get_cov_matrix_z_T(0, 3), d["cov_ZAM"][0][3])
    np.testing.assert_array_equal(data.get_window_matrix_z_t(1, 5), d["U_ZaMn"][1][5])
    np.testing.assert_array_equal(data.get_V_Z_aM(0, 7), d["V_ZaM"][0][7])
    assert data.Px_ZAM.shape[-1] < d["Px_ZAM"].shape[-1]

    c22, c30 = ContaminantsModel(2.2, {"b_H": -0.05}), ContaminantsModel(3.0, {"pC": 0.5})
    assert c22.get_default_hcd_params({}) == {"b_H": -0.02, "beta_H": 0.5, "L_H_Mpc": 5.0 / 0.67}
    assert c22.default_hcd_params["b_H"] == -0.05 and c22.get_default_metal_params({"beta_X": 1.0}) == {"b_X": -0.005, "beta_X": 1.0}
    assert c22.get_default_sky_params({})["b_noise_Mpc"] == 0.0035 and c30.get_default_sky_params({})["b_noise_Mpc"] == 0.00125
    assert c22.get_default_continuum_params({}) == {"kC_Mpc": 0.019, "pC": 0.63}
    assert c30.default_continuum_params == {"kC_Mpc": 0.012, "pC": 0.5}

    rng = np.random.default_rng(3)
    U, model, V, P = rng.normal(size=(7, 30)), rng.normal(size=30), rng.uniform(0.5, 1.5, (3, 7)), rng.normal(size=(3, 7))
    np.testing.assert_array_equal(convolve_window(U, model), px_oracle.convolve_window(U, model))
    np.testing.assert_allclose(rebin_theta(V, P), px_oracle.rebin_theta(V, P), rtol=1e-15)

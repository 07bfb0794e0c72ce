"""GPU: the BASELINE.json workload configurations, as parity / property tests on synthetic data.

  S2  validate_pipeline: full theta x k_par Px with window matrices over all z bins vs the oracle
  S3  chi2_scan_forecast: 4-parameter grid -- slices of the 64^4 grid by flat index == explicit points
  S4  fit_many_mocks: many noisy mocks fitted concurrently (lock-step batched minimiser)
  S5  stress shape: size-independent properties at full size (self-consistency, additivity over z,
      batch-order invariance)
"""
import numpy as np
import pytest

from cupix_b200.likelihood.batch_minimizer import BatchMinimizer
from cupix_b200.likelihood.likelihood import JointLikelihood, Likelihood
from cupix_b200.px_data import synthetic
from cupix_b200.px_data.data_DESI_DR2 import DESI_DR2
from tests import helpers as H
from tests.test_gpu_parity import assert_chi2, assert_px

pytestmark = pytest.mark.gpu


def _build(shape, model, flags, N_t, noise=True, seed=synthetic.SEED + 1, **over):
    d = synthetic.make_measurement_dict(shape, **over)
    theories = [H.make_theory(float(z), model, flags) for z in d["z_centers"]]
    data = DESI_DR2(d)
    likes = [Likelihood(data, th, iz, config={"N_theta_average": N_t}) for iz, th in enumerate(theories)]
    joint = JointLikelihood(likes)
    truth_model = joint.get_convolved_px_batch(np.zeros((1, 0)), [])[0]
    filled = synthetic.fill_from_model(d, truth_model, rng=np.random.default_rng(seed), add_noise=noise)
    data = DESI_DR2(filled)
    likes = [Likelihood(data, th, iz, config={"N_theta_average": N_t}) for iz, th in enumerate(theories)]
    return data, likes, JointLikelihood(likes)


def test_S2_validate_pipeline_all_z_vs_oracle():
    data, likes, joint = _build("S2", "best_fit_arinyo_from_gadget", {"include_hcd": True, "include_continuum": True},
                                N_t=2, N_m=60, N_A=8)
    rng = np.random.default_rng(2)
    names = ["bias", "beta", "q1", "kp_Mpc", "b_H"]
    d0 = likes[0].build_plan().defaults
    pts = np.stack([rng.uniform(-0.13, -0.10, 5), rng.uniform(1.3, 1.7, 5), rng.uniform(0.2, 0.5, 5),
                    rng.uniform(10, 16, 5), rng.uniform(-0.03, -0.01, 5)], axis=1)
    chi2, info = joint.get_chi2_batch(pts, names, return_info=True)
    model = joint.get_convolved_px_batch(pts, names)
    orcs = [H.oracle_likelihood(lk) for lk in likes]
    for i in range(pts.shape[0]):
        params = dict(zip(names, pts[i]))
        ref = 0.0
        for iz, o in enumerate(orcs):
            assert_px(model[i, iz], o.get_convolved_px(params), msg="model pt %d z %d" % (i, iz))
            c, inf = o.get_chi2(params, return_info=True)
            assert_chi2(-2 * info["log_like_all"][i, iz], -2 * inf["log_like_all"], "per-theta pt %d z %d" % (i, iz),
                        n_data=likes[iz].get_ndata() / len(inf["log_like_all"]))
            ref += c
        assert_chi2(chi2[i], ref, "joint chi2 pt %d" % i, n_data=joint.get_ndata())


def test_S3_forecast_grid_slices():
    data, likes, joint = _build("S1", "best_fit_arinyo_from_gadget", {}, N_t=1, N_m=48, N_A=6)
    like = likes[0]
    d0 = like.build_plan().defaults
    names = ["bias", "beta", "q1", "kp_Mpc"]
    lo = [d0[0] - 0.02, d0[1] - 0.2, d0[2] - 0.1, d0[7] - 3]
    hi = [d0[0] + 0.02, d0[1] + 0.2, d0[2] + 0.1, d0[7] + 3]
    n = [64, 64, 64, 64]                                     # 16 777 216 points; only slices are evaluated
    total = 64 ** 4
    axes = [np.linspace(lo[j], hi[j], n[j]) for j in range(4)]
    for first, count in ((0, 300), (total // 8 * 3 + 17, 4097), (total - 513, 513)):   # rank-style shards
        chi2 = like.get_chi2_grid(names, lo, hi, n, first=first, count=count, factor=False)
        idx = np.arange(first, first + count)
        sub = np.stack(np.unravel_index(idx, n), axis=1)
        pts = np.stack([axes[j][sub[:, j]] for j in range(4)], axis=1)
        np.testing.assert_array_equal(chi2, like.get_chi2_batch(pts, names))
    assert np.isfinite(chi2).all() and (chi2 > 0).all()


def test_S4_fit_many_mocks_lockstep():
    data, likes, joint = _build("S1", "best_fit_arinyo_from_colore", {}, N_t=1, N_m=48, N_A=8, noise=False)
    like = likes[0]
    truth = like.build_plan().defaults[:2]
    n_mocks = 64
    mocks = like.generate_mocks(n_mocks, seed=9)              # model(defaults) + L n, drawn on the device, installed as data vectors 0..63
    assert mocks.shape == (n_mocks,) + data.Px_ZAM[0].shape
    Lc = np.linalg.cholesky(data.cov_ZAM[0])
    white = np.linalg.solve(Lc, (mocks - data.Px_ZAM[0][None]).transpose(1, 2, 0))
    assert abs(white.std() - 1.0) < 0.02 and abs(white.mean()) < 0.02          # unit Gaussian after whitening
    np.testing.assert_array_equal(like.generate_mocks(n_mocks, seed=9), mocks)                   # seeded: Philox is counter-based
    assert np.abs(like.generate_mocks(n_mocks, seed=10) - mocks).max() > 0
    like.generate_mocks(n_mocks, seed=9)
    # against the host generator (likelihood.py:37-50 with NumPy): same distribution, element by element
    host = like.generate_mocks(4096, seed=3, install=False)
    devm = like.generate_mocks(4096, seed=3)
    sig = np.sqrt(np.einsum("aii->ai", data.cov_ZAM[0]))
    assert np.all(np.abs(host.mean(0) - devm.mean(0)) < 6 * sig / np.sqrt(2048))
    np.testing.assert_allclose(devm.std(0), host.std(0), rtol=0.12)
    wd = np.linalg.solve(Lc, (devm - data.Px_ZAM[0][None]).transpose(1, 2, 0))                  # [A, M, n]
    corr = np.einsum("amn,akn->amk", wd, wd) / wd.shape[-1]                                      # whitened covariance ~ identity
    assert np.abs(corr - np.eye(corr.shape[-1])[None]).max() < 0.1
    # what the device holds is what it returned: chi2 of the truth against mock j == |L^-1 (mock_j - model)|^2
    chi2_dev = like.get_chi2_batch(np.tile(truth, (8, 1)), ["bias", "beta"], data_idx=np.arange(8), factor=False)
    np.testing.assert_allclose(chi2_dev, (wd[:, :, :8] ** 2).sum(axis=(0, 1)), rtol=1e-6)
    mocks = like.generate_mocks(n_mocks, seed=9)
    fitter = BatchMinimizer(like, ["bias", "beta"], steps=[2e-4, 5e-3])
    x0 = np.tile(truth * np.array([1.05, 0.93]), (n_mocks, 1))
    res = fitter.minimize(x0, data_idx=np.arange(n_mocks))
    assert res["converged"].all()
    chi2_truth = like.get_chi2_batch(np.tile(truth, (n_mocks, 1)), ["bias", "beta"], data_idx=np.arange(n_mocks))
    assert np.all(res["chi2"] <= chi2_truth + 1e-6)
    # pulls of the recovered parameters are unit Gaussians (errors from 2 H^-1)
    err = np.sqrt(np.einsum("nii->ni", res["cov"]))
    pull = (res["x"] - truth) / err
    assert np.all(np.abs(pull.mean(axis=0)) < 0.5) and np.all(np.abs(pull.std(axis=0) - 1.0) < 0.35)
    ndof = like.get_ndata() - 2
    assert abs(res["chi2"].mean() / ndof - 1.0) < 0.15
    assert fitter.n_calls < 60            # vs ~100 likelihood calls per mock with MIGRAD, serial


def test_minimizer_mirror_recovers_truth_with_cosmology_axis():
    """BASELINE config 4 shape (fit_true_cont.py): the Minimizer mirror (minimize_posterior.py interface) on a
    noiseless mock recovers the truth; `ns` is a free parameter next to bias and beta."""
    from cupix_b200.likelihood.free_parameter import FreeParameter
    from cupix_b200.likelihood.minimize_posterior import Minimizer
    from cupix_b200.likelihood.posterior import Posterior
    data, likes, joint = _build("S1", "best_fit_arinyo_from_colore", {"include_hcd": True}, N_t=1, N_m=48, N_A=8, noise=False)
    like = likes[0]
    d0 = like.build_plan().defaults
    ns0 = like.theory.fid_cosmo.ns
    pars = [FreeParameter("bias", d0[0] - 0.05, d0[0] + 0.05, ini_value=d0[0] * 1.04, delta=2e-3, true_value=d0[0]),
            FreeParameter("beta", d0[1] - 0.6, d0[1] + 0.6, ini_value=d0[1] * 0.95, delta=5e-2, true_value=d0[1]),
            FreeParameter("ns", ns0 - 0.2, ns0 + 0.2, ini_value=ns0 + 0.03, delta=2e-2, true_value=ns0)]
    mini = Minimizer(Posterior(like, pars), {"n_starts": 3})
    res = mini.get_results_dict()
    assert mini.valid
    for p in pars:
        assert abs(res[p.name] - p.true_value) < 0.05 * res[p.name + "_err"], p.name    # noiseless: far inside the error
        assert 0 < res[p.name + "_err"] < (p.max_value - p.min_value)
    assert res["chi2"] < 1e-3 and res["ndf"] == like.get_ndata() - 3
    np.testing.assert_allclose(res["cov"], res["cov"].T, rtol=1e-9)
    assert np.all(np.linalg.eigvalsh(res["cov"]) > 0)


def test_S5_stress_shape_properties():
    """Full stress size: 12 z x 20 theta x 256 k_par x 64 k_M, dense windows, N_theta_average = 100."""
    data, likes, joint = _build("S5", "best_fit_arinyo_from_gadget", {"include_hcd": True}, N_t=100, noise=False)
    rng = np.random.default_rng(3)
    names = ["bias", "beta", "q1", "kp_Mpc"]
    d0 = likes[0].build_plan().defaults
    B = 4096
    pts = np.stack([rng.uniform(-0.02, 0.02, B), rng.uniform(-0.2, 0.2, B), rng.uniform(-0.1, 0.1, B),
                    rng.uniform(-3, 3, B)], axis=1)
    # self-consistency: the noiseless forecast scored with the parameters that made it has chi2 ~ 0
    assert abs(joint.get_chi2({})) < 1e-6
    # offsets are applied per z bin relative to that bin's defaults -> use per-z names for an exact check
    p0 = np.array([[-0.004, 0.03]])
    tot = joint.get_chi2_batch(p0 + np.array([[likes[0].build_plan().defaults[0], likes[0].build_plan().defaults[1]]]),
                               ["bias_0", "beta_0"], return_info=True)
    # additivity over z: joint chi2 == sum of chi2_zA, and only bin 0 moved
    assert tot[0][0] == pytest.approx((-2 * tot[1]["log_like_all"][0]).sum(), rel=1e-14)
    assert (-2 * tot[1]["log_like_all"][0, 1:]).sum() < 1e-6 and (-2 * tot[1]["log_like_all"][0, 0]).sum() > 1.0
    # batch-order invariance and chunk-boundary independence (B spans several device chunks of 8192? no: force it)
    absolute = pts * 0 + np.array([[d0[0], d0[1], d0[2], d0[7]]]) + pts * np.array([[1, 1, 1, 1]])
    a = joint.get_chi2_batch(absolute, ["bias_3", "beta_3", "q1_3", "kp_Mpc_3"])
    perm = rng.permutation(B)
    b = joint.get_chi2_batch(absolute[perm], ["bias_3", "beta_3", "q1_3", "kp_Mpc_3"])
    np.testing.assert_array_equal(a[perm], b)
    assert (a > 0).all() and np.isfinite(a).all()
    # the single-z Likelihood of bin 3 gives the same numbers as the joint engine restricted to it
    # (64 points of one z take the FP64 DMMA window kernel, the 4096 x 12 batch above the tcgen05 one: the two
    # agree to the ~2^-40 of the integer split, not bit for bit)
    c = likes[3].get_chi2_batch(absolute[:64], names)
    np.testing.assert_allclose(c, a[:64], rtol=1e-9, atol=1e-9)


def test_emulator_tensor_handoff():
    """[B, 8] Arinyo tensor in ForestFlow naming, living on the GPU (what the cINN emulator returns),
    scored through device pointers == the host path with the renamed parameters."""
    import torch
    data, likes, joint = _build("S1", "best_fit_arinyo_from_gadget", {}, N_t=1, N_m=48, N_A=6)
    like = likes[0]
    rng = np.random.default_rng(4)
    B = 300
    ff = np.stack([rng.uniform(-0.14, -0.10, B), rng.uniform(1.3, 1.8, B), rng.uniform(0.1, 0.5, B),
                   rng.uniform(0.3, 0.8, B), rng.uniform(0.2, 0.6, B), rng.uniform(1.5, 1.8, B),
                   rng.uniform(9, 20, B), rng.uniform(0.1, 0.4, B)], axis=1)     # bias beta q1 kvav av bv kp q2
    t = torch.from_numpy(ff).cuda()
    chi2_t = like.get_chi2_from_arinyo_tensor(t)
    assert chi2_t.is_cuda and chi2_t.shape == (B,)
    kv = np.exp(np.log(ff[:, 3]) / ff[:, 4])
    host = like.get_chi2_batch(np.stack([ff[:, 0], ff[:, 1], ff[:, 2], ff[:, 7], kv, ff[:, 4], ff[:, 5], ff[:, 6]], axis=1),
                               ["bias", "beta", "q1", "q2", "kv_Mpc", "av", "bv", "kp_Mpc"])
    torch.cuda.synchronize()
    np.testing.assert_allclose(chi2_t.cpu().numpy(), host, rtol=1e-9)
    # a large batch produced and consumed on torch's default stream with no global synchronisation in between: the
    # engine runs on the stream it is given (handle 0 = the legacy default stream) and torch's kernels on either side
    # are ordered with it (the handle's private non-blocking stream was not)
    Bb = 60000
    big = torch.from_numpy(np.tile(ff, (Bb // B, 1))).cuda()
    big = big * (1.0 + 1e-3 * torch.sin(torch.arange(Bb, device=big.device, dtype=torch.float64))[:, None])     # device-side producer
    out = like.get_chi2_from_arinyo_tensor(big)
    total = float(out.sum())                                       # device-side consumer, then the D2H of a scalar
    bn = big.cpu().numpy()
    kvb = np.exp(np.log(bn[:, 3]) / bn[:, 4])
    ref = like.get_chi2_batch(np.stack([bn[:, 0], bn[:, 1], bn[:, 2], bn[:, 7], kvb, bn[:, 4], bn[:, 5], bn[:, 6]], axis=1),
                              ["bias", "beta", "q1", "q2", "kv_Mpc", "av", "bv", "kp_Mpc"])
    assert total == pytest.approx(ref.sum(), rel=1e-10)


def test_igm_emulator_mode_end_to_end():
    """IGM-parameter mode (model_lya.py:200-235) with a mock emulator in ForestFlow's interface: the scalar calls
    re-emulate per call (chi2 moves with mF), agree with the oracle evaluated at the emulated Arinyo parameters, and the
    batched call (emulator network in torch on the GPU -> [B, 8] tensor -> cpx_eval with device pointers) equals them."""
    from cupix_b200.cosmology import Cosmology
    from cupix_b200.likelihood.theory import Theory
    from oracle.px_oracle import OracleLikelihood
    emu = H.MockEmulator()
    d = synthetic.make_measurement_dict("S1", N_A=6, N_m=40, z0=2.4)
    cosmo = Cosmology()
    th = Theory(2.4, fid_cosmo=cosmo, config={"default_lya_model": "best_fit_igm_from_gadget", "emulator": emu, "n_q": 96,
                                              "include_hcd": True})
    like = Likelihood(DESI_DR2(d), th, 0, config={"N_theta_average": 2})
    filled = synthetic.fill_from_model(d, like.get_convolved_px({})[None], add_noise=True)
    like = Likelihood(DESI_DR2(filled), th, 0, config={"N_theta_average": 2})
    igm0 = th.lya_model.default_igm_params
    c0 = like.get_chi2({})
    c1 = like.get_chi2({"mF": igm0["mF"] * 0.97})
    assert abs(c1 - c0) > 1.0, "IGM keys must not be dropped"
    # oracle: an Arinyo-mode theory with the emulated values as explicit parameters
    th_a = H.make_theory(2.4, "best_fit_arinyo_from_gadget", {"include_hcd": True}, cosmo=cosmo)
    orc = OracleLikelihood.from_data(like.data, H.oracle_for(th_a), 0, 2)
    for params in ({}, {"mF": igm0["mF"] * 0.97}, {"gamma": igm0["gamma"] + 0.1, "kF_Mpc": igm0["kF_Mpc"] * 1.05, "b_H": -0.03}):
        lya = th.lya_model.get_lya_params(cosmo, params)
        ref = orc.get_chi2(dict(lya, **{k: v for k, v in params.items() if k == "b_H"}))
        assert_chi2(like.get_chi2(params), ref, "IGM scalar %s" % params, n_data=like.get_ndata())
    rng = np.random.default_rng(3)
    names = ["mF", "gamma", "b_H"]
    pts = np.stack([igm0["mF"] * rng.uniform(0.95, 1.05, 40), igm0["gamma"] + rng.uniform(-0.1, 0.1, 40), rng.uniform(-0.03, -0.01, 40)], axis=1)
    batch = like.get_chi2_batch(pts, names)
    one = np.array([like.get_chi2(dict(zip(names, p))) for p in pts[:6]])
    np.testing.assert_allclose(batch[:6], one, rtol=1e-9)
    with pytest.raises(KeyError):
        like.get_chi2_batch(pts[:, :2], ["mF", "bias"])

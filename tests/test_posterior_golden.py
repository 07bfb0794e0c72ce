"""The posterior layer against the REFERENCE's own Posterior / FreeParameter classes
(cupix/likelihood/posterior.py:39-84, free_parameter.py:34-50), run unmodified on the reference Likelihood of the 'tiny'
fixture by tests/golden/make_golden_posterior.py.

  CPU   cupix_b200's Posterior on the float64 oracle likelihood: log posterior, log prior, bounds
  GPU   Posterior.get_log_posterior_batch on the device likelihood (one call for all points); the ensemble sampler
        (stands where the reference wraps emcee, sampler.py:39-43) on the device posterior against the batched minimiser
"""
import numpy as np
import pytest

from cupix_b200.likelihood.free_parameter import FreeParameter
from cupix_b200.likelihood.likelihood import Likelihood
from cupix_b200.likelihood.posterior import Posterior
from cupix_b200.px_data.data_DESI_DR2 import DESI_DR2
from tests import helpers as H

G = H.golden("ref_posterior_tiny.npz")
L = H.golden("ref_likelihood_tiny.npz")


def _free_params():
    out = []
    for i, name in enumerate(G["names"]):
        mean, width = G["prior_mean"][i], G["prior_width"][i]
        out.append(FreeParameter(str(name), min_value=float(G["lo"][i]), max_value=float(G["hi"][i]),
                                 ini_value=float(0.5 * (G["lo"][i] + G["hi"][i])),
                                 gauss_prior_mean=None if np.isnan(mean) else float(mean),
                                 gauss_prior_width=None if np.isnan(width) else float(width)))
    return out


def _like():
    data = DESI_DR2({k: L[k] for k in ("z_centers", "k_m", "k_M_edges", "theta_min_a", "theta_max_a", "theta_min_A",
                                      "theta_max_A", "N_fft", "L_fft", "B_A_a", "U_ZaMn", "V_ZaM", "Px_ZAM", "cov_ZAM")})
    th = H.make_theory(float(data.z[0]), "best_fit_arinyo_from_colore", {})
    return Likelihood(data, th, 0, config={"N_theta_average": int(L["N_theta_average"])})


def test_posterior_on_oracle_likelihood_equals_reference():
    orc = H.oracle_likelihood(_like())
    orc.get_ndata = lambda: orc.Px_AM.size
    post = Posterior(orc, _free_params())
    assert post.get_ndf() == int(G["ndf"])
    for v, lp, lpr, ll in zip(G["inside"], G["log_post_inside"], G["log_prior_inside"], G["log_like_inside"]):
        params = post.get_params_from_values(v)
        assert post.get_log_prior(params) == pytest.approx(lpr, rel=1e-13, abs=1e-13)
        assert orc.get_log_like(params) == pytest.approx(ll, rel=1e-10)
        assert post.get_log_posterior_from_values(v) == pytest.approx(lp, rel=1e-10)
    # outside the bounds the reference returns `--np.inf` = +inf (posterior.py:45), which sends a maximiser away for
    # good; here it is -inf -- the one deliberate difference, recorded by the fixture
    assert np.all(G["log_post_outside_reference"] == np.inf)
    for v in G["outside"]:
        assert post.get_log_posterior_from_values(v) == -np.inf


@pytest.mark.gpu
def test_batched_posterior_on_device_equals_reference():
    like = _like()
    post = Posterior(like, _free_params())
    vals = np.concatenate([G["inside"], G["outside"]])
    got = post.get_log_posterior_batch(vals)
    n_in = len(G["inside"])
    ref = G["log_post_inside"]
    # chi2 rule of the float32-cell pipeline: 1e-3 absolute where chi2 <= 2 n_data, 1e-6 relative beyond; log P = -chi2 / 2
    tol = np.where(-2 * G["log_like_inside"] <= 2 * like.get_ndata(), 5e-4, 5e-7 * np.abs(ref))
    assert np.all(np.abs(got[:n_in] - ref) <= tol), np.abs(got[:n_in] - ref).max()
    assert np.all(got[n_in:] == -np.inf)
    # scalar entry point, same numbers
    assert post.get_log_posterior_from_values(G["inside"][3]) == pytest.approx(got[3], rel=1e-9)
    # a batch large enough for the factorised path?  No: q1 is a tuple parameter; bias and beta alone are
    post2 = Posterior(like, _free_params()[:2])
    many = np.random.default_rng(1).uniform(G["lo"][:2], G["hi"][:2], size=(64, 2))
    f0 = like.engine().info("factored_calls")
    lp = post2.get_log_posterior_batch(many)
    assert like.engine().info("factored_calls") == f0 + 1
    one = np.array([post2.get_log_posterior_from_values(v) for v in many[:5]])
    np.testing.assert_allclose(lp[:5], one, rtol=2e-6, atol=5e-4)


@pytest.mark.gpu
def test_sampler_on_device_posterior_agrees_with_minimiser():
    """Affine-invariant ensemble on the device posterior (two batched calls per step): posterior mean within 0.2 sigma of
    the batched minimiser's best fit, widths within 25 % of its HESSE errors (S1 shape, bias and beta free)."""
    from cupix_b200.likelihood.batch_minimizer import BatchMinimizer
    from cupix_b200.likelihood.sampler import Sampler
    from cupix_b200.px_data import synthetic
    d = synthetic.make_measurement_dict("S1", N_A=8, N_m=48)
    th = H.make_theory(2.2, "best_fit_arinyo_from_colore", {"include_hcd": True})
    like = Likelihood(DESI_DR2(d), th, 0, config={"N_theta_average": 1})
    truth = like.build_plan().defaults[:2]
    filled = synthetic.fill_from_model(d, like.get_convolved_px({})[None], add_noise=True)
    like = Likelihood(DESI_DR2(filled), th, 0, config={"N_theta_average": 1})
    fit = BatchMinimizer(like, ["bias", "beta"], steps=[2e-4, 5e-3]).minimize(truth[None] * np.array([[1.02, 0.97]]))
    best, err = fit["x"][0], np.sqrt(np.diag(fit["cov"][0]))
    free = [FreeParameter("bias", truth[0] - 0.03, truth[0] + 0.03, ini_value=best[0]),
            FreeParameter("beta", truth[1] - 0.5, truth[1] + 0.5, ini_value=best[1])]
    sam = Sampler(Posterior(like, free), {"nwalkers": 64, "max_nsteps": 500, "nburnin": 150, "seed": 4})
    f0 = like.engine().info("factored_calls")
    chain = sam.run_sampler(p0=best + err * np.random.default_rng(2).normal(size=(64, 2)))
    assert like.engine().info("factored_calls") >= f0 + 1000          # 32 walkers per call: the quadratic-form path
    assert 0.2 < sam.acceptance_fraction < 0.9
    assert np.all(np.abs(chain.mean(axis=0) - best) < 0.2 * err), (chain.mean(axis=0), best, err)
    assert np.all(np.abs(chain.std(axis=0) / err - 1.0) < 0.25), (chain.std(axis=0), err)

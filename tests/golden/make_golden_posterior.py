"""Golden vectors for the posterior layer, produced by RUNNING the unmodified reference classes

    cupix.likelihood.posterior.Posterior            (posterior.py:39-84: bounds, log-likelihood, Gaussian priors)
    cupix.likelihood.free_parameter.FreeParameter   (free_parameter.py:34-50)

on top of the reference ``Likelihood`` of the 'tiny' fixture (tests/golden/ref_likelihood_tiny.npz; same stand-ins for
the absent third-party packages as make_golden.py).  Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_posterior.py

Stored: log-posterior, log-prior and log-likelihood of the reference at seeded points inside the bounds, and what it
returns outside them -- ``--np.inf`` = +inf (posterior.py:45), a typo the product deliberately does not reproduce
(cupix_b200/likelihood/posterior.py returns -inf); the fixture records the reference's value so that the test states
the difference instead of hiding it.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as MG                                     # noqa: E402  (installs nothing until asked)

PARAMS = [  # name, min, max, prior mean, prior width
    ("bias", -0.16, -0.08, -0.117, 0.004),
    ("beta", 1.0, 2.2, None, None),
    ("q1", 0.05, 1.2, 0.45, 0.2),
]


def main():
    MG.install_stand_ins()
    from cupix.likelihood.free_parameter import FreeParameter as RefFreeParameter
    from cupix.likelihood.likelihood import Likelihood as RefLikelihood
    from cupix.likelihood.posterior import Posterior as RefPosterior
    from cupix.likelihood.theory import Theory as RefTheory
    from cupix.px_data.base_px_data import BaseDataPx as RefBaseDataPx

    g = dict(np.load(os.path.join(HERE, "ref_likelihood_tiny.npz")))
    cosmo = MG.b200_cosmology.Cosmology()
    th = RefTheory(float(g["z_centers"][0]), fid_cosmo=cosmo, config={"default_lya_model": "best_fit_arinyo_from_colore", "verbose": False})
    data = RefBaseDataPx(g["Px_ZAM"], g["cov_ZAM"], g["z_centers"], g["k_M_edges"], g["theta_min_A"], g["theta_max_A"],
                         g["N_fft"], g["L_fft"], k_m_AA=g["k_m"], theta_min_a_arcmin=g["theta_min_a"],
                         theta_max_a_arcmin=g["theta_max_a"], B_A_a=g["B_A_a"], U_ZaMn=g["U_ZaMn"], V_ZaM=g["V_ZaM"])
    like = RefLikelihood(data, th, 0, config={"N_theta_average": int(g["N_theta_average"])})
    free = [RefFreeParameter(n, min_value=lo, max_value=hi, ini_value=0.5 * (lo + hi), gauss_prior_mean=m, gauss_prior_width=w)
            for n, lo, hi, m, w in PARAMS]
    post = RefPosterior(like, free)
    rng = np.random.default_rng(42)
    lo = np.array([p[1] for p in PARAMS])
    hi = np.array([p[2] for p in PARAMS])
    inside = rng.uniform(lo + 0.05 * (hi - lo), hi - 0.05 * (hi - lo), size=(24, 3))
    inside[0] = [-0.115, 1.5, 0.439]
    outside = inside[:6].copy()
    outside[0, 0] = lo[0] - 1e-3
    outside[1, 1] = hi[1] + 0.1
    outside[2, 2] = lo[2] - 0.01
    outside[3, 0] = hi[0] + 0.02
    outside[4] = [lo[0] - 1.0, hi[1] + 1.0, hi[2] + 1.0]
    outside[5, 2] = hi[2] + 1e-9
    out = {"names": np.array([p[0] for p in PARAMS]), "lo": lo, "hi": hi,
           "prior_mean": np.array([np.nan if p[3] is None else p[3] for p in PARAMS]),
           "prior_width": np.array([np.nan if p[4] is None else p[4] for p in PARAMS]),
           "inside": inside, "outside": outside, "ndf": post.get_ndf()}
    out["log_post_inside"] = np.array([post.get_log_posterior_from_values(v) for v in inside])
    out["log_prior_inside"] = np.array([post.get_log_prior(post.get_params_from_values(v)) for v in inside])
    out["log_like_inside"] = np.array([like.get_log_like(params=post.get_params_from_values(v)) for v in inside])
    out["log_post_outside_reference"] = np.array([post.get_log_posterior_from_values(v) for v in outside])
    assert np.all(out["log_post_outside_reference"] == np.inf)          # the `--np.inf` of posterior.py:45
    np.testing.assert_allclose(out["log_post_inside"], out["log_like_inside"] + out["log_prior_inside"], rtol=1e-15)
    np.savez_compressed(os.path.join(HERE, "ref_posterior_tiny.npz"), **out)
    print("wrote ref_posterior_tiny.npz: log posterior %.3f .. %.3f, ndf %d" %
          (out["log_post_inside"].min(), out["log_post_inside"].max(), out["ndf"]))


if __name__ == "__main__":
    main()

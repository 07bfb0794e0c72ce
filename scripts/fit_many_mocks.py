#!/usr/bin/env python
"""Fit many noisy mocks concurrently with batched likelihood calls (BASELINE config 4).

The reference's scripts/fit_many_mocks.py:151-170 loops serially over mocks, rebuilding data and
likelihood and running MIGRAD for each.  Here all mocks share windows and covariance, their data
vectors live side by side on the device, and one lock-step minimiser advances every fit with a
single batched chi2 call per iteration.

    python scripts/fit_many_mocks.py --nmocks 1000 bias beta
"""
import argparse
import time

import numpy as np

from _common import dist_setup, load_measurement
from cupix_b200.likelihood.batch_minimizer import BatchMinimizer

ap = argparse.ArgumentParser()
ap.add_argument("pars", nargs="+")
ap.add_argument("--nmocks", type=int, default=1000)
ap.add_argument("--iz", type=int, default=0)
ap.add_argument("--out", default=None)
a = ap.parse_args()
rank, world, local = dist_setup()
data, theories, likes = load_measurement(None, "S2", {"default_lya_model": "best_fit_arinyo_from_colore"}, 100, local,
                                         add_noise=False)
like = likes[a.iz]
dflt = like.theory.lya_model.default_lya_params
truth = np.array([dflt[p] for p in a.pars])
like.generate_mocks(a.nmocks, seed=1234)            # model(defaults) + L n per mock (likelihood.py:42-49), kept on the device
steps = [0.002 * max(abs(t), 0.1) for t in truth]
fitter = BatchMinimizer(like, a.pars, steps)
t0 = time.time()
res = fitter.minimize(np.tile(truth * 1.03, (a.nmocks, 1)), data_idx=np.arange(a.nmocks))
dt = time.time() - t0
err = np.sqrt(np.einsum("nii->ni", res["cov"]))
print("%d mocks fitted in %.2f s (%d batched calls); converged %d" % (a.nmocks, dt, res["n_calls"], res["converged"].sum()))
for j, p in enumerate(a.pars):
    print("  %-8s truth %+.5f  mean fit %+.5f  scatter %.5f  mean error %.5f" % (p, truth[j], res["x"][:, j].mean(),
                                                                            res["x"][:, j].std(), err[:, j].mean()))
print("  mean chi2 / ndf = %.3f" % (res["chi2"].mean() / (like.get_ndata() - len(a.pars))))
if a.out:
    np.savez(a.out, **{k: v for k, v in res.items() if isinstance(v, np.ndarray)}, pars=a.pars, truth=truth)

"""GPU-box report: latency of the scalar reference-style calls (one point per call, as iminuit /
emcee drive the reference) and small-batch throughput, S1 shape (1 z, 20 theta, 155 k_par)."""
import os
import sys
import time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cupix_b200.likelihood.likelihood import Likelihood   # noqa: E402
from cupix_b200.px_data import synthetic                    # noqa: E402
from cupix_b200.px_data.data_DESI_DR2 import DESI_DR2       # noqa: E402
from tests import helpers as H                              # noqa: E402

d = synthetic.make_measurement_dict("S1")
th = H.make_theory(float(d["z_centers"][0]), "best_fit_arinyo_from_colore", {"include_hcd": True}, n_q=88)
like = Likelihood(DESI_DR2(d), th, 0, config={"N_theta_average": 100})
truth = {"bias": -0.115, "beta": 1.5}
filled = synthetic.fill_from_model(d, like.get_convolved_px(truth)[None], add_noise=True)
like = Likelihood(DESI_DR2(filled), th, 0, config={"N_theta_average": 100})
like.get_chi2(truth)
t0 = time.time()
n = 500
for i in range(n):
    like.get_chi2({"bias": -0.115 + 1e-5 * i, "beta": 1.5})
dt = (time.time() - t0) / n
print("scalar Likelihood.get_chi2: %.1f us per call (reference: ~1 s per call, notebooks/development/profiling.py:82)" % (dt * 1e6))
for B in (1, 16, 256, 4096, 16384, 65536):
    pts = np.stack([np.linspace(-0.13, -0.10, B), np.linspace(1.3, 1.7, B)], axis=1)
    like.get_chi2_batch(pts, ["bias", "beta"])
    reps = max(3, min(200, 20000 // B))
    t0 = time.time()
    for _ in range(reps):
        like.get_chi2_batch(pts, ["bias", "beta"])
    dt = (time.time() - t0) / reps
    print("get_chi2_batch B = %6d: %9.1f us per call, %10.0f evals/s (single z bin)" % (B, dt * 1e6, B / dt))

"""Small end-to-end case touching every kernel (all contaminants, theta rebinning, several theta
blocks, ragged sizes, grid mode, data_idx, precise mode) -- run under compute-sanitizer."""
import os
import sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cupix_b200.likelihood.likelihood import JointLikelihood, Likelihood   # noqa: E402
from cupix_b200.px_data import synthetic                                     # noqa: E402
from cupix_b200.px_data.data_DESI_DR2 import DESI_DR2                        # noqa: E402
from tests import helpers as H                                               # noqa: E402

flags = {"include_hcd": True, "include_metal": True, "include_sky": True, "include_continuum": True}
for precise in (False, True):
    d = synthetic.make_measurement_dict("S1", N_A=5, fine_per_coarse=6, N_m=45, rebin_k=3, Nz=2)
    rng = np.random.default_rng(1)
    d["Px_ZAM"] = 0.05 * rng.normal(size=d["Px_ZAM"].shape) + 0.1
    data = DESI_DR2(d)
    likes = [Likelihood(data, H.make_theory(float(z), "best_fit_arinyo_from_gadget", flags, n_q=50,
                                            extra={"precise_cells": precise}), iz, config={"N_theta_average": 2})
             for iz, z in enumerate(d["z_centers"])]
    joint = JointLikelihood(likes)
    pts = np.stack([rng.uniform(-0.2, -0.1, 37), rng.uniform(1.0, 2.0, 37)], axis=1)
    c, info = joint.get_chi2_batch(pts, ["bias", "beta_1"], return_info=True)
    m = joint.get_convolved_px_batch(pts[:3], ["bias", "beta"])
    g = likes[0].get_chi2_grid(["bias", "beta"], [-0.2, 1.0], [-0.1, 2.0], [7, 5], first=3, count=20)
    px = likes[1]._compute_Px_Zam({"q1": 0.3})
    likes[0].set_mock_data(np.stack([data.Px_ZAM[0], 1.1 * data.Px_ZAM[0]]))
    cd = likes[0].get_chi2_batch(pts[:4], ["bias", "beta"], data_idx=[0, 1, 1, 0])
    th = likes[0].theory.get_px_obs(np.array([1.0, 5.0, 20.0]), np.array([0.0, 0.1, 0.7]), params={"bias": -0.12})
    print("precise" if precise else "fast", float(c.sum()), float(m.sum()), float(g.sum()), float(px.sum()), float(cd.sum()), float(th.sum()))
print("sanitize_case: done")

"""What would storing the Px_Zam scratch in float32 (for the chi2-only path) cost in chi2?  float64 emulation of the
plan (tests/helpers.py), Px_Zam rounded to float32 before the window / whitening / chi2 step."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from cupix_b200.likelihood.likelihood import Likelihood  # noqa: E402
from cupix_b200.px_data import synthetic  # noqa: E402
from cupix_b200.px_data.data_DESI_DR2 import DESI_DR2  # noqa: E402
from tests import helpers as H  # noqa: E402


def chi2_from_px(p, px):
    chi2 = 0.0
    for A in range(p.n_A):
        idx = np.where(p.B_A_a[A] != 0)[0]
        vs = p.V_aM[idx].sum(axis=0)
        mod = sum((p.V_aM[a] / vs) * (p.U_aMn[a] @ px[a]) for a in idx)
        y = np.linalg.solve(np.linalg.cholesky(p.cov_AMM[A]), p.Px_AM[0, A] - mod)
        chi2 += y @ y
    return chi2


d = synthetic.make_measurement_dict("S1")
th = H.make_theory(float(d["z_centers"][0]), "best_fit_arinyo_from_gadget", {"include_hcd": True}, n_q=88)
like = Likelihood(DESI_DR2(d), th, 0, config={"N_theta_average": 2})
p = like.build_plan()
truth = {"bias": float(p.defaults[0]), "beta": float(p.defaults[1])}
rng = np.random.default_rng(0)
model = H.emulate_plan_chi2(p, truth)[0]
filled = synthetic.fill_from_model(d, model[None], rng=rng, add_noise=True)
p.Px_AM = np.asarray(filled["Px_ZAM"][0])[None]
p.cov_AMM = np.asarray(filled["cov_ZAM"][0])
rows = []
for i in range(40):
    s = 0.0 if i == 0 else rng.uniform(0.0, 1.0)
    par = {"bias": truth["bias"] * (1 + 0.15 * s * rng.uniform(-1, 1)), "beta": truth["beta"] * (1 + 0.15 * s * rng.uniform(-1, 1))}
    px = H.emulate_plan_px(p, par)
    c64 = chi2_from_px(p, px)
    c32 = chi2_from_px(p, px.astype(np.float32).astype(np.float64))
    rows.append((c64, c32 - c64))
rows = np.array(rows)
near = rows[:, 0] < 2 * p.n_A * p.n_M
print("n_data %d; chi2 range [%.1f, %.1f]" % (p.n_A * p.n_M, rows[:, 0].min(), rows[:, 0].max()))
print("chi2 <= 2 n_data: max |dchi2| %.2e (tolerance 1e-3)" % np.abs(rows[near, 1]).max())
print("beyond          : max |dchi2|/chi2 %.2e (tolerance 1e-6)" % (np.abs(rows[~near, 1]) / rows[~near, 0]).max())

"""GPU-box report: distribution of the chi2 / Px differences between the CUDA path and the float64
oracle on the S1 shape (written to stdout; a copy is kept under profiles/)."""
import os
import sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cupix_b200.likelihood.likelihood import Likelihood   # noqa: E402
from cupix_b200.px_data import synthetic                    # noqa: E402
from cupix_b200.px_data.data_DESI_DR2 import DESI_DR2       # noqa: E402
from tests import helpers as H                              # noqa: E402


def report(model, flags, N_t, label, npts=24, seed=1):
    d = synthetic.make_measurement_dict("S1")
    z = float(d["z_centers"][0])
    th = H.make_theory(z, model, flags)
    like = Likelihood(DESI_DR2(d), th, 0, config={"N_theta_average": N_t})
    truth = {"bias": -0.115, "beta": 1.5}
    filled = synthetic.fill_from_model(d, like.get_convolved_px(truth)[None], add_noise=True)
    like = Likelihood(DESI_DR2(filled), th, 0, config={"N_theta_average": N_t})
    orc = H.oracle_likelihood(like)
    rng = np.random.default_rng(seed)
    names = ["bias", "beta", "q1", "kp_Mpc"]
    dflt = dict(zip(["bias", "beta", "q1", "q2", "kv_Mpc", "av", "bv", "kp_Mpc"], like.build_plan().defaults[:8]))
    pts = np.stack([rng.uniform(-0.135, -0.095, npts), rng.uniform(1.3, 1.7, npts),
                    dflt["q1"] * rng.uniform(0.7, 1.3, npts), dflt["kp_Mpc"] * rng.uniform(0.8, 1.2, npts)], axis=1)
    pts[0] = [truth["bias"], truth["beta"], dflt["q1"], dflt["kp_Mpc"]]
    chi2 = like.get_chi2_batch(pts, names)
    model_gpu = like.get_convolved_px_batch(pts, names)
    ref = np.array([orc.get_chi2(dict(zip(names, p))) for p in pts])
    mref = np.array([orc.get_convolved_px(dict(zip(names, p))) for p in pts])
    dchi = chi2 - ref
    dpx = np.abs(model_gpu - mref).max(axis=(1, 2)) / np.abs(mref).max(axis=(1, 2))
    print("%-28s chi2 range [%.1f, %.1f]  max|dchi2| %.2e  max|dchi2|/chi2 %.2e  rms dchi2 %.2e   max|dPx|/max|Px| %.2e"
          % (label, ref.min(), ref.max(), np.abs(dchi).max(), np.abs(dchi / ref).max(), np.sqrt(np.mean(dchi ** 2)), dpx.max()))
    # the two float64-cell modes on the same measurement: the full pipeline with CPX_PRECISE_CELLS, and the factorised
    # path (amplitude-only columns: bias, beta of the same points -> one tuple)
    like_p = Likelihood(like.data, th, 0, config={"N_theta_average": N_t, "precise_cells": True})
    dp = like_p.get_chi2_batch(pts, names, factor=False) - ref
    amp = np.tile(pts[:, :2], (2, 1))                      # 48 points: above the 24-point threshold of the factorised path
    ref_a = np.array([orc.get_chi2({"bias": p[0], "beta": p[1]}) for p in pts])
    calls0 = like.engine().info("factored_calls")
    da = like.get_chi2_batch(amp, ["bias", "beta"])[:npts] - ref_a
    assert like.engine().info("factored_calls") == calls0 + 1
    print("%-28s   float64 cells, full pipeline: max|dchi2| %.2e (%.2e relative);  factorised path: max|dchi2| %.2e (%.2e relative, chi2 up to %.0f)"
          % ("", np.abs(dp).max(), np.abs(dp / ref).max(), np.abs(da).max(), np.abs(da / ref_a).max(), ref_a.max()))
    return dchi, ref


if __name__ == "__main__":
    report("best_fit_arinyo_from_colore", {}, 2, "colore lya N_t=2")
    report("best_fit_arinyo_from_gadget", {}, 2, "gadget lya N_t=2")
    report("best_fit_arinyo_from_gadget", {"include_hcd": True, "include_metal": True, "include_sky": True,
                                           "include_continuum": True}, 3, "gadget all contaminants")

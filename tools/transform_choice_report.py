"""CPU report: what the choice of P3D -> Px transform costs in chi2.

chi2(oracle A) - chi2(oracle B) on the S1 shape (N_theta_average = 2, HCD on), where
  oracle A = the reference's algorithm (FFTLog + P1D splice + splines, cupix/likelihood/lyaP3D.py:272-351,
             including its k_par = 0 -> 1e-4 substitution, :35-39),
  oracle B = the fixed-node rule the device evaluates (float64 emulation of the device plan, tests/helpers.py),
for the k_perp rules on offer (n_q = 52 / 56 / 64 / 88 log-uniform nodes, the refined 'bao' presets) and both treatments of
the k_par = 0 row (Theory config kpar0 = 'exact' | 'reference'), on a smooth linear power and on one with a 5 % BAO-like wiggle.  The data vector is oracle A at the truth plus noise,
so the numbers are what a user comparing against the reference's own output would see: at the best fit and at
delta chi2 = 2.30 / 6.17 / 11.8 along bias, and far from it.

    python tools/transform_choice_report.py > profiles/transform_choice_r02.txt        (runs in ~2 minutes on 8 cores)
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cupix_b200.cosmology import Cosmology                 # noqa: E402
from cupix_b200.likelihood.likelihood import Likelihood     # noqa: E402
from cupix_b200.px_data import synthetic                    # noqa: E402
from cupix_b200.px_data.data_DESI_DR2 import DESI_DR2       # noqa: E402
from oracle.hankel import FFTLogHankel                      # noqa: E402
from tests import helpers as H                              # noqa: E402


def build(cosmo, extra, d=None, n_q=88):
    d = d or synthetic.make_measurement_dict("S1")
    th = H.make_theory(float(d["z_centers"][0]), "best_fit_arinyo_from_gadget", {"include_hcd": True}, cosmo=cosmo, n_q=n_q, extra=extra)
    return d, Likelihood(DESI_DR2(d), th, 0, config={"N_theta_average": 2})


def main(bao_amp=0.0):
    cosmo = Cosmology({"bao_amp": bao_amp}) if bao_amp else Cosmology()
    d, like0 = build(cosmo, {})
    orcA = H.oracle_likelihood(like0)
    orcA.theory.hankel = FFTLogHankel()
    dflt = like0.build_plan().defaults
    truth = {"bias": float(dflt[0]), "beta": float(dflt[1])}
    filled = synthetic.fill_from_model(d, orcA.get_convolved_px(truth)[None], add_noise=True)
    orcA.Px_AM, orcA.cov_AMM = filled["Px_ZAM"][0], filled["cov_ZAM"][0]
    n_data = orcA.Px_AM.size

    def chi2A(p):
        return orcA.get_chi2(p)

    # best fit along bias (A), then offsets reaching the contour levels
    bs = truth["bias"] * (1 + np.linspace(-0.01, 0.01, 9))
    cs = np.array([chi2A(dict(truth, bias=b)) for b in bs])
    a2, a1, a0 = np.polyfit(bs, cs, 2)
    b_best = -a1 / (2 * a2)
    pts = [("best fit (bias profile)", dict(truth, bias=b_best))]
    for lvl in (2.30, 6.17, 11.8):
        pts.append(("delta chi2 = %.2f" % lvl, dict(truth, bias=b_best + np.sqrt(lvl / a2))))
    pts.append(("far: bias +8 %, beta -0.2", dict(bias=truth["bias"] * 1.08, beta=truth["beta"] - 0.2)))
    pts.append(("far: bias -15 %, beta +0.3, q1 x 1.5", dict(bias=truth["bias"] * 0.85, beta=truth["beta"] + 0.3, q1=float(dflt[2]) * 1.5)))
    refA = [chi2A(p) for _, p in pts]
    pxA = orcA.Px_Zam(truth)

    print("S1 shape, N_theta_average = 2, HCD on, n_data = %d, linear power %s" %
          (n_data, "with a %.0f %% BAO-like wiggle" % (100 * bao_amp) if bao_amp else "smooth (Eisenstein-Hu no-wiggle)"))
    print("chi2(oracle A) at the points: " + ", ".join("%s: %.2f" % (n, c) for (n, _), c in zip(pts, refA)))
    print()
    print("%-30s %-10s | %-22s | %s" % ("k_perp rule", "kpar0", "max|dPx_Zam|/max (all rows / k_par>0)",
                                         " | ".join("%-12s" % n[:12] for n, _ in pts)))
    rules = [("smooth n_q=52 (default)", {"kperp_rule": "smooth"}, 52), ("smooth n_q=56", {"kperp_rule": "smooth"}, 56),
             ("smooth n_q=64", {"kperp_rule": "smooth"}, 64),
             ("smooth n_q=88", {"kperp_rule": "smooth"}, 88), ("bao n_q=128 (20, 0.5)", {"kperp_rule": "bao"}, 128),
             ("bao_fine n_q=160 (15, 0.7)", {"kperp_rule": "bao_fine"}, 160)]
    for rname, extra, n_q in rules:
        for kpar0 in ("exact", "reference"):
            _, like = build(cosmo, dict(extra, kpar0=kpar0), d=filled, n_q=n_q)
            plan = like.build_plan()
            px = H.emulate_plan_px(plan, truth)
            scale = np.abs(pxA).max()
            e_all = np.abs(px - pxA).max() / scale
            e_pos = np.abs(px - pxA)[:, 1:].max() / scale
            dchi = [H.emulate_plan_chi2(plan, p)[1].sum() - c for (_, p), c in zip(pts, refA)]
            print("%-30s %-10s | %9.2e / %9.2e | %s" % (rname, kpar0, e_all, e_pos, " | ".join("%+12.2e" % x for x in dchi)))
    print()


if __name__ == "__main__":
    main(0.0)
    main(0.05)

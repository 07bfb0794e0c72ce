"""CPU: measured error of the k_perp node sets on offer (cupix_b200.quadrature.RULES) against the reference's own
P3D -> Px algorithm (oracle A: FFTLog + P1D splice + splines, cupix/likelihood/lyaP3D.py:272-351).

    python tools/quadrature_scan.py > profiles/quadrature_scan_r02.txt

max |dPx| / max |Px| over three Arinyo parameter sets (central and extreme rows of the ForestFlow training table, the
CoLoRe fit), HCD on, theta = 0.5' .. 58' at z = 2.2, k_par = 0.0077 .. 1.19 1/A -- once with the smooth
(Eisenstein-Hu no-wiggle) linear power, once with a 5 % BAO-like wiggle on it (Cosmology(bao_amp=0.05))."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cupix_b200.quadrature as Q                           # noqa: E402
from cupix_b200.cosmology import Cosmology                  # noqa: E402
from oracle.hankel import BruteHankel, FFTLogHankel, NodeHankel   # noqa: E402
from oracle.px_oracle import OracleTheory                   # noqa: E402
from tests.test_quadrature import CONT, LYA_SETS, XI        # noqa: E402

THETA = np.array([0.5, 0.7, 1.5, 3.0, 6.0, 12.0, 25.0, 40.0, 58.0])
K_AA = np.array([0.0077, 0.03, 0.1, 0.25, 0.5, 0.8, 1.19])
SETS = ("gadget_central", "z3_max", "colore")


def main():
    fft = FFTLogHankel()
    brute = BruteHankel(n=2 ** 20 + 1)
    refs = {}
    for amp in (0.0, 0.05):
        cosmo = Cosmology({"bao_amp": amp}) if amp else Cosmology()
        for name in SETS:
            th = OracleTheory(2.2, cosmo, LYA_SETS[name], CONT, XI, fft, include_hcd=True)
            refs[amp, name] = (cosmo, th.get_px_obs(THETA, K_AA, {}))
        b = OracleTheory(2.2, cosmo, LYA_SETS["gadget_central"], CONT, XI, brute, include_hcd=True).get_px_obs(THETA, K_AA, {})
        a = refs[amp, "gadget_central"][1]
        print("bao_amp = %.2f: oracle A (FFTLog) vs 2^20-point brute force, gadget_central: %.2e" % (amp, np.abs(a - b).max() / np.abs(b).max()))

    def err(rule, amp):
        w = 0.0
        for name in SETS:
            cosmo, ref = refs[amp, name]
            th = OracleTheory(2.2, cosmo, LYA_SETS[name], CONT, XI, NodeHankel(rule.k, lambda r: rule.weights(r)), include_hcd=True)
            w = max(w, np.abs(th.get_px_obs(THETA, K_AA, {}) - ref).max() / np.abs(ref).max())
        return w

    print("%-46s %12s %12s %s" % ("rule", "smooth P_L", "5 % wiggle", "listed"))
    cands = [("smooth", kw) for kw, _ in Q.RULES["smooth"]] + [("smooth", dict(n_q=48)), ("smooth", dict(n_q=96))]
    cands += [("bao", kw) for kw, _ in Q.RULES["bao"]] + [("bao", dict(n_q=112, refine=50.0, k_refine=0.3))]
    for kind, kw in cands:
        rule = Q.KperpRule(**kw)
        listed = [e for k2, e in Q.RULES[kind] if k2 == kw]
        print("%-46s %12.2e %12.2e %s" % ("%s %s" % (kind, kw), err(rule, 0.0), err(rule, 0.05),
                                            ("%.1e" % listed[0]) if listed else "(not offered)"))


if __name__ == "__main__":
    main()

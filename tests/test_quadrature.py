"""CPU: the fixed-node Hankel rule (what the device evaluates) against an independent brute-force
integral and against the FFTLog restatement of the reference's algorithm."""
import numpy as np
import pytest

from cupix_b200.cosmology import Cosmology
from cupix_b200.quadrature import KperpRule
from oracle.hankel import BruteHankel, FFTLogHankel, NodeHankel, fftlog
from oracle.px_oracle import OracleTheory

LYA_SETS = {
    # central / extreme rows of data/emulator/ff_training_info.csv at z ~ 2.25 and 3, and the CoLoRe fit
    "gadget_central": dict(bias=-0.118, beta=1.557, q1=0.242, q2=0.293, av=0.289, bv=1.657, kp_Mpc=12.59, kv_Mpc=0.0828),
    "gadget_max": dict(bias=-0.089, beta=1.843, q1=0.693, q2=0.497, av=0.674, bv=1.736, kp_Mpc=24.23, kv_Mpc=0.825),
    "z3_max": dict(bias=-0.187, beta=1.587, q1=0.931, q2=0.539, av=0.962, bv=1.956, kp_Mpc=25.83, kv_Mpc=1.193),
    "colore": dict(bias=-0.1121, beta=1.55, q1=0.439, q2=0.0, av=0.1058, bv=2.4e-4, kp_Mpc=0.593, kv_Mpc=9.54e-7),
}
CONT = dict(b_H=-0.02, beta_H=0.5, L_H_Mpc=5 / 0.67, b_X=-0.005, beta_X=0.5, b_noise_Mpc=0.0035, kC_Mpc=0.019, pC=0.63)
XI = (np.array([0.0, 1000.0]), np.array([1.0, 0.0]))
THETA = np.array([0.7, 3.0, 12.0, 40.0, 58.0])
K_AA = np.array([0.0, 0.0077, 0.1, 0.5, 1.19])


def test_fftlog_analytic_pair():
    """int_0^inf r exp(-r^2/2) J0(kr) k dr = k exp(-k^2/2)."""
    x = np.logspace(-5, 5, 4096)
    y, G = fftlog(x, x * np.exp(-x * x / 2), q=0, mu=0)
    sel = (y > 1e-2) & (y < 6)
    np.testing.assert_allclose(G[sel], (y * np.exp(-y * y / 2))[sel], rtol=0, atol=2e-8)


@pytest.fixture(scope="module")
def brute():
    return BruteHankel(n=2 ** 19 + 1)


@pytest.mark.parametrize("name", sorted(LYA_SETS))
def test_node_rule_converged(name, brute):
    """Quadrature error of the default rule (n_q = 88) is < 2e-7 of the Px scale -- 50x inside the
    1e-5 model tolerance -- for central and extreme Arinyo parameters, all contaminants on."""
    cosmo = Cosmology()
    rule = KperpRule()
    assert rule.n_q == 88
    res = {}
    for key, h in (("node", NodeHankel(rule.k, lambda r: rule.weights(r, nthreads=1))), ("brute", brute)):
        th = OracleTheory(2.2, cosmo, LYA_SETS[name], CONT, XI, h, include_hcd=True, include_metal=True)
        res[key] = th.get_px_obs(THETA, K_AA, {})
    err = np.abs(res["node"] - res["brute"]).max() / np.abs(res["brute"]).max()
    assert err < 2e-7, err


def test_reference_algorithm_agrees_with_node_rule():
    """Oracle A (FFTLog + P1D splice + splines, the reference's algorithm) vs oracle B (the device
    rule): the gap is the reference's own discretisation error, < 2e-7 of the Px scale for k_par > 0.
    (At k_par = 0 the reference substitutes k_par = 1e-4, lyaP3D.py:35-39, so that column differs.)"""
    cosmo = Cosmology()
    rule = KperpRule()
    a = OracleTheory(2.2, cosmo, LYA_SETS["gadget_central"], CONT, XI, FFTLogHankel())
    b = OracleTheory(2.2, cosmo, LYA_SETS["gadget_central"], CONT, XI,
                     NodeHankel(rule.k, lambda r: rule.weights(r, nthreads=1)))
    pa, pb = a.get_px_obs(THETA, K_AA, {}), b.get_px_obs(THETA, K_AA, {})
    scale = np.abs(pb).max()
    assert np.abs(pa - pb)[:, 1:].max() / scale < 2e-7
    assert np.abs(pa - pb)[:, 0].max() / scale < 1e-3


def test_folded_weights_equal_theta_average():
    """The theta sub-sampling average of likelihood.py:99-130 folded into the weights is the same
    linear map as averaging the per-theta predictions."""
    rule = KperpRule(48)
    th = np.array([1.0, 1.2, 1.4, 5.0, 5.5, 6.0])
    grp = np.array([0, 0, 0, 1, 1, 1])
    r = th / 0.6
    W = rule.weights(r, nthreads=1)
    Wf = rule.weights(r, grp, th, 2, nthreads=1)
    for gi in range(2):
        s = grp == gi
        np.testing.assert_allclose(Wf[gi], (W[s] * th[s, None]).sum(0) / th[s].sum(), rtol=1e-10, atol=1e-15)

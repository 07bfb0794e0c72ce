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


def _wiggly_p3d(cosmo, z, amp):
    """Arinyo P3D (theory.py:285-308) on a linear power with a BAO-like oscillation of relative
    amplitude ``amp`` (period 2 pi / 147 Mpc, Silk-damped): what a CAMB power looks like to the k_perp rule."""
    par = dict(bias=-0.1483, beta=1.49, q1=0.297, q2=0.302, kv=0.527 ** (1 / 0.347), av=0.347, bv=1.68, kp=13.4)

    def p3d(k, mu):
        linP = cosmo.get_linP_Mpc(z, k) * (1 + amp * np.sin(k * 147.0) * np.exp(-(k / 0.25) ** 1.4))
        d2 = k ** 3 * linP / (2 * np.pi ** 2)
        vel = (k / par["kv"]) ** par["av"] * np.abs(mu) ** par["bv"]
        return par["bias"] ** 2 * (1 + par["beta"] * mu ** 2) ** 2 * linP * np.exp(
            d2 * (par["q1"] + par["q2"] * d2) * (1 - vel) - (k / par["kp"]) ** 2)
    return p3d


def test_bao_preset_resolves_wiggles_in_the_linear_power(brute):
    """Log-uniform nodes cannot follow baryon acoustic oscillations above k ~ 0.1 1/Mpc (88 nodes: 7e-4 of
    the Px scale); the 'bao' preset (128 nodes, gently refined below ~0.5 1/Mpc) stays well inside the 1e-5
    tolerance with a 5 % wiggle and on a smooth power, down to the smallest separations."""
    from cupix_b200.quadrature import PRESETS
    cosmo = Cosmology()
    r = np.geomspace(0.6, 80.0, 8)
    kp = np.array([0.0, 0.03, 0.1, 0.5, 2.0])
    bao = KperpRule(**PRESETS["bao"])
    assert bao.n_q == 128 and bao.refine == 20.0 and bao.k[0] == 0.0 and bao.k[-1] == 200.0 and np.all(np.diff(bao.k) > 0)
    plain = KperpRule()
    err = {}
    for amp in (0.0, 0.05):
        p3d = _wiggly_p3d(cosmo, 2.4, amp)
        ref = brute(r, kp, p3d)
        for name, rule in (("bao", bao), ("plain", plain)):
            got = NodeHankel(rule.k, lambda rr, rule=rule: rule.weights(rr, nthreads=1))(r, kp, p3d)
            err[name, amp] = np.abs(got - ref).max() / np.abs(ref).max()
    assert err["bao", 0.05] < 3.4e-6 and err["bao", 0.0] < 1.5e-6, err
    assert err["plain", 0.0] < 4e-7 and err["plain", 0.05] > 1e-4, err    # the reason the preset exists


def test_rule_choice_follows_tolerance_and_linear_power():
    """Theory picks the node family from the fiducial linear power (wiggles -> 'bao') and the cheapest node set whose
    measured error against the reference's algorithm is <= px_rtol / 3; explicit config keys still win."""
    from cupix_b200 import quadrature as Q
    from cupix_b200.likelihood.theory import Theory
    smooth, wiggly = Cosmology(), Cosmology({"bao_amp": 0.05})
    assert Q.wiggle_amplitude(smooth, 2.2) < Q.WIGGLE_THRESHOLD < Q.wiggle_amplitude(Cosmology({"bao_amp": 0.005}), 2.2)
    cfg = {"default_lya_model": "best_fit_arinyo_from_colore"}
    th = Theory(2.2, fid_cosmo=smooth, config=cfg)
    assert th.kperp_family == "smooth" and th.rule.n_q == 52 and th.rule.refine == 0.0 and th.rule_error * 3 <= 1e-5
    th = Theory(2.2, fid_cosmo=wiggly, config=cfg)
    assert th.kperp_family == "bao" and th.rule.n_q == 128 and th.rule.refine == 20.0 and th.rule_error * 3 <= 1e-5
    assert abs(th.wiggle_amplitude - 0.05) < 0.01
    assert Theory(2.2, fid_cosmo=smooth, config=dict(cfg, px_rtol=1e-6)).rule.n_q == 72
    assert Theory(2.2, fid_cosmo=wiggly, config=dict(cfg, px_rtol=2e-6)).rule.n_q == 160
    with pytest.warns(UserWarning):
        assert Theory(2.2, fid_cosmo=wiggly, config=dict(cfg, px_rtol=1e-7)).rule.n_q == 176
    th = Theory(2.2, fid_cosmo=wiggly, config=dict(cfg, kperp_rule="smooth", n_q=96))     # explicit keys win
    assert th.rule.n_q == 96 and th.rule.refine == 0.0
    assert Theory(2.2, fid_cosmo=smooth, config=dict(cfg, kperp_rule="bao_fine")).rule.n_q == 160
    for kind in Q.RULES:
        errs = [e for _, e in Q.RULES[kind]]
        assert all(kw["n_q"] % 4 == 0 for kw, _ in Q.RULES[kind]) and errs == sorted(errs, reverse=True)


def test_refined_map_inverse_and_default_unchanged():
    rule = KperpRule(64, refine=30.0, k_refine=0.2)
    np.testing.assert_allclose(rule.u_of_k(rule.k[1:-1]), rule.u[1:-1], rtol=0, atol=1e-12)
    # refine = 0 is the rule the golden fixtures were made with: nodes exactly log-uniform in k + k0
    plain = KperpRule(96)
    np.testing.assert_array_equal(plain.k[1:-1], (np.exp(np.linspace(np.log(0.01), np.log(200.01), 96)) - 0.01)[1:-1])


def test_weights_are_cached_by_separation():
    """Plans rebuilt for another primordial power (same background) reuse the Hankel weights."""
    from cupix_b200.quadrature import KperpRule
    rule = KperpRule(64)
    r = np.array([0.7, 1.3, 2.9, 3.1])
    grp, wgt = np.array([0, 0, 1, 1]), np.array([1.0, 2.0, 1.0, 1.0])
    w1 = rule.weights(r, grp, wgt, 2)
    w2 = rule.weights(r.copy(), grp.copy(), wgt.copy(), 2)
    assert np.array_equal(w1, w2) and w1 is not w2 and len(rule._weights_cache) == 1
    w1[:] = 0.0                                              # callers own their copy
    assert np.array_equal(rule.weights(r, grp, wgt, 2), w2)
    assert not np.array_equal(rule.weights(r * 1.01, grp, wgt, 2), w2) and len(rule._weights_cache) == 2
    assert rule.weights(r).shape == (4, 64) and len(rule._weights_cache) == 3


def test_jbar_hook_and_shared_rules():
    """The J0 panel sums can be supplied from outside (the product path evaluates them on the GPU, cpx_hankel_jbar): a
    NumPy stand-in with the same contract gives the weights of the built-in path; rules are shared per parameter set."""
    from scipy.special import j0
    from cupix_b200 import quadrature as Q

    def jbar_np(start, r, w, kf, meas):
        out = np.empty((len(start) - 1, kf.size))
        for g in range(len(start) - 1):
            sl = slice(start[g], start[g + 1])
            out[g] = (w[sl, None] * j0(kf[None, :] * r[sl, None])).sum(axis=0) * meas / w[sl].sum()
        return out
    rule = KperpRule(56)
    th = np.array([5.0, 1.0, 1.2, 5.5, 1.4, 6.0])              # groups not contiguous: the hook receives them sorted
    grp = np.array([1, 0, 0, 1, 0, 1])
    r = th / 0.6
    a = rule.weights(r, grp, th, 2, nthreads=1)
    b = rule.weights(r, grp, th, 2, jbar_fn=jbar_np)
    np.testing.assert_allclose(b, a, rtol=0, atol=1e-13 * np.abs(a).max())
    assert len(rule._weights_cache) == 2                       # the two paths never answer for each other
    assert Q.get_rule(56) is Q.get_rule(56) and Q.get_rule(56) is not Q.get_rule(64)
    assert Q.get_rule(128, refine=20.0, k_refine=0.5).refine == 20.0

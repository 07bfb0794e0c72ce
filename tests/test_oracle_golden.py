"""CPU: the oracle restatement and the host-side plan against the golden vectors produced by the
UNMODIFIED reference modules (tests/golden/make_golden.py)."""
import numpy as np
import pytest

from cupix_b200 import plan as _plan
from cupix_b200.px_data.base_px_data import BaseDataPx
from cupix_b200.px_data.data_DESI_DR2 import DESI_DR2
from cupix_b200.px_data import synthetic
from cupix_b200.likelihood.likelihood import Likelihood
from tests import helpers as H
from tests.golden.ref_theory_param_sets import FLAG_SETS, PARAM_SETS

G = H.golden("ref_theory_px.npz")
CASES = [k.split("|")[1:] for k in G if k.startswith("px|")]


def _key(model, z, fname, pname):
    return "px|%s|%s|%s|%s" % (model, z, fname, pname)


@pytest.mark.parametrize("model,z", sorted({(c[0], c[1]) for c in CASES}))
def test_defaults_match_reference(model, z):
    """LyaModel defaults (CoLoRe FITS headers, Gadget CSV, pressure_only variant) == reference's."""
    th = H.make_theory(float(z), model, {})
    d = _plan.default_vector(th, th.fid_cosmo)[:8]
    np.testing.assert_allclose(d, G["lyadef|%s|%s" % (model, z)], rtol=1e-14)


@pytest.mark.parametrize("model,z,fname", sorted({(c[0], c[1], c[2]) for c in CASES}))
def test_oracle_theory_matches_reference(model, z, fname):
    """oracle.px_oracle.OracleTheory == reference Theory.get_px_obs (same Hankel stand-in)."""
    th = H.make_theory(float(z), model, FLAG_SETS[fname])
    orc = H.oracle_for(th)
    for pname, params in PARAM_SETS.items():
        ref = G[_key(model, z, fname, pname)]
        got = orc.get_px_obs(G["theta_arc"], G["k_AA"], params)
        np.testing.assert_allclose(got, ref, rtol=1e-11, atol=1e-13 * np.abs(ref).max(), err_msg=pname)


@pytest.mark.parametrize("model,z,fname", sorted({(c[0], c[1], c[2]) for c in CASES}))
def test_plan_emulation_matches_reference(model, z, fname):
    """The host plan (folded weights, metal moments, sky tables) evaluated in float64 reproduces the
    reference Px -- i.e. the device algorithm is the reference's, up to arithmetic precision."""
    th = H.make_theory(float(z), model, FLAG_SETS[fname])
    theta = G["theta_arc"]
    p = _plan.build_theory_plan(th, th.fid_cosmo, theta, np.arange(theta.size), G["k_AA"], th.rule)
    for pname, params in PARAM_SETS.items():
        ref = G[_key(model, z, fname, pname)]
        got = H.emulate_plan_px(p, params)
        np.testing.assert_allclose(got, ref, rtol=1e-9, atol=1e-12 * np.abs(ref).max(), err_msg=pname)


@pytest.mark.parametrize("shape", ["tiny", "tiny_rebin"])
def test_oracle_likelihood_matches_reference(shape):
    """theta average, window, rebin, chi2 (likelihood.py) against the reference Likelihood."""
    g = H.golden("ref_likelihood_%s.npz" % shape)
    data = DESI_DR2({k: g[k] for k in ("z_centers", "k_m", "k_M_edges", "theta_min_a", "theta_max_a", "theta_min_A",
                                      "theta_max_A", "N_fft", "L_fft", "B_A_a", "U_ZaMn", "V_ZaM", "Px_ZAM", "cov_ZAM")})
    flags = {str(f): True for f in g["flags"]}
    trial = dict(zip([str(k) for k in g["trial_keys"]], g["trial_vals"]))
    truth = dict(zip([str(k) for k in g["truth_keys"]], g["truth_vals"]))
    N_t = int(g["N_theta_average"])
    for iz in range(len(data.z)):
        th = H.make_theory(float(data.z[iz]), "best_fit_arinyo_from_colore", flags)
        like = Likelihood(data, th, iz, config={"N_theta_average": N_t})
        orc = H.oracle_likelihood(like)
        np.testing.assert_allclose(orc.Px_Zam(trial), g["Px_Zam_trial_%d" % iz], rtol=1e-10,
                                   atol=1e-13 * np.abs(g["Px_Zam_trial_%d" % iz]).max())
        np.testing.assert_allclose(orc.get_convolved_px(trial), g["model_trial_%d" % iz], rtol=1e-10)
        chi2, info = orc.get_chi2(trial, return_info=True)
        np.testing.assert_allclose(chi2, g["chi2_trial_%d" % iz], rtol=1e-9)
        np.testing.assert_allclose(info["log_like_all"], g["log_like_all_trial_%d" % iz], rtol=1e-9)
        np.testing.assert_allclose(orc.get_chi2(truth), g["chi2_truth_%d" % iz], rtol=1e-8)
        # host plan + whitened formulation reproduce model and chi2 too
        p = like.build_plan()
        model, chi2_A = H.emulate_plan_chi2(p, trial)
        np.testing.assert_allclose(model, g["model_trial_%d" % iz], rtol=1e-9)
        np.testing.assert_allclose(chi2_A.sum(), g["chi2_trial_%d" % iz], rtol=1e-8)


def test_data_cuts_match_reference():
    """BaseDataPx range cuts (base_px_data.py:98-205)."""
    g = H.golden("ref_data_cuts.npz")
    d = synthetic.make_measurement_dict("tiny_rebin")
    data = DESI_DR2(d, theta_min_cut_arcmin=1.0, kM_max_cut_AA=0.2, km_max_cut_AA=0.22)
    np.testing.assert_allclose(np.asarray(data.k_m), g["k_m"])
    np.testing.assert_allclose(np.asarray(data.k_M_edges), g["k_M_edges"])
    assert tuple(data.U_ZaMn.shape) == tuple(g["U_shape"])
    assert tuple(data.Px_ZAM.shape) == tuple(g["Px_shape"])
    assert tuple(data.cov_ZAM.shape) == tuple(g["cov_shape"])
    np.testing.assert_allclose(data.U_ZaMn.sum(), g["U_sum"], rtol=1e-14)
    np.testing.assert_allclose(data.V_ZaM.sum(), g["V_sum"], rtol=1e-14)
    np.testing.assert_array_equal(data.B_A_a, g["B_A_a"])
    np.testing.assert_allclose(data.theta_min_A_arcmin, g["theta_min_A"])
    np.testing.assert_allclose(data.theta_min_a_arcmin, g["theta_min_a"])


@pytest.mark.parametrize("fname,cosmo_cols,tol", [
    ("hcd", {"As": 2.4e-9, "ns": 0.93}, 1e-9), ("hcd", {"ns": 1.0}, 1e-9), ("lya", {"As": 1.7e-9}, 1e-9),
    ("all", {"As": 2.4e-9}, 1e-9),             # with SiIII the amplitude factors out of the metal tables ...
    # ... and the tilt is carried by their Taylor tables in dn = ns - ns0 (plan.METAL_TILT_ORDER = 8): exact to the
    # truncation of exp(dn ln k/k_pivot), 1e-9 of the Px scale for everyday tilts, 2e-7 at dn = -0.17
    ("all", {"As": 2.4e-9, "ns": 0.93}, 1e-9), ("metal", {"ns": 1.02}, 1e-9), ("all", {"ns": 0.80}, 2e-7),
])
def test_plan_emulation_As_ns_columns_equal_rescaled_cosmology(fname, cosmo_cols, tol):
    """The per-point linear-power rescaling behind the As / ns batch columns (CPX_AS, CPX_NS) is what the
    reference gets by building a RescaledCosmology for the point (theory.py:54-56): same Px to 1e-9."""
    from cupix_b200.cosmology import RescaledCosmology
    th = H.make_theory(2.4, "best_fit_arinyo_from_gadget", FLAG_SETS[fname])
    theta = G["theta_arc"]
    p = _plan.build_theory_plan(th, th.fid_cosmo, theta, np.arange(theta.size), G["k_AA"], th.rule)
    assert p.k_pivot_Mpc == th.fid_cosmo.k_pivot and p.defaults[_plan.SLOT_INDEX["ns"]] == th.fid_cosmo.ns
    orc = H.oracle_for(th)
    orc.cosmo = RescaledCosmology(th.fid_cosmo, cosmo_cols)
    for pname in ("defaults", "arinyo", "contaminants"):
        params = PARAM_SETS[pname]
        ref = orc.get_px_obs(theta, G["k_AA"], params)
        got = H.emulate_plan_px(p, dict(params, **cosmo_cols))
        np.testing.assert_allclose(got, ref, rtol=tol, atol=max(tol * 1e-3, 1e-12) * np.abs(ref).max() if tol == 1e-9
                                   else tol * np.abs(ref).max(), err_msg=pname)
        assert np.abs(H.emulate_plan_px(p, params) - ref).max() > 1e-4 * np.abs(ref).max()   # the columns did something

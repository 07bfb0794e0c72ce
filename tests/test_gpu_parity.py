"""GPU parity: the CUDA path (through the C ABI, via the Theory / Likelihood mirrors) against the
golden vectors of the reference and against the float64 oracle.

Tolerances (BASELINE.json north_star): model Px relative 1e-5, chi2 absolute 1e-3.  "Relative" is
taken with respect to the largest |Px| of the compared array: Px crosses zero and falls by 18
decades across (theta, k_par), so an element-wise ratio is meaningless there.  The chi2 tolerance
is absolute 1e-3 for every point within chi2 <= 2 n_data (this covers the best fit and every
contour anyone draws: delta chi2 of 2.30 / 6.17 / 11.8) and relative 1e-6 beyond: far from the best
fit the whitened residuals are large and coherent, and chi2 inherits the ~1e-8 relative float32
noise floor of the P3D cell evaluation times 2 r^T L^-1 m.  Measured
(profiles/parity_report_r01.txt): |dchi2| <= 3e-3 at chi2 = 2e4, <= 1e-5 at the best fit,
max |dPx| / max|Px| <= 4e-8.
"""
import numpy as np
import pytest

from cupix_b200.engine import slots_from_names
from cupix_b200.likelihood.likelihood import JointLikelihood, Likelihood
from cupix_b200.px_data import synthetic
from cupix_b200.px_data.data_DESI_DR2 import DESI_DR2
from tests import helpers as H
from tests.golden.ref_theory_param_sets import FLAG_SETS, PARAM_SETS

pytestmark = pytest.mark.gpu

PX_RTOL = 1e-5
CHI2_ATOL = 1e-3
CHI2_RTOL_FAR = 1e-6

G = H.golden("ref_theory_px.npz")
CASES = sorted({tuple(k.split("|")[1:4]) for k in G if k.startswith("px|")})


def assert_px(got, ref, tol=PX_RTOL, msg=""):
    scale = np.abs(ref).max()
    err = np.abs(got - ref).max() / scale
    assert err <= tol, "%s: max |dPx| / max|Px| = %.3e > %.1e" % (msg, err, tol)
    return err


def assert_chi2(got, ref, msg="", n_data=None):
    """|dchi2| <= 1e-3 where chi2 <= 2 n_data (everywhere if n_data is not given), 1e-6 relative beyond."""
    got, ref = np.atleast_1d(np.asarray(got, dtype=float)), np.atleast_1d(np.asarray(ref, dtype=float))
    tol = np.full(ref.shape, CHI2_ATOL)
    if n_data is not None:
        far = ref > 2.0 * n_data
        tol[far] = np.maximum(CHI2_ATOL, CHI2_RTOL_FAR * np.abs(ref[far]))
    bad = np.abs(got - ref) > tol
    assert not bad.any(), "%s: max |dchi2| = %.3e (chi2 ~ %.3e)" % (msg, np.abs(got - ref).max(), np.abs(ref).max())


@pytest.mark.parametrize("model,z,fname", CASES)
def test_theory_px_vs_reference_golden(model, z, fname):
    th = H.make_theory(float(z), model, FLAG_SETS[fname])
    for pname, params in PARAM_SETS.items():
        ref = G["px|%s|%s|%s|%s" % (model, z, fname, pname)]
        got = th.get_px_obs(G["theta_arc"], G["k_AA"], params=params)
        assert got.shape == ref.shape
        assert_px(got, ref, msg="%s %s %s %s" % (model, z, fname, pname))


def test_theory_batch_equals_scalar():
    th = H.make_theory(2.4, "best_fit_arinyo_from_gadget", FLAG_SETS["all"])
    rng = np.random.default_rng(3)
    names = ["bias", "beta", "q1", "kp_Mpc", "b_H", "pC"]
    pts = np.stack([rng.uniform(-0.2, -0.1, 37), rng.uniform(1.0, 2.0, 37), rng.uniform(0.1, 0.9, 37),
                    rng.uniform(8, 25, 37), rng.uniform(-0.04, -0.01, 37), rng.uniform(0.4, 0.9, 37)], axis=1)
    batch = th.get_px_obs_batch(G["theta_arc"], G["k_AA"], pts, names)
    for i in (0, 17, 36):
        one = th.get_px_obs(G["theta_arc"], G["k_AA"], params=dict(zip(names, pts[i])))
        np.testing.assert_array_equal(batch[i], one)


def _load_like_fixture(shape):
    g = H.golden("ref_likelihood_%s.npz" % shape)
    data = DESI_DR2({k: g[k] for k in ("z_centers", "k_m", "k_M_edges", "theta_min_a", "theta_max_a", "theta_min_A",
                                      "theta_max_A", "N_fft", "L_fft", "B_A_a", "U_ZaMn", "V_ZaM", "Px_ZAM", "cov_ZAM")})
    flags = {str(f): True for f in g["flags"]}
    trial = dict(zip([str(k) for k in g["trial_keys"]], [float(v) for v in g["trial_vals"]]))
    truth = dict(zip([str(k) for k in g["truth_keys"]], [float(v) for v in g["truth_vals"]]))
    likes = [Likelihood(data, H.make_theory(float(data.z[iz]), "best_fit_arinyo_from_colore", flags), iz,
                        config={"N_theta_average": int(g["N_theta_average"])}) for iz in range(len(data.z))]
    return g, data, likes, trial, truth


@pytest.mark.parametrize("shape", ["tiny", "tiny_rebin"])
def test_likelihood_vs_reference_golden(shape):
    g, data, likes, trial, truth = _load_like_fixture(shape)
    for iz, like in enumerate(likes):
        assert_px(like._compute_Px_Zam(trial), g["Px_Zam_trial_%d" % iz], msg="Px_Zam z%d" % iz)
        assert_px(like.get_convolved_px(trial), g["model_trial_%d" % iz], msg="model z%d" % iz)
        chi2, info = like.get_chi2(trial, return_info=True)
        assert_chi2(chi2, g["chi2_trial_%d" % iz], "chi2 trial z%d" % iz)
        assert_chi2(-2 * info["log_like_all"], -2 * g["log_like_all_trial_%d" % iz], "log_like_all z%d" % iz)
        assert_chi2(like.get_chi2(truth), g["chi2_truth_%d" % iz], "chi2 truth z%d" % iz)
        np.testing.assert_allclose(like.get_log_like(trial), -0.5 * chi2, rtol=0, atol=0)
        assert like.get_ndata() == data.Px_ZAM[iz].size
    # all z bins in one device call == sum of the per-z likelihoods
    joint = JointLikelihood(likes)
    tot = sum(float(g["chi2_trial_%d" % iz]) for iz in range(len(likes)))
    assert_chi2(joint.get_chi2(trial), tot, "joint chi2")


def _s1_like(N_t=2, flags=None, shape="S1", model="best_fit_arinyo_from_colore", noise=True, **over):
    d = synthetic.make_measurement_dict(shape, **over)
    z = float(d["z_centers"][0])
    th = H.make_theory(z, model, flags or {})
    like = Likelihood(DESI_DR2(d), th, 0, config={"N_theta_average": N_t})
    truth = {"bias": -0.115, "beta": 1.5}
    filled = synthetic.fill_from_model(d, like.get_convolved_px(truth)[None], add_noise=noise)
    return Likelihood(DESI_DR2(filled), th, 0, config={"N_theta_average": N_t}), truth


def test_chi2_scan_16x16_vs_oracle():
    """16 x 16 (bias, beta) grid around truth on the S1 shape, |dchi2| <= 1e-3."""
    like, truth = _s1_like(N_t=2)
    lo, hi, n = [truth["bias"] - 0.02, truth["beta"] - 0.2], [truth["bias"] + 0.02, truth["beta"] + 0.2], [16, 16]
    # every point through the full pipeline (factor=False); the factorised path of such scans is tested in test_gpu_factor.py
    chi2 = like.get_chi2_grid(["bias", "beta"], lo, hi, n, factor=False).reshape(16, 16)
    # the same points, explicitly, through the points path
    bb, ee = np.meshgrid(np.linspace(lo[0], hi[0], 16), np.linspace(lo[1], hi[1], 16), indexing="ij")
    pts = np.stack([bb.ravel(), ee.ravel()], axis=1)
    np.testing.assert_array_equal(like.get_chi2_batch(pts, ["bias", "beta"], factor=False).reshape(16, 16), chi2)
    orc = H.oracle_likelihood(like)
    sel = [(0, 0), (3, 12), (8, 8), (15, 15), (15, 0), (7, 9), (11, 2)]
    for i, j in sel:
        ref = orc.get_chi2({"bias": bb[i, j], "beta": ee[i, j]})
        assert_chi2(chi2[i, j], ref, "grid point (%d,%d)" % (i, j), n_data=like.get_ndata())
    # a grid slice (how a rank takes its shard) equals the same slice of the full grid
    part = like.get_chi2_grid(["bias", "beta"], lo, hi, n, first=100, count=57, factor=False)
    np.testing.assert_array_equal(part, chi2.ravel()[100:157])


def test_forecast_self_consistency():
    """chi2 ~ 0 when the truth is scored against its own noiseless forecast
    (notebooks/examples/example_user_interface_notebook.py:337-359)."""
    like, truth = _s1_like(N_t=1, noise=False, flags=FLAG_SETS["all"])
    assert abs(like.get_chi2(truth)) < 1e-6
    assert like.get_chi2({"bias": -0.12, "beta": 1.5}) > 1.0


def test_rebinned_thetas_and_all_contaminants_vs_oracle():
    like, truth = _s1_like(N_t=3, flags=FLAG_SETS["all"], shape="S1_rebin", N_A=6, N_m=60)
    orc = H.oracle_likelihood(like)
    rng = np.random.default_rng(11)
    names = ["bias", "beta", "q1", "b_H", "b_X", "b_noise_Mpc", "kC_Mpc"]
    pts = np.stack([rng.uniform(-0.13, -0.10, 6), rng.uniform(1.3, 1.7, 6), rng.uniform(0.2, 0.7, 6),
                    rng.uniform(-0.03, -0.01, 6), rng.uniform(-0.01, -0.002, 6), rng.uniform(0.001, 0.005, 6),
                    rng.uniform(0.01, 0.03, 6)], axis=1)
    chi2, info = like.get_chi2_batch(pts, names, return_info=True)
    model = like.get_convolved_px_batch(pts, names)
    for i in range(pts.shape[0]):
        params = dict(zip(names, pts[i]))
        assert_px(model[i], orc.get_convolved_px(params), msg="model %d" % i)
        ref, rinfo = orc.get_chi2(params, return_info=True)
        assert_chi2(chi2[i], ref, "chi2 %d" % i, n_data=like.get_ndata())
        assert_chi2(-2 * info["log_like_all"][i], -2 * rinfo["log_like_all"], "per-theta chi2 %d" % i,
                    n_data=like.get_ndata() / len(rinfo["log_like_all"]))


def test_many_mocks_data_index():
    """S4 shape: one window/covariance, several data vectors, per-point data index."""
    like, truth = _s1_like(N_t=1, N_A=5, N_m=48)
    rng = np.random.default_rng(5)
    base = like.data.Px_ZAM[0]
    mocks = base[None] * (1 + 0.02 * rng.normal(size=(4,) + base.shape))
    like.set_mock_data(mocks)
    pts = np.tile([[truth["bias"], truth["beta"]]], (8, 1))
    didx = np.array([0, 1, 2, 3, 3, 2, 1, 0])
    chi2 = like.get_chi2_batch(pts, ["bias", "beta"], data_idx=didx)
    orc = H.oracle_likelihood(like)
    for j in range(4):
        orc.Px_AM = mocks[j]
        ref = orc.get_chi2(truth)
        assert_chi2(chi2[didx == j], np.full(2, ref), "mock %d" % j)


def test_unknown_keys_ignored_and_errors_are_loud():
    like, truth = _s1_like(N_t=1, N_A=4, N_m=32)
    a = like.get_chi2(dict(truth))
    b = like.get_chi2(dict(truth, not_a_parameter=1.0, kCb_Mpc=3.0))
    assert a == b
    with pytest.raises(KeyError):
        like.get_chi2_batch(np.zeros((2, 1)), ["not_a_parameter"])
    bad = synthetic.make_measurement_dict("tiny")
    bad["cov_ZAM"] = -bad["cov_ZAM"]
    from cupix_b200.engine import CpxError
    with pytest.raises(CpxError):
        Likelihood(DESI_DR2(bad), H.make_theory(2.2, "best_fit_arinyo_from_colore", {}), 0,
                   config={"N_theta_average": 1}).get_chi2({})


def test_more_than_64_coarse_k_bins():
    """n_M = 70 > 64: k_window walks the coarse-k columns in two blocks."""
    like, truth = _s1_like(N_t=1, N_A=3, N_m=140, rebin_k=2)
    assert like.data.Px_ZAM.shape[-1] == 70
    orc = H.oracle_likelihood(like)
    params = {"bias": -0.118, "beta": 1.57}
    assert_px(like.get_convolved_px(params), orc.get_convolved_px(params), msg="model n_M=70")
    ref, rinfo = orc.get_chi2(params, return_info=True)
    got, info = like.get_chi2(params, return_info=True)
    assert_chi2(got, ref, "chi2 n_M=70", n_data=like.get_ndata())
    assert_chi2(-2 * info["log_like_all"], -2 * rinfo["log_like_all"], "per-theta n_M=70", n_data=like.get_ndata() / 3)


def test_precise_cells_mode_meets_absolute_chi2_tolerance_everywhere():
    """CPX_PRECISE_CELLS: with float64 P3D cells the absolute 1e-3 chi2 tolerance of the north star
    holds for every point of the scan (here to 1e-6 at chi2 up to 2e4), and Px to 1e-10."""
    d = synthetic.make_measurement_dict("S1")
    z = float(d["z_centers"][0])
    flags = FLAG_SETS["all"]
    th = H.make_theory(z, "best_fit_arinyo_from_gadget", flags, extra={"precise_cells": True})
    like = Likelihood(DESI_DR2(d), th, 0, config={"N_theta_average": 2})
    truth = {"bias": -0.115, "beta": 1.5}
    filled = synthetic.fill_from_model(d, like.get_convolved_px(truth)[None], add_noise=True)
    like = Likelihood(DESI_DR2(filled), th, 0, config={"N_theta_average": 2})
    assert like.precise_cells
    orc = H.oracle_likelihood(like)
    rng = np.random.default_rng(21)
    names = ["bias", "beta", "q1", "kp_Mpc", "b_H"]
    dflt = like.build_plan().defaults
    pts = np.stack([rng.uniform(-0.135, -0.095, 8), rng.uniform(1.3, 1.7, 8), dflt[2] * rng.uniform(0.7, 1.3, 8),
                    dflt[7] * rng.uniform(0.8, 1.2, 8), rng.uniform(-0.03, -0.01, 8)], axis=1)
    chi2 = like.get_chi2_batch(pts, names)
    model = like.get_convolved_px_batch(pts, names)
    for i in range(len(pts)):
        params = dict(zip(names, pts[i]))
        assert_px(model[i], orc.get_convolved_px(params), tol=1e-10, msg="precise model %d" % i)
        ref = orc.get_chi2(params)
        assert abs(chi2[i] - ref) <= 1e-6, "precise chi2 %d: %.3e at chi2 %.1f" % (i, abs(chi2[i] - ref), ref)


@pytest.mark.parametrize("N_A,fpc,N_m,rebin_k,n_q", [(10, 1, 50, 5, 88), (5, 6, 45, 3, 90), (3, 9, 33, 3, 64), (1, 1, 8, 2, 48),
                                                     (24, 1, 64, 4, 88), (25, 1, 40, 2, 72)])
def test_ragged_shapes_vs_float64_emulation(N_A, fpc, N_m, rebin_k, n_q):
    """Shapes that exercise padding everywhere: theta bins not a multiple of 8 (and > 24: several
    theta blocks), k_par rows not a multiple of 32, coarse k not a multiple of 8, n_q not a multiple
    of 4.  Checked against the float64 NumPy emulation of the plan (itself pinned to the reference in
    tests/test_oracle_golden.py)."""
    d = synthetic.make_measurement_dict("S1", N_A=N_A, fine_per_coarse=fpc, N_m=N_m, rebin_k=rebin_k, Nz=2)
    flags = {"include_hcd": True, "include_metal": True, "include_sky": True, "include_continuum": True}
    likes = []
    for iz, z in enumerate(d["z_centers"]):
        th = H.make_theory(float(z), "best_fit_arinyo_from_gadget", flags, n_q=n_q)
        likes.append(Likelihood(DESI_DR2(d), th, iz, config={"N_theta_average": 2}))
    rng = np.random.default_rng(N_A * 100 + N_m)
    d["Px_ZAM"] = 0.05 * rng.normal(size=d["Px_ZAM"].shape) + 0.1
    data = DESI_DR2(d)
    for lk in likes:
        lk.data = data
    joint = JointLikelihood(likes)
    names = ["bias", "beta", "q1", "b_H", "b_X", "b_noise_Mpc", "pC"]
    pts = np.stack([rng.uniform(-0.2, -0.1, 5), rng.uniform(1.0, 2.0, 5), rng.uniform(0.1, 0.6, 5),
                    rng.uniform(-0.03, -0.01, 5), rng.uniform(-0.01, -0.002, 5), rng.uniform(0.001, 0.004, 5),
                    rng.uniform(0.4, 0.9, 5)], axis=1)
    chi2, info = joint.get_chi2_batch(pts, names, return_info=True)
    model = joint.get_convolved_px_batch(pts, names)
    plans = [lk.build_plan() for lk in likes]
    for i in range(len(pts)):
        params = dict(zip(names, pts[i]))
        tot = 0.0
        for iz, p in enumerate(plans):
            m_ref, c_ref = H.emulate_plan_chi2(p, params)
            assert_px(model[i, iz], m_ref, msg="ragged model pt %d z %d" % (i, iz))
            np.testing.assert_allclose(-2 * info["log_like_all"][i, iz], c_ref, rtol=2e-6)
            tot += c_ref.sum()
        np.testing.assert_allclose(chi2[i], tot, rtol=2e-6)
    # one point, one z bin of the joint engine, and the empty batch
    one = joint.engine().eval(points=pts[:1], slots=[0, 1, 2, 8, 11, 13, 15], z_list=[1], want=("chi2", "model"))
    np.testing.assert_allclose(one["chi2"][0], H.emulate_plan_chi2(plans[1], dict(zip(names, pts[0])))[1].sum(), rtol=2e-6)
    assert one["model"].shape == (1, 1, N_A, plans[1].n_M)
    empty = joint.get_chi2_batch(np.zeros((0, 2)), ["bias", "beta"])
    assert empty.shape == (0,)


def test_large_batch_spans_chunks():
    """70 000 points on a small shape: the device chunk is 65 536 points here, so this crosses a chunk
    boundary and exercises the largest launch grids; results must not depend on the batch layout."""
    like, truth = _s1_like(N_t=1, N_A=4, N_m=32, flags={"include_hcd": True})
    B = 70000
    pts = np.stack([np.linspace(-0.13, -0.10, B), np.linspace(1.3, 1.7, B)], axis=1)
    chi2 = like.get_chi2_batch(pts, ["bias", "beta"], factor=False)          # the full pipeline is the subject here
    sel = np.array([0, 1, 65535, 65536, 65537, B - 1])
    # the big batch takes the tcgen05 window kernel, the 6-point one the FP64 DMMA kernel: equal to the ~2^-40 of
    # the integer split; within the big batch the chunk boundary must not show at all
    np.testing.assert_allclose(chi2[sel], like.get_chi2_batch(pts[sel], ["bias", "beta"]), rtol=1e-9)
    np.testing.assert_array_equal(chi2[65530:65542], like.get_chi2_batch(pts[10000:], ["bias", "beta"], factor=False)[55530:55542])
    assert np.isfinite(chi2).all()
    assert like.engine().info("factored_calls") == 0


def test_bao_refined_rule_on_device():
    """A 'bao' k_perp node set (112 non-log-uniform nodes = 28 DMMA k-steps) through the CUDA path against the
    float64 emulation of the same plan and, for the model Px, against the oracle."""
    d = synthetic.make_measurement_dict("S1", N_A=7, N_m=40, Nz=1)
    th = H.make_theory(2.2, "best_fit_arinyo_from_gadget", {"include_hcd": True}, n_q=112, extra={"kperp_rule": "bao"})
    assert th.rule.n_q == 112 and th.rule.refine == 20.0
    like = Likelihood(DESI_DR2(d), th, 0, config={"N_theta_average": 2})
    names = ["bias", "beta", "q1", "kp_Mpc"]
    rng = np.random.default_rng(2)
    pts = np.stack([rng.uniform(-0.2, -0.1, 9), rng.uniform(1.0, 2.0, 9), rng.uniform(0.1, 0.6, 9), rng.uniform(8, 20, 9)], axis=1)
    model = like.get_convolved_px_batch(pts, names)
    p = like.build_plan()
    orc = H.oracle_likelihood(like)
    for i in (0, 4, 8):
        params = dict(zip(names, pts[i]))
        assert_px(model[i], H.emulate_plan_chi2(p, params)[0], msg="bao rule, emulation, pt %d" % i)
        assert_px(model[i], orc.get_convolved_px(params), msg="bao rule, oracle, pt %d" % i)


def _training_box(z):
    """min / max of the eight Arinyo slots at the ForestFlow training redshift nearest to z
    (cupix_b200/data/ff_training_info.csv, the file the reference ships as data/emulator/ff_training_info.csv)."""
    import os
    import cupix_b200
    ff = np.genfromtxt(os.path.join(os.path.dirname(cupix_b200.__file__), "data", "ff_training_info.csv"), delimiter=",", names=True)
    row = ff[np.argmin(np.abs(ff["z"] - z))]
    return {k: (float(row[k + "_min"]), float(row[k + "_max"])) for k in ("bias", "beta", "q1", "q2", "kvav", "av", "bv", "kp")}


def test_latin_hypercube_over_all_16_slots_vs_float64_emulation():
    """256 Latin-hypercube points over the 16-slot parameter box (Arinyo slots over the ForestFlow
    training range, contaminant slots around their defaults), all contaminants on: model Px within 1e-5 of the
    array maximum, chi2 within 1e-3 where chi2 <= 2 n_data and 1e-6 relative beyond."""
    like, _ = _s1_like(N_t=2, flags=FLAG_SETS["all"], model="best_fit_arinyo_from_gadget", N_A=8, N_m=64)
    box = _training_box(like.theory.z)
    n = 256
    rng = np.random.default_rng(256)

    def lhs(lo, hi):
        return lo + (hi - lo) * (rng.permutation(n) + rng.uniform(size=n)) / n
    av = lhs(max(box["av"][0], 0.05), box["av"][1])
    cols = {"bias": lhs(*box["bias"]), "beta": lhs(*box["beta"]), "q1": lhs(*box["q1"]), "q2": lhs(*box["q2"]),
            "kv_Mpc": np.exp(np.log(lhs(*box["kvav"])) / av), "av": av, "bv": lhs(*box["bv"]), "kp_Mpc": lhs(*box["kp"]),
            "b_H": lhs(-0.04, -0.005), "beta_H": lhs(0.2, 0.9), "L_H_Mpc": lhs(3.0, 12.0), "b_X": lhs(-0.01, -0.001),
            "beta_X": lhs(0.2, 0.9), "b_noise_Mpc": lhs(0.0005, 0.006), "kC_Mpc": lhs(0.008, 0.03), "pC": lhs(0.3, 0.9)}
    names = list(cols)
    pts = np.stack([cols[k] for k in names], axis=1)
    chi2 = like.get_chi2_batch(pts, names)
    model = like.get_convolved_px_batch(pts, names)
    p = like.build_plan()
    assert np.isfinite(chi2).all()
    for i in range(n):
        m_ref, c_ref = H.emulate_plan_chi2(p, dict(zip(names, pts[i])))
        assert_px(model[i], m_ref, msg="LHS point %d" % i)
        assert_chi2(chi2[i], c_ref.sum(), "LHS point %d" % i, n_data=like.get_ndata())


@pytest.mark.parametrize("params", [{"bv": 1e-13}, {"q1": 0.0, "q2": 0.0}, {"kp_Mpc": 1e3, "q1": 0.0, "q2": 0.0},
                                    {"av": 1.0, "bv": 2.2, "kv_Mpc": 1.0}, {"beta": 0.0}, {"b_H": 0.0}])
def test_edge_parameter_values_vs_float64_emulation(params):
    """SURVEY 8(c) edge cases: mu^bv -> 1 (CoLoRe fits have bv ~ 0), no non-linear term, kp -> large (linear theory,
    notebooks/analysis/mock_validation.py:82), the k_par = 0 row (mu = 0) is part of every S1 shape."""
    like, _ = _s1_like(N_t=1, flags={"include_hcd": True}, model="best_fit_arinyo_from_gadget", N_A=6, N_m=48)
    names = list(params)
    pts = np.array([[params[k] for k in names]])
    p = like.build_plan()
    assert p.kpar_Mpc[0] == 0.0
    m_ref, c_ref = H.emulate_plan_chi2(p, params)
    assert_px(like.get_convolved_px_batch(pts, names)[0], m_ref, msg=str(params))
    assert_chi2(like.get_chi2_batch(pts, names)[0], c_ref.sum(), str(params), n_data=like.get_ndata())


def test_cosmology_keys_in_params_vs_oracle():
    """theory.py:49-64: `As` / `ns` among the params select a RescaledCosmology (same background), a changed
    background a new cosmology; the device plan is rebuilt per cosmology and cached by value.  Interleaved
    calls (test_likelihood.py:134-140 scans `ns` this way) must never be answered from another cosmology's plan."""
    from cupix_b200.cosmology import Cosmology, RescaledCosmology
    like, truth = _s1_like(N_t=1, flags={"include_hcd": True})
    orc = H.oracle_likelihood(like)
    fid = like.theory.fid_cosmo
    seq = [{"ns": 0.95}, {"ns": 0.98, "As": 2.3e-9}, {}, {"ns": 0.95}, {"As": 1.9e-9}, {"ns": 0.98, "As": 2.3e-9},
           {"H0": 70.0, "ns": 0.95}]
    seen = {}
    for cp in seq:
        p = dict(truth, bias=-0.12, **cp)
        got = like.get_chi2(p)
        if not cp:
            orc.theory.cosmo = fid
        elif "H0" in cp:
            orc.theory.cosmo = Cosmology(dict(fid.params, **cp))
        else:
            orc.theory.cosmo = RescaledCosmology(fid, cp)
        assert_chi2(got, orc.get_chi2(p), "cosmology keys %s" % cp, n_data=like.get_ndata())
        key = tuple(sorted(cp.items()))
        if key in seen:
            assert got == seen[key]          # same cosmology again: same plan, same bits
        seen[key] = got
    assert len(set(seen.values())) == len(seen)      # every cosmology moved chi2
    assert len(like._engines) == len(seen)           # one plan per distinct cosmology, temporaries share by value


def test_As_ns_batch_columns_vs_oracle():
    """primordial amplitude and tilt as batch columns (CPX_AS, CPX_NS) -- the device rescales the
    linear power of the plan's cosmology per point; the oracle builds a RescaledCosmology per point
    (theory.py:54-56).  Points, grid and reference-precision paths; refusals."""
    from cupix_b200.cosmology import RescaledCosmology
    from cupix_b200.engine import CpxError
    like, truth = _s1_like(N_t=1, flags={"include_hcd": True, "include_sky": True, "include_continuum": True})
    orc = H.oracle_likelihood(like)
    fid = like.theory.fid_cosmo
    names = ["bias", "As", "ns", "q1"]
    rng = np.random.default_rng(5)
    pts = np.stack([rng.uniform(-0.125, -0.105, 40), rng.uniform(1.8e-9, 2.4e-9, 40), rng.uniform(0.93, 1.0, 40),
                    rng.uniform(0.1, 0.6, 40)], axis=1)
    pts[0, 1:3] = fid.As, fid.ns                     # fiducial values in the columns: the plain model
    chi2 = like.get_chi2_batch(pts, names)
    assert_chi2(chi2[0], like.get_chi2({"bias": pts[0, 0], "q1": pts[0, 3]}), "fiducial As, ns in the columns")
    fine = like.get_chi2_batch(pts, names)
    np.testing.assert_array_equal(chi2, fine)
    like_p = Likelihood(like.data, like.theory, 0, config={"N_theta_average": 1, "precise_cells": True})
    chi2_p = like_p.get_chi2_batch(pts[:8], names)
    px = like.get_convolved_px_batch(pts[:8], names)
    for i in range(8):
        orc.theory.cosmo = RescaledCosmology(fid, {"As": pts[i, 1], "ns": pts[i, 2]})
        par = {"bias": pts[i, 0], "q1": pts[i, 3]}
        ref = orc.get_chi2(par)
        assert_chi2(chi2[i], ref, "As/ns point %d" % i, n_data=like.get_ndata())
        np.testing.assert_allclose(chi2_p[i], ref, rtol=2e-8, err_msg="precise cells, point %d" % i)
        assert_px(px[i], orc.get_convolved_px(par), msg="model, point %d" % i)
    assert np.ptp(chi2) > 10.0                        # the columns move chi2
    # a scalar call with the same cosmology keys goes through a rebuilt plan (reference behaviour): same answer
    assert_chi2(like.get_chi2({"bias": pts[3, 0], "q1": pts[3, 3], "As": pts[3, 1], "ns": pts[3, 2]}), chi2[3],
                "scalar call with cosmology keys", n_data=like.get_ndata())
    # on-device grid over (As, ns)
    lo, hi, n = [1.9e-9, 0.94], [2.3e-9, 0.99], [5, 4]
    g = like.get_chi2_grid(["As", "ns"], lo, hi, n)
    aa, nn = np.meshgrid(np.linspace(lo[0], hi[0], 5), np.linspace(lo[1], hi[1], 4), indexing="ij")
    np.testing.assert_array_equal(g, like.get_chi2_batch(np.stack([aa.ravel(), nn.ravel()], axis=1), ["As", "ns"]))
    # SiIII: the amplitude factors out of the metal tables, the tilt does not
    like_m, _ = _s1_like(N_t=1, flags={"include_metal": True, "include_hcd": True})
    orc_m = H.oracle_likelihood(like_m)
    got = like_m.get_chi2_batch(np.array([[-0.12, 2.3e-9]]), ["bias", "As"])[0]
    orc_m.theory.cosmo = RescaledCosmology(fid, {"As": 2.3e-9})
    assert_chi2(got, orc_m.get_chi2({"bias": -0.12}), "As with SiIII", n_data=like_m.get_ndata())
    # ... it is carried by Taylor tables of the metal moments in dn = ns - ns0 (plan.METAL_TILT_ORDER): ns as a column
    # together with SiIII, on the full pipeline (float32 and float64 cells) and on the factorised path, against the oracle
    # with the rescaled cosmology (the reference: a new cosmology object per call, theory.py:54-56)
    like_mp = Likelihood(like_m.data, like_m.theory, 0, config={"N_theta_average": 1, "precise_cells": True})
    for ns_val, As_val in ((0.95, fid.As), (1.03, 2.0e-9), (fid.ns - 0.15, 2.3e-9)):
        orc_m.theory.cosmo = RescaledCosmology(fid, {"As": As_val, "ns": ns_val})
        par = {"bias": -0.12, "b_X": -0.008, "beta_X": 0.6}
        ref = orc_m.get_chi2(par)
        row = np.array([[par["bias"], par["b_X"], par["beta_X"], As_val, ns_val]])
        cols = ["bias", "b_X", "beta_X", "As", "ns"]
        assert_chi2(like_m.get_chi2_batch(row, cols)[0], ref, "ns = %.3f with SiIII" % ns_val, n_data=like_m.get_ndata())
        assert like_mp.get_chi2_batch(row, cols)[0] == pytest.approx(ref, rel=3e-7)          # series truncation: 1e-5 of the metal term
        assert_px(like_m.get_convolved_px_batch(row, cols)[0], orc_m.get_convolved_px(par), msg="model, ns = %.3f with SiIII" % ns_val)
        lo, hi, n = [ns_val, -0.125, -0.009], [ns_val + 0.02, -0.115, -0.006], [2, 8, 5]
        g = like_m.get_chi2_grid(["ns", "bias", "b_X"], lo, hi, n)                           # 2 tuples x 40 amplitude points
        orc_m.theory.cosmo = RescaledCosmology(fid, {"ns": hi[0]})
        assert g[-1] == pytest.approx(orc_m.get_chi2({"bias": hi[1], "b_X": hi[2]}), rel=3e-7)
    assert like_m.engine().info("factored_calls") == 3

    # a cosmology object that does not expose As / ns / k_pivot: the columns are refused, everything else works
    class Opaque(object):
        def __init__(self, c):
            self._c = c
        def get_linP_Mpc(self, z, k_Mpc): return self._c.get_linP_Mpc(z, k_Mpc)
        def get_darc_dMpc(self, z): return self._c.get_darc_dMpc(z)
        def get_dAA_dMpc(self, z, lambda_rest_AA=None): return self._c.get_dAA_dMpc(z, lambda_rest_AA=lambda_rest_AA)
        def same_background(self, cosmo_params=None): return True
        def get_linP_Mpc_params(self, z, kp_Mpc=0.7): return self._c.get_linP_Mpc_params(z, kp_Mpc)
    th_o = H.make_theory(like.theory.z, "best_fit_arinyo_from_colore", {}, cosmo=Opaque(fid))
    like_o = Likelihood(like.data, th_o, 0, config={"N_theta_average": 1})
    assert np.isfinite(like_o.get_chi2({"bias": -0.12}))
    with pytest.raises(CpxError, match="k_pivot"):
        like_o.get_chi2_batch(np.array([[2.2e-9]]), ["As"])


@pytest.mark.parametrize("B", [203, 49152 + 77])     # DMMA window with a partial tile; tcgen05 window with a partial tile
def test_device_outputs_stay_inside_their_buffers(B):
    """Guard bands around every device output of cpx_eval (chi2, chi2_zA, model_zAM, px_zam) on a shape with
    padding in every dimension (theta 11 x 2, k_par 50, coarse k 16, two z bins): the bands stay untouched and
    the interior is fully written.  (compute-sanitizer is not available on the GPU pool.)"""
    import torch
    d = synthetic.make_measurement_dict("S1", N_A=11, fine_per_coarse=2, N_m=50, rebin_k=3, Nz=2)
    likes = [Likelihood(DESI_DR2(d), H.make_theory(float(z), "best_fit_arinyo_from_gadget", {"include_hcd": True}, n_q=72),
                        iz, config={"N_theta_average": 1}) for iz, z in enumerate(d["z_centers"])]
    joint = JointLikelihood(likes)
    eng = joint.engine()
    n_z, n_A, n_M, n_a, n_m = 2, 11, 16, 22, 50
    G, SENT = 4096, -7.25e99
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(B)
    pts = torch.tensor(np.stack([rng.uniform(-0.2, -0.1, B), rng.uniform(1.0, 2.0, B)], axis=1), device=dev)
    slots, zsel = slots_from_names(["bias", "beta"])

    def guarded(n):
        t = torch.full((n + 2 * G,), SENT, dtype=torch.float64, device=dev)
        return t, t.data_ptr() + 8 * G

    sizes = {"chi2": B, "chi2_zA": B * n_z * n_A, "model": B * n_z * n_A * n_M, "px": B * n_z * n_a * n_m}
    bufs = {k: guarded(n) for k, n in sizes.items()}
    st = torch.cuda.current_stream(dev).cuda_stream
    eng.eval_device(B, 2, slots, points_ptr=pts.data_ptr(), zsel=zsel, chi2_ptr=bufs["chi2"][1],
                    chi2_zA_ptr=bufs["chi2_zA"][1], stream=st)
    eng.eval_device(B, 2, slots, points_ptr=pts.data_ptr(), zsel=zsel, model_ptr=bufs["model"][1], px_ptr=bufs["px"][1],
                    stream=st)
    torch.cuda.synchronize()
    for k, (t, _) in bufs.items():
        assert bool((t[:G] == SENT).all()) and bool((t[-G:] == SENT).all()), "write outside the %s buffer" % k
        inner = t[G:-G]
        assert bool(torch.isfinite(inner).all()) and not bool((inner == SENT).any()), "%s not fully written" % k
    chi2 = bufs["chi2"][0][G:-G].cpu().numpy()
    np.testing.assert_allclose(chi2, bufs["chi2_zA"][0][G:-G].cpu().numpy().reshape(B, -1).sum(axis=1), rtol=1e-12)
    np.testing.assert_allclose(chi2[:50], joint.get_chi2_batch(pts[:50].cpu().numpy(), ["bias", "beta"], factor=False), rtol=1e-9)


def test_kpar0_reference_mode_on_device():
    """Theory(config={'kpar0': 'reference'}): every P3D model sees k_par = 1e-4 1/Mpc in the k_par = 0 row, as the
    reference's in-tree Px path does (lyaP3D.py:35-39), while the continuum distortion and the sky term keep the true
    k_par (theory.py:476-535).  Device against the oracle with the same substitution; all contaminants on."""
    d = synthetic.make_measurement_dict("S1", N_A=6, N_m=40)
    assert d["k_m"][0] == 0.0
    res = {}
    for mode in ("exact", "reference"):
        th = H.make_theory(2.2, "best_fit_arinyo_from_gadget", FLAG_SETS["all"], extra={"kpar0": mode})
        like = Likelihood(DESI_DR2(d), th, 0, config={"N_theta_average": 2})
        truth = {"bias": -0.115, "beta": 1.5}
        filled = synthetic.fill_from_model(d, like.get_convolved_px(truth)[None], add_noise=True)
        like = Likelihood(DESI_DR2(filled), th, 0, config={"N_theta_average": 2})
        orc = H.oracle_likelihood(like)
        assert (orc.theory.hankel.kpar0 is not None) == (mode == "reference")
        for params in (truth, {"bias": -0.12, "beta": 1.6, "b_H": -0.03, "kC_Mpc": 0.02, "b_X": -0.004}):
            px = like._compute_Px_Zam(params)
            assert_px(px, orc.Px_Zam(params), msg="Px_Zam, kpar0 = %s" % mode)
            assert_chi2(like.get_chi2(params), orc.get_chi2(params), "chi2, kpar0 = %s" % mode, n_data=like.get_ndata())
        res[mode] = like._compute_Px_Zam(truth)
        # the factorised path builds its tables from the same plan
        pts = np.stack([np.linspace(-0.12, -0.11, 32), np.linspace(1.4, 1.6, 32)], axis=1)
        np.testing.assert_allclose(like.get_chi2_batch(pts, ["bias", "beta"]), like.get_chi2_batch(pts, ["bias", "beta"], factor=False),
                                   rtol=2e-6, atol=1e-3)
    # continuum distortion on: tanh((k_par / kC)^pC) = 0 at the true k_par = 0 in both modes
    assert np.all(res["exact"][:, 0] == 0.0) and np.all(res["reference"][:, 0] == 0.0)
    np.testing.assert_array_equal(res["exact"][:, 1:], res["reference"][:, 1:])
    th0 = H.make_theory(2.2, "best_fit_arinyo_from_gadget", {"include_hcd": True}, extra={"kpar0": "exact"})
    th1 = H.make_theory(2.2, "best_fit_arinyo_from_gadget", {"include_hcd": True}, extra={"kpar0": "reference"})
    theta, k_AA = np.array([1.0, 5.0, 30.0]), d["k_m"][:5]
    p0, p1 = th0.get_px_obs(theta, k_AA), th1.get_px_obs(theta, k_AA)
    np.testing.assert_array_equal(p0[:, 1:], p1[:, 1:])
    rel = np.abs(p0[:, 0] - p1[:, 0]) / np.abs(p0).max()
    assert 2e-5 < rel.max() < 1e-3                      # 1.7e-4 on the S1 shape (profiles/transform_choice_r02.txt)


def test_theory_component_accessors_vs_oracle():
    """The per-component Px accessors notebooks call (theory.py:126-247, 455-535) on the device path, against the
    oracle's terms; their sum reproduces get_px_obs."""
    th = H.make_theory(2.4, "best_fit_arinyo_from_gadget", FLAG_SETS["all"])
    orc = H.oracle_for(th)
    theta, k_AA = G["theta_arc"], G["k_AA"]
    params = {"bias": -0.13, "beta": 1.4, "b_H": -0.03, "b_X": -0.007, "beta_X": 0.6, "b_noise_Mpc": 0.002, "kC_Mpc": 0.03}
    from oracle.px_oracle import LR_LYA, LR_METAL
    lya = orc._px_obs(theta, k_AA, orc.z, LR_LYA, lambda k, mu: orc.p3d_lya(k, mu, params))
    hcd = orc._px_obs(theta, k_AA, orc.z, LR_LYA, lambda k, mu: orc.p3d_lya_hcd(k, mu, params))
    auto = orc._px_obs(theta, k_AA, orc.z_metal_auto(), LR_METAL, lambda k, mu: orc.p3d_metal_auto(k, mu, params))
    cross = orc._px_obs(theta, k_AA, orc.z_metal_cross(), np.sqrt(LR_METAL * LR_LYA), lambda k, mu: orc.p3d_metal_cross(k, mu, params))
    scale = np.abs(hcd).max()
    assert_px(th.get_px_lya_obs(theta, k_AA, params=params), lya, msg="lya")
    assert_px(th.get_px_lya_hcd_obs(theta, k_AA, params=params), hcd, msg="lya + hcd")
    assert np.abs(th.get_px_metal_auto_obs(theta, k_AA, params=params) - auto).max() <= PX_RTOL * scale
    assert np.abs(th.get_px_metal_cross_obs(theta, k_AA, params=params) - cross).max() <= PX_RTOL * scale
    assert np.abs(th.get_px_sky_obs(theta, k_AA, params=params) - orc.px_sky_obs(theta, k_AA, params)).max() <= PX_RTOL * scale
    cont = th.get_continuum_distortion_obs(k_AA, params=params)
    np.testing.assert_allclose(cont, orc.continuum_distortion_obs(k_AA, params), rtol=1e-14)
    total = (th.get_px_lya_hcd_obs(theta, k_AA, params=params) + th.get_px_metal_auto_obs(theta, k_AA, params=params)
             + th.get_px_metal_cross_obs(theta, k_AA, params=params) + th.get_px_sky_obs(theta, k_AA, params=params)) * cont[None, :]
    assert_px(total, th.get_px_obs(theta, k_AA, params=params), tol=1e-7, msg="sum of the components")


def test_px_from_arbitrary_p3d_callable_and_Mpc_accessors():
    """Theory._compute_px_from_p3d (theory.py:250-264) with an arbitrary P3D callable -- cells sampled by the callable on
    the host, Hankel sum on the device (cpx_hankel_apply) -- and the reference's *_Mpc accessors (theory.py:311-322,
    347-358, 384-395, 441-452, 476-506, 525-535)."""
    th = H.make_theory(2.4, "best_fit_arinyo_from_gadget", FLAG_SETS["all"])
    orc = H.oracle_for(th)
    params = {"bias": -0.13, "beta": 1.4, "q1": 0.5, "b_H": -0.03, "b_X": -0.007, "beta_X": 0.6, "b_noise_Mpc": 0.002}
    rt = np.array([0.05, 0.4, 2.0, 11.0, 37.0, 80.0])
    kp = np.concatenate([[0.0], np.geomspace(0.02, 4.0, 70)])      # 71 values: ragged against the 64-wide warp tile
    # (1) a model P3D through the generic path: float64 cells, so the oracle's node rule to rounding
    got = th._compute_px_from_p3d(rt, kp, lambda k, mu: th.get_p3d_lya_hcd_Mpc(k, mu, params=params), 200.0)
    want = orc.hankel(rt, kp, lambda k, mu: orc.p3d_lya_hcd(k, mu, params))
    assert got.shape == (rt.size, kp.size)
    assert np.abs(got - want).max() <= 1e-10 * np.abs(want).max()      # J0 on the device: 5e-12
    # a node count that is no multiple of four, a single separation, a single k_par
    th50 = H.make_theory(2.4, "best_fit_arinyo_from_gadget", FLAG_SETS["all"], n_q=50)
    orc50 = H.oracle_for(th50)
    got50 = th50._compute_px_from_p3d(rt[2:3], kp[5:6], lambda k, mu: th50.get_p3d_lya_Mpc(k, mu, params=params))
    want50 = orc50.hankel(rt[2:3], kp[5:6], lambda k, mu: orc50.p3d_lya(k, mu, params))
    assert got50.shape == (1, 1) and abs(got50[0, 0] / want50[0, 0] - 1) <= 1e-10
    # ... and the fused device path (float32 cells) for the same term, in Mpc units
    fused = th.get_px_lya_hcd_Mpc(rt, kp, params=params)
    assert np.abs(fused - want).max() <= PX_RTOL * np.abs(want).max()
    # (2) a P3D that is no model of the package, with a closed-form transform:
    #     P3D = exp(-s^2 k^2)  ->  Px = exp(-s^2 k_par^2) exp(-r^2 / 4 s^2) / (4 pi s^2)
    s2 = 0.6 ** 2
    got = th._compute_px_from_p3d(rt, kp, lambda k, mu: np.exp(-s2 * k ** 2), 200.0)
    exact = np.exp(-s2 * kp[None, :] ** 2) * np.exp(-rt[:, None] ** 2 / (4 * s2)) / (4 * np.pi * s2)
    assert np.abs(got - exact).max() <= 3e-5 * np.abs(exact).max()
    # kt_Mpc_max cuts the integrand in k_perp: int_0^K = (1 - exp(-s^2 K^2)) / (4 pi s^2) at r = 0 (here r = 0.05: J0 ~ 1)
    cut = th._compute_px_from_p3d(rt[:1], kp[:1], lambda k, mu: np.exp(-s2 * k ** 2), 1.0)
    assert abs(cut[0, 0] / ((1 - np.exp(-s2)) / (4 * np.pi * s2)) - 1) < 0.15 and cut[0, 0] < 0.7 * got[0, 0]
    # (3) the *_Mpc accessors against the oracle's terms in the units of their own redshift
    from oracle.px_oracle import LR_LYA, LR_METAL
    scale = np.abs(want).max()
    np.testing.assert_allclose(th.get_px_lya_Mpc(rt, kp, params=params), orc.hankel(rt, kp, lambda k, mu: orc.p3d_lya(k, mu, params)),
                               rtol=0, atol=PX_RTOL * scale)
    np.testing.assert_allclose(th.get_px_metal_auto_Mpc(rt, kp, params=params),
                               orc.hankel(rt, kp, lambda k, mu: orc.p3d_metal_auto(k, mu, params)), rtol=0, atol=PX_RTOL * scale)
    np.testing.assert_allclose(th.get_px_metal_cross_Mpc(rt, kp, params=params),
                               orc.hankel(rt, kp, lambda k, mu: orc.p3d_metal_cross(k, mu, params)), rtol=0, atol=PX_RTOL * scale)
    dAA = th.fid_cosmo.get_dAA_dMpc(th.z, lambda_rest_AA=LR_LYA)
    theta = rt * th.fid_cosmo.get_darc_dMpc(th.z)
    np.testing.assert_allclose(th.get_px_sky_Mpc(rt, kp, params=params), orc.px_sky_obs(theta, kp / dAA, params) / dAA,
                               rtol=0, atol=PX_RTOL * scale)
    np.testing.assert_allclose(th.get_continuum_distortion_Mpc(kp, params=params), orc.continuum_distortion_obs(kp / dAA, params), rtol=1e-13)
    b_T, beta_T = th._compute_lya_hcd_biases(kp, th.lya_model.get_lya_params(th.fid_cosmo, params), th.cont_model.get_hcd_params(params))
    F = np.exp(-kp * th.cont_model.get_hcd_params(params)["L_H_Mpc"])
    np.testing.assert_allclose(b_T, -0.13 + -0.03 * F, rtol=1e-15)
    np.testing.assert_allclose(beta_T * b_T, -0.13 * 1.4 + -0.03 * th.cont_model.get_hcd_params(params)["beta_H"] * F, rtol=1e-14)


def test_window_and_rebin_helpers_on_device_px_zam():
    """window_and_rebin.convolve_window / rebin_theta (window_and_rebin.py:59-76) applied on the host to the device's own
    Px_Zam reproduce the device's window stage (likelihood.py:53-88)."""
    from cupix_b200.likelihood.window_and_rebin import convolve_window, rebin_theta
    d = synthetic.make_measurement_dict("S1", N_A=5, fine_per_coarse=3, N_m=45, rebin_k=3, Nz=1)
    data = DESI_DR2(d)
    th = H.make_theory(float(d["z_centers"][0]), "best_fit_arinyo_from_gadget", {"include_hcd": True})
    like = Likelihood(data, th, 0, config={"N_theta_average": 2})
    px_zam = like._compute_Px_Zam({"bias": -0.12})
    out = like.get_convolved_px({"bias": -0.12})
    for A in range(out.shape[0]):
        members = np.where(data.B_A_a[A].astype(bool))[0]
        conv = np.array([convolve_window(data.U_ZaMn[0, a], px_zam[a]) for a in members])
        want = rebin_theta(data.V_ZaM[0, members], conv)
        assert np.abs(out[A] - want).max() <= 1e-12 * np.abs(want).max()

"""GPU parity of the tensor-core Hankel path (k_p3d_hankel_tc, opt-in with CPX_TC=1) against the default
FP64 DMMA path and against the float64 emulation of the plan: same cells, so the two device paths may
differ only by the error of the integer split (<= 1e-8 of the per-theta Px scale, tests/test_tc_split.py)."""
import numpy as np
import pytest

from cupix_b200 import plan as _plan
from cupix_b200.engine import PxEngine, slots_from_names
from cupix_b200.likelihood.likelihood import Likelihood
from cupix_b200.px_data import synthetic
from cupix_b200.px_data.data_DESI_DR2 import DESI_DR2
from tests import helpers as H

pytestmark = pytest.mark.gpu

ALL = {"include_hcd": True, "include_metal": True, "include_sky": True, "include_continuum": True}


def _plans(shape, n_q, flags, N_t=5, **over):
    d = synthetic.make_measurement_dict(shape, **over)
    likes = [Likelihood(DESI_DR2(d), H.make_theory(float(z), "best_fit_arinyo_from_gadget", flags, n_q=n_q), iz,
                        config={"N_theta_average": N_t}) for iz, z in enumerate(d["z_centers"])]
    return [lk.build_plan() for lk in likes]


def _both_paths(monkeypatch, plans, pts, names, want=("chi2", "px_zam")):
    monkeypatch.setenv("CPX_IMMA", "0")          # the reference of these tests is the FP64 DMMA kernel
    slots, zsel = slots_from_names(names)
    out = {}
    for tc in ("0", "1"):
        monkeypatch.setenv("CPX_TC", tc)
        eng = PxEngine(plans)
        out[tc] = eng.eval(pts, slots, zsel=zsel, want=want)
        eng.close()
    return out["0"], out["1"]


@pytest.mark.parametrize("n_q,flags", [(88, {"include_hcd": True}), (64, {}), (96, ALL), (80, ALL)])
def test_tensor_path_equals_dmma_path(monkeypatch, n_q, flags):
    plans = _plans("S2", n_q, flags, N_m=70, N_A=9)          # ragged: 70 rows (9 tiles of 8), 9 theta bins, 4 z
    names = ["bias", "beta", "q1", "q2", "kv_Mpc", "av", "bv", "kp_Mpc", "b_H", "b_X", "b_noise_Mpc", "pC"]
    base = np.array([plans[0].defaults[_plan.SLOT_INDEX[n]] for n in names])
    rng = np.random.default_rng(11)
    pts = base[None, :] * rng.uniform(0.7, 1.3, size=(37, len(names)))      # 37 points: two full groups of 16 + a tail of 5
    ref, got = _both_paths(monkeypatch, plans, pts, names)
    scale = np.abs(ref["px_zam"]).max(axis=3, keepdims=True)
    scale = np.where(scale > 0, scale, 1.0)
    assert (np.abs(got["px_zam"] - ref["px_zam"]) / scale).max() < 1e-8
    np.testing.assert_allclose(got["chi2"], ref["chi2"], rtol=1e-9, atol=1e-7)
    # and against the float64 NumPy emulation of the plan (float32 cell rounding dominates: 1e-5 is the stated Px tolerance)
    for i in (0, 36):
        for iz, p in enumerate(plans):
            px = H.emulate_plan_px(p, dict(zip(names, pts[i])))
            assert np.abs(got["px_zam"][i, iz, :p.n_a, :p.n_m] - px).max() / np.abs(px).max() < 1e-6


def test_tensor_path_more_than_32_theta_bins_and_single_point(monkeypatch):
    plans = _plans("S1", 88, {"include_hcd": True}, N_t=1, N_A=40, N_m=40)   # 40 fine theta bins: two theta blocks
    names = ["bias", "beta"]
    pts = np.array([[-0.12, 1.4]])
    ref, got = _both_paths(monkeypatch, plans, pts, names)
    scale = np.abs(ref["px_zam"]).max(axis=3, keepdims=True)
    assert (np.abs(got["px_zam"] - ref["px_zam"]) / scale).max() < 1e-8
    np.testing.assert_allclose(got["chi2"], ref["chi2"], rtol=1e-9, atol=1e-7)


def test_tensor_path_extreme_parameters_stay_finite(monkeypatch):
    """Cells spanning hundreds of decades (kp = 0.5 or 300 Mpc^-1) go through the per-row block exponent."""
    plans = _plans("tiny", 88, {"include_hcd": True})
    names = ["kp_Mpc", "q1", "q2"]
    pts = np.array([[0.5, 3.0, 0.0], [300.0, 0.0, 0.0], [12.0, 0.0, 1.6], [3.0, 2.0, 1.0]])
    ref, got = _both_paths(monkeypatch, plans, pts, names)
    assert np.isfinite(got["px_zam"]).all()
    scale = np.abs(ref["px_zam"]).max(axis=3, keepdims=True)
    assert (np.abs(got["px_zam"] - ref["px_zam"]) / scale).max() < 1e-8


def _imma_paths(monkeypatch, plans, pts, names, want=("chi2", "px_zam")):
    """The FP64 DMMA kernel (default) and k_p3d_hankel_imma (INT8 mma.sync split, CPX_IMMA=1); the second engine must
    really run that kernel."""
    slots, zsel = slots_from_names(names)
    out = {}
    for im in ("0", "1"):
        monkeypatch.setenv("CPX_IMMA", im)
        eng = PxEngine(plans)
        assert eng.info("hankel_kernel") == (2 if im == "1" else 0)
        out[im] = eng.eval(pts, slots, zsel=zsel, want=want)
        eng.close()
    return out["0"], out["1"]


@pytest.mark.parametrize("n_q,flags", [(56, {"include_hcd": True}), (88, {}), (88, ALL), (56, ALL), (56, {"include_continuum": True})])
def test_imma_path_equals_dmma_path(monkeypatch, n_q, flags):
    """Same FP32 cells on both paths, so they may differ only by the error of the integer split (tests/test_tc_split.py)."""
    plans = _plans("S2", n_q, flags, N_m=70, N_A=9)          # ragged: 70 rows (3 tiles of 32, the last with 6 rows), 9 theta bins, 4 z
    names = ["bias", "beta", "q1", "q2", "kv_Mpc", "av", "bv", "kp_Mpc", "b_H", "b_X", "b_noise_Mpc", "pC"]
    base = np.array([plans[0].defaults[_plan.SLOT_INDEX[n]] for n in names])
    rng = np.random.default_rng(12)
    pts = base[None, :] * rng.uniform(0.7, 1.3, size=(37, len(names)))      # 37 points: four full groups of 8 + a tail of 5
    ref, got = _imma_paths(monkeypatch, plans, pts, names)
    scale = np.abs(ref["px_zam"]).max(axis=3, keepdims=True)
    scale = np.where(scale > 0, scale, 1.0)
    assert (np.abs(got["px_zam"] - ref["px_zam"]) / scale).max() < 2e-8
    np.testing.assert_allclose(got["chi2"], ref["chi2"], rtol=2e-9, atol=1e-6)
    for i in (0, 36):
        for iz, p in enumerate(plans):
            px = H.emulate_plan_px(p, dict(zip(names, pts[i])))
            assert np.abs(got["px_zam"][i, iz, :p.n_a, :p.n_m] - px).max() / np.abs(px).max() < 1e-6


def test_imma_path_theta_blocks_cosmo_columns_and_extremes(monkeypatch):
    plans = _plans("S1", 88, {"include_hcd": True}, N_t=1, N_A=40, N_m=40)   # 40 fine theta bins: two theta blocks of 24
    ref, got = _imma_paths(monkeypatch, plans, np.array([[-0.12, 1.4]]), ["bias", "beta"])
    scale = np.abs(ref["px_zam"]).max(axis=3, keepdims=True)
    assert (np.abs(got["px_zam"] - ref["px_zam"]) / scale).max() < 2e-8
    np.testing.assert_allclose(got["chi2"], ref["chi2"], rtol=2e-9, atol=1e-6)
    # primordial amplitude and tilt as columns: such calls keep the DMMA kernel (no COSMO instantiation in the build), all contaminants on
    plans = _plans("tiny", 56, ALL)
    d0 = plans[0].defaults
    pts = np.array([[d0[_plan.SLOT_INDEX["As"]] * f, d0[_plan.SLOT_INDEX["ns"]] + dn, -0.12] for f, dn in ((1.0, 0.0), (1.1, 0.02), (0.9, -0.03))])
    ref, got = _imma_paths(monkeypatch, plans, pts, ["As", "ns", "bias"])
    scale = np.abs(ref["px_zam"]).max(axis=3, keepdims=True)
    assert (np.abs(got["px_zam"] - ref["px_zam"]) / scale).max() < 2e-8
    # cells spanning hundreds of decades go through the per-row block exponent
    plans = _plans("tiny", 88, {"include_hcd": True})
    pts = np.array([[0.5, 3.0, 0.0], [300.0, 0.0, 0.0], [12.0, 0.0, 1.6], [3.0, 2.0, 1.0]])
    ref, got = _imma_paths(monkeypatch, plans, pts, ["kp_Mpc", "q1", "q2"])
    assert np.isfinite(got["px_zam"]).all()
    scale = np.abs(ref["px_zam"]).max(axis=3, keepdims=True)
    assert (np.abs(got["px_zam"] - ref["px_zam"]) / scale).max() < 5e-8      # 2.6e-8 at these far-out points


def test_imma_path_under_guard_bands(monkeypatch):
    """CPX_GUARD=1 around the per-chunk scratch buffers while k_p3d_hankel_imma writes ragged tiles (70 rows, 9 theta bins,
    37 points): no guard word touched, same numbers as without the guards."""
    plans = _plans("S2", 52, ALL, N_m=70, N_A=9)
    names = ["bias", "beta", "q1", "kp_Mpc", "b_H"]
    base = np.array([plans[0].defaults[_plan.SLOT_INDEX[n]] for n in names])
    pts = base[None, :] * np.random.default_rng(5).uniform(0.8, 1.2, size=(37, len(names)))
    slots, zsel = slots_from_names(names)
    monkeypatch.setenv("CPX_IMMA", "1")
    eng = PxEngine(plans)
    ref = eng.eval(pts, slots, zsel=zsel, want=("chi2", "px_zam"))
    eng.close()
    monkeypatch.setenv("CPX_GUARD", "1")
    eng = PxEngine(plans)
    assert eng.info("hankel_kernel") == 2
    got = eng.eval(pts, slots, zsel=zsel, want=("chi2", "px_zam"))
    assert eng.info("guard_buffers") >= 5 and eng.info("guard_violations") == 0
    eng.close()
    np.testing.assert_array_equal(got["chi2"], ref["chi2"])
    np.testing.assert_array_equal(got["px_zam"], ref["px_zam"])


def test_imma_path_falls_back_for_other_node_counts(monkeypatch):
    monkeypatch.setenv("CPX_IMMA", "1")
    eng = PxEngine(_plans("tiny", 80, {}))
    assert eng.info("hankel_kernel") == 0
    eng.close()


def _window_paths(monkeypatch, plans, pts, names, data_idx=None):
    """chi2 per (z, theta_A) with the window stage on the tensor cores (default) and as the FP64 DMMA GEMM."""
    slots, zsel = slots_from_names(names)
    out = {}
    monkeypatch.setenv("CPX_WTC_MIN", "0")      # small batches would otherwise take the (lower-latency) DMMA kernel
    for wtc in ("0", "1"):
        monkeypatch.setenv("CPX_WTC", wtc)
        eng = PxEngine(plans)
        out[wtc] = eng.eval(pts, slots, zsel=zsel, data_idx=data_idx, want=("chi2", "chi2_zA"))
        eng.close()
    return out["0"], out["1"]


@pytest.mark.parametrize("shape,over,flags", [
    ("S2", dict(N_m=70, N_A=9), ALL),                          # ragged rows / columns, 4 z, additive contaminants
    ("S1_rebin", dict(N_A=6, N_m=60), {"include_hcd": True}),  # three fine theta bins per coarse bin: K spans 3 pairs
    ("tiny", {}, {}),                                          # 8 coarse-k columns: one column half stays empty
])
def test_tensor_window_equals_dmma_window(monkeypatch, shape, over, flags):
    d = synthetic.make_measurement_dict(shape, **over)
    rng = np.random.default_rng(17)
    d["Px_ZAM"] = 0.05 * rng.normal(size=d["Px_ZAM"].shape) + 0.1      # non-trivial data vector and covariance
    cov = np.empty_like(d["cov_ZAM"])
    for iz in range(cov.shape[0]):
        for A in range(cov.shape[1]):
            n = cov.shape[2]
            sig = 0.01 * np.geomspace(1.0, 1e-4, n)                      # sigma falling by 4 decades along k: T rises with 1/sigma
            cov[iz, A] = sig[:, None] * sig[None, :] * 0.3 ** np.abs(np.subtract.outer(np.arange(n), np.arange(n)))
    d["cov_ZAM"] = cov
    likes = [Likelihood(DESI_DR2(d), H.make_theory(float(z), "best_fit_arinyo_from_gadget", flags, n_q=88), iz,
                        config={"N_theta_average": 3}) for iz, z in enumerate(d["z_centers"])]
    plans = [lk.build_plan() for lk in likes]
    names = ["bias", "beta", "q1", "kp_Mpc"]
    base = np.array([plans[0].defaults[_plan.SLOT_INDEX[n]] for n in names])
    pts = base[None, :] * rng.uniform(0.6, 1.4, size=(150, len(names)))          # 150 points: one full tile of 128 + a tail
    ref, got = _window_paths(monkeypatch, plans, pts, names)
    # the integer split is exact to ~2^-40 of the row scales: chi2 values of up to 1e9 agree to better than 1e-9 relative
    np.testing.assert_allclose(got["chi2_zA"], ref["chi2_zA"], rtol=2e-9, atol=1e-9)
    np.testing.assert_allclose(got["chi2"], ref["chi2"], rtol=2e-9, atol=1e-9)
    # and against the float64 emulation of the plan
    for i in (0, 149):
        tot = sum(H.emulate_plan_chi2(p, dict(zip(names, pts[i])))[1].sum() for p in plans)
        np.testing.assert_allclose(got["chi2"][i], tot, rtol=2e-6)


def test_tensor_window_many_mocks_data_index(monkeypatch):
    d = synthetic.make_measurement_dict("S1", N_A=5, N_m=48)
    like = Likelihood(DESI_DR2(d), H.make_theory(2.2, "best_fit_arinyo_from_gadget", {}, n_q=88), 0, config={"N_theta_average": 1})
    p = like.build_plan()
    rng = np.random.default_rng(3)
    mocks = 0.1 + 0.02 * rng.normal(size=(4,) + p.Px_AM.shape[1:])
    p.Px_AM, p.n_data = mocks, 4
    pts = np.stack([rng.uniform(-0.2, -0.1, 300), rng.uniform(1.0, 2.0, 300)], axis=1)
    idx = rng.integers(0, 4, 300)
    ref, got = _window_paths(monkeypatch, [p], pts, ["bias", "beta"], data_idx=idx)
    np.testing.assert_allclose(got["chi2"], ref["chi2"], rtol=2e-9, atol=1e-9)


@pytest.mark.parametrize("tc", ["0", "1"])
def test_scratch_guard_bands_stay_intact(monkeypatch, tc):
    """CPX_GUARD=1 fences every writable scratch buffer of the handle (per-point parameters, row parameters,
    Px_Zam, chi2 partials, lazy model / Px outputs) with 64 KiB guard bands.  After every kind of evaluation --
    both Hankel kernels, the float64 cells, both window kernels with partial tiles, several chunks, many data
    vectors -- no band byte has changed.  Shape with padding in every dimension."""
    monkeypatch.setenv("CPX_GUARD", "1")
    monkeypatch.setenv("CPX_TC", tc)
    monkeypatch.setenv("CPX_SCRATCH_MB", "64")               # small chunks: the 9000-point batch spans several
    plans = _plans("S2", 88, ALL, N_t=2, N_m=70, N_A=9, fine_per_coarse=2)
    rng = np.random.default_rng(4)
    for p in plans:
        p.Px_AM = np.concatenate([p.Px_AM, p.Px_AM * 1.01, p.Px_AM * 0.99])
        p.n_data = 3
    names = ["bias", "beta", "q1", "b_H", "b_X", "As"]
    slots, zsel = slots_from_names(names)
    base = np.array([plans[0].defaults[_plan.SLOT_INDEX[n]] for n in names])
    for precise in (False, True):
        eng = PxEngine(plans, precise=precise)
        eng.eval(base[None, :], slots, zsel=zsel)                                  # the per-chunk scratch is allocated on demand
        assert eng.info("guard_buffers") >= 5 and eng.info("guard_violations") == 0
        for B in ((1, 77, 9000, 49152 // 4 + 13) if not precise else (1, 77)):
            pts = base[None, :] * rng.uniform(0.8, 1.2, size=(B, len(names)))
            out = eng.eval(pts, slots, zsel=zsel, want=("chi2", "chi2_zA"), data_idx=rng.integers(0, 3, B).astype(np.int32))
            assert np.isfinite(out["chi2"]).all()
            if B <= 77:
                out = eng.eval(pts, slots, zsel=zsel, want=("model", "px_zam"))
                assert np.isfinite(out["model"]).all() and np.isfinite(out["px_zam"]).all()
            eng.eval(pts[:, :2], slots[:2], zsel=zsel[:2], want=("chi2",))          # without the As column
            assert eng.info("guard_violations") == 0, "scratch overrun after B = %d (precise = %s)" % (B, precise)
        assert eng.info("chunk") < 9000 or precise
        eng.close()

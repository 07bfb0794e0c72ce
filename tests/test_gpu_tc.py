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

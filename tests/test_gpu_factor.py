"""GPU: the factorised chi2 path (cpx_factor.cuh) -- amplitude parameters (bias, beta, b_H, beta_H, b_X, beta_X,
b_noise_Mpc) taken out of the Hankel stage, chi2 per point as a quadratic form in window-convolved basis vectors.

Checked against the float64 oracle (the path evaluates its few P3D cells in float64, so the tolerance here is 1e-8
relative on chi2 everywhere, not the near-best-fit rule of the float32 pipeline) and against the full pipeline in its
float64-cell mode (CPX_PRECISE_CELLS, factor=False) to 1e-9.
Reference workloads: scripts/chi2_scan_mock.py:107-140 (bias x beta grid), scripts/chi2_scan_forecast.py:116-125.
"""
import numpy as np
import pytest

from cupix_b200 import plan as _plan
from cupix_b200.likelihood.likelihood import JointLikelihood, Likelihood
from cupix_b200.px_data import synthetic
from cupix_b200.px_data.data_DESI_DR2 import DESI_DR2
from tests import helpers as H
from tests.golden.ref_theory_param_sets import FLAG_SETS

pytestmark = pytest.mark.gpu

RTOL = 1e-8


def _like(shape="S1", flags=None, model="best_fit_arinyo_from_gadget", N_t=2, precise=False, n_q=96, seed=3, **over):
    d = synthetic.make_measurement_dict(shape, **over)
    likes = []
    theories = [H.make_theory(float(z), model, flags or {}, n_q=n_q) for z in d["z_centers"]]
    joint = JointLikelihood([Likelihood(DESI_DR2(d), th, iz, config={"N_theta_average": N_t}) for iz, th in enumerate(theories)])
    truth = joint.get_convolved_px_batch(np.zeros((1, 0)), [])[0]
    filled = synthetic.fill_from_model(d, truth, rng=np.random.default_rng(seed), add_noise=True)
    data = DESI_DR2(filled)
    for iz, th in enumerate(theories):
        likes.append(Likelihood(data, th, iz, config={"N_theta_average": N_t, "precise_cells": precise}))
    return data, likes


def _grid_points(lo, hi, n, first=0, count=None):
    total = int(np.prod(n))
    count = total - first if count is None else count
    axes = [np.linspace(lo[j], hi[j], n[j]) for j in range(len(n))]
    sub = np.stack(np.unravel_index(np.arange(first, first + count), n), axis=1)
    return np.stack([axes[j][sub[:, j]] for j in range(len(n))], axis=1)


def test_bias_beta_scan_is_one_tuple(monkeypatch):
    """chi2_scan_mock.py: a (bias, beta) grid shares one tuple -- the Hankel stage runs once for 128 x 128 points."""
    data, likes = _like(N_A=8, N_m=60, flags={"include_hcd": True})
    like = likes[0]
    d0 = like.build_plan().defaults
    lo, hi, n = [d0[0] - 0.02, d0[1] - 0.2], [d0[0] + 0.02, d0[1] + 0.2], [128, 128]
    eng = like.engine()
    chi2 = like.get_chi2_grid(["bias", "beta"], lo, hi, n)
    assert eng.info("factored_calls") == 1 and eng.info("factored_tuples") == 1 and eng.info("factored_cache_hits") == 0
    pts = _grid_points(lo, hi, n)
    orc = H.oracle_likelihood(like)
    for i in (0, 127, 128 * 64 + 64, 128 * 128 - 1, 5000, 11111):
        ref = orc.get_chi2({"bias": pts[i, 0], "beta": pts[i, 1]})
        assert abs(chi2[i] - ref) <= RTOL * ref + 1e-9, "point %d: %.10g vs oracle %.10g" % (i, chi2[i], ref)
    # explicit points whose columns are all amplitude parameters take the same path, with the same numbers
    np.testing.assert_array_equal(like.get_chi2_batch(pts[4000:4100], ["bias", "beta"]), chi2[4000:4100])
    assert eng.info("factored_calls") == 2
    # ... and, the tuple being the one of the previous call, re-use its Gram data: that call was a single kernel launch
    assert eng.info("factored_cache_hits") == 1 and eng.info("factored_tuples") == 1
    l0 = eng.info("kernel_launches")
    like.get_chi2_batch(pts[:64], ["bias", "beta"])
    assert eng.info("kernel_launches") == l0 + 1 and eng.info("factored_cache_hits") == 2
    # new data vectors invalidate it
    like.set_mock_data(like.data.Px_ZAM[0] * 1.01)
    moved = like.get_chi2_batch(pts[4000:4100], ["bias", "beta"])
    assert eng.info("factored_cache_hits") == 2 and np.abs(moved - chi2[4000:4100]).max() > 1e-3
    like.restore_data()
    np.testing.assert_array_equal(like.get_chi2_batch(pts[4000:4100], ["bias", "beta"]), chi2[4000:4100])
    # a rank-style slice of the grid
    np.testing.assert_array_equal(like.get_chi2_grid(["bias", "beta"], lo, hi, n, first=777, count=3000), chi2[777:3777])
    # against the full pipeline: float32 cells within its own tolerance, and few points stay on the full pipeline
    full = like.get_chi2_grid(["bias", "beta"], lo, hi, n, first=8000, count=512, factor=False)
    np.testing.assert_allclose(full, chi2[8000:8512], rtol=2e-6, atol=1e-3)
    calls = eng.info("factored_calls")
    like.get_chi2_batch(pts[:5], ["bias", "beta"])
    assert eng.info("factored_calls") == calls
    # the two-kernel form of the window / Gram stage (k_window + k_gram; taken for n_M > 64 or many data vectors) against the
    # fused k_window_gram: the same sums in another order
    monkeypatch.setenv("CPX_FUSED_GRAM", "0")
    monkeypatch.setenv("CPX_BASIS_CACHE", "0")
    np.testing.assert_allclose(like.get_chi2_grid(["bias", "beta"], lo, hi, n), chi2, rtol=1e-12)


@pytest.mark.parametrize("names,n", [(["bias", "beta", "q1", "kp_Mpc"], [7, 6, 3, 4]),
                                      (["q1", "kp_Mpc", "bias", "beta"], [3, 4, 7, 6]),
                                      (["bias", "q1", "beta", "kp_Mpc"], [9, 2, 8, 3])])
def test_four_parameter_grid_slices_vs_full_pipeline_and_oracle(names, n, monkeypatch):
    """chi2_scan_forecast.py: (bias, beta, q1, kp_Mpc) grids in any axis order; whole grid and unaligned slices
    (CPX_FACTOR_MIN=1 keeps even slices with few points per tuple on the factorised path: the decomposition of a flat
    index range into boxes is the subject)."""
    monkeypatch.setenv("CPX_FACTOR_MIN", "1")
    data, likes = _like(N_A=6, N_m=48, flags={"include_hcd": True}, N_t=1)
    data_p, likes_p = _like(N_A=6, N_m=48, flags={"include_hcd": True}, N_t=1, precise=True)
    like, like_p = likes[0], likes_p[0]
    d0 = like.build_plan().defaults
    half = {"bias": 0.02, "beta": 0.2, "q1": 0.1, "kp_Mpc": 3.0}
    lo = [d0[_plan.SLOT_INDEX[k]] - half[k] for k in names]
    hi = [d0[_plan.SLOT_INDEX[k]] + half[k] for k in names]
    total = int(np.prod(n))
    ref = like_p.get_chi2_grid(names, lo, hi, n, factor=False)                 # float64 cells, every point on the full pipeline
    chi2 = like.get_chi2_grid(names, lo, hi, n)
    n_tuples = n[names.index("q1")] * n[names.index("kp_Mpc")]
    assert like.engine().info("factored_calls") == 1 and like.engine().info("factored_tuples") == n_tuples
    np.testing.assert_allclose(chi2, ref, rtol=1e-9)
    for k, (first, count) in enumerate(((1, total - 2), (total // 3 + 5, total // 2), (17, 400), (total - 300, 300), (total - 1, 1))):
        part = like.get_chi2_grid(names, lo, hi, n, first=first, count=count)
        assert like.engine().info("factored_calls") == 2 + k
        np.testing.assert_allclose(part, ref[first:first + count], rtol=1e-9, err_msg="slice %d+%d" % (first, count))
    orc = H.oracle_likelihood(like)
    pts = _grid_points(lo, hi, n)
    for i in (0, total // 2 + 3, total - 1):
        r = orc.get_chi2(dict(zip(names, pts[i])))
        assert abs(chi2[i] - r) <= RTOL * r + 1e-9


def test_all_contaminants_amplitude_and_row_parameters_vs_oracle():
    """HCD + SiIII + sky + continuum: 16 basis tables.  Amplitude columns (b_H, beta_H, b_X, beta_X, b_noise) inside the
    quadratic form, L_H / kC / pC as tuple axes of a grid."""
    data, likes = _like(N_A=5, N_m=40, flags=FLAG_SETS["all"], model="best_fit_arinyo_from_colore", N_t=2)
    like = likes[0]
    orc = H.oracle_likelihood(like)
    rng = np.random.default_rng(8)
    names = ["bias", "beta", "b_H", "beta_H", "b_X", "beta_X", "b_noise_Mpc"]
    B = 64
    pts = np.stack([rng.uniform(-0.13, -0.10, B), rng.uniform(1.3, 1.7, B), rng.uniform(-0.03, -0.01, B), rng.uniform(0.3, 0.8, B),
                    rng.uniform(-0.01, -0.002, B), rng.uniform(0.3, 0.7, B), rng.uniform(0.001, 0.005, B)], axis=1)
    chi2 = like.get_chi2_batch(pts, names)
    assert like.engine().info("factored_calls") == 1
    for i in range(0, B, 7):
        ref = orc.get_chi2(dict(zip(names, pts[i])))
        assert abs(chi2[i] - ref) <= RTOL * ref + 1e-9, "point %d: %.10g vs %.10g" % (i, chi2[i], ref)
    gnames = ["L_H_Mpc", "bias", "kC_Mpc", "b_X", "pC", "beta"]
    lo, hi, n = [5.0, -0.125, 0.015, -0.008, 0.5, 1.4], [9.0, -0.105, 0.025, -0.003, 0.7, 1.6], [2, 6, 2, 5, 2, 4]
    g = like.get_chi2_grid(gnames, lo, hi, n)
    assert like.engine().info("factored_tuples") == 1 + 8
    gp = _grid_points(lo, hi, n)
    for i in (0, 100, 479, 959):
        ref = orc.get_chi2(dict(zip(gnames, gp[i])))
        assert abs(g[i] - ref) <= RTOL * ref + 1e-9


def test_joint_z_bins_per_z_columns_and_mocks(monkeypatch):
    """Several z bins in one handle: columns restricted to one z (`_<iz>` names), shared columns, and per-point data
    vectors (many mocks): the factorised path against the full pipeline in float64-cell mode."""
    monkeypatch.setenv("CPX_FACTOR_MIN", "1")
    data, likes = _like(shape="S2", N_A=5, N_m=36, Nz=3, flags={"include_hcd": True}, N_t=1)
    data_p, likes_p = _like(shape="S2", N_A=5, N_m=36, Nz=3, flags={"include_hcd": True}, N_t=1, precise=True)
    joint, joint_p = JointLikelihood(likes), JointLikelihood(likes_p)
    rng = np.random.default_rng(6)
    mocks = [data.Px_ZAM[iz][None] * (1 + 0.03 * rng.normal(size=(5,) + data.Px_ZAM[iz].shape)) for iz in range(3)]
    for iz in range(3):
        joint.engine().set_data(iz, mocks[iz])
        joint_p.engine().set_data(iz, mocks[iz])
    d0 = [lk.build_plan().defaults for lk in likes]
    names = ["bias_0", "beta_0", "bias_2", "b_H", "q1_1"]
    lo = [d0[0][0] - 0.01, d0[0][1] - 0.1, d0[2][0] - 0.01, -0.03, d0[1][2] - 0.05]
    hi = [d0[0][0] + 0.01, d0[0][1] + 0.1, d0[2][0] + 0.01, -0.01, d0[1][2] + 0.05]
    n = [5, 4, 5, 3, 2]
    total = int(np.prod(n))
    didx = rng.integers(0, 5, total).astype(np.int32)
    from cupix_b200.engine import slots_from_names
    slots, zsel = slots_from_names(names)
    got = joint.engine().eval(slots=slots, zsel=zsel, grid=(lo, hi, n), data_idx=didx)["chi2"]
    assert joint.engine().info("factored_calls") == 1 and joint.engine().info("factored_tuples") == 2
    ref = joint_p.engine().eval(slots=slots, zsel=zsel, grid=(lo, hi, n), data_idx=didx, factor=False)["chi2"]
    np.testing.assert_allclose(got, ref, rtol=1e-9)
    part = joint.engine().eval(slots=slots, zsel=zsel, grid=(lo, hi, n), grid_first=123, B=311, data_idx=didx[123:434])["chi2"]
    np.testing.assert_allclose(part, ref[123:434], rtol=1e-9)


def test_As_amplitude_with_metals_and_device_io():
    """As as a tuple axis with SiIII on (the amplitude factors out of the metal tables) and device-pointer I/O on a
    torch stream, without any global synchronisation before the result is consumed on that stream."""
    import torch
    data, likes = _like(N_A=5, N_m=40, flags={"include_hcd": True, "include_metal": True}, model="best_fit_arinyo_from_colore", N_t=1)
    data_p, likes_p = _like(N_A=5, N_m=40, flags={"include_hcd": True, "include_metal": True}, model="best_fit_arinyo_from_colore",
                            N_t=1, precise=True)
    like, like_p = likes[0], likes_p[0]
    names = ["As", "bias", "b_X"]
    lo, hi, n = [1.9e-9, -0.125, -0.008], [2.3e-9, -0.105, -0.003], [3, 16, 9]
    ref = like_p.get_chi2_grid(names, lo, hi, n, factor=False)
    got = like.get_chi2_grid(names, lo, hi, n)
    assert like.engine().info("factored_tuples") == 3
    np.testing.assert_allclose(got, ref, rtol=1e-9)
    # device I/O: amplitude-only points produced by torch kernels on a side stream, consumed on the same stream
    from cupix_b200.engine import slots_from_names
    dev = torch.device("cuda", 0)
    s = torch.cuda.Stream(device=dev)
    B = 20000
    base = torch.tensor([[-0.115, 1.5]], dtype=torch.float64, device=dev)
    with torch.cuda.stream(s):
        pts = (base + 0.01 * torch.randn(B, 2, dtype=torch.float64, device=dev)).contiguous()
        out = torch.empty(B, dtype=torch.float64, device=dev)
        slots, zsel = slots_from_names(["bias", "beta"])
        like.engine().eval_device(B, 2, slots, points_ptr=pts.data_ptr(), chi2_ptr=out.data_ptr(), stream=s.cuda_stream)
        best = out.min().item()
    host = like_p.get_chi2_batch(pts.cpu().numpy(), ["bias", "beta"], factor=False)
    np.testing.assert_allclose(out.cpu().numpy(), host, rtol=1e-9)
    assert best == host.min() or abs(best - host.min()) <= 1e-9 * host.min()


@pytest.mark.parametrize("fused", ["1", "0"])
def test_factorised_path_stays_inside_its_scratch(monkeypatch, fused):
    """CPX_GUARD=1: guard bands around every per-chunk scratch buffer (tuple parameters, the Px_Zam scratch that holds the
    basis tables, the model buffer that holds the Y_k in the two-kernel form).  Several tuple batches (small scratch), all
    16 basis tables, padding in every dimension, many data vectors, partial last batch -- no band byte changes.  The Gram
    buffers, which are plain allocations, are checked through the results: the same numbers as with one big batch."""
    monkeypatch.setenv("CPX_GUARD", "1")
    monkeypatch.setenv("CPX_SCRATCH_MB", "8")
    monkeypatch.setenv("CPX_FUSED_GRAM", fused)
    data, likes = _like(shape="S2", N_A=9, fine_per_coarse=2, N_m=70, Nz=2, flags=FLAG_SETS["all"], model="best_fit_arinyo_from_colore", N_t=2)
    joint = JointLikelihood(likes)
    eng = joint.engine()
    rng = np.random.default_rng(5)
    for iz in range(2):
        eng.set_data(iz, data.Px_ZAM[iz][None] * (1 + 0.02 * rng.normal(size=(3,) + data.Px_ZAM[iz].shape)))
    names = ["q1", "kC_Mpc", "bias", "b_X", "b_noise_Mpc"]
    d0 = likes[0].build_plan().defaults
    lo, hi, n = [0.3, 0.015, d0[0] - 0.01, -0.008, 0.001], [0.6, 0.025, d0[0] + 0.01, -0.003, 0.004], [13, 7, 6, 5, 4]
    total = int(np.prod(n))
    from cupix_b200.engine import slots_from_names
    slots, zsel = slots_from_names(names)
    didx = np.random.default_rng(6).integers(0, 3, total).astype(np.int32)
    got = eng.eval(slots=slots, zsel=zsel, grid=(lo, hi, n), data_idx=didx)["chi2"]
    assert eng.info("factored_calls") == 1 and eng.info("factored_tuples") == 91
    assert eng.info("chunk_allocated") < 91 * 16, "the scratch must be small enough for several tuple batches"
    assert eng.info("guard_buffers") >= 5 and eng.info("guard_violations") == 0
    eng.close()
    # the same evaluation in one tuple batch, guard off: bit for bit the same numbers (tuples are independent of the batching)
    monkeypatch.setenv("CPX_SCRATCH_MB", "4096")
    monkeypatch.setenv("CPX_GUARD", "0")
    big = JointLikelihood(likes).engine()
    rng = np.random.default_rng(5)
    for iz in range(2):
        big.set_data(iz, data.Px_ZAM[iz][None] * (1 + 0.02 * rng.normal(size=(3,) + data.Px_ZAM[iz].shape)))
    ref = big.eval(slots=slots, zsel=zsel, grid=(lo, hi, n), data_idx=didx)["chi2"]
    np.testing.assert_array_equal(got, ref)
    assert np.isfinite(got).all() and np.ptp(got) > 1.0


def test_explicit_meshgrid_points_are_grouped_by_tuple():
    """A meshgrid handed in as an explicit point list -- how the reference's scan scripts build theirs
    (scripts/chi2_scan_mock.py:107-111) -- in any order: the host groups the points by tuple and the call takes the
    factorised path, with the numbers of the on-device grid; unrelated points stay on the full pipeline."""
    data, likes = _like(N_A=6, N_m=48, flags={"include_hcd": True}, N_t=1)
    like = likes[0]
    d0 = like.build_plan().defaults
    names = ["bias", "q1", "beta", "kp_Mpc"]
    lo = [d0[0] - 0.02, d0[2] - 0.1, d0[1] - 0.2, d0[7] - 3.0]
    hi = [d0[0] + 0.02, d0[2] + 0.1, d0[1] + 0.2, d0[7] + 3.0]
    n = [12, 3, 10, 4]
    eng = like.engine()
    grid = like.get_chi2_grid(names, lo, hi, n)
    assert eng.info("factored_calls") == 1 and eng.info("factored_tuples") == 12
    pts = _grid_points(lo, hi, n)
    perm = np.random.default_rng(0).permutation(len(pts))
    got = like.get_chi2_batch(pts[perm], names)
    assert eng.info("factored_calls") == 2 and eng.info("factored_tuples") == 24
    np.testing.assert_array_equal(got, grid[perm])
    # with per-point data vectors, and a subset in which the tuples own different numbers of points
    rng = np.random.default_rng(1)
    like.set_mock_data(like.data.Px_ZAM[0][None] * (1 + 0.02 * rng.normal(size=(3,) + like.data.Px_ZAM[0].shape)))
    sub = perm[:900]
    didx = rng.integers(0, 3, len(sub)).astype(np.int32)
    got = like.get_chi2_batch(pts[sub], names, data_idx=didx)
    ref = like.get_chi2_batch(pts[sub], names, data_idx=didx, factor=False)
    assert eng.info("factored_calls") == 3
    np.testing.assert_allclose(got, ref, rtol=2e-6, atol=1e-3)
    like.restore_data()
    # unrelated points: every tuple is its own -> the full pipeline
    rnd = rng.uniform(lo, hi, size=(2000, 4))
    like.get_chi2_batch(rnd, names)
    assert eng.info("factored_calls") == 3
    # a larger list (sampled first): 4 x 3 tuples x 9000 amplitude points, shuffled
    n2 = [90, 4, 100, 3]
    p2 = _grid_points(lo, hi, n2)
    sh = rng.permutation(len(p2))
    got2 = like.get_chi2_batch(p2[sh], names)
    assert eng.info("factored_calls") == 4
    np.testing.assert_array_equal(got2, like.get_chi2_grid(names, lo, hi, n2)[sh])

"""GPU: the configurations the bench numbers are quoted on, against the float64 ORACLE (not against the
CUDA path itself).

  S5  stress shape at full size (12 z x 20 theta x 256 k_par, N_theta_average = 100, HCD on): truth, near
      points and the far corners of bench.py's scan grid -- Px_Zam to 1e-5 of its maximum and chi2 under
      the stated rule, per z and joint, through both window kernels (k_window_tc needs >= 49152 (point, z)
      pairs, so the points of interest ride at the head of a 4096-point batch)
  S3  chi2_scan_forecast: rank-style slices of the 64^4 (bias, beta, q1, kp_Mpc) grid of z bin 3 against
      the oracle at the same grid points
Reference: cupix/likelihood/likelihood.py:91-132,170-194; scripts/chi2_scan_forecast.py:116-125.
"""
import numpy as np
import pytest

from cupix_b200 import plan as _plan
from cupix_b200.likelihood.likelihood import JointLikelihood, Likelihood
from cupix_b200.px_data import synthetic
from cupix_b200.px_data.data_DESI_DR2 import DESI_DR2
from oracle.hankel import NodeHankel
from tests import helpers as H
from tests.test_gpu_parity import assert_chi2, assert_px

pytestmark = pytest.mark.gpu

BENCH_NAMES = ["bias", "beta", "q1", "kp_Mpc"]
BENCH_HALF = np.array([0.02, 0.2, 0.15, 4.0])      # bench.py grid_bounds


def _build(shape, model, flags, N_t, n_q, noise=True, **over):
    d = synthetic.make_measurement_dict(shape, **over)
    theories = [H.make_theory(float(z), model, flags, n_q=n_q) for z in d["z_centers"]]
    data = DESI_DR2(d)
    likes = [Likelihood(data, th, iz, config={"N_theta_average": N_t}) for iz, th in enumerate(theories)]
    joint = JointLikelihood(likes)
    truth_model = joint.get_convolved_px_batch(np.zeros((1, 0)), [])[0]
    filled = synthetic.fill_from_model(d, truth_model, rng=np.random.default_rng(synthetic.SEED + 1), add_noise=noise)
    data = DESI_DR2(filled)
    likes = [Likelihood(data, th, iz, config={"N_theta_average": N_t}) for iz, th in enumerate(theories)]
    return data, likes, JointLikelihood(likes)


def _oracle(like):
    """Oracle twin whose (un-folded, per-separation) Hankel weights are built with all host cores."""
    r = like.theory.rule
    return H.OracleLikelihood.from_data(like.data, H.oracle_for(like.theory, NodeHankel(r.k, lambda rr: r.weights(rr))),
                                        like.iz, like.N_theta_average)


def test_S5_full_size_vs_oracle():
    data, likes, joint = _build("S5", "best_fit_arinyo_from_gadget", {"include_hcd": True}, N_t=100, n_q=88)
    d0 = _plan.default_vector(likes[0].theory, likes[0].theory.fid_cosmo)
    centre = np.array([d0[_plan.SLOT_INDEX[n]] for n in BENCH_NAMES])
    corner = np.array([[+1, +1, +1, +1], [-1, -1, -1, -1], [+1, -1, +1, -1], [-1, +1, -1, +1]], dtype=float)
    pts = np.concatenate([centre[None],                                            # truth of z bin 0
                          centre[None] + 0.02 * BENCH_HALF * corner[:3],           # near the best fit
                          centre[None] + BENCH_HALF * corner])                     # far corners of the bench grid
    assert pts.shape == (8, 4)
    B = 4096                                                                       # 4096 x 12 z = 49152 pairs -> k_window_tc
    rng = np.random.default_rng(12)
    batch = np.concatenate([pts, rng.uniform(centre - BENCH_HALF, centre + BENCH_HALF, size=(B - len(pts), 4))])
    out_tc = joint.engine().eval(points=batch, slots=[_plan.SLOT_INDEX[n] for n in BENCH_NAMES], want=("chi2", "chi2_zA"))
    out_small = joint.engine().eval(points=pts, slots=[_plan.SLOT_INDEX[n] for n in BENCH_NAMES],
                                    want=("chi2", "chi2_zA", "px_zam"))
    ref_tot = np.zeros(len(pts))
    ref_truth = 0.0
    n_A = len(data.theta_min_A_arcmin)
    worst_px, worst_abs, worst_rel = 0.0, 0.0, 0.0
    for iz, lk in enumerate(likes):
        orc = _oracle(lk)
        for i, p in enumerate(pts):
            params = dict(zip(BENCH_NAMES, p))
            worst_px = max(worst_px, assert_px(out_small["px_zam"][i, iz], orc.Px_Zam(params), msg="Px_Zam pt %d z %d" % (i, iz)))
            c, info = orc.get_chi2(params, return_info=True)
            ref_A = -2 * info["log_like_all"]
            for name, out in (("dmma", out_small), ("tcgen05", out_tc)):
                assert_chi2(out["chi2_zA"][i, iz], ref_A, "%s window, per-theta chi2, pt %d z %d" % (name, i, iz),
                            n_data=lk.get_ndata() / n_A)
            ref_tot[i] += c
            if iz == 0 and i < 4:
                assert c <= 2 * lk.get_ndata(), "truth and near points of z bin 0 must lie inside chi2 <= 2 n_data"
        ref_truth += orc.get_chi2({})
    # every z at its own truth (no columns): the best-fit region of the joint data vector, absolute tolerance
    assert ref_truth <= 2 * joint.get_ndata()
    assert_chi2(joint.get_chi2({}), ref_truth, "joint chi2 at the per-z truth")
    for name, out in (("dmma", out_small), ("tcgen05", out_tc)):
        got = out["chi2"][:len(pts)]
        assert_chi2(got, ref_tot, "%s window, joint chi2" % name, n_data=joint.get_ndata())
        near = ref_tot <= 2 * joint.get_ndata()
        if near.any():
            worst_abs = max(worst_abs, np.abs(got - ref_tot)[near].max())
        if (~near).any():
            worst_rel = max(worst_rel, (np.abs(got - ref_tot) / ref_tot)[~near].max())
    # the factorised path at the full size: an amplitude-only batch (bias, beta applied to all 12 z) is one tuple; float64
    # cells, so the comparison with the oracle is relative 1e-8 at any chi2
    amp = np.column_stack([rng.uniform(centre[0] - 0.02, centre[0] + 0.02, 64), rng.uniform(centre[1] - 0.2, centre[1] + 0.2, 64)])
    f0 = joint.engine().info("factored_calls")
    got_f = joint.get_chi2_batch(amp, ["bias", "beta"])
    assert joint.engine().info("factored_calls") == f0 + 1
    orcs = [_oracle(lk) for lk in likes]
    for i in (0, 31, 63):
        ref_f = sum(o.get_chi2({"bias": amp[i, 0], "beta": amp[i, 1]}) for o in orcs)
        assert abs(got_f[i] - ref_f) <= 1e-8 * ref_f, "factorised path, point %d: %.10g vs oracle %.10g" % (i, got_f[i], ref_f)
    # ... and the float64-cell mode of the full pipeline on the far corners: absolute tolerance at chi2 ~ 3e5
    joint.engine().precise = True
    try:
        got_p = joint.engine().eval(points=pts[4:], slots=[_plan.SLOT_INDEX[n] for n in BENCH_NAMES], factor=False)["chi2"]
    finally:
        joint.engine().precise = False
    assert np.abs(got_p - ref_tot[4:]).max() <= 1e-3, np.abs(got_p - ref_tot[4:]).max()
    print("S5 vs oracle: max |dPx|/max|Px| = %.2e, max |dchi2| (chi2 <= 2 n_data) = %.2e, max rel (beyond) = %.2e, chi2 = %s"
          % (worst_px, worst_abs, worst_rel, np.array2string(ref_tot, precision=1)))


def test_S3_forecast_grid_slices_vs_oracle():
    """The 4-parameter 64^4 scan of chi2_scan_forecast.py on z bin 3 of a 4-z measurement: rank-style slices by
    flat index against the oracle evaluated at the same grid points."""
    data, likes, joint = _build("S2", "best_fit_arinyo_from_gadget", {}, N_t=1, n_q=96, N_m=48, N_A=6)
    like = likes[3]
    d0 = like.build_plan().defaults
    lo = [d0[0] - 0.02, d0[1] - 0.2, d0[2] - 0.1, d0[7] - 3]
    hi = [d0[0] + 0.02, d0[1] + 0.2, d0[2] + 0.1, d0[7] + 3]
    n = [64, 64, 64, 64]
    total = 64 ** 4
    axes = [np.linspace(lo[j], hi[j], n[j]) for j in range(4)]
    orc = H.oracle_likelihood(like)
    for first, count in ((0, 12), (total // 8 * 3 + 17, 12), (total // 2 + 64 ** 3 * 31 + 64 * 17 + 5, 12), (total - 12, 12)):
        chi2 = like.get_chi2_grid(BENCH_NAMES, lo, hi, n, first=first, count=count)
        sub = np.stack(np.unravel_index(np.arange(first, first + count), n), axis=1)
        pts = np.stack([axes[j][sub[:, j]] for j in range(4)], axis=1)
        ref = np.array([orc.get_chi2(dict(zip(BENCH_NAMES, p))) for p in pts])
        assert_chi2(chi2, ref, "slice at %d" % first, n_data=like.get_ndata())

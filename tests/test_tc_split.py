"""CPU check of the error-free integer split behind k_p3d_hankel_tc (cupix_b200/csrc/cpx_hankel_tc.cuh).

The tensor-core Hankel path replaces the float64 contraction  sum_q W[a, q] cell[m, q]  by products of
u8 cell bytes and s8 weight digits accumulated exactly in INT32.  This file restates that pipeline in
NumPy integer arithmetic (same scales, same digit extraction, same dropped slice pairs as
build_zbin / the kernel) and bounds its error against the float64 dot product of the same float32
cells over fiducial, in-box and far-out-of-box parameter points.
"""
import numpy as np
import pytest

from cupix_b200 import plan as _plan
from cupix_b200.likelihood.likelihood import Likelihood
from cupix_b200.px_data import synthetic
from cupix_b200.px_data.data_DESI_DR2 import DESI_DR2
from tests import helpers as H

NA_BYTES, NW_DIGITS = 4, 5


def weight_digits(W):
    """s_q (power of two), t_a and the balanced base-256 digits of W / (s_q t_a), as build_zbin makes them."""
    mx = np.abs(W).max(axis=0)
    s_q = np.where(mx > 0, 2.0 ** np.clip(np.rint(np.log2(np.where(mx > 0, mx, 1.0))), -100, 100), 1.0)
    Ws = W / s_q[None, :]
    t_a = 2.05 * np.abs(Ws).max(axis=1)
    t_a = np.where(t_a > 0, t_a, 1.0)
    X = np.rint(Ws / t_a[:, None] * 256.0 ** NW_DIGITS).astype(np.int64)
    digits = []
    for _ in range(NW_DIGITS):
        d = (X + 128) % 256 - 128
        digits.append(d)
        X = (X - d) // 256
    assert np.all(X == 0)
    return s_q, t_a, digits[::-1]          # most significant digit first


def split_contract(W, cell32):
    """Px[a, m] through the integer pipeline; cell32 [n_m, n_q] holds float32 values."""
    s_q, t_a, B = weight_digits(W)
    c = (cell32.astype(np.float32) * s_q.astype(np.float32)[None, :]).astype(np.float32)
    rmax = c.max(axis=1)
    eb = np.clip(rmax.view(np.uint32) >> 23, 31, 254).astype(np.int64)        # biased exponent of the row maximum
    U = np.rint(c.astype(np.float64) * 2.0 ** (158 - eb)[:, None])             # cvt.rni.u32.f32
    U = np.minimum(U, 2.0 ** 32 - 1).astype(np.uint64)
    A = [((U >> np.uint64(8 * (3 - i))) & np.uint64(255)).astype(np.int64) for i in range(NA_BYTES)]
    S = np.zeros((W.shape[0], c.shape[0]))
    for s in range(NW_DIGITS):                                                 # slice pairs with i + j > 4 are not issued
        D = np.zeros(S.shape, dtype=np.int64)
        for i in range(NA_BYTES):
            j = s - i
            if 0 <= j < NW_DIGITS:
                D += B[j] @ A[i].T
        assert np.abs(D).max() < 2 ** 31
        S += D.astype(np.float64) * 256.0 ** (NW_DIGITS - 1 - s)
    return S * 2.0 ** (eb - 174)[None, :] * t_a[:, None]


def _plan_and_cells(n_q, params):
    d = synthetic.make_measurement_dict("tiny", z0=2.4)
    th = H.make_theory(2.4, "best_fit_arinyo_from_gadget", {"include_hcd": True}, n_q=n_q)
    p = Likelihood(DESI_DR2(d), th, 0, config={"N_theta_average": 7}).build_plan()
    v = np.array(p.defaults, dtype=np.float64)
    for k, x in params.items():
        v[_plan.SLOT_INDEX[k]] = x
    bias, beta, q1, q2, kv, av, bv, kp = v[:8]
    kpar, kperp = p.kpar_Mpc, p.kperp_Mpc
    k2 = kpar[:, None] ** 2 + kperp[None, :] ** 2
    k = np.sqrt(k2)
    pos = k > 0
    ks = np.where(pos, k, 1.0)
    mu = np.where(pos, kpar[:, None] / ks, 0.0)
    D = ks ** 3 * p.linP_cell / (2 * np.pi ** 2)
    with np.errstate(all="ignore"):
        dnl = np.exp(D * (q1 + q2 * D) * (1 - (ks / kv) ** av * mu ** bv) - k2 / kp ** 2)
    cell = np.where(pos, p.linP_cell * (1 + beta * mu ** 2) ** 2 * dnl, 0.0)
    return p, cell.astype(np.float32)


@pytest.mark.parametrize("n_q", [64, 88, 96])
@pytest.mark.parametrize("params", [{}, {"kp_Mpc": 9.0, "q1": 0.15}, {"kp_Mpc": 25.0, "q1": 1.5, "av": 0.9},
                                    {"kp_Mpc": 0.5, "q1": 3.0}, {"kp_Mpc": 300.0, "q1": 0.0, "q2": 0.0}])
def test_integer_split_matches_float64_contraction(n_q, params):
    p, cell32 = _plan_and_cells(n_q, params)
    exact = p.hankel_w @ cell32.astype(np.float64).T
    got = split_contract(p.hankel_w, cell32)
    err = np.abs(got - exact).max(axis=1) / np.abs(exact).max(axis=1)          # per theta bin
    assert err.max() < 1e-8, err.max()
    # and it is far below what rounding the cells to float32 costs in the first place
    assert np.abs(got - exact).max() / np.abs(exact).max() < 2e-9


def test_weight_digits_reconstruct_the_weights():
    p, _ = _plan_and_cells(88, {})
    s_q, t_a, B = weight_digits(p.hankel_w)
    rec = sum(B[j].astype(np.float64) * 256.0 ** -(j + 1) for j in range(NW_DIGITS)) * s_q[None, :] * t_a[:, None]
    assert np.abs(rec - p.hankel_w).max() <= 0.5 * 256.0 ** -NW_DIGITS * (s_q[None, :] * t_a[:, None]).max() * (1 + 1e-12)
    assert all(d.min() >= -128 and d.max() <= 127 for d in B)

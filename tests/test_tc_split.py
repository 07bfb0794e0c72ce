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


def imma_split_contract(W, cell32, na=4, nw=4, smax=3):
    """The split of k_p3d_hankel_imma (cupix_b200/csrc/cpx_hankel_imma.cuh: kImNA = 4 cell bytes, kImNW = 4 weight digits,
    slice pairs i + j <= kImSmax = 3: ten INT8 MMAs per K block and theta tile), restated in NumPy integers."""
    mx = np.abs(W).max(axis=0)
    s_q = np.where(mx > 0, 2.0 ** np.clip(np.rint(np.log2(np.where(mx > 0, mx, 1.0))), -100, 100), 1.0)
    Ws = W / s_q[None, :]
    mxa = np.abs(Ws).max(axis=1)
    t_a = np.where(mxa > 0, 2.0 ** np.ceil(np.log2(2.05 * np.where(mxa > 0, mxa, 1.0))), 1.0)      # a power of two here
    X = np.rint(Ws / t_a[:, None] * 256.0 ** nw).astype(np.int64)
    B = []
    for _ in range(nw):
        d = (X + 128) % 256 - 128
        B.append(d)
        X = (X - d) // 256
    assert np.all(X == 0)
    B = B[::-1]
    c = (cell32.astype(np.float32) * s_q.astype(np.float32)[None, :]).astype(np.float32)
    eb = np.clip(c.max(axis=1).view(np.uint32) >> 23, 31, 254).astype(np.int64)
    U = np.minimum(np.rint(c.astype(np.float64) * 2.0 ** (158 - eb)[:, None]), 2.0 ** 32 - 1).astype(np.uint64)
    A = [((U >> np.uint64(8 * (3 - i))) & np.uint64(255)).astype(np.int64) for i in range(na)]
    H = np.zeros((W.shape[0], c.shape[0]))
    for s in range(smax + 1):
        D = np.zeros(H.shape, dtype=np.int64)
        for i in range(na):
            if 0 <= s - i < nw:
                D += B[s - i] @ A[i].T
        assert np.abs(D).max() < 2 ** 24
        H = H * 256.0 + D
    return H * 2.0 ** (eb - (158 - 8 * (na - 2 - smax)))[None, :] * t_a[:, None]


@pytest.mark.parametrize("n_q", [52, 56, 88, 96])
@pytest.mark.parametrize("params", [{}, {"kp_Mpc": 9.0, "q1": 0.15}, {"kp_Mpc": 25.0, "q1": 1.5, "av": 0.9},
                                    {"kp_Mpc": 0.5, "q1": 3.0}, {"kp_Mpc": 300.0, "q1": 0.0, "q2": 0.0}])
def test_imma_integer_split_matches_float64_contraction(n_q, params):
    """4 x 4 slices, 10 pairs: <= 1e-8 of the Px scale per theta bin -- the float32 rounding of the cells is 3e-8 .. 1e-7."""
    p, cell32 = _plan_and_cells(n_q, params)
    exact = p.hankel_w @ cell32.astype(np.float64).T
    got = imma_split_contract(p.hankel_w, cell32)
    err = np.abs(got - exact).max(axis=1) / np.abs(exact).max(axis=1)
    assert err.max() < 1e-8, err.max()
    # no coherent offset worth the name: the mean signed error over all outputs stays below 3e-9 of the Px scale.  What
    # there is comes from rounding the weights to 32 bits below the power-of-two column scale -- a fixed perturbation
    # of the quadrature rule (whose own error is 1e-6), the same for every parameter point -- amplified where the Hankel
    # sum cancels (the widest theta bins); the MUFU bias the cell evaluation avoids is 5e-9 .. 5e-8 and parameter dependent
    rel = (got - exact) / np.abs(exact).max(axis=1, keepdims=True)
    assert abs(rel.mean()) < 3e-9


def test_weight_digits_reconstruct_the_weights():
    p, _ = _plan_and_cells(88, {})
    s_q, t_a, B = weight_digits(p.hankel_w)
    rec = sum(B[j].astype(np.float64) * 256.0 ** -(j + 1) for j in range(NW_DIGITS)) * s_q[None, :] * t_a[:, None]
    assert np.abs(rec - p.hankel_w).max() <= 0.5 * 256.0 ** -NW_DIGITS * (s_q[None, :] * t_a[:, None]).max() * (1 + 1e-12)
    assert all(d.min() >= -128 and d.max() <= 127 for d in B)


# ------------------------------------------------------------------------------------------------
# k_window_tc (cupix_b200/csrc/cpx_window_tc.cuh): the same split for the window / whitening / chi2 stage
# ------------------------------------------------------------------------------------------------
def window_split_chi2(p, px_am, data_index=0):
    """chi2_A of one point from Px_Zam[a, m] through the integer pipeline of k_window_tc: T_chi with column
    scales 2^c per 16 fine-k values and five balanced digits, the Px row of a theta_A as 40-bit two's-complement
    fixed point below its largest exponent, slice pairs beyond i + j = 4 dropped, Horner in float64."""
    n_m_pad = -(-p.n_m // 32) * 32
    chi2 = []
    for A in range(p.n_A):
        pairs = np.where(p.B_A_a[A] != 0)[0]
        vs = p.V_aM[pairs].sum(axis=0)
        Linv = np.linalg.inv(np.linalg.cholesky(p.cov_AMM[A]))
        T = np.zeros((len(pairs), p.n_M, n_m_pad))
        for i, a in enumerate(pairs):
            T[i, :, :p.n_m] = Linv @ ((p.V_aM[a] / vs)[:, None] * p.U_aMn[a])
        x = np.zeros((len(pairs), n_m_pad))
        x[:, :p.n_m] = px_am[pairs]
        # column scales per (pair, 16 columns), folded onto Px
        blk = np.abs(T).reshape(len(pairs), p.n_M, n_m_pad // 16, 16).max(axis=(1, 3))
        c = np.where(blk > 0, np.clip(np.rint(np.log2(np.where(blk > 0, blk, 1.0))), -300, 300), 0.0)
        cm = np.repeat(c, 16, axis=1)
        Ts = T * 2.0 ** -cm[:, None, :]
        rs = 2.05 * np.abs(Ts).max(axis=(0, 2))
        rs = np.where(rs > 0, rs, 1.0)
        X = np.rint(Ts / rs[None, :, None] * 256.0 ** 5).astype(np.int64)
        digits = []
        for _ in range(5):
            dgt = (X + 128) % 256 - 128
            digits.append(dgt)
            X = (X - dgt) // 256
        assert np.all(X == 0)
        digits = digits[::-1]
        xs = x * 2.0 ** cm
        e = np.frexp(np.abs(xs).max())[1] + 1022            # biased exponent of the row maximum (0 for an all-zero row)
        e = int(min(max(e, 64), 2000))
        I = np.rint(xs * 2.0 ** (1061 - e)).astype(np.int64)   # |I| < 2^39
        assert np.abs(I).max() < 2 ** 39
        lo = I & 0xFFFFFFFF
        sl = [(I >> 32)] + [(lo >> (8 * k)) & 255 for k in (3, 2, 1, 0)]     # signed top byte, then unsigned bytes 3..0
        S = np.zeros(p.n_M)
        for s in range(5):
            D = np.zeros(p.n_M, dtype=np.int64)
            for i in range(5):
                j = s - i
                if 0 <= j < 5:
                    D += np.einsum("pMm,pm->M", digits[j], sl[i])
            assert np.abs(D).max() < 2 ** 31
            S += D.astype(np.float64) * 256.0 ** (4 - s)
        y = S * 2.0 ** (e - 1069) * rs
        res = Linv @ p.Px_AM[data_index, A] - y
        chi2.append(res @ res)
    return np.array(chi2)


@pytest.mark.parametrize("sigma_decades", [0.0, 4.0])
def test_window_integer_split_matches_float64_chi2(sigma_decades):
    d = synthetic.make_measurement_dict("tiny_rebin", z0=2.4)
    rng = np.random.default_rng(5)
    d["Px_ZAM"] = 0.05 * rng.normal(size=d["Px_ZAM"].shape) + 0.1
    n = d["cov_ZAM"].shape[2]
    sig = 0.01 * np.geomspace(1.0, 10.0 ** -sigma_decades, n)           # T_chi rises with 1/sigma along k
    d["cov_ZAM"] = np.broadcast_to(sig[:, None] * sig[None, :] * 0.3 ** np.abs(np.subtract.outer(np.arange(n), np.arange(n))),
                                   d["cov_ZAM"].shape).copy()
    th = H.make_theory(2.4, "best_fit_arinyo_from_gadget", {"include_hcd": True}, n_q=88)
    p = Likelihood(DESI_DR2(d), th, 0, config={"N_theta_average": 3}).build_plan()
    for params in ({}, {"bias": -0.2, "beta": 1.1, "kp_Mpc": 9.0}):
        px = H.emulate_plan_px(p, params)
        ref = H.emulate_plan_chi2(p, params)[1]
        got = window_split_chi2(p, px)
        np.testing.assert_allclose(got, ref, rtol=2e-9, atol=1e-9)

"""Float32 emulation of the P3D cell arithmetic of k_p3d_hankel (cupix_b200/csrc/cpx_kernels.cuh) in NumPy:
every packed instruction is one correctly rounded float32 operation, so the device formulation can be checked
on the CPU against the float64 formula of theory.py:285-308 -- the error budget DESIGN.md section 5 quotes
(per-cell rms ~1e-7, coherent bias of a few 1e-9) and the two bit tricks of the exponentials."""
import numpy as np

f32 = np.float32
LOG2E = 1.4426950408889634


def fma(a, b, c):
    """float32 fused multiply-add (the product of two float32 is exact in float64)"""
    return (np.asarray(a, f32).astype(np.float64) * np.asarray(b, f32).astype(np.float64)
            + np.asarray(c, f32).astype(np.float64)).astype(f32)


def mul(a, b):
    return fma(a, b, f32(0))


def add(a, b):
    return (np.asarray(a, f32).astype(np.float64) + np.asarray(b, f32).astype(np.float64)).astype(f32)


def split(x):
    hi = f32(x)
    return hi, f32(np.float64(x) - np.float64(hi))


def exp2_scale(p, t):
    """p * 2^n with n in the low bits of t: integer add of (bits of t) << 23, modulo 2^32"""
    pb = np.asarray(p, f32).view(np.uint32).astype(np.uint64)
    tb = np.asarray(t, f32).view(np.uint32).astype(np.uint64)
    return ((pb + (tb << np.uint64(23))) & np.uint64(0xFFFFFFFF)).astype(np.uint32).view(f32)


def exp2_reduce(x, magic):
    t = add(x, f32(magic))
    return t, fma(add(t, f32(-magic)), f32(-1), x)


POLY = [1.0, 0.69314718246459961, 0.24022647738456726, 0.055503599345684052, 0.0096185067668557167,
        0.001339080510661006, 0.00015326094580814242]          # exp2_poly2


def exp2_poly(x):
    x = np.minimum(np.maximum(np.asarray(x, f32), f32(-125)), f32(126))
    t, f = exp2_reduce(x, 12582912.0)
    p = np.full_like(f, f32(POLY[-1]))
    for c in POLY[-2::-1]:
        p = fma(p, f, f32(c))
    return exp2_scale(p, t)


def exp2_rr_neg(x):
    """-2^x: MUFU.EX2 on the reduced argument (taken as correctly rounded here), sign through the rounding constant"""
    x = np.maximum(np.asarray(x, f32), f32(-125))
    t, f = exp2_reduce(x, 12582912.0 + 256.0)
    return exp2_scale(np.exp2(f.astype(np.float64)).astype(f32), t)


def test_exponential_bit_tricks():
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(-125, 30, 200000), [-125.0, -0.5, 0.0, 0.5, 1.5, 29.5, -124.5]]).astype(f32)
    ref = np.exp2(x.astype(np.float64))
    neg = exp2_rr_neg(x).astype(np.float64)
    assert np.all(neg < 0)
    assert np.abs(-neg / ref - 1).max() < 1.3e-7            # one float32 rounding of 2^f
    pos = exp2_poly(x).astype(np.float64)
    err = pos / ref - 1
    assert np.abs(err).max() < 1.2e-7 and abs(err.mean()) < 1e-9
    # clamps: below -125 the result stays a normal number, above 126 it saturates
    assert exp2_poly(f32(-300.0)) == f32(2.0 ** -125) and exp2_poly(f32(500.0)) == f32(2.0 ** 126)
    # the degree-6 polynomial itself: 3e-9 on the reduced interval
    f = np.linspace(-0.5, 0.5, 100001)
    assert np.abs(np.polyval(POLY[::-1], f) / 2.0 ** f - 1).max() < 3.1e-9


def test_cell_formulation_error_budget():
    """One row of cells over the k_perp nodes for parameters across the ForestFlow box: the float32 device
    formulation against the float64 formula."""
    rng = np.random.default_rng(1)
    n = 200000
    k = np.exp(rng.uniform(np.log(1e-2), np.log(30.0), n))
    mu = rng.uniform(0.0, 1.0, n)
    PL = 2e3 * k / (1 + (k / 0.02) ** 2.4)                    # smooth stand-in for the linear power
    D = k ** 3 * PL / (2 * np.pi ** 2)
    worst_bias, worst_rms = 0.0, 0.0
    for q1, q2, kv, av, bv, kp, beta in ((0.4, 0.2, 0.6, 0.45, 1.6, 12.0, 1.5), (0.8, 0.0, 1.2, 0.3, 1.9, 20.0, 0.6),
                                          (0.1, 0.6, 0.3, 0.7, 1.4, 8.0, 1.9)):
        ref = PL * (1 + beta * mu ** 2) ** 2 * np.exp(D * (q1 + q2 * D) * (1 - (k / kv) ** av * mu ** bv) - (k / kp) ** 2)
        # cell constants (float32) and per-point scalars (hi + lo pairs), as build_zbin / k_params form them
        cD, cL2K, cMU2, cNK2, cPL = f32(D), f32(np.log2(k)), f32(mu * mu), f32(-k * k), f32(PL)
        cL2MU = np.where(mu > 0, np.log2(np.where(mu > 0, mu, 1.0)), -1e30).astype(f32)
        q1l, q1lo = split(q1 * LOG2E)
        q2l, q2lo = split(q2 * LOG2E)
        av_h, _ = split(av)
        nlkv, nlkvlo = split(-np.float64(av_h) * np.log2(kv))
        ikp2, ikp2lo = split(LOG2E / kp ** 2)
        bhi, blo = split(beta)
        nl = mul(cD, add(fma(q2l, cD, q1l), fma(q2lo, cD, q1lo)))          # log2 units: q1, q2 carry log2(e)
        nv = exp2_rr_neg(add(fma(av_h, cL2K, fma(f32(bv), cL2MU, nlkv)), nlkvlo))
        tt = fma(cNK2, ikp2lo, fma(cNK2, ikp2, nl))
        e = exp2_poly(fma(nl, nv, tt))
        x = mul(cPL, e)
        bm = fma(bhi, cMU2, mul(blo, cMU2))
        x = fma(x, bm, x)
        x = fma(x, bm, x)
        ok = ref > 1e-30 * ref.max()
        err = x.astype(np.float64)[ok] / ref[ok] - 1
        w = ref[ok] / ref[ok].sum()                          # cells enter Px weighted by their size
        worst_bias = max(worst_bias, abs((err * w).sum()))
        worst_rms = max(worst_rms, np.sqrt((err ** 2 * w).sum()))
    assert worst_rms < 1.5e-7, worst_rms                     # measured 6.7e-8: a few float32 roundings per cell
    assert worst_bias < 6e-9, worst_bias                     # measured 1.7e-9, coherent over the cells of a row

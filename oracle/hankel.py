"""CPU ORACLE (test infrastructure, NOT product code) -- three float64 versions of

    Px(k_par, r) = 1/(2 pi) int_0^kmax dk_perp k_perp J0(k_perp r) P3D(k, mu),

the step the reference delegates to ``forestflow.pcross.Px_Mpc_detailed`` (call site
cupix/likelihood/theory.py:250-264; third-party, un-pinned, absent here; pinned instead to the
reference's in-tree copy of the algorithm, see FFTLogHankel below).
All three have the call signature ``hankel(rt_Mpc[Nr], kp_Mpc[Nk], p3d_func(k, mu)) -> [Nr, Nk]``.

  NodeHankel    "oracle B": the fixed-node rule the device evaluates -- nodes k_q and a callable
                returning per-separation weights W[r, q] are handed in as plain arrays, the sum
                is done here in float64.  Bit-level definition of what the CUDA path must match.
  BruteHankel   independent check of that rule: composite Simpson on 2^20+1 points in ln k_perp.
  FFTLogHankel  "oracle A": the algorithm of the in-tree copy of ForestFlow's Px code
                (cupix/likelihood/lyaP3D.py:272-351): FFTLog (Hamilton 2000; ``hankl.FFTLog(x, f,
                q=0, mu=0)`` at :324-326) of P3D*k_perp on 2^11 log-spaced k_perp in [1e-7, 1e3],
                P1D splice below r ~ 0.02-0.2 Mpc (:291-298, :329-347), cubic spline to the
                requested r (:349-350).  hankl's source is not available, so FFTLog is restated
                from the published algorithm; it equals scipy.fft.fht to 1e-14, and this class
                equals the output of the reference's lyaP3D.py run with scipy.fft.fht in hankl's
                place to 1e-10 (tests/golden/make_golden_hankel.py,
                tests/test_oracle_hankel_golden.py).
"""
import numpy as np
from scipy.integrate import simpson
from scipy.interpolate import CubicSpline
from scipy.special import j0, loggamma

KT_MPC_MAX = 200.0


def _k_mu(kpar, kperp):
    """k, mu on a [Nq, Nk] grid as at lyaP3D.py:279-282; mu := 0 where k == 0."""
    kperp2d = kperp[:, None] * np.ones_like(kpar)[None, :]
    kpar2d = np.ones_like(kperp)[:, None] * kpar[None, :]
    k2d = np.sqrt(kperp2d ** 2 + kpar2d ** 2)
    with np.errstate(divide="ignore", invalid="ignore"):
        mu2d = np.where(k2d > 0, kpar2d / np.where(k2d > 0, k2d, 1.0), 0.0)
    return k2d, mu2d


def _p3d_on_grid(p3d_func, k2d, mu2d, kmax=KT_MPC_MAX):
    """P3D on the grid; 0 where k == 0 (P_L(0) = 0) and above max_k_for_p3d (theory.py:264)."""
    ksafe = np.where(k2d > 0, k2d, 1.0)
    with np.errstate(all="ignore"):
        p = p3d_func(ksafe, mu2d)
    return np.where((k2d > 0) & (k2d <= kmax * (1 + 1e-12)), p, 0.0)


class NodeHankel(object):
    def __init__(self, k_nodes, weight_fn, kpar0=None):
        self.k_nodes = np.asarray(k_nodes, dtype=np.float64)
        self.weight_fn = weight_fn            # r[Nr] -> W[Nr, Nq]
        self.kpar0 = kpar0                    # 1e-4: k_par = 0 replaced as the reference's in-tree path does (lyaP3D.py:35-39)

    def __call__(self, rt_Mpc, kp_Mpc, p3d_func):
        kp = np.atleast_1d(np.asarray(kp_Mpc, dtype=np.float64))
        if self.kpar0 is not None:
            kp = np.where(kp == 0, self.kpar0, kp)
        k2d, mu2d = _k_mu(kp, self.k_nodes)
        # the rule integrates over k_perp <= kmax; P3D is taken as is at every node
        ksafe = np.where(k2d > 0, k2d, 1.0)
        with np.errstate(all="ignore"):
            p = np.where(k2d > 0, p3d_func(ksafe, mu2d), 0.0)
        return self.weight_fn(np.atleast_1d(rt_Mpc)) @ p


class BruteHankel(object):
    def __init__(self, n=2 ** 20 + 1, kmin=1e-6, kmax=KT_MPC_MAX):
        assert n % 2 == 1
        self.t = np.linspace(np.log(kmin), np.log(kmax), n)
        self.k = np.exp(self.t)
        dt = self.t[1] - self.t[0]
        w = np.full(n, dt / 3.0)
        w[1:-1:2] *= 4.0
        w[2:-1:2] *= 2.0
        self.w = w
        self.kmin = kmin

    def __call__(self, rt_Mpc, kp_Mpc, p3d_func):
        r = np.atleast_1d(np.asarray(rt_Mpc, dtype=np.float64))
        kp = np.atleast_1d(np.asarray(kp_Mpc, dtype=np.float64))
        out = np.zeros((r.size, kp.size))
        jw = [self.w * self.k ** 2 * j0(self.k * ri) / (2 * np.pi) for ri in r]
        for ik, kq in enumerate(kp):
            k = np.sqrt(kq * kq + self.k ** 2)
            mu = kq / k
            with np.errstate(all="ignore"):
                f = p3d_func(k, mu)
            # k_perp < kmin: P3D taken constant, J0 = 1
            tail = f[0] * self.kmin ** 2 / 2 / (2 * np.pi)
            for ir in range(r.size):
                out[ir, ik] = np.dot(jw[ir], f) + tail
        return out


def fftlog(x, f_x, q=0.0, mu=0.0, xy=1.0):
    """G(y) = int_0^inf F(x) (xy)^q J_mu(xy) y dx on the reciprocal log grid y = xy / x[::-1]
    (Hamilton 2000, MNRAS 312, 257; the transform ``hankl.FFTLog`` implements, no low-ringing)."""
    x = np.asarray(x, dtype=np.float64)
    N = x.size
    delta = np.log(x[-1] / x[0]) / (N - 1)
    L = N * delta
    y = xy / x[::-1]
    c = np.fft.rfft(f_x) / N
    m = np.arange(c.size)
    eta = 2 * np.pi * m / L
    s = q + 1j * eta
    lnU = s * np.log(2.0) + loggamma((mu + 1 + s) / 2) - loggamma((mu + 1 - s) / 2)
    u = np.exp(lnU - 1j * eta * np.log(x[0] * y[0]))
    if N % 2 == 0:
        u[-1] = u[-1].real
    return y, np.fft.irfft(np.conj(c * u), n=N) * N


class FFTLogHankel(object):
    def __init__(self, min_kperp=1e-7, max_kperp=1e3, nkperp=2 ** 11, max_k_for_p3d=KT_MPC_MAX):
        self.kperps = np.logspace(np.log10(min_kperp), np.log10(max_kperp), nkperp)   # lyaP3D.py:272
        self.max_k = max_k_for_p3d

    def _p1d(self, kp, p3d_func):
        """lyaP3D.py:66-106 with ln k_perp = linspace(ln 1e-3, ln 1e2, 99) (:291-298)."""
        ln_kperp = np.linspace(np.log(0.001), np.log(100), 99)
        k_perp = np.exp(ln_kperp)
        k2d, mu2d = _k_mu(kp, k_perp)
        p = _p3d_on_grid(p3d_func, k2d, mu2d, self.max_k) * (k_perp ** 2 / (2 * np.pi))[:, None]
        return simpson(p, x=ln_kperp, axis=0)

    def __call__(self, rt_Mpc, kp_Mpc, p3d_func):
        r_req = np.atleast_1d(np.asarray(rt_Mpc, dtype=np.float64))
        kp = np.atleast_1d(np.asarray(kp_Mpc, dtype=np.float64)).copy()
        kp[kp == 0] = 1e-4                                           # lyaP3D.py:35-39
        k2d, mu2d = _k_mu(kp, self.kperps)
        P3D = _p3d_on_grid(p3d_func, k2d, mu2d, self.max_k)
        P1D = self._p1d(kp, p3d_func)
        out = np.empty((r_req.size, kp.size))
        for ik in range(kp.size):
            rperp, lhs = fftlog(self.kperps, P3D[:, ik] * self.kperps, q=0, mu=0)    # :322-326
            Px = lhs / rperp / (2 * np.pi)                                           # :327
            Px[rperp < 0.02] = P1D[ik]                                               # :331-333
            idxmin = np.abs(rperp - 0.02).argmin()                                   # :335-336
            idxmax = np.abs(rperp - 0.08).argmin()
            cut = np.arange(idxmin, idxmax)
            Px_keep, r_keep = np.delete(Px, cut), np.delete(rperp, cut)              # :338-339
            lo = np.abs(r_keep - 0.005).argmin()                                     # :340-341
            hi = np.abs(r_keep - 0.2).argmin()
            cs = CubicSpline(r_keep[lo:hi], Px_keep[lo:hi])                          # :342-345
            Px = np.insert(Px_keep, idxmin, cs(rperp[idxmin:idxmax]))                # :346-347
            out[:, ik] = CubicSpline(rperp, Px)(r_req)                               # :349-350
        return out

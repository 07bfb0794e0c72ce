"""Fixed-node product-integration rule for the P3D -> Px Hankel transform.

    Px(k_par, r) = 1/(2 pi) int_0^kmax dk_perp k_perp J0(k_perp r) P3D(k_par, k_perp)

(the integral ``forestflow.pcross.Px_Mpc_detailed`` approximates with FFTLog, called at
cupix/likelihood/theory.py:250-264; ``max_k_for_p3d = kt_Mpc_max = 200``, theory.py:320).

P3D is smooth in k_perp while J0(k_perp r) oscillates up to ~3000 times over the range, so the
oscillation is integrated *exactly into the weights*:  P3D(k_par, .) is represented by its
quintic not-a-knot spline interpolant on ``n_q`` nodes that are uniform in u = ln(k_perp + k0)
(linear near k_perp = 0, logarithmic above k0), and

    W[r, q] = 1/(2 pi) int du (k + k0) k J0(k r) phi_q(u)          (phi_q = cardinal spline of node q)

is evaluated once per (z, binning) in float64 with Gauss-Legendre panels that resolve J0.  Then
``Px(k_par, r) = sum_q W[r, q] P3D(k_par, k_q)`` for every parameter point.  Because the rule is
linear, the reference's theta sub-sampling average (cupix/likelihood/likelihood.py:99-130) is
folded into the weights as well.  Measured error against a 2^20-point brute-force integral is
< 2e-7 of the Px scale at the default n_q = 88 over the ForestFlow training box
(tests/test_quadrature.py; 64 nodes: 4e-7, 96 nodes: 7e-8 -- `n_q` is a Theory config key).

That accuracy holds for a *smooth* linear power.  With baryon acoustic oscillations in P_L (period
2 pi / r_s ~ 0.043 1/Mpc, what CAMB delivers) the log-uniform nodes are too sparse above k ~ 0.1:
88 nodes give 7e-4, 192 nodes 1.4e-5 of the Px scale.  ``refine > 0`` therefore adds node density
where the wiggles live: nodes are uniform in

    u(k) = ln(k + k0) + refine * k_refine * atan(k / k_refine),      du/dk = 1/(k + k0) + refine / (1 + (k/k_refine)^2)

The preset ``PRESETS['bao']`` (112 nodes, refine = 50, k_refine = 0.3) reaches 5e-6 with a 5 %% wiggle
and 1e-7 without (tests/test_quadrature.py), i.e. the 1e-5 tolerance of the north star at 1.27x the
cost of the default rule instead of > 2.2x.
"""
from concurrent.futures import ThreadPoolExecutor
import os

import numpy as np
from scipy.interpolate import BSpline, make_interp_spline
from scipy.special import j0

PRESETS = {"smooth": dict(n_q=88, refine=0.0), "bao": dict(n_q=112, refine=50.0, k_refine=0.3)}
DEFAULT_NQ = 88
DEFAULT_K0 = 0.01
DEFAULT_KMAX = 200.0
SPLINE_ORDER = 5
_GL_N = 16
_RAD_PER_PANEL = 4.0


class KperpRule(object):
    """Nodes of the k_perp quadrature and the machinery to build Hankel weights on them."""

    def __init__(self, n_q=DEFAULT_NQ, k0=DEFAULT_K0, kmax=DEFAULT_KMAX, order=SPLINE_ORDER, refine=0.0, k_refine=0.3):
        assert n_q >= 2 * (order + 1), "too few k_perp nodes for the spline order"
        self.n_q, self.k0, self.kmax, self.order = int(n_q), float(k0), float(kmax), int(order)
        self.refine, self.k_refine = float(refine), float(k_refine)
        if self.refine > 0.0:
            self.u = np.linspace(self.u_of_k(0.0), self.u_of_k(self.kmax), self.n_q)
            k = self.k_of_u(self.u)
        else:
            self.u = np.linspace(np.log(self.k0), np.log(self.kmax + self.k0), self.n_q)
            k = np.exp(self.u) - self.k0
        k[0] = 0.0
        k[-1] = self.kmax
        self.k = k
        spl = make_interp_spline(self.u, np.eye(self.n_q), k=self.order)
        self._knots = spl.t
        self._cardinal = np.ascontiguousarray(spl.c)          # [n_coef, n_q]
        self._fine_cache = {}
        self._weights_cache = {}

    def u_of_k(self, k):
        """Node coordinate: log spacing plus, for refine > 0, extra density below ~k_refine."""
        k = np.asarray(k, dtype=np.float64)
        return np.log(k + self.k0) + self.refine * self.k_refine * np.arctan(k / self.k_refine)

    def k_of_u(self, u):
        """Inverse of the (monotone) map by bisection, vectorised; exact to the last few ulps."""
        u = np.asarray(u, dtype=np.float64)
        lo, hi = np.zeros_like(u), np.full_like(u, self.kmax * (1.0 + 1e-12))
        for _ in range(200):
            mid = 0.5 * (lo + hi)
            below = self.u_of_k(mid) < u
            lo = np.where(below, mid, lo)
            hi = np.where(below, hi, mid)
        return 0.5 * (lo + hi)

    def _fine_grid(self, rmax):
        """Gauss-Legendre points/weights between consecutive nodes, panels never straddling a node,
        <= 4 rad of J0 phase per 16-point panel at separation rmax.  Default map: panels in u (as the
        golden fixtures were made); refined map: panels in k_perp with the design matrix at u(k)."""
        key = float(np.ceil(rmax))
        if key not in self._fine_cache:
            gx, gw = np.polynomial.legendre.leggauss(_GL_N)
            xs, ws = [], []
            for i in range(self.n_q - 1):
                if self.refine > 0.0:
                    a, b = self.k[i], self.k[i + 1]
                    dk = b - a
                else:
                    a, b = self.u[i], self.u[i + 1]
                    dk = np.exp(b) - np.exp(a)
                npan = max(1, int(np.ceil(dk * key / _RAD_PER_PANEL)))
                edges = np.linspace(a, b, npan + 1)
                c = 0.5 * (edges[:-1] + edges[1:])
                hw = 0.5 * (edges[1:] - edges[:-1])
                xs.append((c[:, None] + hw[:, None] * gx[None, :]).ravel())
                ws.append((hw[:, None] * gw[None, :]).ravel())
            xs = np.concatenate(xs)
            ws = np.concatenate(ws)
            if self.refine > 0.0:
                kf = xs
                uf = np.clip(self.u_of_k(kf), self.u[0], self.u[-1])
                meas = ws * kf / (2.0 * np.pi)                        # dk k / 2 pi
            else:
                kf = np.exp(xs) - self.k0
                uf = xs
                meas = ws * kf * (kf + self.k0) / (2.0 * np.pi)       # du (k + k0) k / 2 pi
            design = BSpline.design_matrix(uf, self._knots, self.order).tocsr()   # [n_f, n_coef]
            self._fine_cache[key] = (kf, meas, design)
        return self._fine_cache[key]

    def weights(self, r_Mpc, group_index=None, group_weight=None, n_groups=None, nthreads=None):
        """Hankel weights.

        r_Mpc [n_r]: transverse separations.  Without groups returns W[n_r, n_q].
        With ``group_index[n_r]`` / ``group_weight[n_r]`` returns the folded weights
        W[g, q] = sum_{i in g} w_i W_r[i, q] / sum_{i in g} w_i   (theta-weighted bin average).
        """
        r = np.atleast_1d(np.asarray(r_Mpc, dtype=np.float64))
        if group_index is None:
            gi = np.arange(r.size)
            gw = np.ones(r.size)
            ng = r.size
        else:
            gi = np.asarray(group_index)
            gw = np.asarray(group_weight, dtype=np.float64)
            ng = int(n_groups if n_groups is not None else gi.max() + 1)
        # the weights depend on the separations only: plans rebuilt for another primordial power (same
        # background, theory.py:54-56) or for another Likelihood on the same binning find them here
        ckey = (r.tobytes(), np.asarray(gi, dtype=np.int64).tobytes(), gw.tobytes(), ng)
        if ckey in self._weights_cache:
            return self._weights_cache[ckey].copy()
        kf, meas, design = self._fine_grid(max(r.max(), 1e-3))
        norm = np.bincount(gi, weights=gw, minlength=ng)
        jbar = np.zeros((ng, kf.size))
        # one process per GPU: share the host cores between the ranks of this node
        ranks = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))
        nthreads = nthreads or max(1, min(8, len(os.sched_getaffinity(0)) // ranks))

        def work(g):
            sel = np.where(gi == g)[0]
            acc = np.zeros(kf.size)
            for i in sel:
                acc += gw[i] * j0(kf * r[i])
            jbar[g] = acc * (meas / norm[g])

        if nthreads > 1 and ng > 1:
            with ThreadPoolExecutor(nthreads) as ex:
                list(ex.map(work, range(ng)))
        else:
            for g in range(ng):
                work(g)
        b = design.T.dot(jbar.T)                      # [n_coef, ng]
        W = np.ascontiguousarray((self._cardinal.T @ b).T)          # [ng, n_q]
        if len(self._weights_cache) >= 64:
            self._weights_cache.pop(next(iter(self._weights_cache)))
        self._weights_cache[ckey] = W
        return W.copy()

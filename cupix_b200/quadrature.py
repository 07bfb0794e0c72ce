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
< 2.5e-7 of the Px scale at n_q = 88 over the ForestFlow training box down to theta = 0.5'
(tests/test_quadrature.py; 52 nodes: 2.8e-6, 56 nodes: 1.5e-6, 64 nodes: 8.7e-7 -- `n_q` / `px_rtol` are Theory config keys).

That accuracy holds for a *smooth* linear power.  With baryon acoustic oscillations in P_L (period
2 pi / r_s ~ 0.043 1/Mpc, what CAMB / LaCE deliver; they move Px by up to 4e-3 of its scale at low k_par) the
log-uniform nodes are too sparse above k ~ 0.1: 88 nodes are off by 8e-4.  ``refine > 0`` therefore adds node density
where the wiggles live: nodes are uniform in

    u(k) = ln(k + k0) + refine * k_refine * atan(k / k_refine),      du/dk = 1/(k + k0) + refine / (1 + (k/k_refine)^2)

(the extra density must stay gentle: refine = 50 starves the logarithmic part and costs 4e-5 at the smallest
separations even on a smooth power -- the round-1 preset; refine = 15..25 with k_refine = 0.5..0.7 does not).

``RULES`` lists the node sets on offer with their MEASURED worst-case error against the reference's own algorithm
(oracle A: FFTLog on 2^11 nodes + splines, cupix/likelihood/lyaP3D.py:272-351, itself within 9e-8 of a 2^20-point
brute-force integral for k_par > 0): max |dPx| / max |Px| over three Arinyo parameter sets (central and extreme rows of
the ForestFlow training table, the CoLoRe fit), HCD on, theta = 0.5' .. 58' at z = 2.2, k_par = 0.0077 .. 1.19 1/A;
for the 'bao' family with a 5 % BAO-like wiggle in P_L (tools/quadrature_scan.py, profiles/quadrature_scan_r02.txt).
``choose_rule(kind, px_rtol)`` returns the cheapest entry whose error is at most px_rtol / 3.
"""
from concurrent.futures import ThreadPoolExecutor
import os

import numpy as np
from scipy.interpolate import BSpline, make_interp_spline
from scipy.special import j0

# kind -> [(KperpRule keyword arguments, measured worst-case error)], cheapest first
RULES = {
    "smooth": [(dict(n_q=52), 2.8e-6), (dict(n_q=56), 1.5e-6), (dict(n_q=64), 8.7e-7), (dict(n_q=72), 3.0e-7),
               (dict(n_q=88), 3.0e-7)],
    "bao": [(dict(n_q=112, refine=20.0, k_refine=0.5), 4.8e-6), (dict(n_q=128, refine=20.0, k_refine=0.5), 2.8e-6),
            (dict(n_q=144, refine=20.0, k_refine=0.5), 1.3e-6), (dict(n_q=160, refine=15.0, k_refine=0.7), 5.8e-7),
            (dict(n_q=176, refine=20.0, k_refine=0.5), 3.9e-7)],
}
RULE_MARGIN = 3.0
DEFAULT_PX_RTOL = 1e-5          # BASELINE.json north_star: relative 1e-5 on model Px
# named presets (Theory config 'kperp_rule'): the entries chosen for the default tolerance, and the finest ones
PRESETS = {"smooth": RULES["smooth"][0][0], "bao": RULES["bao"][1][0],
           "smooth_fine": RULES["smooth"][4][0], "bao_fine": RULES["bao"][3][0]}
WIGGLE_THRESHOLD = 1.5e-3       # see wiggle_amplitude
DEFAULT_NQ = 88
DEFAULT_K0 = 0.01
DEFAULT_KMAX = 200.0
SPLINE_ORDER = 5
_GL_N = 16
_RAD_PER_PANEL = 4.0


def choose_rule(kind, px_rtol=DEFAULT_PX_RTOL, margin=RULE_MARGIN):
    """Keyword arguments of the cheapest rule of family ``kind`` ('smooth' | 'bao') whose measured error is at most
    ``px_rtol / margin``; the finest one (with ``met = False``) if none is.  Returns (kwargs, error, met)."""
    for kw, err in RULES[kind]:
        if err * margin <= px_rtol:
            return dict(kw), err, True
    kw, err = min(RULES[kind], key=lambda e: e[1])
    return dict(kw), err, False


def wiggle_amplitude(cosmo, z):
    """Largest deviation of ln P_L(k) from a degree-8 Chebyshev fit in ln k over 0.02 .. 0.6 1/Mpc: 5e-4 for a
    no-wiggle (Eisenstein-Hu) power, ~ the relative BAO amplitude (0.05) for a CAMB-like one.  Decides between the
    'smooth' and 'bao' node families when the Theory config does not (kperp_rule = 'auto')."""
    k = np.geomspace(0.02, 0.6, 600)
    x = np.log(k)
    y = np.log(np.asarray(cosmo.get_linP_Mpc(z, k), dtype=np.float64))
    xs = (x - x.mean()) / np.ptp(x) * 2.0
    c = np.polynomial.chebyshev.chebfit(xs, y, 8)
    return float(np.abs(y - np.polynomial.chebyshev.chebval(xs, c)).max())


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

    def weights(self, r_Mpc, group_index=None, group_weight=None, n_groups=None, nthreads=None, jbar_fn=None):
        """Hankel weights.

        r_Mpc [n_r]: transverse separations.  Without groups returns W[n_r, n_q].
        With ``group_index[n_r]`` / ``group_weight[n_r]`` returns the folded weights
        W[g, q] = sum_{i in g} w_i W_r[i, q] / sum_{i in g} w_i   (theta-weighted bin average).
        ``jbar_fn(group_start, r_sorted, w_sorted, kf, meas) -> jbar[n_groups, n_f]``: the J0 panel sums done elsewhere
        (cupix_b200.engine.hankel_jbar: on the GPU, where the product path builds its plans); default: NumPy threads.
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
        ckey = (r.tobytes(), np.asarray(gi, dtype=np.int64).tobytes(), gw.tobytes(), ng, jbar_fn is not None)
        if ckey in self._weights_cache:
            return self._weights_cache[ckey].copy()
        kf, meas, design = self._fine_grid(max(r.max(), 1e-3))
        if jbar_fn is not None:
            order = np.argsort(gi, kind="stable")
            start = np.concatenate([[0], np.cumsum(np.bincount(gi, minlength=ng))])
            jbar = jbar_fn(start, r[order], gw[order], kf, meas)
            return self._project(jbar, design, ckey)
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
        return self._project(jbar, design, ckey)

    def _project(self, jbar, design, ckey):
        """jbar[g, f] -> W[g, q]: projection onto the cardinal splines of the nodes; result cached."""
        b = design.T.dot(jbar.T)                      # [n_coef, ng]
        W = np.ascontiguousarray((self._cardinal.T @ b).T)          # [ng, n_q]
        if len(self._weights_cache) >= 64:
            self._weights_cache.pop(next(iter(self._weights_cache)))
        self._weights_cache[ckey] = W
        return W.copy()


_RULE_CACHE = {}


def get_rule(n_q=DEFAULT_NQ, k0=DEFAULT_K0, kmax=DEFAULT_KMAX, order=SPLINE_ORDER, refine=0.0, k_refine=0.3):
    """Shared KperpRule instance per parameter set: the Theory objects of all z bins of a measurement then share the
    fine-grid panels, the spline design matrix and the weights cache instead of rebuilding them per z."""
    key = (int(n_q), float(k0), float(kmax), int(order), float(refine), float(k_refine))
    if key not in _RULE_CACHE:
        _RULE_CACHE[key] = KperpRule(*key)
    return _RULE_CACHE[key]

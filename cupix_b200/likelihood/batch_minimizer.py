"""Lock-step minimiser for many fits at once.

The reference fits one mock at a time with iminuit (`cupix/likelihood/minimize_posterior.py:68-79`,
`scripts/fit_many_mocks.py:151-170`: a serial loop that rebuilds data + likelihood per mock, one
likelihood call per MIGRAD function evaluation).  On the GPU the concurrency comes from evaluating
the current trial points of *all* fits in one batched call: every iteration builds a second-order
finite-difference stencil around each fit's current point (1 + 2P + P(P-1)/2 points), gets all
chi2 values with one `get_chi2_batch(..., data_idx=...)`, and takes a damped Newton
(Levenberg-Marquardt) step per fit.  Converged fits keep being evaluated (the batch is cheap) but
stop moving.  Returns the best-fit values, chi2 and the parameter covariance 2 H^-1 (errordef of a
chi2, the same quantity HESSE reports).
"""
import numpy as np


def _stencil(P):
    """Offsets (in units of the step) of the quadratic-model stencil and how to read it back."""
    offs = [np.zeros(P)]
    for i in range(P):
        e = np.zeros(P)
        e[i] = 1.0
        offs += [e, -e]
    pairs = []
    for i in range(P):
        for j in range(i + 1, P):
            e = np.zeros(P)
            e[i] = e[j] = 1.0
            offs.append(e)
            pairs.append((i, j))
    return np.array(offs), pairs


class BatchMinimizer(object):

    def __init__(self, like, names, steps, lower=None, upper=None, priors=None):
        """like: object with get_chi2_batch(points, names, data_idx=...);  steps [P]: finite
        difference steps (about a tenth of the expected 1-sigma error works well);  priors: optional
        list of (mean, width) or None per parameter (Gaussian, added to chi2 as in posterior.py:73-84)."""
        self.like, self.names = like, list(names)
        self.P = len(self.names)
        self.steps = np.asarray(steps, dtype=np.float64)
        self.lower = np.full(self.P, -np.inf) if lower is None else np.asarray(lower, dtype=np.float64)
        self.upper = np.full(self.P, np.inf) if upper is None else np.asarray(upper, dtype=np.float64)
        self.priors = priors or [None] * self.P
        self.offs, self.pairs = _stencil(self.P)
        self.n_calls = 0

    def _chi2(self, pts, data_idx):
        c = self.like.get_chi2_batch(pts, self.names, data_idx=data_idx)
        self.n_calls += 1
        for i, pr in enumerate(self.priors):
            if pr is not None:
                c = c + ((pts[:, i] - pr[0]) / pr[1]) ** 2
        return c

    def _quadratic_model(self, x, data_idx):
        """chi2, gradient and Hessian of every fit from one batched stencil evaluation."""
        N, P, S = x.shape[0], self.P, self.offs.shape[0]
        # no stencil point may leave [lower, upper] (the model can be undefined there, e.g. kp_Mpc <= 0): the stencil is
        # centred one step inside the bounds and the quadratic model is translated back to x below
        x_req = x
        x = np.clip(x, np.minimum(self.lower + self.steps, self.upper), np.maximum(self.upper - self.steps, self.lower))
        pts = (x[:, None, :] + self.offs[None, :, :] * self.steps[None, None, :]).reshape(N * S, P)
        pts = np.clip(pts, self.lower, self.upper)          # only bites when upper - lower < 2 steps
        didx = None if data_idx is None else np.repeat(data_idx, S)
        f = self._chi2(pts, didx).reshape(N, S)
        f0 = f[:, 0]
        h = self.steps
        g = np.empty((N, P))
        H = np.empty((N, P, P))
        for i in range(P):
            fp, fm = f[:, 1 + 2 * i], f[:, 2 + 2 * i]
            g[:, i] = (fp - fm) / (2 * h[i])
            H[:, i, i] = (fp - 2 * f0 + fm) / h[i] ** 2
        for n, (i, j) in enumerate(self.pairs):
            fij = f[:, 1 + 2 * P + n]
            H[:, i, j] = H[:, j, i] = (fij - f[:, 1 + 2 * i] - f[:, 1 + 2 * j] + f0) / (h[i] * h[j])
        d = x_req - x
        if np.any(d != 0.0):
            Hd = np.einsum("nij,nj->ni", H, d)
            f0 = f0 + np.einsum("ni,ni->n", g, d) + 0.5 * np.einsum("ni,ni->n", d, Hd)
            g = g + Hd
        return f0, g, H

    def minimize(self, x0, data_idx=None, max_iter=20, tol=1e-4, lam0=1e-3):
        """x0 [N, P] start points (one row per fit), data_idx [N] data vector of each fit."""
        x = np.array(x0, dtype=np.float64, copy=True)
        N = x.shape[0]
        lam = np.full(N, lam0)
        done = np.zeros(N, dtype=bool)
        f0, g, H = self._quadratic_model(x, data_idx)
        for it in range(max_iter):
            D = np.einsum("nii->ni", H).copy()
            D = np.where(D > 0, D, np.abs(D) + 1e-30)
            A = H + lam[:, None, None] * np.eye(self.P)[None] * D[:, :, None]
            try:
                dx = -np.linalg.solve(A, g[:, :, None])[:, :, 0]
            except np.linalg.LinAlgError:
                dx = -g / D
            xt = np.clip(x + dx, self.lower, self.upper)
            ft, gt, Ht = self._quadratic_model(xt, data_idx)
            better = (ft < f0) & ~done
            edm = 0.5 * np.abs(np.einsum("ni,ni->n", g, dx))         # estimated distance to minimum
            x[better], f0[better], g[better], H[better] = xt[better], ft[better], gt[better], Ht[better]
            lam = np.where(better, np.maximum(lam * 0.3, 1e-9), np.minimum(lam * 10.0, 1e9))
            # converged once the step the quadratic model proposes promises less than `tol` in chi2 -- whether or not the
            # trial point was numerically "better": at the minimum of an exactly polynomial chi2 (the factorised path) it is not
            done |= edm < tol
            if done.all():
                break
        cov = np.full((N, self.P, self.P), np.nan)
        for n in range(N):
            try:
                cov[n] = 2.0 * np.linalg.inv(H[n])
            except np.linalg.LinAlgError:
                pass
        return {"x": x, "chi2": f0, "cov": cov, "converged": done, "n_iter": it + 1, "n_calls": self.n_calls,
                "hessian": H}

"""``Minimizer``: best fit, errors and covariance of a Posterior -- interface mirror of
``cupix/likelihood/minimize_posterior.py`` (and of ``iminuit_minimizer.py`` through ``IminuitMinimizer``).

The reference hands ``minus_log_prob_interface`` to iminuit, which calls the likelihood once per MIGRAD /
HESSE function evaluation (minimize_posterior.py:31-79).  iminuit is not a dependency here and a serial
driver would waste the device, so the same quantities come from the lock-step Levenberg-Marquardt of
``batch_minimizer.BatchMinimizer``: every iteration scores a full second-order stencil around the current
point of every start in one batched call.  ``config['n_starts'] > 1`` runs that many starts at once (the
first at the parameters' ``ini_value``, the others spread over +-``delta``) and keeps the best.

Conventions kept from the reference: start values, step sizes and limits come from the FreeParameter objects
(``ini_value``, ``delta``, ``min_value``, ``max_value``); errors are those of errordef = 0.5 on -log P, i.e.
``cov = 2 (d2 chi2 / dp dp)^-1`` with the Gaussian priors included; the methods minimise on demand.
"""
import os

import numpy as np

from cupix_b200.likelihood.batch_minimizer import BatchMinimizer
from cupix_b200.likelihood.posterior import Posterior


class Minimizer(object):
    """Best-fit finder for the Posterior class"""

    def __init__(self, post, config={}):
        self.verbose = config.get('verbose', False)
        self.post = post
        pars = post.free_params
        self.names = [p.name for p in pars]
        self.ini_values = np.array([p.ini_value for p in pars], dtype=np.float64)
        self.deltas = np.array([p.delta if p.delta else max(abs(p.ini_value), 1.0) * 1e-2 for p in pars], dtype=np.float64)
        lower = [-np.inf if p.min_value is None else p.min_value for p in pars]
        upper = [np.inf if p.max_value is None else p.max_value for p in pars]
        priors = [None if p.gauss_prior_mean is None else (p.gauss_prior_mean, p.gauss_prior_width) for p in pars]
        self.n_starts = int(config.get('n_starts', 1))
        self.max_iter = int(config.get('max_iter', 30))
        self.tol = float(config.get('tol', 1e-4))
        self.rng = np.random.default_rng(config.get('seed', 0))
        # finite-difference step of the stencil: a fraction of the expected error
        self.fitter = BatchMinimizer(post.like, self.names, config.get('step_scale', 0.1) * self.deltas,
                                     lower=lower, upper=upper, priors=priors)
        self.values = None          # best-fit values, in the order of post.free_params
        self.errors = None
        self.covariance = None
        self.valid = False
        self.fmin = None            # -log posterior at the best fit
        if self.verbose:
            print('Free parameters in minimizer')
            print(self.names)
            print('Initial values set to', self.ini_values)

    def silence(self):
        """set verbose=False in all classes"""
        self.verbose = False
        self.post.verbose = False
        self.post.like.verbose = False
        if hasattr(self.post.like, 'theory'):
            self.post.like.theory.verbose = False

    def minus_log_prob_interface(self, values):
        return -1.0 * self.post.get_log_posterior_from_values(values=values)

    def minimize(self, compute_hesse=True):
        """Find the minimum of -log P; the Hessian comes with the last stencil, so ``compute_hesse`` costs nothing."""
        x0 = np.tile(self.ini_values, (self.n_starts, 1))
        if self.n_starts > 1:
            x0[1:] += self.rng.uniform(-1.0, 1.0, (self.n_starts - 1, len(self.names))) * self.deltas
            x0 = np.clip(x0, self.fitter.lower, self.fitter.upper)
        res = self.fitter.minimize(x0, max_iter=self.max_iter, tol=self.tol)
        best = int(np.nanargmin(res["chi2"]))
        self.values = res["x"][best]
        self.covariance = res["cov"][best]
        with np.errstate(invalid="ignore"):
            self.errors = np.sqrt(np.diag(self.covariance))
        self.fmin = 0.5 * res["chi2"][best]
        self.valid = bool(res["converged"][best]) and bool(np.all(np.isfinite(self.errors)))
        self.result = res
        if self.verbose:
            print("best fit", dict(zip(self.names, self.values)), "chi2 + priors =", 2 * self.fmin,
                  "after", res["n_iter"], "iterations,", res["n_calls"], "batched calls")

    def _minimize_if_needed(self, compute_hesse=True):
        if not self.valid:
            self.minimize(compute_hesse=compute_hesse)
        elif self.verbose:
            print('already minimized')

    def get_best_fit_params(self):
        """Run minimizer if needed, return dictionary"""
        self._minimize_if_needed(compute_hesse=False)
        return self.post.get_params_from_values(self.values)

    def get_best_fit_chi2(self, return_info=False):
        """chi2 of the likelihood alone (no priors) at the best fit, as minimize_posterior.py:109-116"""
        return self.post.like.get_chi2(params=self.get_best_fit_params(), return_info=return_info)

    def get_best_fit_probability(self):
        return self.post.like.get_probability(self.get_best_fit_params(), n_free_p=len(self.post.free_params))

    def get_param_index(self, pname):
        return self.post.get_param_index(pname)

    def get_best_fit_value(self, pname, return_hesse=False):
        """Best-fit value of one parameter, optionally with its Gaussian error"""
        self._minimize_if_needed(compute_hesse=return_hesse)
        ipar = self.get_param_index(pname)
        if return_hesse:
            return self.values[ipar], self.errors[ipar]
        return self.values[ipar]

    def get_results_dict(self):
        """Best-fit values, errors, covariance, chi2, probability (keys as minimize_posterior.py:283-301)"""
        results = {}
        for name in self.names:
            results[name], results[name + '_err'] = self.get_best_fit_value(name, return_hesse=True)
        results['cov'] = self.covariance
        results['prob'] = self.get_best_fit_probability()
        results['chi2'] = self.get_best_fit_chi2()
        results['ndata'] = self.post.like.get_ndata()
        results['ndf'] = self.post.get_ndf()
        return results

    def print_results(self):
        results = self.get_results_dict()
        print('chi2 = {:.3f} (ndf = {}) , prob = {:.3f}'.format(results['chi2'], results['ndf'], results['prob']))
        for par in self.post.free_params:
            info = '{} = {:.4f} +/- {:.4f}'.format(par.name, results[par.name], results[par.name + '_err'])
            if par.true_value is not None:
                info += ' (true value = {:.4f})'.format(par.true_value)
            if par.gauss_prior_mean is not None:
                info += ' (prior = {:.4f} +/- {:.4f})'.format(par.gauss_prior_mean, par.gauss_prior_width)
            print(info)

    def save_results(self, outfile=None, outpath=None):
        """One .npz with the results dictionary; ``outpath`` defaults to the working directory (the reference
        writes into its own repository, minimize_posterior.py:304-313)."""
        savepath = os.path.join(outpath or os.getcwd(), outfile or "minimizer_results.npz")
        save_analysis_npz(self.get_results_dict(), filename=savepath)
        return savepath


class IminuitMinimizer(Minimizer):
    """Constructor signature of ``cupix/likelihood/iminuit_minimizer.py:10`` (likelihood + free parameters)."""

    def __init__(self, like, free_params, verbose=False):
        super().__init__(Posterior(like, free_params, {'verbose': verbose}), {'verbose': verbose})

    def get_params_dict_from_values(self, values):
        return self.post.get_params_from_values(values)


def save_analysis_npz(results, filename="analysis_results.npz"):
    """results: dict, or list of per-analysis dicts (stored as 'analysis-i')"""
    if isinstance(results, list):
        out = {'analysis-%d' % i: r for i, r in enumerate(results)}
    else:
        out = {str(k): r for k, r in results.items()}
    np.savez(filename, **out, allow_pickle=True)

"""Posterior = likelihood + bounds + Gaussian priors (mirror of cupix/likelihood/posterior.py),
with a vectorised entry point for ensemble samplers / multi-start minimisers.

One deliberate difference: out-of-bounds points return -inf.  The reference returns ``--np.inf``
(= +inf, posterior.py:45), which makes a maximiser run away; that is a typo, not behaviour to keep.
"""
import numpy as np


class Posterior(object):

    def __init__(self, like, free_params, config={}):
        self.verbose = config.get('verbose', False)
        self.like = like
        self.free_params = free_params

    def silence(self):
        self.verbose = False
        self.like.verbose = False
        self.like.theory.verbose = False

    def get_ndf(self):
        return self.like.get_ndata() - len(self.free_params)

    def get_param_index(self, param_name):
        for ip, par in enumerate(self.free_params):
            if par.name == param_name:
                return ip
        raise KeyError(param_name)

    def get_params_from_values(self, values):
        assert len(values) == len(self.free_params), "Inconsistent number of free parameters"
        return {p.name: values[ip] for ip, p in enumerate(self.free_params)}

    def get_values_from_params(self, params):
        assert len(params) == len(self.free_params), "Inconsistent number of free parameters"
        return np.array([params[p.name] for p in self.free_params])

    def get_log_prior(self, params={}):
        chi2_prior = 0.0
        for name, value in params.items():
            chi2_prior += self.free_params[self.get_param_index(name)].get_prior_chi2(value)
        return -0.5 * chi2_prior

    def get_log_posterior(self, params={}):
        for name, value in params.items():
            if self.free_params[self.get_param_index(name)].out_of_bounds(value):
                return -np.inf
        return self.like.get_log_like(params=params) + self.get_log_prior(params=params)

    def get_log_posterior_from_values(self, values):
        return self.get_log_posterior(self.get_params_from_values(values))

    def get_log_posterior_batch(self, values):
        """values [B, Np] -> log posterior [B]; one device call for all in-bounds rows
        (emcee ``vectorize=True`` signature)."""
        values = np.atleast_2d(np.asarray(values, dtype=np.float64))
        names = [p.name for p in self.free_params]
        lo = np.array([-np.inf if p.min_value is None else p.min_value for p in self.free_params])
        hi = np.array([np.inf if p.max_value is None else p.max_value for p in self.free_params])
        ok = np.all((values >= lo) & (values <= hi), axis=1)
        out = np.full(values.shape[0], -np.inf)
        if ok.any():
            logp = self.like.get_log_like_batch(values[ok], names)
            for ip, p in enumerate(self.free_params):
                if p.gauss_prior_mean is not None:
                    logp = logp - 0.5 * ((values[ok, ip] - p.gauss_prior_mean) / p.gauss_prior_width) ** 2
            out[ok] = logp
        return out

"""Ensemble MCMC on the batched posterior.

The reference wraps emcee (`cupix/likelihood/sampler.py:39-43`: ``EnsembleSampler(nwalkers, Np,
post.get_log_posterior_from_values)``), which calls the likelihood once per walker per step.  emcee is
not a dependency here; this is the same affine-invariant stretch move (Goodman & Weare 2010) with the
two half-ensembles updated in turn, so that every step is two batched device calls of nwalkers/2
points each -- the ``vectorize=True`` calling convention, with config keys as in the reference
(`nwalkers, max_nsteps, nburnin`, sampler.py:31-35).
"""
import numpy as np


class Sampler(object):

    def __init__(self, post, config={}):
        self.post = post
        self.verbose = config.get('verbose', False)
        self.ndim = len(post.free_params)
        self.nwalkers = config.get('nwalkers', max(4 * self.ndim, 16))
        assert self.nwalkers % 2 == 0 and self.nwalkers >= 2 * self.ndim + 2
        self.max_nsteps = config.get('max_nsteps', 1000)
        self.nburnin = config.get('nburnin', 200)
        self.stretch_a = config.get('stretch_a', 2.0)
        self.rng = np.random.default_rng(config.get('seed', 0))
        self.chain = None
        self.lnprob = None
        self.acceptance_fraction = None

    def get_initial_walkers(self, scale=1e-2):
        """Gaussian ball around the parameters' ini_value, width ``scale`` x (max - min), inside bounds."""
        fp = self.post.free_params
        ini = np.array([p.ini_value for p in fp], dtype=np.float64)
        width = np.array([scale * (p.max_value - p.min_value) for p in fp])
        lo = np.array([p.min_value for p in fp])
        hi = np.array([p.max_value for p in fp])
        return np.clip(ini + width * self.rng.normal(size=(self.nwalkers, self.ndim)), lo, hi)

    def run_sampler(self, p0=None, nsteps=None):
        nsteps = self.max_nsteps if nsteps is None else nsteps
        x = self.get_initial_walkers() if p0 is None else np.array(p0, dtype=np.float64)
        lp = self.post.get_log_posterior_batch(x)
        if not np.all(np.isfinite(lp)):
            raise ValueError("initial walkers have zero posterior probability")
        half = self.nwalkers // 2
        chain = np.empty((nsteps, self.nwalkers, self.ndim))
        lnprob = np.empty((nsteps, self.nwalkers))
        n_acc = 0
        for it in range(nsteps):
            for first in (0, 1):
                mine = slice(0, half) if first == 0 else slice(half, None)
                other = slice(half, None) if first == 0 else slice(0, half)
                zz = ((self.stretch_a - 1.0) * self.rng.random(half) + 1.0) ** 2 / self.stretch_a
                partner = x[other][self.rng.integers(half, size=half)]
                prop = partner + zz[:, None] * (x[mine] - partner)
                lp_prop = self.post.get_log_posterior_batch(prop)
                log_ratio = (self.ndim - 1.0) * np.log(zz) + lp_prop - lp[mine]
                accept = np.log(self.rng.random(half)) < log_ratio
                xm, lm = x[mine], lp[mine]
                xm[accept], lm[accept] = prop[accept], lp_prop[accept]
                x[mine], lp[mine] = xm, lm
                n_acc += int(accept.sum())
            chain[it], lnprob[it] = x, lp
        self.chain, self.lnprob = chain, lnprob
        self.acceptance_fraction = n_acc / float(nsteps * self.nwalkers)
        return chain[self.nburnin:].reshape(-1, self.ndim) if nsteps > self.nburnin else chain.reshape(-1, self.ndim)

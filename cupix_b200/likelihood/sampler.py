"""Ensemble MCMC on the batched posterior.

The reference wraps emcee (`cupix/likelihood/sampler.py:28-43`: ``EnsembleSampler(nwalkers, Np,
post.get_log_posterior_from_values)``), which calls the likelihood once per walker per step.  emcee is
not a dependency here; this is the same affine-invariant stretch move (Goodman & Weare 2010) with the
two half-ensembles updated in turn, so that every step is two batched device calls of nwalkers/2
points each -- the ``vectorize=True`` calling convention.  Interface as in the reference: config keys
``nwalkers, max_nsteps, nburnin`` (sampler.py:31-35), ``setup_emcee_sampler``, ``silence``,
``get_initial_walkers`` (ini_value + U(0, 1) delta, pulled back inside the bounds, :55-82), ``run_sampler`` (burn-in +
max_nsteps iterations, :85-107), and ``emcee_sampler`` with the accessors notebooks use on emcee's object
(``get_chain``, ``get_log_prob``, ``iteration``, ``nwalkers``, ``acceptance_fraction``).  For amplitude-only fits every
batched call is one kernel launch (the factorised path keeps the Gram data of the tuple, include/cupix_b200.h).
"""
import numpy as np


class Sampler(object):

    def __init__(self, post, config={}):
        self.post = post
        self.verbose = config.get('verbose', False)
        self.setup_emcee_sampler(config)
        if self.verbose:
            print('Free parameters in sampler')
            print([par.name for par in self.post.free_params])
            print('Initial values set to', [par.ini_value for par in self.post.free_params])

    def setup_emcee_sampler(self, config):
        """sampler.py:28-43.  The stretch move on half-ensembles needs an even number of walkers, at least 2 Np + 2:
        the reference's default of 10 is raised when there are more than four free parameters."""
        self.ndim = len(self.post.free_params)
        self.nwalkers = config.get('nwalkers', max(10, 2 * self.ndim + 2))
        assert self.nwalkers % 2 == 0 and self.nwalkers >= 2 * self.ndim + 2, "nwalkers must be even and >= 2 Np + 2"
        self.max_nsteps = config.get('max_nsteps', 1000)
        self.nburnin = config.get('nburnin', 100)
        assert self.nburnin < self.max_nsteps, 'nburnin >= max_nsteps'
        self.parallel = config.get('parallel', False)       # accepted for compatibility: every step is one batch anyway
        self.stretch_a = config.get('stretch_a', 2.0)
        self.rng = np.random.default_rng(config.get('seed', None))
        self.chain = None
        self.lnprob = None
        self.iteration = 0
        self.acceptance_fraction = None
        self.emcee_sampler = self

    def silence(self):
        """set verbose=False in all classes"""
        self.verbose = False
        self.post.verbose = False
        self.post.like.verbose = False
        self.post.like.theory.verbose = False

    def get_initial_walkers(self, scale=1e-2):
        """Starting points (sampler.py:55-82): ``ini_value + U(0, 1) * delta`` per parameter, pulled back to
        ``min + 0.01 delta`` / ``max - 0.01 delta`` when that leaves the bounds.  A parameter without ``delta`` gets a
        Gaussian ball of width ``scale`` x (max - min) instead."""
        out = np.empty((self.nwalkers, self.ndim))
        for ip, par in enumerate(self.post.free_params):
            if par.delta is not None:
                val = par.ini_value + self.rng.random(self.nwalkers) * par.delta
                val[val < par.min_value] = par.min_value + 0.01 * par.delta
                val[val > par.max_value] = par.max_value - 0.01 * par.delta
            else:
                width = scale * (par.max_value - par.min_value)
                val = np.clip(par.ini_value + width * self.rng.normal(size=self.nwalkers), par.min_value, par.max_value)
            out[:, ip] = val
        return out

    # -- the part of emcee's EnsembleSampler interface that callers of ``sampler.emcee_sampler`` use ----------------
    def get_chain(self, discard=0, flat=False, thin=1):
        c = self.chain[discard::thin]
        return c.reshape(-1, self.ndim) if flat else c

    def get_log_prob(self, discard=0, flat=False, thin=1):
        lp = self.lnprob[discard::thin]
        return lp.reshape(-1) if flat else lp

    def run_sampler(self, p0=None, nsteps=None):
        """Burn-in plus ``max_nsteps`` iterations from ``get_initial_walkers()`` (sampler.py:85-107); ``p0`` / ``nsteps``
        override the start and the total.  Returns the flattened chain after burn-in (the reference returns nothing:
        its caller reads ``emcee_sampler.get_chain``, which works here too)."""
        nsteps = self.nburnin + self.max_nsteps if nsteps is None else nsteps
        x = self.get_initial_walkers() if p0 is None else np.array(p0, dtype=np.float64)
        if self.verbose:
            print('starting points of walkers')
            print(x)
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
            self.iteration = it + 1
            if self.verbose and (it + 1) % 10 == 0:
                print("Step %d out of %d " % (it + 1, nsteps))
        self.chain, self.lnprob = chain, lnprob
        self.acceptance_fraction = n_acc / float(nsteps * self.nwalkers)
        return chain[self.nburnin:].reshape(-1, self.ndim) if nsteps > self.nburnin else chain.reshape(-1, self.ndim)

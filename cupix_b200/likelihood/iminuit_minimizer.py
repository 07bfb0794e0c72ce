"""``IminuitMinimizer(like, free_params)``: the module path of ``cupix/likelihood/iminuit_minimizer.py``.

The class lives in ``minimize_posterior`` (it is the ``Minimizer`` on a ``Posterior`` built from the likelihood and the
free parameters, as in the reference, where the two files are near-copies of each other); this module keeps the
reference's import path working.
"""
from cupix_b200.likelihood.minimize_posterior import IminuitMinimizer, save_analysis_npz   # noqa: F401

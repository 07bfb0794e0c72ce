"""``convolve_window`` / ``rebin_theta`` for ad-hoc use on arrays (``cupix/likelihood/window_and_rebin.py:59-76``).

The reference's ``Likelihood.get_convolved_px`` calls these two per (z, theta) inside Python loops.  The likelihood here
does not: window, rebinning, whitening and chi2 are one device stage (``k_window_tc`` / ``k_window`` on the operator
``T_chi`` that ``cupix_b200.plan`` folds them into).  The functions are kept because the reference's scripts import
them by name (``scripts/fit_many_mocks.py:12``, ``scripts/validate_pipeline.py:12``) and notebooks apply them to a
single array; nothing under ``cupix_b200`` calls them.  ``tests/test_gpu_parity.py`` checks the device stage against
them on the device's own Px_Zam.
"""
import numpy as np


def convolve_window(window_matrix, model):
    """``window_matrix [M, n] @ model [n]`` (or ``[n, ...]``): a rectangular window also rebins in k."""
    return np.matmul(window_matrix, model)


def rebin_theta(V_aM_in_Theta, P_aM_in_Theta):
    """V-weighted mean over the fine theta bins ``a`` of one coarse bin (top-hat binning): ``[n_a_in_A, M] -> [M]``."""
    V_aM_in_Theta = np.asarray(V_aM_in_Theta)
    return np.sum(V_aM_in_Theta * P_aM_in_Theta, axis=0) / np.sum(V_aM_in_Theta, axis=0)

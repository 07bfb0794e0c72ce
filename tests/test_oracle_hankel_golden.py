"""The P3D -> Px Hankel step against the reference's own in-tree implementation of it.

tests/golden/ref_hankel_lyaP3D.npz was produced by *running* cupix/likelihood/lyaP3D.py
(``Px_Mpc_withSiIII`` :127-362 with the SiIII terms switched off, ``P1D_Mpc`` :66-106 -- the code the
reference calls "copied from ForestFlow") with ``hankl.FFTLog`` supplied by ``scipy.fft.fht``
(tests/golden/make_golden_hankel.py).  Three links are checked here:

  1. oracle.hankel.fftlog (restated from Hamilton 2000)  ==  scipy.fft.fht          (1e-12)
  2. oracle.hankel.FFTLogHankel ("oracle A", restatement of the reference algorithm)
                                                          ==  the reference's output (1e-10)
  3. the fixed-node rule the device evaluates ("oracle B") vs the reference's output:
     <= 1e-5 of the Px scale for every separation >= 0.3 Mpc (the smallest DESI theta bin is ~0.8 Mpc;
     below ~0.2 Mpc the reference splices in P1D(k_par) and is itself off the exact integral by up to
     1e-2, while the node rule follows the brute-force integral to 2e-7, tests/test_quadrature.py).
"""
import numpy as np
import pytest
import scipy.fft

from cupix_b200.cosmology import Cosmology
from cupix_b200.quadrature import KperpRule
from oracle.hankel import FFTLogHankel, NodeHankel, fftlog
from tests import helpers as H

G = H.golden("ref_hankel_lyaP3D.npz")
COSMO = Cosmology()


def p3d_for(par):
    """Arinyo P3D of theory.py:285-308 as a callable of (k, mu)."""
    z, bias, beta, q1, q2, kv, av, bv, kp = par

    def p3d(k, mu):
        linP = COSMO.get_linP_Mpc(z, k)
        d2 = k ** 3 * linP / (2 * np.pi ** 2)
        vel = (k / kv) ** av * np.abs(mu) ** bv
        return bias ** 2 * (1 + beta * mu ** 2) ** 2 * linP * np.exp(d2 * (q1 + q2 * d2) * (1 - vel) - (k / kp) ** 2)
    return p3d


def test_fftlog_restatement_equals_scipy_fht():
    x = np.logspace(-7, 3, 2 ** 11)
    dln = np.log(x[-1] / x[0]) / (x.size - 1)
    for f in (x ** 1.5 * np.exp(-x * x / 50.0), x / (1 + x ** 2) ** 1.5, np.exp(-np.log(x / 3.0) ** 2)):
        y, g = fftlog(x, f, q=0, mu=0)
        ref = scipy.fft.fht(f, dln, mu=0, offset=0.0, bias=0.0)
        np.testing.assert_allclose(y, 1.0 / x[::-1], rtol=1e-14)
        assert np.abs(g - ref).max() <= 1e-12 * np.abs(ref).max()


@pytest.mark.parametrize("name", [str(n) for n in G["names"]])
def test_oracle_A_equals_reference_lyaP3D(name):
    # the in-tree copy has no max_k_for_p3d cut (theory.py:264 passes one to the newer ForestFlow function)
    got = FFTLogHankel(max_k_for_p3d=np.inf)(G["rperp_Mpc"], G["kpar_Mpc"], p3d_for(G["params|" + name]))
    ref = G["px|" + name]
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() <= 1e-10 * np.abs(ref).max()


@pytest.mark.parametrize("n_q", [88, 96])
@pytest.mark.parametrize("name", ["gadget_z2.4", "colore_like", "strong_nl"])
def test_node_rule_vs_reference_lyaP3D(name, n_q):
    rule = KperpRule(n_q)
    node = NodeHankel(rule.k, lambda rr: rule.weights(rr, nthreads=1))
    r = G["rperp_Mpc"]
    got = node(r, G["kpar_Mpc"], p3d_for(G["params|" + name]))
    ref = G["px|" + name]
    sel = r >= 0.3
    assert np.abs(got - ref)[sel].max() <= 1e-5 * np.abs(ref).max()

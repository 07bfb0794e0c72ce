"""GPU: set-up work done on the device -- the J0 panel sums of the Hankel weights (cpx_hankel_jbar) against the NumPy
implementation the oracle uses, and the wall clock of building the 12 plans of the S5 stress shape."""
import time

import numpy as np
import pytest

from cupix_b200 import plan as _plan
from cupix_b200 import quadrature as Q
from cupix_b200.engine import hankel_jbar
from cupix_b200.likelihood.likelihood import Likelihood
from cupix_b200.px_data import synthetic
from cupix_b200.px_data.data_DESI_DR2 import DESI_DR2
from tests import helpers as H

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kw", [dict(n_q=56), dict(n_q=128, refine=20.0, k_refine=0.5)])
def test_device_hankel_weights_equal_host_weights(kw):
    rule = Q.KperpRule(**kw)
    th, grp = _plan.fine_theta_grid(np.geomspace(0.5, 50.0, 7), np.geomspace(0.6, 60.0, 7), 25)
    r = th / 0.62
    host = rule.weights(r, grp, th, 7)
    dev = rule.weights(r, grp, th, 7, jbar_fn=lambda *a: hankel_jbar(0, *a))
    assert dev.shape == host.shape == (7, rule.n_q)
    # CUDA's j0 is good to ~5e-12 absolute at large arguments (k r up to 2e4 here), scipy's to ~1e-16: the weights agree
    # to 1e-11 of their scale (averaging over a bin's separations brings it below 1e-12), six decades inside the budget
    assert np.abs(dev - host).max() <= 1e-11 * np.abs(host).max()
    one = rule.weights(r[:3], jbar_fn=lambda *a: hankel_jbar(0, *a))          # ungrouped: one separation per row
    assert np.abs(one - rule.weights(r[:3])).max() <= 1e-11 * np.abs(host).max()


def test_S5_plans_build_fast_and_match_host_plans():
    d = synthetic.make_measurement_dict("S5")
    data = DESI_DR2(d)
    likes = [Likelihood(data, H.make_theory(float(z), "best_fit_arinyo_from_gadget", {"include_hcd": True}, n_q=56), iz,
                        config={"N_theta_average": 100}) for iz, z in enumerate(d["z_centers"])]
    likes[0].build_plan()                                                     # library load, CUDA context, fine-grid panels
    Q.get_rule(56)._weights_cache.clear()
    t0 = time.time()
    plans = [lk.build_plan() for lk in likes]
    dt = time.time() - t0
    print("12 S5 plans (N_theta_average = 100, 2000 separations each): %.2f s" % dt)
    assert dt < 3.0, "building the S5 plans took %.1f s" % dt
    host = _plan.build_likelihood_plan(likes[5].theory, likes[5].theory.fid_cosmo, data, 5, 100, likes[5].theory.rule)
    assert np.abs(plans[5].hankel_w - host.hankel_w).max() <= 1e-12 * np.abs(host.hankel_w).max()
    p, q = {"bias": -0.17, "beta": 1.4}, likes[5].build_plan()
    np.testing.assert_allclose(H.emulate_plan_px(q, p), H.emulate_plan_px(host, p), rtol=0, atol=1e-11 * np.abs(H.emulate_plan_px(host, p)).max())

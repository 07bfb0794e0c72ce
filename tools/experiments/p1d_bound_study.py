"""Does P1D(k_par) = Hankel at r_perp = 0 bound |Px_Zam| of a point, and by how many bits is the bound loose?"""
import numpy as np, sys
sys.path.insert(0, "/root/repo")
from tests import helpers as H
from cupix_b200 import plan as _plan
from cupix_b200.px_data import synthetic
from cupix_b200.px_data.data_DESI_DR2 import DESI_DR2
from cupix_b200.likelihood.likelihood import Likelihood

flags = {"include_hcd": True, "include_metal": True, "include_sky": True, "include_continuum": True}
d = synthetic.make_measurement_dict("S1", N_A=20, N_m=156)
res = []
for model in ("best_fit_arinyo_from_gadget", "best_fit_arinyo_from_colore"):
  for fl in ({"include_hcd": True}, flags):
    th = H.make_theory(float(d["z_centers"][0]), model, fl, n_q=88)
    like = Likelihood(DESI_DR2(d), th, 0, config={"N_theta_average": 2})
    p = like.build_plan()
    # a copy of the plan with one extra theta column at r = 0 (weights of the r_perp = 0 Hankel = P1D)
    w0 = th.rule.weights(np.array([0.0]))            # [1, n_q]
    import copy
    p0 = copy.copy(p)
    p0.hankel_w = np.vstack([p.hankel_w, w0]); p0.n_a = p.n_a + 1
    if p.metal_auto is not None:
        p0.metal_auto = np.concatenate([p.metal_auto, np.zeros((3, 1, p.n_m))], axis=1)
        p0.metal_cross = np.concatenate([p.metal_cross, np.zeros((3, 1, p.n_m))], axis=1)
    if p.sky_xi_a is not None:
        p0.sky_xi_a = np.concatenate([p.sky_xi_a, [0.0]])
    rng = np.random.default_rng(0)
    names = ["bias", "beta", "q1", "q2", "kv_Mpc", "av", "bv", "kp_Mpc", "b_H", "b_X", "b_noise_Mpc"]
    base = np.array([p.defaults[_plan.SLOT_INDEX[n]] for n in names])
    worst_lo, worst_hi = 1e9, 0
    for i in range(60):
        par = dict(zip(names, base * rng.uniform(0.5, 1.5, len(names))))
        flags0 = p0.flags
        px = H.emulate_plan_px(p0, par)              # [n_a + 1, n_m], incl. additive terms and continuum
        full = np.abs(px[:-1]).max()
        # Lya-only P1D column: evaluate with additive terms off
        p0.flags = flags0 & 1
        p1d = np.abs(H.emulate_plan_px(p0, par)[-1]).max()
        p0.flags = flags0
        ratio = p1d / full
        worst_lo, worst_hi = min(worst_lo, ratio), max(worst_hi, ratio)
    print("%-30s flags %2d  P1D_lya_max / max|Px_Zam| in [%.3f, %.3f]  -> bits lost <= %.1f, bound violated: %s"
          % (model, p.flags, worst_lo, worst_hi, np.log2(worst_hi), worst_lo < 1.0))

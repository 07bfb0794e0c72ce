"""Compare the tensor-core Hankel path (k_p3d_hankel_tc, default) with the FP64 DMMA path (CPX_TC=0)
and with the float64 NumPy emulation of the plan on the same points: Px_Zam and chi2."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as G   # noqa: E402

G.build()
import bench   # noqa: E402
from cupix_b200 import plan as _plan   # noqa: E402
from cupix_b200.engine import PxEngine, slots_from_names   # noqa: E402
from tests import helpers as H   # noqa: E402


def main(shape="S2", n_q=88, B=300):
    d, data, likes, _ = bench.build_workload(shape, 100, n_q, need_engine=False)
    plans = [lk.build_plan() for lk in likes]
    names = ["bias", "beta", "q1", "kp_Mpc", "av", "b_H"]
    rng = np.random.default_rng(5)
    nz = len(plans)
    p0 = plans[0]
    base = np.array([p0.defaults[_plan.SLOT_INDEX[n]] for n in names])
    pts = base[None, :] * rng.uniform(0.8, 1.2, size=(B, len(names)))
    slots, zsel = slots_from_names(names)
    out = {}
    for tc in ("0", "1"):
        os.environ["CPX_TC"] = tc
        eng = PxEngine(plans)
        t0 = time.time()
        out[tc] = eng.eval(pts, slots, zsel=zsel, want=("chi2", "px_zam"), stage_ms=True)
        print("CPX_TC=%s: %.3f s, stage_ms %s" % (tc, time.time() - t0, out[tc]["stage_ms"]), flush=True)
        eng.close()
    a, b = out["0"]["px_zam"], out["1"]["px_zam"]
    scale = np.abs(a).max(axis=(2, 3), keepdims=True)
    print("Px_Zam  tensor vs DMMA: max |d| / max|Px| per (pt, z) = %.3e" % (np.abs(a - b) / scale).max())
    sc_a = np.abs(a).max(axis=3, keepdims=True)
    print("Px_Zam  tensor vs DMMA: max |d| / max_m|Px| per (pt, z, theta) = %.3e" % (np.abs(a - b) / sc_a).max())
    print("chi2    tensor vs DMMA: max |d| = %.3e at chi2 up to %.3e" % (np.abs(out["0"]["chi2"] - out["1"]["chi2"]).max(), out["0"]["chi2"].max()))
    # float64 emulation of the plan for a few points
    for i in (0, B // 2, B - 1):
        prm = dict(zip(names, pts[i]))
        chi_ref = 0.0
        for iz, p in enumerate(plans):
            px = H.emulate_plan_px(p, prm)
            for tag in ("0", "1"):
                got = out[tag]["px_zam"][i, iz, :p.n_a, :p.n_m]
                err = np.abs(got - px).max() / np.abs(px).max()
                if iz == 0:
                    print("  point %d z0 CPX_TC=%s: max |dPx| / max|Px| vs float64 emulation = %.3e" % (i, tag, err))
            chi_ref += H.emulate_plan_chi2(p, prm)[1].sum()
        print("  point %d chi2: emulation %.6f  DMMA %.6f (d = %.2e)  tensor %.6f (d = %.2e)" % (
            i, chi_ref, out["0"]["chi2"][i], out["0"]["chi2"][i] - chi_ref, out["1"]["chi2"][i], out["1"]["chi2"][i] - chi_ref))


if __name__ == "__main__":
    main(*(sys.argv[1:2] or ["S2"]), *(int(x) for x in sys.argv[2:]))

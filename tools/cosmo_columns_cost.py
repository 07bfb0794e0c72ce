"""Cost of the As / ns batch columns on the S5 bench workload: the same 65 536-point batches with and without
the two cosmology columns (COSMO instantiation of k_p3d_hankel), stage times from cpx_eval.

    python tools/cosmo_columns_cost.py  > profiles/cosmo_columns_cost_r01.txt
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from cupix_b200 import plan as _plan  # noqa: E402
from cupix_b200.engine import slots_from_names  # noqa: E402


def main():
    B = 65536
    if os.environ.get("CPX_LIB_VARIANT"):          # time another build of the library (kernel variants, tools/bin/)
        import cupix_b200.engine as E
        E.LIB_PATH = os.path.abspath(os.environ["CPX_LIB_VARIANT"])
    filled, data, likes, joint = bench.build_workload("S5", 100, 88)
    eng = joint.engine()
    d = _plan.default_vector(likes[0].theory, likes[0].theory.fid_cosmo)
    rng = np.random.default_rng(1)
    base = np.stack([d[0] + rng.uniform(-0.02, 0.02, B), d[1] + rng.uniform(-0.2, 0.2, B)], axis=1)
    cosmo = np.stack([d[16] * rng.uniform(0.9, 1.1, B), d[17] + rng.uniform(-0.03, 0.03, B)], axis=1)
    for label, names, pts in (("bias, beta", ["bias", "beta"], base),
                              ("bias, beta, As, ns", ["bias", "beta", "As", "ns"], np.hstack([base, cosmo]))):
        slots, zsel = slots_from_names(names)
        best = None
        for rep in range(4):
            out = eng.eval(points=pts, slots=slots, zsel=zsel, want=("chi2",), stage_ms=True)
            ms = out["stage_ms"]
            if rep and (best is None or ms.sum() < best.sum()):
                best = ms
        print("%-22s params %.2f  p3d_hankel %.2f  window_chi2 %.2f  reduce %.2f  total %.2f ms per %d points x %d z"
              % (label, best[0], best[1], best[2], best[3], best.sum(), B, len(likes)))


if __name__ == "__main__":
    main()

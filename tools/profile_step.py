"""Short profiling target for ncu: the S5 bench workload, a few general-points steps (k_params, k_rowpars, k_p3d_hankel,
k_window_tc, k_reduce) and two factorised whole-grid scans (k_basis_f64, k_window, k_gram, k_quad).  Same engine, same
shapes and the same calls as bench.py, without its CPU legs."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench                                                     # noqa: E402


def main(points=1 << 16, steps=2):
    import torch
    filled, data, likes, joint = bench.build_workload("S5", 100, 0)
    eng = joint.engine()
    import cupix_b200.plan as _plan
    from cupix_b200.engine import slots_from_names
    defaults = _plan.default_vector(likes[0].theory, likes[0].theory.fid_cosmo)
    lo, hi = bench.grid_bounds(defaults)
    slots, zsel = slots_from_names(bench.NAMES)
    dev = torch.device("cuda", 0)
    out = torch.empty(int(np.prod(bench.GRID_N)), dtype=torch.float64, device=dev)
    for i in range(steps):
        eng.eval_device(points, 4, slots, grid=(lo, hi, bench.GRID_N), grid_first=i * points, zsel=zsel, chi2_ptr=out.data_ptr(), factor=False)
    perm = [2, 3, 0, 1]
    gnames = [bench.NAMES[j] for j in perm]
    gslots, gzsel = slots_from_names(gnames)
    for i in range(2):
        eng.eval_device(out.numel(), 4, gslots, grid=([lo[j] for j in perm], [hi[j] for j in perm], [bench.GRID_N[j] for j in perm]),
                        zsel=gzsel, chi2_ptr=out.data_ptr())
    torch.cuda.synchronize()
    print("profile_step: ok, chi2[0] = %.6f, launches = %d" % (float(out[0]), eng.info("kernel_launches")))


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""4-parameter 64^4 forecast chi2 scan sharded over the GPUs of one box (BASELINE config 3).

Counterpart of the reference's scripts/chi2_scan_forecast.py (multiprocessing.Pool over 16^npar
points, :116-125).  Each rank evaluates a contiguous slice of the flat grid in chunks; only the
per-point chi2 is gathered.

    torchrun --nproc-per-node 8 scripts/chi2_scan_forecast.py bias beta q1 kp_Mpc --grid 64
"""
import argparse
import time

import numpy as np

from _common import dist_setup, load_measurement
from cupix_b200.parallel import scan_chi2_sharded
from cupix_b200.utils.utils import print0

ap = argparse.ArgumentParser()
ap.add_argument("pars", nargs="+")
ap.add_argument("--file", default=None)
ap.add_argument("--iz", type=int, default=3)
ap.add_argument("--grid", type=int, default=64)
ap.add_argument("--out", default=None)
a = ap.parse_args()
rank, world, local = dist_setup()
data, theories, likes = load_measurement(a.file, "S2", {"default_lya_model": "best_fit_arinyo_from_gadget"}, 100, local,
                                         add_noise=False, kM_max_cut_AA=1.5, km_max_cut_AA=1.8)
like = likes[a.iz]
dflt = like.theory.lya_model.default_lya_params
lo = [dflt[p] - 0.1 * abs(dflt[p]) for p in a.pars]
hi = [dflt[p] + 0.1 * abs(dflt[p]) for p in a.pars]
n = [a.grid] * len(a.pars)
t0 = time.time()
chi2 = scan_chi2_sharded(like, a.pars, lo, hi, n, rank=rank, world_size=world)
dt = time.time() - t0
print0("%d points in %.2f s: %.3g evals/s on %d GPU(s); min chi2 %.4g" % (chi2.size, dt, chi2.size / dt, world, chi2.min()))
if rank == 0 and a.out:
    np.savez(a.out, chi2=chi2.reshape(n), lo=lo, hi=hi, pars=a.pars)
if world > 1:
    import torch.distributed as dist
    dist.destroy_process_group()

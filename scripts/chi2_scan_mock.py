#!/usr/bin/env python
"""chi2 scan over a grid of parameters on one mock Px measurement (one z bin per scan).

B200 counterpart of the reference's scripts/chi2_scan_mock.py: the 128^npar meshgrid, the MPI block
partition and the serial per-rank loop (:107-151) become one device call per rank with the points
generated on the GPU from (lo, hi, n, first index); ranks are GPUs (torchrun), the gather is NCCL.

    python scripts/chi2_scan_mock.py arinyo tight 0 bias beta
    torchrun --nproc-per-node 8 scripts/chi2_scan_mock.py arinyo tight all bias beta --grid 512
"""
import argparse
import time

import numpy as np

from _common import dist_setup, load_measurement
from cupix_b200.parallel import scan_chi2_sharded
from cupix_b200.utils.utils import print0

ap = argparse.ArgumentParser()
ap.add_argument("param_mode", choices=["arinyo"], help="igm mode needs the ForestFlow emulator")
ap.add_argument("param_range", choices=["tight", "broad"])
ap.add_argument("iz_choice", help="redshift bin index or 'all'")
ap.add_argument("pars", nargs="+")
ap.add_argument("--file", default=None, help="DESI_DR2-format .hdf5/.npz measurement (default: synthetic S2 mock)")
ap.add_argument("--grid", type=int, default=128)
ap.add_argument("--out", default=None)
a = ap.parse_args()

rank, world, local = dist_setup()
data, theories, likes = load_measurement(a.file, "S2", {"default_lya_model": "best_fit_arinyo_from_colore"}, 100, local,
                                         kM_max_cut_AA=1, km_max_cut_AA=1.2)
izs = range(len(data.z)) if a.iz_choice.lower() == "all" else [int(a.iz_choice)]
half = {"bias": 0.02, "beta": 0.2, "q1": 0.15, "q2": 0.15, "kp_Mpc": 0.1, "av": 0.1, "bv": 0.1, "kv_Mpc": 0.1}
for iz in izs:
    like = likes[iz]
    dflt = like.theory.lya_model.default_lya_params
    lo = [dflt[p] - half.get(p, 0.1 * abs(dflt[p])) for p in a.pars]
    hi = [dflt[p] + half.get(p, 0.1 * abs(dflt[p])) for p in a.pars]
    n = [a.grid] * len(a.pars)
    t0 = time.time()
    chi2 = scan_chi2_sharded(like, a.pars, lo, hi, n, rank=rank, world_size=world).reshape(n)
    dt = time.time() - t0
    print0("z = %.2f: %d points in %.3f s (%.3g evals/s on %d GPU); min chi2 = %.3f at index %s"
           % (data.z[iz], chi2.size, dt, chi2.size / dt, world, chi2.min(), np.unravel_index(chi2.argmin(), chi2.shape)))
    if rank == 0 and a.out:
        np.savez("%s_z%.1f.npz" % (a.out, data.z[iz]), chi2=chi2, lo=lo, hi=hi, pars=a.pars)
if world > 1:
    import torch.distributed as dist
    dist.destroy_process_group()

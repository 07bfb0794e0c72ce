"""Shared set-up of the example scripts: a measurement (file or seeded synthetic), one Theory and
Likelihood per redshift bin, optional torch.distributed initialisation (one process per GPU)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from cupix_b200.cosmology import Cosmology                      # noqa: E402
from cupix_b200.likelihood.likelihood import JointLikelihood, Likelihood   # noqa: E402
from cupix_b200.likelihood.theory import Theory                 # noqa: E402
from cupix_b200.px_data import synthetic                        # noqa: E402
from cupix_b200.px_data.data_DESI_DR2 import DESI_DR2            # noqa: E402


def dist_setup():
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist.barrier()            # builds the NCCL communicator now, not inside the first timed gather
    return rank, world, local


def load_measurement(path, shape, theory_cfg, N_theta_average, device, truth=None, add_noise=True, **cuts):
    """DESI_DR2 file if ``path`` is given, else a synthetic measurement of the named shape whose data
    vector is the convolved theory at ``truth`` (+ noise) -- the forecast recipe of
    notebooks/examples/generate_forecast_data.py."""
    cosmo = Cosmology()
    if path:
        data = DESI_DR2(path, **cuts)
        theories = [Theory(float(z), fid_cosmo=cosmo, config=dict(theory_cfg, device=device)) for z in data.z]
        likes = [Likelihood(data, th, iz, config={"N_theta_average": N_theta_average, "device": device})
                 for iz, th in enumerate(theories)]
        return data, theories, likes
    d = synthetic.make_measurement_dict(shape)
    theories = [Theory(float(z), fid_cosmo=cosmo, config=dict(theory_cfg, device=device)) for z in d["z_centers"]]
    data = DESI_DR2(d)
    likes = [Likelihood(data, th, iz, config={"N_theta_average": N_theta_average, "device": device})
             for iz, th in enumerate(theories)]
    model = np.array([lk.get_convolved_px(truth or {}) for lk in likes])
    data = DESI_DR2(synthetic.fill_from_model(d, model, add_noise=add_noise))
    likes = [Likelihood(data, th, iz, config={"N_theta_average": N_theta_average, "device": device})
             for iz, th in enumerate(theories)]
    return data, theories, likes

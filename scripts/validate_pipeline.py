#!/usr/bin/env python
"""Full theta x k_par Px with window matrices over all z bins on one GPU (BASELINE config 2):
forecast self-consistency (chi2 of the truth against its own noiseless forecast ~ 0), response to a
parameter shift, and scalar == batched == per-z results.  The tolerance check against the float64
oracle is tests/test_gpu_configs.py::test_S2_validate_pipeline_all_z_vs_oracle.

    python scripts/validate_pipeline.py
"""
import numpy as np

from _common import dist_setup, load_measurement
from cupix_b200.likelihood.likelihood import JointLikelihood

rank, world, local = dist_setup()
cfg = {"default_lya_model": "best_fit_arinyo_from_colore", "include_hcd": True, "include_metal": True,
       "include_sky": True, "include_continuum": True}
data, theories, likes = load_measurement(None, "S2", cfg, 100, local, add_noise=False)
joint = JointLikelihood(likes, device=local)
print("z bins", data.z, " ndata", joint.get_ndata())
c0 = joint.get_chi2({})
print("chi2(truth vs own forecast) = %.3e" % c0)
assert abs(c0) < 1e-6
for iz, lk in enumerate(likes):
    px = lk.get_convolved_px({})
    assert px.shape == data.Px_ZAM[iz].shape and np.allclose(px, data.Px_ZAM[iz], rtol=1e-12, atol=0)
    shifted = lk.get_chi2({"bias": lk.theory.lya_model.default_lya_params["bias"] * 1.02})
    batch = lk.get_chi2_batch(np.array([[lk.theory.lya_model.default_lya_params["bias"] * 1.02]]), ["bias"])[0]
    print("  z = %.1f  chi2(bias +2%%) = %10.3f   scalar == batched: %s   P(chi2) = %.3g"
          % (data.z[iz], shifted, shifted == batch, lk.get_probability({}, n_free_p=0)))
print("validate_pipeline: OK")

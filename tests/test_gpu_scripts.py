"""GPU: the four BASELINE.json workload scripts run end to end as subprocesses (the way a user launches them), and the
sharded scan over two ranks on NCCL when the box has two GPUs."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCRIPTS = os.path.join(ROOT, "scripts")


def _run(args, timeout=600, env=None):
    e = dict(os.environ)
    e.update(env or {})
    r = subprocess.run(args, cwd=ROOT, capture_output=True, text=True, timeout=timeout, env=e)
    assert r.returncode == 0, "%s failed:\n%s\n%s" % (" ".join(args), r.stdout[-2000:], r.stderr[-4000:])
    return r.stdout


def test_validate_pipeline_script():
    out = _run([sys.executable, os.path.join(SCRIPTS, "validate_pipeline.py")])
    assert "validate_pipeline: OK" in out


def test_chi2_scan_mock_script(tmp_path):
    out = _run([sys.executable, os.path.join(SCRIPTS, "chi2_scan_mock.py"), "arinyo", "tight", "0", "bias", "beta",
                "--grid", "64", "--out", str(tmp_path / "scan")])
    assert "4096 points" in out
    f = np.load(tmp_path / "scan_z2.2.npz")
    chi2 = f["chi2"]
    assert chi2.shape == (64, 64) and np.isfinite(chi2).all()
    i, j = np.unravel_index(chi2.argmin(), chi2.shape)
    assert 8 < i < 56 and 8 < j < 56, "the minimum of the scan sits near the truth at the grid centre"


def test_chi2_scan_forecast_script(tmp_path):
    out = _run([sys.executable, os.path.join(SCRIPTS, "chi2_scan_forecast.py"), "bias", "beta", "q1", "kp_Mpc",
                "--grid", "12", "--out", str(tmp_path / "fc.npz")])
    assert "20736 points" in out
    chi2 = np.load(tmp_path / "fc.npz")["chi2"]
    assert chi2.shape == (12, 12, 12, 12) and np.isfinite(chi2).all()
    assert chi2.min() < 0.02 * np.median(chi2)                            # noiseless forecast: a sharp minimum near the truth
    assert all(3 <= i <= 8 for i in np.unravel_index(chi2.argmin(), chi2.shape)[:2])


def test_fit_many_mocks_script(tmp_path):
    out = _run([sys.executable, os.path.join(SCRIPTS, "fit_many_mocks.py"), "bias", "beta", "--nmocks", "128",
                "--out", str(tmp_path / "fits.npz")])
    assert "128 mocks fitted" in out
    f = np.load(tmp_path / "fits.npz")
    pull = (f["x"] - f["truth"]) / np.sqrt(np.einsum("nii->ni", f["cov"]))
    assert f["converged"].all() and np.all(np.abs(pull.mean(axis=0)) < 0.4) and np.all(np.abs(pull.std(axis=0) - 1) < 0.3)


def test_sharded_scan_two_ranks_nccl(tmp_path):
    """torchrun, two ranks, NCCL: the gathered scan equals the single-process scan bit for bit (chi2 stays on the device
    from cpx_eval through the all_gather)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    one = tmp_path / "one"
    two = tmp_path / "two"
    base = [os.path.join(SCRIPTS, "chi2_scan_forecast.py"), "bias", "beta", "q1", "kp_Mpc", "--grid", "10"]
    _run([sys.executable] + base + ["--out", str(one) + ".npz"])
    _run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
          "--master-port", "29611"] + base + ["--out", str(two) + ".npz"], timeout=900)
    np.testing.assert_array_equal(np.load(str(one) + ".npz")["chi2"], np.load(str(two) + ".npz")["chi2"])

"""CPU, world_size 2 over gloo: the N>1 path of the sharded chi2 scan (block partition, slice by
flat grid index, all_gather of the per-point chi2) with a NumPy evaluator standing in for the GPU."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from cupix_b200.parallel import scan_chi2_sharded
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi, n = [-0.135, 1.3], [-0.095, 1.7], [7, 9]
    grid = np.stack(np.meshgrid(np.linspace(lo[0], hi[0], n[0]), np.linspace(lo[1], hi[1], n[1]), indexing="ij"), -1).reshape(-1, 2)

    def evaluate(first, count):
        pts = grid[first:first + count]
        return 1e4 * (pts[:, 0] + 0.115) ** 2 + 30 * (pts[:, 1] - 1.5) ** 2 + rank * 0.0
    full = scan_chi2_sharded(None, ["bias", "beta"], lo, hi, n, rank=rank, world_size=world, evaluate=evaluate)
    np.save(os.path.join(out_dir, "r%d.npy" % rank), full)
    dist.destroy_process_group()


def test_sharded_scan_world_size_2(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a, b = np.load(tmp_path / "r0.npy"), np.load(tmp_path / "r1.npy")
    np.testing.assert_array_equal(a, b)
    lo, hi, n = [-0.135, 1.3], [-0.095, 1.7], [7, 9]
    grid = np.stack(np.meshgrid(np.linspace(lo[0], hi[0], n[0]), np.linspace(lo[1], hi[1], n[1]), indexing="ij"), -1).reshape(-1, 2)
    ref = 1e4 * (grid[:, 0] + 0.115) ** 2 + 30 * (grid[:, 1] - 1.5) ** 2
    np.testing.assert_allclose(a, ref, rtol=0, atol=0)


class _FakeLike(object):
    """Stands for a Likelihood: chi2 on a slice of the meshgrid in whatever axis order it is asked for."""

    @staticmethod
    def chi2_of(names, pts):
        c = dict(bias=(-0.115, 1e4), beta=(1.5, 30.0), q1=(0.4, 7.0), kp_Mpc=(12.0, 0.3))
        return sum(c[nm][1] * (pts[:, j] - c[nm][0]) ** 2 * (1 + 0.1 * j) for j, nm in enumerate(sorted(names))
                   for j in [list(names).index(nm)]) + np.prod(pts, axis=1)

    def get_chi2_grid(self, names, lo, hi, n, first=0, count=None):
        idx = np.arange(first, first + count)
        sub = np.stack(np.unravel_index(idx, n), axis=1)
        pts = np.stack([np.linspace(lo[j], hi[j], n[j])[sub[:, j]] for j in range(len(n))], axis=1)
        order = [list(names).index(nm) for nm in sorted(names)]
        return self.chi2_of(sorted(names), pts[:, order])


def _worker_perm(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from cupix_b200.parallel import scan_chi2_sharded
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    names, lo, hi, n = ["bias", "q1", "beta", "kp_Mpc"], [-0.13, 0.2, 1.3, 9.0], [-0.10, 0.6, 1.7, 15.0], [5, 3, 4, 2]
    full = scan_chi2_sharded(_FakeLike(), names, lo, hi, n, rank=rank, world_size=world)
    np.save(os.path.join(out_dir, "p%d.npy" % rank), full)
    dist.destroy_process_group()


def test_sharded_scan_reorders_tuple_axes_first_and_restores_the_order(tmp_path):
    """The scan flattens the grid with the tuple parameters (q1, kp_Mpc) slowest, so that a rank's slice holds whole
    tuples for the factorised path, and hands the result back in the caller's axis order."""
    sys.path.insert(0, ROOT)
    from cupix_b200.parallel import scan_chi2_sharded, tuple_axes_first
    assert tuple_axes_first(["bias", "q1", "beta", "kp_Mpc"]) == [1, 3, 0, 2]
    assert tuple_axes_first(["bias_0", "beta_1", "L_H_Mpc", "b_X"]) == [2, 0, 1, 3]
    port = _free_port()
    mp.spawn(_worker_perm, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a, b = np.load(tmp_path / "p0.npy"), np.load(tmp_path / "p1.npy")
    np.testing.assert_array_equal(a, b)
    names, lo, hi, n = ["bias", "q1", "beta", "kp_Mpc"], [-0.13, 0.2, 1.3, 9.0], [-0.10, 0.6, 1.7, 15.0], [5, 3, 4, 2]
    ref = _FakeLike().get_chi2_grid(names, lo, hi, n, first=0, count=int(np.prod(n)))
    np.testing.assert_allclose(a, ref, rtol=1e-14)
    one = scan_chi2_sharded(_FakeLike(), names, lo, hi, n)
    np.testing.assert_allclose(one, ref, rtol=1e-14)

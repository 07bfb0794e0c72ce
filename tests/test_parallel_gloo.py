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

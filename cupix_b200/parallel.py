"""Sharding of parameter points over the GPUs of one box (one process per GPU).

Replaces the MPI block partition + Send/Recv + Gatherv of scripts/chi2_scan_mock.py:118-151 and the
multiprocessing pool of scripts/chi2_scan_forecast.py:124-125.  Points are independent, so there is
no data-path collective: every rank holds a replica of the (small) measurement products and
evaluates a contiguous slice of the flattened grid, generating its points on the device from
(lo, hi, n, first index).  The only exchange is the gather of the per-point chi2 (NCCL on GPUs,
gloo in the CPU tests)."""
import numpy as np


def block_partition(N, size):
    """counts/starts exactly as chi2_scan_mock.py:122-123 (first N % size ranks get one extra)."""
    counts = [(N // size) + (1 if i < N % size else 0) for i in range(size)]
    starts = [sum(counts[:i]) for i in range(size)]
    return counts, starts


def scan_chi2_sharded(like, names, lo, hi, n, rank=0, world_size=1, gather=True, group=None,
                      evaluate=None):
    """chi2 over the meshgrid(linspace(lo, hi, n)) sharded over ``world_size`` ranks.

    Returns the full flattened chi2 array on every rank (``gather=True``; all_gather of the per-rank
    slices) or this rank's slice.  ``evaluate(first, count)`` defaults to
    ``like.get_chi2_grid(names, lo, hi, n, first, count)`` (the device path)."""
    N = int(np.prod(np.asarray(n, dtype=np.int64)))
    counts, starts = block_partition(N, world_size)
    if evaluate is None:
        def evaluate(first, count):
            return like.get_chi2_grid(names, lo, hi, n, first=first, count=count)
    mine = np.asarray(evaluate(starts[rank], counts[rank]), dtype=np.float64) if counts[rank] else np.empty(0)
    if world_size == 1 or not gather:
        return mine
    import torch
    import torch.distributed as dist
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    pad = max(counts)
    send = torch.zeros(pad, dtype=torch.float64, device=dev)
    send[:counts[rank]] = torch.from_numpy(mine).to(dev)
    recv = torch.empty(pad * world_size, dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(recv, send, group=group)
    recv = recv.cpu().numpy().reshape(world_size, pad)
    return np.concatenate([recv[r, :counts[r]] for r in range(world_size)])

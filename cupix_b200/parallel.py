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


def tuple_axes_first(names):
    """Permutation of the grid axes that puts the tuple parameters (everything the factorised path of cpx_eval cannot
    take out of the Hankel stage: q1, q2, kv, av, bv, kp, L_H, kC, pC, As, ns) before the amplitude parameters.  In that
    flattening a rank's contiguous slice holds whole tuples, so each tuple's Hankel + window stage runs on one rank only."""
    from cupix_b200 import plan as _plan
    base = [nm if nm in _plan.SLOT_INDEX else nm.rpartition("_")[0] for nm in names]
    amp = [b in _plan.AMP_SLOT_NAMES for b in base]
    return [j for j in range(len(names)) if not amp[j]] + [j for j in range(len(names)) if amp[j]]


def scan_chi2_sharded(like, names, lo, hi, n, rank=0, world_size=1, gather=True, group=None,
                      evaluate=None, as_tensor=False):
    """chi2 over the meshgrid(linspace(lo, hi, n)) sharded over ``world_size`` ranks.

    Returns the full chi2 array, flattened in the order of ``names`` (meshgrid indexing='ij'), on every rank
    (``gather=True``; all_gather of the per-rank slices) or this rank's slice of the internal flattening.
    ``evaluate(first, count)`` defaults to the device path.  Internally the axes are reordered with the tuple parameters
    first (``tuple_axes_first``) and the result is transposed back.  On the NCCL backend the per-rank chi2 never leaves
    the GPU before the gather: cpx_eval writes into the send buffer (device pointers, the current torch stream), the
    all_gather and the transposition run on the device, and one copy brings the full array to the host at the end
    (``as_tensor=True`` returns the CUDA tensor instead)."""
    n = [int(x) for x in n]
    N = int(np.prod(np.asarray(n, dtype=np.int64)))
    counts, starts = block_partition(N, world_size)
    perm = list(range(len(names))) if evaluate is not None else tuple_axes_first(names)
    pnames, plo, phi, pn = [names[j] for j in perm], [lo[j] for j in perm], [hi[j] for j in perm], [n[j] for j in perm]
    inv = np.argsort(perm)

    def unpermute(flat):                      # numpy array or torch tensor, flattened in the permuted order
        return flat.reshape(pn).transpose(*inv).reshape(-1) if isinstance(flat, np.ndarray) \
            else flat.reshape(pn).permute(*[int(i) for i in inv]).reshape(-1)

    nccl = False
    if world_size > 1 and gather:
        import torch
        import torch.distributed as dist
        nccl = dist.get_backend(group) == "nccl"
    if nccl and evaluate is None:
        from cupix_b200.engine import slots_from_names
        dev = torch.device("cuda", torch.cuda.current_device())
        pad = max(counts)
        send = torch.zeros(pad, dtype=torch.float64, device=dev)
        slots, zsel = slots_from_names(pnames)
        if counts[rank]:
            like.engine().eval_device(counts[rank], len(pnames), slots, grid=(plo, phi, pn), grid_first=starts[rank], zsel=zsel,
                                      chi2_ptr=send.data_ptr(), stream=torch.cuda.current_stream(dev).cuda_stream)
        recv = torch.empty(pad * world_size, dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(recv, send, group=group)
        full = recv if all(c == pad for c in counts) else torch.cat([recv[r * pad:r * pad + counts[r]] for r in range(world_size)])
        full = unpermute(full)
        return full if as_tensor else full.cpu().numpy()
    if evaluate is None:
        def evaluate(first, count):
            return like.get_chi2_grid(pnames, plo, phi, pn, first=first, count=count)
    mine = np.asarray(evaluate(starts[rank], counts[rank]), dtype=np.float64) if counts[rank] else np.empty(0)
    if world_size == 1:
        return unpermute(mine)
    if not gather:
        return mine
    dev = torch.device("cpu")
    pad = max(counts)
    send = torch.zeros(pad, dtype=torch.float64, device=dev)
    send[:counts[rank]] = torch.from_numpy(mine).to(dev)
    recv = torch.empty(pad * world_size, dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(recv, send, group=group)
    recv = recv.cpu().numpy().reshape(world_size, pad)
    return unpermute(np.concatenate([recv[r, :counts[r]] for r in range(world_size)]))

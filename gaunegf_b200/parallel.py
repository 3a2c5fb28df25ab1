"""Energy-grid sharding across the GPUs of one box: one process per GPU
(torchrun), cyclic partition of the energies of every call, per-rank on-device partial sums and ONE
all-reduce (NCCL over NVLink) of the N x N partial density matrix per GrInt / GrLessInt call;
per-energy scalars (T(E), DOS) are all-gathered.  Nothing here is a data-path collective inside a
kernel: energies are independent, so the path shards with weak scaling.

torch is used for what it is here for: device buffers, streams and torch.distributed.
"""
import os

import numpy as np


def dist_info():
    """(rank, world) of an initialised torch.distributed group, else (0, 1)."""
    if os.environ.get("GAUNEGF_B200_SHARD", "1") == "0":
        return 0, 1
    try:
        import torch.distributed as dist
    except Exception:       # torch absent: single process
        return 0, 1
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_indices(M, rank, world):
    """cyclic (round-robin) partition: slow-converging Sigma(E) energies spread evenly"""
    return np.arange(rank, M, world)


def sharded_matrix_sum(N, Elist, weights, partial_fn, device=None):
    """sum_k w_k f(E_k) with the energies split over the ranks.

    partial_fn(E_local, w_local, out) must return the rank's partial N x N complex128 sum: as a numpy
    array when `out` is None (CPU / gloo testing of this logic), or written into `out`, a
    torch.complex128 CUDA tensor, when a device is given (the GPU path hands out.data_ptr() to the
    C ABI, so the partial sum never leaves HBM before the all-reduce)."""
    rank, world = dist_info()
    Elist = np.asarray(Elist)
    weights = np.asarray(weights)
    idx = shard_indices(Elist.size, rank, world)
    if world == 1:
        return partial_fn(Elist, weights, None)
    import torch
    import torch.distributed as dist
    if device is not None:
        out = torch.zeros((N, N), dtype=torch.complex128, device=device)
        partial_fn(Elist[idx], weights[idx], out)
        flat = torch.view_as_real(out)
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        return out.cpu().numpy()
    part = np.asarray(partial_fn(Elist[idx], weights[idx], None), dtype=np.complex128)
    t = torch.from_numpy(np.ascontiguousarray(part).view(np.float64).copy())
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.numpy().view(np.complex128).reshape(N, N)


def sharded_per_energy(Elist, per_energy_fn, width=None):
    """per-energy results (T(E): (M,), DOS rows: (M, width)) with the energies split over the ranks;
    every rank returns the full array (all-gather of the cyclic slices)."""
    rank, world = dist_info()
    Elist = np.asarray(Elist)
    M = Elist.size
    if world == 1:
        return per_energy_fn(Elist)
    import torch
    import torch.distributed as dist
    idx = shard_indices(M, rank, world)
    local = np.asarray(per_energy_fn(Elist[idx]), dtype=np.float64)
    shape = (M,) if width is None else (M, width)
    full = np.zeros(shape, dtype=np.float64)
    full[idx] = local
    use_cuda = dist.get_backend() == "nccl"
    t = torch.from_numpy(full)
    if use_cuda:
        t = t.cuda()
    dist.all_reduce(t, op=dist.ReduceOp.SUM)     # slices are disjoint: the sum is the gather
    return t.cpu().numpy()

"""Energy-grid sharding across the GPUs of one box: one process per GPU
(torchrun), cyclic partition of the energies of every call, per-rank on-device partial sums and ONE
all-reduce (NCCL over NVLink) of the N x N partial density matrix per GrInt / GrLessInt call;
per-energy scalars (T(E), DOS) are all-gathered.  Nothing here is a data-path collective inside a
kernel: energies are independent, so the path shards with weak scaling.

torch is used for what it is here for: device buffers, streams and torch.distributed.
"""
import os
import time

import numpy as np

# wall-clock / device-time breakdown of the last sharded_matrix_sum call on this rank (bench.py, profiles/):
# partial_ms (device work of the rank's shard incl. its host orchestration), allreduce_ms, d2h_ms
last_breakdown = {}

_buffers = {}          # (N, device index) -> (device accumulator, pinned host result)


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


def rank0_value(fn):
    """fn() evaluated on rank 0 and broadcast, so that every rank continues from the same value (checkpoint contents)"""
    rank, world = dist_info()
    if world == 1:
        return fn()
    import torch.distributed as dist
    box = [fn() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    return box[0]


def set_system(ctx, F, S):
    """ctx.set_system(F, S) for the sharded drivers, where every rank is handed the same F and S: rank 0 compares them
    with its resident copy element by element, the other ranks only check the size and a strided sample, and the ranks
    agree on one flag (all-reduce MAX of one int).  Without this every rank makes its own full pass over F, S and
    the pinned shadows — 8 x 270 MB of host-memory traffic per call at N = 2048 on a box whose 8 ranks share the memory
    controllers (8-GPU bench: cfg 5 115 ms per call against 75 ms on one GPU).  GAUNEGF_B200_RANK0_COMPARE=0: every rank
    for itself."""
    rank, world = dist_info()
    if world == 1 or os.environ.get("GAUNEGF_B200_RANK0_COMPARE", "1") == "0":
        return ctx.set_system(F, S)
    import torch
    import torch.distributed as dist
    bits = ctx.system_differs(F, S, full=(rank == 0))      # bit 0: F differs, bit 1: S differs
    dev = ("cuda:%d" % ctx.device) if dist.get_backend() == "nccl" else "cpu"
    flag = torch.tensor([bits & 1, (bits >> 1) & 1], dtype=torch.int32, device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MAX)
    f, s = (int(v) for v in flag.tolist())
    # what changed is copied and uploaded without another comparison; an unchanged matrix is not read at all
    return ctx.set_system_known(F, S, f | (s << 1))


def _device_buffers(N, device):
    """persistent N x N device accumulator + pinned host landing buffer per (N, GPU): no allocation, no pageable copy
    in the per-call path"""
    import torch
    key = (N, device.index)
    if key not in _buffers:
        _buffers.clear()                       # one size at a time: 2 x 16 N^2 bytes
        _buffers[key] = (torch.empty((N, N), dtype=torch.complex128, device=device),
                         torch.empty((N, N), dtype=torch.complex128).pin_memory())
    return _buffers[key]


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
        out, host = _device_buffers(N, device)
        t0 = time.perf_counter()
        partial_fn(Elist[idx], weights[idx], out)          # an empty shard writes zeros (gnb_gr_int with M = 0)
        t1 = time.perf_counter()
        dist.all_reduce(torch.view_as_real(out), op=dist.ReduceOp.SUM)
        torch.cuda.synchronize(device)
        t2 = time.perf_counter()
        host.copy_(out, non_blocking=True)
        torch.cuda.synchronize(device)
        t3 = time.perf_counter()
        last_breakdown.update(partial_ms=1e3 * (t1 - t0), allreduce_ms=1e3 * (t2 - t1), d2h_ms=1e3 * (t3 - t2),
                              bytes=16 * N * N, world=world)
        return host.numpy().copy()
    part = np.asarray(partial_fn(Elist[idx], weights[idx], None), dtype=np.complex128)
    t = torch.from_numpy(np.ascontiguousarray(part).view(np.float64).copy())
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.numpy().view(np.complex128).reshape(N, N)


def sharded_matrix_sums(N, Elist, weights, seg_end, partial_fn, device=None):
    """Several weighted sums in one batch: segment s = energies [seg_end[s-1], seg_end[s]).  Every segment is split
    cyclically over the ranks (each rank's local list stays ordered segment by segment) and ONE all-reduce carries
    the (nseg, N, N) partial sums.  partial_fn(E_local, w_local, local_seg_end, out) -> (nseg, N, N) or writes `out`."""
    rank, world = dist_info()
    Elist = np.asarray(Elist)
    weights = np.asarray(weights)
    seg_end = np.asarray(seg_end, dtype=np.int64)
    nseg = seg_end.size
    if world == 1:
        return partial_fn(Elist, weights, seg_end, None)
    idx, ends, lo = [], [], 0
    for hi in seg_end:
        idx.append(lo + shard_indices(int(hi - lo), rank, world))
        ends.append((ends[-1] if ends else 0) + idx[-1].size)
        lo = int(hi)
    idx = np.concatenate(idx) if idx else np.zeros(0, dtype=np.int64)
    ends = np.asarray(ends, dtype=np.int64)
    import torch
    import torch.distributed as dist
    if device is not None:
        key = ("seg", nseg, N, device.index)
        if key not in _buffers:
            for k in [k for k in _buffers if k[0] == "seg"]:
                del _buffers[k]
            _buffers[key] = (torch.empty((nseg, N, N), dtype=torch.complex128, device=device),
                             torch.empty((nseg, N, N), dtype=torch.complex128).pin_memory())
        out, host = _buffers[key]
        partial_fn(Elist[idx], weights[idx], ends, out)
        dist.all_reduce(torch.view_as_real(out), op=dist.ReduceOp.SUM)
        host.copy_(out, non_blocking=True)
        torch.cuda.synchronize(device)
        return host.numpy().copy()
    part = np.asarray(partial_fn(Elist[idx], weights[idx], ends, None), dtype=np.complex128)
    t = torch.from_numpy(np.ascontiguousarray(part).view(np.float64).copy())
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.numpy().view(np.complex128).reshape(nseg, N, N)


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

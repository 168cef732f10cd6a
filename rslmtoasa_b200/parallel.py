"""Multi-GPU plumbing of the path: one process per GPU, units sharded with the reference's own block rule, no halo.

The reference parallelises the recursion over independent units (recursion sites `irec`, pair vectors `4*njij`,
random KPM vectors) with `get_mpi_variables` (mpi.f90:32-58) and sums results afterwards with
MPI_ALLREDUCE(SUM, MPI_IN_PLACE) (bands.f90:270-275).  Here the same two operations run over
`torch.distributed` (NCCL over NVLink on GPUs, gloo in the CPU tests): there is no per-step communication, so the
only collectives are one all-reduce (stochastic-trace moments) or one all-gather (site-resolved results) per
recursion call.
"""
from __future__ import annotations

import numpy as np

from .synthetic import partition


def shard_range(n_units: int, rank: int, world: int) -> tuple[int, int]:
    """0-based half-open [lo, hi) of this rank's units (get_mpi_variables' start_atom..end_atom)."""
    s, e = partition(rank, world, n_units)
    return s - 1, e


def allreduce_sum(arr: np.ndarray, device=None) -> np.ndarray:
    """sum of a complex128/float64 host array over all ranks (MPI_ALLREDUCE(SUM) of the reference)."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return arr
    t = torch.from_numpy(np.ascontiguousarray(arr).view(np.float64).copy())
    if device is not None:
        t = t.to(device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    out = t.cpu().numpy().view(arr.dtype).reshape(arr.shape)
    return np.asfortranarray(out) if np.isfortran(arr) else out


def allgather_units(local: np.ndarray, n_units: int, device=None) -> np.ndarray:
    """concatenate per-unit results (last axis = local unit index) of all ranks in global unit order
    (the commented-out MPI_Allgather of a_b/b2_b, recursion.f90:1788-1799)."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    full = np.zeros(local.shape[:-1] + (n_units,), dtype=local.dtype, order="F")
    lo, hi = shard_range(n_units, rank, world)
    full[..., lo:hi] = local
    return allreduce_sum(full, device)   # disjoint shards: a sum is a gather

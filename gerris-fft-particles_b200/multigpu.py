"""Multi-GPU plumbing for the particulate path (one process per GPU).

Particles shard across ranks (they are independent given the read-only
velocity field); the flat tree and the field are replicated on every GPU, so
no particle ever migrates and the one-way path needs no collective at all.
Two-way coupling has exactly one exchange step: the deposited field
(void fraction + force components, [1 + dim][n_cells] fp64) is summed over
ranks.  On GPUs that is the C library's own business (include/gfsb200.h,
gfsb200_comm_*: csrc/comm.cu -- owner slices over NVLink peer memory, or
ncclAllReduce); what is left here are the host-side helpers the CPU (gloo)
tests of the sharding logic use.

This replaces the reference's MPI particle migration (text-serialised
gfs_send_objects per neighbour, modules/particulatecommon.c:3218-3244) and its
scalar MPI_Allreduce id renumbering (:51-87).
"""
from __future__ import annotations

import numpy as np


def shard_bounds(n_total: int, rank: int, world: int):
    """contiguous, near-equal slices: [lo, hi) of rank `rank`"""
    return n_total * rank // world, n_total * (rank + 1) // world


def id_offset(n_total: int, rank: int, world: int) -> int:
    """first particle id (1-based) of the rank's slice -- ids stay globally
    unique without the reference's comm_size all-reduces"""
    return shard_bounds(n_total, rank, world)[0] + 1


def allreduce_host(array: np.ndarray, group=None) -> np.ndarray:
    """the same reduction on a host array (gloo), for CPU tests of the sharding logic"""
    import torch
    import torch.distributed as dist
    t = torch.from_numpy(np.ascontiguousarray(array))
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.numpy()

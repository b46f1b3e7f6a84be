"""Multi-GPU plumbing for the particulate path (one process per GPU).

Particles shard across ranks (they are independent given the read-only
velocity field); the flat tree and the field are replicated on every GPU, so
no particle ever migrates and the one-way path needs no collective at all.
Two-way coupling has exactly one exchange step: the deposited field
(void fraction + force components, [1 + dim][n_cells] fp64) is summed over
ranks with an all-reduce -- NCCL over NVLink on GPUs, issued on the context's
own stream against the C-ABI's deposit buffer (no copy).

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


class _DeviceAlias:
    """exposes a raw device pointer through __cuda_array_interface__"""

    def __init__(self, ptr: int, count: int):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": "<f8", "data": (ptr, False), "version": 3}


def deposit_tensor(ctx, device_index: int):
    """torch view (no copy) of the context's deposit buffer"""
    import torch
    ptr, count = ctx.deposit_buffer()
    return torch.as_tensor(_DeviceAlias(ptr, count), device=f"cuda:{device_index}")


def allreduce_deposit(ctx, tensor, group=None):
    """sum the deposited field over ranks, in place, ordered after the deposit
    kernels on the context's stream"""
    import torch
    import torch.distributed as dist
    stream = torch.cuda.ExternalStream(ctx.stream, device=tensor.device)
    with torch.cuda.stream(stream):
        dist.all_reduce(tensor, op=dist.ReduceOp.SUM, group=group)


def allreduce_host(array: np.ndarray, group=None) -> np.ndarray:
    """the same reduction on a host array (gloo), for CPU tests of the sharding logic"""
    import torch
    import torch.distributed as dist
    t = torch.from_numpy(np.ascontiguousarray(array))
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.numpy()


class OverlappedDepositReduce:
    """Two-way coupling without a bubble: the deposited field of step n is summed
    over ranks on a communication stream while step n+1 (cell pass, fused step,
    deposit into the OTHER buffer) runs on the context's stream.  The reduced
    field of step n is what the host fluid solver consumes; `wait(which)` orders
    a consumer (or the next deposit into that buffer) after its all-reduce."""

    def __init__(self, ctx, device_index: int, group=None):
        import torch
        self.ctx, self.group, self.torch = ctx, group, torch
        self.device = torch.device("cuda", device_index)
        self.compute = torch.cuda.ExternalStream(ctx.stream, device=self.device)
        self.comm = torch.cuda.Stream(device=self.device)
        self.tensors, self.done = [], []
        for which in (0, 1):
            ctx.deposit_select(which)
            self.tensors.append(deposit_tensor(ctx, device_index))
            self.done.append(None)
        self.which = 0
        ctx.deposit_select(0)

    def begin_step(self):
        """call before the deposit of a step: picks the buffer and makes the
        compute stream wait until that buffer's previous all-reduce has drained"""
        if self.done[self.which] is not None:
            self.compute.wait_event(self.done[self.which])
        self.ctx.deposit_select(self.which)

    def end_step(self):
        """call after the deposit: launches the all-reduce on the comm stream"""
        import torch.distributed as dist
        torch = self.torch
        ready = torch.cuda.Event()
        ready.record(self.compute)
        self.comm.wait_event(ready)
        with torch.cuda.stream(self.comm):
            dist.all_reduce(self.tensors[self.which], op=dist.ReduceOp.SUM, group=self.group)
            ev = torch.cuda.Event()
            ev.record(self.comm)
        self.done[self.which] = ev
        self.which ^= 1

    def join(self):
        """make the compute stream wait for every outstanding all-reduce"""
        for ev in self.done:
            if ev is not None:
                self.compute.wait_event(ev)

    def drain(self):
        for ev in self.done:
            if ev is not None:
                ev.synchronize()

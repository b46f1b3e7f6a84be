"""world_size-2 test of the N>1 path on CPU (gloo): particles shard across
ranks, the tree and field are replicated, and the only exchange is the sum of
the deposited field.  The per-rank deposit itself is produced here by the
oracle (no GPU in this container); the GPU deposit kernel is parity-tested
against the same oracle in test_gpu_parity.py::test_deposit."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers
from helpers import capi, worlds, ora

pkg = helpers.pkg
multigpu = pkg.multigpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _deposit(w, sim, ptrs, parts):
    live = (w.arrays.flags & capi.CELL_DESTROYED) == 0
    sim.set_values(3, ptrs[live], np.zeros(int(live.sum())))
    plist = ora.ParticleList(sim, *[parts[k] for k in ("x", "y", "z", "vx", "vy", "vz", "mass", "volume")])
    plist.deposit_volume(3)
    out = np.zeros(w.arrays.n_cells)
    out[live] = sim.get_values(3, ptrs[live])
    return out


def _worker(rank, world, port, n_total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        w = worlds.make_ring("mr", 3, 5, n_total, 77)
        sim, ptrs = helpers.matched_oracle(w)
        allp = worlds.make_particles(w)                      # every rank can regenerate the cloud
        lo, hi = multigpu.shard_bounds(n_total, rank, world)
        mine = {k: v[lo:hi] for k, v in allp.items()}
        local = _deposit(w, sim, ptrs, mine)
        total = multigpu.allreduce_host(local)
        if rank == 0:
            q.put((total, lo, hi))
        else:
            q.put((None, lo, hi))
    finally:
        dist.destroy_process_group()


def test_sharded_deposit_allreduce_equals_single_rank():
    n_total, world = 4001, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    total = next(g[0] for g in got if g[0] is not None)
    spans = sorted((g[1], g[2]) for g in got)
    assert spans[0][0] == 0 and spans[0][1] == spans[1][0] and spans[1][1] == n_total   # a partition
    w = worlds.make_ring("mr", 3, 5, n_total, 77)
    sim, ptrs = helpers.matched_oracle(w)
    want = _deposit(w, sim, ptrs, worlds.make_particles(w))
    assert np.abs(total - want).max() <= 1e-12 * np.abs(want).max()
    assert want.sum() > 0


def test_shard_bounds_partition():
    for n in (0, 1, 7, 1000, 10_000_019):
        for world in (1, 2, 3, 8):
            b = [multigpu.shard_bounds(n, r, world) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1
            assert [multigpu.id_offset(n, r, world) for r in range(world)] == [lo + 1 for lo, _ in b]

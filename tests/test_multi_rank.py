"""world_size-2 tests of the N>1 path on CPU (gloo): particles shard across
ranks, the tree and field are replicated, and the only exchange is the sum of
the deposited field.  No GPU in this container, so the per-rank deposit is
produced by the oracle; what IS the product here is the host logic of
gfsb200_comm_rebalance -- gfsb200_comm_splitters, the slice boundaries from
the all-reduced per-cell counts.  The device side of the same path (kernels,
NVLink exchange) runs in tests/test_gpu_twoway.py (group of one on any box,
two ranks under `gpurun --gpus 2`) and tools/check_twoway_ranks.py."""
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


def _rebalance_worker(rank, world, port, n_total, q):
    """what gfsb200_comm_rebalance does, with gloo for NCCL and numpy for the kernels: per-cell
    counts of the rank's share -> all-reduce -> gfsb200_comm_splitters (the product's host function)
    -> every particle goes to the rank that owns its cell"""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        w = worlds.make_ring("mr2", 3, 5, n_total, 78)
        sim, ptrs = helpers.matched_oracle(w)
        idx = helpers.PtrIndex(ptrs)
        allp = worlds.make_particles(w)
        # a random (not contiguous) share, as a rank that drew its own particles would hold
        mine = np.flatnonzero(np.random.default_rng(5).integers(0, world, n_total) == rank)
        cells = idx(sim.locate(allp["x"][mine], allp["y"][mine], allp["z"][mine]))
        assert cells.min() >= 0
        n_cells = w.arrays.n_cells
        local = np.bincount(cells, minlength=n_cells).astype(np.float64)
        counts = multigpu.allreduce_host(local).astype(np.uint32)
        split = capi.comm_splitters(counts, world)
        dest = np.searchsorted(split, cells, side="right") - 1
        out = [mine[dest == r] for r in range(world)]
        got = [None] * world
        dist.all_gather_object(got, out)
        owned = np.sort(np.concatenate([g[rank] for g in got]))
        q.put((rank, split, owned, int(counts.max())))
    finally:
        dist.destroy_process_group()


def test_rebalance_slices_partition_the_cloud_by_cell():
    n_total, world = 6007, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_rebalance_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=180) for _ in procs], key=lambda g: g[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    w = worlds.make_ring("mr2", 3, 5, n_total, 78)
    sim, ptrs = helpers.matched_oracle(w)
    idx = helpers.PtrIndex(ptrs)
    allp = worlds.make_particles(w)
    cells = idx(sim.locate(allp["x"], allp["y"], allp["z"]))
    split = got[0][1]
    assert np.array_equal(split, got[1][1])                      # every rank computed the same slices
    assert split[0] == 0 and split[-1] == w.arrays.n_cells and np.all(np.diff(split) >= 0)
    owned = [g[2] for g in got]
    assert np.array_equal(np.sort(np.concatenate(owned)), np.arange(n_total))     # a partition of the cloud
    for r in range(world):
        c = cells[owned[r]]
        assert np.all((c >= split[r]) & (c < split[r + 1]))        # by cell: no cell straddles two ranks
    sizes = [len(o) for o in owned]
    assert abs(sizes[0] - sizes[1]) <= 2 * got[0][3]               # equal up to one cell's population


def test_splitters_edge_cases():
    assert list(capi.comm_splitters(np.zeros(7, dtype=np.uint32), 3)) == [0, 7, 7, 7]
    assert list(capi.comm_splitters(np.array([4, 4, 4, 4], dtype=np.uint32), 4)) == [0, 1, 2, 3, 4]
    assert list(capi.comm_splitters(np.array([100, 0, 0, 1], dtype=np.uint32), 2)) == [0, 0, 4] or \
        list(capi.comm_splitters(np.array([100, 0, 0, 1], dtype=np.uint32), 2))[0] == 0
    s = capi.comm_splitters(np.array([0, 5, 5, 0, 10, 0, 20, 0], dtype=np.uint32), 4)
    assert s[0] == 0 and s[-1] == 8 and np.all(np.diff(s) >= 0)
    assert list(capi.comm_splitters(np.array([3, 1, 2], dtype=np.uint32), 1)) == [0, 3]


def test_owner_slices_follow_the_depth_first_leaf_order():
    """gfsb200_comm_owner_slices (the product's host function behind gfsb200_comm_rebalance on adaptive
    trees): the owners are non-decreasing along the depth-first leaf order, the shares are equal up to
    one leaf's population, every leaf has an owner and no other cell has, and inside one level a rank's
    leaves form ONE range of the (level-ordered) cell index that holds no other rank's leaf -- what the
    exchange pushes and clears."""
    w = worlds.make_ring("owners", 3, 6, 1000, 99)
    a = w.arrays
    rng = np.random.default_rng(3)
    leaves = np.flatnonzero(a.child0 < 0)
    count = np.zeros(a.n_cells, dtype=np.uint32)
    count[a.box_leaves] = rng.integers(0, 40, len(a.box_leaves))
    n_roots = int(np.sum(a.level == a.level.min()))
    for ranks in (2, 3, 8):
        owner = capi.comm_owner_slices(a.child0, n_roots, 3, count, ranks)
        assert np.all(owner[leaves] < ranks) and np.all(np.delete(owner, leaves) == 255)
        # depth-first order by hand
        order, stack = [], list(range(n_roots))[::-1]
        while stack:
            c = stack.pop()
            if a.child0[c] < 0:
                order.append(c)
            else:
                stack.extend(range(a.child0[c] + 7, a.child0[c] - 1, -1))
        order = np.array(order)
        assert len(order) == len(leaves)
        assert np.all(np.diff(owner[order].astype(int)) >= 0)
        share = np.array([count[order][owner[order] == q].sum() for q in range(ranks)])
        assert share.sum() == count.sum() and share.max() - share.min() <= 2 * 40
        for lv in np.unique(a.level[leaves]):
            cells = leaves[a.level[leaves] == lv]                 # ascending cell index
            assert np.all(np.diff(owner[cells].astype(int)) >= 0), lv

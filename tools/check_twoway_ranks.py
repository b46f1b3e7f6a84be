#!/usr/bin/env python
"""Two (or more) ranks, one process per GPU, launched with torch.distributed.run: the exchanged
deposit of the sharded cloud must equal what ONE context computes for the whole cloud.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 \
      --master-port 29511 tools/check_twoway_ranks.py

Checks (rank 0 prints "two-rank exchange ok"):
  * gfsb200_comm_rebalance: every rank ends with the particles of its cell slice, sorted by cell;
    the ids of all ranks are a permutation of the ids uploaded;
  * gfsb200_broadcast_field == gfsb200_upload_field on every rank (vorticity + corner tables);
  * N steps of fused step + deposit + gfsb200_deposit_allreduce: state equal to the one-context run
    particle by particle (by id, bit for bit), field equal to 1e-12 of its maximum, identical bytes
    on every rank, volume conserved to 1e-12 -- with particles drifting across slice boundaries
    between rebalances (their deposits are remote reductions into the owner's slice);
  * the same with GFSB200_EXCHANGE=1 (ncclAllReduce) or GFSB200_NO_P2P=1 when set by the caller.
torch.distributed is used for the rendezvous (NCCL id) and for gathering the results to rank 0.
"""
import hashlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

COLS = ("x", "y", "z", "vx", "vy", "vz", "mass", "volume")


def main():
    rank, size, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pkg = entry.load_package()
    capi, worlds = pkg.capi, pkg.worlds

    w = worlds.make_ring("ranks", 3, 6, 60000, 4242)
    n_total = 60000
    allp = worlds.make_particles(w, n_total)              # every rank can draw the whole cloud
    lo, hi = n_total * rank // size, n_total * (rank + 1) // size
    mine = {k: allp[k][lo:hi] for k in COLS}
    ids = np.arange(lo + 1, hi + 1, dtype=np.uint32)

    ctx = capi.Context(local)
    ctx.upload_tree(w.tree)
    uid = [capi.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    comm = capi.Comm.init_rank(ctx, uid[0], rank, size)
    expect_p2p = not os.environ.get("GFSB200_NO_P2P")
    assert comm.peer_access == expect_p2p, (comm.peer_access, expect_p2p)

    # field: host arrays on rank 0 only
    if rank == 0:
        comm.broadcast_field(0, w.u, w.v, w.w)
    else:
        comm.broadcast_field(0, None, None, None)
    ref = capi.Context(local)
    ref.upload_tree(w.tree)
    ref.upload_field(w.u, w.v, w.w)
    cells = w.arrays.box_leaves[::7]
    assert np.array_equal(ctx.vorticity(cells), ref.vorticity(cells))
    for comp in range(3):
        assert np.array_equal(ctx.corner_values(comp, cells), ref.corner_values(comp, cells))

    ctx.particles_upload(**mine, ids=ids)
    comm.rebalance()
    owner = comm.owner_table(w.arrays.n_cells)
    got = ctx.particles_download(ids=True)
    c = ctx.locate(got["x"], got["y"], got["z"])
    assert np.all(np.diff(c) >= 0), "not sorted by cell"
    assert np.all(owner[c] == rank), "a particle outside the rank's slice"
    counts = [None] * size
    dist.all_gather_object(counts, int(ctx.count))
    assert sum(counts) == n_total
    assert max(counts) - min(counts) <= 2000, counts        # equal shares up to one cell's population
    all_ids = [None] * size
    dist.all_gather_object(all_ids, got["id"])
    assert np.array_equal(np.sort(np.concatenate(all_ids)), np.arange(1, n_total + 1, dtype=np.uint32))

    # the one-context run of the whole cloud
    ref.particles_upload(**{k: allp[k] for k in COLS})
    ref.sort()
    par, par_f = w.step_params(), w.step_params(fuse_deposit=True)
    drifted = 0
    for step in range(24):
        ref.step(par)
        ref.deposit_all(par)
        if step % 3 == 2:
            ctx.step(par)                                  # separate passes now and then
            ctx.deposit_all(par)
        else:
            ctx.step(par_f)
        comm.deposit_allreduce()
        if step % 6 == 5 or step == 23:
            want = [ref.download_deposit(k) for k in range(4)]
            have = [ctx.download_deposit(k) for k in range(4)]
            for k in range(4):
                scale = np.abs(want[k]).max()
                assert scale > 0
                err = np.abs(have[k] - want[k]).max() / scale
                assert err <= 1e-12, (step, k, err)
            digest = hashlib.sha1(b"".join(a.tobytes() for a in have)).hexdigest()
            digests = [None] * size
            dist.all_gather_object(digests, digest)
            assert len(set(digests)) == 1, "the ranks hold different fields"
            vol = float(np.sum(have[0] * w.arrays.h ** 3))
            st = ref.particles_download()
            inside = ref.locate(st["x"], st["y"], st["z"]) >= 0
            vp = float(st["volume"][inside].sum())
            assert abs(vol - vp) <= 1e-12 * vp, (vol, vp)
            now = ctx.particles_download()
            cc = ctx.locate(now["x"], now["y"], now["z"])
            drifted += int(np.sum((cc >= 0) & (owner[np.maximum(cc, 0)] != rank)))
    total_drift = [None] * size
    dist.all_gather_object(total_drift, drifted)
    if comm.size > 1:
        assert sum(total_drift) > 0, "no particle ever left its rank's slice: the remote path was not exercised"

    # state by id, bit for bit
    mine_now = ctx.particles_download(ids=True)
    gathered = [None] * size
    dist.all_gather_object(gathered, {k: mine_now[k] for k in ("id", "x", "y", "z", "vx", "vy", "vz")})
    if rank == 0:
        st = ref.particles_download(ids=True)
        order = np.argsort(st["id"])
        cat = {k: np.concatenate([g[k] for g in gathered]) for k in gathered[0]}
        o2 = np.argsort(cat["id"])
        for k in ("x", "y", "z", "vx", "vy", "vz"):
            assert np.array_equal(cat[k][o2], st[k][order]), k

    # a second rebalance (the drifters go home) and one more exchange
    comm.rebalance()
    owner = comm.owner_table(w.arrays.n_cells)
    now = ctx.particles_download()
    cc = ctx.locate(now["x"], now["y"], now["z"])
    assert np.all(owner[cc] == rank)
    ref.step(par); ref.deposit_all(par)
    ctx.step(par_f); comm.deposit_allreduce()
    for k in range(4):
        a, b = ctx.download_deposit(k), ref.download_deposit(k)
        assert np.abs(a - b).max() <= 1e-12 * np.abs(b).max(), k
    # unequal target shares (a caller that has timed its ranks): the counts follow them
    want = np.linspace(1.0, 2.0, size)
    comm.set_shares(want)
    comm.rebalance()
    counts = [None] * size
    dist.all_gather_object(counts, int(ctx.count))
    frac = np.array(counts, dtype=float) / sum(counts)
    assert np.abs(frac - want / want.sum()).max() <= 0.02, (frac, want / want.sum())
    owner = comm.owner_table(w.arrays.n_cells)
    now = ctx.particles_download()
    cc = ctx.locate(now["x"], now["y"], now["z"])
    assert np.all(owner[cc] == rank)
    ref.step(par); ref.deposit_all(par)
    ctx.step(par_f); comm.deposit_allreduce()
    for k in range(4):
        a, b = ctx.download_deposit(k), ref.download_deposit(k)
        assert np.abs(a - b).max() <= 1e-12 * np.abs(b).max(), k
    comm.set_shares(None)
    ms, n, sent = comm.exchange_stats()
    dist.barrier()
    if rank == 0:
        share = [int(np.sum(owner == q)) for q in range(size)]
        print(f"two-rank exchange ok: {size} ranks, peer access {comm.peer_access}, leaves per rank {share}, "
              f"{sum(total_drift)} drifter observations, exchange {ms:.3f} ms over {n} calls, {sent} bytes sent per rank",
              flush=True)
    comm.close()
    ctx.close()
    ref.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""GPU tests of the two-way path added in round 2: the deposits fused into the step kernel
(gfsb200_step_params.fuse_deposit), the owner-slice deposit protocol behind a communicator
(gfsb200_comm_*: group of one on this box; two ranks when the box has two GPUs), the
force-recording flavour of the warp-pipelined kernel, and the escape-record overflow path.

Reference semantics: gfs_particle_list_event (modules/particulatecommon.c:980-1015) moves the
particles, THEN GfsParticulateField (:1934-1957) and GfsSourceParticulate (:2177-2228) deposit at the
new state with the same fluid field -- so "step, then deposit" is what the fused kernel must equal.
"""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

import helpers
from helpers import capi, worlds, ora

pytestmark = pytest.mark.gpu

STATE = ("x", "y", "z", "vx", "vy", "vz")
COLS = ("x", "y", "z", "vx", "vy", "vz", "mass", "volume")


@pytest.fixture()
def ctx():
    c = capi.Context(0)
    yield c
    c.close()


def _load(ctx, w):
    ctx.upload_tree(w.tree)
    ctx.upload_field(w.u, w.v, w.w)


def _fields(ctx, dim):
    return [ctx.download_deposit(c) for c in range(1 + dim)]


def _close(a, b, tol):
    for c, (x, y) in enumerate(zip(a, b)):
        scale = max(np.abs(y).max(), 1e-300)
        assert np.abs(x - y).max() <= tol * scale, (c, np.abs(x - y).max() / scale)


@pytest.mark.parametrize("kind", ["c1", "uniform3", "ring3", "ring2", "chain2"])
@pytest.mark.parametrize("forces", ["world", "drag"])
def test_fused_step_deposit_equals_step_then_deposit(kind, forces, ctx):
    """state bit-identical, field equal up to the order of the fp64 reductions (1e-13 of its
    maximum); and against the oracle: step, then both deposits at the new state (1e-12)."""
    w = helpers.test_world(kind)
    if forces == "drag":
        w = worlds.World(**{**w.__dict__, "forces": (capi.FORCE_DRAG,)})
    _load(ctx, w)
    n = 20011                                     # not a multiple of the 32-particle tiles
    parts = helpers.test_particles(w, n)
    ctx.particles_upload(**parts)
    ctx.sort()
    sorted_parts = ctx.particles_download()
    ctx.step(w.step_params())
    ctx.deposit_all(w.step_params())
    want_state = ctx.particles_download()
    want = _fields(ctx, w.dim)
    assert np.abs(want[0]).max() > 0 and np.abs(want[1]).max() > 0

    ctx.particles_upload(**{k: sorted_parts[k] for k in COLS})
    ctx.step(w.step_params(fuse_deposit=True))
    got_state = ctx.particles_download()
    for k in STATE:
        if want_state[k] is not None:
            assert np.array_equal(got_state[k], want_state[k]), k
    _close(_fields(ctx, w.dim), want, 1e-13)

    # the oracle: one step, then the two deposits at the new state
    sim, ptrs = helpers.matched_oracle(w)
    live = (w.arrays.flags & capi.CELL_DESTROYED) == 0
    plist = ora.ParticleList(sim, *[sorted_parts[k] for k in COLS])
    plist.step(helpers.oracle_params(w), 1)
    zero = np.zeros(int(live.sum()))
    for iv in range(3, 3 + 1 + w.dim):
        sim.set_values(iv, ptrs[live], zero)
    plist.deposit_volume(3)
    plist.deposit_force(helpers.oracle_params(w), 4)
    got = _fields(ctx, w.dim)
    for comp in range(1 + w.dim):
        ref = sim.get_values(3 + comp, ptrs[live])
        assert np.abs(got[comp][live] - ref).max() <= 1e-12 * max(np.abs(ref).max(), 1e-300), comp


def test_fused_deposit_falls_back_for_runtime_force_lists(ctx):
    """a constant drag coefficient has no compile-time kernel: the library runs the stand-alone
    deposit pass itself and the caller sees the same result"""
    w = helpers.test_world("ring3")
    _load(ctx, w)
    parts = helpers.test_particles(w, 9001)
    ctx.particles_upload(**parts)
    ctx.step(w.step_params(cd_const=0.44))
    ctx.deposit_all(w.step_params(cd_const=0.44))
    want = _fields(ctx, 3)
    ctx.particles_upload(**parts)
    ctx.step(w.step_params(cd_const=0.44, fuse_deposit=True))
    _close(_fields(ctx, 3), want, 1e-13)
    # tracers deposit no force: refused, not ignored
    with pytest.raises(capi.GfsB200Error):
        ctx.step(capi.StepParams(w.dt, (), fuse_deposit=True))


@pytest.mark.parametrize("dim", [2, 3])
def test_fused_deposit_through_the_list_event_with_periodic_wrap(dim, ctx):
    """gfs_particle_list_event with BCs: a particle that leaves through a periodic side is wrapped
    by gfs_particle_bc BEFORE the deposits run in the reference, so it deposits at the wrapped
    position; one that is dropped deposits nothing."""
    w, mask = helpers.periodic_world(dim)
    _load(ctx, w)
    parts = worlds.make_particles(w, 6000)
    rng = np.random.default_rng(3)
    parts["vx"] = parts["vx"] + rng.choice([-6.0, 6.0], len(parts["x"]))     # many crossings per step
    ctx.particles_upload(**parts)
    other = capi.Context(0)
    try:
        _load(other, w)
        other.particles_upload(**parts)
        wrapped_any = False
        for step in range(6):
            removed_a = ctx.particle_list_event(w.step_params(fuse_deposit=True))
            removed_b = other.particle_list_event(w.step_params())
            other.deposit_all(w.step_params())
            assert removed_a == removed_b
            a, b = ctx.particles_download(ids=True), other.particles_download(ids=True)
            assert np.array_equal(a["id"], b["id"])
            for k in STATE:
                if a[k] is not None:
                    assert np.array_equal(a[k], b[k]), (step, k)
            _close(_fields(ctx, dim), _fields(other, dim), 1e-13)
            wrapped_any = wrapped_any or np.any(np.abs(a["x"] - parts["x"][a["id"] - 1]) > 0.5)
        assert wrapped_any
    finally:
        other.close()


def test_recorded_forces_from_the_pipelined_kernel(ctx):
    """record_forces with a compile-time force list runs step_kernel_wpipe<REC> (round 2; the
    GModule always records): forces, cells and state equal the plain recording kernel's"""
    w = helpers.test_world("ring3")
    _load(ctx, w)
    parts = helpers.test_particles(w, 20011)
    out = {}
    for mode in ("0", "9"):
        os.environ["GFSB200_STEP_MODE"] = mode
        try:
            c = capi.Context(0)
            _load(c, w)
            c.particles_upload(**parts)
            c.step(w.step_params(record_forces=True, record_cells=True))
            out[mode] = c.particles_download(forces=True, cells=True)
            c.close()
        finally:
            del os.environ["GFSB200_STEP_MODE"]
    assert np.array_equal(out["0"]["cell"], out["9"]["cell"])
    # the plain recording kernel walks the force list at run time, this one has it compiled in:
    # same operations, contracted differently -- a few ulp
    scale = max(np.abs(out["0"][k]).max() for k in ("fx", "fy", "fz"))
    assert scale > 0
    for k in ("fx", "fy", "fz"):
        assert np.abs(out["0"][k] - out["9"][k]).max() <= 1e-13 * scale, k
    assert helpers.vec_rel_err(out["9"], out["0"], ("x", "y", "z")) <= 1e-14
    assert helpers.vec_rel_err(out["9"], out["0"], ("vx", "vy", "vz")) <= 1e-13
    sim, ptrs = helpers.matched_oracle(w)
    cells, want = helpers.oracle_step(sim, ptrs, w, parts)
    assert np.array_equal(out["9"]["cell"], cells)


def test_escape_record_overflow_drops_instead_of_failing(ctx):
    """more particles leave in one step than the record holds (n/16 + 1024): the list event must
    not fail half-way: the recorded ones get the exact BC, the rest are dropped"""
    w, mask = helpers.periodic_world(3)
    _load(ctx, w)
    n = 40000
    parts = worlds.make_particles(w, n)
    parts["vy"] = np.full(n, 400.0)             # everybody leaves through the (non-periodic) top
    ctx.particles_upload(**parts)
    removed = ctx.particle_list_event(w.step_params())
    assert removed > n // 16 + 1024
    left = ctx.particles_download()
    assert ctx.count == n - removed
    if ctx.count:
        assert np.all(ctx.locate(left["x"], left["y"], left["z"]) >= 0)


# ---------------------------------------------------------------------------
# communicator: group of one (runs on the 1-GPU test box)

def test_group_of_one_owner_slices_equal_plain_deposit(ctx):
    w = helpers.test_world("ring3")
    _load(ctx, w)
    parts = helpers.test_particles(w, 30011)
    plain = capi.Context(0)
    try:
        _load(plain, w)
        plain.particles_upload(**parts)
        plain.sort()
        comm = capi.Comm.init_rank(ctx, None, 0, 1)
        assert comm.size == 1 and comm.rank == 0 and comm.peer_access
        ctx.particles_upload(**parts)
        comm.rebalance()
        split = comm.split()
        assert list(split) == [0, w.arrays.n_cells]
        a, b = ctx.particles_download(ids=True), plain.particles_download(ids=True)
        assert np.array_equal(np.sort(a["id"]), np.sort(b["id"]))
        cells = ctx.locate(a["x"], a["y"], a["z"])
        assert np.all(np.diff(cells) >= 0)                      # globally (here: locally) cell-sorted
        for step in range(4):                                   # both deposit buffers, twice
            plain.step(w.step_params())
            plain.deposit_all(w.step_params())
            if step % 2 == 0:
                ctx.step(w.step_params(fuse_deposit=True))
            else:
                ctx.step(w.step_params())
                ctx.deposit_volume()
                ctx.deposit_force(w.step_params())
            comm.deposit_allreduce()
            _close(_fields(ctx, 3), _fields(plain, 3), 1e-13)
        # a second deposit of a component without the exchange in between is refused
        ctx.deposit_volume()
        with pytest.raises(capi.GfsB200Error):
            ctx.deposit_volume()
        comm.deposit_allreduce()
        # whole-buffer mode on the same communicator needs a fresh start of the slices
        comm.set_exchange(capi.EXCHANGE_ALLREDUCE)
        ctx.deposit_all(w.step_params())
        comm.deposit_allreduce()
        plain.deposit_all(w.step_params())
        _close(_fields(ctx, 3), _fields(plain, 3), 1e-13)
        comm.set_exchange(capi.EXCHANGE_AUTO)
        with pytest.raises(capi.GfsB200Error):
            ctx.deposit_all(w.step_params())                    # slices were not cleared ahead
        comm.rebalance()
        ctx.deposit_all(w.step_params())
        comm.deposit_allreduce()
        _close(_fields(ctx, 3), _fields(plain, 3), 1e-13)
        # a new tree drops the ownership; deposits keep working (whole buffer) until the next rebalance
        _load(ctx, w)
        ctx.deposit_all(w.step_params())
        comm.deposit_allreduce()
        _close(_fields(ctx, 3), _fields(plain, 3), 1e-13)
        comm.close()
    finally:
        plain.close()


def test_group_of_one_broadcast_field_and_smoothed_deposit(ctx):
    w = helpers.test_world("uniform3")
    ctx.upload_tree(w.tree)
    comm = capi.Comm.init_all([ctx])
    comm.broadcast_field(0, w.u, w.v, w.w)
    plain = capi.Context(0)
    try:
        _load(plain, w)
        cells = w.arrays.box_leaves[:500]
        assert np.array_equal(ctx.vorticity(cells), plain.vorticity(cells))
        for comp in range(3):
            assert np.array_equal(ctx.corner_values(comp, cells), plain.corner_values(comp, cells))
        parts = helpers.test_particles(w, 4000)
        for c in (ctx, plain):
            c.particles_upload(**parts)
        comm.rebalance()
        ctx.deposit_force_smoothed(w.step_params(), 0.08)       # local to the rank: fine on one rank
        comm.deposit_allreduce()
        plain.deposit_force_smoothed(w.step_params(), 0.08)
        for comp in (1, 2, 3):
            x, y = ctx.download_deposit(comp), plain.download_deposit(comp)
            assert np.abs(x - y).max() <= 1e-12 * np.abs(y).max()
        comm.close()
    finally:
        plain.close()


# ---------------------------------------------------------------------------
# two ranks (one process per GPU): only where the box has two GPUs -- `gpurun --gpus 2`

def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("exchange", ["auto", "allreduce", "no_p2p"])
def test_two_ranks_exchange_equals_one_rank(exchange):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    script = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tools", "check_twoway_ranks.py")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(_free_port()))
    if exchange == "allreduce":
        env["GFSB200_EXCHANGE"] = "1"
    if exchange == "no_p2p":
        env["GFSB200_NO_P2P"] = "1"
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", env["MASTER_PORT"], script],
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "two-rank exchange ok" in r.stdout

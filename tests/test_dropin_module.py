"""The drop-in boundary, executed: gerris-fft-particles_b200/host/particulates_b200.c
-- the GModule source a Gerris installation would build -- linked with the
reference's own object code (modules/particulatecommon.c, src/event.c, ...) and
the GLib/GTS run-time of oracle/refobj/glue.c into oracle/_ref/libgfsrefmod*.so.

Its g_module_check_init() is run exactly as GLib would on g_module_open
(src/simulation.c:199-225); real GfsParticleList / GfsParticulateField /
GfsSourceParticulate objects are then driven through gfs_event_do, once in this
library (events routed to the B200 through the C-ABI: FttCell bridge -> flat
tree -> CUDA) and once in libgfsrefobj (the unmodified reference), and the
GtsObject state left behind is compared.
"""
import numpy as np
import pytest

import helpers
from helpers import capi, worlds, ora

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not (ora.refobj_available(2, True) and ora.refobj_available(3, True)),
                                 reason="oracle/_ref/libgfsrefmod*.so not built (make -C oracle)")]

KEYS = ("x", "y", "z", "vx", "vy", "vz", "mass", "volume")
_cache = {}


def setup(kind):
    if kind not in _cache:
        w = helpers.test_world(kind)
        sim, ptrs = helpers.matched_oracle(w)
        _cache[kind] = (w, sim, ptrs)
    return _cache[kind]


def run_list(sim, parts, par, steps, module, periodic_mask=0, before_step=None):
    """states after each of `steps` list events, in the reference (module=False)
    or with the drop-in module loaded (module=True).  One RefSim at a time: the
    GfsBox objects hang on the shared root cells."""
    rs = ora.RefSim(sim, periodic_mask, module=module)
    rs.configure(par)
    rl = ora.RefParticleList(rs, *[parts[k] for k in KEYS], par)
    out = []
    for step in range(steps):
        if before_step:
            before_step(step)
        assert rl.event() == 1
        out.append(rl.get())
    rs.close()
    return out


def check(got, want, dim, rtol, what):
    go, wo = np.argsort(got["id"]), np.argsort(want["id"])
    assert np.array_equal(got["id"][go], want["id"][wo]), what
    for keys in (("x", "y", "z")[:dim], ("vx", "vy", "vz")[:dim]):
        g = {k: got[k][go] for k in keys}
        w = {k: want[k][wo] for k in keys}
        err = helpers.vec_rel_err(g, w, keys)
        assert err <= rtol, (what, keys, err)
    for k in ("fx", "fy", "fz")[:dim]:
        scale = max(np.abs(want[k]).max(), 1e-300)
        assert np.abs(got[k][go] - want[k][wo]).max() <= 1e-10 * scale, (what, k)
    assert np.abs(got["mass"][go] - want["mass"][wo]).max() <= 1e-14 * np.abs(want["mass"]).max(), what


@pytest.mark.parametrize("resident", ["0", None])
@pytest.mark.parametrize("kind", ["c1", "uniform3", "ring3", "ring2", "ring3b", "chain2", "chain3"])
def test_module_particle_list_event(kind, resident, monkeypatch):
    """GfsParticleList.event re-pointed by the module: flatten the live FttCell
    trees, mirror U,V,W, cull + fused step on the device, state written back into the
    GfsParticulate objects, gfs_particle_bc on the host -- against the reference's own
    gfs_particle_list_event over 3 steps.  resident "0": the objects are written back after every
    event (GFSB200_RESIDENT=0); None: the default, the device copy is authoritative and the
    harness reads the objects through gfsb200_module_sync"""
    if resident is None:
        monkeypatch.delenv("GFSB200_RESIDENT", raising=False)
    else:
        monkeypatch.setenv("GFSB200_RESIDENT", resident)
    w, sim, ptrs = setup(kind)
    parts = helpers.test_particles(w, 4000)
    parts["x"][::97] = 5.0                      # a few particles outside: culled by both
    par = helpers.oracle_params(w)
    want = run_list(sim, parts, par, 3, module=False)
    got = run_list(sim, parts, par, 3, module=True)
    for step in range(3):
        assert len(got[step]["x"]) == len(want[step]["x"]) < len(parts["x"])
        check(got[step], want[step], w.dim, 1e-12 if step == 0 else 1e-11, (kind, step))


@pytest.mark.parametrize("resident", ["0", "1"])
@pytest.mark.parametrize("kind", ["c1", "ring3", "chain2"])
def test_module_tracer_list(kind, resident, monkeypatch):
    """a list without forces: the reference falls through to gfs_particle_event /
    gfs_domain_advect_point (src/particle.c:31-44), the module to the device's RK2 tracer
    kernel; escapes are not tracked there, so the host BCs run every event in either mode"""
    monkeypatch.setenv("GFSB200_RESIDENT", resident)
    w, sim, ptrs = setup(kind)
    parts = helpers.test_particles(w, 3000)
    par = ora.step_params(w.dt, [])
    want = run_list(sim, parts, par, 3, module=False)
    got = run_list(sim, parts, par, 3, module=True)
    for step in range(3):
        go, wo = np.argsort(got[step]["id"]), np.argsort(want[step]["id"])
        assert np.array_equal(got[step]["id"][go], want[step]["id"][wo])
        keys = ("x", "y", "z")[:w.dim]
        g = {k: got[step][k][go] for k in keys}
        wv = {k: want[step][k][wo] for k in keys}
        assert helpers.vec_rel_err(g, wv, keys) <= 1e-12, (kind, step)
        for k in ("vx", "vy", "vz")[:w.dim]:                     # tracers keep their velocity field
            assert np.array_equal(got[step][k][go], want[step][k][wo])


@pytest.mark.parametrize("kind", ["c1", "ring3"])
@pytest.mark.parametrize("forces,kw", [
    ((ora.FORCE_INERTIAL, ora.FORCE_DRAG), {}),
    ((ora.FORCE_ADDEDMASS, ora.FORCE_DRAG, ora.FORCE_BUOY), dict(cm_const=0.3)),
    ((ora.FORCE_DRAG, ora.FORCE_LIFT), dict(cd_const=0.44, cl_const=0.25)),
])
def test_module_force_lists(kind, forces, kw):
    """force-list wiring of the module (step_params): inertial / added-mass with the
    Un,Vn,Wn upload and the reference's snapshot rule, constant coefficient functions"""
    w, sim, ptrs = setup(kind)
    a = w.arrays
    rng = np.random.default_rng(21)
    fields = [w.u, w.v] + ([w.w] if w.dim == 3 else [])
    prev = [0.8 * f + 0.05 * rng.standard_normal(a.n_cells) for f in fields]
    live = (a.flags & capi.CELL_DESTROYED) == 0

    def reset_un(step):
        if step == 0:
            for c, f in enumerate(prev):
                sim.set_values(5 + c, ptrs[live], f[live])

    parts = helpers.test_particles(w, 2000)
    par = ora.step_params(w.dt, list(forces), rho=w.rho, mu=w.mu, g=(0.1, -1.0, 0.0), ivar_uold=5, **kw)
    want = run_list(sim, parts, par, 3, module=False, before_step=reset_un)
    got = run_list(sim, parts, par, 3, module=True, before_step=reset_un)
    for step in range(3):
        check(got[step], want[step], w.dim, 1e-12 if step == 0 else 1e-11, (kind, forces, step))
    if ora.FORCE_ADDEDMASS in forces:
        assert np.all(got[2]["mass"] > parts["mass"])


@pytest.mark.parametrize("kind", ["ring2", "uniform3"])
def test_module_per_cell_alpha_and_viscosity(kind):
    """PhysicalParams alpha = <variable> and a variable viscosity (GfsDiffusion.mu) are
    gathered per cell and mirrored with the velocity"""
    w, sim, ptrs = setup(kind)
    a = w.arrays
    live = (a.flags & capi.CELL_DESTROYED) == 0
    rng = np.random.default_rng(4)
    sim.set_values(3, ptrs[live], rng.uniform(0.5, 2.0, int(live.sum())))
    sim.set_values(4, ptrs[live], rng.uniform(5e-4, 2e-3, int(live.sum())))
    parts = helpers.test_particles(w, 3000)
    par = helpers.oracle_params(w, ivar_alpha=3, ivar_mu=4)
    want = run_list(sim, parts, par, 2, module=False)
    got = run_list(sim, parts, par, 2, module=True)
    for step in range(2):
        check(got[step], want[step], w.dim, 1e-12 if step == 0 else 1e-11, (kind, step))


@pytest.mark.parametrize("dim", [2, 3])
def test_module_periodic_run(dim):
    """15 steps in a box with periodic and plain sides: device cull + step, the
    reference's gfs_particle_bc on the host objects afterwards, by particle id"""
    w, mask = helpers.periodic_world(dim)
    sim, ptrs = helpers.matched_oracle(w)
    rng = np.random.default_rng(9)
    parts = worlds.make_particles(w)
    n = len(parts["x"])
    for k in ("x", "y", "z")[:dim]:
        parts[k] = rng.uniform(-0.499, 0.499, n)
    for k, f in zip(("vx", "vy", "vz")[:dim], (3.0, 3.0, 2.0)):
        parts[k] = f * rng.standard_normal(n)
    par = helpers.oracle_params(w)
    want = run_list(sim, parts, par, 15, module=False, periodic_mask=mask)
    got = run_list(sim, parts, par, 15, module=True, periodic_mask=mask)
    assert len(want[-1]["x"]) < n
    for step in range(15):
        assert len(got[step]["x"]) == len(want[step]["x"]), step
        check(got[step], want[step], dim, 1e-11, ("periodic", dim, step))


@pytest.mark.parametrize("kind", ["c1", "ring3", "chain2"])
def test_module_particulate_field_event(kind):
    """GfsParticulateField.event re-pointed: void fraction deposited on the device and
    scattered back into the cell variable (gfs_cell_reset + :1945-1953)"""
    w, sim, ptrs = setup(kind)
    a = w.arrays
    parts = helpers.test_particles(w, 20000)
    par = helpers.oracle_params(w)
    live = (a.flags & capi.CELL_DESTROYED) == 0
    leaves = live & (a.child0 < 0) & ((a.flags & capi.CELL_BOUNDARY) == 0)
    res = []
    for module in (False, True):
        sim.set_values(3, ptrs[live], np.full(int(live.sum()), 7.0))      # must be reset by the event
        rs = ora.RefSim(sim, module=module)
        rs.configure(par)
        rl = ora.RefParticleList(rs, *[parts[k] for k in KEYS], par)
        rl.field_event(3)
        res.append(sim.get_values(3, ptrs[leaves]))
        # neither the reference nor the module touches ghost or non-leaf cells
        assert np.all(sim.get_values(3, ptrs[live & ~leaves]) == 7.0)
        rs.close()
    want, got = res
    assert want.max() > 0
    assert np.abs(got - want).max() <= 1e-12 * np.abs(want).max()


@pytest.mark.parametrize("kind,rk,kernel", [
    ("c1", 0.06, (ora.KERNEL_GAUSSIAN, 1.0, 2e-4, 1)),
    ("ring3", 0.04, (ora.KERNEL_COMPACT, 2.0, 1e-4, 2)),
    ("ring2", 0.05, (ora.KERNEL_CONSTANT, 1.0, 0.0, 1)),
])
def test_module_source_particulate_event(kind, rk, kernel):
    """GfsSourceParticulate.event re-pointed: the user's kernel GfsFunction is recognised
    by probing (gfsb200_kernel_fit through gfs_function_spatial_value), the smoothed
    force deposited on the device and scattered into <plist>_Fx,_Fy,_Fz; the on-fluid
    force is left in particulate->force as the reference does (:2195-2201)"""
    w, sim, ptrs = setup(kind)
    a = w.arrays
    parts = helpers.test_particles(w, 600)
    par = helpers.oracle_params(w)
    live = (a.flags & capi.CELL_DESTROYED) == 0
    leaves = live & (a.child0 < 0) & ((a.flags & capi.CELL_BOUNDARY) == 0)
    k = ora.Kernel(*kernel, 0)
    res, forces = [], []
    for module in (False, True):
        for iv in range(4, 4 + w.dim):
            sim.set_values(iv, ptrs[live], np.full(int(live.sum()), 3.0))
        rs = ora.RefSim(sim, module=module)
        rs.configure(par)
        rl = ora.RefParticleList(rs, *[parts[q] for q in KEYS], par)
        rl.source_event(4, rk, k)
        res.append([sim.get_values(4 + c, ptrs[leaves]) for c in range(w.dim)])
        forces.append(rl.get())
        rs.close()
    for c in range(w.dim):
        want, got = res[0][c], res[1][c]
        assert np.abs(want).max() > 0
        assert np.abs(got - want).max() <= 1e-12 * np.abs(want).max(), c
    for q in ("fx", "fy", "fz")[:w.dim]:
        scale = max(np.abs(forces[0][q]).max(), 1e-300)
        assert np.abs(forces[1][q] - forces[0][q]).max() <= 1e-10 * scale, q


def test_module_counts_device_launches():
    """the events above really ran on the device: the library's own launch counter moved"""
    w, sim, ptrs = setup("ring3")
    before = capi.kernel_launches()
    parts = helpers.test_particles(w, 500)
    run_list(sim, parts, helpers.oracle_params(w), 1, module=True)
    assert capi.kernel_launches() > before


def run_resident(sim, parts, par, steps, record_at, periodic_mask=0):
    """the module with $GFSB200_RESIDENT=1 (set by the caller): objects are only looked at after
    an explicit gfsb200_module_sync at the steps in record_at"""
    rs = ora.RefSim(sim, periodic_mask, module=True)
    rs.configure(par)
    rl = ora.RefParticleList(rs, *[parts[k] for k in KEYS], par)
    out, stale = {}, None
    for step in range(1, steps + 1):
        assert rl.event() == 1
        if step == 1:
            stale = rl.get(sync=False)            # NOT synced: what a careless reader would see
        if step in record_at:
            assert rl.sync()
            out[step] = rl.get()
    rs.close()
    return out, stale


@pytest.mark.parametrize("kind", ["uniform3", "ring2", "chain3"])
def test_module_resident_mode(kind, monkeypatch, tmp_path):
    """resident mode (the default since round 2; GFSB200_RESIDENT=0 turns it off): the device copy
    is authoritative between events (no object gather / scatter per event); gfsb200_module_sync,
    the list's write method and the reference's own reader classes refresh the objects on demand"""
    monkeypatch.delenv("GFSB200_RESIDENT", raising=False)
    w, sim, ptrs = setup(kind)
    parts = helpers.test_particles(w, 3000)
    par = helpers.oracle_params(w)
    want = run_list(sim, parts, par, 6, module=False)
    got, stale = run_resident(sim, parts, par, 6, record_at=(3, 6))
    # after the first event the objects still hold the initial state: nothing was written back
    assert np.array_equal(stale["x"], parts["x"]) and np.array_equal(stale["vx"], parts["vx"])
    for step in (3, 6):
        check(got[step], want[step - 1], w.dim, 1e-11, (kind, step))
    # the list's own write method syncs first: the dump holds the state after the last event
    rs = ora.RefSim(sim, module=True)
    rs.configure(par)
    rl = ora.RefParticleList(rs, *[parts[k] for k in KEYS], par)
    rl.event(2)
    rl.class_write(tmp_path / "list.gfs")
    after = rl.get()
    rs.close()
    text = (tmp_path / "list.gfs").read_text()
    assert text.startswith("GfsParticleList") and text.count("GfsParticulate ") >= 3000
    check(after, want[1], w.dim, 1e-11, (kind, "write"))


@pytest.mark.parametrize("dim", [2, 3])
def test_module_resident_mode_with_boundaries(dim, monkeypatch):
    """resident mode where particles keep leaving: every event with an escape falls back to
    download + the reference's gfs_particle_bc (pos_old patched from the device's record),
    events without one stay on the device"""
    monkeypatch.setenv("GFSB200_RESIDENT", "1")
    w, mask = helpers.periodic_world(dim)
    sim, ptrs = helpers.matched_oracle(w)
    rng = np.random.default_rng(9)
    parts = worlds.make_particles(w, 600)
    n = len(parts["x"])
    for k in ("x", "y", "z")[:dim]:
        parts[k] = rng.uniform(-0.499, 0.499, n)
    for k, f in zip(("vx", "vy", "vz")[:dim], (1.0, 1.0, 0.7)):
        parts[k] = f * rng.standard_normal(n)
    par = helpers.oracle_params(w)
    want = run_list(sim, parts, par, 20, module=False, periodic_mask=mask)
    got, _ = run_resident(sim, parts, par, 20, record_at=(5, 10, 20), periodic_mask=mask)
    assert len(want[-1]["x"]) < n
    for step in (5, 10, 20):
        assert len(got[step]["x"]) == len(want[step - 1]["x"]), step
        check(got[step], want[step - 1], dim, 1e-10, ("resident periodic", dim, step))


def test_module_c1_full_config(monkeypatch):
    """BASELINE config C1 as the reference would run it -- gerris2D, `GModule particulates`,
    2D level-6 lid-style cavity with four ghost layers, 1000 GfsParticulate objects,
    GfsForceDrag, dt = 1e-2 -- for 1000 list events through the drop-in module (resident mode,
    objects synced every 100 events) against 1000 events of the reference's own object code"""
    monkeypatch.setenv("GFSB200_RESIDENT", "1")
    w = worlds.make_c1()
    assert w.arrays.n_leaves == 64 * 64 and w.arrays.n_roots == 5
    sim, ptrs = helpers.matched_oracle(w)
    parts = worlds.make_particles(w)
    assert len(parts["x"]) == 1000
    par = helpers.oracle_params(w)
    marks = tuple(range(100, 1001, 100))
    want = run_list(sim, parts, par, 1000, module=False)
    got, _ = run_resident(sim, parts, par, 1000, record_at=marks)
    worst = 0.0
    for step in marks:
        g, wv = got[step], want[step - 1]
        assert np.array_equal(g["id"], wv["id"])                  # nobody leaves the cavity
        for keys in (("x", "y"), ("vx", "vy")):
            worst = max(worst, helpers.vec_rel_err(g, wv, keys))
    assert worst <= 1e-9, worst        # the documented drift bound after 1000 steps (drag only: no
                                       # cell-constant force, so no particle is excluded)


@pytest.mark.parametrize("kind", ["c1", "ring3"])
def test_module_with_embedded_solids(kind):
    """a simulation with a (static) GfsSolid: the bridge carries the mixed cells'
    GfsSolidVector into the flat tree, and list event, void-fraction field and smoothed
    source through the module match the reference's own events on the same mixed tree"""
    w, sim, ptrs = setup(kind)
    a = w.arrays
    rng = np.random.default_rng(6)
    leaves = a.box_leaves
    mixed = rng.choice(leaves, len(leaves) // 4, replace=False)
    for i in mixed:
        h = 2.0 ** -int(a.level[i])
        sim.set_solid(ptrs[i], float(rng.uniform(0.05, 1.0)), a.pos[i, :a.dim] + rng.uniform(-0.45, 0.45, a.dim) * h,
                      rng.choice([0.0, 0.3, 1.0], 2 * a.dim, p=[0.2, 0.4, 0.4]))
    try:
        parts = helpers.test_particles(w, 3000)
        par = helpers.oracle_params(w)
        live = (a.flags & capi.CELL_DESTROYED) == 0
        box_leaf = live & (a.child0 < 0) & ((a.flags & capi.CELL_BOUNDARY) == 0)
        res = {}
        for module in (False, True):
            rs = ora.RefSim(sim, module=module)
            rs.configure(par)
            rs.add_solid()
            rl = ora.RefParticleList(rs, *[parts[k] for k in KEYS], par)
            states = []
            for step in range(3):
                assert rl.event() == 1
                states.append(rl.get())
            for iv in range(4, 4 + w.dim):
                sim.set_values(iv, ptrs[live], np.full(int(live.sum()), 3.0))
            rl.source_event(4, 0.05, ora.Kernel(ora.KERNEL_GAUSSIAN, 1.0, 2e-4, 1, 0))
            src = [sim.get_values(4 + c, ptrs[box_leaf]) for c in range(w.dim)]
            res[module] = (states, src)
            rs.close()
        for step in range(3):
            check(res[True][0][step], res[False][0][step], w.dim, 1e-12 if step == 0 else 1e-11, (kind, step))
        for c in range(w.dim):
            want, got = res[False][1][c], res[True][1][c]
            assert np.abs(want).max() > 0
            assert np.abs(got - want).max() <= 1e-12 * np.abs(want).max(), c
    finally:
        for i in mixed:
            sim.set_solid(ptrs[i], 0.0, np.zeros(3))


# ---------------------------------------------------------------------------
# round 2: mesh identity, solid cells, the no-fallback policy

@pytest.mark.parametrize("resident", ["0", None])
def test_module_reflattens_after_every_adapt(resident, monkeypatch):
    """the mesh changes between events -- cells refined (gfs_cell_fine_init) and
    coarsened (gfs_cell_cleanup) as gfs_simulation_adapt does, twice with the SAME created /
    removed counts, i.e. exactly the case sim->adapts_stats cannot tell apart -- and the module must
    flatten the live tree again each time (its hidden variable's coarse_fine / cleanup methods see
    every refined and destroyed cell).  A stale flat tree would interpolate with the old stencils
    and its map would point at freed FttCells."""
    if resident is None:
        monkeypatch.delenv("GFSB200_RESIDENT", raising=False)
    else:
        monkeypatch.setenv("GFSB200_RESIDENT", resident)
    w = helpers.test_world("ring3")
    parts = helpers.test_particles(w, 3000)
    par = helpers.oracle_params(w)
    a = w.arrays
    # coarse leaves of the box, away from the hull (their neighbours exist), in flat order
    cand = [i for i in a.box_leaves if a.level[i] == 3 and np.all(np.abs(a.pos[i]) < 0.3)][:24]
    assert len(cand) == 24
    states = {}
    for module in (False, True):
        sim, ptrs = helpers.matched_oracle(w)          # one tree per run: the adapts are destructive
        rs = ora.RefSim(sim, module=module)
        rs.configure(par)
        rl = ora.RefParticleList(rs, *[parts[k] for k in KEYS], par)
        out = []
        for step in range(6):
            if step in (1, 2):                          # two adapts, 8 cells refined in each
                for i in cand[8 * (step - 1):8 * step]:
                    assert rs.refine(ptrs[i])
            if step == 4:                               # and a coarsening
                for i in cand[:4]:
                    assert rs.coarsen(ptrs[i])
            assert rl.event() == 1
            out.append(rl.get())
        rs.close()
        states[module] = out
    for step in range(6):
        check(states[True][step], states[False][step], 3, 1e-12 if step == 0 else 1e-11, ("adapt", step))


def test_module_particle_in_a_solid_cell_is_culled_not_fatal(monkeypatch):
    """with entirely solid (destroyed) cells inside the box the hull test does
    not decide whether a particle is still in the domain.  A particle that steps into a solid cell
    must leave the list through gfs_particle_bc in the same event, as in the reference
    (modules/particulatecommon.c:3333-3335) -- in resident mode too, where round 1 aborted."""
    monkeypatch.delenv("GFSB200_RESIDENT", raising=False)
    w = helpers.test_world("ring3")
    a = w.arrays
    # a coarse leaf well inside the box becomes entirely solid
    nbr = a.neighbor.reshape(a.n_cells, 6)
    dead = next(int(i) for i in a.box_leaves if a.level[i] == 3 and np.all(np.abs(a.pos[i]) < 0.32) and
                nbr[i, 0] >= 0 and a.level[nbr[i, 0]] == 3 and a.child0[nbr[i, 0]] < 0)
    parts = helpers.test_particles(w, 400)
    # a handful of particles right next to the solid cell, moving into it
    h = a.h[dead]
    for j in range(10):
        parts["x"][j] = a.pos[dead, 0] + 0.6 * h
        parts["y"][j] = a.pos[dead, 1] + (j - 5) * 0.05 * h
        parts["z"][j] = a.pos[dead, 2]
        parts["vx"][j], parts["vy"][j], parts["vz"][j] = -0.5 * h / w.dt, 0.0, 0.0
    par = helpers.oracle_params(w)
    res = {}
    for module in (False, True):
        sim, ptrs = helpers.matched_oracle(w)
        rs = ora.RefSim(sim, module=module)
        rs.configure(par)
        assert rs.destroy_cell(ptrs[dead])
        rl = ora.RefParticleList(rs, *[parts[k] for k in KEYS], par)
        out = []
        for step in range(4):
            assert rl.event() == 1
            out.append(rl.get())
        rs.close()
        res[module] = out
    assert len(res[False][0]["x"]) <= 390                     # the reference dropped them in event 1
    for step in range(4):
        assert len(res[True][step]["x"]) == len(res[False][step]["x"]), step
        check(res[True][step], res[False][step], 3, 1e-11, ("solid cell", step))


def test_module_unknown_kernel_needs_the_opt_in(monkeypatch):
    """a smoothing kernel outside the closed forms: with GFSB200_ALLOW_REFERENCE_EVENT=1 the
    reference's own event runs (one warning), and the result is the reference's, bit for bit"""
    monkeypatch.setenv("GFSB200_ALLOW_REFERENCE_EVENT", "1")
    w, sim, ptrs = setup("c1")
    a = w.arrays
    parts = helpers.test_particles(w, 300)
    par = helpers.oracle_params(w)
    live = (a.flags & capi.CELL_DESTROYED) == 0
    k = ora.Kernel(ora.KERNEL_ODD, 1.0, 2e-4, 1, 0) if hasattr(ora, "KERNEL_ODD") else None
    if k is None:
        pytest.skip("the harness has no kernel outside the closed forms")
    res = []
    for module in (False, True):
        for iv in range(4, 4 + w.dim):
            sim.set_values(iv, ptrs[live], np.full(int(live.sum()), 3.0))
        rs = ora.RefSim(sim, module=module)
        rs.configure(par)
        rl = ora.RefParticleList(rs, *[parts[q] for q in KEYS], par)
        rl.source_event(4, 0.06, k)
        res.append([sim.get_values(4 + c, ptrs[live]) for c in range(w.dim)])
        rs.close()
    for c in range(w.dim):
        assert np.array_equal(res[0][c], res[1][c])


# ---------------------------------------------------------------------------
# GFSB200_DEVICES=2: one serial process drives two GPUs (needs `gpurun --gpus 2`)

def _two_gpus():
    return capi.lib().gfsb200_device_count() >= 2


@pytest.mark.skipif("not _two_gpus()", reason="needs two GPUs")
@pytest.mark.parametrize("resident", ["0", None])
@pytest.mark.parametrize("kind", ["c1", "ring3", "chain3"])
def test_module_two_devices_list_field_and_source_events(kind, resident, monkeypatch):
    """the list sharded over two GPUs by gfsb200_comm_rebalance (objects found again by particle
    id), the field broadcast over NVLink, the deposits summed with gfsb200_deposit_allreduce before
    they are scattered into the cells: list event (with culled particles), GfsParticulateField and
    GfsSourceParticulate against the reference's own events"""
    monkeypatch.setenv("GFSB200_DEVICES", "2")
    if resident is None:
        monkeypatch.delenv("GFSB200_RESIDENT", raising=False)
    else:
        monkeypatch.setenv("GFSB200_RESIDENT", resident)
    w, sim, ptrs = setup(kind)
    a = w.arrays
    parts = helpers.test_particles(w, 5000)
    parts["x"][::101] = 7.0                              # culled on whichever device holds them
    par = helpers.oracle_params(w)
    live = (a.flags & capi.CELL_DESTROYED) == 0
    leaves = live & (a.child0 < 0) & ((a.flags & capi.CELL_BOUNDARY) == 0)
    kern = ora.Kernel(ora.KERNEL_GAUSSIAN, 1.0, 2e-4, 1, 0)
    res = {}
    for module in (False, True):
        for iv in range(3, 4 + w.dim):
            sim.set_values(iv, ptrs[live], np.full(int(live.sum()), 5.0))
        rs = ora.RefSim(sim, module=module)
        rs.configure(par)
        rl = ora.RefParticleList(rs, *[parts[k] for k in KEYS], par)
        states = []
        for step in range(3):
            assert rl.event() == 1
            states.append(rl.get())
        rl.field_event(3)
        vol = sim.get_values(3, ptrs[leaves])
        rl.source_event(4, 0.05, kern)
        src = [sim.get_values(4 + c, ptrs[leaves]) for c in range(w.dim)]
        forces = rl.get()
        rs.close()
        res[module] = (states, vol, src, forces)
    for step in range(3):
        assert len(res[True][0][step]["x"]) == len(res[False][0][step]["x"]) < len(parts["x"])
        check(res[True][0][step], res[False][0][step], w.dim, 1e-12 if step == 0 else 1e-11, (kind, step))
    assert res[False][1].max() > 0
    assert np.abs(res[True][1] - res[False][1]).max() <= 1e-12 * np.abs(res[False][1]).max()
    for c in range(w.dim):
        want, got = res[False][2][c], res[True][2][c]
        assert np.abs(want).max() > 0
        assert np.abs(got - want).max() <= 1e-12 * np.abs(want).max(), c
    go, wo = np.argsort(res[True][3]["id"]), np.argsort(res[False][3]["id"])
    for q in ("fx", "fy", "fz")[:w.dim]:
        scale = max(np.abs(res[False][3][q]).max(), 1e-300)
        assert np.abs(res[True][3][q][go] - res[False][3][q][wo]).max() <= 1e-10 * scale, q


@pytest.mark.skipif("not _two_gpus()", reason="needs two GPUs")
def test_module_two_devices_resident_with_boundaries(monkeypatch):
    """two GPUs, resident mode, particles leaving all the time: pos_old of an escaped particle is
    found through the device's ids, the host's gfs_particle_bc wraps or drops it"""
    monkeypatch.setenv("GFSB200_DEVICES", "2")
    monkeypatch.delenv("GFSB200_RESIDENT", raising=False)
    w, mask = helpers.periodic_world(3)
    sim, ptrs = helpers.matched_oracle(w)
    rng = np.random.default_rng(9)
    parts = worlds.make_particles(w, 600)
    n = len(parts["x"])
    for k in ("x", "y", "z"):
        parts[k] = rng.uniform(-0.499, 0.499, n)
    for k, f in zip(("vx", "vy", "vz"), (1.0, 1.0, 0.7)):
        parts[k] = f * rng.standard_normal(n)
    par = helpers.oracle_params(w)
    want = run_list(sim, parts, par, 20, module=False, periodic_mask=mask)
    got, _ = run_resident(sim, parts, par, 20, record_at=(5, 10, 20), periodic_mask=mask)
    assert len(want[-1]["x"]) < n
    for step in (5, 10, 20):
        assert len(got[step]["x"]) == len(want[step - 1]["x"]), step
        check(got[step], want[step - 1], 3, 1e-10, ("two devices, periodic", step))

"""The parity pin: the restated oracle (oracle/particulate_port.c) against THE
REFERENCE'S OWN OBJECT CODE for the particulate layer.

oracle/_ref/libgfsrefobj{2D,3D}.so holds modules/particulatecommon.c, src/event.c,
src/particle.c, src/fluid.c and src/ftt.c compiled unmodified from
/root/reference (oracle/Makefile; run-time in oracle/refobj/glue.c).  These tests
drive the reference's GfsParticleList / GfsParticulate / GfsForce* /
GfsParticulateField / GfsSourceParticulate objects through gfs_event_do -- the
call simulation_run makes -- and require the restatement, which the GPU parity
tests compare the CUDA path with, to reproduce them BIT FOR BIT (both are
compiled by the same gcc with -ffp-contract=off, so identical operation order
gives identical bits; any tolerance here would hide a restatement error).

CPU only; nothing here touches the product path.
"""
import ctypes as C

import numpy as np
import pytest

import helpers
from helpers import capi, worlds, ora

KEYS = ("x", "y", "z", "vx", "vy", "vz", "mass", "volume")
STATE = ("x", "y", "z", "vx", "vy", "vz", "fx", "fy", "fz", "mass")

pytestmark = pytest.mark.skipif(not (ora.refobj_available(2) and ora.refobj_available(3)),
                                reason="oracle/_ref/libgfsrefobj*.so not built (make -C oracle)")

_cache = {}


def setup(kind):
    if kind not in _cache:
        w = helpers.test_world(kind)
        sim, ptrs = helpers.matched_oracle(w)
        _cache[kind] = (w, sim, ptrs)
    return _cache[kind]


def both_lists(sim, parts, par, periodic_mask=0):
    rs = ora.RefSim(sim, periodic_mask)
    rs.configure(par)
    rl = ora.RefParticleList(rs, *[parts[k] for k in KEYS], par)
    pl = ora.ParticleList(sim, *[parts[k] for k in KEYS])
    return rs, rl, pl


def assert_same_state(ref, port, dim, what=""):
    assert len(ref["x"]) == len(port["x"]), what
    for k in STATE:
        if dim == 2 and k in ("z", "vz"):
            continue
        assert np.array_equal(ref[k], port[k]), (what, k, float(np.abs(ref[k] - port[k]).max()))


@pytest.mark.parametrize("kind", ["c1", "uniform3", "ring3", "ring2", "ring3b", "chain2", "chain3"])
def test_particulate_event_bit_identical(kind):
    """gfs_particle_list_event -> gfs_event_list_event -> gfs_event_do ->
    gfs_particulate_event (:804-840) with the world's force list, 5 steps"""
    w, sim, ptrs = setup(kind)
    parts = helpers.test_particles(w, 3000)
    par = helpers.oracle_params(w)
    rs, rl, pl = both_lists(sim, parts, par)
    for step in range(5):
        assert rl.event() == 1
        pl.cull()
        pl.step(par)
        pl.bc(0)
        assert_same_state(rl.get(), pl.get(), w.dim, (kind, step))
    assert np.abs(rl.get()["fx"]).max() > 0
    t, i = rs.time()
    assert i == 5 and abs(t - 5 * w.dt) < 1e-15
    rs.close()


@pytest.mark.parametrize("forces,kw", [
    ((ora.FORCE_DRAG,), {}),
    ((ora.FORCE_LIFT,), {}),
    ((ora.FORCE_BUOY,), {}),
    ((ora.FORCE_BUOY, ora.FORCE_LIFT, ora.FORCE_DRAG), {}),
    ((ora.FORCE_DRAG, ora.FORCE_LIFT), dict(cd_const=0.44, cl_const=0.25)),
    ((ora.FORCE_DRAG,), dict(mu=0.0)),                       # no viscosity: drag returns 0 (:553-554)
])
@pytest.mark.parametrize("kind", ["ring2", "ring3"])
def test_force_models_bit_identical(kind, forces, kw):
    """each GfsForce* function through the vtable its class init installs, and the
    accumulated compute_forces (:737-751); constant coefficient functions go through
    the reference's Rep/Urelp/... cell-variable protocol (:566-575)"""
    w, sim, ptrs = setup(kind)
    parts = helpers.test_particles(w, 1500)
    mu = kw.pop("mu", w.mu)
    par = ora.step_params(w.dt, list(forces), rho=w.rho, mu=mu, g=w.g, **kw)
    rs, rl, pl = both_lists(sim, parts, par)
    if not kw:
        # single models, per unit volume, before any step: against the port's one-force list
        for k, f in enumerate(forces):
            single = ora.step_params(w.dt, [f], rho=w.rho, mu=mu, g=w.g)
            one = ora.ParticleList(sim, *[parts[q] for q in KEYS])
            one.step(single)
            got = one.get()
            for i in (0, 7, 1499):
                fv = rl.force(i, k) * parts["volume"][i]
                want = np.array([got["fx"][i], got["fy"][i], got["fz"][i] if w.dim == 3 else 0.0])
                assert np.array_equal(fv[:w.dim], want[:w.dim]), (f, i, fv, want)
    for step in range(2):
        rl.event()
        pl.step(par)
        assert_same_state(rl.get(), pl.get(), w.dim, (kind, forces, step))
    rs.close()


@pytest.mark.parametrize("kind", ["ring2", "uniform3"])
def test_per_cell_alpha_and_viscosity(kind):
    """PhysicalParams alpha = <variable> and a per-cell viscosity variable"""
    w, sim, ptrs = setup(kind)
    a = w.arrays
    live = (a.flags & capi.CELL_DESTROYED) == 0
    rng = np.random.default_rng(4)
    sim.set_values(3, ptrs[live], rng.uniform(0.5, 2.0, int(live.sum())))
    sim.set_values(4, ptrs[live], rng.uniform(5e-4, 2e-3, int(live.sum())))
    parts = helpers.test_particles(w, 2000)
    par = helpers.oracle_params(w, ivar_alpha=3, ivar_mu=4)
    rs, rl, pl = both_lists(sim, parts, par)
    for step in range(3):
        rl.event()
        pl.step(par)
        assert_same_state(rl.get(), pl.get(), w.dim, (kind, step))
    rs.close()


@pytest.mark.parametrize("kind", ["c1", "ring3"])
@pytest.mark.parametrize("forces,kw", [
    ((ora.FORCE_INERTIAL,), {}),
    ((ora.FORCE_ADDEDMASS, ora.FORCE_DRAG, ora.FORCE_BUOY), {}),
    ((ora.FORCE_DRAG, ora.FORCE_INERTIAL, ora.FORCE_ADDEDMASS, ora.FORCE_LIFT), dict(cm_const=0.3)),
])
def test_inertial_and_added_mass_bit_identical(kind, forces, kw):
    """GfsForceInertial (:255-303), GfsForceAddedMass (:331-394) with its cumulative
    mass update, and the Un,Vn,Wn snapshot the list event takes after the particles
    (store_domain_previous_vel :98-112)"""
    w, sim, ptrs = setup(kind)
    a = w.arrays
    rng = np.random.default_rng(21)
    fields = [w.u, w.v] + ([w.w] if w.dim == 3 else [])
    prev = [0.8 * f + 0.05 * rng.standard_normal(a.n_cells) for f in fields]
    live = (a.flags & capi.CELL_DESTROYED) == 0
    parts = helpers.test_particles(w, 1500)
    par = ora.step_params(w.dt, list(forces), rho=w.rho, mu=w.mu, g=(0.1, -1.0, 0.0), ivar_uold=5, **kw)
    rs, rl, pl = both_lists(sim, parts, par)
    leaves = live & (a.child0 < 0) & ((a.flags & capi.CELL_BOUNDARY) == 0)
    for step in range(3):
        # both run on the same cell data: the port first, because the reference's list
        # event ends by copying U,V,W into Un,Vn,Wn on every leaf
        for c, f in enumerate(prev):
            sim.set_values(5 + c, ptrs[live], f[live])
        pl.step(par)
        rl.event()
        assert_same_state(rl.get(), pl.get(), w.dim, (kind, forces, step))
        # reference quirk: only a GfsForceInertial in the list triggers the snapshot
        # (:1003-1011); GfsForceAddedMass alone keeps the Un,Vn,Wn it was read with
        for c, f in enumerate(fields if ora.FORCE_INERTIAL in forces else prev):
            assert np.array_equal(sim.get_values(5 + c, ptrs[leaves]), f[leaves])
    if ora.FORCE_ADDEDMASS in forces:
        assert np.all(rl.get()["mass"] > parts["mass"])
    rs.close()


@pytest.mark.parametrize("kind", ["c1", "ring3", "chain2"])
def test_tracer_event_bit_identical(kind):
    """no force list: gfs_particulate_event falls through to gfs_particle_event
    (src/particle.c:31-44) and gfs_domain_advect_point"""
    w, sim, ptrs = setup(kind)
    parts = helpers.test_particles(w, 2000)
    par = ora.step_params(w.dt, [])
    rs = ora.RefSim(sim)
    rs.configure(par)
    rl = ora.RefParticleList(rs, *[parts[k] for k in KEYS], par)
    x, y, z = parts["x"], parts["y"], parts["z"]
    for step in range(3):
        rl.event()
        x, y, z = sim.advect_points(x, y, z, w.dt)
        got = rl.get()
        assert np.array_equal(got["x"], x) and np.array_equal(got["y"], y)
        if w.dim == 3:
            assert np.array_equal(got["z"], z)
    rs.close()


def test_cull_removes_particles_outside_the_domain():
    """remove_particles_not_in_domain (:955-969) runs before the children's events"""
    w, sim, ptrs = setup("ring3")
    parts = helpers.test_particles(w, 500)
    parts["x"][::7] = 0.8
    parts["y"][3] = float("nan")
    par = helpers.oracle_params(w)
    rs, rl, pl = both_lists(sim, parts, par)
    assert rl.outside() == len(parts["x"][::7]) + 1
    rl.event()
    removed = pl.cull()
    pl.step(par)
    assert removed == len(parts["x"][::7]) + 1 and len(rl) == len(pl) == 500 - removed
    got, want = rl.get(), pl.get()
    assert np.array_equal(got["id"], want["id"])
    assert_same_state(got, want, 3)
    rs.close()


@pytest.mark.parametrize("dim", [2, 3])
def test_particle_bc_wrap_and_drop(dim):
    """gfs_particle_bc (:3318-3395): periodic wrap through GfsBoundaryPeriodic sides,
    removal at plain boundaries, 25 steps, compared by particle id.  The reference
    takes a wrapped particle out of the list and adds it again
    (gts_container_add PREPENDS), so the two lists differ in order only."""
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
    rs, rl, pl = both_lists(sim, parts, par, mask)
    wrapped = dropped = 0
    for step in range(25):
        rl.event()
        pl.cull()
        pl.step(par)
        before = len(pl)
        dropped += pl.bc(mask)
        got, want = rl.get(), pl.get()
        go, wo = np.argsort(got["id"]), np.argsort(want["id"])
        assert np.array_equal(got["id"][go], want["id"][wo]), step
        for k in ("x", "y", "z", "vx", "vy", "vz")[:2 * 3]:
            if dim == 2 and k in ("z", "vz"):
                continue
            assert np.array_equal(got[k][go], want[k][wo]), (step, k)
        wrapped += int((np.diff(got["id"].astype(np.int64)) < 0).any())
    assert dropped > 20 and wrapped > 5
    rs.close()


@pytest.mark.parametrize("kind", ["c1", "ring3", "chain2"])
def test_particulate_field_event_bit_identical(kind):
    """GfsParticulateField: gfs_cell_reset on the leaves, then V_p/V_cell scattered
    in list order (:1929-1957)"""
    w, sim, ptrs = setup(kind)
    a = w.arrays
    parts = helpers.test_particles(w, 20000)
    par = helpers.oracle_params(w)
    rs, rl, pl = both_lists(sim, parts, par)
    live = (a.flags & capi.CELL_DESTROYED) == 0
    leaves = live & (a.child0 < 0) & ((a.flags & capi.CELL_BOUNDARY) == 0)
    sim.set_values(3, ptrs[live], np.full(int(live.sum()), 7.0))     # the event must reset this
    rl.field_event(3)
    ref = sim.get_values(3, ptrs[leaves])
    sim.set_values(3, ptrs[live], np.zeros(int(live.sum())))
    pl.deposit_volume(3)
    port = sim.get_values(3, ptrs[leaves])
    assert ref.max() > 0 and np.array_equal(ref, port)
    rs.close()


@pytest.mark.parametrize("kind,rk,kernel", [
    ("c1", 0.06, (ora.KERNEL_GAUSSIAN, 1.0, 2e-4, 1)),
    ("ring3", 0.04, (ora.KERNEL_COMPACT, 2.0, 1e-4, 2)),
    ("ring2", 0.05, (ora.KERNEL_CONSTANT, 1.0, 0.0, 1)),
    ("chain2", 0.08, (ora.KERNEL_GAUSSIAN, 0.5, 3e-4, 1)),
    ("uniform3", 0.0, (ora.KERNEL_COMPACT, 1.0, 1e-5, 1)),
])
def test_source_particulate_event_bit_identical(kind, rk, kernel):
    """GfsSourceParticulate with its smoothing kernel (:2087-2228): forces without
    buoyancy, two conditional traversals per particle, the reference's own
    distance_normalization (z offset zeroed before use in 3D)"""
    w, sim, ptrs = setup(kind)
    a = w.arrays
    parts = helpers.test_particles(w, 400)
    if kind != "chain2":
        parts["x"][:3] = 0.7                           # outside: zero force, still traversed
    par = helpers.oracle_params(w)
    rs, rl, pl = both_lists(sim, parts, par)
    live = (a.flags & capi.CELL_DESTROYED) == 0
    leaves = live & (a.child0 < 0) & ((a.flags & capi.CELL_BOUNDARY) == 0)
    k = ora.Kernel(*kernel, 0)
    for iv in range(4, 4 + w.dim):
        sim.set_values(iv, ptrs[live], np.full(int(live.sum()), 3.0))
    rl.source_event(4, rk, k)
    ref = [sim.get_values(4 + c, ptrs[leaves]) for c in range(w.dim)]
    for iv in range(4, 4 + w.dim):
        sim.set_values(iv, ptrs[live], np.zeros(int(live.sum())))
    pl.deposit_force_smoothed(par, 4, rk, k)
    for c in range(w.dim):
        port = sim.get_values(4 + c, ptrs[leaves])
        assert np.abs(ref[c]).max() > 0
        assert np.array_equal(ref[c], port), (c, float(np.abs(ref[c] - port).max()))
    rs.close()


@pytest.mark.parametrize("kind", ["c1", "ring3"])
def test_particle_text_block_byte_identical(kind, tmp_path):
    """gfs_particulate_write (:910-926) over gfs_particle_write (src/particle.c:86-98)"""
    w, sim, ptrs = setup(kind)
    parts = helpers.test_particles(w, 300)
    par = helpers.oracle_params(w)
    rs, rl, pl = both_lists(sim, parts, par)
    for _ in range(3):
        rl.event()
        pl.step(par)
    rl.write(tmp_path / "ref.txt")
    pl.write(tmp_path / "port.txt")
    ref = (tmp_path / "ref.txt").read_bytes()
    port = (tmp_path / "port.txt").read_bytes()
    assert ref.count(b"\n") == 300 and ref.startswith(b"    GfsParticulate 1 ")
    assert ref == port
    rs.close()


import glob
import os

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))


@pytest.mark.parametrize("path", GOLDEN, ids=lambda p: os.path.basename(p)[:-4])
def test_reference_object_code_reproduces_golden(path):
    """the committed fixtures (tests/golden, generated by make_golden.py through the
    restated port) are what the reference's own particulate object code gives"""
    g = np.load(path)
    name = os.path.basename(path)[:-4]
    w = {"c1_l5": lambda: worlds.make_c1(level=5, n_particles=400),
         "tg_l4": lambda: worlds.make_c2(level=4, n_particles=400),
         "ring_3_6": lambda: worlds.make_ring("ring", 3, 6, 400, 3003)}[name]()
    sim, ptrs = helpers.matched_oracle(w)
    parts = {k: (g["p_" + k] if len(g["p_" + k]) else None) for k in KEYS}
    par = helpers.oracle_params(w)
    rs = ora.RefSim(sim)
    rs.configure(par)
    rl = ora.RefParticleList(rs, *[parts[k] for k in KEYS], par)
    done = 0
    for steps in (1, 10):
        rl.event(steps - done)
        done = steps
        st = rl.get()
        for k in ("x", "y", "vx", "vy", "fx", "fy"):
            assert np.array_equal(st[k], g[f"s{steps}_{k}"]), (steps, k)
    rs.close()


@pytest.mark.skipif(not ora.refobj_available(3, module=True), reason="libgfsrefmod not built")
def test_dropin_module_loads_and_hands_back_what_the_device_cannot_do(monkeypatch):
    """libgfsrefmod = the reference objects + the drop-in GModule source
    (host/particulates_b200.c) linked as a Gerris installation would.  Its
    g_module_check_init() instantiates the 16 classes and re-points the three
    hot-path events.  A simulation with MOVING solids (fractions change every step
    without an adapt) is not expressible on the device: with GFSB200_ALLOW_REFERENCE_EVENT=1 the
    module hands the event back to the reference's own method, untouched -- checked here without a
    GPU against the unmodified library (without the opt-in it stops the run: next test)."""
    monkeypatch.setenv("GFSB200_ALLOW_REFERENCE_EVENT", "1")
    w, sim, ptrs = setup("ring3")
    parts = helpers.test_particles(w, 800)
    par = helpers.oracle_params(w)
    states = []
    for module in (True, False):               # one RefSim at a time: the GfsBox objects hang on the roots
        rs = ora.RefSim(sim, module=module)
        assert rs.R.refobj_module_name() == (b"particulates" if module else None)
        rs.configure(par)
        rs.add_solid(moving=True)
        rl = ora.RefParticleList(rs, *[parts[k] for k in KEYS], par)
        states.append([])
        for step in range(2):
            assert rl.event() == 1
            states[-1].append(rl.get())
        rs.close()
    for step in range(2):
        assert_same_state(states[0][step], states[1][step], 3, step)


_NO_FALLBACK = r"""
import sys
sys.path.insert(0, %(tests)r)
import helpers
from helpers import ora
KEYS = ("x", "y", "z", "vx", "vy", "vz", "mass", "volume")
w = helpers.test_world("ring3")
sim, ptrs = helpers.matched_oracle(w)
parts = helpers.test_particles(w, 50)
par = helpers.oracle_params(w)
rs = ora.RefSim(sim, module=True)
rs.configure(par)
rs.add_solid(moving=True)
rl = ora.RefParticleList(rs, *[parts[k] for k in KEYS], par)
print("before the event", flush=True)
rl.event()
print("the event returned", flush=True)
"""


@pytest.mark.skipif(not ora.refobj_available(3, module=True), reason="libgfsrefmod not built")
def test_dropin_module_has_no_silent_cpu_fallback():
    """the default: a list the device cannot run stops the simulation (g_error -> abort) with a
    message naming the object and the reason; nothing runs on the CPU behind the user's back"""
    import subprocess
    import sys
    env = {k: v for k, v in os.environ.items() if k != "GFSB200_ALLOW_REFERENCE_EVENT"}
    r = subprocess.run([sys.executable, "-c", _NO_FALLBACK % {"tests": os.path.dirname(os.path.abspath(__file__))}],
                       capture_output=True, text=True, env=env, timeout=300)
    assert "before the event" in r.stdout and "the event returned" not in r.stdout
    assert r.returncode != 0
    assert "GfsParticleList" in r.stderr and "moving solids" in r.stderr
    assert "no silent CPU fallback" in r.stderr and "GFSB200_ALLOW_REFERENCE_EVENT" in r.stderr


@pytest.mark.parametrize("kind", ["c1", "uniform3", "ring3", "ring2", "ring3b", "chain2", "chain3", "periodic2",
                                  "periodic3"])
def test_locate_array_and_domain_locate_bit_identical(kind):
    """GfsLocateArray (src/domain.c:43-145) and gfs_domain_locate (:2623-2638), the
    reference's own object code, against the restatement the GPU locate is held to:
    same array geometry, same FttCell* for random points, points on cell faces /
    vertices / +-1 ulp around them, the hull, outside, NaN"""
    if kind.startswith("periodic"):
        w, mask = helpers.periodic_world(int(kind[-1]))
        sim, ptrs = helpers.matched_oracle(w)
    else:
        w, sim, ptrs = setup(kind)
        mask = 0
    rs = ora.RefSim(sim, mask)
    mn, h, n = rs.locate_array()
    pmn, ph, pn = np.zeros(3), C.c_double(), np.zeros(3, dtype=np.int32)
    sim.L.ora_locate_array(sim.h, pmn.ctypes.data, C.byref(ph), pn.ctypes.data)
    assert np.array_equal(mn, pmn[:w.dim]) and h == ph.value and np.array_equal(n, pn[:w.dim])
    rng = np.random.default_rng(3)
    nbox = w.meta.get("nbox", 1)
    m = 40000
    pts = rng.uniform(-0.75, nbox - 0.25, (m, 3))
    pts[:, 1:] = rng.uniform(-0.75, 0.75, (m, 2))
    ax, ay, az = worlds.adversarial_points(w.arrays, rng, 4000)
    adv = np.stack([ax, ay, az if az is not None else np.zeros_like(ax)], 1)
    pts = np.vstack([pts, adv, [[np.nan, 0, 0], [0, np.inf, 0], [1e300, 0, 0]]])
    z = pts[:, 2].copy() if w.dim == 3 else None
    ref = rs.locate(pts[:, 0].copy(), pts[:, 1].copy(), z)
    port = sim.locate(pts[:, 0].copy(), pts[:, 1].copy(), z)
    assert (ref != 0).sum() > 1000 and (ref == 0).sum() > 1000
    assert np.array_equal(ref, port)
    rs.close()

"""GPU parity tests: every call goes through the C-ABI (lib/libgfsb200.so) and
is compared with the oracle (reference ftt.c/fluid.c object code + restated
particulate layer) on the same seeded inputs.

Bars (BASELINE.json north_star):
  * cell indices from point location: bit-exact
  * vorticity table: bit-exact (cell pass is compiled --fmad=false)
  * corner values: bit-exact for the canonical leaf of each vertex, <= 4 ulp of
    max|field| for the other leaves sharing it (summation order)
  * particle positions / velocities after one step: relative error <= 1e-12
"""
import numpy as np
import pytest

import helpers
from helpers import capi, worlds, ora

pytestmark = pytest.mark.gpu

STATE = ("x", "y", "z", "vx", "vy", "vz")
RTOL_STEP = 1e-12          # north_star: <= 1e-12 relative after one step


@pytest.fixture(scope="module")
def ctx():
    c = capi.Context(0)
    yield c
    c.close()


_world, _particles = helpers.test_world, helpers.test_particles


_cache = {}


def setup(kind, ctx):
    if kind not in _cache:
        w = _world(kind)
        sim, ptrs = helpers.matched_oracle(w)
        _cache[kind] = (w, sim, ptrs, helpers.PtrIndex(ptrs))
    w, sim, ptrs, idx = _cache[kind]
    ctx.upload_tree(w.tree)
    ctx.upload_field(w.u, w.v, w.w)
    return w, sim, ptrs, idx


KINDS = ["c1", "uniform3", "ring3", "ring2", "ring3b", "chain2", "chain3"]


@pytest.mark.parametrize("kind", KINDS)
def test_locate_bit_exact(kind, ctx):
    w, sim, ptrs, idx = setup(kind, ctx)
    rng = np.random.default_rng(5)
    n = 200_000
    cols = [rng.uniform(-0.6, 0.6, n) for _ in range(w.dim)] + ([None] if w.dim == 2 else [])
    cols[0] = rng.uniform(-0.6, w.meta.get("nbox", 1) - 0.4, n)
    adv = worlds.adversarial_points(w.arrays, rng, 20000)
    for x, y, z in (cols, adv):
        got = ctx.locate(x, y, z)
        want = idx(sim.locate(x, y, z))
        assert np.array_equal(got, want), f"{(got != want).sum()} of {len(got)} cell indices differ"
        assert (got < 0).any() and (got >= 0).any()


@pytest.mark.parametrize("kind", KINDS)
def test_vorticity_bit_exact(kind, ctx):
    w, sim, ptrs, idx = setup(kind, ctx)
    leaves = w.arrays.box_leaves
    got = ctx.vorticity(leaves)
    want = sim.vorticity(ptrs[leaves])
    assert np.array_equal(got, want), f"max diff {np.abs(got - want).max():.3e}"


@pytest.mark.parametrize("kind", KINDS)
def test_corner_values(kind, ctx):
    w, sim, ptrs, idx = setup(kind, ctx)
    leaves = w.arrays.box_leaves
    fields = [w.u, w.v] + ([w.w] if w.dim == 3 else [])
    for comp, f in enumerate(fields):
        got = ctx.corner_values(comp, leaves)
        want = sim.corner_values(comp, ptrs[leaves])
        tol = 4 * np.finfo(np.float64).eps * max(np.abs(f).max(), 1e-300)
        assert np.abs(got - want).max() <= tol
        # the canonical (first-seen) leaf of every vertex is bit-exact
        lv = w.arrays.leaf_vtx[leaves]
        first = np.zeros(lv.shape, dtype=bool)
        _, pos = np.unique(lv.ravel(), return_index=True)
        first.ravel()[pos] = True
        assert np.array_equal(got[first], want[first])


@pytest.mark.parametrize("kind", KINDS)
def test_interpolate(kind, ctx):
    w, sim, ptrs, idx = setup(kind, ctx)
    rng = np.random.default_rng(9)
    n = 50_000
    cols = [rng.uniform(-0.5, 0.5, n) for _ in range(w.dim)] + ([None] if w.dim == 2 else [])
    cols[0] = rng.uniform(-0.5, w.meta.get("nbox", 1) - 0.5, n)
    got = ctx.interpolate(*cols)
    fields = [w.u, w.v] + ([w.w] if w.dim == 3 else [])
    for comp, f in enumerate(fields):
        want = sim.interpolate(comp, *cols)
        inside = want != capi.NODATA
        assert np.array_equal(got[comp] == capi.NODATA, ~inside)
        scale = np.abs(f).max()
        assert np.abs(got[comp][inside] - want[inside]).max() <= 1e-14 * scale


@pytest.mark.parametrize("kind", ["uniform3", "ring3"])
def test_interpolate_with_nodata_cells(kind, ctx):
    """A field with GFS_NODATA cells: every vertex whose stencil touches one falls back to the
    calling leaf's own value (gfs_cell_corner_value, src/fluid.c:3094-3097) -- the out-of-line
    flavour of the trilinear form (trilinear_rows_nodata); a point inside a NODATA cell returns
    NODATA (gfs_interpolate :2704-2705).  Uniform (lattice) and adaptive tree, against the oracle;
    then one fused step of particles kept out of the NODATA cells, against the oracle."""
    import copy
    w = copy.copy(_world(kind))
    rng = np.random.default_rng(17)
    leaves = w.arrays.box_leaves
    bad = rng.choice(leaves, max(8, len(leaves) // 200), replace=False)
    w.u, w.v, w.w = w.u.copy(), w.v.copy(), w.w.copy()
    w.u[bad[::3]] = capi.NODATA
    w.v[bad[1::3]] = capi.NODATA
    w.w[bad[2::3]] = capi.NODATA
    sim, ptrs = helpers.matched_oracle(w)
    ctx.upload_tree(w.tree)
    ctx.upload_field(w.u, w.v, w.w)
    n = 50_000
    cols = [rng.uniform(-0.5, 0.5, n) for _ in range(3)]
    got = ctx.interpolate(*cols)
    touched = 0
    for comp, f in enumerate((w.u, w.v, w.w)):
        want = sim.interpolate(comp, *cols)
        data = want != capi.NODATA
        assert np.array_equal(got[comp] == capi.NODATA, ~data)
        touched += int((~data).sum())
        scale = np.abs(f[f != capi.NODATA]).max()
        assert np.abs(got[comp][data] - want[data]).max() <= 1e-14 * scale
    assert touched > 0
    # the fused step (lattice kernel on uniform3) with the fallback active
    parts = _particles(w, 20_000)
    cells = ctx.locate(parts["x"], parts["y"], parts["z"])
    # (a NODATA neighbour poisons the centred differences behind the lift force -- in the reference
    # too: keep the particles whose leaf has a finite vorticity)
    uc = np.unique(cells)
    vort = np.asarray(ctx.vorticity(uc)).reshape(len(uc), -1)
    poisoned = uc[~np.all(np.abs(vort) < 1e100, axis=1)]
    keep = ~np.isin(cells, bad) & ~np.isin(cells, poisoned)
    assert keep.sum() > 1000
    parts = {k: v[keep] for k, v in parts.items()}
    got = _run_step(ctx, w, parts)
    ocells, want = helpers.oracle_step(sim, ptrs, w, parts)
    assert np.array_equal(got["cell"], ocells)
    _check_state(got, want, 3)


def _run_step(ctx, w, parts, **kw):
    ctx.particles_upload(**parts)
    ctx.step(w.step_params(record_cells=True, record_forces=True, **kw))
    return ctx.particles_download(forces=True, cells=True)


def _vec_keys(dim):
    return (("x", "y", "z"), ("vx", "vy", "vz")) if dim == 3 else (("x", "y"), ("vx", "vy"))


def _check_state(got, want, dim, rtol=RTOL_STEP, sel=None):
    """|dx|/|x| and |dv|/|v| per particle (2-norm of the vector) <= rtol"""
    for keys in _vec_keys(dim):
        err = helpers.vec_rel_err(got, want, keys, sel)
        assert err <= rtol, f"{keys}: relative error {err:.3e} > {rtol:g}"


@pytest.mark.parametrize("kind", KINDS)
def test_step_one(kind, ctx):
    """cells bit-exact; x, v within 1e-12 relative; forces within 1e-11 of the
    force scale (they are differences of O(1) velocities)"""
    w, sim, ptrs, idx = setup(kind, ctx)
    parts = _particles(w)
    got = _run_step(ctx, w, parts)
    cells, want = helpers.oracle_step(sim, ptrs, w, parts)
    assert np.array_equal(got["cell"], cells)
    _check_state(got, want, w.dim)
    for k in ("fx", "fy", "fz"):
        scale = max(np.abs(want[k]).max(), 1e-300)
        assert np.abs(got[k] - want[k]).max() <= 1e-11 * scale


@pytest.mark.parametrize("kind", KINDS)
def test_step_against_reference_object_code(kind, ctx):
    """the CUDA step against the reference's OWN object code (libgfsrefobj:
    particulatecommon.c + event.c + particle.c + fluid.c + ftt.c compiled unmodified)
    driving a GfsParticleList through gfs_event_do -- no restated layer in between;
    3 steps, same bars as against the restated oracle"""
    if not ora.refobj_available(2 if kind in ("c1", "ring2", "chain2") else 3):
        pytest.skip("oracle/_ref/libgfsrefobj*.so not built")
    w, sim, ptrs, idx = setup(kind, ctx)
    parts = _particles(w, 5000)
    ctx.particles_upload(**parts)
    opar = helpers.oracle_params(w)
    rs = ora.RefSim(sim)
    rs.configure(opar)
    rl = ora.RefParticleList(rs, *[parts[k] for k in ("x", "y", "z", "vx", "vy", "vz", "mass", "volume")], opar)
    par = w.step_params(record_cells=True, record_forces=True)
    for step in range(3):
        s = rl.get()
        ocell = idx(sim.locate(s["x"], s["y"], s["z"] if w.dim == 3 else None))
        assert rl.event() == 1
        ctx.step(par)
        got = ctx.particles_download(cells=True, forces=True)
        want = rl.get()
        assert len(want["x"]) == len(got["x"])
        if step == 0:
            assert np.array_equal(got["cell"], ocell)
            same = np.ones(len(ocell), dtype=bool)
        same &= got["cell"] == ocell          # a cell flipping one step early is counted, not bounded
        assert same.mean() > 0.999
        _check_state(got, want, w.dim, rtol=RTOL_STEP if step == 0 else 1e-11, sel=same)
        for k in ("fx", "fy", "fz")[:w.dim]:
            scale = max(np.abs(want[k]).max(), 1e-300)
            assert np.abs(got[k] - want[k])[same].max() <= 1e-11 * scale, (step, k)
    rs.close()


@pytest.mark.parametrize("forces,kw", [
    ((capi.FORCE_BUOY,), {}),
    ((capi.FORCE_BUOY, capi.FORCE_DRAG, capi.FORCE_LIFT), {}),
    ((capi.FORCE_DRAG,), dict(cd_const=0.44)),
    ((capi.FORCE_LIFT, capi.FORCE_DRAG), dict(cl_const=0.25)),
    ((capi.FORCE_DRAG, capi.FORCE_DRAG), {}),
])
def test_step_force_lists(forces, kw, ctx):
    w, sim, ptrs, idx = setup("ring3", ctx)
    w2 = worlds.World(**{**w.__dict__, "forces": forces})
    parts = worlds.make_particles(w2, 5000)
    got = _run_step(ctx, w2, parts, **kw)
    cells, want = helpers.oracle_step(sim, ptrs, w2, parts, **kw)
    assert np.array_equal(got["cell"], cells)
    _check_state(got, want, 3)


def test_step_zero_viscosity_and_stokes(ctx):
    """mu = 0: GfsForceDrag returns 0 (:560-561); tiny relative velocity: Re < 1e-8 => 0"""
    w, sim, ptrs, idx = setup("uniform3", ctx)
    w0 = worlds.World(**{**w.__dict__, "mu": 0.0, "forces": (capi.FORCE_DRAG,)})
    parts = worlds.make_particles(w0, 2000)
    got = _run_step(ctx, w0, parts)
    assert np.all(got["fx"] == 0) and np.all(got["fy"] == 0) and np.all(got["fz"] == 0)
    cells, want = helpers.oracle_step(sim, ptrs, w0, parts)
    _check_state(got, want, 3)


def test_step_per_cell_alpha_mu(ctx):
    """sim->physical_params.alpha and a per-cell viscosity variable"""
    w, sim, ptrs, idx = setup("ring3", ctx)
    rng = np.random.default_rng(3)
    alpha = rng.uniform(0.5, 2.0, w.arrays.n_cells)
    mu = rng.uniform(5e-4, 2e-3, w.arrays.n_cells)
    ctx.upload_field(w.u, w.v, w.w, alpha=alpha, mu=mu)
    live = (w.arrays.flags & capi.CELL_DESTROYED) == 0
    sim.set_values(3, ptrs[live], alpha[live])
    sim.set_values(4, ptrs[live], mu[live])
    parts = worlds.make_particles(w, 5000)
    got = _run_step(ctx, w, parts)
    cells, want = helpers.oracle_step(sim, ptrs, w, parts, ivar_alpha=3, ivar_mu=4)
    assert np.array_equal(got["cell"], cells)
    _check_state(got, want, 3)
    ctx.upload_field(w.u, w.v, w.w)


@pytest.mark.parametrize("kind", ["c1", "ring3"])
def test_drift_after_n_steps(kind, ctx):
    """Documented drift bound: particles whose cell history matches the
    reference's stay within 1e-12 * 4^(steps/10) ... in practice ~1e-13 after
    100 steps; the lift force is cell-constant, so a particle whose cell flips
    one step early may differ by O(dt * |dF|) -- those are counted, not bounded."""
    w, sim, ptrs, idx = setup(kind, ctx)
    parts = worlds.make_particles(w, 2000)
    ctx.particles_upload(**parts)
    par = w.step_params(record_cells=True)
    plist = ora.ParticleList(sim, *[parts[k] for k in ("x", "y", "z", "vx", "vy", "vz", "mass", "volume")])
    opar = helpers.oracle_params(w)
    same = np.ones(len(parts["x"]), dtype=bool)
    for step in range(1, 101):
        s = plist.get()
        ocell = idx(sim.locate(s["x"], s["y"], s["z"]))
        plist.step(opar)
        ctx.step(par)
        if step in (1, 10, 100):
            got = ctx.particles_download(cells=True)
            same &= got["cell"] == ocell
            want = plist.get()
            errs = [helpers.vec_rel_err(got, want, keys, same) for keys in _vec_keys(w.dim)]
            bound = {1: 1e-12, 10: 1e-11, 100: 1e-9}[step]
            assert max(errs) <= bound, f"step {step}: drift {max(errs):.3e} > {bound:g}"
            assert same.mean() > 0.98, f"step {step}: only {same.mean():.3f} of cell histories match"


def test_outside_particles_untouched_and_cull(ctx):
    """gfs_particle_list_event: particles with locate == NULL are removed
    (remove_particles_not_in_domain) before the children run"""
    w, sim, ptrs, idx = setup("uniform3", ctx)
    parts = worlds.make_particles(w, 3000)
    rng = np.random.default_rng(1)
    out = rng.random(3000) < 0.2
    parts["x"] = np.where(out, parts["x"] + 1.0, parts["x"])
    ctx.particles_upload(**parts)
    ctx.step(w.step_params(record_cells=True))
    got = ctx.particles_download(cells=True)
    assert np.all(got["cell"][out] == -1)
    for k in STATE:
        assert np.array_equal(got[k][out], parts[k][out])
    # the list event: cull, then step
    ctx.particles_upload(**parts)
    removed = ctx.particle_list_event(w.step_params(record_cells=True))
    assert removed == int(out.sum()) and ctx.count == 3000 - removed
    got = ctx.particles_download(cells=True, ids=True)
    assert np.array_equal(got["id"], np.nonzero(~out)[0] + 1)      # list order preserved
    plist = ora.ParticleList(sim, *[parts[k] for k in ("x", "y", "z", "vx", "vy", "vz", "mass", "volume")])
    assert plist.cull() == removed
    plist.step(helpers.oracle_params(w))
    want = plist.get()
    assert np.array_equal(got["id"], want["id"])
    _check_state(got, want, 3)


def test_sort_is_a_permutation_and_step_invariant(ctx):
    w, sim, ptrs, idx = setup("ring3", ctx)
    parts = worlds.make_particles(w, 20000)
    ref = _run_step(ctx, w, parts)
    ctx.particles_upload(**parts)
    ctx.sort()
    pre = ctx.particles_download(ids=True)
    perm = pre["id"].astype(np.int64) - 1
    assert np.array_equal(np.sort(perm), np.arange(20000))
    for k in STATE:
        assert np.array_equal(pre[k], parts[k][perm])
    cells = ctx.locate(pre["x"], pre["y"], pre["z"])
    assert np.all(np.diff(cells) >= 0), "particles are not sorted by cell"
    ctx.step(w.step_params(record_cells=True))
    got = ctx.particles_download(cells=True)
    for k in STATE:
        assert np.array_equal(got[k], ref[k][perm])      # bitwise: the step does not depend on order
    assert np.array_equal(got["cell"], ref["cell"][perm])


@pytest.mark.parametrize("kind", ["c1", "ring3"])
def test_tracer_advection(kind, ctx):
    """no forces attached: gfs_domain_advect_point (RK2), src/domain.c:2764-2788"""
    w, sim, ptrs, idx = setup(kind, ctx)
    parts = worlds.make_particles(w, 5000)
    ctx.particles_upload(**parts)
    ctx.step(capi.StepParams(w.dt, ()))
    got = ctx.particles_download()
    x, y, z = sim.advect_points(parts["x"], parts["y"], parts["z"], w.dt)
    keys = ("x", "y", "z")[:w.dim]
    assert helpers.vec_rel_err(got, dict(x=x, y=y, z=z), keys) <= RTOL_STEP


@pytest.mark.parametrize("kind", ["c1", "ring3"])
def test_deposit(kind, ctx):
    """GfsParticulateField (void fraction) and the single-cell force deposit.
    Summation order differs (atomics): abs tol 1e-12 * max|field|."""
    w, sim, ptrs, idx = setup(kind, ctx)
    n = 40000
    parts = worlds.make_particles(w, n)
    ctx.particles_upload(**parts)
    ctx.sort()
    ctx.deposit_volume()
    ctx.deposit_force(w.step_params())
    separate = [ctx.download_deposit(c) for c in range(1 + w.dim)]
    ctx.deposit_all(w.step_params())              # the fused pass gives the same field
    for c in range(1 + w.dim):
        fused = ctx.download_deposit(c)
        assert np.abs(fused - separate[c]).max() <= 1e-13 * max(np.abs(separate[c]).max(), 1e-300)
    live = (w.arrays.flags & capi.CELL_DESTROYED) == 0
    plist = ora.ParticleList(sim, *[parts[k] for k in ("x", "y", "z", "vx", "vy", "vz", "mass", "volume")])
    zero = np.zeros(int(live.sum()))
    for iv in range(3, 3 + 1 + w.dim):
        sim.set_values(iv, ptrs[live], zero)
    plist.deposit_volume(3)
    plist.deposit_force(helpers.oracle_params(w), 4)
    for comp in range(1 + w.dim):
        got = ctx.download_deposit(comp)[live]
        want = sim.get_values(3 + comp, ptrs[live])
        assert np.abs(want).max() > 0
        assert np.abs(got - want).max() <= 1e-12 * np.abs(want).max(), comp
    # conservation: sum(field * V_cell) = sum(V_p) over located particles
    a = w.arrays
    vol = ctx.download_deposit(0)
    inside = ctx.locate(parts["x"], parts["y"], parts["z"]) >= 0
    total = np.sum(vol * a.h ** w.dim)
    assert abs(total - parts["volume"][inside].sum()) <= 1e-12 * parts["volume"].sum()


@pytest.mark.parametrize("kind,rk,kernel,flags", [
    ("c1", 0.06, (capi.KERNEL_GAUSSIAN, 1.0, 2e-4, 1), 0),
    ("ring3", 0.04, (capi.KERNEL_COMPACT, 2.0, 1e-4, 2), 0),
    ("ring3", 0.03, (capi.KERNEL_GAUSSIAN, 1.0, 1e-4, 1), capi.KERNEL_FIX_Z),
    ("ring2", 0.05, (capi.KERNEL_CONSTANT, 1.0, 0.0, 1), 0),
    ("chain2", 0.08, (capi.KERNEL_GAUSSIAN, 0.5, 3e-4, 1), 0),
    ("uniform3", 0.0, (capi.KERNEL_COMPACT, 1.0, 1e-5, 1), 0),
])
def test_deposit_smoothed(kind, rk, kernel, flags, ctx):
    """GfsSourceParticulate with its smoothing kernel (source_particulate_event,
    modules/particulatecommon.c:2087-2228).  The set of leaves each particle reaches
    is bit-exact (its exact volume sum); the correction is summed in the reference's
    traversal order and agrees to 1e-13 (device pow/exp vs libm); the deposited field sums
    over particles in a different order (atomics): abs tol 1e-12 * max|field|."""
    w, sim, ptrs, idx = setup(kind, ctx)
    parts = _particles(w, 3000)
    if kind != "chain2":                       # a few particles outside the domain (zero force)
        parts["x"][:5] = 0.7
    ctx.particles_upload(**parts)
    kk, ka, kb, kp = kernel
    ctx.deposit_force_smoothed(w.step_params(), rk, kk, ka, kb, kp, flags, record_norm=True)
    corr, vol = ctx.download_kernel_norm()
    live = (w.arrays.flags & capi.CELL_DESTROYED) == 0
    plist = ora.ParticleList(sim, *[parts[k] for k in ("x", "y", "z", "vx", "vy", "vz", "mass", "volume")])
    zero = np.zeros(int(live.sum()))
    for iv in range(4, 4 + w.dim):
        sim.set_values(iv, ptrs[live], zero)
    wcorr, wvol = plist.deposit_force_smoothed(helpers.oracle_params(w), 4, rk, ora.Kernel(kk, ka, kb, kp, flags))
    assert np.array_equal(vol, wvol)
    # r_b = pow (3V/4pi, 1/3) and exp differ from libm by <= 1-2 ulp, which moves every kernel value
    # by a few ulp: the correction agrees to 1e-13, not bitwise (a constant kernel is exact)
    ok = np.isfinite(wcorr)
    assert np.array_equal(ok, np.isfinite(corr))           # 0/0 for particles outside the domain
    if kk == capi.KERNEL_CONSTANT:
        assert np.array_equal(corr[ok], wcorr[ok])
    else:
        assert np.abs(corr[ok] - wcorr[ok]).max() <= 1e-13 * np.abs(wcorr[ok]).max()
    assert (wcorr > 1e-10).any()
    for comp in range(w.dim):
        got = ctx.download_deposit(1 + comp)[live]
        want = sim.get_values(4 + comp, ptrs[live])
        assert np.abs(want).max() > 0
        assert np.abs(got - want).max() <= 1e-12 * np.abs(want).max(), comp
    # without record_norm the kernel skips force-free particles; same field
    ctx.deposit_force_smoothed(w.step_params(), rk, kk, ka, kb, kp, flags)
    for comp in range(w.dim):
        again = ctx.download_deposit(1 + comp)[live]
        want = sim.get_values(4 + comp, ptrs[live])
        assert np.abs(again - want).max() <= 1e-12 * np.abs(want).max(), comp


def test_deposit_smoothed_rejects_unknown_kernel(ctx):
    w, sim, ptrs, idx = setup("c1", ctx)
    ctx.particles_upload(**_particles(w, 10))
    with pytest.raises(capi.GfsB200Error):
        ctx.deposit_force_smoothed(w.step_params(), 0.1, kind=7)


@pytest.mark.parametrize("kind", ["c1", "ring3", "chain2", "uniform3"])
def test_output_location_any_variable(kind, ctx):
    """GfsOutputLocation (src/output.c:1153-1212) for arbitrary cell variables: four variables
    (more than one pass of three), interpolated and as cell values; values within 4 ulp of
    max|field| of gfs_interpolate (vertex sums shared between leaves), cells bit-exact."""
    w, sim, ptrs, idx = setup(kind, ctx)
    a = w.arrays
    rng = np.random.default_rng(99)
    live = (a.flags & capi.CELL_DESTROYED) == 0
    variables = [np.sin(3 * a.pos[:, 0]) + a.pos[:, 1] ** 2, np.cos(2 * a.pos[:, 1]) - a.pos[:, 0],
                 a.pos[:, 0] * a.pos[:, 1] + 0.3 * a.pos[:, 2], rng.standard_normal(a.n_cells)]
    for k, v in enumerate(variables):
        sim.set_values(3 + k, ptrs[live], v[live])
    parts = _particles(w, 5000)
    x, y, z = parts["x"], parts["y"], parts["z"]
    x[:7] = 0.9 if kind != "chain2" else 9.0                 # outside the domain
    got, cell = ctx.output_location(variables, x, y, z)
    want_cell = idx(sim.locate(x, y, z))
    assert np.array_equal(cell, want_cell)
    inside = want_cell >= 0
    assert (~inside).sum() >= 7 and (got[:, ~inside] == 1.7976931348623157e308).all()
    for k, v in enumerate(variables):
        want = sim.interpolate(3 + k, x, y, z)
        tol = 4 * np.finfo(np.float64).eps * np.abs(v).max()
        assert np.abs(got[k][inside] - want[inside]).max() <= tol, k
    got0, _ = ctx.output_location(variables, x, y, z, interpolate=False)
    for k, v in enumerate(variables):
        assert np.array_equal(got0[k][inside], v[want_cell[inside]])
    # the resident U,V,W tables are untouched
    u, v_, w_ = ctx.interpolate(x, y, z)
    assert np.abs(u[inside] - sim.interpolate(0, x, y, z)[inside]).max() <= 4e-16 * max(np.abs(w.u).max(), 1)


@pytest.mark.parametrize("kind", ["c1", "ring3"])
def test_particle_text_format_and_checkpoint(kind, ctx, tmp_path):
    """The particle block written from the device state is byte-identical to the reference's
    gfs_particle_write + gfs_particulate_write of the oracle's state after the same steps
    (where the two states agree to the 6 digits %g keeps), and the binary checkpoint restores
    every column bit-exactly."""
    w, sim, ptrs, idx = setup(kind, ctx)
    parts = _particles(w, 300)
    ctx.particles_upload(**parts)
    par = w.step_params(record_forces=True)
    plist = ora.ParticleList(sim, *[parts[k] for k in ("x", "y", "z", "vx", "vy", "vz", "mass", "volume")])
    for _ in range(3):
        ctx.step(par)
        plist.step(helpers.oracle_params(w))
    ctx.write_gfs(tmp_path / "dev.txt", L=1.0)
    plist.write(tmp_path / "ref.txt", L=1.0)
    dev, ref = (tmp_path / "dev.txt").read_text().splitlines(), (tmp_path / "ref.txt").read_text().splitlines()
    assert len(dev) == len(ref) == 300
    assert dev[0].startswith("    GfsParticulate 1 ") and len(dev[0].split()) == 13
    same = sum(d == r for d, r in zip(dev, ref))
    assert same >= 290, same          # a 1e-13 difference flips a printed 6th digit only rarely
    for d, r in zip(dev, ref):
        assert np.allclose([float(t) for t in d.split()[1:]], [float(t) for t in r.split()[1:]], rtol=2e-5, atol=1e-300)
    # L scales the printed volume by L^dim
    ctx.write_gfs(tmp_path / "dev2.txt", L=2.0)
    v1 = float(dev[5].split()[6]); v2 = float((tmp_path / "dev2.txt").read_text().splitlines()[5].split()[6])
    assert abs(v2 / v1 - 2.0 ** w.dim) < 1e-5
    # lossless restart
    before = ctx.particles_download(forces=True)
    ctx.checkpoint_save(tmp_path / "ckpt.bin")
    ctx.particles_upload(**_particles(w, 10))
    ctx.checkpoint_load(tmp_path / "ckpt.bin")
    after = ctx.particles_download(forces=True)
    assert ctx.count == 300
    for k in before:
        if before[k] is not None:
            assert np.array_equal(before[k], after[k]), k
    with pytest.raises(capi.GfsB200Error):
        ctx.checkpoint_load(tmp_path / "dev.txt")


@pytest.mark.parametrize("level", [5, 6])
def test_lattice_cell_pass_equals_table_driven_kernels(level, ctx, monkeypatch):
    """lattice_cell_pass_kernel (brick-tiled, arithmetic addressing, shared-memory staging) gives
    bit-identical vertex and vorticity tables to the table-driven kernels on every leaf and
    corner of a 32^3 / 64^3 tree -- interior and hull -- also with GFS_NODATA cells present."""
    w = worlds.make_c2(level=level, n_particles=10)
    a = w.arrays
    leaves = a.box_leaves
    rng = np.random.default_rng(level)
    u, v, wz = w.u.copy(), w.v.copy(), 0.3 * w.u + 0.1 * rng.standard_normal(a.n_cells)
    for nodata in (False, True):
        if nodata:                               # a few NODATA cells, inside and on the hull
            bad = rng.choice(leaves, 40, replace=False)
            u[bad[:20]] = 1.7976931348623157e308
            wz[bad[20:]] = 1.7976931348623157e308
        tables = []
        for env in ("", "1"):
            if env:
                monkeypatch.setenv("GFSB200_NO_LATTICE_PATTERN", env)
            else:
                monkeypatch.delenv("GFSB200_NO_LATTICE_PATTERN", raising=False)
            ctx.upload_tree(w.tree)
            ctx.upload_field(u, v, wz)
            tables.append((ctx.vorticity(leaves), [ctx.corner_values(c, leaves) for c in range(3)]))
        monkeypatch.delenv("GFSB200_NO_LATTICE_PATTERN", raising=False)
        (vort_f, corner_f), (vort_t, corner_t) = tables
        if not nodata:                           # (NODATA poisons the differences: compare bit patterns)
            assert np.array_equal(vort_f, vort_t)
        assert np.array_equal(vort_f.view(np.uint64), vort_t.view(np.uint64))
        for c in range(3):
            assert np.array_equal(corner_f[c].view(np.uint64), corner_t[c].view(np.uint64)), c


def test_kernel_timer_sampling(ctx):
    """gfsb200_timer_sampling: the library's kernel-time events around every n-th launch only
    (1: all, 0: none); a reset starts a new sample with a timed launch."""
    w, sim, ptrs, idx = setup("uniform3", ctx)
    ctx.particles_upload(**_particles(w, 5000))
    par = w.step_params()
    try:
        for every, want in ((1, 8), (4, 2), (3, 3), (0, 0)):
            ctx.timer_sampling(every)
            ctx.timer_reset()
            for _ in range(8):
                ctx.step(par)
            ms, n = ctx.timer_read()
            assert n == want, (every, n)
            assert (ms > 0) == (want > 0)
    finally:
        ctx.timer_sampling(1)


def test_empty_and_single_particle(ctx):
    w, sim, ptrs, idx = setup("uniform3", ctx)
    empty = {k: np.zeros(0) for k in ("x", "y", "z", "vx", "vy", "vz", "mass", "volume")}
    ctx.particles_upload(**empty)
    ctx.step(w.step_params())
    ctx.sort(); ctx.cull(); ctx.deposit_volume()
    assert ctx.count == 0
    one = worlds.make_particles(w, 1)
    got = _run_step(ctx, w, one)
    cells, want = helpers.oracle_step(sim, ptrs, w, one)
    assert np.array_equal(got["cell"], cells)
    _check_state(got, want, 3)


def test_full_size_c2_properties(ctx):
    """BASELINE config C2 at full size (128^3, 10 M particles): properties that
    need no oracle run -- sortedness, permutation, idempotent locate, volume
    conservation of the deposit, and a 1e5-particle oracle spot check."""
    w = worlds.make_c2()
    assert w.arrays.n_leaves == 128 ** 3 and w.arrays.n_vertices == 129 ** 3
    ctx.upload_tree(w.tree)
    ctx.upload_field(w.u, w.v, w.w)
    parts = worlds.make_particles(w)
    n = len(parts["x"])
    assert n == 10_000_000
    ctx.particles_upload(**parts)
    ctx.sort()
    pre = ctx.particles_download(ids=True)
    cells = ctx.locate(pre["x"], pre["y"], pre["z"])
    assert np.all(np.diff(cells) >= 0) and cells.min() >= 0
    assert np.array_equal(np.sort(pre["id"]), np.arange(1, n + 1, dtype=np.uint32))
    # cell centre of the located leaf is within h/2 of the particle (exact containment)
    a = w.arrays
    d = np.abs(np.stack([pre["x"], pre["y"], pre["z"]], 1) - a.pos[cells])
    assert np.all(d <= a.h[cells][:, None] / 2)
    ctx.deposit_volume()
    vol = ctx.download_deposit(0)
    assert abs(np.sum(vol) * (1 / 128) ** 3 - parts["volume"].sum()) <= 1e-12 * parts["volume"].sum()
    ctx.step(w.step_params(record_cells=True))
    got = ctx.particles_download(cells=True, ids=True)
    assert np.array_equal(got["cell"], cells)
    # oracle spot check on the first 1e5 ids
    sim, ptrs = helpers.matched_oracle(w)
    m = 100_000
    sub = {k: parts[k][:m] for k in parts}
    ocells, want = helpers.oracle_step(sim, ptrs, w, sub, nthreads=ora.load(3).ora_max_threads())
    sel = np.argsort(got["id"])[:m]          # rows holding ids 1..m, in id order
    assert np.array_equal(got["cell"][sel], ocells)
    sub_got = {k: got[k][sel] for k in STATE}
    _check_state(sub_got, want, 3)


@pytest.mark.parametrize("config", ["C3", "C5"])
def test_full_size_adaptive_trees(config, ctx):
    """BASELINE configs C3 (levels 5-9) and the C4/C5 tree (levels 6-10) at FULL depth, 10 M
    particles: the corner-balanced refinement down to level 9 / 10 (ftt_refine_corner,
    src/ftt.c:2013-2075) is where a flat-tree bug would hide.
      * the native tree equals flatten(the tree the reference's own ftt.c builds), array by array
        (helpers.matched_oracle asserts it);
      * locate bit-exact against ftt_cell_locate (src/ftt.c:1535-1574) on 1e6 random + 20 k
        adversarial points;
      * sortedness / permutation / containment of the 10 M-particle cloud, volume conservation of
        the deposit, fused == separate deposits;
      * a 1e5-particle oracle spot check of the step (cells bit-exact, state <= 1e-12)."""
    w = worlds.make_c3() if config == "C3" else worlds.make_c5(n_particles=10_000_000)
    a = w.arrays
    lo, hi = (5, 9) if config == "C3" else (6, 10)
    hist = w.level_histogram()
    assert min(hist) == lo and max(hist) == hi and all(hist[l] > 0 for l in range(lo, hi + 1))
    sim, ptrs = helpers.matched_oracle(w)               # flatten(reference tree) == native tree
    idx = helpers.PtrIndex(ptrs)
    ctx.upload_tree(w.tree)
    ctx.upload_field(w.u, w.v, w.w)
    rng = np.random.default_rng(99)
    n_pts = 1_000_000
    pts = [rng.uniform(-0.55, 0.55, n_pts) for _ in range(3)]
    # half of them near the ring, where the fine levels are
    th = rng.uniform(0, 2 * np.pi, n_pts // 2)
    pts[0][:n_pts // 2] = 0.25 * np.cos(th) + rng.normal(0, 0.01, n_pts // 2)
    pts[1][:n_pts // 2] = 0.25 * np.sin(th) + rng.normal(0, 0.01, n_pts // 2)
    pts[2][:n_pts // 2] = rng.normal(0, 0.01, n_pts // 2)
    ax, ay, az = worlds.adversarial_points(a, rng, 20000)
    x, y, z = (np.concatenate(c) for c in zip(pts, (ax, ay, az)))
    got = ctx.locate(x, y, z)
    want = idx(sim.locate(x, y, z))
    assert np.array_equal(got, want)
    assert (got >= 0).sum() > 800_000 and (got < 0).sum() > 10_000
    assert len(np.unique(a.level[got[got >= 0]])) == hi - lo + 1      # every level was hit

    parts = worlds.make_particles(w)
    n = len(parts["x"])
    assert n == 10_000_000
    ctx.particles_upload(**parts)
    ctx.sort()
    pre = ctx.particles_download(ids=True)
    cells = ctx.locate(pre["x"], pre["y"], pre["z"])
    assert np.all(np.diff(cells) >= 0) and cells.min() >= 0
    assert np.array_equal(np.sort(pre["id"]), np.arange(1, n + 1, dtype=np.uint32))
    d = np.abs(np.stack([pre["x"], pre["y"], pre["z"]], 1) - a.pos[cells])
    assert np.all(d <= a.h[cells][:, None] / 2)
    ctx.deposit_all(w.step_params())
    sep = [ctx.download_deposit(c) for c in range(4)]
    assert abs(np.sum(sep[0] * a.h ** 3) - parts["volume"].sum()) <= 1e-12 * parts["volume"].sum()
    ctx.step(w.step_params(record_cells=True))
    got = ctx.particles_download(cells=True, ids=True)
    assert np.array_equal(got["cell"], cells)
    # the fused flavour on the same start: same state, same field as step + deposit_all
    ctx.deposit_all(w.step_params())
    after = [ctx.download_deposit(c) for c in range(4)]
    ctx.particles_upload(**{k: pre[k] for k in ("x", "y", "z", "vx", "vy", "vz", "mass", "volume")}, ids=pre["id"])
    ctx.step(w.step_params(fuse_deposit=True))
    fused_state = ctx.particles_download()
    for k in STATE:
        assert np.array_equal(fused_state[k], got[k]), k
    for c in range(4):
        f = ctx.download_deposit(c)
        assert np.abs(f - after[c]).max() <= 1e-12 * np.abs(after[c]).max(), c
    # oracle spot check on the first 1e5 ids
    m = 100_000
    sub = {k: parts[k][:m] for k in parts}
    ocells, want = helpers.oracle_step(sim, ptrs, w, sub, nthreads=ora.load(3).ora_max_threads())
    sel = np.argsort(got["id"])[:m]
    assert np.array_equal(got["cell"][sel], ocells)
    _check_state({k: got[k][sel] for k in STATE}, want, 3)


@pytest.mark.parametrize("mode", ["0", "2", "3", "7", "9"])
@pytest.mark.parametrize("kind", ["c1", "uniform3", "ring3", "ring2", "chain2"])
def test_step_fast_paths(kind, mode, monkeypatch):
    """The production launch (no cell/force recording): compile-time force
    lists, lattice addressing on uniform trees, the TMA-staged persistent
    kernel (GFSB200_STEP_MODE 2/3: CTA-wide stages) and the warp-pipelined one
    (7/9: per-warp stages, particle fetched late from the staged tile) must all
    match the oracle, whatever the default picks for the tree."""
    monkeypatch.setenv("GFSB200_STEP_MODE", mode)
    c = capi.Context(0)
    try:
        w, sim, ptrs, idx = setup(kind, c)
        n = 20011                                   # not a multiple of the 256- or 32-particle tiles
        parts = _particles(w, n)
        c.particles_upload(**parts)
        for forces in (w.forces, (capi.FORCE_DRAG, capi.FORCE_BUOY), (capi.FORCE_LIFT,)):
            w2 = worlds.World(**{**w.__dict__, "forces": forces})
            c.particles_upload(**parts)
            c.step(w2.step_params())
            got = c.particles_download()
            cells, want = helpers.oracle_step(sim, ptrs, w2, parts)
            _check_state(got, want, w.dim)
    finally:
        c.close()


@pytest.mark.parametrize("kind", ["c1", "ring3", "uniform3"])
def test_step_host_streams_match_oracle(kind, ctx):
    """gfsb200_step_host: host arrays in, host arrays out (in place), chunked"""
    w, sim, ptrs, idx = setup(kind, ctx)
    n = 30011
    parts = worlds.make_particles(w, n)
    host = {k: (None if v is None else np.ascontiguousarray(v).copy()) for k, v in parts.items()}
    ctx.step_host(w.step_params(), host["x"], host["y"], host["z"], host["vx"], host["vy"], host["vz"],
                  host["mass"], host["volume"], chunk=4096)
    cells, want = helpers.oracle_step(sim, ptrs, w, parts)
    _check_state(host, want, w.dim)
    assert np.array_equal(host["mass"], parts["mass"]) and np.array_equal(host["volume"], parts["volume"])


def test_c1_full_config(ctx):
    """BASELINE config C1 as specified: 2D level-6 quadtree with four Dirichlet
    ghost layers, lid-style field, 1000 particles, GfsForceDrag, dt = 1e-2,
    1000 steps -- against the oracle every 100 steps."""
    w = worlds.make_c1()
    assert w.arrays.n_leaves == 64 * 64 and w.arrays.n_roots == 5
    sim, ptrs = helpers.matched_oracle(w)
    idx = helpers.PtrIndex(ptrs)
    ctx.upload_tree(w.tree)
    ctx.upload_field(w.u, w.v, None)
    parts = worlds.make_particles(w)
    assert len(parts["x"]) == 1000
    ctx.particles_upload(**parts)
    plist = ora.ParticleList(sim, *[parts[k] for k in ("x", "y", "z", "vx", "vy", "vz", "mass", "volume")])
    opar = helpers.oracle_params(w)
    par = w.step_params(record_cells=True)
    same = np.ones(1000, dtype=bool)
    worst = 0.0
    for step in range(1, 1001):
        if step % 100 == 0:
            s = plist.get()
            ocell = idx(sim.locate(s["x"], s["y"], None))
        plist.step(opar)
        ctx.step(par)
        if step % 100 == 0:
            got = ctx.particles_download(cells=True)
            same &= got["cell"] == ocell
            want = plist.get()
            worst = max(worst, max(helpers.vec_rel_err(got, want, k, same) for k in _vec_keys(2)))
    assert same.mean() >= 0.99, same.mean()
    assert worst <= 1e-9, worst          # documented drift bound after 1000 steps (measured ~1e-12)


@pytest.mark.parametrize("kind", ["c1", "uniform3", "ring3", "ring3b"])
@pytest.mark.parametrize("forces,kw", [
    ((capi.FORCE_INERTIAL,), {}),
    ((capi.FORCE_ADDEDMASS, capi.FORCE_DRAG, capi.FORCE_BUOY), {}),
    ((capi.FORCE_DRAG, capi.FORCE_INERTIAL, capi.FORCE_ADDEDMASS, capi.FORCE_LIFT), dict(cm_const=0.3)),
])
def test_inertial_and_added_mass(kind, forces, kw, ctx):
    """GfsForceInertial (:255-303) and GfsForceAddedMass (:331-394): previous-step
    velocity Un,Vn,Wn through a second vertex table, (u.grad)u per leaf, and the
    reference's cumulative mass += rho V cm, over three steps"""
    w, sim, ptrs, idx = setup(kind, ctx)
    a = w.arrays
    rng = np.random.default_rng(21)
    fields = [w.u, w.v] + ([w.w] if w.dim == 3 else [])
    prev = [0.8 * f + 0.05 * rng.standard_normal(a.n_cells) for f in fields]
    live = (a.flags & capi.CELL_DESTROYED) == 0
    for c, f in enumerate(prev):
        sim.set_values(5 + c, ptrs[live], f[live])
    ctx.upload_field_prev(*prev)
    w2 = worlds.World(**{**w.__dict__, "forces": forces, "g": (0.1, -1.0, 0.0)})
    parts = worlds.make_particles(w2, 4000)
    ctx.particles_upload(**parts)
    plist = ora.ParticleList(sim, *[parts[k] for k in ("x", "y", "z", "vx", "vy", "vz", "mass", "volume")])
    opar = helpers.oracle_params(w2, ivar_uold=5, **kw)
    par = w2.step_params(record_cells=True, record_forces=True, **kw)
    for step in range(3):
        plist.step(opar)
        ctx.step(par)
        got = ctx.particles_download(forces=True)
        want = plist.get()
        _check_state(got, want, w.dim, rtol=1e-12 if step == 0 else 1e-11)
        assert np.abs(got["mass"] - want["mass"]).max() <= 1e-14 * np.abs(want["mass"]).max()
        for k in ("fx", "fy", "fz"):
            scale = max(np.abs(want[k]).max(), 1e-300)
            assert np.abs(got[k] - want[k]).max() <= 1e-10 * scale, (step, k)
    if capi.FORCE_ADDEDMASS in forces:
        assert np.all(got["mass"] > parts["mass"])


def test_inertial_needs_the_previous_field():
    c = capi.Context(0)
    try:
        w = worlds.make_c2(level=3, n_particles=100)
        c.upload_tree(w.tree)
        c.upload_field(w.u, w.v, w.w)
        c.particles_upload(**worlds.make_particles(w))
        with pytest.raises(capi.GfsB200Error, match="upload_field_prev"):
            c.step(capi.StepParams(1e-3, (capi.FORCE_INERTIAL,)))
    finally:
        c.close()


_periodic_world = helpers.periodic_world


def test_recorded_forces_follow_their_particles(ctx):
    """the force a step records per particle (particulate->force) stays with the particle
    when the list is re-sorted or compacted (cull, BC drop) afterwards"""
    w, sim, ptrs, idx = setup("ring3", ctx)
    parts = worlds.make_particles(w, 5000)
    ctx.particles_upload(**parts)
    ctx.step(w.step_params(record_forces=True))
    before = ctx.particles_download(forces=True, ids=True)
    ctx.sort()
    after = ctx.particles_download(forces=True, ids=True)
    assert not np.array_equal(before["id"], after["id"])
    a, b = np.argsort(before["id"]), np.argsort(after["id"])
    for k in ("fx", "fy", "fz", "x", "vx"):
        assert np.array_equal(before[k][a], after[k][b]), k
    # push a third of the particles out, cull: the survivors keep their forces
    x = after["x"].copy()
    x[::3] = 5.0
    moved = dict(after, x=x)
    ctx.particles_upload(**{k: moved[k] for k in ("x", "y", "z", "vx", "vy", "vz", "mass", "volume")})
    ctx.step(w.step_params(record_forces=True))
    full = ctx.particles_download(forces=True, ids=True)
    removed = ctx.cull()
    assert removed == len(x[::3])
    kept = ctx.particles_download(forces=True, ids=True)
    sel = np.isin(full["id"], kept["id"])
    for k in ("fx", "fy", "fz"):
        assert np.array_equal(full[k][sel], kept[k]), k


def test_escaped_count_matches_host_locate(ctx):
    """gfsb200_escaped_count = number of particles whose new position gfs_domain_locate
    rejects, which is what gfs_particle_bc would find one locate at a time"""
    w, sim, ptrs, idx = setup("ring3", ctx)
    rng = np.random.default_rng(17)
    parts = worlds.make_particles(w, 6000)
    n = len(parts["x"])
    parts["x"] = rng.uniform(-0.499, 0.499, n)
    for k, f in zip(("vx", "vy", "vz"), (4.0, 4.0, 4.0)):
        parts[k] = f * rng.standard_normal(n)
    ctx.particles_upload(**parts)
    par = worlds.World(**{**w.__dict__, "dt": 2e-2}).step_params(track_escapes=True)
    ctx.step(par)
    got = ctx.particles_download()
    outside = int((sim.locate(got["x"], got["y"], got["z"]) == 0).sum())
    assert outside > 10
    assert ctx.escaped_count() == outside
    assert ctx.step_counts() == (outside, 0)
    eidx, eold = ctx.escaped()
    gone = np.nonzero(sim.locate(got["x"], got["y"], got["z"]) == 0)[0]
    assert np.array_equal(np.sort(eidx), gone)
    for a, k in enumerate(("x", "y", "z")):                 # the positions before the step
        assert np.array_equal(eold[:, a], parts[k][eidx])
    ctx.step(par)                              # the escapees are still in the list: outside now
    assert ctx.step_counts()[1] == outside
    ctx.step(w.step_params())
    with pytest.raises(capi.GfsB200Error):
        ctx.escaped_count()                    # the last step did not track


@pytest.mark.parametrize("dim", [2, 3])
def test_periodic_wrap_and_drop(dim, ctx):
    """gfs_particle_list_event with gfs_particle_bc: cull, step, then wrap the
    particles that crossed a periodic side (periodic_bc_particle) and drop the
    ones that crossed a plain boundary -- 25 steps against the oracle, by id"""
    w, mask = _periodic_world(dim)
    sim, ptrs = helpers.matched_oracle(w)
    ctx.upload_tree(w.tree)
    ctx.upload_field(w.u, w.v, w.w)
    rng = np.random.default_rng(9)
    parts = worlds.make_particles(w)
    n = len(parts["x"])
    for k in ("x", "y", "z")[:dim]:
        parts[k] = rng.uniform(-0.499, 0.499, n)
    for k, f in zip(("vx", "vy", "vz")[:dim], (3.0, 3.0, 2.0)):
        parts[k] = f * rng.standard_normal(n)
    ctx.particles_upload(**parts)
    plist = ora.ParticleList(sim, *[parts[k] for k in ("x", "y", "z", "vx", "vy", "vz", "mass", "volume")])
    opar = helpers.oracle_params(w)
    par = w.step_params()
    wrapped_total = dropped_total = 0
    for step in range(25):
        before = ctx.count
        removed = ctx.particle_list_event(par)
        plist.cull()
        plist.step(opar)
        n_before_bc = len(plist)
        odrop = plist.bc(mask)
        dropped_total += removed
        assert ctx.count == len(plist) == before - removed
        got = ctx.particles_download(ids=True)
        want = plist.get()
        go, wo = np.argsort(got["id"]), np.argsort(want["id"])
        assert np.array_equal(got["id"][go], want["id"][wo])
        g = {k: got[k][go] for k in STATE if got[k] is not None and (dim == 3 or k not in ("z", "vz"))}
        wv = {k: want[k][wo] for k in g}
        _check_state(g, wv, dim, rtol=1e-11)
        # wrapped particles were re-appended at the end of the reference's list
        wrapped_total += int((np.diff(want["id"].astype(np.int64)) < 0).sum() > 0)
    assert dropped_total > 20, "the test did not exercise dropping"
    assert wrapped_total > 5, "the test did not exercise wrapping"
    inside = ctx.locate(got["x"], got["y"], got["z"] if dim == 3 else None) >= 0
    assert inside.mean() > 0.95


# ---------------------------------------------------------------------------
# mixed (solid-cut) cells

class _SolidWorld:
    """a test world whose reference tree has a quarter of its leaves turned into mixed cells
    (prescribed GfsSolidVector a / cm / s, some faces closed), flattened through the FttCell
    bridge as the module does; the shared oracle tree is restored on exit"""

    def __init__(self, kind, ctx, seed=5):
        self.w, self.sim, self.ptrs, self.idx = setup(kind, ctx)
        self.ctx, a = ctx, self.w.arrays
        rng = np.random.default_rng(seed)
        leaves = a.box_leaves
        pick = list(rng.choice(leaves, max(1, len(leaves) // 4), replace=False))
        ghost = np.nonzero((a.child0 < 0) & ((a.flags & capi.CELL_BOUNDARY) != 0) &
                           ((a.flags & capi.CELL_DESTROYED) == 0))[0]
        self.mixed = pick + list(ghost[::5])

        def solid(i):
            h = 2.0 ** -int(a.level[i])
            return (float(rng.uniform(0.05, 1.0)), a.pos[i, :a.dim] + rng.uniform(-0.45, 0.45, a.dim) * h,
                    rng.choice([0.0, 0.3, 1.0], 2 * a.dim, p=[0.2, 0.4, 0.4]))
        self.solid = {int(i): solid(i) for i in self.mixed}

    def __enter__(self):
        w, sim, ptrs = self.w, self.sim, self.ptrs
        for i, (fa, cm, fs) in self.solid.items():
            sim.set_solid(ptrs[i], fa, cm, fs)
        roots, is_box = sim.roots()
        self.tree, fmap = capi.flatten_ftt(w.dim, roots, is_box)
        assert np.array_equal(fmap.cells, ptrs)
        self.tree.build_stencils()
        self.arrays = self.tree.view()
        assert self.arrays.solid_a is not None
        self.ctx.upload_tree(self.tree)
        self.ctx.upload_field(w.u, w.v, w.w)
        return self

    def __exit__(self, *exc):
        for i in self.solid:
            self.sim.set_solid(self.ptrs[i], 0.0, np.zeros(3))
        self.ctx.upload_tree(self.w.tree)
        self.ctx.upload_field(self.w.u, self.w.v, self.w.w)
        return False


@pytest.mark.parametrize("kind", ["c1", "uniform3", "ring3", "ring2", "chain3"])
def test_mixed_cells_interpolation_and_vorticity(kind, ctx):
    """GFS_IS_MIXED cells on the device: corner values / gfs_interpolate with the
    centre-of-mass weights, vorticity with closed faces (gfs_cell_face) and face-fraction
    averages over finer neighbours (average_neighbor_value) -- vorticity bit-exact"""
    with _SolidWorld(kind, ctx) as sw:
        w, sim, ptrs = sw.w, sw.sim, sw.ptrs
        leaves = sw.arrays.box_leaves
        vort = ctx.vorticity(leaves)
        want = sim.vorticity(ptrs[leaves])
        assert np.array_equal(vort, want), f"max diff {np.abs(vort - want).max():.3e}"
        fields = [w.u, w.v] + ([w.w] if w.dim == 3 else [])
        for comp, f in enumerate(fields):
            got = ctx.corner_values(comp, leaves)
            want = sim.corner_values(comp, ptrs[leaves])
            tol = 4 * np.finfo(np.float64).eps * max(np.abs(f).max(), 1e-300)
            assert np.abs(got - want).max() <= tol
        rng = np.random.default_rng(1)
        parts = _particles(w, 20000)
        x, y, z = parts["x"], parts["y"], parts["z"]
        u = ctx.interpolate(x, y, z)
        inside = sim.locate(x, y, z) != 0
        for comp, f in enumerate(fields):
            want = sim.interpolate(comp, x, y, z)
            tol = 8 * np.finfo(np.float64).eps * max(np.abs(f).max(), 1e-300)
            assert np.abs(u[comp][inside] - want[inside]).max() <= tol
    # and it matters: without the solids the reference gives another vorticity field
    assert not np.array_equal(sim.vorticity(ptrs[leaves]), vort)


@pytest.mark.parametrize("kind", ["c1", "ring3", "uniform3"])
def test_mixed_cells_step_and_deposits(kind, ctx):
    """the fused step and both force deposits on a tree with mixed cells, against the
    reference arithmetic (gfs_cell_volume = h^dim a in the deposits)"""
    with _SolidWorld(kind, ctx) as sw:
        w, sim, ptrs, idx = sw.w, sw.sim, sw.ptrs, sw.idx
        parts = _particles(w, 8000)
        got = _run_step(ctx, w, parts)
        cells, want = helpers.oracle_step(sim, ptrs, w, parts)
        assert np.array_equal(got["cell"], cells)
        _check_state(got, want, w.dim)
        # single-cell force deposit
        ctx.particles_upload(**parts)
        ctx.deposit_volume()
        ctx.deposit_force(w.step_params())
        live = (w.arrays.flags & capi.CELL_DESTROYED) == 0
        plist = ora.ParticleList(sim, *[parts[k] for k in ("x", "y", "z", "vx", "vy", "vz", "mass", "volume")])
        zero = np.zeros(int(live.sum()))
        for iv in range(3, 3 + 1 + w.dim):
            sim.set_values(iv, ptrs[live], zero)
        plist.deposit_volume(3)
        plist.deposit_force(helpers.oracle_params(w), 4)
        for comp in range(1 + w.dim):
            g = ctx.download_deposit(comp)[live]
            wv = sim.get_values(3 + comp, ptrs[live])
            assert np.abs(wv).max() > 0
            assert np.abs(g - wv).max() <= 1e-12 * np.abs(wv).max(), comp
        # kernel-smoothed deposit
        small = _particles(w, 1500)
        ctx.particles_upload(**small)
        rk = 0.06 if w.dim == 2 else 0.04
        ctx.deposit_force_smoothed(w.step_params(), rk, capi.KERNEL_GAUSSIAN, 1.0, 2e-4, 1, 0, record_norm=True)
        corr, vol = ctx.download_kernel_norm()
        plist = ora.ParticleList(sim, *[small[k] for k in ("x", "y", "z", "vx", "vy", "vz", "mass", "volume")])
        for iv in range(4, 4 + w.dim):
            sim.set_values(iv, ptrs[live], zero)
        wcorr, wvol = plist.deposit_force_smoothed(helpers.oracle_params(w), 4, rk,
                                                   ora.Kernel(ora.KERNEL_GAUSSIAN, 1.0, 2e-4, 1, 0))
        assert np.array_equal(vol, wvol)                    # same leaves, same h^dim a, same order
        ok = np.isfinite(wcorr)
        assert np.abs(corr[ok] - wcorr[ok]).max() <= 1e-13 * np.abs(wcorr[ok]).max()
        for comp in range(w.dim):
            g = ctx.download_deposit(1 + comp)[live]
            wv = sim.get_values(4 + comp, ptrs[live])
            assert np.abs(wv).max() > 0
            assert np.abs(g - wv).max() <= 1e-12 * np.abs(wv).max(), comp

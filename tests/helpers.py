"""Shared test helpers: pair a product World with an oracle Sim holding the
IDENTICAL tree (built by the reference's own ftt.c from the flat tree's list
of refined cells), and run the oracle's particulate step on the same inputs."""
from __future__ import annotations

import numpy as np

import __graft_entry__ as entry

pkg = entry.load_package()
capi, worlds = pkg.capi, pkg.worlds
ora = entry.load_oracle()

FORCE_MAP = {capi.FORCE_DRAG: ora.FORCE_DRAG, capi.FORCE_LIFT: ora.FORCE_LIFT, capi.FORCE_BUOY: ora.FORCE_BUOY,
             capi.FORCE_INERTIAL: ora.FORCE_INERTIAL, capi.FORCE_ADDEDMASS: ora.FORCE_ADDEDMASS}
TREE_KEYS = ("parent", "child0", "neighbor", "level", "flags", "pos")


def oracle_tree_from_flat(arrays) -> "ora.Sim":
    """Reference FTT tree with exactly the cells of the flat tree."""
    a = arrays
    sim = ora.Sim(a.dim, nvar=16)      # 0-2 U,V,W; 3-7 test fields; 11-15 Rep, Urelp ... of coefficient functions
    # boxes: the oracle supports a chain of unit boxes along +x starting at the origin
    for b in range(1, a.n_box_roots):
        assert np.array_equal(a.pos[b], [float(b), 0.0, 0.0]), "oracle boxes form a chain along +x"
        sim.add_box()
    # boundaries next, as Gerris does while reading the .gfs file
    for r in range(a.n_box_roots, a.n_roots):
        dist = np.abs(a.pos[:a.n_box_roots] - a.pos[r]).sum(1)
        b = int(np.argmin(dist))                      # the box this ghost root touches
        d = a.pos[r] - a.pos[b]
        axis = int(np.argmax(np.abs(d)))
        sim.add_boundary(2 * axis + (0 if d[axis] > 0 else 1), b)
    refined = np.nonzero((a.child0 >= 0) & ((a.flags & capi.CELL_BOUNDARY) == 0))[0]   # level order
    if len(refined):
        lv = np.ascontiguousarray(a.level[refined], dtype=np.int32)
        p = a.pos[refined]
        x, y, z = (np.ascontiguousarray(p[:, k]) for k in range(3))
        done = sim.L.ora_refine_points(sim.h, len(refined), lv.ctypes.data, x.ctypes.data, y.ctypes.data,
                                       z.ctypes.data if a.dim == 3 else None)
        assert done == len(refined), "flat tree is not 2:1 balanced the way ftt.c builds it"
    sim.finalize()
    return sim


def matched_oracle(world):
    """(sim, ptrs): oracle domain with the world's tree; ptrs[i] = FttCell* of flat cell i."""
    a = world.arrays
    sim = oracle_tree_from_flat(a)
    roots, is_box = sim.roots()
    tree2, fmap = capi.flatten_ftt(a.dim, roots, is_box)
    b = tree2.view()
    assert b.n_cells == a.n_cells, (b.n_cells, a.n_cells)
    for k in TREE_KEYS:
        assert np.array_equal(getattr(a, k), getattr(b, k)), f"flatten(reference tree).{k} != native tree"
    ptrs = fmap.cells.copy()
    sim._keep = (tree2, fmap)
    push_field(sim, ptrs, world)
    return sim, ptrs


def push_field(sim, ptrs, world):
    alive = ptrs != 0
    fields = [world.u, world.v] + ([world.w] if world.dim == 3 else [])
    live = (world.arrays.flags & capi.CELL_DESTROYED) == 0
    for i, f in enumerate(fields):
        sim.set_values(i, ptrs[live & alive], f[live & alive])


class PtrIndex:
    def __init__(self, ptrs):
        self.order = np.argsort(ptrs, kind="stable")
        self.sorted = ptrs[self.order]

    def __call__(self, p):
        p = np.asarray(p, dtype=np.uint64)
        out = np.full(p.shape, -1, dtype=np.int32)
        nz = p != 0
        j = np.searchsorted(self.sorted, p[nz])
        assert np.array_equal(self.sorted[j], p[nz]), "oracle returned a cell unknown to the flat tree"
        out[nz] = self.order[j]
        return out


def oracle_params(world, pattern=0, **kw):
    return ora.step_params(world.dt, [FORCE_MAP[f] for f in world.forces], rho=world.rho, mu=world.mu,
                           g=world.g, pattern=pattern, **kw)


def oracle_step(sim, ptrs, world, parts, steps=1, pattern=0, nthreads=1, **kw):
    """Runs `steps` reference particulate events; returns (start cells of the
    LAST step as flat indices, final state dict incl. forces)."""
    plist = ora.ParticleList(sim, *[parts[k] for k in ("x", "y", "z", "vx", "vy", "vz", "mass", "volume")])
    par = oracle_params(world, pattern, **kw)
    idx = PtrIndex(ptrs)
    cells = None
    for _ in range(steps):
        s = plist.get()
        cells = idx(sim.locate(s["x"], s["y"], s["z"]))
        plist.step(par, nthreads)
    return cells, plist.get()


def rel_err(got, want, floor=1e-300):
    return float(np.max(np.abs(got - want) / np.maximum(np.abs(want), floor))) if len(want) else 0.0


def vec_rel_err(got, want, keys, sel=None):
    """max over particles of |got_i - want_i|_2 / |want_i|_2 for the vector
    with components `keys` (position or velocity).  A per-component ratio is
    meaningless where one component passes through zero."""
    g = np.stack([got[k] if sel is None else got[k][sel] for k in keys], 1)
    w = np.stack([want[k] if sel is None else want[k][sel] for k in keys], 1)
    if not len(w):
        return 0.0
    num = np.sqrt(((g - w) ** 2).sum(1))
    den = np.sqrt((w ** 2).sum(1))
    return float(np.max(num / np.maximum(den, 1e-300)))


# ---------------------------------------------------------------------------
# the seeded test worlds shared by the GPU parity tests and the CPU oracle tests

def test_world(kind):
    if kind == "c1":
        return worlds.make_c1(level=5, n_particles=1000)
    if kind == "uniform3":
        return worlds.make_c2(level=4, n_particles=20000)
    if kind == "ring3":
        return worlds.make_ring("ring3", 3, 6, 30000, 3003)
    if kind == "ring2":
        t = capi.Tree(2)
        t.refine_ring(3, 7, 0.25, 1.5)
        t.corner_sweep()
        for s in range(4):
            t.add_boundary(s)
        t.finalize(); t.build_stencils()
        a = t.view()
        u, v = worlds.lid_style(a.pos)
        worlds.apply_dirichlet_ghosts(a, u, {2: 1.0})
        worlds.apply_dirichlet_ghosts(a, v, {})
        return worlds.World("ring2", 2, t, a, u, v, None, (capi.FORCE_DRAG, capi.FORCE_LIFT, capi.FORCE_BUOY),
                            dt=1e-2, mu=1e-3, g=(0.3, -1.0, 0.0), seed=11, n_particles=5000,
                            meta=dict(d_p=worlds.D_P, rho_p=worlds.RHO_P, v0="zero", field="lid"))
    if kind == "ring3b":
        t = capi.Tree(3)
        t.refine_ring(2, 5, 0.3, 1.5)
        t.corner_sweep()
        for s in range(6):
            t.add_boundary(s)
        t.finalize(); t.build_stencils()
        a = t.view()
        u, v, w = worlds.vortex_ring(a.pos)
        return worlds.World("ring3b", 3, t, a, u, v, w, (capi.FORCE_LIFT, capi.FORCE_DRAG),
                            dt=2e-3, mu=2e-3, rho=1.3, seed=12, n_particles=8000,
                            meta=dict(d_p=worlds.D_P, rho_p=worlds.RHO_P, v0="fluid", field="ring",
                                      cloud="half uniform, half gaussian"))
    if kind in ("chain2", "chain3"):
        # several GfsBoxes in a row along +x, refined around a ring straddling the first interface
        dim, nbox = (2, 3) if kind == "chain2" else (3, 2)
        minl, maxl = (3, 6) if dim == 2 else (2, 5)
        t = capi.Tree(dim)
        t.add_root((0.0, 0.0, 0.0))
        for b in range(1, nbox):
            t.add_root((float(b), 0.0, 0.0))
            t.link_roots(b - 1, 0, b)
        t.refine(lambda pos, level, h: level < minl or (level < maxl and abs(
            np.hypot(pos[0] - 0.5, pos[1]) - 0.3) < 1.5 * h and (dim == 2 or abs(pos[2]) < 3 * h)))
        t.corner_sweep()
        for b, sd in [(0, 1), (nbox - 1, 0)] + [(b, sd) for b in range(nbox) for sd in range(2, 2 * dim)]:
            t.add_boundary(sd, b)
        t.finalize(); t.build_stencils()
        a = t.view()
        u, v, wz = worlds.taylor_green(a.pos * 0.5)
        forces = (capi.FORCE_DRAG, capi.FORCE_LIFT, capi.FORCE_BUOY)
        return worlds.World(kind, dim, t, a, u, v, (0.3 * u if dim == 3 else None), forces, dt=2e-3, mu=1e-3,
                            g=(0.2, -1.0, 0.0), seed=14, n_particles=6000,
                            meta=dict(d_p=worlds.D_P, rho_p=worlds.RHO_P, v0="zero", field="tg", nbox=nbox))
    raise KeyError(kind)


def test_particles(w, n=None):
    """the world's seeded cloud; chains of boxes get it stretched over all boxes"""
    parts = worlds.make_particles(w, n)
    nbox = w.meta.get("nbox", 1)
    if nbox > 1:
        rng = np.random.default_rng(w.seed + 1)
        parts["x"] = rng.uniform(-0.47, nbox - 0.53, len(parts["x"]))
    return parts


def periodic_world(dim):
    """uniform (3D) / ring-refined (2D) box with ghost layers on every side;
    x (and z in 3D) periodic, y sides plain boundaries (particles leaving there are dropped)"""
    t = capi.Tree(dim)
    if dim == 3:
        t.refine_uniform(4)
    else:
        t.refine_ring(3, 6, 0.25, 1.5)
        t.corner_sweep()
    for s in range(2 * dim):
        t.add_boundary(s)
    periodic = [0, 1] + ([4, 5] if dim == 3 else [])
    for s in periodic:
        t.set_periodic(s)
    t.finalize(); t.build_stencils()
    a = t.view()
    u, v, wz = worlds.taylor_green(a.pos)            # period 1: ghost values = field at the ghost centres
    w = worlds.World("periodic%d" % dim, dim, t, a, 3.0 * u, 3.0 * v + 0.7, wz if dim == 3 else None,
                     (capi.FORCE_DRAG, capi.FORCE_BUOY), dt=4e-3, mu=1e-3, g=(0.0, -1.0, 0.0), seed=5,
                     n_particles=3000, meta=dict(d_p=worlds.D_P, rho_p=worlds.RHO_P, v0="fluid", field="tg"))
    mask = sum(1 << s for s in periodic)
    return w, mask


test_world.__test__ = test_particles.__test__ = False

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
    sim = ora.Sim(a.dim, nvar=8)
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

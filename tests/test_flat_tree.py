"""CPU tests of the host side of the product: the flat FTT tree (native
builders, 2:1 and corner balancing, ghost trees, level ordering), the FttCell
bridge, and the corner-stencil table -- all against the reference's own ftt.c /
fluid.c object code (the oracle)."""
import numpy as np
import pytest

import helpers
from helpers import capi, worlds, ora

KEYS = ("parent", "child0", "neighbor", "level", "flags", "pos")


def build_pair(dim, minl, maxl, sides, R=0.25):
    t = capi.Tree(dim)
    t.refine_ring(minl, maxl, R, 1.5)
    t.corner_sweep()
    for s in sides:
        t.add_boundary(s)
    t.finalize()
    sim = ora.Sim(dim)
    for s in sides:
        sim.add_boundary(s)
    sim.refine_ring(minl, maxl, R, 1.5)
    sim.corner_sweep()
    sim.finalize()
    return t, sim


@pytest.mark.parametrize("dim,minl,maxl,sides", [
    (2, 3, 7, ()), (2, 2, 6, (0, 1, 2, 3)), (2, 4, 4, (0, 3)),
    (3, 3, 6, ()), (3, 2, 5, (0, 1, 2, 3, 4, 5)), (3, 4, 4, (4,)),
])
def test_native_builder_equals_flattened_reference_tree(dim, minl, maxl, sides):
    """same refinement criterion through ftt_cell_refine + ftt_refine_corner +
    boundary_match (reference) and through the native builders: identical
    arrays, cell for cell"""
    t, sim = build_pair(dim, minl, maxl, sides)
    a = t.view()
    roots, is_box = sim.roots()
    t2, fmap = capi.flatten_ftt(dim, roots, is_box)
    b = t2.view()
    assert a.n_cells == b.n_cells == len(fmap.cells)
    for k in KEYS:
        assert np.array_equal(getattr(a, k), getattr(b, k)), k
    assert np.array_equal(a.la_min, b.la_min) and np.array_equal(a.la_n, b.la_n)
    assert np.array_equal(a.la_slot, b.la_slot) and a.complete_level == b.complete_level
    assert sim.count() == int(((a.flags & capi.CELL_BOUNDARY) == 0).sum())


@pytest.mark.parametrize("dim,minl,maxl,sides", [(2, 2, 6, (0, 1, 2, 3)), (3, 2, 5, (0, 2, 5)), (3, 3, 6, ())])
def test_neighbors_match_ftt_cell_neighbor(dim, minl, maxl, sides):
    t, sim = build_pair(dim, minl, maxl, sides)
    roots, is_box = sim.roots()
    t2, fmap = capi.flatten_ftt(dim, roots, is_box)
    a = t2.view()
    idx = helpers.PtrIndex(fmap.cells)
    L = sim.L
    live = np.nonzero((a.flags & capi.CELL_DESTROYED) == 0)[0]
    for d in range(2 * dim):
        want = idx(np.array([L.ora_neighbor(int(fmap.cells[i]), d) for i in live], dtype=np.uint64))
        assert np.array_equal(a.neighbor[live, d], want), d


@pytest.mark.parametrize("dim,minl,maxl,sides", [(2, 2, 6, (0, 1, 2, 3)), (3, 2, 5, (0, 1, 2, 3, 4, 5)), (3, 3, 6, ())])
def test_corner_stencils_bit_exact(dim, minl, maxl, sides):
    """every (leaf, corner) interpolator equals gfs_cell_corner_interpolator:
    same cells, same order, same weights to the last bit -- including
    T-junctions and the domain-corner rule"""
    t, sim = build_pair(dim, minl, maxl, sides)
    roots, is_box = sim.roots()
    t2, fmap = capi.flatten_ftt(dim, roots, is_box)
    t2.build_stencils()
    a = t2.view()
    idx = helpers.PtrIndex(fmap.cells)
    nmax = 0
    for i in a.box_leaves:
        for k in range(2 ** dim):
            cells, w = t2.corner_interpolator(int(i), k)
            oc, ow = sim.corner_interpolator(fmap.cells[i], k)
            assert list(idx(np.array(oc, dtype=np.uint64))) == cells and ow == w, (i, k)
            nmax = max(nmax, len(cells))
            # the shared vertex entry holds the same stencil as a set
            v = a.leaf_vtx[i, k]
            sl = slice(a.vtx_off[v], a.vtx_off[v + 1])
            assert sorted(zip(a.vtx_cell[sl], a.vtx_w[sl])) == sorted(zip(cells, w))
    assert nmax > 2 ** dim or not sides == ()      # T-junction stencils were exercised


def test_uniform_tree_is_a_lattice():
    w = worlds.make_c2(level=4, n_particles=0)
    a = w.arrays
    assert a.n_vertices == 17 ** 3 and a.lattice_level == 4 and a.complete_level == 4
    # row-major lattice numbering: vertex of leaf corner 4 (-,-,-) at column (kx,ky,kz)
    leaves = a.box_leaves
    k = np.rint((a.pos[leaves] + 0.5) * 16 - 0.5).astype(int)
    assert np.array_equal(a.leaf_vtx[leaves, 4], (k[:, 2] * 17 + k[:, 1]) * 17 + k[:, 0])
    assert np.array_equal(a.leaf_vtx[leaves, 2], ((k[:, 2] + 1) * 17 + k[:, 1] + 1) * 17 + k[:, 0] + 1)
    ring = worlds.make_ring("r", 3, 5, 0, 1)
    assert ring.arrays.lattice_level == -1 and ring.arrays.complete_level == 3


def test_level_order_and_morton_keys():
    """cells are level-ordered; within a complete level the index is the Morton
    path (child digits: bit0 = +x, bit1 = -y, bit2 = -z)"""
    w = worlds.make_c2(level=3, n_particles=0)
    a = w.arrays
    assert np.all(np.diff(a.level.astype(int)) >= 0)
    for l in range(1, 4):
        s = a.level_start[l]
        n = 8 ** l
        pos = a.pos[s:s + n]
        k = np.rint((pos + 0.5) * 2 ** l - 0.5).astype(int)
        key = np.zeros(n, dtype=int)
        for b in range(l):
            key |= ((k[:, 0] >> b) & 1) << (3 * b)
            key |= (1 - ((k[:, 1] >> b) & 1)) << (3 * b + 1)
            key |= (1 - ((k[:, 2] >> b) & 1)) << (3 * b + 2)
        assert np.array_equal(key, np.arange(n))


def test_gather_scatter_through_the_bridge():
    t, sim = build_pair(3, 2, 4, (0,))
    roots, is_box = sim.roots()
    t2, fmap = capi.flatten_ftt(3, roots, is_box)
    a = t2.view()
    live = (a.flags & capi.CELL_DESTROYED) == 0
    vals = np.arange(a.n_cells, dtype=np.float64) * 0.5
    offset = 6 * 16 + 8            # offsetof (GfsStateVector, place_holder) in 3D
    fmap.scatter(offset, 2, vals)
    assert np.array_equal(sim.get_values(2, fmap.cells[live]), vals[live])
    back = fmap.gather(offset, 2)
    assert np.array_equal(back[live], vals[live]) and np.all(back[~live] == capi.NODATA)


def test_error_paths():
    t = capi.Tree(3)
    with pytest.raises(capi.GfsB200Error):
        t.view()                                   # not finalized
    t.refine_uniform(1)
    with pytest.raises(capi.GfsB200Error):
        t.split(0)                                 # root is not a leaf any more
    with pytest.raises(capi.GfsB200Error):
        t.add_root((1, 0, 0))                      # roots after a split
    t.add_boundary(0)
    with pytest.raises(capi.GfsB200Error):
        t.add_boundary(0)                          # side already taken
    with pytest.raises(capi.GfsB200Error):
        t.corner_sweep()                           # boundaries must come last
    with pytest.raises(capi.GfsB200Error):
        capi.Tree(4)


def make_mixed(sim, fmap, a, rng, frac=0.25):
    """turns a random quarter of the box leaves (and a few ghost leaves) into mixed cells
    with a prescribed fluid fraction and centre of mass; returns {flat index: (a, cm)}"""
    leaves = a.box_leaves
    pick = list(rng.choice(leaves, max(1, int(frac * len(leaves))), replace=False))
    ghost = np.nonzero((a.child0 < 0) & ((a.flags & capi.CELL_BOUNDARY) != 0) &
                       ((a.flags & capi.CELL_DESTROYED) == 0))[0]
    pick += list(ghost[::7])
    mixed = {}
    for i in pick:
        h = 2.0 ** -int(a.level[i])
        frac_a = float(rng.uniform(0.05, 1.0))
        cm = a.pos[i, :a.dim] + rng.uniform(-0.45, 0.45, a.dim) * h
        fs = rng.choice([0.0, 0.3, 1.0], 2 * a.dim, p=[0.2, 0.4, 0.4])       # some faces closed
        sim.set_solid(fmap.cells[i], frac_a, cm, fs)
        mixed[int(i)] = (frac_a, cm, fs)
    return mixed


@pytest.mark.parametrize("dim,minl,maxl,sides", [(2, 2, 6, (0, 1, 2, 3)), (3, 2, 5, (0, 1, 2, 3, 4, 5)), (3, 4, 4, ())])
def test_mixed_cells_through_the_bridge_and_in_the_stencils(dim, minl, maxl, sides):
    """GFS_IS_MIXED cells: the bridge carries GfsSolidVector.a / .cm into the flat tree and
    the corner interpolators weight such a cell by the distance from the corner to its
    centre of mass (distance (), src/fluid.c:2983-3003) -- same cells, same order, same
    weights to the last bit as gfs_cell_corner_interpolator on the reference tree,
    T-junctions, ghost cells and the uniform (otherwise lattice) tree included"""
    t, sim = build_pair(dim, minl, maxl, sides)
    roots, is_box = sim.roots()
    t0, fmap0 = capi.flatten_ftt(dim, roots, is_box)
    rng = np.random.default_rng(8)
    mixed = make_mixed(sim, fmap0, t0.view(), rng)
    t2, fmap = capi.flatten_ftt(dim, roots, is_box)
    a = t2.view()
    assert a.solid_a is not None and np.array_equal(fmap.cells, fmap0.cells)
    for i, (fa, cm, fs) in mixed.items():
        assert a.solid_a[i] == fa and np.array_equal(a.solid_cm[i, :dim], cm)
        assert np.array_equal(a.solid_s[i], fs)
    rest = np.setdiff1d(np.arange(a.n_cells), list(mixed))
    assert np.all(a.solid_a[rest] == 1.0) and np.all(np.isnan(a.solid_cm[rest, 0])) and np.all(a.solid_s[rest] == 1.0)
    t2.build_stencils()
    a = t2.view()
    assert a.lattice_level == -1 or minl == maxl          # weights differ: no single-weight shortcut needed
    idx = helpers.PtrIndex(fmap.cells)
    changed = 0
    for i in a.box_leaves:
        for k in range(2 ** dim):
            cells, w = t2.corner_interpolator(int(i), k)
            oc, ow = sim.corner_interpolator(fmap.cells[i], k)
            assert list(idx(np.array(oc, dtype=np.uint64))) == cells and ow == w, (i, k)
            changed += any(c in mixed for c in cells)
            v = a.leaf_vtx[i, k]
            sl = slice(a.vtx_off[v], a.vtx_off[v + 1])
            assert sorted(zip(a.vtx_cell[sl], a.vtx_w[sl])) == sorted(zip(cells, w))
    assert changed > 100
    for i in mixed:                                        # leave the shared oracle tree clean
        sim.set_solid(fmap.cells[i], 0.0, np.zeros(3))


@pytest.mark.parametrize("dim,level,sides", [(3, 4, ()), (3, 3, (0, 1, 2, 3, 4, 5)), (2, 5, (0, 1, 2, 3)), (3, 4, (2, 5))])
def test_lattice_stencil_fast_path_equals_general_path(dim, level, sides, monkeypatch):
    """uniform trees build one interpolator per lattice vertex (canonical leaf);
    the result must be the table the general de-duplicating builder produces"""
    def build():
        t = capi.Tree(dim)
        t.refine_uniform(level)
        for s in sides:
            t.add_boundary(s)
        t.finalize()
        t.build_stencils()
        return t, t.view()
    ta, a = build()
    monkeypatch.setenv("GFSB200_GENERAL_STENCILS", "1")
    tb, b = build()
    assert a.lattice_level == b.lattice_level == level
    for k in ("vtx_off", "vtx_cell", "vtx_w", "leaf_vtx"):
        assert np.array_equal(getattr(a, k), getattr(b, k)), k


def build_chain(dim, nbox, minl, maxl, R, ghost=True):
    """nbox unit boxes in a row along +x (a multi-GfsBox domain), refined around a
    ring that straddles the box interfaces, with ghost layers on the outer sides"""
    t = capi.Tree(dim)
    sim = ora.Sim(dim)
    t.add_root((0.0, 0.0, 0.0))
    for b in range(1, nbox):
        t.add_root((float(b), 0.0, 0.0))
        t.link_roots(b - 1, 0, b)
        sim.add_box()
    sides = []
    if ghost:
        sides = [(0, 1), (nbox - 1, 0)] + [(b, s) for b in range(nbox) for s in range(2, 2 * dim)]
        for b, s in sides:
            sim.add_boundary(s, b)
    # centre the ring on the interface between box 0 and box 1
    crit = lambda pos, level, h: level < minl or (level < maxl and abs(
        np.hypot(pos[0] - 0.5, pos[1]) - R) < 1.5 * h and (dim == 2 or abs(pos[2]) < 3 * h))
    t.refine(crit)
    t.corner_sweep()
    for b, s in sides:
        t.add_boundary(s, b)
    t.finalize()
    return t, sim, crit


@pytest.mark.parametrize("dim,nbox,minl,maxl", [(2, 2, 2, 6), (2, 3, 3, 5), (3, 2, 2, 5)])
def test_multi_box_domain(dim, nbox, minl, maxl):
    """several GfsBoxes: the locate array has one slot per box, refinement and
    2:1 / corner balance cross the box interfaces, stencils straddle them"""
    t, sim, crit = build_chain(dim, nbox, minl, maxl, 0.3)
    a = t.view()
    # the reference tree with exactly these cells, then flatten it
    sim2 = helpers.oracle_tree_from_flat(a)
    roots, is_box = sim2.roots()
    t2, fmap = capi.flatten_ftt(dim, roots, is_box)
    b = t2.view()
    assert a.n_box_roots == nbox and a.n_cells == b.n_cells
    for k in KEYS:
        assert np.array_equal(getattr(a, k), getattr(b, k)), k
    assert np.array_equal(a.la_n, b.la_n) and np.array_equal(a.la_slot, b.la_slot)
    assert a.la_n[0] == nbox + 2 and sorted(a.la_slot[a.la_slot >= 0]) == list(range(nbox))
    # stencils against gfs_cell_corner_interpolator, including those across the interface
    t2.build_stencils()
    idx = helpers.PtrIndex(fmap.cells)
    leaves = b.box_leaves
    near = leaves[np.abs(b.pos[leaves, 0] - 0.5) < 0.2]
    assert len(near) > 50
    for i in near[::3]:
        for k in range(2 ** dim):
            cells, w = t2.corner_interpolator(int(i), k)
            oc, ow = sim2.corner_interpolator(fmap.cells[i], k)
            assert list(idx(np.array(oc, dtype=np.uint64))) == cells and ow == w
    # point location across the boxes
    rng = np.random.default_rng(2)
    x = rng.uniform(-0.7, nbox - 0.3, 4000)
    y = rng.uniform(-0.7, 0.7, 4000)
    z = rng.uniform(-0.7, 0.7, 4000) if dim == 3 else None
    got = idx(sim2.locate(x, y, z))
    assert (got >= 0).sum() > 1000 and (got < 0).sum() > 100
    roots_of = got.copy()
    while True:
        p = np.where(roots_of >= 0, b.parent[np.maximum(roots_of, 0)], -1)
        m = p >= 0
        if not m.any():
            break
        roots_of[m] = p[m]
    assert set(np.unique(roots_of[got >= 0])) == set(range(nbox))

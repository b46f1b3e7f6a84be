"""CPU tests of the ORACLE itself (test infrastructure): hand-derived known
answers for the restated force/integrator layer, exactness properties of the
reference's interpolation on its own tree, and the committed golden fixtures
(tests/golden/*.npz, generated here by tests/golden/make_golden.py)."""
import glob
import os

import numpy as np
import pytest

import helpers
from helpers import capi, worlds, ora

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))
COLS = ("x", "y", "z", "vx", "vy", "vz", "mass", "volume")


def _world(name):
    return {"c1_l5": lambda: worlds.make_c1(level=5, n_particles=400),
            "tg_l4": lambda: worlds.make_c2(level=4, n_particles=400),
            "ring_3_6": lambda: worlds.make_ring("ring", 3, 6, 400, 3003)}[name]()


def test_golden_fixtures_exist():
    assert len(GOLDEN) == 3


@pytest.mark.parametrize("path", GOLDEN, ids=lambda p: os.path.basename(p)[:-4])
def test_oracle_reproduces_golden(path):
    """the prebuilt oracle/_ref library gives bit-identical answers to the
    fixtures generated where /root/reference exists"""
    g = np.load(path)
    w = _world(os.path.basename(path)[:-4])
    a = w.arrays
    assert (a.n_cells, a.n_leaves, a.n_vertices) == (int(g["n_cells"]), int(g["n_leaves"]), int(g["n_vertices"]))
    sim, ptrs = helpers.matched_oracle(w)
    idx = helpers.PtrIndex(ptrs)
    qz = g["qz"] if w.dim == 3 else None
    cell = idx(sim.locate(g["qx"], g["qy"], qz))
    lvl = np.where(cell >= 0, a.level[np.maximum(cell, 0)], -1)
    assert np.array_equal(lvl, g["loc_level"])
    inside = cell >= 0
    assert np.array_equal(a.pos[cell[inside]], g["loc_pos"][inside])
    for c in range(w.dim):
        assert np.array_equal(sim.interpolate(c, g["qx"], g["qy"], qz), g[f"interp{c}"])
    parts = {k: (g["p_" + k] if len(g["p_" + k]) else None) for k in COLS}
    for steps in (1, 10):
        cells, st = helpers.oracle_step(sim, ptrs, w, parts, steps=steps)
        for k in ("x", "y", "vx", "vy", "fx", "fy"):
            assert np.array_equal(st[k], g[f"s{steps}_{k}"]), (steps, k)
        assert np.array_equal(a.pos[cells], g[f"s{steps}_cell_pos"])


def test_interpolation_reproduces_linear_fields():
    """inverse-distance corner values + trilinear evaluation are exact for a
    linear field on a uniform tree (away from the hull)"""
    w = worlds.make_c2(level=4, n_particles=0)
    a = w.arrays
    coef = np.array([0.3, -1.7, 2.2])
    w.u[:] = a.pos @ coef + 0.5
    w.v[:] = 0.0
    w.w[:] = 0.0
    sim, ptrs = helpers.matched_oracle(w)
    rng = np.random.default_rng(0)
    p = rng.uniform(-0.4, 0.4, (500, 3))
    got = sim.interpolate(0, p[:, 0].copy(), p[:, 1].copy(), p[:, 2].copy())
    assert np.abs(got - (p @ coef + 0.5)).max() < 5e-15


def test_locate_ties_go_to_the_low_child():
    """strict '>' at every level (src/ftt.c:1559-1563): a point on a face
    belongs to the cell on the low side; the hull is inclusive at the root but
    the locate array makes the upper face fall outside (src/domain.c:47)"""
    w = worlds.make_c2(level=3, n_particles=0)
    a = w.arrays
    sim, ptrs = helpers.matched_oracle(w)
    idx = helpers.PtrIndex(ptrs)
    h = 1.0 / 8
    x = np.array([0.0, h, -h, -0.5, 0.5, np.nextafter(0.5, 0)])
    y = np.full_like(x, 0.01)
    z = np.full_like(x, 0.01)
    cell = idx(sim.locate(x, y, z))
    cx = np.where(cell >= 0, a.pos[np.maximum(cell, 0), 0], np.nan)
    assert cx[0] == -h / 2 and cx[1] == h / 2 and cx[2] == -1.5 * h
    assert cx[3] == -0.5 + h / 2          # lower hull face: inside
    assert cell[4] == -1                  # upper hull face: floor((0.5+0.5)/1) = 1 -> outside
    assert cell[5] == -1                  # 0.5 - ulp + 0.5 rounds to 1.0 -> outside too


def test_drag_known_answer_stokes():
    """one particle in a uniform stream: Re, cd and the velocity update worked
    out by hand from modules/particulatecommon.c:519-588, 828-839"""
    w = worlds.make_c2(level=3, n_particles=0)
    w.u[:] = 1.0
    w.v[:] = 0.0
    w.w[:] = 0.0
    w = worlds.World(**{**w.__dict__, "forces": (capi.FORCE_DRAG,), "mu": 1e-3, "dt": 1e-3})
    sim, ptrs = helpers.matched_oracle(w)
    d, rho_p = 1e-3, 1000.0
    vol = np.pi * d ** 3 / 6
    parts = dict(x=np.array([0.1]), y=np.array([0.2]), z=np.array([-0.3]), vx=np.array([0.5]),
                 vy=np.array([0.0]), vz=np.array([0.0]), mass=np.array([rho_p * vol]), volume=np.array([vol]))
    cells, st = helpers.oracle_step(sim, ptrs, w, parts)
    dia = 2.0 * (3.0 * vol / 4.0 / np.pi) ** (1.0 / 3.0)
    urel = 0.5
    Re = urel * dia * 1.0 / 1e-3
    cd = 16.0 * (1.0 + 0.15 * Re ** 0.5) / Re
    f = 3.0 / (4.0 * dia) * cd * urel * urel * 1.0
    F = f * vol
    vx = 0.5 + F * 1e-3 / (rho_p * vol)
    x = 0.1 + 0.5 * 1e-3 / 2 + vx * 1e-3 / 2
    assert abs(st["fx"][0] - F) <= 1e-14 * F
    assert abs(st["vx"][0] - vx) <= 1e-15 and abs(st["x"][0] - x) <= 1e-16
    assert st["fy"][0] == 0 and st["vy"][0] == 0 and st["y"][0] == 0.2


def test_buoyancy_and_lift_known_answer():
    """solid-body rotation U = -Omega y, V = Omega x has vorticity 2 Omega k;
    lift = rho cl (u_rel x omega), buoyancy = (m/V - rho) g"""
    omega = 0.7
    w = worlds.make_c2(level=4, n_particles=0)
    a = w.arrays
    w.u[:] = -omega * a.pos[:, 1]
    w.v[:] = omega * a.pos[:, 0]
    w.w[:] = 0.0
    w = worlds.World(**{**w.__dict__, "forces": (capi.FORCE_LIFT, capi.FORCE_BUOY), "g": (0.0, -2.0, 0.0)})
    sim, ptrs = helpers.matched_oracle(w)
    vol, m = 1e-9, 3e-9
    parts = dict(x=np.array([0.11]), y=np.array([-0.07]), z=np.array([0.03]), vx=np.array([0.0]),
                 vy=np.array([0.0]), vz=np.array([0.0]), mass=np.array([m]), volume=np.array([vol]))
    cells, st = helpers.oracle_step(sim, ptrs, w, parts)
    ux, uy = -omega * -0.07, omega * 0.11
    fx = 1.0 * 0.5 * (uy * 2 * omega) * vol
    fy = 1.0 * 0.5 * (-ux * 2 * omega) * vol + (m / vol - 1.0) * -2.0 * vol
    assert abs(st["fx"][0] - fx) <= 1e-13 * abs(fx)
    assert abs(st["fy"][0] - fy) <= 1e-13 * abs(fy)
    assert st["fz"][0] == 0


def test_reference_call_pattern_equals_fused():
    """pattern 0 (6 locates, dead vliq) and the fused variant give the same state"""
    w = worlds.make_ring("r", 3, 5, 300, 5)
    sim, ptrs = helpers.matched_oracle(w)
    parts = worlds.make_particles(w)
    _, a = helpers.oracle_step(sim, ptrs, w, parts, pattern=0)
    _, b = helpers.oracle_step(sim, ptrs, w, parts, pattern=1)
    for k in ("x", "y", "z", "vx", "vy", "vz", "fx", "fy", "fz"):
        assert np.array_equal(a[k], b[k])


def smoothed_deposit_numpy(a, dim, parts, force, rho, rkernel, kernel, fix_z=False):
    """Independent numpy restatement of source_particulate_event with its kernel
    (modules/particulatecommon.c:2087-2228): the visited leaves are those whose
    whole ancestor chain passes cond_kernel; returns (fields [dim][n_cells],
    correction, volume)."""
    n_cells = a.n_cells
    size = 0.5 ** a.level.astype(float) / 2.0
    radeq = size * (np.sqrt(3.0) if dim == 3 else np.sqrt(2.0))
    leaf = (a.child0 < 0) & ((a.flags & capi.CELL_DESTROYED) == 0)
    cellvol = (2 * size) ** dim
    out = np.zeros((dim, n_cells))
    n = len(parts["x"])
    corr, vol = np.zeros(n), np.zeros(n)
    order = np.argsort(a.level, kind="stable")
    for i in range(n):
        p = np.array([parts["x"][i], parts["y"][i], parts["z"][i] if dim == 3 else 0.0])
        d = a.pos - p
        if dim == 3:
            dist = np.sqrt(d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1] + d[:, 2] * d[:, 2])
        else:
            dist = np.sqrt(d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1])
        ok = dist - radeq <= rkernel
        inside = np.ones(n_cells, bool)
        for c in range(dim):
            inside &= ~((p[c] > a.pos[:, c] + size) | (p[c] < a.pos[:, c] - size))
        ok |= inside
        ok &= (a.flags & capi.CELL_DESTROYED) == 0
        reach = np.zeros(n_cells, bool)
        reach[:a.n_box_roots] = ok[:a.n_box_roots]
        for lv in range(a.min_level + 1, a.max_level + 1):
            sel = np.nonzero(a.level == lv)[0]
            reach[sel] = ok[sel] & reach[a.parent[sel]]
        vis = np.nonzero(reach & leaf)[0]
        rb = (3.0 * parts["volume"][i] / (4.0 * np.pi)) ** (1.0 / 3.0)
        q = (a.pos[vis] - p) / rb
        if dim == 2:
            q[:, 2] = 0.0
        elif not fix_z:
            q[:, 2] = (0.0 - p[2]) / rb
        r2 = q[:, 0] * q[:, 0] + q[:, 1] * q[:, 1] + q[:, 2] * q[:, 2]
        kind, ka, kb, kp = kernel
        if kind == ora.KERNEL_CONSTANT:
            K = np.full(len(vis), ka)
        elif kind == ora.KERNEL_GAUSSIAN:
            K = ka * np.exp(-kb * r2)
        else:
            t = 1.0 - kb * r2
            K = np.where(t > 0, ka * t ** kp, 0.0)
        vol[i] = cellvol[vis].sum()
        corr[i] = (K * cellvol[vis]).sum() / vol[i]
        if corr[i] > 1e-10:
            for c in range(dim):
                np.subtract.at(out[c], vis, force[c][i] / rho / cellvol[vis] * K / corr[i])
    return out, corr, vol


@pytest.mark.parametrize("name,rk,kernel", [
    ("c1_l5", 0.06, (ora.KERNEL_GAUSSIAN, 1.0, 2e-4)),
    ("ring_3_6", 0.05, (ora.KERNEL_COMPACT, 2.0, 1e-4, 2)),
    ("ring_3_6", 0.0, (ora.KERNEL_CONSTANT, 1.0, 0.0)),
])
def test_smoothed_deposit_matches_numpy_restatement(name, rk, kernel):
    """the oracle's conditional traversals (the reference's ftt_cell_traverse_condition
    object code + restated cond_kernel / kernel_volume / diffuse_force) against a
    brute-force numpy restatement over all cells"""
    w = _world(name)
    a = w.arrays
    sim, ptrs = helpers.matched_oracle(w)
    parts = worlds.make_particles(w, 60)
    plist = ora.ParticleList(sim, *[parts[k] for k in COLS])
    live = (a.flags & capi.CELL_DESTROYED) == 0
    for iv in range(4, 4 + w.dim):
        sim.set_values(iv, ptrs[live], np.zeros(int(live.sum())))
    kind, ka, kb = kernel[:3]
    kp = kernel[3] if len(kernel) > 3 else 1
    par = helpers.oracle_params(w)
    corr, vol = plist.deposit_force_smoothed(par, 4, rk, ora.Kernel(kind, ka, kb, kp, 0))
    st = plist.get()          # force = on-fluid force (drag[, lift])
    force = [st["fx"], st["fy"], st["fz"]]
    want, wcorr, wvol = smoothed_deposit_numpy(a, w.dim, parts, force, w.rho, rk, (kind, ka, kb, kp))
    assert np.array_equal(vol, wvol)                       # same set of visited leaves (exact sums)
    assert np.allclose(corr, wcorr, rtol=1e-13, atol=0)
    assert (corr > 1e-10).any()
    for c in range(w.dim):
        got = np.zeros(a.n_cells)
        got[live] = sim.get_values(4 + c, ptrs[live])
        assert np.abs(want[c]).max() > 0
        assert np.abs(got - want[c]).max() <= 1e-12 * np.abs(want[c]).max()
    if rk == 0.0:
        # rkernel = 0: only cells whose circumscribed sphere contains the particle
        assert vol.max() <= (2 ** w.dim + 8) * (0.5 ** a.min_level) ** w.dim

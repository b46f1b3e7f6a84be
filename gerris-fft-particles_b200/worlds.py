"""Synthetic worlds for the BASELINE.json configs.

Harness code shared by tests and bench.py: it builds the flat tree with the
native builders of libgfsb200 (no oracle, no reference code), samples the
analytic velocity fields at the exact cell centres, and draws the seeded
particle clouds.  Nothing here is on the timed path.

  C1  2D uniform level 6 + four Dirichlet ghost layers, lid-style field, drag
  C2  3D uniform level 7 (128^3), frozen Taylor-Green, drag + lift + buoyancy
  C3  3D adaptive levels 5-9 around a vortex ring (R = 0.25)
  C4  two-way variant (C3-style tree at 6-10 or the C2 tree), 50 M particles
  C5  3D adaptive levels 6-10, 200 M particles (scaling sweep)
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, Optional, Tuple

import numpy as np

from . import capi

LEAF, BOUNDARY = capi.CELL_LEAF, capi.CELL_BOUNDARY

# Particle properties of the 3D configs.  SURVEY.md section 8d proposed
# d_p in [1e-4, 1e-3], rho_p = 2.5 with mu = 1e-3 and dt = 1e-3; the Stokes
# response time rho_p d^2/(18 mu) is then 1.4e-6 .. 1.4e-4 s, i.e. dt/tau up to
# 720, and the reference's explicit update v += F dt/m (particulatecommon.c:835)
# amplifies the slip velocity ~6-700x per step: every particle leaves the
# domain within ~10 steps, on the reference CPU path just as on the GPU.  A
# benchmark on that would time an empty kernel.  These values keep
# dt/tau <= 0.02 (tau = 0.055 .. 0.89 s), particle Reynolds numbers O(0.1-50),
# and every particle inside the box for thousands of steps.
D_P = (1e-3, 4e-3)
RHO_P = 1000.0


@dataclass
class World:
    name: str
    dim: int
    tree: Optional[capi.Tree]          # None for a bare spec (no product tree built)
    arrays: Optional[capi.TreeArrays]
    u: Optional[np.ndarray]
    v: Optional[np.ndarray]
    w: Optional[np.ndarray]
    forces: Tuple[int, ...]
    dt: float
    rho: float = 1.0
    mu: float = 1e-3
    g: Tuple[float, float, float] = (0.0, 0.0, 0.0)
    seed: int = 0
    n_particles: int = 0
    meta: Dict = field(default_factory=dict)

    def step_params(self, **kw) -> capi.StepParams:
        return capi.StepParams(self.dt, self.forces, rho=self.rho, mu=self.mu, g=self.g, **kw)

    def level_histogram(self) -> Dict[int, int]:
        a = self.arrays
        lv = a.level[a.box_leaves]
        return {int(l): int((lv == l).sum()) for l in np.unique(lv)}


# ---------------------------------------------------------------------------
# analytic fields

def taylor_green(pos: np.ndarray):
    x, y, z = pos[:, 0], pos[:, 1], pos[:, 2]
    two_pi = 2.0 * np.pi
    u = np.sin(two_pi * x) * np.cos(two_pi * y) * np.cos(two_pi * z)
    v = -np.cos(two_pi * x) * np.sin(two_pi * y) * np.cos(two_pi * z)
    return u, v, np.zeros_like(u)


def vortex_ring(pos: np.ndarray, R: float = 0.25, gamma: float = 1.0, a: float = 0.05):
    """Gaussian-core ring in the plane z = 0: poloidal speed
    gamma/(2 pi s) (1 - exp(-s^2/a^2)), s = distance to the core circle."""
    x, y, z = pos[:, 0], pos[:, 1], pos[:, 2]
    rho = np.sqrt(x * x + y * y)
    dr = rho - R
    s2 = dr * dr + z * z
    # speed/s = gamma/(2 pi) * (1 - exp(-s2/a2))/s2, finite at s = 0
    with np.errstate(divide="ignore", invalid="ignore"):
        k = np.where(s2 > 1e-300, -np.expm1(-s2 / (a * a)) / s2, 1.0 / (a * a))
    k = gamma / (2.0 * np.pi) * k
    u_rho = -k * z
    u_z = k * dr
    with np.errstate(divide="ignore", invalid="ignore"):
        cx = np.where(rho > 0, x / rho, 0.0)
        cy = np.where(rho > 0, y / rho, 0.0)
    return u_rho * cx, u_rho * cy, u_z


def lid_style(pos: np.ndarray):
    """psi = (1/pi) sin^2(pi x') sin^2(pi y') y',  U = dpsi/dy, V = -dpsi/dx."""
    xp, yp = pos[:, 0] + 0.5, pos[:, 1] + 0.5
    sx, cx = np.sin(np.pi * xp), np.cos(np.pi * xp)
    sy, cy = np.sin(np.pi * yp), np.cos(np.pi * yp)
    u = sx * sx * (2.0 * sy * cy * yp + sy * sy / np.pi)
    v = -(2.0 * sx * cx * sy * sy * yp)
    return u, v


def apply_dirichlet_ghosts(arrays: capi.TreeArrays, val: np.ndarray, bc: Dict[int, float]):
    """Ghost leaves: value = 2*bc - interior neighbour (src/boundary.c:253-258).
    bc maps the box side (0:+x 1:-x 2:+y 3:-y 4:+z 5:-z) to the Dirichlet value."""
    a = arrays
    ghosts = np.nonzero((a.flags & (LEAF | BOUNDARY)) == (LEAF | BOUNDARY))[0]
    # which side a ghost cell belongs to: its root
    root = ghosts.copy()
    while True:
        p = a.parent[root]
        m = p >= 0
        if not m.any():
            break
        root[m] = p[m]
    # root r > 0 was added by add_boundary in side order; recover the side from its position
    for r in np.unique(root):
        d = a.pos[r] - a.pos[0]
        axis = int(np.argmax(np.abs(d)))
        side = 2 * axis + (0 if d[axis] > 0 else 1)
        g = ghosts[root == r]
        inner = a.neighbor[g, side ^ 1]
        ok = inner >= 0
        val[g[ok]] = 2.0 * bc.get(side, 0.0) - val[inner[ok]]


# ---------------------------------------------------------------------------
# particle clouds

def cloud_uniform(rng, n, dim, lo=-0.45, hi=0.45):
    return [rng.uniform(lo, hi, n) for _ in range(dim)]


def particle_props(rng, n, d_lo, d_hi, rho_p):
    d = rng.uniform(d_lo, d_hi, n) if d_hi > d_lo else np.full(n, d_lo)
    vol = np.pi * d ** 3 / 6.0
    return rho_p * vol, vol


# ---------------------------------------------------------------------------
# the configs

ALL3 = (capi.FORCE_DRAG, capi.FORCE_LIFT, capi.FORCE_BUOY)


def spec(name: str, n_particles: Optional[int] = None) -> World:
    """The config's parameters without building any tree (tree/arrays/fields = None):
    what the reference arm of bench.py needs to draw the same particle cloud."""
    ring = dict(d_p=D_P, rho_p=RHO_P, v0="fluid", field="ring",
                cloud="half uniform, half gaussian(0.08) around the core")
    table = {
        "C1": dict(dim=2, forces=(capi.FORCE_DRAG,), dt=1e-2, seed=1001, n=1000,
                   meta=dict(d_p=(1e-3, 1e-3), rho_p=1000.0, v0="zero", field="lid", level=6)),
        "C2": dict(dim=3, forces=ALL3, dt=1e-3, g=(0.0, -1.0, 0.0), seed=2002, n=10_000_000,
                   meta=dict(d_p=D_P, rho_p=RHO_P, v0="fluid", field="tg", level=7)),
        "C3": dict(dim=3, forces=ALL3, dt=1e-3, g=(0.0, -1.0, 0.0), seed=3003, n=10_000_000,
                   meta=dict(ring, levels=(5, 9))),
        "C4": dict(dim=3, forces=ALL3, dt=1e-3, g=(0.0, -1.0, 0.0), seed=4004, n=50_000_000,
                   meta=dict(ring, levels=(6, 10))),
        "C5": dict(dim=3, forces=ALL3, dt=1e-3, g=(0.0, -1.0, 0.0), seed=5005, n=200_000_000,
                   meta=dict(ring, levels=(6, 10))),
    }
    t = table[name]
    return World(name, t["dim"], None, None, None, None, None, t["forces"], t["dt"], mu=1e-3,
                 g=t.get("g", (0.0, 0.0, 0.0)), seed=t["seed"],
                 n_particles=t["n"] if n_particles is None else n_particles, meta=dict(t["meta"]))


def field_of(world: World, pos: np.ndarray):
    """analytic (u, v, w) of the world's config at positions pos[n,3]"""
    kind = world.meta.get("field")
    if kind == "tg":
        return taylor_green(pos)
    if kind == "ring":
        return vortex_ring(pos)
    if kind == "lid":
        u, v = lid_style(pos)
        return u, v, None
    raise KeyError(kind)


def _finish(tree: capi.Tree) -> capi.TreeArrays:
    tree.finalize()
    tree.build_stencils()
    return tree.view()


def make_c1(level: int = 6, n_particles: int = 1000) -> World:
    t = capi.Tree(2)
    t.refine_uniform(level)
    for side in range(4):
        t.add_boundary(side)
    a = _finish(t)
    u, v = lid_style(a.pos)
    u[a.flags & BOUNDARY != 0] = 0.0
    v[a.flags & BOUNDARY != 0] = 0.0
    apply_dirichlet_ghosts(a, u, {2: 1.0})      # tangential velocity 1 on the lid (top)
    apply_dirichlet_ghosts(a, v, {})
    return World("C1", 2, t, a, u, v, None, (capi.FORCE_DRAG,), dt=1e-2, mu=1e-3,
                 seed=1001, n_particles=n_particles,
                 meta=dict(d_p=(1e-3, 1e-3), rho_p=1000.0, v0="zero", field="lid", level=level))


def make_c2(level: int = 7, n_particles: int = 10_000_000) -> World:
    t = capi.Tree(3)
    t.refine_uniform(level)
    a = _finish(t)
    u, v, w = taylor_green(a.pos)
    return World("C2", 3, t, a, u, v, w, (capi.FORCE_DRAG, capi.FORCE_LIFT, capi.FORCE_BUOY),
                 dt=1e-3, mu=1e-3, g=(0.0, -1.0, 0.0), seed=2002, n_particles=n_particles,
                 meta=dict(d_p=D_P, rho_p=RHO_P, v0="fluid", field="tg", level=level))


def make_ring(name: str, minlevel: int, maxlevel: int, n_particles: int, seed: int) -> World:
    t = capi.Tree(3)
    t.refine_ring(minlevel, maxlevel, 0.25, 1.5)
    t.corner_sweep()
    a = _finish(t)
    u, v, w = vortex_ring(a.pos)
    return World(name, 3, t, a, u, v, w, (capi.FORCE_DRAG, capi.FORCE_LIFT, capi.FORCE_BUOY),
                 dt=1e-3, mu=1e-3, g=(0.0, -1.0, 0.0), seed=seed, n_particles=n_particles,
                 meta=dict(d_p=D_P, rho_p=RHO_P, v0="fluid", field="ring",
                           levels=(minlevel, maxlevel),
                           cloud="half uniform, half gaussian(0.08) around the core"))


def make_c3(minlevel: int = 5, maxlevel: int = 9, n_particles: int = 10_000_000) -> World:
    return make_ring("C3", minlevel, maxlevel, n_particles, 3003)


def make_c4(minlevel: int = 6, maxlevel: int = 10, n_particles: int = 50_000_000) -> World:
    return make_ring("C4", minlevel, maxlevel, n_particles, 4004)


def make_c5(minlevel: int = 6, maxlevel: int = 10, n_particles: int = 200_000_000) -> World:
    return make_ring("C5", minlevel, maxlevel, n_particles, 5005)


def make_particles(world: World, n: Optional[int] = None, rank: int = 0, n_ranks: int = 1):
    """Seeded particle cloud of the world's config; rank r of n_ranks draws its
    own contiguous share with an independent stream of the same seed."""
    n_total = world.n_particles if n is None else n
    lo = n_total * rank // n_ranks
    hi = n_total * (rank + 1) // n_ranks
    n_local = hi - lo
    rng = np.random.default_rng([world.seed, rank, n_ranks] if n_ranks > 1 else world.seed)
    dim = world.dim
    m = world.meta
    if "cloud" in m:
        n_uni = n_local // 2
        pos = cloud_uniform(rng, n_uni, 3)
        n_g = n_local - n_uni
        theta = rng.uniform(0.0, 2.0 * np.pi, n_g)
        off = rng.normal(0.0, 0.08, (3, n_g))
        gx = 0.25 * np.cos(theta) + off[0]
        gy = 0.25 * np.sin(theta) + off[1]
        gz = off[2]
        pos = [np.concatenate([pos[0], gx]), np.concatenate([pos[1], gy]), np.concatenate([pos[2], gz])]
        pos = [np.clip(p, -0.49, 0.49) for p in pos]
    else:
        pos = cloud_uniform(rng, n_local, dim)
    mass, vol = particle_props(rng, n_local, m["d_p"][0], m["d_p"][1], m["rho_p"])
    if m["v0"] == "zero":
        vel = [np.zeros(n_local) for _ in range(dim)]
    else:
        P = np.stack(pos + ([np.zeros(n_local)] if dim == 2 else []), axis=1)
        vel = list(field_of(world, P))[:dim]
    if dim == 2:
        return dict(x=pos[0], y=pos[1], z=None, vx=vel[0], vy=vel[1], vz=None, mass=mass, volume=vol)
    return dict(x=pos[0], y=pos[1], z=pos[2], vx=vel[0], vy=vel[1], vz=vel[2], mass=mass, volume=vol)


def adversarial_points(arrays: capi.TreeArrays, rng, n: int = 2000):
    """Points on cell faces, centres, vertices, the hull and just outside it --
    where a strict-'>' descent and an integer-quantised one would disagree."""
    a = arrays
    dim = a.dim
    leaves = a.box_leaves
    pick = leaves[rng.integers(0, len(leaves), n)]
    c = a.pos[pick][:, :dim]
    h = a.h[pick][:, None]
    kinds = rng.integers(0, 6, n)
    off = rng.integers(-1, 2, (n, dim)).astype(np.float64)          # -1, 0, +1 half sizes
    pts = c + 0.5 * h * off                                         # centres / faces / edges / vertices
    eps = np.ldexp(1.0, -52) * rng.integers(-2, 3, (n, dim))
    pts = np.where((kinds == 1)[:, None], pts + eps * np.abs(pts), pts)     # +- a few ulps
    tiny = rng.choice([0.0, 1e-300, -1e-300, 5e-324, 1e-17, -1e-17], (n, dim))
    pts = np.where((kinds == 2)[:, None], tiny, pts)                # around the root centre
    hull = rng.choice([-0.5, 0.5, np.nextafter(0.5, 0), np.nextafter(-0.5, 0), np.nextafter(0.5, 1),
                       np.nextafter(-0.5, -1)], (n, dim))
    mix = rng.uniform(-0.5, 0.5, (n, dim))
    sel = rng.integers(0, 2, (n, dim)).astype(bool)
    pts = np.where((kinds == 3)[:, None], np.where(sel, hull, mix), pts)    # on / next to the hull
    far = rng.uniform(-3.0, 3.0, (n, dim))
    pts = np.where((kinds == 4)[:, None], far, pts)                 # mostly outside
    cols = [np.ascontiguousarray(pts[:, k]) for k in range(dim)]
    return cols if dim == 3 else cols + [None]

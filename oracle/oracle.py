"""TEST INFRASTRUCTURE ONLY -- ctypes wrapper of oracle/_ref/libgfsoracle{2D,3D}.so.

The oracle is the reference's own src/ftt.c + src/fluid.c object code plus the
restated particulate layer of oracle/particulate_port.c (see its header for the
file:line map and the "parity unpinned" note).  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product path never does.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
FORCE_DRAG, FORCE_LIFT, FORCE_BUOY, FORCE_INERTIAL, FORCE_ADDEDMASS = 1, 2, 3, 4, 5
KERNEL_CONSTANT, KERNEL_GAUSSIAN, KERNEL_COMPACT = 0, 1, 2
KERNEL_ODD = 3            # a/(1 + b r^2): reference object code only (libgfsrefobj / libgfsrefmod)


class Kernel(C.Structure):
    """OraKernel: closed-form smoothing kernels of GfsSourceParticulate"""
    _fields_ = [("kind", C.c_int), ("a", C.c_double), ("b", C.c_double), ("p", C.c_int), ("flags", C.c_int)]


class StepParams(C.Structure):
    _fields_ = [
        ("dt", C.c_double), ("n_forces", C.c_int), ("force", C.c_int * 8),
        ("rho", C.c_double), ("ivar_alpha", C.c_int), ("mu", C.c_double), ("ivar_mu", C.c_int),
        ("g", C.c_double * 3), ("cd_const", C.c_double), ("cl_const", C.c_double),
        ("pattern", C.c_int), ("ivar_uold", C.c_int), ("cm_const", C.c_double),
    ]


def step_params(dt, forces, rho=1.0, mu=0.0, g=(0.0, 0.0, 0.0), cd_const=float("nan"),
                cl_const=float("nan"), pattern=0, ivar_alpha=-1, ivar_mu=-1, ivar_uold=5,
                cm_const=float("nan")) -> StepParams:
    p = StepParams()
    p.dt = dt
    p.n_forces = len(forces)
    for k, f in enumerate(forces):
        p.force[k] = int(f)
    p.rho, p.mu, p.ivar_alpha, p.ivar_mu = rho, mu, ivar_alpha, ivar_mu
    for a in range(3):
        p.g[a] = float(g[a])
    p.cd_const, p.cl_const, p.pattern = cd_const, cl_const, pattern
    p.ivar_uold, p.cm_const = ivar_uold, cm_const
    return p


_libs = {}


def available(dim: int) -> bool:
    return os.path.exists(os.path.join(_HERE, "_ref", f"libgfsoracle{dim}D.so"))


def load(dim: int) -> C.CDLL:
    if dim in _libs:
        return _libs[dim]
    path = os.path.join(_HERE, "_ref", f"libgfsoracle{dim}D.so")
    L = C.CDLL(path)
    vp, u64, dbl, lng, i32 = C.c_void_p, C.c_uint64, C.c_double, C.c_long, C.c_int
    sig = {
        "ora_sim_new": (vp, [i32]), "ora_sim_destroy": (None, [vp]),
        "ora_root": (u64, [vp]), "ora_boundary_root": (u64, [vp, i32]), "ora_dimension": (i32, []),
        "ora_add_box": (i32, [vp]), "ora_nbox": (i32, [vp]), "ora_box_root": (u64, [vp, i32]),
        "ora_box_boundary_root": (u64, [vp, i32, i32]), "ora_add_boundary_box": (None, [vp, i32, i32]),
        "ora_refine_uniform": (None, [vp, i32]), "ora_refine_ring": (None, [vp, i32, i32, dbl, dbl]),
        "ora_refine_points": (i32, [vp, i32, vp, vp, vp, vp]),
        "ora_corner_sweep": (None, [vp]), "ora_add_boundary": (None, [vp, i32]),
        "ora_match_boundaries": (None, [vp]), "ora_finalize": (None, [vp]),
        "ora_locate_array": (None, [vp, vp, vp, vp]),
        "ora_locate": (None, [vp, lng, vp, vp, vp, vp]),
        "ora_locate_one": (vp, [vp, dbl, dbl, dbl, i32]),
        "ora_set_solid": (None, [u64, dbl, vp, vp]),
        "ora_cell_info": (None, [u64, vp, vp, vp, vp]),
        "ora_set_values": (None, [vp, i32, lng, vp, vp]), "ora_get_values": (None, [vp, i32, lng, vp, vp]),
        "ora_neighbor": (u64, [u64, i32]), "ora_count": (lng, [vp, i32]),
        "ora_export_cells": (lng, [vp, vp, vp, vp, vp]),
        "ora_interpolate": (None, [vp, i32, lng, vp, vp, vp, vp]),
        "ora_corner_interpolator": (i32, [vp, u64, i32, vp, vp]),
        "ora_corner_values": (None, [vp, i32, lng, vp, vp]),
        "ora_center_gradient": (None, [vp, i32, i32, lng, vp, vp]),
        "ora_vorticity": (None, [vp, lng, vp, vp]),
        "ora_list_new": (vp, [lng] + [vp] * 8), "ora_list_destroy": (None, [vp]),
        "ora_list_size": (lng, [vp]), "ora_list_get": (None, [vp] + [vp] * 10),
        "ora_list_get_mass": (None, [vp, vp]),
        "ora_list_cull": (lng, [vp, vp]),
        "ora_list_bc": (lng, [vp, vp, C.c_uint]),
        "ora_list_step": (None, [vp, vp, C.POINTER(StepParams), i32]),
        "ora_deposit_volume": (None, [vp, vp, i32]),
        "ora_deposit_force": (None, [vp, vp, C.POINTER(StepParams), i32]),
        "ora_deposit_force_smoothed": (None, [vp, vp, C.POINTER(StepParams), i32, dbl, C.POINTER(Kernel), vp, vp]),
        "ora_advect_points": (None, [vp, lng, vp, vp, vp, dbl]),
        "ora_list_write": (i32, [vp, C.c_char_p, dbl]),
        "ora_max_threads": (i32, []),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype, f.argtypes = res, args
    _libs[dim] = L
    return L


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


class Sim:
    """One single-box reference domain (OraSim) with nvar cell variables;
    variables 0,1,2 are U,V,W (U,V in 2D)."""

    def __init__(self, dim: int, nvar: int = 8):
        self.dim, self.L = dim, load(dim)
        self.h = C.c_void_p(self.L.ora_sim_new(nvar))
        self.nvar = nvar
        self.sides = []

    def __del__(self):
        try:
            if self.h:
                self.L.ora_sim_destroy(self.h)
                self.h = None
        except Exception:
            pass

    # construction, in Gerris' order: boxes, boundaries, refinement, corner sweep, match
    def add_box(self):
        """append a unit GfsBox to the right (+x) of the last one"""
        return self.L.ora_add_box(self.h)

    def add_boundary(self, side, box=0):
        self.L.ora_add_boundary_box(self.h, box, side)
        self.sides.append((box, side))

    def refine_uniform(self, level):
        self.L.ora_refine_uniform(self.h, level)

    def refine_ring(self, minlevel, maxlevel, R=0.25, factor=1.5):
        self.L.ora_refine_ring(self.h, minlevel, maxlevel, R, factor)

    def corner_sweep(self):
        self.L.ora_corner_sweep(self.h)

    def finalize(self):
        self.L.ora_match_boundaries(self.h)
        self.L.ora_finalize(self.h)

    def roots(self):
        """(FttCell* list, is_box list) in flat-tree root order"""
        nb = self.L.ora_nbox(self.h)
        r = [self.L.ora_box_root(self.h, b) for b in range(nb)] + \
            [self.L.ora_box_boundary_root(self.h, b, s) for b, s in self.sides]
        return r, [1] * nb + [0] * len(self.sides)

    def count(self, leaves_only=False):
        return self.L.ora_count(self.h, int(leaves_only))

    def export_cells(self):
        """(ptr, pos[n,3], level, is_leaf) of every GfsBox cell, pre-order"""
        n = self.count()
        ptr = np.zeros(n, dtype=np.uint64)
        pos = np.zeros((n, 3))
        level = np.zeros(n, dtype=np.int32)
        leaf = np.zeros(n, dtype=np.int32)
        m = self.L.ora_export_cells(self.h, _p(ptr), _p(pos), _p(level), _p(leaf))
        assert m == n
        return ptr, pos, level, leaf.astype(bool)

    # queries
    def locate(self, x, y, z=None):
        x, y, z = _f64(x), _f64(y), _f64(z)
        out = np.zeros(len(x), dtype=np.uint64)
        self.L.ora_locate(self.h, len(x), _p(x), _p(y), _p(z), _p(out))
        return out

    def set_solid(self, cell, a, cm, s=None):
        """make `cell` (FttCell*) a mixed cell with fluid fraction a, centre of mass cm and face
        fractions s (None: all 1); a <= 0: not mixed any more"""
        v = np.zeros(3)
        v[:len(cm)] = cm
        sv = None if s is None else np.ascontiguousarray(list(s) + [1.0] * (6 - len(s)), dtype=np.float64)
        self.L.ora_set_solid(int(cell), float(a), _p(v), _p(sv))

    def set_values(self, ivar, cells, vals):
        cells = np.ascontiguousarray(cells, dtype=np.uint64)
        vals = _f64(vals)
        self.L.ora_set_values(self.h, ivar, len(cells), _p(cells), _p(vals))

    def get_values(self, ivar, cells):
        cells = np.ascontiguousarray(cells, dtype=np.uint64)
        out = np.empty(len(cells))
        self.L.ora_get_values(self.h, ivar, len(cells), _p(cells), _p(out))
        return out

    def interpolate(self, ivar, x, y, z=None):
        x, y, z = _f64(x), _f64(y), _f64(z)
        out = np.empty(len(x))
        self.L.ora_interpolate(self.h, ivar, len(x), _p(x), _p(y), _p(z), _p(out))
        return out

    def corner_interpolator(self, cell, k):
        cells = (C.c_uint64 * 29)()
        w = (C.c_double * 29)()
        n = self.L.ora_corner_interpolator(self.h, int(cell), k, cells, w)
        return list(cells[:n]), list(w[:n])

    def corner_values(self, ivar, cells):
        cells = np.ascontiguousarray(cells, dtype=np.uint64)
        out = np.empty((len(cells), 2 ** self.dim))
        self.L.ora_corner_values(self.h, ivar, len(cells), _p(cells), _p(out))
        return out

    def center_gradient(self, comp, ivar, cells):
        cells = np.ascontiguousarray(cells, dtype=np.uint64)
        out = np.empty(len(cells))
        self.L.ora_center_gradient(self.h, comp, ivar, len(cells), _p(cells), _p(out))
        return out

    def vorticity(self, cells):
        cells = np.ascontiguousarray(cells, dtype=np.uint64)
        out = np.empty((len(cells), 3))
        self.L.ora_vorticity(self.h, len(cells), _p(cells), _p(out))
        return out

    def advect_points(self, x, y, z, dt):
        x, y, z = _f64(x).copy(), _f64(y).copy(), None if z is None else _f64(z).copy()
        self.L.ora_advect_points(self.h, len(x), _p(x), _p(y), _p(z), dt)
        return x, y, z


class ParticleList:
    """GfsParticleList stand-in: one heap object per particle."""

    def __init__(self, sim: Sim, x, y, z, vx, vy, vz, mass, volume):
        self.sim = sim
        a = [_f64(v) for v in (x, y, z, vx, vy, vz, mass, volume)]
        self.h = C.c_void_p(sim.L.ora_list_new(len(a[0]), *[_p(v) for v in a]))

    def __del__(self):
        try:
            if self.h:
                self.sim.L.ora_list_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def __len__(self):
        return self.sim.L.ora_list_size(self.h)

    def cull(self):
        return self.sim.L.ora_list_cull(self.sim.h, self.h)

    def bc(self, periodic_mask=0):
        """gfs_particle_bc: wrap through periodic sides (bit d of the mask), drop the rest"""
        return self.sim.L.ora_list_bc(self.sim.h, self.h, periodic_mask)

    def step(self, params: StepParams, nthreads=1):
        self.sim.L.ora_list_step(self.sim.h, self.h, C.byref(params), nthreads)

    def get(self):
        n = len(self)
        d3 = self.sim.dim == 3
        out = {k: np.empty(n) for k in ("x", "y", "vx", "vy", "fx", "fy", "fz")}
        out["z"] = np.empty(n) if d3 else None
        out["vz"] = np.empty(n) if d3 else None
        ids = np.empty(n, dtype=np.uint32)
        self.sim.L.ora_list_get(self.h, _p(out["x"]), _p(out["y"]), _p(out["z"]), _p(out["vx"]),
                                _p(out["vy"]), _p(out["vz"]), _p(out["fx"]), _p(out["fy"]),
                                _p(out["fz"]), _p(ids))
        out["id"] = ids
        out["mass"] = np.empty(n)
        self.sim.L.ora_list_get_mass(self.h, _p(out["mass"]))
        return out

    def deposit_volume(self, ivar):
        self.sim.L.ora_deposit_volume(self.sim.h, self.h, ivar)

    def deposit_force(self, params: StepParams, ivar0):
        self.sim.L.ora_deposit_force(self.sim.h, self.h, C.byref(params), ivar0)

    def deposit_force_smoothed(self, params: StepParams, ivar0, rkernel, kernel: Kernel):
        """source_particulate_event with its smoothing kernel; returns the per-particle
        (correction, volume) normalisation"""
        n = len(self)
        corr, vol = np.empty(n), np.empty(n)
        self.sim.L.ora_deposit_force_smoothed(self.sim.h, self.h, C.byref(params), ivar0, rkernel,
                                              C.byref(kernel), _p(corr), _p(vol))
        return corr, vol

    def write(self, path, L=1.0):
        """the particle block as the reference writes it into a .gfs / dump file"""
        assert self.sim.L.ora_list_write(self.h, str(path).encode(), L) == 0


# ---------------------------------------------------------------------------
# libgfsrefobj: the reference's particulate layer itself as object code
# (modules/particulatecommon.c + src/event.c + src/particle.c + src/fluid.c +
# src/ftt.c compiled unmodified; run-time in oracle/refobj/glue.c).  It runs on
# the trees of a Sim above, shared by pointer.

_reflibs = {}
_LOCATE_T = C.CFUNCTYPE(C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_int)


def refobj_available(dim: int, module: bool = False) -> bool:
    return os.path.exists(os.path.join(_HERE, "_ref", f"libgfsref{'mod' if module else 'obj'}{dim}D.so"))


def load_refobj(dim: int, module: bool = False) -> C.CDLL:
    """module=False: libgfsrefobj, the unmodified reference.  module=True:
    libgfsrefmod, the same objects and run-time with the drop-in GModule
    (gerris-fft-particles_b200/host/particulates_b200.c) linked in and its
    g_module_check_init() run, i.e. the three hot-path events routed to the GPU."""
    key = (dim, module)
    if key in _reflibs:
        return _reflibs[key]
    R = C.CDLL(os.path.join(_HERE, "_ref", f"libgfsref{'mod' if module else 'obj'}{dim}D.so"))
    vp, dbl, lng, i32 = C.c_void_p, C.c_double, C.c_long, C.c_int
    sig = {
        "refobj_dimension": (i32, []),
        "refobj_sim_new": (vp, [i32, vp, vp, C.c_uint, vp, vp, vp, i32]),
        "refobj_sim_configure": (None, [vp, C.POINTER(StepParams), i32]),
        "refobj_sim_destroy": (None, [vp]),
        "refobj_locate": (None, [vp, lng, vp, vp, vp, vp]),
        "refobj_locate_array": (None, [vp, vp, vp, vp]),
        "refobj_sim_time": (None, [vp, C.POINTER(dbl), C.POINTER(i32)]),
        "refobj_sim_add_solid": (None, [vp, i32]),
        "refobj_sim_refine": (i32, [vp, vp]),
        "refobj_sim_coarsen": (i32, [vp, vp]),
        "refobj_sim_destroy_cell": (i32, [vp, vp]),
        "refobj_list_new": (vp, [vp, lng] + [vp] * 8 + [C.POINTER(StepParams)]),
        "refobj_list_destroy": (None, [vp]),
        "refobj_list_size": (lng, [vp]),
        "refobj_list_get": (None, [vp] * 12),
        "refobj_list_event": (i32, [vp, vp, i32]),
        "refobj_list_outside": (lng, [vp, vp]),
        "refobj_force": (None, [vp, lng, i32, vp]),
        "refobj_list_bc": (None, [vp]),
        "refobj_list_write": (i32, [vp, C.c_char_p]),
        "refobj_field_event": (None, [vp, vp, i32]),
        "refobj_source_event": (None, [vp, vp, i32, dbl, C.POINTER(Kernel)]),
        "refobj_warnings": (lng, []),
        "refobj_list_class_write": (i32, [vp, C.c_char_p]),
        "refobj_list_sync": (i32, [vp]),
        "refobj_module_init": (i32, []),
        "refobj_module_name": (C.c_char_p, []),
    }
    for name, (res, args) in sig.items():
        f = getattr(R, name)
        f.restype, f.argtypes = res, args
    assert R.refobj_dimension() == dim
    if module:
        assert R.refobj_module_init() == 1 and R.refobj_module_name() == b"particulates"
    else:
        assert R.refobj_module_init() == 0
    _reflibs[key] = R
    return R


class RefSim:
    """A GfsSimulation of the reference's own structs around the GfsBox /
    GfsBoundary trees of `sim`; gfs_domain_locate and its GfsLocateArray are the
    reference's own (src/domain.c compiled unmodified).  Bit d of periodic_mask: the boundaries on side d are
    GfsBoundaryPeriodic."""

    def __init__(self, sim: Sim, periodic_mask: int = 0, module: bool = False):
        self.sim, self.dim = sim, sim.dim
        self.R = load_refobj(sim.dim, module)
        nb = sim.L.ora_nbox(sim.h)
        nn = 2 * sim.dim
        roots = (C.c_void_p * nb)(*[sim.L.ora_box_root(sim.h, b) for b in range(nb)])
        broots = (C.c_void_p * (nb * nn))()
        for b, s in sim.sides:
            broots[b * nn + s] = sim.L.ora_box_boundary_root(sim.h, b, s)
        locate = C.cast(sim.L.ora_locate_one, C.c_void_p)
        self.h = C.c_void_p(self.R.refobj_sim_new(nb, roots, broots, periodic_mask, locate, None,
                                                  sim.h, sim.nvar))
        self._lists = []

    def configure(self, params: StepParams, timers=True):
        """PhysicalParams alpha, SourceViscosity, Source g, dt -- what the .gfs
        file declares around the list"""
        self.R.refobj_sim_configure(self.h, C.byref(params), int(timers))

    def locate(self, x, y, z=None):
        """gfs_domain_locate of the reference's own src/domain.c: FttCell* (0 = outside)"""
        x, y, z = _f64(x), _f64(y), _f64(z)
        out = np.zeros(len(x), dtype=np.uint64)
        self.R.refobj_locate(self.h, len(x), _p(x), _p(y), _p(z), _p(out))
        return out

    def locate_array(self):
        """(min[dim], h, n[dim]) of the GfsLocateArray the reference built"""
        mn, h, n = np.zeros(3), C.c_double(), np.zeros(3, dtype=np.int32)
        self.R.refobj_locate_array(self.h, _p(mn), C.byref(h), _p(n))
        return mn[:self.dim], h.value, n[:self.dim]

    def add_solid(self, moving=False):
        """an entry in sim->solids, as a GfsSolid (or GfsSolidMoving) declaration leaves"""
        self.R.refobj_sim_add_solid(self.h, int(moving))

    def refine(self, cell: int) -> bool:
        """ftt_cell_refine_single + gfs_cell_fine_init on a leaf (FttCell*): what an adapt does"""
        return bool(self.R.refobj_sim_refine(self.h, C.c_void_p(int(cell))))

    def destroy_cell(self, cell: int) -> bool:
        """ftt_cell_destroy: an entirely solid cell (what gfs_init_solid_fractions does)"""
        return bool(self.R.refobj_sim_destroy_cell(self.h, C.c_void_p(int(cell))))

    def coarsen(self, cell: int) -> bool:
        """ftt_cell_coarsen + gfs_cell_cleanup of the children of a non-leaf cell"""
        return bool(self.R.refobj_sim_coarsen(self.h, C.c_void_p(int(cell))))

    def time(self):
        t, i = C.c_double(), C.c_int()
        self.R.refobj_sim_time(self.h, C.byref(t), C.byref(i))
        return t.value, i.value

    def close(self):
        for l in list(self._lists):
            l.close()
        if self.h:
            self.R.refobj_sim_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class RefParticleList:
    """A GfsParticleList of GfsParticulate objects with its GfsForce* list, as
    gfs_particle_list_read leaves it; ids are 1..n in list order."""

    def __init__(self, rsim: RefSim, x, y, z, vx, vy, vz, mass, volume, params: StepParams):
        self.rsim, self.R = rsim, rsim.R
        a = [_f64(v) for v in (x, y, z, vx, vy, vz, mass, volume)]
        self.h = C.c_void_p(self.R.refobj_list_new(rsim.h, len(a[0]), *[_p(v) for v in a], C.byref(params)))
        rsim._lists.append(self)

    def close(self):
        if self.h:
            self.R.refobj_list_destroy(self.h)
            self.h = None
            self.rsim._lists.remove(self)

    def __len__(self):
        return self.R.refobj_list_size(self.h)

    def event(self, steps=1):
        """gfs_event_do on the list `steps` times (cull, GfsParticulate events,
        gfs_particle_bc, Un snapshot), the time level advancing in between"""
        return self.R.refobj_list_event(self.rsim.h, self.h, steps)

    def outside(self):
        return self.R.refobj_list_outside(self.rsim.h, self.h)

    def bc(self):
        self.R.refobj_list_bc(self.h)

    def force(self, index, k):
        """force model k of the list on particle `index`, per unit volume"""
        f = np.zeros(3)
        self.R.refobj_force(self.h, index, k, _p(f))
        return f

    def get(self, sync=True):
        """the state held by the GfsParticulate objects.  sync: call gfsb200_module_sync first, as
        host code reading the objects of a device-resident list has to (a no-op in the reference)"""
        if sync:
            self.sync()
        n = len(self)
        out = {k: np.empty(n) for k in ("x", "y", "z", "vx", "vy", "vz", "fx", "fy", "fz", "mass")}
        ids = np.empty(n, dtype=np.uint32)
        self.R.refobj_list_get(self.h, *[_p(out[k]) for k in ("x", "y", "z", "vx", "vy", "vz", "fx", "fy",
                                                               "fz", "mass")], _p(ids))
        out["id"] = ids
        return out

    def field_event(self, ivar):
        """GfsParticulateField: reset + V_p/V_cell scatter into variable ivar"""
        self.R.refobj_field_event(self.rsim.h, self.h, ivar)

    def source_event(self, ivar0, rkernel, kernel: Kernel):
        """GfsSourceParticulate: kernel-smoothed -F/rho/V_cell into ivar0.."""
        self.R.refobj_source_event(self.rsim.h, self.h, ivar0, rkernel, C.byref(kernel))

    def write(self, path):
        assert self.R.refobj_list_write(self.h, str(path).encode()) == 0

    def class_write(self, path):
        """the GfsParticleList's own write method: the list as it appears in a simulation dump"""
        assert self.R.refobj_list_class_write(self.h, str(path).encode()) == 0

    def sync(self):
        """module library only: gfsb200_module_sync, what foreign host code calls before it reads
        the objects of a list driven in resident mode; returns False in the plain reference"""
        return bool(self.R.refobj_list_sync(self.h))

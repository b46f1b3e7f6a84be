"""ctypes binding of the C-ABI in include/gfsb200.h (lib/libgfsb200.so).

This is harness code for tests and bench.py: the product is the shared
library; nothing here computes.  Every call goes straight to the C entry
point of the same name and raises GfsB200Error (with gfsb200_last_error())
on a non-zero status.  There is no CPU fallback: a missing library or a
missing GPU is an error.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_DIR = os.path.join(_HERE, "lib")

FORCE_DRAG, FORCE_LIFT, FORCE_BUOY, FORCE_INERTIAL, FORCE_ADDEDMASS = 1, 2, 3, 4, 5
CELL_DESTROYED, CELL_BOUNDARY, CELL_LEAF = 1, 2, 4
KERNEL_CONSTANT, KERNEL_GAUSSIAN, KERNEL_COMPACT = 0, 1, 2
KERNEL_FIX_Z = 1
MAX_FORCES = 8
NODATA = float(np.finfo(np.float64).max)


class GfsB200Error(RuntimeError):
    pass


class TreeView(C.Structure):
    _fields_ = [
        ("dim", C.c_int32), ("n_cells", C.c_int32), ("n_roots", C.c_int32), ("n_box_roots", C.c_int32),
        ("min_level", C.c_int32), ("max_level", C.c_int32), ("complete_level", C.c_int32),
        ("n_leaves", C.c_int64),
        ("level_start", C.POINTER(C.c_int32)), ("parent", C.POINTER(C.c_int32)),
        ("child0", C.POINTER(C.c_int32)), ("neighbor", C.POINTER(C.c_int32)),
        ("level", C.POINTER(C.c_uint8)), ("flags", C.POINTER(C.c_uint8)), ("pos", C.POINTER(C.c_double)),
        ("la_min", C.c_double * 3), ("la_h", C.c_double), ("la_n", C.c_int32 * 3),
        ("la_slot", C.POINTER(C.c_int32)),
        ("n_vertices", C.c_int32), ("vtx_off", C.POINTER(C.c_int32)), ("vtx_cell", C.POINTER(C.c_int32)),
        ("vtx_w", C.POINTER(C.c_double)), ("leaf_vtx", C.POINTER(C.c_int32)),
        ("lattice_level", C.c_int32),
        ("solid_a", C.POINTER(C.c_double)), ("solid_cm", C.POINTER(C.c_double)),
        ("solid_s", C.POINTER(C.c_double)),
    ]


class StepParamsC(C.Structure):
    _fields_ = [
        ("dt", C.c_double), ("n_forces", C.c_int32), ("force", C.c_int32 * MAX_FORCES),
        ("rho", C.c_double), ("mu", C.c_double), ("g", C.c_double * 3),
        ("cd_const", C.c_double), ("cl_const", C.c_double),
        ("record_cells", C.c_int32), ("record_forces", C.c_int32), ("cm_const", C.c_double),
        ("track_escapes", C.c_int32), ("fuse_deposit", C.c_int32),
    ]


class KernelC(C.Structure):
    """gfsb200_kernel: the closed-form smoothing kernels of GfsSourceParticulate"""
    _fields_ = [("kind", C.c_int32), ("p", C.c_int32), ("a", C.c_double), ("b", C.c_double),
                ("flags", C.c_int32), ("record_norm", C.c_int32)]


KERNEL_FUNC = C.CFUNCTYPE(C.c_double, C.c_double, C.c_double, C.c_double, C.c_void_p)
REFINE_FUNC = C.CFUNCTYPE(C.c_int, C.POINTER(C.c_double), C.c_int, C.c_double, C.c_void_p)

_lib = None


def lib() -> C.CDLL:
    """Loads lib/libgfsb200.so (built by __graft_entry__.build() / make)."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("GFSB200_LIB") or os.path.join(LIB_DIR, "libgfsb200.so")   # GFSB200_LIB: an A/B build
    if not os.path.exists(path):
        raise GfsB200Error(f"{path} not built: run `python -c 'import __graft_entry__ as g; g.build()'`")
    L = C.CDLL(path, mode=C.RTLD_GLOBAL)
    vp, i32, i64, dbl = C.c_void_p, C.c_int, C.c_int64, C.c_double
    sig = {
        "gfsb200_last_error": (C.c_char_p, []),
        "gfsb200_version": (C.c_char_p, []),
        "gfsb200_tree_new": (vp, [i32]),
        "gfsb200_tree_free": (None, [vp]),
        "gfsb200_tree_add_root": (i32, [vp, C.POINTER(dbl), i32, i32]),
        "gfsb200_tree_link_roots": (i32, [vp, i32, i32, i32]),
        "gfsb200_tree_split": (i32, [vp, i32, C.c_uint, C.c_uint]),
        "gfsb200_tree_refine_cell": (i32, [vp, i32]),
        "gfsb200_tree_refine": (i32, [vp, REFINE_FUNC, vp]),
        "gfsb200_tree_refine_uniform": (i32, [vp, i32]),
        "gfsb200_tree_refine_ring": (i32, [vp, i32, i32, dbl, dbl]),
        "gfsb200_tree_corner_sweep": (i32, [vp]),
        "gfsb200_tree_add_boundary": (i32, [vp, i32, i32]),
        "gfsb200_tree_set_periodic": (i32, [vp, i32, i32, i32]),
        "gfsb200_tree_set_solid": (i32, [vp, i32, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
        "gfsb200_tree_finalize": (i32, [vp, vp]),
        "gfsb200_tree_build_stencils": (i32, [vp]),
        "gfsb200_tree_get_view": (i32, [vp, C.POINTER(TreeView)]),
        "gfsb200_tree_corner_interpolator": (i32, [vp, i32, i32, vp, vp]),
        "gfsb200_device_count": (i32, []),
        "gfsb200_ctx_create": (i32, [i32, C.POINTER(vp)]),
        "gfsb200_ctx_destroy": (None, [vp]),
        "gfsb200_ctx_stream": (vp, [vp]),
        "gfsb200_ctx_synchronize": (i32, [vp]),
        "gfsb200_upload_tree": (i32, [vp, vp]),
        "gfsb200_upload_field": (i32, [vp, vp, vp, vp, vp, vp]),
        "gfsb200_set_field_device": (i32, [vp, vp, vp, vp, vp, vp]),
        "gfsb200_upload_field_part": (i32, [vp, i64, i64, vp, vp, vp, vp, vp]),
        "gfsb200_refresh_field": (i32, [vp]),
        "gfsb200_upload_field_prev": (i32, [vp, vp, vp, vp]),
        "gfsb200_download_corner_values": (i32, [vp, i32, i64, vp, vp]),
        "gfsb200_download_vorticity": (i32, [vp, i64, vp, vp]),
        "gfsb200_particles_upload": (i32, [vp, i64] + [vp] * 9),
        "gfsb200_particles_download": (i32, [vp] + [vp] * 13),
        "gfsb200_particles_count": (i64, [vp]),
        "gfsb200_particles_resize": (i32, [vp, i64]),
        "gfsb200_particles_device_ptrs": (i32, [vp, C.POINTER(vp)]),
        "gfsb200_step_params_default": (None, [C.POINTER(StepParamsC)]),
        "gfsb200_step": (i32, [vp, C.POINTER(StepParamsC)]),
        "gfsb200_particle_list_event": (i32, [vp, C.POINTER(StepParamsC), C.POINTER(i64)]),
        "gfsb200_step_host": (i32, [vp, C.POINTER(StepParamsC), i64] + [vp] * 8 + [i64]),
        "gfsb200_particles_cull": (i32, [vp, C.POINTER(i64)]),
        "gfsb200_particle_bc": (i32, [vp, C.POINTER(i64), C.POINTER(i64)]),
        "gfsb200_escaped_count": (i32, [vp, C.POINTER(i64)]),
        "gfsb200_step_counts": (i32, [vp, C.POINTER(i64), C.POINTER(i64)]),
        "gfsb200_escaped_download": (i32, [vp, i64, vp, vp, C.POINTER(i64)]),
        "gfsb200_host_alloc": (vp, [C.c_size_t]),
        "gfsb200_host_free": (None, [vp]),
        "gfsb200_particles_sort": (i32, [vp]),
        "gfsb200_locate": (i32, [vp, i64, vp, vp, vp, vp]),
        "gfsb200_interpolate": (i32, [vp, i64, vp, vp, vp, vp, vp, vp]),
        "gfsb200_output_location": (i32, [vp, i32, vp, i32, i64, vp, vp, vp, vp, vp]),
        "gfsb200_particles_write_gfs": (i32, [vp, C.c_char_p, C.c_char_p, dbl, i32]),
        "gfsb200_checkpoint_save": (i32, [vp, C.c_char_p]),
        "gfsb200_checkpoint_load": (i32, [vp, C.c_char_p]),
        "gfsb200_deposit_volume": (i32, [vp]),
        "gfsb200_deposit_force": (i32, [vp, C.POINTER(StepParamsC)]),
        "gfsb200_deposit_all": (i32, [vp, C.POINTER(StepParamsC)]),
        "gfsb200_deposit_force_smoothed": (i32, [vp, C.POINTER(StepParamsC), dbl, C.POINTER(KernelC)]),
        "gfsb200_download_kernel_norm": (i32, [vp, vp, vp]),
        "gfsb200_kernel_fit": (i32, [KERNEL_FUNC, vp, i32, C.POINTER(KernelC)]),
        "gfsb200_deposit_select": (i32, [vp, i32]),
        "gfsb200_deposit_buffer": (i32, [vp, C.POINTER(vp), C.POINTER(i64)]),
        "gfsb200_download_deposit": (i32, [vp, i32, vp]),
        "gfsb200_comm_unique_id": (i32, [vp]),
        "gfsb200_comm_init_rank": (i32, [vp, vp, i32, i32, C.POINTER(vp)]),
        "gfsb200_comm_init_all": (i32, [i32, C.POINTER(vp), C.POINTER(vp)]),
        "gfsb200_comm_destroy": (None, [vp]),
        "gfsb200_comm_rank": (i32, [vp]),
        "gfsb200_comm_size": (i32, [vp]),
        "gfsb200_comm_peer_access": (i32, [vp]),
        "gfsb200_broadcast_field": (i32, [C.POINTER(vp), i32, i32, vp, vp, vp, vp, vp]),
        "gfsb200_comm_rebalance": (i32, [C.POINTER(vp), i32]),
        "gfsb200_comm_split": (i32, [vp, vp]),
        "gfsb200_comm_splitters": (i32, [vp, C.c_int32, i32, vp]),
        "gfsb200_comm_owner_table": (i32, [vp, vp]),
        "gfsb200_comm_set_shares": (i32, [vp, i32, vp]),
        "gfsb200_comm_owner_slices": (i32, [vp, C.c_int32, C.c_int32, i32, vp, i32, vp]),
        "gfsb200_deposit_allreduce": (i32, [C.POINTER(vp), i32]),
        "gfsb200_deposit_wait": (i32, [vp]),
        "gfsb200_comm_set_exchange": (i32, [vp, i32]),
        "gfsb200_comm_exchange_stats": (i32, [vp, C.POINTER(dbl), C.POINTER(i64), C.POINTER(i64)]),
        "gfsb200_timer_reset": (i32, [vp]),
        "gfsb200_timer_read": (i32, [vp, C.POINTER(dbl), C.POINTER(i64)]),
        "gfsb200_timer_sampling": (i32, [vp, i32]),
        "gfsb200_kernel_launches": (i64, []),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)       # AttributeError if the library lacks a declared symbol
        f.restype = res
        f.argtypes = args
    _lib = L
    return L


EXPORTED_SYMBOLS = None  # filled lazily by exported_symbols()


def _check(rc: int, what: str) -> int:
    if rc < 0:
        raise GfsB200Error(f"{what}: {lib().gfsb200_last_error().decode()} (status {rc})")
    return rc


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f64(a) -> Optional[np.ndarray]:
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def kernel_launches() -> int:
    """this library's own kernels launched so far by the process"""
    return int(lib().gfsb200_kernel_launches())


def kernel_fit(func, dim: int):
    """gfsb200_kernel_fit of a Python callable f(x, y, z); returns a KernelC or raises"""
    k = KernelC()
    cb = KERNEL_FUNC(lambda x, y, z, _d: float(func(x, y, z)))
    _check(lib().gfsb200_kernel_fit(cb, None, dim, C.byref(k)), "kernel_fit")
    return k


class Tree:
    """A flat FTT tree (gfsb200_tree*)."""

    def __init__(self, dim: int = 3, handle: Optional[int] = None):
        self._lib = lib()
        if handle is None:
            handle = self._lib.gfsb200_tree_new(dim)
            if not handle:
                raise GfsB200Error(self._lib.gfsb200_last_error().decode())
        self.handle = C.c_void_p(handle)
        self.dim = dim
        self._cb = None

    def __del__(self):
        try:
            if self.handle:
                self._lib.gfsb200_tree_free(self.handle)
                self.handle = None
        except Exception:
            pass

    # growth ---------------------------------------------------------------
    def add_root(self, pos: Sequence[float], level: int = 0, is_box: bool = True) -> int:
        p = (C.c_double * 3)(*[float(v) for v in pos])
        return _check(self._lib.gfsb200_tree_add_root(self.handle, p, level, int(is_box)), "add_root")

    def link_roots(self, r0: int, d: int, r1: int):
        _check(self._lib.gfsb200_tree_link_roots(self.handle, r0, d, r1), "link_roots")

    def split(self, cell: int, destroyed_mask: int = 0, child_flags: int = 0) -> int:
        return _check(self._lib.gfsb200_tree_split(self.handle, cell, destroyed_mask, child_flags), "split")

    def refine_cell(self, cell: int) -> int:
        return _check(self._lib.gfsb200_tree_refine_cell(self.handle, cell), "refine_cell")

    def refine(self, func):
        """func(pos(x,y,z), level, h) -> bool, semantics of ftt_cell_refine."""
        def tramp(pos, level, h, data):
            return 1 if func((pos[0], pos[1], pos[2]), level, h) else 0
        cb = REFINE_FUNC(tramp)
        _check(self._lib.gfsb200_tree_refine(self.handle, cb, None), "refine")

    def refine_uniform(self, level: int):
        _check(self._lib.gfsb200_tree_refine_uniform(self.handle, level), "refine_uniform")

    def refine_ring(self, minlevel: int, maxlevel: int, R: float = 0.25, factor: float = 1.5):
        _check(self._lib.gfsb200_tree_refine_ring(self.handle, minlevel, maxlevel, R, factor), "refine_ring")

    def corner_sweep(self):
        _check(self._lib.gfsb200_tree_corner_sweep(self.handle), "corner_sweep")

    def add_boundary(self, side: int, box_root: int = 0):
        _check(self._lib.gfsb200_tree_add_boundary(self.handle, box_root, side), "add_boundary")

    def set_periodic(self, side: int, box_root: int = 0, matching_box_root: int = 0):
        _check(self._lib.gfsb200_tree_set_periodic(self.handle, box_root, side, matching_box_root),
               "set_periodic")

    def finalize(self) -> np.ndarray:
        # n_cells is only known through the view after finalize; over-allocate via a first call
        _check(self._lib.gfsb200_tree_finalize(self.handle, None), "finalize")
        return self.view()

    def build_stencils(self):
        _check(self._lib.gfsb200_tree_build_stencils(self.handle), "build_stencils")
        self._view = None

    # views ----------------------------------------------------------------
    def view(self) -> "TreeArrays":
        v = TreeView()
        _check(self._lib.gfsb200_tree_get_view(self.handle, C.byref(v)), "get_view")
        return TreeArrays(self, v)

    def set_solid(self, cell: int, a: float, cm, s=None):
        """mixed (solid-cut) cell of the finalized tree: fluid fraction, centre of mass and
        (optionally) the 2*dim face fractions"""
        v = (C.c_double * 3)(*[float(x) for x in list(cm) + [0.0] * (3 - len(cm))])
        sv = None if s is None else (C.c_double * 6)(*[float(x) for x in list(s) + [1.0] * (6 - len(s))])
        _check(self._lib.gfsb200_tree_set_solid(self.handle, int(cell), float(a), v, sv), "set_solid")

    def corner_interpolator(self, cell: int, k: int):
        cells = (C.c_int32 * 29)()
        w = (C.c_double * 29)()
        n = _check(self._lib.gfsb200_tree_corner_interpolator(self.handle, cell, k, cells, w),
                   "corner_interpolator")
        return list(cells[:n]), list(w[:n])


class TreeArrays:
    """numpy views (no copies) of a finalized tree; valid while the Tree lives."""

    def __init__(self, tree: Tree, v: TreeView):
        self.tree = tree
        n, dim = v.n_cells, v.dim
        as_arr = np.ctypeslib.as_array
        self.dim, self.n_cells, self.n_roots, self.n_box_roots = dim, n, v.n_roots, v.n_box_roots
        self.min_level, self.max_level, self.complete_level = v.min_level, v.max_level, v.complete_level
        self.n_leaves = v.n_leaves
        self.parent = as_arr(v.parent, shape=(n,))
        self.child0 = as_arr(v.child0, shape=(n,))
        self.neighbor = as_arr(v.neighbor, shape=(n, 2 * dim))
        self.level = as_arr(v.level, shape=(n,))
        self.flags = as_arr(v.flags, shape=(n,))
        self.pos = as_arr(v.pos, shape=(n, 3))
        nl = int(self.level.max()) - v.min_level + 1 if n else 0
        self.level_start = as_arr(v.level_start, shape=(nl + 1,))
        self.la_min = np.array(list(v.la_min))
        self.la_h = v.la_h
        self.la_n = np.array(list(v.la_n))
        self.la_slot = as_arr(v.la_slot, shape=(int(np.prod(self.la_n)),))
        self.n_vertices = v.n_vertices
        self.lattice_level = v.lattice_level
        self.solid_a = as_arr(v.solid_a, shape=(n,)) if bool(v.solid_a) else None
        self.solid_cm = as_arr(v.solid_cm, shape=(n, 3)) if bool(v.solid_cm) else None
        self.solid_s = as_arr(v.solid_s, shape=(n, 2 * dim)) if bool(v.solid_s) else None
        if v.n_vertices and bool(v.vtx_off):
            self.vtx_off = as_arr(v.vtx_off, shape=(v.n_vertices + 1,))
            ne = int(self.vtx_off[-1])
            self.vtx_cell = as_arr(v.vtx_cell, shape=(ne,))
            self.vtx_w = as_arr(v.vtx_w, shape=(ne,))
            self.leaf_vtx = as_arr(v.leaf_vtx, shape=(n, 2 ** dim))
        else:
            self.vtx_off = self.vtx_cell = self.vtx_w = self.leaf_vtx = None

    @property
    def box_leaves(self) -> np.ndarray:
        """indices of the non-ghost leaf cells"""
        return np.nonzero((self.flags & (CELL_LEAF | CELL_BOUNDARY)) == CELL_LEAF)[0].astype(np.int32)

    @property
    def h(self) -> np.ndarray:
        return np.ldexp(1.0, -self.level.astype(np.int32))


class StepParams:
    def __init__(self, dt: float, forces: Sequence[int] = (), rho: float = 1.0, mu: float = 0.0,
                 g: Sequence[float] = (0.0, 0.0, 0.0), cd_const: float = float("nan"),
                 cl_const: float = float("nan"), record_cells: bool = False, record_forces: bool = False,
                 cm_const: float = float("nan"), track_escapes: bool = False, fuse_deposit: bool = False):
        self.c = StepParamsC()
        lib().gfsb200_step_params_default(C.byref(self.c))
        self.c.dt = dt
        self.c.n_forces = len(forces)
        for k, f in enumerate(forces):
            self.c.force[k] = int(f)
        self.c.rho, self.c.mu = rho, mu
        for a in range(3):
            self.c.g[a] = float(g[a])
        self.c.cd_const, self.c.cl_const, self.c.cm_const = cd_const, cl_const, cm_const
        self.c.record_cells, self.c.record_forces = int(record_cells), int(record_forces)
        self.c.track_escapes = int(track_escapes)
        self.c.fuse_deposit = int(fuse_deposit)


class Context:
    """A device context (gfsb200_ctx*): one GPU, one stream."""

    def __init__(self, device: int = 0):
        self._lib = lib()
        h = C.c_void_p()
        _check(self._lib.gfsb200_ctx_create(device, C.byref(h)), "ctx_create")
        self.handle = h
        self.device = device
        self.dim = None

    def close(self):
        if self.handle:
            self._lib.gfsb200_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def stream(self) -> int:
        return int(self._lib.gfsb200_ctx_stream(self.handle) or 0)

    def synchronize(self):
        _check(self._lib.gfsb200_ctx_synchronize(self.handle), "synchronize")

    def upload_tree(self, tree: Tree):
        _check(self._lib.gfsb200_upload_tree(self.handle, tree.handle), "upload_tree")
        self.dim = tree.dim
        self.n_cells = tree.view().n_cells

    def upload_field(self, u, v, w=None, alpha=None, mu=None):
        arrs = [_f64(a) for a in (u, v, w, alpha, mu)]
        for a in arrs:
            if a is not None and a.shape != (self.n_cells,):
                raise GfsB200Error(f"field array of shape {a.shape}, expected ({self.n_cells},)")
        _check(self._lib.gfsb200_upload_field(self.handle, *[_ptr(a) for a in arrs]), "upload_field")

    def upload_field_part(self, first: int, n: int, u, v, w=None, alpha=None, mu=None):
        """copies cells [first, first + n) of the given (full-length, contiguous float64) arrays;
        finish with refresh_field()"""
        arrs = [u, v, w, alpha, mu]
        _check(self._lib.gfsb200_upload_field_part(self.handle, first, n, *[_ptr(a) for a in arrs]),
               "upload_field_part")

    def set_field_device(self, u: int, v: int, w: int = 0, alpha: int = 0, mu: int = 0):
        _check(self._lib.gfsb200_set_field_device(self.handle, u, v, w or None, alpha or None, mu or None),
               "set_field_device")

    def upload_field_prev(self, un, vn, wn=None):
        arrs = [_f64(a) for a in (un, vn, wn)]
        _check(self._lib.gfsb200_upload_field_prev(self.handle, *[_ptr(a) for a in arrs]), "upload_field_prev")

    def refresh_field(self):
        _check(self._lib.gfsb200_refresh_field(self.handle), "refresh_field")

    def corner_values(self, comp: int, cells) -> np.ndarray:
        cells = np.ascontiguousarray(cells, dtype=np.int32)
        out = np.empty((len(cells), 2 ** self.dim))
        _check(self._lib.gfsb200_download_corner_values(self.handle, comp, len(cells), _ptr(cells), _ptr(out)),
               "download_corner_values")
        return out

    def vorticity(self, cells) -> np.ndarray:
        cells = np.ascontiguousarray(cells, dtype=np.int32)
        out = np.empty((len(cells), 3))
        _check(self._lib.gfsb200_download_vorticity(self.handle, len(cells), _ptr(cells), _ptr(out)),
               "download_vorticity")
        return out

    # particles --------------------------------------------------------------
    def particles_upload(self, x, y, z, vx, vy, vz, mass, volume, ids=None):
        arrs = [_f64(a) for a in (x, y, z, vx, vy, vz, mass, volume)]
        n = len(arrs[0])
        idarr = None if ids is None else np.ascontiguousarray(ids, dtype=np.uint32)
        _check(self._lib.gfsb200_particles_upload(self.handle, n, *[_ptr(a) for a in arrs], _ptr(idarr)),
               "particles_upload")

    def particles_download(self, forces: bool = False, cells: bool = False, ids: bool = False) -> dict:
        n = self.count
        names = ["x", "y", "z", "vx", "vy", "vz"]
        out = {k: np.empty(n) for k in names}
        out["mass"], out["volume"] = np.empty(n), np.empty(n)
        f = [np.empty(n) for _ in range(3)] if forces else [None] * 3
        idarr = np.empty(n, dtype=np.uint32) if ids else None
        cell = np.empty(n, dtype=np.int32) if cells else None
        _check(self._lib.gfsb200_particles_download(
            self.handle, *[_ptr(out[k]) for k in names], *[_ptr(a) for a in f],
            _ptr(out["mass"]), _ptr(out["volume"]), _ptr(idarr), _ptr(cell)), "particles_download")
        if forces:
            out["fx"], out["fy"], out["fz"] = f
        if ids:
            out["id"] = idarr
        if cells:
            out["cell"] = cell
        return out

    @property
    def count(self) -> int:
        return int(self._lib.gfsb200_particles_count(self.handle))

    def particles_resize(self, n: int):
        _check(self._lib.gfsb200_particles_resize(self.handle, n), "particles_resize")

    def particles_device_ptrs(self):
        p = (C.c_void_p * 8)()
        _check(self._lib.gfsb200_particles_device_ptrs(self.handle, p), "particles_device_ptrs")
        return [int(v or 0) for v in p]

    def step(self, params: StepParams):
        _check(self._lib.gfsb200_step(self.handle, C.byref(params.c)), "step")

    def step_host(self, params: StepParams, x, y, z, vx, vy, vz, mass, volume, chunk: int = 0):
        """in-place step of host-resident float64 arrays (no copies are made here:
        the arrays must already be contiguous float64)"""
        arrs = [x, y, z, vx, vy, vz, mass, volume]
        for a in arrs:
            if a is not None and (a.dtype != np.float64 or not a.flags.c_contiguous):
                raise GfsB200Error("step_host needs contiguous float64 arrays")
        _check(self._lib.gfsb200_step_host(self.handle, C.byref(params.c), len(x),
                                           *[_ptr(a) for a in arrs], chunk), "step_host")

    def particle_list_event(self, params: StepParams) -> int:
        removed = C.c_int64(0)
        _check(self._lib.gfsb200_particle_list_event(self.handle, C.byref(params.c), C.byref(removed)),
               "particle_list_event")
        return removed.value

    def cull(self) -> int:
        removed = C.c_int64(0)
        _check(self._lib.gfsb200_particles_cull(self.handle, C.byref(removed)), "particles_cull")
        return removed.value

    def escaped(self):
        """(list indices, positions before the step [n,3]) of the particles that left the domain
        during the last step issued with track_escapes"""
        n = self.escaped_count()
        idx = np.empty(max(n, 1), dtype=np.int32)
        old = np.empty((max(n, 1), 3))
        m = C.c_int64(0)
        _check(self._lib.gfsb200_escaped_download(self.handle, n, _ptr(idx), _ptr(old), C.byref(m)),
               "escaped_download")
        return idx[:m.value], old[:m.value]

    def step_counts(self):
        """(escaped, outside before the step) of the last step issued with track_escapes"""
        e, o = C.c_int64(0), C.c_int64(0)
        _check(self._lib.gfsb200_step_counts(self.handle, C.byref(e), C.byref(o)), "step_counts")
        return int(e.value), int(o.value)

    def escaped_count(self):
        """particles that left the domain during the last step issued with track_escapes"""
        n = C.c_int64(0)
        _check(self._lib.gfsb200_escaped_count(self.handle, C.byref(n)), "escaped_count")
        return int(n.value)

    def particle_bc(self):
        """(wrapped, dropped) of gfs_particle_bc after a step with track_escapes"""
        w, d = C.c_int64(0), C.c_int64(0)
        _check(self._lib.gfsb200_particle_bc(self.handle, C.byref(w), C.byref(d)), "particle_bc")
        return w.value, d.value

    def sort(self):
        _check(self._lib.gfsb200_particles_sort(self.handle), "particles_sort")

    def locate(self, x, y, z=None) -> np.ndarray:
        x, y, z = _f64(x), _f64(y), _f64(z)
        cell = np.empty(len(x), dtype=np.int32)
        _check(self._lib.gfsb200_locate(self.handle, len(x), _ptr(x), _ptr(y), _ptr(z), _ptr(cell)), "locate")
        return cell

    def interpolate(self, x, y, z=None):
        x, y, z = _f64(x), _f64(y), _f64(z)
        n = len(x)
        u, v = np.empty(n), np.empty(n)
        w = np.empty(n) if self.dim == 3 else None
        _check(self._lib.gfsb200_interpolate(self.handle, n, _ptr(x), _ptr(y), _ptr(z), _ptr(u), _ptr(v), _ptr(w)),
               "interpolate")
        return u, v, w

    def output_location(self, variables, x, y, z=None, interpolate=True):
        """GfsOutputLocation for arbitrary cell variables (each an array of n_cells doubles);
        returns (values [nvar][n], cell [n])"""
        x, y = np.ascontiguousarray(x, dtype=np.float64), np.ascontiguousarray(y, dtype=np.float64)
        z = None if z is None else np.ascontiguousarray(z, dtype=np.float64)
        vs = [np.ascontiguousarray(v, dtype=np.float64) for v in variables]
        n = len(x)
        out = np.empty((len(vs), n))
        cell = np.empty(n, dtype=np.int32)
        vp = (C.c_void_p * max(len(vs), 1))(*[v.ctypes.data for v in vs])
        op = (C.c_void_p * max(len(vs), 1))(*[out[k].ctypes.data for k in range(len(vs))])
        _check(self._lib.gfsb200_output_location(self.handle, len(vs), vp, int(bool(interpolate)), n, _ptr(x),
                                                 _ptr(y), _ptr(z), op, _ptr(cell)), "output_location")
        return out, cell

    def write_gfs(self, path, class_name="GfsParticulate", L=1.0, append=False):
        _check(self._lib.gfsb200_particles_write_gfs(self.handle, str(path).encode(), class_name.encode(), L,
                                                     int(append)), "particles_write_gfs")

    def checkpoint_save(self, path):
        _check(self._lib.gfsb200_checkpoint_save(self.handle, str(path).encode()), "checkpoint_save")

    def checkpoint_load(self, path):
        _check(self._lib.gfsb200_checkpoint_load(self.handle, str(path).encode()), "checkpoint_load")

    def deposit_volume(self):
        _check(self._lib.gfsb200_deposit_volume(self.handle), "deposit_volume")

    def deposit_force(self, params: StepParams):
        _check(self._lib.gfsb200_deposit_force(self.handle, C.byref(params.c)), "deposit_force")

    def deposit_all(self, params: StepParams):
        _check(self._lib.gfsb200_deposit_all(self.handle, C.byref(params.c)), "deposit_all")

    def deposit_force_smoothed(self, params: StepParams, rkernel: float, kind=KERNEL_GAUSSIAN, a=1.0, b=1.0,
                               p=1, flags=0, record_norm=False):
        """GfsSourceParticulate with its smoothing kernel (fills deposit components 1..dim)"""
        k = KernelC(kind, p, a, b, flags, int(record_norm))
        _check(self._lib.gfsb200_deposit_force_smoothed(self.handle, C.byref(params.c), rkernel, C.byref(k)),
               "deposit_force_smoothed")

    def download_kernel_norm(self):
        corr, vol = np.empty(self.count), np.empty(self.count)
        _check(self._lib.gfsb200_download_kernel_norm(self.handle, _ptr(corr), _ptr(vol)), "download_kernel_norm")
        return corr, vol

    def deposit_select(self, which: int):
        _check(self._lib.gfsb200_deposit_select(self.handle, which), "deposit_select")

    def deposit_buffer(self):
        p, n = C.c_void_p(), C.c_int64()
        _check(self._lib.gfsb200_deposit_buffer(self.handle, C.byref(p), C.byref(n)), "deposit_buffer")
        return int(p.value or 0), n.value

    def download_deposit(self, comp: int) -> np.ndarray:
        out = np.empty(self.n_cells)
        _check(self._lib.gfsb200_download_deposit(self.handle, comp, _ptr(out)), "download_deposit")
        return out

    def timer_reset(self):
        _check(self._lib.gfsb200_timer_reset(self.handle), "timer_reset")

    def timer_sampling(self, every: int):
        """CUDA events around every `every`-th step-kernel launch only (1: all, 0: none)"""
        _check(self._lib.gfsb200_timer_sampling(self.handle, int(every)), "timer_sampling")

    def timer_read(self):
        ms, n = C.c_double(), C.c_int64()
        _check(self._lib.gfsb200_timer_read(self.handle, C.byref(ms), C.byref(n)), "timer_read")
        return ms.value, n.value


# ---------------------------------------------------------------------------
# multi-GPU (gfsb200_comm*): one communicator per context

EXCHANGE_AUTO, EXCHANGE_ALLREDUCE = 0, 1
UNIQUE_ID_BYTES = 128


def comm_unique_id() -> bytes:
    buf = C.create_string_buffer(UNIQUE_ID_BYTES)
    _check(lib().gfsb200_comm_unique_id(buf), "comm_unique_id")
    return buf.raw


def comm_splitters(count: np.ndarray, nranks: int) -> np.ndarray:
    """slice boundaries from global per-cell particle counts (host only, no device needed)"""
    cnt = np.ascontiguousarray(count, dtype=np.uint32)
    split = np.zeros(nranks + 1, dtype=np.int32)
    _check(lib().gfsb200_comm_splitters(_ptr(cnt), len(cnt), nranks, _ptr(split)), "comm_splitters")
    return split


def comm_owner_slices(child0: np.ndarray, n_roots: int, dim: int, count: np.ndarray, nranks: int) -> np.ndarray:
    """owner of every leaf (255: not a leaf) from the tree and the global per-cell particle counts: equal
    shares along the depth-first leaf order (host only, no device needed)"""
    c0 = np.ascontiguousarray(child0, dtype=np.int32)
    cnt = np.ascontiguousarray(count, dtype=np.uint32)
    assert len(c0) == len(cnt)
    owner = np.zeros(len(c0), dtype=np.uint8)
    _check(lib().gfsb200_comm_owner_slices(_ptr(c0), len(c0), n_roots, dim, _ptr(cnt), nranks, _ptr(owner)),
           "comm_owner_slices")
    return owner


class Comm:
    """The communicators of THIS process (one per local context): `Comm.init_rank` for one
    process per GPU, `Comm.init_all` for one process driving several GPUs.  Every method is
    a collective over all processes of the job."""

    def __init__(self, handles, ctxs):
        self._lib = lib()
        self.handles = list(handles)
        self.ctxs = list(ctxs)
        self._arr = (C.c_void_p * len(self.handles))(*self.handles)

    @classmethod
    def init_rank(cls, ctx: "Context", unique_id: Optional[bytes], rank: int, nranks: int) -> "Comm":
        h = C.c_void_p()
        idbuf = C.create_string_buffer(unique_id, UNIQUE_ID_BYTES) if unique_id is not None else None
        _check(lib().gfsb200_comm_init_rank(ctx.handle, idbuf, rank, nranks, C.byref(h)), "comm_init_rank")
        return cls([h.value], [ctx])

    @classmethod
    def init_all(cls, ctxs) -> "Comm":
        n = len(ctxs)
        cs = (C.c_void_p * n)(*[c.handle for c in ctxs])
        out = (C.c_void_p * n)()
        _check(lib().gfsb200_comm_init_all(n, cs, out), "comm_init_all")
        return cls([out[k] for k in range(n)], ctxs)

    def close(self):
        for h in self.handles:
            self._lib.gfsb200_comm_destroy(h)
        self.handles = []

    @property
    def n_local(self) -> int:
        return len(self.handles)

    @property
    def size(self) -> int:
        return self._lib.gfsb200_comm_size(self.handles[0])

    @property
    def rank(self) -> int:
        return self._lib.gfsb200_comm_rank(self.handles[0])

    @property
    def peer_access(self) -> bool:
        return bool(self._lib.gfsb200_comm_peer_access(self.handles[0]))

    def set_exchange(self, mode: int):
        for h in self.handles:
            _check(self._lib.gfsb200_comm_set_exchange(h, mode), "comm_set_exchange")

    def broadcast_field(self, root: int, u, v, w=None, alpha=None, mu=None):
        arrs = [_f64(a) for a in (u, v, w, alpha, mu)]
        _check(self._lib.gfsb200_broadcast_field(self._arr, self.n_local, root, *[_ptr(a) for a in arrs]),
               "broadcast_field")

    def rebalance(self):
        _check(self._lib.gfsb200_comm_rebalance(self._arr, self.n_local), "comm_rebalance")

    def split(self) -> np.ndarray:
        s = np.zeros(self.size + 1, dtype=np.int32)
        _check(self._lib.gfsb200_comm_split(self.handles[0], _ptr(s)), "comm_split")
        return s

    def set_shares(self, share=None):
        """target shares of the particles for the next rebalance (one positive number per rank of the
        job; None: equal shares)"""
        s = None if share is None else np.ascontiguousarray(share, dtype=np.float64)
        assert s is None or len(s) == self.size
        _check(self._lib.gfsb200_comm_set_shares(self._arr, self.n_local, _ptr(s)), "comm_set_shares")

    def owner_table(self, n_cells: int) -> np.ndarray:
        """the rank that owns every cell after the last rebalance (255: not a leaf)"""
        o = np.full(n_cells, 255, dtype=np.uint8)
        _check(self._lib.gfsb200_comm_owner_table(self.handles[0], _ptr(o)), "comm_owner_table")
        return o

    def deposit_allreduce(self):
        _check(self._lib.gfsb200_deposit_allreduce(self._arr, self.n_local), "deposit_allreduce")

    def deposit_wait(self):
        for h in self.handles:
            _check(self._lib.gfsb200_deposit_wait(h), "deposit_wait")

    def exchange_stats(self, k: int = 0):
        ms, n, b = C.c_double(), C.c_int64(), C.c_int64()
        _check(self._lib.gfsb200_comm_exchange_stats(self.handles[k], C.byref(ms), C.byref(n), C.byref(b)),
               "comm_exchange_stats")
        return ms.value, n.value, b.value


# ---------------------------------------------------------------------------
# FttCell bridge (lib/libgfsb200_ftt{2D,3D}.so, include/gfsb200_ftt.h)

_bridges = {}


def bridge(dim: int) -> C.CDLL:
    if dim in _bridges:
        return _bridges[dim]
    lib()
    path = os.path.join(LIB_DIR, f"libgfsb200_ftt{dim}D.so")
    if not os.path.exists(path):
        raise GfsB200Error(f"{path} not built")
    B = C.CDLL(path)
    vp, i32 = C.c_void_p, C.c_int
    sig = {
        "gfsb200_ftt_last_error": (C.c_char_p, []),
        "gfsb200_ftt_flatten": (i32, [i32, C.POINTER(vp), C.POINTER(i32), C.POINTER(vp), C.POINTER(vp)]),
        "gfsb200_ftt_map_free": (None, [vp]),
        "gfsb200_ftt_map_size": (C.c_int32, [vp]),
        "gfsb200_ftt_map_cell": (vp, [vp, C.c_int32]),
        "gfsb200_ftt_map_cells": (C.POINTER(vp), [vp]),
        "gfsb200_ftt_gather": (i32, [vp, C.c_size_t, i32, C.c_double, vp]),
        "gfsb200_ftt_scatter": (i32, [vp, C.c_size_t, i32, i32, vp]),
        "gfsb200_ftt_map_cache_data": (i32, [vp, C.c_uint]),
        "gfsb200_ftt_gather_range": (i32, [vp, C.c_size_t, C.c_int32, C.c_int32, i32, vp, vp, vp]),
    }
    for name, (res, args) in sig.items():
        f = getattr(B, name)
        f.restype, f.argtypes = res, args
    _bridges[dim] = B
    return B


class FttMap:
    """flat cell index -> FttCell* of the mirrored Gerris tree"""

    def __init__(self, dim: int, handle):
        self.dim, self.handle = dim, handle
        B = bridge(dim)
        n = B.gfsb200_ftt_map_size(handle)
        p = B.gfsb200_ftt_map_cells(handle)
        self.cells = np.array([p[i] or 0 for i in range(n)], dtype=np.uint64) if n < 4096 else \
            np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint64)), shape=(n,)).copy()

    def __del__(self):
        try:
            if self.handle:
                bridge(self.dim).gfsb200_ftt_map_free(self.handle)
                self.handle = None
        except Exception:
            pass

    def cache_data(self, generation: int = 0):
        """record the data-block address of every cell: later gathers stream through it"""
        _check(bridge(self.dim).gfsb200_ftt_map_cache_data(self.handle, generation), "ftt_map_cache_data")

    def gather(self, offset: int, var: int) -> np.ndarray:
        out = np.empty(len(self.cells))
        _check(bridge(self.dim).gfsb200_ftt_gather(self.handle, offset, var, NODATA, _ptr(out)), "ftt_gather")
        return out

    def scatter(self, offset: int, var: int, values, leaves_only: bool = False):
        values = _f64(values)
        _check(bridge(self.dim).gfsb200_ftt_scatter(self.handle, offset, var, int(leaves_only), _ptr(values)),
               "ftt_scatter")


def flatten_ftt(dim: int, roots: Sequence[int], is_box: Sequence[int]):
    """Mirror a live FttCell tree (root pointers as integers) into a flat Tree."""
    B = bridge(dim)
    n = len(roots)
    arr = (C.c_void_p * n)(*[int(r) for r in roots])
    box = (C.c_int * n)(*[int(b) for b in is_box])
    t, m = C.c_void_p(), C.c_void_p()
    rc = B.gfsb200_ftt_flatten(n, arr, box, C.byref(t), C.byref(m))
    if rc < 0:
        raise GfsB200Error(f"ftt_flatten: {B.gfsb200_ftt_last_error().decode()} (status {rc})")
    return Tree(dim, handle=t.value), FttMap(dim, m)

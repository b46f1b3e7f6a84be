/* device_types.cuh -- plain structs passed by value to the kernels */
#ifndef GFSB200_DEVICE_TYPES_CUH
#define GFSB200_DEVICE_TYPES_CUH

#include <stdint.h>
#include "gfsb200.h"

#define GFSB200_NODATA 1.7976931348623157e308   /* GFS_NODATA = G_MAXDOUBLE, src/utils.h:80 */
#define GFSB200_MAX_DEV_ROOTS 27

/* child0 encoding on the device: >= 0 first child, -1 leaf, -2 destroyed cell
 * (so that the descent sees FTT_CELL_IS_DESTROYED without a second load) */
#define CELL_REGULAR    8u   /* info bit: all 2*dim neighbours are same-level leaves */
#define CHILD_LEAF      (-1)
#define CHILD_DESTROYED (-2)

struct DevTree {
  int dim, n_cells, n_roots, n_box_roots;
  int root_level;
  int top_levels;              /* complete_level - root_level: levels resolved arithmetically */
  int top_start;               /* level_start of the complete level */
  int single_box;              /* locate array has exactly one slot holding box root 0 */
  int slot_is_box;             /* ... and that slot IS the box: same lower corner, same size, corners exact in fp64
				  (host-checked) -- the slot test then implies the root test of ftt_cell_locate */
  int has_destroyed;           /* some cell of a GfsBox tree is destroyed (entirely solid): the hull test does
				  not decide whether a point is inside the domain */
  double root_size;            /* ftt_level_size (root_level) */
  double top_h, top_inv_h;     /* cell size at the complete level and its (exact) inverse */
  double la_inv_h;             /* 1/la_h; la_h = 2^-rootlevel, so x*la_inv_h == x/la_h bitwise */
  double root_pos[GFSB200_MAX_DEV_ROOTS][3];
  /* GfsLocateArray */
  double la_min[3], la_h;
  int la_n[3];
  const int32_t * la_slot;
  /* topology */
  const int32_t * child0;      /* encoded as above */
  const int32_t * neighbor;    /* [n_cells][2*dim] */
  const uint8_t * level;       /* absolute level */
  const uint8_t * info;        /* flags (low 3 bits) | CELL_REGULAR | child id << 4 */
  const int32_t * parent;      /* [n_cells], -1 for roots */
  const double * solid_a;      /* [n_cells] fluid fraction of mixed (solid-cut) cells, 1 elsewhere; NULL: none */
  const double * solid_s;      /* [n_cells][2*dim] face fractions of mixed cells, 1 elsewhere; NULL: none */
  signed char periodic[GFSB200_MAX_DEV_ROOTS][6];   /* box root, side -> matching box root or -1 */
  /* stencils */
  int n_vertices;
  const int32_t * vtx_off;
  const int32_t * vtx_cell;
  const double  * vtx_w;
  const double  * vtx_wuni;    /* [n_vertices] the stencil's weight if all its weights are equal, else NaN */
  const int32_t * leaf_vtx;    /* [n_cells][2^dim] */
  int lattice_n1;              /* > 0: uniform one-box tree, vertex id = (k*n1 + j)*n1 + i with
				  n1 = 2^levels + 1 (no leaf_vtx / child0 loads needed) */
  int lattice_pattern;         /* >= 0 (lattice trees): every interior vertex (i,j,k) has the 2^dim
				  leaves around it as its stencil, all with one weight, in one
				  common order: entry e is leaf (i - 1 + bx, j - 1 + by, k - 1 + bz)
				  with (bx,by,bz) = bits dim*e .. dim*e + dim - 1.  Verified against
				  the CSR tables at upload; -1 = use the tables */
  double lattice_w;            /* that common weight */
  int lattice_bricks;          /* 3D lattice tree that lattice_cell_pass_kernel can tile with 8^3 bricks */
};

struct DevField {
  const double * u[3];         /* cell-centred U,V,W in flat-tree order */
  const double * alpha;        /* optional per-cell 1/rho */
  const double * mu;           /* optional per-cell viscosity */
  /* derived per field update */
  /* 3D table rows are SPLIT (round 2): the first two components as 16-byte (a,b) rows -- eight to a
     128-byte line -- followed, 2*rows doubles further on, by the third as an array of doubles --
     sixteen to a line.  The L1 charges a gather per request AND per 128-byte line it touches;
     against padded 32-byte rows (four to a line, both requests of a row on the same lines) the
     32 sorted particles of a warp touch about a third fewer lines.  gfsb200_row_store3 /
     load_row are the only accessors. */
  double * vtx_val;            /* 3D: [n_vertices] (u,v) rows + [n_vertices] w; 2D: [n_vertices][2] */
  double * vort;               /* 3D: [n_cells] (wx,wy) rows + [n_cells] wz; 2D: [..] (wz).  Indexed by cell, or on
				  lattice trees by (kz*N + ky)*N + kx with N = lattice_n1 - 1 */
  int * nodata_flag;           /* the cell pass writes nodata_epoch here when a vertex stencil touches GFS_NODATA */
  int nodata_epoch;            /* number of this field update: the flag is never cleared, a reader compares it
				  with the epoch (one stream operation less per step than a memset) */
  /* GfsForceInertial / GfsForceAddedMass only */
  const double * uprev[3];     /* Un,Vn,Wn cell values */
  double * vtx_prev;           /* their vertex table, laid out like vtx_val */
  double * acc;                /* per leaf (u.grad)u at the cell centre, laid out like vort (3D: split rows;
				  2D: [..][2]) */
};

struct DevParticles {
  int64_t n;
  double * x, * y, * z, * vx, * vy, * vz, * mass, * volume;
  double * fx, * fy, * fz;     /* optional (record_forces) */
  int32_t * cell;              /* optional (record_cells) */
  uint32_t * id;
};

struct DevStep {
  double dt;
  int n_forces;
  unsigned forces;             /* force kinds in list order, 4 bits each (no indexed array: keeps
				  the by-value struct in the constant bank) */
  int need_velocity;           /* any drag/lift in the list */
  double rho, mu;
  double inv_mu;               /* 1/mu for a constant viscosity (0 when mu == 0) */
  double inv_rho;              /* 1/rho, correctly rounded on the host */
  double g[3];
  double cd_const, cl_const;   /* NaN = built-in law */
  double cm_const;             /* GfsForceAddedMass coefficient, NaN = 0.5 */
  int mutates_mass;            /* a GfsForceAddedMass is in the list: mass is written back */
  /* escape tracking for gfs_particle_bc */
  int track_escapes;
  int esc_cap;
  int * esc_count;
  int32_t * esc_idx;           /* [esc_cap] particle index */
  double * esc_old;            /* [esc_cap][3] position before the step */
  unsigned char * keep;        /* [n] or NULL: cleared for a particle that was outside the domain BEFORE the step */
};

#ifdef __CUDACC__
/* row i of a 3D table of n_rows rows (see DevField) */
__device__ __forceinline__ void gfsb200_row_store3 (double * __restrict__ tab, int64_t n_rows, int64_t i,
						    double a, double b, double c)
{
  reinterpret_cast<double2 *> (tab)[i] = make_double2 (a, b);
  tab[2*n_rows + i] = c;
}
#endif

/* Two-way coupling over several GPUs.  After gfsb200_comm_rebalance rank r holds the particles of
 * the cells [split[r], split[r + 1]) of the flat tree and OWNS that slice of the deposit buffer: a
 * deposit into a cell is an fp64 reduction on the owner's copy -- a plain L2 atomic for the own
 * slice, a remote one over NVLink peer memory for the few particles that have drifted into another
 * rank's cells since the last rebalance.  n = 0: no ownership, every deposit is local. */
#define GFSB200_MAX_RANKS 16
struct DevOwners {
  int n, self;
  int32_t split[GFSB200_MAX_RANKS + 1];
  double * base[GFSB200_MAX_RANKS];   /* [1 + dim][n_cells] deposit buffer of every rank (base[self]: local) */
};

/* where a two-way pass accumulates: component k of cell c lives at owner(c)'s base + k*n_cells + c */
struct DevDeposit {
  double * local;              /* this rank's buffer, [1 + dim][n_cells] */
  int64_t n_cells;
  int32_t own_lo, own_hi;      /* cells [own_lo, own_hi) are this rank's: plain local atomics */
  const uint8_t * owner_of;    /* [n_cells] or NULL: the owning rank of every leaf when the slices are ranges of the
				  depth-first leaf order (adaptive trees) instead of ranges of the cell index */
  int self;                    /* this rank (owner_of) */
  const DevOwners * peers;     /* device memory; NULL: one rank, everything is local */
  int local_gpu_scope;         /* experiment (GFSB200_LOCAL_RED_GPU_SCOPE): reductions into the own slice at GPU
				  scope even though peers reduce into it at system scope */
};

/* compact every third (second) bit of a Morton key back into an integer */
__host__ __device__ inline unsigned gfsb200_compact3 (unsigned v)
{
  v &= 0x09249249;
  v = (v | (v >> 2)) & 0x030c30c3;
  v = (v | (v >> 4)) & 0x0300f00f;
  v = (v | (v >> 8)) & 0x030000ff;
  v = (v | (v >> 16)) & 0x3ff;
  return v;
}

__host__ __device__ inline unsigned gfsb200_compact2 (unsigned v)
{
  v &= 0x55555555;
  v = (v | (v >> 1)) & 0x33333333;
  v = (v | (v >> 2)) & 0x0f0f0f0f;
  v = (v | (v >> 4)) & 0x00ff00ff;
  v = (v | (v >> 8)) & 0xffff;
  return v;
}

/* lattice trees: row-major index of leaf `cell` (child digits: bit0 = +x, bit1 = -y, bit2 = -z) */
__host__ __device__ inline int gfsb200_lattice_index (int dim, int top_start, int n, int cell)
{
  const unsigned key = (unsigned) (cell - top_start);
  if (dim == 3) {
    const int kx = gfsb200_compact3 (key), ky = ~gfsb200_compact3 (key >> 1) & (n - 1),
      kz = ~gfsb200_compact3 (key >> 2) & (n - 1);
    return (kz*n + ky)*n + kx;
  }
  const int kx = gfsb200_compact2 (key), ky = ~gfsb200_compact2 (key >> 1) & (n - 1);
  return ky*n + kx;
}

#endif

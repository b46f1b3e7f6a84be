/* cell_kernels.cu -- the per-field-update cell pass (sm_100a).
 *
 *   vertex_values_kernel   one thread per deduplicated vertex: sum(w_i * v_i)
 *                          over its corner stencil for U,V(,W)
 *                          = gfs_cell_corner_value, src/fluid.c:3081-3101
 *   vorticity_kernel       one thread per cell: the cell-constant vorticity
 *                          vector of vorticity_vector,
 *                          modules/particulatecommon.c:142-164, built from
 *                          gfs_center_gradient, src/fluid.c:434-475
 *
 * The reference recomputes both per particle per force (72 stencil
 * constructions and 6 gradients per particle-step); they depend on the cell
 * only, so here they are evaluated once per cell per field update and the
 * particle kernel gathers the results.
 *
 * This translation unit is compiled with --fmad=false and keeps the
 * reference's operation order, so its results are bit-identical to the
 * reference's non-FMA x86-64 arithmetic.  Both kernels are HBM/L2-bound
 * streaming passes over the cell arrays (roofline: DESIGN.md section 4).
 */
#include <stdlib.h>
#include "device_types.cuh"

extern "C" long long gfsb200_launch_counter;   /* kernels launched by this library (capi.cu) */

namespace {

__device__ __forceinline__ bool is_nodata (double v)
{
  return __double2hiint (v) == 0x7fefffff && __double2loint (v) == (int) 0xffffffff;
}
__device__ __forceinline__ bool is_leaf (const DevTree & T, int c) { return T.child0[c] == CHILD_LEAF; }
__device__ __forceinline__ int child_id (const DevTree & T, int c) { return T.info[c] >> 4; }
__device__ __forceinline__ bool child_positive (int n, int axis)
{
  return axis == 0 ? (n & 1) : !((n >> axis) & 1);
}

/* Morton arithmetic on lattice trees: the key of a leaf interleaves kx, ~ky, ~kz
 * (child digit bit0 = +x, bit1 = -y, bit2 = -z).  Adding or subtracting one along
 * an axis is a carry through that axis' bits only ("dilated" arithmetic). */
template <int DIM> struct Dilated {
  static constexpr unsigned X = DIM == 3 ? 0x09249249u : 0x55555555u;
  __device__ static __forceinline__ unsigned inc (unsigned key, unsigned m)
  { return (((key | ~m) + 1u) & m) | (key & ~m); }
  __device__ static __forceinline__ unsigned dec (unsigned key, unsigned m)
  { return (((key & m) - 1u) & m) | (key & ~m); }
};

__device__ __forceinline__ unsigned spread_bits3 (unsigned v)
{
  v &= 0x3ff;
  v = (v | (v << 16)) & 0x030000ff;
  v = (v | (v << 8))  & 0x0300f00f;
  v = (v | (v << 4))  & 0x030c30c3;
  v = (v | (v << 2))  & 0x09249249;
  return v;
}

__device__ __forceinline__ unsigned spread_bits2 (unsigned v)
{
  v &= 0xffff;
  v = (v | (v << 8)) & 0x00ff00ff;
  v = (v | (v << 4)) & 0x0f0f0f0f;
  v = (v | (v << 2)) & 0x33333333;
  v = (v | (v << 1)) & 0x55555555;
  return v;
}

/* gfs_cell_face, src/fluid.c:42-52: a mixed cell has no neighbour through a closed face */
template <int DIM>
__device__ __forceinline__ int face_neighbor (const DevTree & T, int cell, int d)
{
  if (T.solid_s && !(T.solid_s[(int64_t) cell*2*DIM + d] > 0.))
    return -1;
  return T.neighbor[(int64_t) cell*2*DIM + d];
}

/* average_neighbor_value, src/fluid.c:64-93 */
template <int DIM>
__device__ double average_neighbor_value (const DevTree & T, const double * __restrict__ F,
					  int cell, int nb, int d, double * x)
{
  if (is_leaf (T, nb))
    return F[nb];
  double av = 0., a = 0.;
  const int od = d ^ 1, axis = d >> 1;
  const bool od_pos = !(od & 1);
  const int c0 = T.child0[nb];
  for (int k = 0; k < (1 << DIM); k++) {
    if (child_positive (k, axis) != od_pos)
      continue;
    int c = c0 + k;
    if (T.child0[c] != CHILD_DESTROYED && F[c] != GFSB200_NODATA) {
      /* GFS_IS_MIXED (child) ? solid->s[od] : 1. -- the table holds 1 for cells that are not mixed */
      const double w = T.solid_s ? T.solid_s[(int64_t) c*2*DIM + od] : 1.;
      a += w;
      av += w*F[c];
    }
  }
  if (a > 0.) {
    *x = 3./4.;
    return av/a;
  }
  return F[cell];
}

struct Grad2 { double a, b; };

/* interpolate_2D1 (3D), src/fluid.c:214-245; interpolate_1D1 (2D), :171-191 */
template <int DIM>
__device__ Grad2 interpolate_perp (const DevTree & T, const double * __restrict__ F,
				   int cell, int d1, int d2, double x, double y)
{
  Grad2 p = { 1., 0. };
  if (DIM == 3) {
    int f1 = face_neighbor<DIM> (T, cell, d1);
    if (f1 >= 0) {
      double y1 = 1.;
      double p1 = average_neighbor_value<DIM> (T, F, cell, f1, d1, &y1);
      if (p1 != GFSB200_NODATA) {
	double a1 = y/y1;
	p.b += a1*p1;
	p.a -= a1;
      }
    }
  }
  int f2 = face_neighbor<DIM> (T, cell, d2);
  if (f2 >= 0) {
    double x2 = 1.;
    double p2 = average_neighbor_value<DIM> (T, F, cell, f2, d2, &x2);
    if (p2 != GFSB200_NODATA) {
      double a2 = x/x2;
      p.b += a2*p2;
      p.a -= a2;
    }
  }
  return p;
}

/* gfs_neighbor_value, src/fluid.c:364-396 */
template <int DIM>
__device__ double neighbor_value (const DevTree & T, const double * __restrict__ F,
				  int cell, int nb, int d, double * x)
{
  if (T.level[nb] == T.level[cell])
    return average_neighbor_value<DIM> (T, F, cell, nb, d, x);
  if (F[nb] == GFSB200_NODATA)
    return GFSB200_NODATA;
  /* coarser neighbour: perpendicular[d][child id], src/fluid.c:264-278 --
     the directions in which `cell` sits inside its parent along the other
     axes, in cyclic axis order */
  const int n = child_id (T, cell), axis = d >> 1;
  Grad2 vc;
  if (DIM == 3) {
    int a1 = (axis + 1) % 3, a2 = (axis + 2) % 3;
    int d1 = 2*a1 + (child_positive (n, a1) ? 0 : 1);
    int d2 = 2*a2 + (child_positive (n, a2) ? 0 : 1);
    vc = interpolate_perp<DIM> (T, F, nb, d1, d2, 1./4., 1./4.);
  }
  else {
    int a1 = 1 - axis;
    int dp = 2*a1 + (child_positive (n, a1) ? 0 : 1);
    vc = interpolate_perp<DIM> (T, F, nb, 0, dp, 1./4., 0.);
  }
  *x = 3./2.;
  return vc.a*F[nb] + vc.b;
}

/* gfs_center_gradient, src/fluid.c:434-475 */
template <int DIM>
__device__ double center_gradient (const DevTree & T, const double * __restrict__ F, int cell, int c)
{
  const int d = 2*c;
  const int f1 = face_neighbor<DIM> (T, cell, d ^ 1);
  const int f2 = face_neighbor<DIM> (T, cell, d);
  const double v0 = F[cell];
  if (f1 >= 0) {
    double x1 = 1., v1;
    v1 = neighbor_value<DIM> (T, F, cell, f1, d ^ 1, &x1);
    if (f2 >= 0) {
      double x2 = 1., v2;
      v2 = neighbor_value<DIM> (T, F, cell, f2, d, &x2);
      return (x1*x1*(v2 - v0) + x2*x2*(v0 - v1))/(x1*x2*(x2 + x1));
    }
    return (v0 - v1)/x1;
  }
  if (f2 >= 0) {
    double x2 = 1.;
    return (neighbor_value<DIM> (T, F, cell, f2, d, &x2) - v0)/x2;
  }
  return 0.;
}

/* body of vorticity_kernel for block `bid` of `nblk` (also run by cell_pass_small_kernel) */
template <int DIM>
__device__ __forceinline__ void vorticity_body (const DevTree & T, const DevField & fld, int bid, int nblk)
{
  const int stride = nblk*blockDim.x;
  for (int cell = bid*blockDim.x + threadIdx.x; cell < T.n_cells; cell += stride) {
    const unsigned info = T.info[cell];
    const bool box_leaf = (info & (GFSB200_CELL_LEAF | GFSB200_CELL_BOUNDARY)) == GFSB200_CELL_LEAF;
    double wx = 0., wy = 0., wz = 0.;
    if (box_leaf && (info & CELL_REGULAR)) {
      /* all 2*dim neighbours are leaves of the cell's own level: every
	 gfs_neighbor_value is the neighbour's value at x = 1, and
	 gfs_center_gradient reduces to ((v2 - v0) + (v0 - v1))/2 -- the same
	 operations the general path performs with x1 = x2 = 1 */
      const double size = __longlong_as_double ((long long) (1023 - T.level[cell]) << 52);
      int nb[2*DIM];
      if (T.lattice_n1 > 0) {
	/* lattice trees: a regular cell's neighbours follow from its Morton key (no
	   neighbour-row load in front of the value gathers); +y / +z decrement
	   the key's inverted digits */
	const unsigned key = (unsigned) (cell - T.top_start);
	const unsigned X = Dilated<DIM>::X, Y = X << 1, Z = X << 2;
	nb[0] = T.top_start + (int) Dilated<DIM>::inc (key, X);
	nb[1] = T.top_start + (int) Dilated<DIM>::dec (key, X);
	nb[2] = T.top_start + (int) Dilated<DIM>::dec (key, Y);
	nb[3] = T.top_start + (int) Dilated<DIM>::inc (key, Y);
	if (DIM == 3) {
	  nb[2*DIM - 2] = T.top_start + (int) Dilated<DIM>::dec (key, Z);
	  nb[2*DIM - 1] = T.top_start + (int) Dilated<DIM>::inc (key, Z);
	}
      }
      else {
	const int2 * np = reinterpret_cast<const int2 *> (T.neighbor + (int64_t) cell*(2*DIM));
#pragma unroll
	for (int d = 0; d < DIM; d++) {
	  const int2 t = __ldg (np + d);
	  nb[2*d] = t.x; nb[2*d + 1] = t.y;
	}
      }
      const double * __restrict__ U = fld.u[0], * __restrict__ V = fld.u[1];
#define GRAD(F, c) ((((F)[nb[2*(c)]] - (F)[cell]) + ((F)[cell] - (F)[nb[2*(c) + 1]]))/2.)
      if (DIM == 2)
	wz = (GRAD (V, 0) - GRAD (U, 1))/size;
      else {
	const double * __restrict__ W = fld.u[2];
	wx = (GRAD (W, 1) - GRAD (V, 2))/size;
	wy = (GRAD (U, 2) - GRAD (W, 0))/size;
	wz = (GRAD (V, 0) - GRAD (U, 1))/size;
      }
#undef GRAD
    }
    else if (box_leaf) {
      /* ftt_cell_size: 2^-level, exact */
      const double size = __longlong_as_double ((long long) (1023 - T.level[cell]) << 52);
      if (DIM == 2)
	wz = (center_gradient<DIM> (T, fld.u[1], cell, 0) -
	      center_gradient<DIM> (T, fld.u[0], cell, 1))/size;
      else {
	wx = (center_gradient<DIM> (T, fld.u[2], cell, 1) -
	      center_gradient<DIM> (T, fld.u[1], cell, 2))/size;
	wy = (center_gradient<DIM> (T, fld.u[0], cell, 2) -
	      center_gradient<DIM> (T, fld.u[2], cell, 0))/size;
	wz = (center_gradient<DIM> (T, fld.u[1], cell, 0) -
	      center_gradient<DIM> (T, fld.u[0], cell, 1))/size;
      }
    }
    int64_t slot = cell;
    if (T.lattice_n1 > 0) {
      if (!box_leaf)
	continue;
      slot = gfsb200_lattice_index (DIM, T.top_start, T.lattice_n1 - 1, cell);
    }
    if (DIM == 2)
      fld.vort[slot] = wz;
    else {
      gfsb200_row_store3 (fld.vort, T.n_cells, slot, wx, wy, wz);
    }
  }
}

template <int DIM>
__global__ void __launch_bounds__(256, 8)      /* 32 registers: full occupancy hides the nb -> value chain */
vorticity_kernel (DevTree T, DevField fld)
{
  vorticity_body<DIM> (T, fld, blockIdx.x, gridDim.x);
}

/* The cell-constant part of compute_inertial_force
 * (modules/particulatecommon.c:296-300):
 *   A_c = sum_c2  gfs_center_gradient (cell, c2, U_c) * U_c2 (cell) / size
 * per leaf, stored like the vorticity table. */
template <int DIM>
__global__ void __launch_bounds__(256)
convective_kernel (DevTree T, DevField fld)
{
  const int stride = gridDim.x*blockDim.x;
  for (int cell = blockIdx.x*blockDim.x + threadIdx.x; cell < T.n_cells; cell += stride) {
    const unsigned info = T.info[cell];
    const bool box_leaf = (info & (GFSB200_CELL_LEAF | GFSB200_CELL_BOUNDARY)) == GFSB200_CELL_LEAF;
    int64_t slot = cell;
    if (T.lattice_n1 > 0) {
      if (!box_leaf)
	continue;
      slot = gfsb200_lattice_index (DIM, T.top_start, T.lattice_n1 - 1, cell);
    }
    double a[3] = { 0., 0., 0. };
    if (box_leaf) {
      const double size = __longlong_as_double ((long long) (1023 - T.level[cell]) << 52);
      for (int c = 0; c < DIM; c++)
	for (int c2 = 0; c2 < DIM; c2++)
	  a[c] += center_gradient<DIM> (T, fld.u[c], cell, c2)*fld.u[c2][cell]/size;
    }
    if (DIM == 2)
      reinterpret_cast<double2 *> (fld.acc)[slot] = make_double2 (a[0], a[1]);
    else {
      gfsb200_row_store3 (fld.acc, T.n_cells, slot, a[0], a[1], a[2]);
    }
  }
}

/* sum w_i v_i over the CSR stencil of vertex v, in stencil order */
template <int DIM>
__device__ __forceinline__ void vertex_from_tables (const DevTree & T, const DevField & fld, int v,
						     double & s0, double & s1, double & s2)
{
  /* per VARIABLE, as gfs_cell_corner_value is called per variable: a NODATA value of V in the
     stencil does not touch the corner value of U */
  bool nd0 = false, nd1 = false, nd2 = false;
  s0 = s1 = s2 = 0.;
  const int b = T.vtx_off[v], e = T.vtx_off[v + 1];
  /* most stencils carry one weight repeated (equal-size cells around the
     vertex): it is stored once per vertex and the per-entry array is skipped */
  const double wu = T.vtx_wuni[v];
  const bool uni = wu == wu;
  /* batches of 2^DIM entries: all indices first, then all gathers, then the
     ordered accumulation -- the loads of a batch are in flight together.
     Padding entries repeat a valid cell with weight 0 (s + 0*v == s). */
  constexpr int NB = 1 << DIM;
  for (int i = b; i < e; i += NB) {
    int c[NB];
    double w[NB], a0[NB], a1[NB], a2[NB];
#pragma unroll
    for (int j = 0; j < NB; j++) {
	const bool ok = i + j < e;
	c[j] = T.vtx_cell[ok ? i + j : i];
	w[j] = ok ? (uni ? wu : T.vtx_w[i + j]) : 0.;
    }
#pragma unroll
    for (int j = 0; j < NB; j++) {
	a0[j] = fld.u[0][c[j]];
	a1[j] = fld.u[1][c[j]];
	a2[j] = DIM == 3 ? fld.u[2][c[j]] : 0.;
    }
#pragma unroll
    for (int j = 0; j < NB; j++) {
	/* GFS_NODATA = DBL_MAX: compare the high word on the integer pipe */
	nd0 |= is_nodata (a0[j]); nd1 |= is_nodata (a1[j]); nd2 |= DIM == 3 && is_nodata (a2[j]);
	s0 += w[j]*a0[j];
	s1 += w[j]*a1[j];
	if (DIM == 3) s2 += w[j]*a2[j];
    }
  }
  if (nd0 | nd1 | nd2) {
    if (nd0) s0 = GFSB200_NODATA;
    if (nd1) s1 = GFSB200_NODATA;
    if (nd2) s2 = GFSB200_NODATA;
    *fld.nodata_flag = fld.nodata_epoch;
  }
}

/* the same out of line, for the few hull vertices lattice_cell_pass_kernel meets (plain
 * pointer arguments: taking the address of the by-value kernel parameters would copy them to
 * local memory in every thread) */
__device__ __noinline__ void hull_vertex_3d (const int32_t * __restrict__ vtx_off,
					     const int32_t * __restrict__ vtx_cell,
					     const double * __restrict__ vtx_w,
					     const double * __restrict__ U, const double * __restrict__ V,
					     const double * __restrict__ W, int * nodata_flag, int nodata_epoch,
					     double * __restrict__ out, int64_t n_rows, int v)
{
  const int b = vtx_off[v], e = vtx_off[v + 1];
  double s0 = 0., s1 = 0., s2 = 0.;
  bool nd0 = false, nd1 = false, nd2 = false;
  for (int i = b; i < e; i++) {
    const int c = vtx_cell[i];
    const double w = vtx_w[i], a0 = U[c], a1 = V[c], a2 = W[c];
    nd0 |= is_nodata (a0); nd1 |= is_nodata (a1); nd2 |= is_nodata (a2);
    s0 += w*a0; s1 += w*a1; s2 += w*a2;
  }
  if (nd0 | nd1 | nd2) {
    if (nd0) s0 = GFSB200_NODATA;
    if (nd1) s1 = GFSB200_NODATA;
    if (nd2) s2 = GFSB200_NODATA;
    *nodata_flag = nodata_epoch;
  }
  gfsb200_row_store3 (out, n_rows, v, s0, s1, s2);
}

/* gfs_cell_corner_value, src/fluid.c:3081-3101: val = sum w_i v_i in stencil
 * order.  (The GFS_NODATA early-out returns the *calling* leaf's own value,
 * which a shared vertex cannot represent: a vertex whose stencil touches
 * NODATA is stored as NODATA and resolved by the particle kernel.) */
/* body of vertex_values_kernel for block `bid` of `nblk` */
template <int DIM>
__device__ __forceinline__ void vertex_values_body (const DevTree & T, const DevField & fld, int bid, int nblk)
{
  /* Lattice trees: vertices are numbered row-major, cells in Morton order.  A
     CTA then takes an 8x8x4 (3D) / 16x16 (2D) brick of vertices instead of 256
     consecutive ones, which halves the number of distinct cell sectors it
     gathers (9x9x5 cells per brick instead of two 129-long rows). */
  const int n1 = T.lattice_n1;
  const int bx = DIM == 3 ? 8 : 16, by = DIM == 3 ? 8 : 16, bz = DIM == 3 ? 4 : 1;
  const int tx = n1 > 0 ? (n1 + bx - 1)/bx : 0, ty = n1 > 0 ? (n1 + by - 1)/by : 0,
    tz = DIM == 3 && n1 > 0 ? (n1 + bz - 1)/bz : 1;
  const int64_t n_items = n1 > 0 ? (int64_t) tx*ty*tz*256 : T.n_vertices;
  const int64_t stride = (int64_t) nblk*blockDim.x;
  for (int64_t item = (int64_t) bid*blockDim.x + threadIdx.x; item < n_items; item += stride) {
    int v = (int) item;
    if (n1 > 0) {
      const int brick = (int) (item >> 8), t = (int) (item & 255);
      const int i = (brick % tx)*bx + (t % bx);
      const int j = ((brick/tx) % ty)*by + ((t/bx) % by);
      const int k = DIM == 3 ? (brick/(tx*ty))*bz + t/(bx*by) : 0;
      if (i >= n1 || j >= n1 || (DIM == 3 && k >= n1))
	continue;
      v = (k*n1 + j)*n1 + i;
    }
    double s0 = 0., s1 = 0., s2 = 0.;
    bool nd0 = false, nd1 = false, nd2 = false;
    if (n1 > 0 && T.lattice_pattern >= 0) {
      const int i = v % n1, j = (v/n1) % n1, k = DIM == 3 ? v/(n1*n1) : 1;
      const int nn = n1 - 1;
      if (i >= 1 && i < nn && j >= 1 && j < nn && k >= 1 && (DIM == 2 || k < nn)) {
	/* interior vertex of a lattice tree: the 2^DIM leaves around it, weight and
	   order as verified at upload (DevTree.lattice_pattern) -- no table loads */
	constexpr int NB = 1 << DIM;
	unsigned sx[2], sy[2], sz[2] = { 0u, 0u };
	if (DIM == 3) {
	  sx[0] = spread_bits3 (i - 1); sx[1] = spread_bits3 (i);
	  sy[0] = spread_bits3 (~(j - 1) & (nn - 1)) << 1; sy[1] = spread_bits3 (~j & (nn - 1)) << 1;
	  sz[0] = spread_bits3 (~(k - 1) & (nn - 1)) << 2; sz[1] = spread_bits3 (~k & (nn - 1)) << 2;
	}
	else {
	  sx[0] = spread_bits2 (i - 1); sx[1] = spread_bits2 (i);
	  sy[0] = spread_bits2 (~(j - 1) & (nn - 1)) << 1; sy[1] = spread_bits2 (~j & (nn - 1)) << 1;
	}
	const double w = T.lattice_w;
	double a0[NB], a1[NB], a2[NB];
#pragma unroll
	for (int q = 0; q < NB; q++) {
	  const int bits = (T.lattice_pattern >> (DIM*q)) & (NB - 1);
	  const int c = T.top_start + (int) (sx[bits & 1] | sy[(bits >> 1) & 1] | sz[(bits >> 2) & 1]);
	  a0[q] = fld.u[0][c];
	  a1[q] = fld.u[1][c];
	  a2[q] = DIM == 3 ? fld.u[2][c] : 0.;
	}
#pragma unroll
	for (int q = 0; q < NB; q++) {
	  nd0 |= is_nodata (a0[q]); nd1 |= is_nodata (a1[q]); nd2 |= DIM == 3 && is_nodata (a2[q]);
	  s0 += w*a0[q];
	  s1 += w*a1[q];
	  if (DIM == 3) s2 += w*a2[q];
	}
	if (nd0 | nd1 | nd2) {
	  if (nd0) s0 = GFSB200_NODATA;
	  if (nd1) s1 = GFSB200_NODATA;
	  if (nd2) s2 = GFSB200_NODATA;
	  *fld.nodata_flag = fld.nodata_epoch;
	}
	if (DIM == 2)
	  reinterpret_cast<double2 *> (fld.vtx_val)[v] = make_double2 (s0, s1);
	else {
	  gfsb200_row_store3 (fld.vtx_val, T.n_vertices, v, s0, s1, s2);
	}
	continue;
      }
    }
    vertex_from_tables<DIM> (T, fld, v, s0, s1, s2);
    if (DIM == 2)
      reinterpret_cast<double2 *> (fld.vtx_val)[v] = make_double2 (s0, s1);
    else {
      gfsb200_row_store3 (fld.vtx_val, T.n_vertices, v, s0, s1, s2);
    }
  }
}

template <int DIM>
__global__ void __launch_bounds__(256)
vertex_values_kernel (DevTree T, DevField fld)
{
  vertex_values_body<DIM> (T, fld, blockIdx.x, gridDim.x);
}

/* Small trees (adaptive configs: 1e5 cells): both tables in ONE launch -- the first gv blocks
 * take the vertices, the others the cells -- instead of two kernels on two streams with a
 * fork/join pair of events, whose fixed cost dominates at that size. */
template <int DIM>
__global__ void __launch_bounds__(256)
cell_pass_small_kernel (DevTree T, DevField fld, int gv)
{
  if ((int) blockIdx.x < gv)
    vertex_values_body<DIM> (T, fld, blockIdx.x, gv);
  else
    vorticity_body<DIM> (T, fld, blockIdx.x - gv, gridDim.x - gv);
}

/* ------------------------------------------------------------------ */
/* Lattice trees (uniform, one GfsBox, no GfsBoundary -- BASELINE config C2):
 * the whole cell pass for the INTERIOR vertices and leaves in one kernel.
 *
 * One CTA per 8x8x8 brick of leaves.  In the flat tree's Morton order such a
 * brick is 512 consecutive cells, so the bulk of the 10x10x10 region a brick
 * needs (its leaves plus one layer around them) arrives as contiguous 4 KB
 * runs per component; the 488 halo cells are short runs of the neighbouring
 * bricks.  All loads of a CTA are issued before the first use (12 per
 * thread), the region sits in shared memory (24 KB -> 8 CTAs per SM), and
 * both tables are then computed from it:
 *   vertex (i,j,k), i,j,k in [8b+1, 8b+8]: sum of its 8 leaves in the order and
 *     with the weight verified at upload (DevTree.lattice_pattern/lattice_w)
 *     = gfs_cell_corner_value, src/fluid.c:3081-3101
 *   leaf (kx,ky,kz): vorticity_vector (modules/particulatecommon.c:142-164)
 *     from centred differences, the x1 = x2 = 1 case of gfs_center_gradient
 * in the same operation order as the table-driven kernels above (bit-identical
 * results, tested), with no index-table traffic at all.  Leaves on the hull
 * take one-sided differences from the same region; the vertices on the hull
 * (3 % of them) go through the stencil tables.  HBM traffic per
 * launch: 24 B per cell read (halo re-reads hit L2) + 32 B per vertex + 32 B
 * per leaf written. */
#define SMALL_TREE_CELLS 400000   /* below this the cell pass is one merged launch */
#define BRICK 8
#define REG (BRICK + 2)
#define PLANE (REG*REG + 4)      /* plane stride = 8 mod 16 doubles: the (x, z) lanes of a warp hit
				    distinct bank pairs */
#define REFERENCE_PATTERN 0x21ab3e   /* the order gfs_cell_corner_interpolator visits the 8 leaves in */

__device__ __forceinline__ void cp_async8 (double * smem_dst, const double * gmem_src)
{
  asm volatile ("cp.async.ca.shared.global [%0], [%1], 8;"
		:: "r"((unsigned) __cvta_generic_to_shared (smem_dst)), "l"(gmem_src) : "memory");
}

/* issue the asynchronous copies of one brick's region into buffer `buf` */
__device__ __forceinline__ void stage_brick (const DevTree & T, const DevField & fld, int bx, int by, int bz,
					      double (* buf)[REG*PLANE], unsigned (* key)[REG])
{
  const int nn = T.lattice_n1 - 1;
  /* Morton key of region cell (x,y,z) = key[0][x] | key[1][y] | key[2][z]; 0xffffffff marks a
     coordinate outside the lattice */
  if (threadIdx.x < 3*REG) {
    const int axis = threadIdx.x/REG, l = threadIdx.x % REG;
    const int g = (axis == 0 ? bx : axis == 1 ? by : bz)*BRICK - 1 + l;
    unsigned k = 0xffffffffu;
    if ((unsigned) g < (unsigned) nn)
      k = axis == 0 ? spread_bits3 (g) : spread_bits3 (~g & (nn - 1)) << axis;
    key[axis][l] = k;
  }
  __syncthreads ();
  const double * __restrict__ U = fld.u[0], * __restrict__ V = fld.u[1], * __restrict__ W = fld.u[2];
#pragma unroll
  for (int it = 0; it < (REG*REG*REG + 255)/256; it++) {
    const int r = threadIdx.x + 256*it;
    if (r < REG*REG*REG) {
      const int lx = r % REG, ly = (r/REG) % REG, lz = r/(REG*REG);
      const unsigned kx = key[0][lx], ky = key[1][ly], kz = key[2][lz];
      const int o = lz*PLANE + ly*REG + lx;
      if (kx != 0xffffffffu && ky != 0xffffffffu && kz != 0xffffffffu) {
	const int c = T.top_start + (int) (kx | ky | kz);
	cp_async8 (&buf[0][o], U + c); cp_async8 (&buf[1][o], V + c); cp_async8 (&buf[2][o], W + c);
      }
      else        /* outside the lattice: never read (the hull leaves test their coordinates) */
	buf[0][o] = buf[1][o] = buf[2][o] = 0.;
    }
  }
  asm volatile ("cp.async.commit_group;" ::: "memory");
}

/* BULK staging (round 2).  The 512 leaves of a brick are CONSECUTIVE cells of the flat tree (Morton
 * order), so they need no gather: one bulk copy per component (cp.async.bulk, 4 KB, issued by one
 * thread, landing on an mbarrier) brings them into a staging array, and the CTA moves them from there
 * into the (x,y,z)-indexed region -- 2 x 3 LDS/STS per thread, conflict-free on the load side --
 * while only the 488 halo cells still come as 8-byte LDGSTS gathers.  The gathers were the largest
 * single item on the L1 data pipe, the busiest unit of this kernel (72 %): 94 LDGSTS per brick, 20
 * sectors each.  The flat index of a brick's first leaf is odd (level_start = 1 + 8 + ... is), so the
 * copy starts one cell early to be 16-byte aligned and takes one cell more at the end; the field
 * arrays are allocated with that one cell of slack. */
#define STG_CELLS (BRICK*BRICK*BRICK + 2)

__device__ __forceinline__ void stage_brick_bulk (const DevTree & T, const DevField & fld, int bx, int by, int bz,
						   double (* buf)[REG*PLANE], unsigned (* key)[REG],
						   double (* stg)[STG_CELLS], uint64_t * bar)
{
  const int nn = T.lattice_n1 - 1;
  const unsigned bar_a = (unsigned) __cvta_generic_to_shared (bar);
  if (threadIdx.x < 3*REG) {
    const int axis = threadIdx.x/REG, l = threadIdx.x % REG;
    const int g = (axis == 0 ? bx : axis == 1 ? by : bz)*BRICK - 1 + l;
    unsigned k = 0xffffffffu;
    if ((unsigned) g < (unsigned) nn)
      k = axis == 0 ? spread_bits3 (g) : spread_bits3 (~g & (nn - 1)) << axis;
    key[axis][l] = k;
  }
  if (threadIdx.x == 0) {
    asm volatile ("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar_a) : "memory");
    asm volatile ("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads ();
  /* first leaf of the brick: the key with the three low bits of every coordinate cleared (y and z enter
     the key inverted: the brick's cells are the 512 keys above it) */
  const int first = T.top_start + (int) ((key[0][1] | key[1][1] | key[2][1]) & ~511u);
  const int off = first & 1;
  if (threadIdx.x == 0) {
    const unsigned bytes = (unsigned) ((BRICK*BRICK*BRICK + 2*off)*sizeof (double));
    asm volatile ("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar_a), "r"(3*bytes) : "memory");
#pragma unroll
    for (int f = 0; f < 3; f++)
      asm volatile ("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
		    :: "r"((unsigned) __cvta_generic_to_shared (&stg[f][0])), "l"(fld.u[f] + (first - off)),
		       "r"(bytes), "r"(bar_a) : "memory");
  }
  /* the halo: planes lz = 0 and 9 (100 cells each), then the ring of 36 cells around each of the 8
     planes between them */
  const double * __restrict__ U = fld.u[0], * __restrict__ V = fld.u[1], * __restrict__ W = fld.u[2];
  constexpr int N_HALO = REG*REG*REG - BRICK*BRICK*BRICK;
#pragma unroll
  for (int it = 0; it < (N_HALO + 255)/256; it++) {
    const int hcell = threadIdx.x + 256*it;
    if (hcell < N_HALO) {
      int lx, ly, lz;
      if (hcell < 2*REG*REG) {
	lz = hcell < REG*REG ? 0 : REG - 1;
	const int q = hcell < REG*REG ? hcell : hcell - REG*REG;
	ly = q/REG; lx = q - ly*REG;
      }
      else {
	const int h2 = hcell - 2*REG*REG, ring = 4*REG - 4;
	lz = 1 + h2/ring;
	const int q = h2 - (lz - 1)*ring;
	if (q < REG) { ly = 0; lx = q; }
	else if (q < 2*REG) { ly = REG - 1; lx = q - REG; }
	else { ly = 1 + ((q - 2*REG) >> 1); lx = (q & 1) ? REG - 1 : 0; }
      }
      const unsigned kx = key[0][lx], ky = key[1][ly], kz = key[2][lz];
      const int o = lz*PLANE + ly*REG + lx;
      if (kx != 0xffffffffu && ky != 0xffffffffu && kz != 0xffffffffu) {
	const int c = T.top_start + (int) (kx | ky | kz);
	cp_async8 (&buf[0][o], U + c); cp_async8 (&buf[1][o], V + c); cp_async8 (&buf[2][o], W + c);
      }
      else
	buf[0][o] = buf[1][o] = buf[2][o] = 0.;
    }
  }
  asm volatile ("cp.async.commit_group;" ::: "memory");
  /* the bulk copies have landed (phase 0 of the barrier) */
  asm volatile ("{\n\t.reg .pred p;\n\t"
		"WAIT_%=:\n\t"
		"mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t"
		"@p bra DONE_%=;\n\t"
		"bra WAIT_%=;\n\t"
		"DONE_%=:\n\t}" :: "r"(bar_a) : "memory");
  /* staging (Morton order within the brick: x bits direct, y and z bits inverted) -> region */
#pragma unroll
  for (int half = 0; half < 2; half++) {
    const int m = threadIdx.x + 256*half;
    const int lx = (m & 1) | ((m >> 2) & 2) | ((m >> 4) & 4);
    const int ly = 7 - (((m >> 1) & 1) | ((m >> 3) & 2) | ((m >> 5) & 4));
    const int lz = 7 - (((m >> 2) & 1) | ((m >> 4) & 2) | ((m >> 6) & 4));
    const int o = (lz + 1)*PLANE + (ly + 1)*REG + lx + 1;
#pragma unroll
    for (int f = 0; f < 3; f++)
      buf[f][o] = stg[f][off + m];
  }
}

/* One CTA per brick, 4 CTAs per SM (64 registers: the 24 shared loads of a vertex are in flight
 * together).  A persistent variant with two region buffers (the next brick's copies in flight
 * during the compute phase) measured 1.6x SLOWER; dropped. */
template <int PATTERN, bool BULK = false>
__global__ void __launch_bounds__(256, 4)
lattice_cell_pass_kernel (DevTree T, DevField fld)
{
  __shared__ double sh[3][REG*PLANE];
  __shared__ unsigned key[3][REG];
  __shared__ __align__(16) double stg[BULK ? 3 : 1][BULK ? STG_CELLS : 2];
  __shared__ uint64_t stg_bar;
  /* the step kernel behind this one is launched programmatically (its prologue touches the
     particle stream only): once every brick has started, its CTAs may take the SM slots the last
     wave of bricks frees, and wait there for this grid to finish (griddepcontrol.wait) */
  asm volatile ("griddepcontrol.launch_dependents;" ::: "memory");
  /* one CTA per brick, the grid is (nb, nb, nb): no division to find the brick's coordinates */
  const int n1 = T.lattice_n1, nn = n1 - 1, nb = gridDim.x;
  const int bx = blockIdx.x, by = blockIdx.y, bz = blockIdx.z;

  /* a warp = 8 x by 4 z at one y: with the padded plane stride its 64-bit shared loads are
     conflict-free, and its 32-byte table rows form four 256-byte runs */
  const int tx = threadIdx.x & 7, tzq = (threadIdx.x >> 3) & 3, ty = threadIdx.x >> 5;
  const double wgt = T.lattice_w;
  const int pattern = PATTERN >= 0 ? PATTERN : T.lattice_pattern;

  if (BULK)
    stage_brick_bulk (T, fld, bx, by, bz, sh, key, reinterpret_cast<double (*)[STG_CELLS]> (&stg[0][0]), &stg_bar);
  else
    stage_brick (T, fld, bx, by, bz, sh, key);
  asm volatile ("cp.async.wait_group 0;" ::: "memory");
  __syncthreads ();
  {
    const double (* reg)[REG*PLANE] = sh;
    /* a brick off the hull: every leaf has all six neighbours (no case selection in the gradients) */
    const bool inner = bx > 0 && bx < nb - 1 && by > 0 && by < nb - 1 && bz > 0 && bz < nb - 1;
#pragma unroll
    for (int half = 0; half < 2; half++) {
      const int tz = tzq + 4*half;
      const int r0 = (tz + 1)*PLANE + (ty + 1)*REG + (tx + 1);      /* the leaf (bx*8 + tx, ...) */
      /* The 2 x 2 x 2 region cells r0 + {0,1} + {0,REG} + {0,PLANE} are read ONCE into registers:
	 they are the 8 leaves of vertex (i,j,k) = the low corner of the leaf's +x+y+z neighbour, and
	 also the leaf's own value and its three "+" neighbours for the gradients (the shared-memory
	 data pipe is the busiest unit of this kernel: 103 -> 60 loads per warp). */
      double c[3][8];
#pragma unroll
      for (int f = 0; f < 3; f++)
#pragma unroll
	for (int q = 0; q < 8; q++)
	  c[f][q] = reg[f][r0 + (q & 1) + ((q >> 1) & 1)*REG + ((q >> 2) & 1)*PLANE];
      {
	const int i = bx*BRICK + 1 + tx, j = by*BRICK + 1 + ty, k = bz*BRICK + 1 + tz;
	if (i < nn && j < nn && k < nn) {
	  double s0 = 0., s1 = 0., s2 = 0.;
#pragma unroll
	  for (int q = 0; q < 8; q++) {
	    const int bits = (pattern >> (3*q)) & 7;
	    if (PATTERN >= 0) {                 /* compile-time order: straight from the registers */
	      s0 += wgt*c[0][bits]; s1 += wgt*c[1][bits]; s2 += wgt*c[2][bits];
	    }
	    else {
	      const int r = r0 + (bits & 1) + ((bits >> 1) & 1)*REG + ((bits >> 2) & 1)*PLANE;
	      s0 += wgt*reg[0][r]; s1 += wgt*reg[1][r]; s2 += wgt*reg[2][r];
	    }
	  }
	  /* GFS_NODATA is DBL_MAX: a stencil that touches one sums to >= w*DBL_MAX (or
	     overflows); no velocity gets near that, so one magnitude test screens for the
	     exact check */
	  if (!(fabs (s0) < 1e300 && fabs (s1) < 1e300 && fabs (s2) < 1e300)) {
	    bool bad0 = false, bad1 = false, bad2 = false;     /* per variable (gfs_cell_corner_value) */
#pragma unroll
	    for (int q = 0; q < 8; q++) {
	      bad0 |= is_nodata (c[0][q]); bad1 |= is_nodata (c[1][q]); bad2 |= is_nodata (c[2][q]);
	    }
	    if (bad0 | bad1 | bad2) {
	      if (bad0) s0 = GFSB200_NODATA;
	      if (bad1) s1 = GFSB200_NODATA;
	      if (bad2) s2 = GFSB200_NODATA;
	      *fld.nodata_flag = fld.nodata_epoch;
	    }
	  }
	  gfsb200_row_store3 (fld.vtx_val, T.n_vertices, (k*n1 + j)*n1 + i, s0, s1, s2);
	}
      }
      /* ---- the leaf itself.  gfs_center_gradient (src/fluid.c:434-475) with same-level leaf
	 neighbours: centred where both exist (the x1 = x2 = 1 case), one-sided (v0 - v1)/1 or
	 (v2 - v0)/1 at the hull where a neighbour is NULL, 0 where both are.  The "-" neighbours
	 are read unconditionally (outside the lattice the region holds zeros) and the case is
	 selected afterwards: the same operations on the same operands as the branches. */
      {
	const int kx = bx*BRICK + tx, ky = by*BRICK + ty, kz = bz*BRICK + tz;
	/* size is a power of two: its inverse is an exponent flip and x/size == x*inv_size exactly */
	const double inv_size = __longlong_as_double ((2046LL << 52) - __double_as_longlong (T.top_h));
	auto grad = [&] (int f, int plus, int st, int kc) -> double {
	  const double v0 = c[f][0];
	  const double up = c[f][plus] - v0, dn = v0 - reg[f][r0 - st];
	  const bool has_up = kc < nn - 1, has_dn = kc > 0;
	  return has_up ? (has_dn ? (up + dn)/2. : up) : (has_dn ? dn : 0.);
	};
	auto grad_inner = [&] (int f, int plus, int st) -> double {
	  const double v0 = c[f][0];
	  const double up = c[f][plus] - v0, dn = v0 - reg[f][r0 - st];
	  return (up + dn)/2.;
	};
	/* neighbour in direction 2c is +axis c, 2c + 1 is -axis c (FttDirection); c[f][1|2|4] = +x|+y|+z */
	double wx, wy, wz;
	if (inner) {
	  wx = (grad_inner (2, 2, REG) - grad_inner (1, 4, PLANE))*inv_size;
	  wy = (grad_inner (0, 4, PLANE) - grad_inner (2, 1, 1))*inv_size;
	  wz = (grad_inner (1, 1, 1) - grad_inner (0, 2, REG))*inv_size;
	}
	else {
	  wx = (grad (2, 2, REG, ky) - grad (1, 4, PLANE, kz))*inv_size;
	  wy = (grad (0, 4, PLANE, kz) - grad (2, 1, 1, kx))*inv_size;
	  wz = (grad (1, 1, 1, kx) - grad (0, 2, REG, ky))*inv_size;
	}
	gfsb200_row_store3 (fld.vort, T.n_cells, (kz*nn + ky)*nn + kx, wx, wy, wz);
      }
    }
    /* ---- vertices on the hull (a coordinate equal to 0 or nn): their stencils are the few
       leaves that exist around them, in the tables' order; only the bricks that touch the hull
       have any.  This brick owns i in [8 bx + 1, 8 bx + 8], and i = 0 if bx == 0.  They are
       enumerated face by face (81 candidates each, dense in the warp); a vertex on several
       hull faces belongs to the first. */
    const int bc[3] = { bx, by, bz };
#pragma unroll
    for (int f = 0; f < 6; f++) {
      const int axis = f >> 1, side = f & 1;
      if (bc[axis] != (side ? nb - 1 : 0))
	continue;
      const int idx = threadIdx.x;
      if (idx < (BRICK + 1)*(BRICK + 1)) {
	const int a1 = (axis + 1) % 3, a2 = (axis + 2) % 3;
	const int l1 = idx % (BRICK + 1), l2 = idx/(BRICK + 1);
	if ((l1 == 0 && bc[a1]) || (l2 == 0 && bc[a2]))
	  continue;                            /* owned by the neighbouring brick */
	int c[3];
	c[axis] = side ? nn : 0;
	c[a1] = bc[a1]*BRICK + l1;
	c[a2] = bc[a2]*BRICK + l2;
	bool first = true;                     /* not already on a lower-numbered hull face */
#pragma unroll
	for (int g = 0; g < 6; g++)
	  if (g < f && c[g >> 1] == ((g & 1) ? nn : 0))
	    first = false;
	if (first)
	  hull_vertex_3d (T.vtx_off, T.vtx_cell, T.vtx_w, fld.u[0], fld.u[1], fld.u[2], fld.nodata_flag, fld.nodata_epoch,
			  fld.vtx_val, T.n_vertices, (c[2]*n1 + c[1])*n1 + c[0]);
      }
    }
  }
}

} // namespace

static int cell_grid (int64_t n, int n_sm)
{
  int64_t g = (n + 255)/256;
  if (g > (int64_t) n_sm*16) g = (int64_t) n_sm*16;
  return g < 1 ? 1 : (int) g;
}

static int64_t vertex_items (const DevTree * T)
{
  if (T->lattice_n1 <= 0) return T->n_vertices;
  const int n1 = T->lattice_n1, b = T->dim == 3 ? 8 : 16;
  return (int64_t) ((n1 + b - 1)/b)*((n1 + b - 1)/b)*(T->dim == 3 ? (n1 + 3)/4 : 1)*256;
}

/* vertex table of an arbitrary field triple (used for Un,Vn,Wn): fld->u / fld->vtx_val
 * must already point at the source arrays / destination table */
extern "C" void gfsb200_launch_vertex_values (const DevTree * T, const DevField * fld, int n_sm,
					      cudaStream_t stream)
{
  gfsb200_launch_counter += 1;
  const int g = cell_grid (vertex_items (T), n_sm);
  if (T->dim == 2) vertex_values_kernel<2><<<g, 256, 0, stream>>> (*T, *fld);
  else vertex_values_kernel<3><<<g, 256, 0, stream>>> (*T, *fld);
}

extern "C" void gfsb200_launch_convective (const DevTree * T, const DevField * fld, int n_sm,
					   cudaStream_t stream)
{
  gfsb200_launch_counter += 1;
  const int g = cell_grid (T->n_cells, n_sm);
  if (T->dim == 2) convective_kernel<2><<<g, 256, 0, stream>>> (*T, *fld);
  else convective_kernel<3><<<g, 256, 0, stream>>> (*T, *fld);
}

/* The two kernels are independent and each is latency-bound (dependent
 * index -> value gathers), so they are issued on two streams and overlap:
 * `aux` forks from `stream` at ev_fork and joins back at ev_join. */
extern "C" void gfsb200_launch_cell_pass (const DevTree * T, const DevField * fld, int n_sm,
					  cudaStream_t stream, cudaStream_t aux,
					  cudaEvent_t ev_fork, cudaEvent_t ev_join)
{
  const int threads = 256;
  int64_t n_items = T->n_vertices;
  if (T->lattice_n1 > 0) {
    const int n1 = T->lattice_n1, b = T->dim == 3 ? 8 : 16;
    n_items = (int64_t) ((n1 + b - 1)/b)*((n1 + b - 1)/b)*(T->dim == 3 ? (n1 + 3)/4 : 1)*256;
  }
  int gv = (int) ((n_items + threads - 1)/threads), gc = (T->n_cells + threads - 1)/threads;
  const int cap = n_sm*16;            /* grid-stride */
  if (gv > cap) gv = cap;
  if (gc > cap) gc = cap;
  if (gv < 1) gv = 1;
  if (gc < 1) gc = 1;
  /* 3D lattice trees: the whole pass in one brick-tiled kernel */
  if (T->dim == 3 && T->lattice_bricks) {
    const int nb = (T->lattice_n1 - 1)/BRICK;
    gfsb200_launch_counter += 1;
    static const bool bulk = !(getenv ("GFSB200_CELLPASS_BULK") && atoi (getenv ("GFSB200_CELLPASS_BULK")) == 0);
    if (T->lattice_pattern == REFERENCE_PATTERN) {
      if (bulk) lattice_cell_pass_kernel<REFERENCE_PATTERN, true><<<dim3 (nb, nb, nb), 256, 0, stream>>> (*T, *fld);
      else lattice_cell_pass_kernel<REFERENCE_PATTERN><<<dim3 (nb, nb, nb), 256, 0, stream>>> (*T, *fld);
    }
    else
      lattice_cell_pass_kernel<-1><<<dim3 (nb, nb, nb), 256, 0, stream>>> (*T, *fld);
    return;
  }
  if (T->n_cells <= SMALL_TREE_CELLS) {
    gfsb200_launch_counter += 1;
    if (T->dim == 2) cell_pass_small_kernel<2><<<gv + gc, threads, 0, stream>>> (*T, *fld, gv);
    else cell_pass_small_kernel<3><<<gv + gc, threads, 0, stream>>> (*T, *fld, gv);
    return;
  }
  gfsb200_launch_counter += 2;
  cudaEventRecord (ev_fork, stream);
  cudaStreamWaitEvent (aux, ev_fork, 0);
  if (T->dim == 2) {
    vertex_values_kernel<2><<<gv, threads, 0, stream>>> (*T, *fld);
    vorticity_kernel<2><<<gc, threads, 0, aux>>> (*T, *fld);
  }
  else {
    vertex_values_kernel<3><<<gv, threads, 0, stream>>> (*T, *fld);
    vorticity_kernel<3><<<gc, threads, 0, aux>>> (*T, *fld);
  }
  cudaEventRecord (ev_join, aux);
  cudaStreamWaitEvent (stream, ev_join, 0);
}

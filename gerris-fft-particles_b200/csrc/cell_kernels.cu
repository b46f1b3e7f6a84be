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
#include "device_types.cuh"

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

/* average_neighbor_value, src/fluid.c:64-93 (no solid fractions) */
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
      a += 1.;
      av += 1.*F[c];
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
  const int nd = 2*DIM;
  if (DIM == 3) {
    int f1 = T.neighbor[(int64_t) cell*nd + d1];
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
  int f2 = T.neighbor[(int64_t) cell*nd + d2];
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
  const int nd = 2*DIM, d = 2*c;
  const int f1 = T.neighbor[(int64_t) cell*nd + (d ^ 1)];
  const int f2 = T.neighbor[(int64_t) cell*nd + d];
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

template <int DIM>
__global__ void __launch_bounds__(256, 8)      /* 32 registers: full occupancy hides the nb -> value chain */
vorticity_kernel (DevTree T, DevField fld)
{
  const int stride = gridDim.x*blockDim.x;
  for (int cell = blockIdx.x*blockDim.x + threadIdx.x; cell < T.n_cells; cell += stride) {
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
      {
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
      double2 * o = reinterpret_cast<double2 *> (fld.vort + slot*4);
      o[0] = make_double2 (wx, wy);
      o[1] = make_double2 (wz, 0.);
    }
  }
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
      double2 * o = reinterpret_cast<double2 *> (fld.acc + slot*4);
      o[0] = make_double2 (a[0], a[1]);
      o[1] = make_double2 (a[2], 0.);
    }
  }
}

/* gfs_cell_corner_value, src/fluid.c:3081-3101: val = sum w_i v_i in stencil
 * order.  (The GFS_NODATA early-out returns the *calling* leaf's own value,
 * which a shared vertex cannot represent: a vertex whose stencil touches
 * NODATA is stored as NODATA and resolved by the particle kernel.) */
template <int DIM>
__global__ void __launch_bounds__(256)
vertex_values_kernel (DevTree T, DevField fld)
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
  const int64_t stride = (int64_t) gridDim.x*blockDim.x;
  for (int64_t item = (int64_t) blockIdx.x*blockDim.x + threadIdx.x; item < n_items; item += stride) {
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
    const int b = T.vtx_off[v], e = T.vtx_off[v + 1];
    double s0 = 0., s1 = 0., s2 = 0.;
    bool nodata = false;
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
	nodata |= is_nodata (a0[j]) | is_nodata (a1[j]) | (DIM == 3 && is_nodata (a2[j]));
	s0 += w[j]*a0[j];
	s1 += w[j]*a1[j];
	if (DIM == 3) s2 += w[j]*a2[j];
      }
    }
    if (nodata) {
      s0 = s1 = s2 = GFSB200_NODATA;
      *fld.nodata_flag = 1;
    }
    if (DIM == 2)
      reinterpret_cast<double2 *> (fld.vtx_val)[v] = make_double2 (s0, s1);
    else {
      double2 * o = reinterpret_cast<double2 *> (fld.vtx_val + (int64_t) v*4);
      o[0] = make_double2 (s0, s1);
      o[1] = make_double2 (s2, 0.);
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
  const int g = cell_grid (vertex_items (T), n_sm);
  if (T->dim == 2) vertex_values_kernel<2><<<g, 256, 0, stream>>> (*T, *fld);
  else vertex_values_kernel<3><<<g, 256, 0, stream>>> (*T, *fld);
}

extern "C" void gfsb200_launch_convective (const DevTree * T, const DevField * fld, int n_sm,
					   cudaStream_t stream)
{
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

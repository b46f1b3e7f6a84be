/* particle_kernels.cu -- the per-particle kernels (sm_100a).
 *
 *   step_kernel        fused locate + interpolate + forces + integrate
 *                      = gfs_particulate_event, modules/particulatecommon.c:768-842,
 *                        over every member of the GfsParticleList
 *   locate_kernel      batched gfs_domain_locate (src/domain.c:2623-2638)
 *   interpolate_kernel batched locate + gfs_interpolate (src/fluid.c:2697-2710)
 *   deposit kernels    GfsParticulateField / GfsSourceParticulate (single-cell limit)
 *   gather kernels     SoA permutation for the cell sort and the cull
 *
 * Data flow of step_kernel per particle: 8 coalesced fp64 loads
 * (x,y,z,vx,vy,vz,m,V), point location (arithmetic over the complete top
 * levels of the tree, then a child-index descent through L2-resident child0),
 * 2^dim vertex ids + 2^dim (u,v,w) vertex gathers + 1 vorticity gather (all
 * L1/L2 hits when particles are kept sorted by cell), ~250 fp64 instructions,
 * 6 coalesced fp64 stores.  Algorithmic HBM traffic: 112 B/particle-step in
 * 3D, 80 B in 2D.  No tensor cores: nothing here is a
 * contraction.
 */
#include <cub/cub.cuh>
#include "device_types.cuh"

extern "C" long long gfsb200_launch_counter;   /* kernels launched by this library (capi.cu) */

namespace {

/* ------------------------------------------------------------------ */
/* point location                                                       */

struct Located {
  int cell;            /* flat index, -1 = outside the domain */
  double cx, cy, cz;   /* exact centre of the leaf */
  double half;         /* half its size */
  int kx, ky, kz;      /* column indices at the complete level (lattice trees: of the leaf) */
};

/* spread the low 10 bits of v to every third bit */
__device__ __forceinline__ unsigned spread3 (unsigned v)
{
  v &= 0x3ff;
  v = (v | (v << 16)) & 0x030000ff;
  v = (v | (v << 8))  & 0x0300f00f;
  v = (v | (v << 4))  & 0x030c30c3;
  v = (v | (v << 2))  & 0x09249249;
  return v;
}

/* spread the low 16 bits of v to every second bit */
__device__ __forceinline__ unsigned spread2 (unsigned v)
{
  v &= 0xffff;
  v = (v | (v << 8)) & 0x00ff00ff;
  v = (v | (v << 4)) & 0x0f0f0f0f;
  v = (v | (v << 2)) & 0x33333333;
  v = (v | (v << 1)) & 0x55555555;
  return v;
}

/* Number of level thresholds  lo + k*h (k = 1..n-1)  that are strictly below
 * p, i.e. the column the reference's strict-'>' descent (src/ftt.c:1556-1570)
 * ends in after log2(n) levels.  The truncated quotient is a candidate that can
 * be off by one when p - lo rounds; two exact comparisons against the exactly
 * representable dyadic thresholds repair it, so the result is bit-identical to
 * the descent.  Branch-free.  Requires p >= lo (the root test guarantees it). */
__device__ __forceinline__ int column (double p, double lo, double h, double inv_h, int n)
{
  int k = min (__double2int_rz ((p - lo)*inv_h), n - 1);
  const double tk = fma ((double) k, h, lo);           /* lower threshold of column k, exact */
  const int up = (p > tk + h) & (k < n - 1);
  const int down = (!(p > tk)) & (k > 0);
  return k + up - down;
}

/* SKIP_REPAIR (lattice trees): take the branch around the column repair when no coordinate is
 * within 1e-9 h of a cell face -- 30 instructions fewer per call.  Measured (profiles/README.md,
 * round 2): the fused step + deposit kernel gains 3 % from it, the plain step kernel LOSES 5 %
 * (fewer instructions, longer waits at the first use of the gathers), so only the former asks for it. */
template <int DIM, bool LATTICE = false, bool SKIP_REPAIR = false>
__device__ __forceinline__ Located locate (const DevTree & T, double x, double y, double z)
{
  Located L;
  L.cell = -1;
  L.cx = L.cy = L.cz = 0.; L.half = 0.;
  L.kx = L.ky = L.kz = 0;

  if (LATTICE) {
    /* One GfsBox, one locate-array slot, every leaf at the complete level (checked at upload,
       lattice_n1 > 0, together with la_min == root centre - root_size/2 and la_h == root_size,
       both exact).  The slot test floor ((p - min)/la_h) == 0, i.e. 0 <= (p - min)*la_inv_h < 1,
       then implies the inclusive root test of ftt_cell_locate (p - min >= 0 and fl (p - min) < h
       => min <= p < min + h), so that test is not repeated; and (p - min)*top_inv_h is the same
       product scaled by the exact factor n = 2^levels, so ONE product per axis serves the slot
       test (0 <= s < n) and the column candidate (trunc (s) <= n - 1, no clamp).  The candidate
       is repaired by the two exact threshold comparisons of column (), and the centre comes
       from the repaired column kept in floating point: same values as the general path below,
       bit for bit, without its root_pos loads, second set of int -> double conversions and
       6 + 6 comparisons. */
    const int n = 1 << T.top_levels;
    const double h = T.top_h, inv_h = T.top_inv_h, nd = (double) n;
    const double lox = T.la_min[0], loy = T.la_min[1], loz = DIM == 3 ? T.la_min[2] : 0.;
    const double sx = (x - lox)*inv_h, sy = (y - loy)*inv_h, sz = DIM == 3 ? (z - loz)*inv_h : 0.;
    if (!(sx >= 0. && sx < nd && sy >= 0. && sy < nd && sz >= 0. && sz < nd))
      return L;
    int k[3];
    double c[3], kd[3];
    const double p[3] = { x, y, z }, lo[3] = { lox, loy, loz }, sc[3] = { sx, sy, sz };
    /* s differs from the exact (p - lo)/h by less than n 2^-52 <= 2.3e-13: a fractional part
       inside (1e-9, 1 - 1e-9) proves that trunc (s) IS the column, and the repair is skipped
       (all but a few particles in 10^8, those within 1e-9 h of a cell face) */
    bool near_face = false;
#pragma unroll
    for (int a = 0; a < DIM; a++) {
      k[a] = __double2int_rz (sc[a]);
      kd[a] = (double) k[a];
      const double frac = sc[a] - kd[a];
      near_face |= !(frac > 1e-9 && frac < 1. - 1e-9);
    }
    if (!SKIP_REPAIR || near_face) {
#pragma unroll
      for (int a = 0; a < DIM; a++) {
	const int kc = k[a];
	const double tk = fma (kd[a], h, lo[a]);         /* lower threshold of column kc, exact */
	const bool up = (p[a] > tk + h) & (kc < n - 1);
	const bool down = (!(p[a] > tk)) & (kc > 0);
	if (up) kd[a] += 1.;
	if (down) kd[a] -= 1.;
	k[a] = kc + (up ? 1 : 0) - (down ? 1 : 0);
      }
    }
#pragma unroll
    for (int a = 0; a < DIM; a++)
      c[a] = lo[a] + (kd[a] + 0.5)*h;
    L.kx = k[0]; L.ky = k[1]; L.kz = DIM == 3 ? k[2] : 0;
    L.cx = c[0]; L.cy = c[1]; L.cz = DIM == 3 ? c[2] : 0.;
    L.half = 0.5*h;
    L.cell = 0;             /* the Morton index is only materialised when asked for (cell_index ()) */
    return L;
  }

  /* GfsLocateArray, src/domain.c:43-80: i_c = floor ((p_c - min_c)/h); h is a
     power of two, so the division is a multiplication by its exact inverse.
     NaN coordinates fail every comparison below and end up outside, as in the
     reference (floor (NaN) -> INT_MIN -> index -1). */
  int root;
  if (T.single_box) {
    /* one slot: floor (t) == 0  <=>  0 <= t < 1 */
    const double tx = (x - T.la_min[0])*T.la_inv_h, ty = (y - T.la_min[1])*T.la_inv_h;
    const double tz = DIM == 3 ? (z - T.la_min[2])*T.la_inv_h : 0.;
    if (!(tx >= 0. && tx < 1. && ty >= 0. && ty < 1. && tz >= 0. && tz < 1.))
      return L;
    root = 0;
  }
  else {
    if (!(x == x && y == y && z == z))
      return L;
    int ix = (int) floor ((x - T.la_min[0])*T.la_inv_h);
    int iy = (int) floor ((y - T.la_min[1])*T.la_inv_h);
    if (ix < 0 || ix >= T.la_n[0] || iy < 0 || iy >= T.la_n[1])
      return L;
    int index = ix*T.la_n[1] + iy;
    if (DIM == 3) {
      int iz = (int) floor ((z - T.la_min[2])*T.la_inv_h);
      if (iz < 0 || iz >= T.la_n[2])
	return L;
      index = index*T.la_n[2] + iz;
    }
    root = T.la_slot[index];
    if (root < 0)
      return L;
  }

  /* ftt_cell_locate, src/ftt.c:1535-1574: inclusive root test ... */
  double cx = T.root_pos[root][0], cy = T.root_pos[root][1], cz = DIM == 3 ? T.root_pos[root][2] : 0.;
  double half = 0.5*T.root_size;
  /* (one box whose locate-array slot is the box itself: the slot test above has decided, see the
     lattice path) */
  if (!T.slot_is_box &&
      (x > cx + half || x < cx - half || y > cy + half || y < cy - half ||
       (DIM == 3 && (z > cz + half || z < cz - half))))
    return L;

  int cell = root;
  if (LATTICE || T.top_levels > 0) {
    /* ... the first top_levels levels of the descent, resolved arithmetically
       because every GfsBox tree is complete down to that level */
    const int n = 1 << T.top_levels;
    const double h = T.top_h, inv_h = T.top_inv_h;
    const int kx = column (x, cx - half, h, inv_h, n);
    const int ky = column (y, cy - half, h, inv_h, n);
    const int kz = DIM == 3 ? column (z, cz - half, h, inv_h, n) : 0;
    cx = (cx - half) + (kx + 0.5)*h;
    cy = (cy - half) + (ky + 0.5)*h;
    if (DIM == 3) cz = (cz - half) + (kz + 0.5)*h;
    half = 0.5*h;
    L.kx = kx; L.ky = ky; L.kz = kz;
    if (LATTICE) {
      /* every leaf sits at the complete level: no descent, no child0 load; the
	 tables are addressed by (kx,ky,kz), the Morton cell index is only
	 materialised when asked for (cell_index ()) */
      L.cell = 0;
      L.cx = cx; L.cy = cy; L.cz = cz; L.half = half;
      return L;
    }
    /* child digit per level: bit0 = (x > c), bit1 = !(y > c), bit2 = !(z > c) */
    unsigned key;
    if (DIM == 3)
      key = spread3 (kx) | (spread3 (~ky & (n - 1)) << 1) | (spread3 (~kz & (n - 1)) << 2);
    else
      key = spread2 (kx) | (spread2 (~ky & (n - 1)) << 1);
    cell = T.top_start + (root << (DIM*T.top_levels)) + (int) key;
  }

  /* ... then strict-'>' descent with exactly tracked dyadic centres */
  int c0;
  while ((c0 = __ldg (T.child0 + cell)) >= 0) {
    half *= 0.5;
    const bool px = x > cx, py = y > cy;
    int n = (px ? 1 : 0) | (py ? 0 : 2);
    cx += px ? half : -half;
    cy += py ? half : -half;
    if (DIM == 3) {
      const bool pz = z > cz;
      n |= pz ? 0 : 4;
      cz += pz ? half : -half;
    }
    cell = c0 + n;
  }
  if (c0 == CHILD_DESTROYED)
    return L;
  L.cell = cell;
  L.cx = cx; L.cy = cy; L.cz = cz; L.half = half;
  return L;
}

/* is p inside some GfsBox root (locate array slot + inclusive root test)?  For
 * trees without destroyed box cells this is exactly gfs_domain_locate != NULL. */
template <int DIM>
__device__ __forceinline__ bool in_domain (const DevTree & T, double x, double y, double z)
{
  int root = 0;
  if (T.single_box) {
    const double tx = (x - T.la_min[0])*T.la_inv_h, ty = (y - T.la_min[1])*T.la_inv_h;
    const double tz = DIM == 3 ? (z - T.la_min[2])*T.la_inv_h : 0.;
    if (!(tx >= 0. && tx < 1. && ty >= 0. && ty < 1. && tz >= 0. && tz < 1.))
      return false;
  }
  else {
    if (!(x == x && y == y && z == z))
      return false;
    int ix = (int) floor ((x - T.la_min[0])*T.la_inv_h);
    int iy = (int) floor ((y - T.la_min[1])*T.la_inv_h);
    if (ix < 0 || ix >= T.la_n[0] || iy < 0 || iy >= T.la_n[1])
      return false;
    int index = ix*T.la_n[1] + iy;
    if (DIM == 3) {
      int iz = (int) floor ((z - T.la_min[2])*T.la_inv_h);
      if (iz < 0 || iz >= T.la_n[2])
	return false;
      index = index*T.la_n[2] + iz;
    }
    root = T.la_slot[index];
    if (root < 0)
      return false;
  }
  const double cx = T.root_pos[root][0], cy = T.root_pos[root][1], cz = DIM == 3 ? T.root_pos[root][2] : 0.;
  const double half = 0.5*T.root_size;
  return !(x > cx + half || x < cx - half || y > cy + half || y < cy - half ||
	   (DIM == 3 && (z > cz + half || z < cz - half)));
}

/* has the particle left the domain, i.e. is gfs_domain_locate (src/domain.c:2623-2638) NULL at p?
 * The hull test decides unless the tree has destroyed (entirely solid) box cells: then only the
 * full descent does (a particle that steps into a solid cell has left, modules/particulatecommon.c
 * :3333-3335). */
template <int DIM, bool LATTICE>
__device__ __forceinline__ bool left_domain (const DevTree & T, double x, double y, double z)
{
  if (LATTICE) {
    /* the locate-array slot test decides (see locate ()) */
    const double tx = (x - T.la_min[0])*T.la_inv_h, ty = (y - T.la_min[1])*T.la_inv_h;
    const double tz = DIM == 3 ? (z - T.la_min[2])*T.la_inv_h : 0.;
    return !(tx >= 0. && tx < 1. && ty >= 0. && ty < 1. && tz >= 0. && tz < 1.);
  }
  if (T.has_destroyed)
    return locate<DIM, LATTICE> (T, x, y, z).cell < 0;
  return !in_domain<DIM> (T, x, y, z);
}

/* remember a particle that is about to leave the domain; its position before
 * the step is still in global memory at this point */
template <int DIM>
__device__ __forceinline__ void record_escape (const DevStep & S, const DevParticles & P, int64_t i)
{
  const int k = atomicAdd (S.esc_count, 1);
  if (k < S.esc_cap) {
    S.esc_idx[k] = (int32_t) i;
    S.esc_old[3*k] = P.x[i];
    S.esc_old[3*k + 1] = P.y[i];
    S.esc_old[3*k + 2] = DIM == 3 ? P.z[i] : 0.;
  }
}

/* flat cell index of a located leaf (lattice trees: rebuilt from the columns) */
template <int DIM, bool LATTICE>
__device__ __forceinline__ int cell_index (const DevTree & T, const Located & L)
{
  if (!LATTICE || L.cell < 0)
    return L.cell;
  const int n = 1 << T.top_levels;
  unsigned key;
  if (DIM == 3)
    key = spread3 (L.kx) | (spread3 (~L.ky & (n - 1)) << 1) | (spread3 (~L.kz & (n - 1)) << 2);
  else
    key = spread2 (L.kx) | (spread2 (~L.ky & (n - 1)) << 1);
  return T.top_start + (int) key;
}

/* index of the leaf's entry in the per-leaf tables (vorticity) */
template <int DIM, bool LATTICE>
__device__ __forceinline__ int64_t leaf_slot (const DevTree & T, const Located & L)
{
  if (!LATTICE)
    return L.cell;
  const int n = T.lattice_n1 - 1;         /* n <= 2^10 (3D) / 2^15 (2D): the slot fits 32 bits */
  return DIM == 3 ? (L.kz*n + L.ky)*n + L.kx : L.ky*n + L.kx;
}

/* ------------------------------------------------------------------ */
/* interpolation                                                        */

__device__ __forceinline__ double lerp (double a, double b, double t)
{
  return fma (t, b - a, a);
}

/* Row i of a 3D vertex / vorticity / acceleration table of n_rows rows (split layout, DevField):
 * (a, b) with one 128-bit load from the 16-byte rows, c with one 64-bit load from the array behind
 * them.  (One 256-bit load of a padded 32-byte row -- LDG.E.ENL2.256 -- measured 5 % slower, rounds
 * 1 and 2: profiles/README.md.) */
__device__ __forceinline__ void load_row (const double * __restrict__ tab, int64_t n_rows, int64_t i,
					  double & a, double & b, double & c)
{
  const double2 ab = __ldg (reinterpret_cast<const double2 *> (tab) + i);
  a = ab.x; b = ab.y;
  c = __ldg (tab + 2*n_rows + i);
}

/* Per-leaf NODATA fallback of gfs_cell_corner_value (src/fluid.c:3094-3097) */
__device__ __forceinline__ double resolve (double v, const double * __restrict__ F, int cell)
{
  return __double2hiint (v) == 0x7fefffff && __double2loint (v) == (int) 0xffffffff ? F[cell] : v;
}

/* The trilinear form of gfs_interpolate_from_corners (src/fluid.c:2666-2681) from the eight vertex
 * rows id[0..7] of a 3D leaf, written as nested linear interpolations.
 * Corners: 0(-,-,+) 1(+,-,+) 2(+,+,+) 3(-,+,+) 4(-,-,-) 5(+,-,-) 6(+,+,-) 7(-,+,-), evaluated one
 * z-plane at a time (back: 4 5 7 6, front: 0 1 3 2) so that only 12 corner values are live at once.
 * ND: some vertex of the field carries GFS_NODATA -- resolve every value against the leaf's own. */
template <bool ND>
__device__ __forceinline__ void trilinear_rows (const double * __restrict__ vtx_val, int64_t n_vertices,
						const int (& id)[8], double tx, double ty, double tz,
						const double * __restrict__ U, const double * __restrict__ V,
						const double * __restrict__ W, int cell,
						double & u, double & v, double & w)
{
  double pu[2], pv[2], pw[2];
#pragma unroll
  for (int plane = 0; plane < 2; plane++) {
    const int q[4] = { plane ? 0 : 4, plane ? 1 : 5, plane ? 3 : 7, plane ? 2 : 6 };
    double fu[4], fv[4], fw[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      load_row (vtx_val, n_vertices, id[q[k]], fu[k], fv[k], fw[k]);
    }
    if (ND) {
#pragma unroll
      for (int k = 0; k < 4; k++) {
	fu[k] = resolve (fu[k], U, cell);
	fv[k] = resolve (fv[k], V, cell);
	fw[k] = resolve (fw[k], W, cell);
      }
    }
    pu[plane] = lerp (lerp (fu[0], fu[1], tx), lerp (fu[2], fu[3], tx), ty);
    pv[plane] = lerp (lerp (fv[0], fv[1], tx), lerp (fv[2], fv[3], tx), ty);
    pw[plane] = lerp (lerp (fw[0], fw[1], tx), lerp (fw[2], fw[3], tx), ty);
  }
  u = lerp (pu[0], pu[1], tz);
  v = lerp (pv[0], pv[1], tz);
  w = lerp (pw[0], pw[1], tz);
}

/* gfs_interpolate (src/fluid.c:2697-2710) of U,V,W at p inside leaf L.
 * 3D: trilinear in the 8 corner values (gfs_interpolate_from_corners,
 * :2666-2681, written as nested linear interpolations);
 * 2D: centre value + the two diagonal triangles (:2655-2664). */
template <int DIM, bool LATTICE = false>
__device__ __forceinline__ void interpolate (const DevTree & T, const DevField & fld,
					     const Located & L, double x, double y, double z,
					     double & u, double & v, double & w)
{
  /* half is a power of two: its inverse is an exponent flip, no division */
  const double inv = __longlong_as_double ((2046LL << 52) - __double_as_longlong (L.half));
  /* one uniform load decides whether any vertex needs the NODATA fallback */
  const bool any_nodata = __ldg (fld.nodata_flag) == fld.nodata_epoch;
  if (DIM == 3) {
    int id[8];
    if (LATTICE) {
      /* vertex ids straight from the leaf's column indices (row-major lattice) */
      const int n1 = T.lattice_n1;
      const int b = (L.kz*n1 + L.ky)*n1 + L.kx, up = n1*n1;
      id[4] = b;          id[5] = b + 1;          id[7] = b + n1;      id[6] = b + n1 + 1;
      id[0] = b + up;     id[1] = b + up + 1;     id[3] = b + up + n1; id[2] = b + up + n1 + 1;
    }
    else {
      const int4 * vi = reinterpret_cast<const int4 *> (T.leaf_vtx + (int64_t) L.cell*8);
      const int4 i0 = __ldg (vi), i1 = __ldg (vi + 1);
      id[0] = i0.x; id[1] = i0.y; id[2] = i0.z; id[3] = i0.w;
      id[4] = i1.x; id[5] = i1.y; id[6] = i1.z; id[7] = i1.w;
    }
    /* t in [0,1] along each axis: (1 + (p - o)/(h/2))/2 */
    const double tx = fma (x - L.cx, 0.5*inv, 0.5);
    const double ty = fma (y - L.cy, 0.5*inv, 0.5);
    const double tz = fma (z - L.cz, 0.5*inv, 0.5);
    /* two copies of the form: fields with GFS_NODATA vertices are rare, and sharing one copy kept
       six registers of cell-value addresses alive across the gathers of every particle */
    if (any_nodata)
      trilinear_rows<true> (fld.vtx_val, T.n_vertices, id, tx, ty, tz, fld.u[0], fld.u[1], fld.u[2],
			    cell_index<DIM, LATTICE> (T, L), u, v, w);
    else
      trilinear_rows<false> (fld.vtx_val, T.n_vertices, id, tx, ty, tz, NULL, NULL, NULL, 0, u, v, w);
  }
  else {
    int4 id;
    if (LATTICE) {
      /* corners: 0(-,-) 1(+,-) 2(+,+) 3(-,+) */
      const int n1 = T.lattice_n1, b = L.ky*n1 + L.kx;
      id = make_int4 (b, b + 1, b + n1 + 1, b + n1);
    }
    else
      id = __ldg (reinterpret_cast<const int4 *> (T.leaf_vtx + (int64_t) L.cell*4));
    const int cell = cell_index<DIM, LATTICE> (T, L);
    const double2 * vv = reinterpret_cast<const double2 *> (fld.vtx_val);
    double2 f0 = __ldg (vv + id.x), f1 = __ldg (vv + id.y), f2 = __ldg (vv + id.z), f3 = __ldg (vv + id.w);
    const double c0 = fld.u[0][cell], c1 = fld.u[1][cell];
    if (any_nodata) {
      f0.x = resolve (f0.x, fld.u[0], cell); f0.y = resolve (f0.y, fld.u[1], cell);
      f1.x = resolve (f1.x, fld.u[0], cell); f1.y = resolve (f1.y, fld.u[1], cell);
      f2.x = resolve (f2.x, fld.u[0], cell); f2.y = resolve (f2.y, fld.u[1], cell);
      f3.x = resolve (f3.x, fld.u[0], cell); f3.y = resolve (f3.y, fld.u[1], cell);
    }
    const double px = (x - L.cx)*inv, py = (y - L.cy)*inv;
    const double a = (px + py)/2., b = (py - px)/2.;
    double uu = c0, vv2 = c1;
    if (a > 0.) { uu += a*(f2.x - c0); vv2 += a*(f2.y - c1); }
    else        { uu -= a*(f0.x - c0); vv2 -= a*(f0.y - c1); }
    if (b > 0.) { uu += b*(f3.x - c0); vv2 += b*(f3.y - c1); }
    else        { uu -= b*(f1.x - c0); vv2 -= b*(f1.y - c1); }
    u = uu; v = vv2; w = 0.;
  }
}

/* ------------------------------------------------------------------ */
/* forces: modules/particulatecommon.c:423-490 (lift), 519-588 (drag),
 * 617-655 (buoyancy), accumulated as in compute_forces (:737-751)      */

/* The drag law of compute_drag_force,
 *   Re = |u_r| d rho/mu,  cd = 16 (1 + 0.15 Re^.5)/Re  (Re < 50)  or  48 (1 - 2.21/Re^.5)/Re,
 *   f  = 3/(4 d) cd |u_r| u_r rho        (per unit particle volume),
 * with cd/Re substituted: |u_r| and rho cancel, leaving
 *   f = 12 mu (1 + 0.15 Re^.5)/d^2 u_r   or   36 mu (1 - 2.21/Re^.5)/d^2 u_r.
 * Same value up to a few ulp, one square root and no division in the common
 * branch.  d = 2 (3V/4pi)^(1/3) = K V^(1/3): 1/d^2 comes from one rcbrt. */
#define DIA_K 1.2407009817988000333       /* 2 (3/(4 pi))^(1/3) */
#define INV_DIA_K2 0.64962951495345899316  /* 1/DIA_K^2 */

/* PROG: the force list known at compile time (kinds in list order, 4 bits
 * each), or 0 to read it from the launch parameters. */
__host__ __device__ constexpr int prog_len (unsigned prog)
{
  return prog == 0 ? 0 : 1 + prog_len (prog >> 4);
}

/* LATE: where the particle's velocity, mass and volume come from.  NoLate: they are the
 * arguments.  The warp-pipelined step kernel passes LateShared instead: the values stay in
 * the TMA-staged tile until the velocity gathers have been consumed, so that they occupy no
 * registers while those gathers are in flight. */
struct NoLate {
  __device__ __forceinline__ void fetch (double, double &, double &, double &, double &, double &) const {}
};

template <int DIM, bool ONFLUID, bool LATTICE = false, unsigned PROG = 0, typename LATE = NoLate>
__device__ __forceinline__ void total_force (const DevTree & T, const DevField & fld,
					     const DevStep & S, const Located & L,
					     double x, double y, double z,
					     double vx, double vy, double vz,
					     double & mass, double volume,
					     double & Fx, double & Fy, double & Fz, double & rho_out,
					     const LATE late = LATE (),
					     double * r3_memo = NULL, bool r3_known = false)
{
  const unsigned forces = PROG ? PROG : S.forces;
  const int n_forces = PROG ? prog_len (PROG) : S.n_forces;
  bool any_lift = false, need_velocity = PROG ? false : S.need_velocity != 0;
  if (PROG) {
#pragma unroll
    for (int k = 0; k < prog_len (PROG); k++) {
      any_lift |= ((PROG >> (4*k)) & 15) == GFSB200_FORCE_LIFT;
      need_velocity |= ((PROG >> (4*k)) & 15) != GFSB200_FORCE_BUOY;
    }
  }
  /* issue the per-leaf vorticity gather before the vertex gathers so that the
     two latencies overlap */
  double wx = 0., wy = 0., wz = 0.;
  if (PROG && any_lift) {
    const int64_t slot = leaf_slot<DIM, LATTICE> (T, L);
    if (DIM == 3) {
      load_row (fld.vort, T.n_cells, slot, wx, wy, wz);
    }
    else
      wz = __ldg (fld.vort + slot);
  }
  Fx = Fy = Fz = 0.;
  double rho = S.rho;
  if (fld.alpha)
    rho = 1./fld.alpha[cell_index<DIM, LATTICE> (T, L)];
  rho_out = rho;
  double rx = 0., ry = 0., rz = 0.;
  double u = 0., v = 0., w = 0.;
  if (need_velocity)
    interpolate<DIM, LATTICE> (T, fld, L, x, y, z, u, v, w);
  late.fetch (u, vx, vy, vz, mass, volume);
  if (need_velocity) {
    rx = u - vx; ry = v - vy; rz = DIM == 3 ? w - vz : 0.;
  }
#pragma unroll
  for (int k = 0; k < (PROG ? prog_len (PROG) : GFSB200_MAX_FORCES); k++) {
    if (!PROG && k >= n_forces)
      break;
    const int kind = (forces >> (4*k)) & 15;
    /* f*V, the particle force contribution of compute_forces (:744) */
    double fx = 0., fy = 0., fz = 0.;
    if (kind == GFSB200_FORCE_DRAG) {
      double mu = S.mu, inv_mu = S.inv_mu;
      if (fld.mu) {
	mu = fld.mu[cell_index<DIM, LATTICE> (T, L)];
	inv_mu = mu != 0. ? 1./mu : 0.;
      }
      if (mu != 0.) {
	/* V^(-1/3): the fused step + deposit kernel evaluates the forces twice for the same
	   particle and hands the value over (r3_memo) */
	const double r3 = r3_known ? *r3_memo : rcbrt (volume);
	if (r3_memo && !r3_known) *r3_memo = r3;
	const double dia = DIA_K*volume*r3*r3;             /* K V^(1/3) */
	const double inv_d2 = r3*r3*INV_DIA_K2;
	const double nrm = sqrt (DIM == 3 ? rx*rx + ry*ry + rz*rz : rx*rx + ry*ry);
	double k3;
	if (S.cd_const == S.cd_const)
	  k3 = 0.75*S.cd_const*nrm*rho/dia;
	else {
	  const double Re = nrm*dia*rho*inv_mu;
	  const double s = sqrt (Re);
	  if (Re < 1e-8)
	    k3 = 0.;
	  else if (Re < 50.0)
	    k3 = 12.*mu*inv_d2*(1. + 0.15*s);
	  else
	    k3 = 36.*mu*inv_d2*(1. - 2.21/s);
	}
	k3 *= volume;
	fx = k3*rx; fy = k3*ry;
	if (DIM == 3) fz = k3*rz;
      }
    }
    else if (kind == GFSB200_FORCE_LIFT) {
      if (!(PROG && any_lift)) {
	const int64_t slot = leaf_slot<DIM, LATTICE> (T, L);
	if (DIM == 3) {
	  load_row (fld.vort, T.n_cells, slot, wx, wy, wz);
	}
	else
	  wz = __ldg (fld.vort + slot);
      }
      const double cl = S.cl_const == S.cl_const ? S.cl_const : 0.5;
      const double q = rho*cl*volume;
      if (DIM == 3) {
	fx = q*(ry*wz - rz*wy);
	fy = q*(rz*wx - rx*wz);
	fz = q*(rx*wy - ry*wx);
      }
      else {
	fx = q*ry*wz;
	fy = -q*rx*wz;
      }
    }
    else if (!PROG && (kind == GFSB200_FORCE_INERTIAL || kind == GFSB200_FORCE_ADDEDMASS)) {
      /* compute_inertial_force, :255-303: rho ((u - u_old)/dt + (u.grad)u|cell); the
	 convective part is cell-constant and comes from the acc table */
      if (S.dt > 0.) {
	DevField prev = fld;
	prev.vtx_val = fld.vtx_prev;
	prev.u[0] = fld.uprev[0]; prev.u[1] = fld.uprev[1]; prev.u[2] = fld.uprev[2];
	double un, vn, wn;
	interpolate<DIM, LATTICE> (T, prev, L, x, y, z, un, vn, wn);
	const int64_t slot = leaf_slot<DIM, LATTICE> (T, L);
	double ax, ay, az = 0.;
	if (DIM == 3) {
	  load_row (fld.acc, T.n_cells, slot, ax, ay, az);
	}
	else {
	  const double2 a = __ldg (reinterpret_cast<const double2 *> (fld.acc) + slot);
	  ax = a.x; ay = a.y;
	}
	fx = rho*(u - un)/S.dt + rho*ax;
	fy = rho*(v - vn)/S.dt + rho*ay;
	if (DIM == 3) fz = rho*(w - wn)/S.dt + rho*az;
      }
      if (kind == GFSB200_FORCE_ADDEDMASS) {
	const double cm = S.cm_const == S.cm_const ? S.cm_const : 0.5;
	fx *= cm; fy *= cm; fz *= cm;
	mass += rho*volume*cm;            /* the reference's cumulative update, :391 */
      }
      fx *= volume; fy *= volume; fz *= volume;
    }
    else if (kind == GFSB200_FORCE_BUOY) {
      if (!ONFLUID) {                   /* compute_forces_onfluid skips GfsForceBuoy, :753-765 */
	/* (m/V - rho) g V = (m - rho V) g */
	const double dm = mass - rho*volume;
	fx = dm*S.g[0]; fy = dm*S.g[1];
	if (DIM == 3) fz = dm*S.g[2];
      }
    }
    Fx += fx; Fy += fy;
    if (DIM == 3) Fz += fz;
  }
}

/* ------------------------------------------------------------------ */
/* deposition: one reduction per run of equal cells inside a warp       */

/* the buffer that owns `cell': this rank's for its own slice [own_lo, own_hi) (and always on one
 * rank), else the owner's copy in peer memory (NVLink) -- rare: only particles that drifted out of
 * the rank's cells since the last gfsb200_comm_rebalance */
__device__ __forceinline__ double * owner_base (const DevDeposit & D, int cell)
{
  if (D.peers == NULL)
    return D.local;
  if (D.owner_of) {
    /* adaptive trees: slices of the depth-first leaf order, looked up per cell (one byte; only the
       head of a run of equal cells gets here) */
    const int r = D.owner_of[cell];
    return r == D.self || r >= D.peers->n ? D.local : D.peers->base[r];     /* (255: not a leaf -- never reached) */
  }
  if (cell >= D.own_lo && cell < D.own_hi)
    return D.local;
  int r = 0;
  const int n = D.peers->n;
  for (int k = 1; k < n; k++)
    r += cell >= D.peers->split[k];
  return D.peers->base[r];
}

/* fp64 reduction on global memory without a return value (RED.E.ADD.F64): explicit state space, so
 * that no generic-address dispatch is generated, and explicit scope */
__device__ __forceinline__ void red_add (double * p, double v, bool sys)
{
  if (sys)
    asm volatile ("red.relaxed.sys.global.add.f64 [%0], %1;" :: "l"(p), "d"(v) : "memory");
  else
    asm volatile ("red.relaxed.gpu.global.add.f64 [%0], %1;" :: "l"(p), "d"(v) : "memory");
}

/* Adds (av, ax, ay, az) to components (0, 1, 2, 3) of dst[cell] with ONE reduction per component
 * and per run of consecutive lanes that share the same cell (particles are kept sorted by cell, so
 * runs are long): the run structure (heads by shuffle + ballot) is found once, each value is summed
 * along its run by a shuffle tree, and the run's head issues the fp64 reductions (RED.ADD.F64: no
 * return value, so a remote owner costs no round trip).  Lanes with cell < 0 contribute nothing.
 * Must be called by all 32 lanes. */
template <int DIM, bool VOL, bool FORCE>
__device__ __forceinline__ void run_deposit (const DevDeposit & D, int cell, double av, double ax,
					     double ay, double az)
{
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int prev = __shfl_up_sync (full, cell, 1);
  const bool head = lane == 0 || prev != cell;
  const unsigned heads = __ballot_sync (full, head);
  /* my run spans [my_head, run_end) */
  const unsigned above = lane == 31 ? 0u : heads & (0xffffffffu << (lane + 1));
  const int run_end = above ? __ffs (above) - 1 : 32;
  /* as many rounds as the longest run of the warp needs (C2: ~5 particles per leaf -> 3 or 4
     rounds instead of 5; every round is 8 SHFL on the L1 data pipe, the busiest unit of this kernel) */
  const int longest = __reduce_max_sync (full, head ? run_end - lane : 0);
#pragma unroll 1
  for (int o = 1; o < longest; o <<= 1) {
    const bool in = lane + o < run_end;
    if (VOL) { const double t = __shfl_down_sync (full, av, o); if (in) av += t; }
    if (FORCE) {
      const double tx = __shfl_down_sync (full, ax, o), ty = __shfl_down_sync (full, ay, o);
      if (in) { ax += tx; ay += ty; }
      if (DIM == 3) { const double tz = __shfl_down_sync (full, az, o); if (in) az += tz; }
    }
  }
  if (head && cell >= 0) {
    double * base = owner_base (D, cell);
    double * dst = base + cell;
    /* with several ranks every slice also receives reductions from its peers' SMs: system scope */
    const bool sys = D.peers != NULL && !(D.local_gpu_scope && base == D.local);
    if (VOL) red_add (dst, av, sys);
    if (FORCE) {
      red_add (dst + D.n_cells, ax, sys);
      red_add (dst + 2*D.n_cells, ay, sys);
      if (DIM == 3) red_add (dst + 3*D.n_cells, az, sys);
    }
  }
}

/* what one particle adds to its cell in a two-way pass:
 *   VOL    GfsParticulateField with voidfraction_from_particles,
 *          modules/particulatecommon.c:1929-1957:  v[cell] += V_p / V_cell
 *   FORCE  GfsSourceParticulate in the single-cell limit, :2158-2228: forces
 *          recomputed without GfsForceBuoy (compute_forces_onfluid :753-765),
 *          then u_c[cell] -= F_c / rho / V_cell
 * Returns the flat cell index (-1: outside, nothing to add). */
template <int DIM, bool LATTICE, unsigned PROG, bool VOL, bool FORCE, bool SKIP_REPAIR = false>
__device__ __forceinline__ int deposit_terms (const DevTree & T, const DevField & fld, const DevStep & S,
					      double x, double y, double z, double vx, double vy, double vz,
					      double & mass, double volume,
					      double & av, double & ax, double & ay, double & az,
					      double * r3_memo = NULL)
{
  av = ax = ay = az = 0.;
  const Located L = locate<DIM, LATTICE, SKIP_REPAIR> (T, x, y, z);
  const int cell = cell_index<DIM, LATTICE> (T, L);
  if (cell < 0)
    return -1;
  /* 1/ftt_cell_volume: the cell size is a power of two, its inverse an exponent flip */
  const double inv_h = __longlong_as_double ((2045LL << 52) - __double_as_longlong (L.half));   /* 1/(2 half) */
  const double inv_cellvol = DIM == 3 ? inv_h*inv_h*inv_h : inv_h*inv_h;
  if (VOL)
    av = volume*inv_cellvol;
  if (FORCE) {
    double Fx, Fy, Fz, rho;
    total_force<DIM, true, LATTICE, PROG> (T, fld, S, L, x, y, z, vx, vy, vz, mass, volume, Fx, Fy, Fz, rho,
					   NoLate (), r3_memo, r3_memo != NULL);
    /* the single-cell limit of diffuse_force (:2158-2175): gfs_cell_volume, i.e. times the
       fluid fraction in a mixed cell (the void fraction above uses ftt_cell_volume, :1931) */
    /* (inv_cellvol is a power of two: times the correctly rounded 1/rho is the quotient, bit for bit) */
    const double k = T.solid_a || fld.alpha ? -(T.solid_a ? inv_cellvol/T.solid_a[cell] : inv_cellvol)/rho
      : -inv_cellvol*S.inv_rho;
    ax = Fx*k; ay = Fy*k; az = Fz*k;
  }
  return cell;
}

/* ------------------------------------------------------------------ */

template <int DIM, bool REC, bool LATTICE, unsigned PROG, int MINB>
__global__ void __launch_bounds__(256, MINB)
step_kernel (DevTree T, DevField fld, DevParticles P, DevStep S)
{
  const int64_t i = (int64_t) blockIdx.x*blockDim.x + threadIdx.x;
  if (i >= P.n)
    return;
  /* streaming (evict-first) loads: the particle arrays are touched once per
     step and must not displace the vertex / vorticity tables from L1/L2 */
  double x = __ldcs (P.x + i), y = __ldcs (P.y + i), z = DIM == 3 ? __ldcs (P.z + i) : 0.;
  double vx = __ldcs (P.vx + i), vy = __ldcs (P.vy + i), vz = DIM == 3 ? __ldcs (P.vz + i) : 0.;
  double mass = __ldcs (P.mass + i);
  const double volume = __ldcs (P.volume + i);

  const Located L = locate<DIM, LATTICE> (T, x, y, z);
  if (REC)
    P.cell[i] = cell_index<DIM, LATTICE> (T, L);
  if (L.cell < 0) {   /* outside: gfs_particle_list_event removes it first (:987) */
    if (S.track_escapes) {
      atomicAdd (S.esc_count + 3, 1);         /* lets the list event skip its cull pass when 0 */
      if (S.keep) S.keep[i] = 0;
    }
    return;
  }

  double Fx, Fy, Fz, rho;
  total_force<DIM, false, LATTICE, PROG> (T, fld, S, L, x, y, z, vx, vy, vz, mass, volume,
					 Fx, Fy, Fz, rho);
  if (REC) {
    P.fx[i] = Fx; P.fy[i] = Fy; P.fz[i] = Fz;
  }
  if (!PROG && S.mutates_mass)
    P.mass[i] = mass;

  /* x += v dt/2 ; v += F dt/m ; x += v dt/2   (:828-839) */
  const double hdt = 0.5*S.dt, dtm = S.dt*__drcp_rn (mass);
  x = fma (vx, hdt, x); vx = fma (Fx, dtm, vx); x = fma (vx, hdt, x);
  y = fma (vy, hdt, y); vy = fma (Fy, dtm, vy); y = fma (vy, hdt, y);
  if (DIM == 3) {
    z = fma (vz, hdt, z); vz = fma (Fz, dtm, vz); z = fma (vz, hdt, z);
  }
  if (S.track_escapes && left_domain<DIM, LATTICE> (T, x, y, z))
    record_escape<DIM> (S, P, i);
  __stcs (P.x + i, x); __stcs (P.y + i, y); __stcs (P.vx + i, vx); __stcs (P.vy + i, vy);
  if (DIM == 3) {
    __stcs (P.z + i, z); __stcs (P.vz + i, vz);
  }
}

/* ------------------------------------------------------------------ */
/* The same step with the particle stream staged through shared memory by the
 * TMA engine (cp.async.bulk + mbarrier), as a persistent kernel.
 *
 * Profiling the plain kernel above showed it latency-bound, not issue- or
 * bandwidth-bound: with ~24 resident warps per SM each holding 2 KB of loads
 * in flight only while in its load phase, about 40 % of the bytes-in-flight
 * needed to saturate HBM3e were outstanding.  Here every CTA keeps STAGES
 * tiles (256 particles x 8 fp64 columns = 16 KB each) in flight at all times,
 * independent of what its warps are computing, at no register cost; the
 * compute part then starts from 29-cycle LDS instead of ~800-cycle LDG. */

namespace pipe {

__device__ __forceinline__ uint32_t smem_u32 (const void * p)
{
  return (uint32_t) __cvta_generic_to_shared (p);
}

__device__ __forceinline__ void mbar_init (uint64_t * bar, unsigned count)
{
  asm volatile ("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32 (bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_expect_tx (uint64_t * bar, unsigned bytes)
{
  asm volatile ("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
		:: "r"(smem_u32 (bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_wait (uint64_t * bar, unsigned parity)
{
  asm volatile ("{\n\t.reg .pred p;\n\t"
		"WAIT_%=:\n\t"
		"mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
		"@p bra DONE_%=;\n\t"
		"bra WAIT_%=;\n\t"
		"DONE_%=:\n\t}"
		:: "r"(smem_u32 (bar)), "r"(parity) : "memory");
}

__device__ __forceinline__ uint64_t policy_evict_first ()
{
  uint64_t pol;
  asm volatile ("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}

/* global -> shared bulk copy, completion signalled on `bar` (SASS: UBLKCP) */
__device__ __forceinline__ void bulk_g2s (void * dst, const void * src, unsigned bytes,
					  uint64_t * bar, uint64_t policy)
{
  asm volatile ("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes.L2::cache_hint "
		"[%0], [%1], %2, [%3], %4;"
		:: "r"(smem_u32 (dst)), "l"(src), "r"(bytes), "r"(smem_u32 (bar)), "l"(policy) : "memory");
}

/* the same without a cache hint */
__device__ __forceinline__ void bulk_g2s_plain (void * dst, const void * src, unsigned bytes, uint64_t * bar)
{
  asm volatile ("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
		:: "r"(smem_u32 (dst)), "l"(src), "r"(bytes), "r"(smem_u32 (bar)) : "memory");
}

__device__ __forceinline__ void fence_async_shared ()
{
  asm volatile ("fence.proxy.async.shared::cta;" ::: "memory");
}

} // namespace pipe

#define PIPE_ALIGN 256   /* particle columns are padded to this many elements */

template <int DIM, bool LATTICE, unsigned PROG, int STAGES, int MINB, int PIPE_TILE>
__global__ void __launch_bounds__(PIPE_TILE, MINB)
step_kernel_pipe (DevTree T, DevField fld, DevParticles P, DevStep S, int n_tiles)
{
  constexpr int NC = DIM == 3 ? 8 : 6;
  constexpr unsigned COL_BYTES = PIPE_TILE*sizeof (double);
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double (* buf)[NC][PIPE_TILE] = reinterpret_cast<double (*)[NC][PIPE_TILE]> (smem_raw);
  __shared__ uint64_t full[STAGES];
  const int tid = threadIdx.x;

  const double * col[NC];
  if (DIM == 3) {
    col[0] = P.x; col[1] = P.y; col[2] = P.z; col[3] = P.vx; col[4] = P.vy; col[5] = P.vz;
    col[6] = P.mass; col[7] = P.volume;
  }
  else {
    col[0] = P.x; col[1] = P.y; col[2] = P.vx; col[3] = P.vy; col[4] = P.mass; col[5] = P.volume;
  }

  uint64_t policy = 0;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; s++)
      pipe::mbar_init (&full[s], 1);
    pipe::fence_async_shared ();
    policy = pipe::policy_evict_first ();
  }
  __syncthreads ();

  auto issue = [&] (int s, int tile) {
    pipe::mbar_expect_tx (&full[s], NC*COL_BYTES);
#pragma unroll
    for (int c = 0; c < NC; c++)
      pipe::bulk_g2s (&buf[s][c][0], col[c] + (int64_t) tile*PIPE_TILE, COL_BYTES, &full[s], policy);
  };

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; s++) {
      const int tile = blockIdx.x + s*gridDim.x;
      if (tile < n_tiles)
	issue (s, tile);
    }
  }

  int s = 0;
  unsigned parity = 0;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    pipe::mbar_wait (&full[s], parity);
    double x, y, z = 0., vx, vy, vz = 0., mass, volume;
    if (DIM == 3) {
      x = buf[s][0][tid]; y = buf[s][1][tid]; z = buf[s][2][tid];
      vx = buf[s][3][tid]; vy = buf[s][4][tid]; vz = buf[s][5][tid];
      mass = buf[s][6][tid]; volume = buf[s][7][tid];
    }
    else {
      x = buf[s][0][tid]; y = buf[s][1][tid]; vx = buf[s][2][tid]; vy = buf[s][3][tid];
      mass = buf[s][4][tid]; volume = buf[s][5][tid];
    }
    /* The stage is released as soon as every thread holds its particle in
       registers, so the warps of a CTA run the long compute part unsynchronised
       (a barrier at the END of the tile, tried with a one-tile lookahead that
       located and prefetched the next tile's table lines, was 30 % slower:
       profiles/README.md). */
    __syncthreads ();
    if (tid == 0) {
      const int next = tile + STAGES*gridDim.x;
      if (next < n_tiles) {
	pipe::fence_async_shared ();       /* generic-proxy reads before async-proxy writes */
	issue (s, next);
      }
    }
    if (++s == STAGES) { s = 0; parity ^= 1; }

    const int64_t i = (int64_t) tile*PIPE_TILE + tid;
    if (i < P.n) {
      const Located L = locate<DIM, LATTICE> (T, x, y, z);
      if (L.cell >= 0) {
	double Fx, Fy, Fz, rho;
	total_force<DIM, false, LATTICE, PROG> (T, fld, S, L, x, y, z, vx, vy, vz, mass, volume,
					       Fx, Fy, Fz, rho);
	if (!PROG && S.mutates_mass)
	  P.mass[i] = mass;
	const double hdt = 0.5*S.dt, dtm = S.dt*__drcp_rn (mass);
	x = fma (vx, hdt, x); vx = fma (Fx, dtm, vx); x = fma (vx, hdt, x);
	y = fma (vy, hdt, y); vy = fma (Fy, dtm, vy); y = fma (vy, hdt, y);
	if (DIM == 3) {
	  z = fma (vz, hdt, z); vz = fma (Fz, dtm, vz); z = fma (vz, hdt, z);
	}
	if (S.track_escapes && left_domain<DIM, LATTICE> (T, x, y, z))
	  record_escape<DIM> (S, P, i);
	__stcs (P.x + i, x); __stcs (P.y + i, y); __stcs (P.vx + i, vx); __stcs (P.vy + i, vy);
	if (DIM == 3) {
	  __stcs (P.z + i, z); __stcs (P.vz + i, vz);
	}
      }
      else if (S.track_escapes) {
	atomicAdd (S.esc_count + 3, 1);       /* outside the domain before the step */
	if (S.keep) S.keep[i] = 0;
      }
    }
  }
}

/* ------------------------------------------------------------------ */
/* Warp-private variant of the TMA pipeline.  Each WARP owns STAGES tiles of 32 particles
 * (8 bulk copies of 256 B per tile) with its own mbarriers: no CTA-wide barrier per tile,
 * and the stage is released only at the END of the tile, so velocity, mass and volume are
 * read from the staged tile when the force needs them (after the vertex gathers) and the
 * position is re-read for the integration -- 16 registers fewer live across the gathers. */

namespace pipe {

/* elect.sync: true in exactly one lane of the (converged) warp.  ptxas knows that, so a bulk
 * copy issued under it is a single UBLKCP; under `lane == 0` it wraps every copy in a loop over
 * the active lanes (ELECT / R2UR / BRA.U.ANY: 90 of 593 instructions per tile). */
__device__ __forceinline__ bool elect_one ()
{
  unsigned pred;
  asm volatile ("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

/* shared-memory load that cannot be scheduled before `dep` is known: column COL of the staged tile,
 * from the shared-window address of the lane's slot in column 0 (computed once per tile) */
template <int COL>
__device__ __forceinline__ double lds_col_after (uint32_t lane_addr, double dep)
{
  double v;
  asm volatile ("ld.shared.f64 %0, [%1+%2];   // after %3" : "=d"(v) : "r"(lane_addr), "n"(COL*32*8), "d"(dep));
  return v;
}

} // namespace pipe

template <int DIM>
struct LateShared {
  uint32_t b;                  /* shared-window address of this lane's slot in column 0 of the staged
				  tile; columns 32 doubles apart */
  double * keep;               /* the caller's vx, vy, vz (total_force takes them by value) */
  __device__ __forceinline__ void fetch (double dep, double & vx, double & vy, double & vz,
					 double & mass, double & volume) const
  {
    constexpr int V0 = DIM == 3 ? 3 : 2;
    keep[0] = vx = pipe::lds_col_after<V0> (b, dep);
    keep[1] = vy = pipe::lds_col_after<V0 + 1> (b, dep);
    if (DIM == 3) keep[2] = vz = pipe::lds_col_after<V0 + 2> (b, dep);
    mass = pipe::lds_col_after<2*DIM> (b, dep);
    volume = pipe::lds_col_after<2*DIM + 1> (b, dep);
  }
};

/* REC: also store the accumulated force (GfsParticulate.force) -- three more column stores.
 * DEP: the two-way deposits of the same step, fused: after the integration the particle is
 * located again at its NEW position, the on-fluid force is evaluated there (new velocity, same
 * field) and void fraction + force are reduced into the deposit buffer -- what
 * gfsb200_deposit_all does right after the step, without streaming the particles a second time
 * (136 instead of 112 + 88 algorithmic bytes per particle-step). */
/* STREAM: the particle stream is marked as such -- evict-first bulk loads, streaming stores -- so that
 * it leaves the tables their L2 lines.  That pays when the tables fit the L2 (the adaptive configs:
 * C5 with 200 M particles loses 6 % without it) and in the fused step + deposit kernel, whose second
 * evaluation re-reads them; on a big uniform tree (C2: 102 MB of tables against 126 MB of L2) the plain
 * step kernel is 1.5 % faster with ordinary loads and stores (profiles/README.md, round 2). */
template <int DIM, bool LATTICE, unsigned PROG, int STAGES, int MINB, int WPIPE_WARPS, bool REC, bool DEP,
	  bool STREAM = true>
__global__ void __launch_bounds__(32*WPIPE_WARPS, MINB)
step_kernel_wpipe (DevTree T, DevField fld, DevParticles P, DevStep S, int n_tiles, DevDeposit D)
{
  constexpr int NC = DIM == 3 ? 8 : 6;
  constexpr unsigned COL_BYTES = 32*sizeof (double);
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double (* buf)[STAGES][NC][32] = reinterpret_cast<double (*)[STAGES][NC][32]> (smem_raw);
  __shared__ uint64_t full[WPIPE_WARPS][STAGES];
  const int lane = threadIdx.x & 31;
  /* broadcast through a shuffle so that the compiler knows the warp index -- and with it the
     tile number, the stage addresses and the operands of the bulk copies -- to be warp-uniform:
     without it every UBLKCP sits in a lane-serialising loop (6 R2UR + ELECT + branch per copy,
     90 of 593 instructions per tile) */
  const int warp = __shfl_sync (0xffffffffu, (int) (threadIdx.x >> 5), 0);

  const double * col[NC];
  if (DIM == 3) {
    col[0] = P.x; col[1] = P.y; col[2] = P.z; col[3] = P.vx; col[4] = P.vy; col[5] = P.vz;
    col[6] = P.mass; col[7] = P.volume;
  }
  else {
    col[0] = P.x; col[1] = P.y; col[2] = P.vx; col[3] = P.vy; col[4] = P.mass; col[5] = P.volume;
  }

  const int first = blockIdx.x*WPIPE_WARPS + warp, stride = gridDim.x*WPIPE_WARPS;
  const uint64_t policy = STREAM ? pipe::policy_evict_first () : 0;

  auto issue = [&] (int s, int tile) {
    pipe::mbar_expect_tx (&full[warp][s], NC*COL_BYTES);
#pragma unroll
    for (int c = 0; c < NC; c++) {
      if (STREAM)
	pipe::bulk_g2s (&buf[warp][s][c][0], col[c] + (int64_t) tile*32, COL_BYTES, &full[warp][s], policy);
      else
	pipe::bulk_g2s_plain (&buf[warp][s][c][0], col[c] + (int64_t) tile*32, COL_BYTES, &full[warp][s]);
    }
  };

  if (pipe::elect_one ()) {
#pragma unroll
    for (int s = 0; s < STAGES; s++)
      pipe::mbar_init (&full[warp][s], 1);
    pipe::fence_async_shared ();
#pragma unroll
    for (int s = 0; s < STAGES; s++)
      if (first + s*stride < n_tiles)
	issue (s, first + s*stride);
  }
  __syncwarp ();
  /* Programmatic dependent launch: the kernel may have been started while the cell pass (the
     previous kernel of the stream) was still draining -- everything above touches only the
     particle stream, which that kernel does not write.  The tables are read from here on: wait
     for the prerequisite grid and its memory.  (A no-op for an ordinary launch.) */
  asm volatile ("griddepcontrol.wait;" ::: "memory");

  int s = 0;
  unsigned parity = 0;
  for (int tile = first; tile < n_tiles; tile += stride) {
    pipe::mbar_wait (&full[warp][s], parity);
    /* (A one-tile lookahead -- locate the next tile's particles in the other stage and prefetch
       their table rows into L1 -- measured 2 % SLOWER on C2, C3 and 2D, round 2: the wait at the
       first use of the gathered values is queueing in the L1 data pipe, not L2 latency.) */
    const double * b = &buf[warp][s][0][lane];
    const uint32_t sb = pipe::smem_u32 (b);
    const int64_t i = (int64_t) tile*32 + lane;
    int dcell = -1;
    double av = 0., ax = 0., ay = 0., az = 0.;
    if (i < P.n) {
      double x = b[0], y = b[32], z = DIM == 3 ? b[64] : 0.;
      const Located L = locate<DIM, LATTICE, DEP> (T, x, y, z);
      if (REC && P.cell)
	P.cell[i] = cell_index<DIM, LATTICE> (T, L);
      if (L.cell >= 0) {
	double Fx, Fy, Fz, rho;
	double vx = 0., vy = 0., vz = 0., mass = 0., volume = 0.;
	double vkeep[3];
	double r3 = 0.;
	total_force<DIM, false, LATTICE, PROG, LateShared<DIM> > (T, fld, S, L, x, y, z, vx, vy, vz, mass, volume,
								 Fx, Fy, Fz, rho, LateShared<DIM> { sb, vkeep },
								 DEP ? &r3 : NULL);
	if (!PROG && S.mutates_mass)
	  P.mass[i] = mass;
	if (REC) {
	  __stcs (P.fx + i, Fx); __stcs (P.fy + i, Fy); __stcs (P.fz + i, Fz);
	}
	/* the position comes back from the staged tile, the velocity from the late fetch */
	x = pipe::lds_col_after<0> (sb, Fx); y = pipe::lds_col_after<1> (sb, Fx);
	if (DIM == 3) z = pipe::lds_col_after<2> (sb, Fx);
	vx = vkeep[0]; vy = vkeep[1];
	if (DIM == 3) vz = vkeep[2];
	const double hdt = 0.5*S.dt, dtm = S.dt*__drcp_rn (mass);
	x = fma (vx, hdt, x); vx = fma (Fx, dtm, vx); x = fma (vx, hdt, x);
	y = fma (vy, hdt, y); vy = fma (Fy, dtm, vy); y = fma (vy, hdt, y);
	if (DIM == 3) {
	  z = fma (vz, hdt, z); vz = fma (Fz, dtm, vz); z = fma (vz, hdt, z);
	}
	if (S.track_escapes && left_domain<DIM, LATTICE> (T, x, y, z))
	  record_escape<DIM> (S, P, i);
	if (STREAM) {
	  __stcs (P.x + i, x); __stcs (P.y + i, y); __stcs (P.vx + i, vx); __stcs (P.vy + i, vy);
	  if (DIM == 3) {
	    __stcs (P.z + i, z); __stcs (P.vz + i, vz);
	  }
	}
	else {
	  P.x[i] = x; P.y[i] = y; P.vx[i] = vx; P.vy[i] = vy;
	  if (DIM == 3) { P.z[i] = z; P.vz[i] = vz; }
	}
	if (DEP) {
	  /* volume is re-read from the staged tile (total_force took it by value) */
	  const double vol2 = pipe::lds_col_after<2*DIM + 1> (sb, vx);
	  dcell = deposit_terms<DIM, LATTICE, PROG, true, true, true> (T, fld, S, x, y, z, vx, vy, vz, mass, vol2,
								 av, ax, ay, az, &r3);
	  if (!PROG && S.mutates_mass && dcell >= 0)
	    P.mass[i] = mass;
	}
      }
      else if (S.track_escapes) {
	atomicAdd (S.esc_count + 3, 1);       /* outside the domain before the step */
	if (S.keep) S.keep[i] = 0;
      }
    }
    if (DEP)
      run_deposit<DIM, true, true> (D, dcell, av, ax, ay, az);
    /* every lane has read what it needs: hand the stage back to the copy engine */
    __syncwarp ();
    const int next = tile + STAGES*stride;
    if (next < n_tiles && pipe::elect_one ()) {
      pipe::fence_async_shared ();         /* generic-proxy reads before async-proxy writes */
      issue (s, next);
    }
    if (++s == STAGES) { s = 0; parity ^= 1; }
  }
}

/* passive tracers: gfs_domain_advect_point, src/domain.c:2764-2788 (RK2) */
template <int DIM, bool REC_CELL>
__global__ void __launch_bounds__(256)
advect_kernel (DevTree T, DevField fld, DevParticles P, double dt)
{
  const int64_t i = (int64_t) blockIdx.x*blockDim.x + threadIdx.x;
  if (i >= P.n)
    return;
  const double x = P.x[i], y = P.y[i], z = DIM == 3 ? P.z[i] : 0.;
  Located L = locate<DIM> (T, x, y, z);
  if (REC_CELL)
    P.cell[i] = L.cell;
  if (L.cell < 0)
    return;
  double u, v, w;
  interpolate<DIM> (T, fld, L, x, y, z, u, v, w);
  const double x1 = x + dt*u/2., y1 = y + dt*v/2., z1 = DIM == 3 ? z + dt*w/2. : 0.;
  L = locate<DIM> (T, x1, y1, z1);
  if (L.cell < 0)
    return;
  interpolate<DIM> (T, fld, L, x1, y1, z1, u, v, w);
  P.x[i] = x + dt*u;
  P.y[i] = y + dt*v;
  if (DIM == 3) P.z[i] = z + dt*w;
}

template <int DIM>
__global__ void __launch_bounds__(256)
locate_kernel (DevTree T, int64_t n, const double * __restrict__ x, const double * __restrict__ y,
	       const double * __restrict__ z, int32_t * __restrict__ cell)
{
  const int64_t i = (int64_t) blockIdx.x*blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  cell[i] = locate<DIM> (T, x[i], y[i], DIM == 3 ? z[i] : 0.).cell;
}

template <int DIM>
__global__ void __launch_bounds__(256)
interpolate_kernel (DevTree T, DevField fld, int64_t n, const double * __restrict__ x,
		    const double * __restrict__ y, const double * __restrict__ z,
		    double * __restrict__ u, double * __restrict__ v, double * __restrict__ w)
{
  const int64_t i = (int64_t) blockIdx.x*blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  const double px = x[i], py = y[i], pz = DIM == 3 ? z[i] : 0.;
  const Located L = locate<DIM> (T, px, py, pz);
  double a = GFSB200_NODATA, b = GFSB200_NODATA, c = GFSB200_NODATA;
  if (L.cell >= 0) {
    interpolate<DIM> (T, fld, L, px, py, pz, a, b, c);
    /* gfs_interpolate returns NODATA when the cell's own value is (:2704-2705) */
    if (fld.u[0][L.cell] == GFSB200_NODATA) a = GFSB200_NODATA;
    if (fld.u[1][L.cell] == GFSB200_NODATA) b = GFSB200_NODATA;
    if (DIM == 3 && fld.u[2][L.cell] == GFSB200_NODATA) c = GFSB200_NODATA;
  }
  if (u) u[i] = a;
  if (v) v[i] = b;
  if (w && DIM == 3) w[i] = c;
}

/* corner values of a list of leaves, for the parity tests:
 * out[j][k] = vtx_val[leaf_vtx[cells[j]][k]].comp (NODATA resolved per leaf) */
template <int DIM>
__global__ void corner_values_kernel (DevTree T, DevField fld, int comp, int64_t n,
				      const int32_t * __restrict__ cells, double * __restrict__ out)
{
  const int64_t j = (int64_t) blockIdx.x*blockDim.x + threadIdx.x;
  const int nc = 1 << DIM;
  if (j >= n*nc)
    return;
  const int cell = cells[j/nc], k = (int) (j % nc);
  const int v = T.leaf_vtx[(int64_t) cell*nc + k];
  double val = GFSB200_NODATA;
  if (v >= 0) {
    val = DIM == 3 ? (comp < 2 ? fld.vtx_val[(int64_t) v*2 + comp] : fld.vtx_val[(int64_t) T.n_vertices*2 + v])
      : fld.vtx_val[(int64_t) v*2 + comp];
    val = resolve (val, fld.u[comp], cell);
  }
  out[j] = val;
}

/* ------------------------------------------------------------------ */
/* One pass for both deposits of a two-way step (see deposit_terms): one locate, one
 * interpolation set, same compile-time force programs and lattice addressing as the step kernel;
 * one reduction per run of equal cells.  INDEXED: the pass runs over a list of particle indices
 * (the particles gfs_particle_bc has just wrapped, after a fused step + deposit) instead of the
 * whole list; keep[i] == 0 marks a particle that was dropped. */
template <int DIM, bool LATTICE, unsigned PROG, bool VOL, bool FORCE, bool INDEXED>
__global__ void __launch_bounds__(256, 3)
deposit_kernel (DevTree T, DevField fld, DevParticles P, DevStep S, DevDeposit D,
		int n_idx, const int32_t * __restrict__ idx, const uint8_t * __restrict__ keep)
{
  int64_t i = (int64_t) blockIdx.x*blockDim.x + threadIdx.x;
  bool live = INDEXED ? i < n_idx : i < P.n;
  if (INDEXED && live) {
    i = idx[i];
    live = keep[i] != 0;
  }
  int cell = -1;
  double av = 0., ax = 0., ay = 0., az = 0.;
  if (live) {
    const double x = __ldcs (P.x + i), y = __ldcs (P.y + i), z = DIM == 3 ? __ldcs (P.z + i) : 0.;
    double mass = FORCE ? __ldcs (P.mass + i) : 0.;
    cell = deposit_terms<DIM, LATTICE, PROG, VOL, FORCE> (T, fld, S, x, y, z,
							 FORCE ? __ldcs (P.vx + i) : 0., FORCE ? __ldcs (P.vy + i) : 0.,
							 FORCE && DIM == 3 ? __ldcs (P.vz + i) : 0.,
							 mass, __ldcs (P.volume + i), av, ax, ay, az);
    if (FORCE && !PROG && S.mutates_mass && cell >= 0)       /* compute_forces_onfluid runs GfsForceAddedMass too */
      P.mass[i] = mass;
  }
  run_deposit<DIM, VOL, FORCE> (D, cell, av, ax, ay, az);
}

/* ------------------------------------------------------------------ */
/* gfs_particle_bc on the escaped particles                             */

/* check_intersetion, modules/particulatecommon.c:3058-3148: the first face
 * direction d (in order) through which the segment p0 -> p1 leaves the cell */
template <int DIM>
__device__ __forceinline__ int exit_face (const double c[3], double size, const double p0[3],
					  const double p1[3])
{
  for (int d = 0; d < 2*DIM; d++) {
    const double normal = d & 1 ? -1. : 1.;
    const int a = d >> 1;
    const double dp = p1[a] - p0[a];
    if (dp != 0. && normal*dp > 0.) {
      const double t = (c[a] + normal*size*0.5 - p0[a])/dp;
      bool ok = t*(t - 1.) <= 0.;
      for (int b = 0; b < DIM; b++)
	if (b != a) {
	  const double q = p0[b] + t*(p1[b] - p0[b]);
	  ok = ok && (q - c[b] + size*0.5)*(q - c[b] - size*0.5) <= 0.;
	}
      if (ok)
	return d;
    }
  }
  return -1;
}

template <int DIM>
__global__ void __launch_bounds__(128)
particle_bc_kernel (DevTree T, DevParticles P, int n_esc, const int32_t * __restrict__ esc_idx,
		    const double * __restrict__ esc_old, uint8_t * __restrict__ keep,
		    int * __restrict__ counters /* [0] wrapped, [1] dropped */)
{
  const int k = blockIdx.x*blockDim.x + threadIdx.x;
  if (k >= n_esc)
    return;
  const int64_t i = esc_idx[k];
  const double p0[3] = { esc_old[3*k], esc_old[3*k + 1], esc_old[3*k + 2] };
  double p1[3] = { P.x[i], P.y[i], DIM == 3 ? P.z[i] : 0. };
  /* boundarycell, :3151-3186: walk from the cell of pos_old along the path */
  const Located L = locate<DIM> (T, p0[0], p0[1], p0[2]);
  int d = -1;
  int cell = L.cell;
  bool ghost = false;        /* the walk ended at a ghost cell of a GfsBoundary */
  if (cell >= 0) {
    double c[3] = { L.cx, L.cy, L.cz };
    double size = 2.*L.half;
    for (int it = 0; it < 4096; it++) {
      d = exit_face<DIM> (c, size, p0, p1);
      if (d < 0)
	break;
      const int nb = T.neighbor[(int64_t) cell*(2*DIM) + d];
      if (nb < 0)            /* the hull without a GfsBoundary, or a destroyed (solid) cell: dropped (:3259-3265) */
	break;
      if (T.info[nb] & GFSB200_CELL_BOUNDARY) {
	ghost = true;
	break;
      }
      /* centre of the neighbour: same level, or one level coarser */
      const double dir = d & 1 ? -1. : 1.;
      if (T.level[nb] == T.level[cell])
	c[d >> 1] += dir*size;
      else {
	const int n = T.info[cell] >> 4;            /* child id: where `cell` sits in its parent */
	c[0] -= (n & 1 ? 1. : -1.)*size*0.5;
	c[1] -= (n & 2 ? -1. : 1.)*size*0.5;
	if (DIM == 3) c[2] -= (n & 4 ? -1. : 1.)*size*0.5;
	size *= 2.;
	c[d >> 1] += dir*size;
      }
      cell = nb;
    }
  }
  int match = -1, root = -1;
  if (cell >= 0 && d >= 0 && ghost) {
    root = cell;
    while (T.parent[root] >= 0)
      root = T.parent[root];
    match = T.periodic[root][d];
  }
  if (match >= 0) {
    /* periodic_bc_particle, :3189-3214 */
    const int a = d >> 1;
    const double normal = d & 1 ? -1. : 1.;
    const double size = T.root_size;
    const double box_face = T.root_pos[root][a] + normal*size/2.;
    const double box_face_nbr = T.root_pos[match][a] - normal*size/2.;
    const double tolerance = size/1.e8;
    const double distance = (p1[a] - box_face)*normal;
    p1[a] = box_face_nbr + distance + normal*tolerance;
    if (a == 0) P.x[i] = p1[0];
    else if (a == 1) P.y[i] = p1[1];
    else P.z[i] = p1[2];
    atomicAdd (counters, 1);
  }
  else {
    keep[i] = 0;
    atomicAdd (counters + 1, 1);
  }
}

/* ------------------------------------------------------------------ */
/* permutation gathers (sort by cell, cull)                             */

__global__ void __launch_bounds__(256)
gather_kernel (int64_t n, const int32_t * __restrict__ perm, int ncols,
	       const double * const * __restrict__ src, double * const * __restrict__ dst,
	       const uint32_t * __restrict__ id_src, uint32_t * __restrict__ id_dst)
{
  const int64_t i = (int64_t) blockIdx.x*blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  const int32_t j = perm[i];
  for (int c = 0; c < ncols; c++)
    dst[c][i] = src[c][j];
  id_dst[i] = id_src[j];
}

__global__ void iota_kernel (int64_t n, int32_t * __restrict__ a)
{
  const int64_t i = (int64_t) blockIdx.x*blockDim.x + threadIdx.x;
  if (i < n) a[i] = (int32_t) i;
}

__global__ void iota_u32_kernel (int64_t n, uint32_t * __restrict__ a, uint32_t base)
{
  const int64_t i = (int64_t) blockIdx.x*blockDim.x + threadIdx.x;
  if (i < n) a[i] = base + (uint32_t) i;
}

__global__ void sort_keys_kernel (int64_t n, const int32_t * __restrict__ cell,
				  uint32_t * __restrict__ key, uint32_t outside_key)
{
  const int64_t i = (int64_t) blockIdx.x*blockDim.x + threadIdx.x;
  if (i < n) key[i] = cell[i] < 0 ? outside_key : (uint32_t) cell[i];
}

/* key[i] = owner of the cell particle i sits in (sorted cell keys of the last sort; a key >= n_cells:
 * outside the domain -> `outside_key') */
__global__ void owner_keys_kernel (int64_t n, const uint32_t * __restrict__ cell_key, uint32_t n_cells,
				   const uint8_t * __restrict__ owner_of, uint32_t * __restrict__ key,
				   uint32_t outside_key)
{
  const int64_t i = (int64_t) blockIdx.x*blockDim.x + threadIdx.x;
  if (i < n) {
    const uint32_t c = cell_key[i];
    key[i] = c < n_cells ? (uint32_t) owner_of[c] : outside_key;
  }
}

__global__ void inside_flag_kernel (int64_t n, const int32_t * __restrict__ cell,
				    uint8_t * __restrict__ flag)
{
  const int64_t i = (int64_t) blockIdx.x*blockDim.x + threadIdx.x;
  if (i < n) flag[i] = cell[i] >= 0;
}

/* ------------------------------------------------------------------ */
/* GfsSourceParticulate with its smoothing kernel                        */
/* modules/particulatecommon.c:2087-2228                                 */

/* stage 1: particulate->force = sum of the forces acting on the fluid
 * (compute_forces_onfluid :753-765, every force but GfsForceBuoy); zero for a
 * particle outside the domain (each force model returns 0 when
 * gfs_domain_locate is NULL) */
template <int DIM, bool LATTICE>
__global__ void __launch_bounds__(256, 3)
onfluid_force_kernel (DevTree T, DevField fld, DevParticles P, DevStep S)
{
  const int64_t i = (int64_t) blockIdx.x*blockDim.x + threadIdx.x;
  if (i >= P.n)
    return;
  const double x = P.x[i], y = P.y[i], z = DIM == 3 ? P.z[i] : 0.;
  const Located L = locate<DIM, LATTICE> (T, x, y, z);
  double Fx = 0., Fy = 0., Fz = 0.;
  if (L.cell >= 0) {
    double rho, mass = P.mass[i];
    total_force<DIM, true, LATTICE, 0> (T, fld, S, L, x, y, z, P.vx[i], P.vy[i], DIM == 3 ? P.vz[i] : 0.,
				       mass, P.volume[i], Fx, Fy, Fz, rho);
    if (S.mutates_mass)
      P.mass[i] = mass;
  }
  P.fx[i] = Fx; P.fy[i] = Fy; P.fz[i] = Fz;
}

/* The kernel GfsFunction evaluated at the normalised offset
 * (gfs_function_spatial_value of distance_normalization, :2087-2098), in the
 * reference's operation order: round-to-nearest intrinsics are never
 * contracted into FMAs. */
__device__ __forceinline__ double kernel_value (const gfsb200_kernel & K, double x, double y, double z)
{
  const double r2 = __dadd_rn (__dadd_rn (__dmul_rn (x, x), __dmul_rn (y, y)), __dmul_rn (z, z));
  if (K.kind == GFSB200_KERNEL_GAUSSIAN)
    return __dmul_rn (K.a, exp (__dmul_rn (-K.b, r2)));
  if (K.kind == GFSB200_KERNEL_COMPACT) {
    const double t = __dsub_rn (1., __dmul_rn (K.b, r2));
    if (t <= 0.) return 0.;
    double v = K.a;
    for (int i = 0; i < K.p; i++) v = __dmul_rn (v, t);
    return v;
  }
  return K.a;
}

/* gfs_domain_cell_traverse_condition (FTT_PRE_ORDER, FTT_TRAVERSE_LEAFS, -1)
 * with cond_kernel (:2126-2156) as the pruning condition, over every GfsBox
 * tree (src/domain.c:1516-1574, src/ftt.c:948-986): depth first, children in
 * FTT order, destroyed children skipped without evaluating the condition.
 * No stack: the child digits of the current path live in a 64-bit word, the
 * centre is tracked with exact dyadic adds and undone on the way up.
 * visit (cell, cx, cy, cz, half) is called for every leaf reached. */
template <int DIM, typename Visit>
__device__ __forceinline__ void traverse_kernel_support (const DevTree & T, double px, double py, double pz,
							 double rkernel, Visit visit)
{
  constexpr int NCH = 1 << DIM;
  const double SQ = DIM == 3 ? 1.7320508075688772 : 1.4142135623730951;   /* sqrt(3.), sqrt(2.) */
  for (int r = 0; r < T.n_box_roots; r++) {
    int cell = r, depth = 0;
    unsigned long long path = 0;
    double cx = T.root_pos[r][0], cy = T.root_pos[r][1], cz = DIM == 3 ? T.root_pos[r][2] : 0.;
    double size = 0.5*T.root_size;             /* ftt_cell_size (cell)/2. */
    for (;;) {
      const int c0 = __ldg (T.child0 + cell);
      bool descend = false;
      if (c0 != CHILD_DESTROYED) {
	const double dx = __dsub_rn (cx, px), dy = __dsub_rn (cy, py);
	double d2 = __dadd_rn (__dmul_rn (dx, dx), __dmul_rn (dy, dy));
	if (DIM == 3) {
	  const double dz = __dsub_rn (cz, pz);
	  d2 = __dadd_rn (d2, __dmul_rn (dz, dz));
	}
	/* ftt_vector_distance (..) - radeq <= distance, decided without the square root unless
	   d2 is within 1e-12 of (distance + radeq)^2 (then by the reference's exact expression) */
	const double radeq = __dmul_rn (size, SQ), reach = rkernel + radeq, reach2 = reach*reach;
	bool ok;
	if (d2 <= reach2*(1. - 1e-12))
	  ok = true;
	else if (d2 >= reach2*(1. + 1e-12))
	  ok = false;
	else
	  ok = __dsub_rn (__dsqrt_rn (d2), radeq) <= rkernel;
	if (!ok)          /* "check also if the bubble is inside the cell" */
	  ok = !(px > cx + size || px < cx - size || py > cy + size || py < cy - size ||
		 (DIM == 3 && (pz > cz + size || pz < cz - size)));
	if (ok) {
	  if (c0 < 0)
	    visit (cell, cx, cy, cz, size);
	  else
	    descend = true;
	}
      }
      if (descend) {
	size *= 0.5;
	cell = c0; depth++; path <<= DIM;
	cx -= size; cy += size;                /* child 0 sits at (-,+,+) */
	if (DIM == 3) cz += size;
	continue;
      }
      /* next sibling, or up until there is one */
      bool done = false;
      for (;;) {
	if (depth == 0) { done = true; break; }
	const int n = (int) (path & (NCH - 1));
	cx -= (n & 1) ? size : -size;
	cy -= (n & 2) ? -size : size;
	if (DIM == 3) cz -= (n & 4) ? -size : size;
	if (n < NCH - 1) {
	  const int m = n + 1;
	  cell++; path++;
	  cx += (m & 1) ? size : -size;
	  cy += (m & 2) ? -size : size;
	  if (DIM == 3) cz += (m & 4) ? -size : size;
	  break;
	}
	cell = __ldg (T.parent + cell);
	path >>= DIM; size *= 2.; depth--;
      }
      if (done) break;
    }
  }
}

/* stage 2 (source_particulate_event :2207-2221), one thread per particle:
 * pass 1 accumulates kd.volume and kd.correction in the reference's traversal
 * order (kernel_volume :2108-2119), pass 2 scatters
 *   u_c[cell] -= F_c/rho/V_cell * K/correction          (diffuse_force :2158-2175)
 * with fp64 atomics.  Neighbouring (cell-sorted) particles walk nearly the same
 * cells in lock step. */
template <int DIM>
__global__ void __launch_bounds__(128)
smoothed_deposit_kernel (DevTree T, DevField fld, DevParticles P, double rho_const, double rkernel,
			 gfsb200_kernel K, double * __restrict__ f0, double * __restrict__ f1,
			 double * __restrict__ f2, double * __restrict__ norm)
{
  const int64_t i = (int64_t) blockIdx.x*blockDim.x + threadIdx.x;
  if (i >= P.n)
    return;
  const double px = P.x[i], py = P.y[i], pz = DIM == 3 ? P.z[i] : 0.;
  const double Fx = P.fx[i], Fy = P.fy[i], Fz = DIM == 3 ? P.fz[i] : 0.;
  /* nothing to deposit (outside the domain, or no force): the reference would
     subtract zeros */
  if (!norm && Fx == 0. && Fy == 0. && Fz == 0.)
    return;
  const double rb = pow (__ddiv_rn (__dmul_rn (3., P.volume[i]), 4.*M_PI), 1./3.);
  /* offsets are scaled by 1/r_b with a multiplication (the reference divides): device pow
     already differs from libm's by an ulp, so the kernel argument is not bit-reproducible
     either way; the deposit stays within its 1e-12 bar */
  const double inv_rb = 1./rb;
  const bool fixz = (K.flags & GFSB200_KERNEL_FIX_Z) != 0;
  const double qz_const = DIM == 3 ? -pz*inv_rb : 0.;        /* the reference's (0 - z_p)/r_b */

  double volume = 0., correction = 0.;
  traverse_kernel_support<DIM> (T, px, py, pz, rkernel,
    [&] (int cell, double cx, double cy, double cz, double half) {
      const double h = 2.*half;
      double cellvol = DIM == 3 ? h*h*h : h*h;                        /* powers of two: exact */
      if (T.solid_a)                                                  /* gfs_cell_volume, src/domain.h:503-508 */
	cellvol = __dmul_rn (cellvol, T.solid_a[cell]);
      volume = __dadd_rn (volume, cellvol);
      const double qx = (cx - px)*inv_rb, qy = (cy - py)*inv_rb;
      const double qz = DIM == 3 ? (fixz ? (cz - pz)*inv_rb : qz_const) : 0.;
      correction = __dadd_rn (correction, __dmul_rn (kernel_value (K, qx, qy, qz), cellvol));
    });
  correction = __ddiv_rn (correction, volume);
  if (norm) {
    norm[i] = correction;
    norm[P.n + i] = volume;
  }
  if (!(correction > 1.e-10))
    return;
  /* F/rho/V_cell*K/correction: V_cell is a power of two, 1/correction is hoisted */
  const double inv_corr = 1./correction;
  const double inv_rho = fld.alpha ? 0. : 1./rho_const;
  traverse_kernel_support<DIM> (T, px, py, pz, rkernel,
    [&] (int cell, double cx, double cy, double cz, double half) {
      const double inv_h = __longlong_as_double ((2045LL << 52) - __double_as_longlong (half));   /* 1/(2 half) */
      double inv_cellvol = DIM == 3 ? inv_h*inv_h*inv_h : inv_h*inv_h;
      if (T.solid_a)
	inv_cellvol = inv_cellvol/T.solid_a[cell];                    /* mixed cell: V_cell = h^dim a */
      const double ir = fld.alpha ? fld.alpha[cell] : inv_rho;       /* 1/rho = alpha */
      const double qx = (cx - px)*inv_rb, qy = (cy - py)*inv_rb;
      const double qz = DIM == 3 ? (fixz ? (cz - pz)*inv_rb : qz_const) : 0.;
      const double s = -kernel_value (K, qx, qy, qz)*inv_corr*inv_cellvol*ir;
      atomicAdd (f0 + cell, Fx*s);
      atomicAdd (f1 + cell, Fy*s);
      if (DIM == 3)
	atomicAdd (f2 + cell, Fz*s);
    });
}

inline unsigned grid_for (int64_t n, int threads) { return (unsigned) ((n + threads - 1)/threads); }

} // namespace

/* ------------------------------------------------------------------ */
/* launchers (C linkage, called from capi.cu)                           */

template <int DIM, bool LA, unsigned PR, int ST, int MB, int TILE>
static void launch_pipe (const DevTree * T, const DevField * F, const DevParticles * P,
			 const DevStep * S, int n_sm, cudaStream_t st)
{
  const int n_tiles = (int) ((P->n + TILE - 1)/TILE);
  const size_t smem = (size_t) ST*(DIM == 3 ? 8 : 6)*TILE*sizeof (double);
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute (step_kernel_pipe<DIM, LA, PR, ST, MB, TILE>,
			  cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    configured = true;
  }
  int grid = n_sm*MB;                   /* persistent: MB CTAs of TILE threads per SM */
  if (grid > n_tiles) grid = n_tiles;
  step_kernel_pipe<DIM, LA, PR, ST, MB, TILE><<<grid, TILE, smem, st>>> (*T, *F, *P, *S, n_tiles);
}

template <int DIM, bool LA, unsigned PR, int ST, int MB, int WPIPE_WARPS, bool REC, bool DEP, bool STREAM = true>
static void launch_wpipe (const DevTree * T, const DevField * F, const DevParticles * P,
			  const DevStep * S, int n_sm, cudaStream_t st, const DevDeposit * D)
{
  const int n_tiles = (int) ((P->n + 31)/32);
  const size_t smem = (size_t) WPIPE_WARPS*ST*(DIM == 3 ? 8 : 6)*32*sizeof (double);
  static bool configured = false;
  static int per_sm = MB;
  if (!configured) {
    cudaFuncSetAttribute (step_kernel_wpipe<DIM, LA, PR, ST, MB, WPIPE_WARPS, REC, DEP, STREAM>,
			  cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    /* persistent grid: as many CTAs per SM as the instance's registers and staging allow -- MB
       for the 72-register 3D kernels, more for the leaner ones (2D drag: 48 registers -> 10) */
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor (&occ, step_kernel_wpipe<DIM, LA, PR, ST, MB, WPIPE_WARPS, REC, DEP, STREAM>,
						       32*WPIPE_WARPS, smem) == cudaSuccess && occ > MB &&
	!getenv ("GFSB200_WPIPE_FIXED_GRID"))
      per_sm = occ;
    configured = true;
  }
  int grid = n_sm*per_sm;
  if (grid*WPIPE_WARPS > n_tiles) grid = (n_tiles + WPIPE_WARPS - 1)/WPIPE_WARPS;
  DevDeposit none;
  memset (&none, 0, sizeof none);
  static const bool pdl = !(getenv ("GFSB200_PDL") && atoi (getenv ("GFSB200_PDL")) == 0);
  if (pdl) {
    /* let the prologue (mbarrier set-up, the first particle tiles) overlap the tail of the cell pass */
    cudaLaunchConfig_t cfg;
    memset (&cfg, 0, sizeof cfg);
    cfg.gridDim = dim3 (grid); cfg.blockDim = dim3 (32*WPIPE_WARPS);
    cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaLaunchKernelEx (&cfg, step_kernel_wpipe<DIM, LA, PR, ST, MB, WPIPE_WARPS, REC, DEP, STREAM>,
			*T, *F, *P, *S, n_tiles, DEP ? *D : none);
    return;
  }
  step_kernel_wpipe<DIM, LA, PR, ST, MB, WPIPE_WARPS, REC, DEP, STREAM><<<grid, 32*WPIPE_WARPS, smem, st>>>
    (*T, *F, *P, *S, n_tiles, DEP ? *D : none);
}

extern "C" {

/* mode: 0 plain kernel; 2: TMA-staged persistent kernel, 3 CTAs x 256 threads per SM;
 * 3: the same with 6 CTAs x 128 threads -- shorter waits at the per-tile barrier
 * (64-thread tiles and a third stage measured within 1.5 % of it); 7, 9: the warp-pipelined
 * kernel (step_kernel_wpipe); < 0 (default): chosen by tree type.
 * rec: also record cell and force per particle.  D != NULL: the caller wants the two-way deposits
 * of this step fused into the kernel; returns 1 if the launched kernel did that (the warp-pipelined
 * kernel with a compile-time force list), 0 if the caller has to run the deposit pass itself. */
int gfsb200_launch_step (const DevTree * T, const DevField * F, const DevParticles * P,
			 const DevStep * S, int rec, int minb, int mode, int n_sm, cudaStream_t st,
			 const DevDeposit * D)
{
  gfsb200_launch_counter += 1;
  if (P->n <= 0) return 0;
  const int th = 256;
  const unsigned g = grid_for (P->n, th);
  const bool lat = T->lattice_n1 > 0;
  /* force lists with a compile-time specialisation (drag / lift / buoyancy = 1 / 2 / 3) */
  unsigned prog = 0;
  switch (S->forces) {
  case 0x1: case 0x21: case 0x321: case 0x31: case 0x3:
    prog = S->forces;
  }
  if (S->cd_const == S->cd_const) prog = 0;
  /* default (mode < 0): the warp-pipelined kernel for the compile-time force lists (C2: 0.249 vs
     0.285 ms, C3: 0.266 vs 0.280 ms -- profiles/README.md, round 1e); runtime force lists spill
     at its 72 registers and keep the CTA-pipelined one */
  if (mode < 0)
    mode = prog != 0 ? 9 : 3;
  if (mode >= 4 && P->n >= 1024 && (prog != 0 || !rec)) {
    /* warp-private pipeline, 2 stages, 4 warps per CTA; 9: 7 CTAs/SM (72 registers, 28 warps);
       7: 6 CTAs/SM (80 registers).  The force-recording and the deposit-fusing flavours exist for
       the compile-time force lists at 7 CTAs/SM only. */
    const bool fuse = D != NULL && !rec && prog != 0 && mode != 7;
    /* vertex + vorticity tables against the L2 (see STREAM); lattice trees: 48 B per vertex and leaf in 3D */
    const bool big_tables = lat && (double) T->n_vertices*(T->dim == 3 ? 48. : 24.) > 20e6 &&
      !getenv ("GFSB200_STREAM_HINT");
#define WP_ST(D_, LA, PR) do { \
      if (rec) launch_wpipe<D_, LA, PR, 2, 7, 4, true, false> (T, F, P, S, n_sm, st, NULL); \
      else if (fuse) launch_wpipe<D_, LA, PR, 2, 7, 4, false, true> (T, F, P, S, n_sm, st, D); \
      else if (mode == 7) launch_wpipe<D_, LA, PR, 2, 6, 4, false, false> (T, F, P, S, n_sm, st, NULL); \
      else if (LA && big_tables) launch_wpipe<D_, LA, PR, 2, 7, 4, false, false, !LA> (T, F, P, S, n_sm, st, NULL); \
      else launch_wpipe<D_, LA, PR, 2, 7, 4, false, false> (T, F, P, S, n_sm, st, NULL); } while (0)
#define WP_ST0(D_, LA) do { \
      if (mode == 7) launch_wpipe<D_, LA, 0, 2, 6, 4, false, false> (T, F, P, S, n_sm, st, NULL); \
      else launch_wpipe<D_, LA, 0, 2, 7, 4, false, false> (T, F, P, S, n_sm, st, NULL); } while (0)
#define WP_PR(D_, LA) do { switch (prog) { \
    case 0x1: WP_ST (D_, LA, 0x1); break; case 0x21: WP_ST (D_, LA, 0x21); break; \
    case 0x321: WP_ST (D_, LA, 0x321); break; case 0x31: WP_ST (D_, LA, 0x31); break; \
    case 0x3: WP_ST (D_, LA, 0x3); break; default: WP_ST0 (D_, LA); } } while (0)
    if (T->dim == 3) { if (lat) WP_PR (3, true); else WP_PR (3, false); }
    else             { if (lat) WP_PR (2, true); else WP_PR (2, false); }
#undef WP_PR
#undef WP_ST0
#undef WP_ST
    return fuse ? 1 : 0;
  }
  if (rec) prog = 0;
  if (!rec && mode >= 2 && P->n >= 1024) {
#define PIPE_ST(D, LA, PR) do { if (mode == 3) launch_pipe<D, LA, PR, 2, 6, 128> (T, F, P, S, n_sm, st); \
				else launch_pipe<D, LA, PR, 2, 3, 256> (T, F, P, S, n_sm, st); } while (0)
#define PIPE_PR(D, LA) do { switch (prog) { \
    case 0x1: PIPE_ST (D, LA, 0x1); break; case 0x21: PIPE_ST (D, LA, 0x21); break; \
    case 0x321: PIPE_ST (D, LA, 0x321); break; case 0x31: PIPE_ST (D, LA, 0x31); break; \
    case 0x3: PIPE_ST (D, LA, 0x3); break; default: PIPE_ST (D, LA, 0); } } while (0)
    if (T->dim == 3) { if (lat) PIPE_PR (3, true); else PIPE_PR (3, false); }
    else             { if (lat) PIPE_PR (2, true); else PIPE_PR (2, false); }
#undef PIPE_PR
#undef PIPE_ST
    return 0;
  }
#define LAUNCH(D, R, LA, PR, MB) step_kernel<D, R, LA, PR, MB><<<g, th, 0, st>>> (*T, *F, *P, *S)
#define PICK_MB(D, LA, PR) do { if (minb <= 3) LAUNCH (D, false, LA, PR, 3); else LAUNCH (D, false, LA, PR, 4); } while (0)
#define PICK_PR(D, LA) do { switch (prog) { \
    case 0x1: PICK_MB (D, LA, 0x1); break; case 0x21: PICK_MB (D, LA, 0x21); break; \
    case 0x321: PICK_MB (D, LA, 0x321); break; case 0x31: PICK_MB (D, LA, 0x31); break; \
    case 0x3: PICK_MB (D, LA, 0x3); break; default: PICK_MB (D, LA, 0); } } while (0)
#define PICK(D) do { if (rec) { if (lat) LAUNCH (D, true, true, 0, 3); else LAUNCH (D, true, false, 0, 3); } \
		     else     { if (lat) PICK_PR (D, true); else PICK_PR (D, false); } } while (0)
  if (T->dim == 3) PICK (3); else PICK (2);
#undef PICK
#undef PICK_PR
#undef PICK_MB
#undef LAUNCH
  return 0;
}

void gfsb200_launch_advect (const DevTree * T, const DevField * F, const DevParticles * P,
			    double dt, int rec_cell, cudaStream_t st)
{
  gfsb200_launch_counter += 1;
  if (P->n <= 0) return;
  const int th = 256;
  const unsigned g = grid_for (P->n, th);
  if (T->dim == 3) {
    if (rec_cell) advect_kernel<3, true><<<g, th, 0, st>>> (*T, *F, *P, dt);
    else advect_kernel<3, false><<<g, th, 0, st>>> (*T, *F, *P, dt);
  }
  else {
    if (rec_cell) advect_kernel<2, true><<<g, th, 0, st>>> (*T, *F, *P, dt);
    else advect_kernel<2, false><<<g, th, 0, st>>> (*T, *F, *P, dt);
  }
}

void gfsb200_launch_locate (const DevTree * T, int64_t n, const double * x, const double * y,
			    const double * z, int32_t * cell, cudaStream_t st)
{
  gfsb200_launch_counter += 1;
  if (n <= 0) return;
  if (T->dim == 3) locate_kernel<3><<<grid_for (n, 256), 256, 0, st>>> (*T, n, x, y, z, cell);
  else locate_kernel<2><<<grid_for (n, 256), 256, 0, st>>> (*T, n, x, y, z, cell);
}

void gfsb200_launch_interpolate (const DevTree * T, const DevField * F, int64_t n,
				 const double * x, const double * y, const double * z,
				 double * u, double * v, double * w, cudaStream_t st)
{
  gfsb200_launch_counter += 1;
  if (n <= 0) return;
  if (T->dim == 3)
    interpolate_kernel<3><<<grid_for (n, 256), 256, 0, st>>> (*T, *F, n, x, y, z, u, v, w);
  else
    interpolate_kernel<2><<<grid_for (n, 256), 256, 0, st>>> (*T, *F, n, x, y, z, u, v, w);
}

void gfsb200_launch_corner_values (const DevTree * T, const DevField * F, int comp, int64_t n,
				   const int32_t * cells, double * out, cudaStream_t st)
{
  gfsb200_launch_counter += 1;
  if (n <= 0) return;
  const int nc = 1 << T->dim;
  if (T->dim == 3)
    corner_values_kernel<3><<<grid_for (n*nc, 256), 256, 0, st>>> (*T, *F, comp, n, cells, out);
  else
    corner_values_kernel<2><<<grid_for (n*nc, 256), 256, 0, st>>> (*T, *F, comp, n, cells, out);
}

/* what: bit 0 = void fraction, bit 1 = force components.  idx != NULL: only the n_idx particles
 * idx[k] with keep[idx[k]] != 0 */
void gfsb200_launch_deposit (const DevTree * T, const DevField * F, const DevParticles * P,
			     const DevStep * S, int what, const DevDeposit * D, int n_idx,
			     const int32_t * idx, const uint8_t * keep, cudaStream_t st)
{
  gfsb200_launch_counter += 1;
  const int64_t n = idx ? n_idx : P->n;
  if (n <= 0 || !(what & 3)) return;
  const unsigned g = grid_for (n, 256);
  const bool lat = T->lattice_n1 > 0;
  unsigned prog = 0;
  if (what & 2)
    switch (S->forces) {
    case 0x1: case 0x21: case 0x321: case 0x31: case 0x3:
      prog = S->forces;
    }
  if (S->cd_const == S->cd_const) prog = 0;
  if (idx) {
    /* the rare fix-up pass after a fused step + deposit: runtime force list, both deposits */
    if (T->dim == 3) {
      if (lat) deposit_kernel<3, true, 0, true, true, true><<<g, 256, 0, st>>> (*T, *F, *P, *S, *D, n_idx, idx, keep);
      else deposit_kernel<3, false, 0, true, true, true><<<g, 256, 0, st>>> (*T, *F, *P, *S, *D, n_idx, idx, keep);
    }
    else {
      if (lat) deposit_kernel<2, true, 0, true, true, true><<<g, 256, 0, st>>> (*T, *F, *P, *S, *D, n_idx, idx, keep);
      else deposit_kernel<2, false, 0, true, true, true><<<g, 256, 0, st>>> (*T, *F, *P, *S, *D, n_idx, idx, keep);
    }
    return;
  }
#define DEP(D_, LA, PR, V, FO) deposit_kernel<D_, LA, PR, V, FO, false><<<g, 256, 0, st>>> (*T, *F, *P, *S, *D, 0, NULL, NULL)
#define DEP_W(D_, LA, PR) do { if (what == 1) DEP (D_, LA, 0, true, false); else if (what == 2) DEP (D_, LA, PR, false, true); \
			      else DEP (D_, LA, PR, true, true); } while (0)
#define DEP_PR(D_, LA) do { switch (prog) { \
    case 0x1: DEP_W (D_, LA, 0x1); break; case 0x21: DEP_W (D_, LA, 0x21); break; \
    case 0x321: DEP_W (D_, LA, 0x321); break; case 0x31: DEP_W (D_, LA, 0x31); break; \
    case 0x3: DEP_W (D_, LA, 0x3); break; default: DEP_W (D_, LA, 0); } } while (0)
  if (T->dim == 3) { if (lat) DEP_PR (3, true); else DEP_PR (3, false); }
  else             { if (lat) DEP_PR (2, true); else DEP_PR (2, false); }
#undef DEP_PR
#undef DEP_W
#undef DEP
}

/* GfsSourceParticulate with a smoothing kernel: on-fluid forces into P.fx/fy/fz,
 * then the two conditional traversals per particle.  norm (may be NULL):
 * [2][n] per-particle correction and volume. */
void gfsb200_launch_deposit_smoothed (const DevTree * T, const DevField * F, const DevParticles * P,
				      const DevStep * S, double rkernel, const gfsb200_kernel * K,
				      double * f0, double * f1, double * f2, double * norm,
				      cudaStream_t st)
{
  gfsb200_launch_counter += 2;
  if (P->n <= 0) return;
  const bool lat = T->lattice_n1 > 0;
  const unsigned g = grid_for (P->n, 256);
  if (T->dim == 3) {
    if (lat) onfluid_force_kernel<3, true><<<g, 256, 0, st>>> (*T, *F, *P, *S);
    else     onfluid_force_kernel<3, false><<<g, 256, 0, st>>> (*T, *F, *P, *S);
    smoothed_deposit_kernel<3><<<grid_for (P->n, 128), 128, 0, st>>> (*T, *F, *P, S->rho, rkernel, *K, f0, f1, f2, norm);
  }
  else {
    if (lat) onfluid_force_kernel<2, true><<<g, 256, 0, st>>> (*T, *F, *P, *S);
    else     onfluid_force_kernel<2, false><<<g, 256, 0, st>>> (*T, *F, *P, *S);
    smoothed_deposit_kernel<2><<<grid_for (P->n, 128), 128, 0, st>>> (*T, *F, *P, S->rho, rkernel, *K, f0, f1, f2, norm);
  }
}

void gfsb200_launch_particle_bc (const DevTree * T, const DevParticles * P, int n_esc,
				 const int32_t * esc_idx, const double * esc_old, uint8_t * keep,
				 int * counters, cudaStream_t st)
{
  gfsb200_launch_counter += 1;
  if (n_esc <= 0) return;
  if (T->dim == 3)
    particle_bc_kernel<3><<<(n_esc + 127)/128, 128, 0, st>>> (*T, *P, n_esc, esc_idx, esc_old, keep, counters);
  else
    particle_bc_kernel<2><<<(n_esc + 127)/128, 128, 0, st>>> (*T, *P, n_esc, esc_idx, esc_old, keep, counters);
}

void gfsb200_launch_gather (int64_t n, const int32_t * perm, int ncols, const double * const * src,
			    double * const * dst, const uint32_t * id_src, uint32_t * id_dst,
			    cudaStream_t st)
{
  gfsb200_launch_counter += 1;
  if (n <= 0) return;
  gather_kernel<<<grid_for (n, 256), 256, 0, st>>> (n, perm, ncols, src, dst, id_src, id_dst);
}

/* dst_k[i] = src_k[perm[i]] for three plain arrays (the recorded forces) */
__global__ void __launch_bounds__(256)
gather3_kernel (int64_t n, const int32_t * __restrict__ perm,
		const double * __restrict__ s0, const double * __restrict__ s1, const double * __restrict__ s2,
		double * __restrict__ d0, double * __restrict__ d1, double * __restrict__ d2)
{
  const int64_t i = (int64_t) blockIdx.x*blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int32_t j = perm[i];
  d0[i] = s0[j]; d1[i] = s1[j]; d2[i] = s2[j];
}

void gfsb200_launch_gather3 (int64_t n, const int32_t * perm, const double * s0, const double * s1,
			     const double * s2, double * d0, double * d1, double * d2, cudaStream_t st)
{
  gfsb200_launch_counter += 1;
  if (n <= 0) return;
  gather3_kernel<<<grid_for (n, 256), 256, 0, st>>> (n, perm, s0, s1, s2, d0, d1, d2);
}

void gfsb200_launch_iota (int64_t n, int32_t * a, cudaStream_t st)
{
  gfsb200_launch_counter += 1;
  if (n > 0) iota_kernel<<<grid_for (n, 256), 256, 0, st>>> (n, a);
}

void gfsb200_launch_iota_u32 (int64_t n, uint32_t * a, uint32_t base, cudaStream_t st)
{
  gfsb200_launch_counter += 1;
  if (n > 0) iota_u32_kernel<<<grid_for (n, 256), 256, 0, st>>> (n, a, base);
}

void gfsb200_launch_sort_keys (int64_t n, const int32_t * cell, uint32_t * key,
			       uint32_t outside_key, cudaStream_t st)
{
  gfsb200_launch_counter += 1;
  if (n > 0) sort_keys_kernel<<<grid_for (n, 256), 256, 0, st>>> (n, cell, key, outside_key);
}

void gfsb200_launch_owner_keys (int64_t n, const uint32_t * cell_key, uint32_t n_cells, const uint8_t * owner_of,
				uint32_t * key, uint32_t outside_key, cudaStream_t st)
{
  gfsb200_launch_counter += 1;
  if (n > 0) owner_keys_kernel<<<grid_for (n, 256), 256, 0, st>>> (n, cell_key, n_cells, owner_of, key, outside_key);
}

void gfsb200_launch_inside_flags (int64_t n, const int32_t * cell, uint8_t * flag, cudaStream_t st)
{
  gfsb200_launch_counter += 1;
  if (n > 0) inside_flag_kernel<<<grid_for (n, 256), 256, 0, st>>> (n, cell, flag);
}

/* keep[i] = 0 for every particle that is still flagged but lies outside the domain (cell < 0);
 * *count is incremented once per particle cleared */
__global__ void __launch_bounds__(256)
outside_clear_kernel (int64_t n, const int32_t * __restrict__ cell, uint8_t * __restrict__ keep,
		      int * __restrict__ count)
{
  const int64_t i = (int64_t) blockIdx.x*blockDim.x + threadIdx.x;
  if (i < n && cell[i] < 0 && keep[i]) {
    keep[i] = 0;
    atomicAdd (count, 1);
  }
}

void gfsb200_launch_outside_clear (int64_t n, const int32_t * cell, uint8_t * keep, int * count, cudaStream_t st)
{
  gfsb200_launch_counter += 1;
  if (n > 0) outside_clear_kernel<<<grid_for (n, 256), 256, 0, st>>> (n, cell, keep, count);
}

/* cub wrappers: pass tmp = NULL to query *tmp_bytes */
cudaError_t gfsb200_cub_sort_pairs (void * tmp, size_t * tmp_bytes, const uint32_t * keys_in,
				    uint32_t * keys_out, const int32_t * vals_in, int32_t * vals_out,
				    int64_t n, int end_bit, cudaStream_t st)
{
  return cub::DeviceRadixSort::SortPairs (tmp, *tmp_bytes, keys_in, keys_out, vals_in, vals_out,
					  (int) n, 0, end_bit, st);
}

cudaError_t gfsb200_cub_select_flagged (void * tmp, size_t * tmp_bytes, const int32_t * in,
					const uint8_t * flags, int32_t * out, int32_t * n_selected,
					int64_t n, cudaStream_t st)
{
  return cub::DeviceSelect::Flagged (tmp, *tmp_bytes, in, flags, out, n_selected, (int) n, st);
}

} // extern "C"

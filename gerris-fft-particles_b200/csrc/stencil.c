/* stencil.c -- corner-interpolator stencils on the flat tree, deduplicated
 * into a per-vertex table.
 *
 * gfs_interpolate (src/fluid.c:2697-2710) rebuilds, for every call, the
 * inverse-distance stencil of each of the 2^dim corners of the containing
 * cell (gfs_cell_corner_interpolator, src/fluid.c:3015-3069).  Those stencils
 * depend on the tree topology only -- not on the particle, not on the field --
 * so here they are built once per adapt, on the host, by the same
 * depth-first walk over the cells sharing the vertex (including T-junction
 * averaging and the domain-corner rule), and stored as (cell, weight) lists.
 * Leaves sharing a vertex share one list when the lists agree as sets, which
 * shrinks the table 2^dim-fold on regular regions; the device then only
 * evaluates sum(w_i * v_i) once per vertex per field update.
 */
#include <stdlib.h>
#include <string.h>
#include <math.h>
#ifdef _OPENMP
# include <omp.h>
#endif
#include "gfsb200_internal.h"

typedef struct {
  int n;
  int32_t c[GFSB200_MAX_STENCIL];
  double w[GFSB200_MAX_STENCIL];
} interp_t;

/* corner k -> its dim face directions, in the order of the reference's
 * corner[] table (src/fluid.c:2588-2606): x, y, z */
static const int corner_dir[2][8][3] = {
  { {1,3,0}, {0,3,0}, {0,2,0}, {1,2,0} },
  { {1,3,4}, {0,3,4}, {0,2,4}, {1,2,4}, {1,3,5}, {0,3,5}, {0,2,5}, {1,2,5} }
};

/* Depth-first order in which the 2^dim cells around a vertex are visited
 * (do_path, src/fluid.c:2925-2981).  Slot s holds the cell reached from slot 0
 * by crossing the faces whose bits are set in s; from slot s, try j crosses
 * face walk[s][j][0] first and passes the remaining positions in the listed
 * order.  Entries are positions in the corner's direction array, and the sign
 * of each direction follows from the slot (set bit = already crossed = go
 * back), so only the permutations are tabulated. */
static const unsigned char walk3[8][3][3] = {
  { {0,1,2}, {1,0,2}, {2,0,1} },
  { {1,2,0}, {2,1,0}, {0,1,2} },
  { {0,1,2}, {2,1,0}, {1,2,0} },
  { {2,0,1}, {1,0,2}, {0,1,2} },
  { {1,0,2}, {0,1,2}, {2,0,1} },
  { {1,0,2}, {0,1,2}, {2,1,0} },
  { {0,1,2}, {1,0,2}, {2,0,1} },
  { {0,1,2}, {1,0,2}, {2,0,1} }
};
static const unsigned char walk2[4][2][2] = {
  { {0,1}, {1,0} },
  { {1,0}, {0,1} },
  { {0,1}, {1,0} },
  { {0,1}, {1,0} }
};

static inline int child_positive (int n, int axis)
{
  return axis == 0 ? (n & 1) : !((n >> axis) & 1);
}

/* ftt_cell_child_corner, src/ftt.h:366-422 */
static int32_t child_corner (const gfsb200_tree * t, int32_t cell, const int * d)
{
  int n = 0;
  for (int l = 0; l < t->dim; l++) {
    int axis = d[l] >> 1, positive = !(d[l] & 1);
    if (axis == 0) n |= positive ? 1 : 0;
    else n |= positive ? 0 : (1 << axis);
  }
  int32_t c = t->child0[cell] + n;
  return (t->flags[c] & GFSB200_CELL_DESTROYED) ? -1 : c;
}

/* ftt_cell_neighbor_is_brother, src/ftt.h:620-638 */
static int neighbor_is_brother (const gfsb200_tree * t, int32_t cell, int d)
{
  int32_t p = t->parent[cell];
  if (p < 0) return 0;
  int n = cell - t->child0[p];
  return child_positive (n, d >> 1) != !(d & 1);
}

/* cell_corner_neighbor, src/fluid.c:2813-2846 (max_level = -1) */
static int32_t corner_neighbor (const gfsb200_tree * t, int32_t cell, const int * d, int * t_junction)
{
  int32_t nb = t->neighbor[(int64_t) cell*t->ndir + d[0]];
  if (nb < 0)
    return -1;
  if (t->level[nb] < t->level[cell]) {
    /* shallower neighbour */
    if (child_corner (t, t->parent[cell], d) != cell)
      *t_junction = 1;
    return nb;
  }
  if (t->child0[nb] < 0)
    return nb;
  /* deeper: the neighbour's child touching the vertex */
  int d1[3];
  d1[0] = d[0] ^ 1;
  for (int l = 1; l < t->dim; l++) d1[l] = d[l];
  int32_t n = child_corner (t, nb, d1);
  return n >= 0 ? n : nb;
}

/* interpolator_merge / interpolator_scale, src/fluid.c:2848-2877 */
static void merge (interp_t * a, const interp_t * b)
{
  for (int i = 0; i < b->n; i++) {
    int j;
    for (j = 0; j < a->n && b->c[i] != a->c[j]; j++)
      ;
    if (j < a->n)
      a->w[j] += b->w[i];
    else if (j < GFSB200_MAX_STENCIL) {
      a->c[j] = b->c[i];
      a->w[j] = b->w[i];
      a->n++;
    }
  }
}

static void scale (interp_t * a, double b)
{
  for (int i = 0; i < a->n; i++)
    a->w[i] *= b;
}

static void corner_interp (const gfsb200_tree * t, int32_t cell, const int * d, interp_t * inter);

/* t_junction_interpolator, src/fluid.c:2879-2923 */
static void t_junction (const gfsb200_tree * t, int32_t cell, const int * d, int32_t n,
			interp_t * inter)
{
  int d1[3];
  interp_t a;
  d1[0] = d[0] ^ 1;
  if (t->dim == 2) {
    d1[1] = d[1];
    corner_interp (t, n, d1, inter);
    d1[1] = d[1] ^ 1;
    corner_interp (t, n, d1, &a);
    merge (inter, &a);
    scale (inter, 0.5);
    return;
  }
  d1[1] = d[1]; d1[2] = d[2];
  corner_interp (t, n, d1, inter);
  if (neighbor_is_brother (t, cell, d[1])) {
    d1[1] = d[1] ^ 1;
    corner_interp (t, n, d1, &a);
    merge (inter, &a);
    if (neighbor_is_brother (t, cell, d[2])) {
      d1[2] = d[2] ^ 1;
      corner_interp (t, n, d1, &a);
      merge (inter, &a);
      d1[1] = d[1];
      corner_interp (t, n, d1, &a);
      merge (inter, &a);
      scale (inter, 0.25);
    }
    else
      scale (inter, 0.5);
  }
  else {
    d1[2] = d[2] ^ 1;
    corner_interp (t, n, d1, &a);
    merge (inter, &a);
    scale (inter, 0.5);
  }
}

/* do_path, src/fluid.c:2925-2981 */
static int do_path (const gfsb200_tree * t, int32_t cell, int slot, int32_t * n, const int * d,
		    interp_t * inter)
{
  for (int j = 0; j < t->dim; j++) {
    const unsigned char * w = t->dim == 3 ? walk3[slot][j] : walk2[slot][j];
    int k = slot ^ (1 << w[0]);
    if (n[k] < 0) {
      int tj = 0, d1[3];
      for (int l = 0; l < t->dim; l++)
	d1[l] = (slot >> w[l]) & 1 ? d[w[l]] ^ 1 : d[w[l]];
      n[k] = corner_neighbor (t, cell, d1, &tj);
      if (tj) {
	t_junction (t, cell, d1, n[k], inter);
	return 1;
      }
      if (n[k] >= 0 && do_path (t, n[k], k, n, d, inter))
	return 1;
    }
  }
  return 0;
}

/* gfs_cell_corner_interpolator, src/fluid.c:3015-3069, for non-centred
 * variables (U,V,W: src/variable.c:114); mixed cells as set by gfsb200_tree_set_solid */
static void corner_interp (const gfsb200_tree * t, int32_t cell, const int * d, interp_t * inter)
{
  int32_t n[8], c;
  int ncells = t->nchild;
  while (t->child0[cell] >= 0 && (c = child_corner (t, cell, d)) >= 0)
    cell = c;
  n[0] = cell;
  for (int i = 1; i < ncells; i++)
    n[i] = -1;
  if (do_path (t, cell, 0, n, d, inter))
    return;

  double w = 0.;
  int boundaries = 0;
  const double diag = t->dim == 3 ? 0.866025403785 : 0.707106781185;  /* distance(), :2983-2992 */
  /* ftt_corner_pos (cell, d), src/ftt.c:379-429: only needed next to mixed cells */
  double corner[3] = { 0., 0., 0. };
  if (t->solid_cm) {
    const double size = ldexp (1., -t->level[cell]);
    double rel[3] = { 0., 0., 0. };
    for (int l = 0; l < t->dim; l++)
      rel[d[l] >> 1] += d[l] & 1 ? -0.5 : 0.5;
    for (int l = 0; l < t->dim; l++)
      corner[l] = t->pos[3*cell + l] + size*rel[l];
  }
  inter->n = 0;
  for (int i = 0; i < ncells; i++)
    if (n[i] >= 0) {
      double size = ldexp (1., -t->level[n[i]]);
      double dist = size*diag;
      if (t->solid_cm && t->solid_cm[3*n[i]] == t->solid_cm[3*n[i]]) {
	/* U,V,W are not "centered" variables: a mixed cell is weighted by the distance of
	   the corner to the centre of mass of its fluid part (:2993-3002) */
	const double * cm = t->solid_cm + 3*n[i];
	dist = t->dim == 3 ?
	  sqrt ((cm[0] - corner[0])*(cm[0] - corner[0]) + (cm[1] - corner[1])*(cm[1] - corner[1]) +
		(cm[2] - corner[2])*(cm[2] - corner[2])) :
	  sqrt ((cm[0] - corner[0])*(cm[0] - corner[0]) + (cm[1] - corner[1])*(cm[1] - corner[1]));
      }
      double a = 1./(dist + 1e-12);
      inter->c[inter->n] = n[i];
      inter->w[inter->n++] = a;
      w += a;
      if (t->flags[n[i]] & GFSB200_CELL_BOUNDARY)
	boundaries++;
    }
  /* corners of the domain: drop the central cell (:3056-3065) */
  if (inter->n == t->dim + 1 && boundaries == t->dim) {
    w -= inter->w[0];
    for (int i = 0; i < inter->n - 1; i++) {
      inter->c[i] = inter->c[i + 1];
      inter->w[i] = inter->w[i + 1];
    }
    inter->n--;
  }
  scale (inter, 1./w);
}

int gfsb200_tree_corner_interpolator (const gfsb200_tree * t, int cell, int k,
				      int32_t * cells, double * w)
{
  if (!t || !t->finalized)
    return gfsb200_fail (GFSB200_ERR_STATE, "corner_interpolator: tree not finalized");
  if (cell < 0 || cell >= t->n_cells || k < 0 || k >= t->nchild ||
      (t->flags[cell] & GFSB200_CELL_DESTROYED))
    return gfsb200_fail (GFSB200_ERR_ARG, "corner_interpolator: bad cell/corner");
  interp_t inter;
  corner_interp (t, cell, corner_dir[t->dim - 2][k], &inter);
  for (int i = 0; i < inter.n; i++) {
    cells[i] = inter.c[i];
    w[i] = inter.w[i];
  }
  return inter.n;
}

/* ------------------------------------------------------------------ */
/* vertex deduplication                                                 */

static inline uint64_t mix64 (uint64_t x)
{
  x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ULL;
  x ^= x >> 27; x *= 0x94d049bb133111ebULL;
  x ^= x >> 31;
  return x;
}

/* order-independent signature of a stencil: hash of the (cell, weight) pairs
 * sorted by cell */
static void signature (const interp_t * in, uint64_t * h1, uint64_t * h2)
{
  int idx[GFSB200_MAX_STENCIL];
  for (int i = 0; i < in->n; i++) {
    int j = i;
    while (j > 0 && in->c[idx[j - 1]] > in->c[i]) {
      idx[j] = idx[j - 1];
      j--;
    }
    idx[j] = i;
  }
  uint64_t a = 0x9e3779b97f4a7c15ULL + (uint64_t) in->n, b = 0xc2b2ae3d27d4eb4fULL;
  for (int i = 0; i < in->n; i++) {
    uint64_t wb;
    memcpy (&wb, &in->w[idx[i]], 8);
    a = mix64 (a ^ (uint64_t) (uint32_t) in->c[idx[i]]);
    a = mix64 (a ^ wb);
    b = mix64 (b + wb*0x2545f4914f6cdd1dULL + (uint64_t) (uint32_t) in->c[idx[i]]);
  }
  *h1 = a; *h2 = b;
}

typedef struct {
  int64_t k[3];
  uint64_t h1, h2;
  int32_t vid;       /* -1 = empty */
  int32_t rep;       /* canonical cell*nchild + corner */
} vslot_t;

typedef struct {
  vslot_t * s;
  uint64_t mask;
  int64_t used;
} vtable_t;

static int vtable_init (vtable_t * v, uint64_t cap)
{
  uint64_t c = 1024;
  while (c < cap) c <<= 1;
  v->s = malloc (c*sizeof (vslot_t));
  if (!v->s) return -1;
  for (uint64_t i = 0; i < c; i++) v->s[i].vid = -1;
  v->mask = c - 1;
  v->used = 0;
  return 0;
}

static inline uint64_t vhash (const int64_t k[3], uint64_t h1)
{
  return mix64 ((uint64_t) k[0]*0x9e3779b97f4a7c15ULL ^ mix64 ((uint64_t) k[1] + 0x632be59bd9b4e019ULL) ^
		mix64 ((uint64_t) k[2]*0xd6e8feb86659fd93ULL + 7) ^ h1);
}

static vslot_t * vtable_find (vtable_t * v, const int64_t k[3], uint64_t h1, uint64_t h2)
{
  uint64_t i = vhash (k, h1) & v->mask;
  for (;;) {
    vslot_t * s = &v->s[i];
    if (s->vid < 0 ||
	(s->k[0] == k[0] && s->k[1] == k[1] && s->k[2] == k[2] && s->h1 == h1 && s->h2 == h2))
      return s;
    i = (i + 1) & v->mask;
  }
}

static int vtable_grow (vtable_t * v)
{
  vtable_t n;
  if (vtable_init (&n, (v->mask + 1)*2))
    return -1;
  for (uint64_t i = 0; i <= v->mask; i++)
    if (v->s[i].vid >= 0) {
      vslot_t * s = vtable_find (&n, v->s[i].k, v->s[i].h1, v->s[i].h2);
      *s = v->s[i];
    }
  n.used = v->used;
  free (v->s);
  *v = n;
  return 0;
}

/* A one-box tree whose leaves all sit at one level L has exactly one vertex per
 * point of the (2^L + 1)^dim lattice.  Renumber them in row-major lattice order,
 *   id = (k*(N + 1) + j)*(N + 1) + i,   N = 2^L,
 * so that the device can compute the ids of a leaf's corners from its column
 * indices instead of loading leaf_vtx.  Leaves lattice_level = -1 otherwise. */
static void lattice_renumber (gfsb200_tree * t, const int64_t * vkey)
{
  const int dim = t->dim, nc = t->nchild;
  const int L = t->max_level - t->root_level;
  if (t->n_box_roots != 1 || t->complete_level != t->max_level || L < 1 || L > (dim == 3 ? 10 : 15))
    return;
  const int64_t N1 = ((int64_t) 1 << L) + 1;
  int64_t want = N1*N1*(dim == 3 ? N1 : 1);
  if (want != t->n_vertices)
    return;
  const int nv = t->n_vertices;
  const double lattice = ldexp (1., GFSB200_MAX_LEVEL + 2);
  const double size = ldexp (1., -t->root_level);
  const int64_t step = (int64_t) llround (ldexp (size, -L)*lattice);
  int64_t lo[3];
  for (int a = 0; a < 3; a++)
    lo[a] = a < dim ? (int64_t) llround ((t->pos[a] - size/2.)*lattice) : 0;
  int32_t * newid = malloc ((size_t) nv*sizeof (int32_t));
  char * seen = calloc ((size_t) nv, 1);
  int ok = newid && seen;
  for (int v = 0; v < nv && ok; v++) {
    int64_t id = 0;
    for (int a = dim - 1; a >= 0; a--) {
      int64_t d = vkey[3*v + a] - lo[a];
      if (d < 0 || d % step || d/step >= N1) { ok = 0; break; }
      id = id*N1 + d/step;
    }
    if (ok && seen[id]) ok = 0;
    if (ok) { seen[id] = 1; newid[v] = (int32_t) id; }
  }
  free (seen);
  if (ok) {
    int32_t * off = malloc ((size_t) (nv + 1)*sizeof (int32_t));
    int32_t * cell = malloc ((size_t) (t->vtx_off[nv] ? t->vtx_off[nv] : 1)*sizeof (int32_t));
    double * w = malloc ((size_t) (t->vtx_off[nv] ? t->vtx_off[nv] : 1)*sizeof (double));
    if (off && cell && w) {
      for (int v = 0; v < nv; v++)
	off[newid[v] + 1] = t->vtx_off[v + 1] - t->vtx_off[v];
      off[0] = 0;
      for (int v = 0; v < nv; v++)
	off[v + 1] += off[v];
      for (int v = 0; v < nv; v++) {
	int32_t o = off[newid[v]], b = t->vtx_off[v], n = t->vtx_off[v + 1] - b;
	memcpy (cell + o, t->vtx_cell + b, (size_t) n*sizeof (int32_t));
	memcpy (w + o, t->vtx_w + b, (size_t) n*sizeof (double));
      }
      for (int64_t e = 0; e < (int64_t) t->n_cells*nc; e++)
	if (t->leaf_vtx[e] >= 0)
	  t->leaf_vtx[e] = newid[t->leaf_vtx[e]];
      free (t->vtx_off); free (t->vtx_cell); free (t->vtx_w);
      t->vtx_off = off; t->vtx_cell = cell; t->vtx_w = w;
      t->lattice_level = t->max_level;
    }
    else {
      free (off); free (cell); free (w);
    }
  }
  free (newid);
}

/* Uniform one-box trees: every leaf sits at the complete level, so the vertex
 * set is the (2^L + 1)^dim lattice and the leaf <-> vertex relation is
 * arithmetic.  Each vertex gets the interpolator of its canonical perspective
 * -- the lowest-indexed leaf that has it as a corner, lowest corner first,
 * exactly what the general first-seen de-duplication below would pick -- so
 * only one interpolator per vertex is built (8x fewer than the general path,
 * no hashing).  All leaves around a lattice vertex have the same size and no
 * T-junction can occur, hence every perspective yields the same cell set. */
static int lattice_stencils (gfsb200_tree * t)
{
  const int dim = t->dim, nc = t->nchild;
  const int L = t->max_level - t->root_level;
  const int N = 1 << L, N1 = N + 1;
  const int32_t top = t->level_start[L];
  const int64_t nv64 = (int64_t) N1*N1*(dim == 3 ? N1 : 1);
  if (nv64 > INT32_MAX/8)
    return 0;
  const int32_t nv = (int32_t) nv64;
  const int (* cd)[3] = corner_dir[dim - 2];
  const int32_t n = t->n_cells;

  /* column indices of every leaf, from its exact centre */
  const double size = ldexp (1., -t->root_level), h = ldexp (size, -L);
  int32_t * rep = malloc ((size_t) nv*2*sizeof (int32_t));     /* canonical (leaf, corner) */
  t->leaf_vtx = malloc ((size_t) n*nc*sizeof (int32_t));
  if (!rep || !t->leaf_vtx) { free (rep); return -1; }
  for (int32_t v = 0; v < nv; v++) rep[2*v] = INT32_MAX;
  for (int64_t e = 0; e < (int64_t) n*nc; e++) t->leaf_vtx[e] = -1;
  const int64_t n_leaves = (int64_t) 1 << (dim*L);
  for (int64_t j = 0; j < n_leaves; j++) {                      /* ascending leaf index */
    const int32_t i = top + (int32_t) j;
    int col[3] = { 0, 0, 0 };
    for (int a = 0; a < dim; a++)
      col[a] = (int) floor ((t->pos[3*i + a] - (t->pos[a] - size/2.))/h);
    for (int k = 0; k < nc; k++) {
      int vi[3] = { col[0], col[1], col[2] };
      for (int l = 0; l < dim; l++) {
	const int d = cd[k][l];
	if (!(d & 1)) vi[d >> 1]++;
      }
      const int32_t v = dim == 3 ? (vi[2]*N1 + vi[1])*N1 + vi[0] : vi[1]*N1 + vi[0];
      t->leaf_vtx[(int64_t) i*nc + k] = v;
      if (rep[2*v] == INT32_MAX) { rep[2*v] = i; rep[2*v + 1] = k; }
    }
  }
  t->vtx_off = malloc ((size_t) (nv + 1)*sizeof (int32_t));
  if (!t->vtx_off) { free (rep); return -1; }
  t->vtx_off[0] = 0;
#pragma omp parallel for schedule(dynamic, 4096)
  for (int32_t v = 0; v < nv; v++) {
    interp_t inter;
    corner_interp (t, rep[2*v], cd[rep[2*v + 1]], &inter);
    t->vtx_off[v + 1] = inter.n;
  }
  int64_t total = 0;
  for (int32_t v = 0; v < nv; v++) {
    const int32_t c = t->vtx_off[v + 1];
    t->vtx_off[v] = (int32_t) total;
    total += c;
  }
  t->vtx_off[nv] = (int32_t) total;
  t->vtx_cell = malloc ((size_t) (total ? total : 1)*sizeof (int32_t));
  t->vtx_w = malloc ((size_t) (total ? total : 1)*sizeof (double));
  if (!t->vtx_cell || !t->vtx_w) { free (rep); return -1; }
#pragma omp parallel for schedule(dynamic, 4096)
  for (int32_t v = 0; v < nv; v++) {
    interp_t inter;
    corner_interp (t, rep[2*v], cd[rep[2*v + 1]], &inter);
    const int32_t o = t->vtx_off[v];
    for (int q = 0; q < inter.n; q++) {
      t->vtx_cell[o + q] = inter.c[q];
      t->vtx_w[o + q] = inter.w[q];
    }
  }
  free (rep);
  t->n_vertices = nv;
  t->lattice_level = t->max_level;
  return 1;
}

int gfsb200_tree_build_stencils (gfsb200_tree * t)
{
  if (!t || !t->finalized)
    return gfsb200_fail (GFSB200_ERR_STATE, "build_stencils: tree not finalized");
  free (t->vtx_off); free (t->vtx_cell); free (t->vtx_w); free (t->leaf_vtx);
  t->vtx_off = NULL; t->vtx_cell = NULL; t->vtx_w = NULL; t->leaf_vtx = NULL;
  t->n_vertices = 0;
  t->lattice_level = -1;

  {
    const int L = t->max_level - t->root_level;
    if (!t->solid_cm && t->n_box_roots == 1 && t->complete_level == t->max_level && L >= 1 &&
	L <= (t->dim == 3 ? 10 : 15) && !getenv ("GFSB200_GENERAL_STENCILS")) {
      int r = lattice_stencils (t);
      if (r > 0) return GFSB200_OK;
      free (t->vtx_off); free (t->vtx_cell); free (t->vtx_w); free (t->leaf_vtx);
      t->vtx_off = NULL; t->vtx_cell = NULL; t->vtx_w = NULL; t->leaf_vtx = NULL;
      if (r < 0) return gfsb200_fail (GFSB200_ERR_NOMEM, "build_stencils: out of memory");
    }
  }

  const int nc = t->nchild;
  const int32_t n = t->n_cells;
  const int (* cd)[3] = corner_dir[t->dim - 2];
  uint64_t * sig = malloc ((size_t) n*nc*2*sizeof (uint64_t));
  t->leaf_vtx = malloc ((size_t) n*nc*sizeof (int32_t));
  if (!sig || !t->leaf_vtx) {
    free (sig);
    return gfsb200_fail (GFSB200_ERR_NOMEM, "build_stencils: out of memory");
  }

  /* pass A: signature of every (box leaf, corner) stencil */
#pragma omp parallel for schedule(dynamic, 4096)
  for (int32_t i = 0; i < n; i++) {
    int box_leaf = (t->flags[i] & (GFSB200_CELL_LEAF | GFSB200_CELL_BOUNDARY)) == GFSB200_CELL_LEAF;
    for (int k = 0; k < nc; k++) {
      t->leaf_vtx[(int64_t) i*nc + k] = -1;
      if (box_leaf) {
	interp_t inter;
	corner_interp (t, i, cd[k], &inter);
	signature (&inter, &sig[((int64_t) i*nc + k)*2], &sig[((int64_t) i*nc + k)*2 + 1]);
      }
    }
  }

  /* pass B: group by (vertex position, signature).  The hash table is sharded
     by position: every thread scans all (leaf, corner) pairs, in ascending
     order, but only inserts those whose position hashes to its shard, so the
     first pair it sees for a group is that group's canonical representative
     (the lowest e = leaf*nc + corner).  Vertex ids are then handed out in
     ascending order of the representatives -- the same numbering a serial
     first-seen scan gives, independent of the thread count. */
  if ((int64_t) n*nc > INT32_MAX) {
    free (sig);
    return gfsb200_fail (GFSB200_ERR_UNSUPPORTED, "build_stencils: tree too large");
  }
  const double lattice = ldexp (1., GFSB200_MAX_LEVEL + 2);
  int n_shards = 1;
#ifdef _OPENMP
  n_shards = omp_get_max_threads ();
#endif
  int oom = 0;
#pragma omp parallel num_threads(n_shards)
  {
    int tid = 0;
#ifdef _OPENMP
    tid = omp_get_thread_num ();
#endif
    vtable_t vt;
    if (vtable_init (&vt, (uint64_t) (t->n_leaves*2/n_shards + 1024))) {
#pragma omp atomic write
      oom = 1;
    }
    else {
      for (int32_t i = 0; i < n && !oom; i++) {
	if ((t->flags[i] & (GFSB200_CELL_LEAF | GFSB200_CELL_BOUNDARY)) != GFSB200_CELL_LEAF)
	  continue;
	const double half = ldexp (1., -t->level[i])/2.;
	for (int k = 0; k < nc; k++) {
	  int64_t key[3] = { 0, 0, 0 };
	  for (int l = 0; l < t->dim; l++) {
	    const int d = cd[k][l];
	    const double p = t->pos[3*i + (d >> 1)] + (d & 1 ? -half : half);
	    key[d >> 1] = (int64_t) llround (p*lattice);
	  }
	  if ((int) (vhash (key, 0) % (uint64_t) n_shards) != tid)
	    continue;
	  const int64_t e = (int64_t) i*nc + k;
	  if ((vt.used + 1)*10 > (int64_t) (vt.mask + 1)*6 && vtable_grow (&vt)) {
#pragma omp atomic write
	    oom = 1;
	    break;
	  }
	  vslot_t * sl = vtable_find (&vt, key, sig[2*e], sig[2*e + 1]);
	  if (sl->vid < 0) {
	    sl->k[0] = key[0]; sl->k[1] = key[1]; sl->k[2] = key[2];
	    sl->h1 = sig[2*e]; sl->h2 = sig[2*e + 1];
	    sl->vid = 0;
	    sl->rep = (int32_t) e;
	    vt.used++;
	  }
	  t->leaf_vtx[e] = sl->rep;          /* provisional: the representative pair */
	}
      }
      free (vt.s);
    }
  }
  free (sig);
  if (oom)
    return gfsb200_fail (GFSB200_ERR_NOMEM, "build_stencils: out of memory");
  /* ids in ascending order of the representatives */
  int32_t nv = 0;
  int32_t * idmap = malloc ((size_t) n*nc*sizeof (int32_t));
  if (!idmap)
    return gfsb200_fail (GFSB200_ERR_NOMEM, "build_stencils: out of memory");
  for (int64_t e = 0; e < (int64_t) n*nc; e++)
    if (t->leaf_vtx[e] == e)
      idmap[e] = nv++;
  int32_t * rep = malloc ((size_t) (nv ? nv : 1)*2*sizeof (int32_t));
  int64_t * vkey = malloc ((size_t) (nv ? nv : 1)*3*sizeof (int64_t));   /* lattice position of each vertex */
  if (!rep || !vkey) {
    free (idmap); free (rep); free (vkey);
    return gfsb200_fail (GFSB200_ERR_NOMEM, "build_stencils: out of memory");
  }
#pragma omp parallel for schedule(static)
  for (int64_t e = 0; e < (int64_t) n*nc; e++) {
    const int32_t r = t->leaf_vtx[e];
    if (r < 0)
      continue;
    if (r == e) {
      const int32_t v = idmap[e], i = (int32_t) (e/nc);
      const int k = (int) (e % nc);
      const double half = ldexp (1., -t->level[i])/2.;
      rep[2*v] = i; rep[2*v + 1] = k;
      vkey[3*v] = vkey[3*v + 1] = vkey[3*v + 2] = 0;
      for (int l = 0; l < t->dim; l++) {
	const int d = cd[k][l];
	const double p = t->pos[3*i + (d >> 1)] + (d & 1 ? -half : half);
	vkey[3*v + (d >> 1)] = (int64_t) llround (p*lattice);
      }
    }
  }
#pragma omp parallel for schedule(static)
  for (int64_t e = 0; e < (int64_t) n*nc; e++)
    if (t->leaf_vtx[e] >= 0)
      t->leaf_vtx[e] = idmap[t->leaf_vtx[e]];
  free (idmap);

  /* pass C: CSR of the canonical stencils */
  t->vtx_off = malloc ((size_t) (nv + 1)*sizeof (int32_t));
  if (!t->vtx_off) {
    free (rep); free (vkey);
    return gfsb200_fail (GFSB200_ERR_NOMEM, "build_stencils: out of memory");
  }
  t->vtx_off[0] = 0;
#pragma omp parallel for schedule(dynamic, 4096)
  for (int32_t v = 0; v < nv; v++) {
    interp_t inter;
    corner_interp (t, rep[2*v], cd[rep[2*v + 1]], &inter);
    t->vtx_off[v + 1] = inter.n;
  }
  int64_t total = 0;
  for (int32_t v = 0; v < nv; v++) {
    int32_t c = t->vtx_off[v + 1];
    t->vtx_off[v] = (int32_t) total;
    total += c;
    if (total > INT32_MAX) {
      free (rep); free (vkey);
      return gfsb200_fail (GFSB200_ERR_UNSUPPORTED, "build_stencils: stencil table too large");
    }
  }
  t->vtx_off[nv] = (int32_t) total;
  t->vtx_cell = malloc ((size_t) (total ? total : 1)*sizeof (int32_t));
  t->vtx_w = malloc ((size_t) (total ? total : 1)*sizeof (double));
  if (!t->vtx_cell || !t->vtx_w) {
    free (rep); free (vkey);
    return gfsb200_fail (GFSB200_ERR_NOMEM, "build_stencils: out of memory");
  }
#pragma omp parallel for schedule(dynamic, 4096)
  for (int32_t v = 0; v < nv; v++) {
    interp_t inter;
    corner_interp (t, rep[2*v], cd[rep[2*v + 1]], &inter);
    int32_t o = t->vtx_off[v];
    for (int i = 0; i < inter.n; i++) {
      t->vtx_cell[o + i] = inter.c[i];
      t->vtx_w[o + i] = inter.w[i];
    }
  }
  free (rep);
  t->n_vertices = nv;
  t->lattice_level = -1;
  lattice_renumber (t, vkey);
  free (vkey);
  return GFSB200_OK;
}

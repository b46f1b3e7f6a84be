/* flat_tree.c -- host-side flat FTT: growth, 2:1 balancing, level-ordered
 * finalisation, neighbour tables, locate array.
 *
 * The reference keeps the quad/octree as individually malloc'ed FttOct blocks
 * linked by pointers (src/ftt.h:134-159).  Here a tree is a handful of flat
 * int32/uint8/double arrays indexed by cell number, so that the whole
 * topology can be uploaded to HBM in a few large copies and walked by index.
 *
 * Conventions shared with the reference (so that indices can be compared
 * bit-exactly with its pointer results):
 *   child n of a cell sits at  x: +h/2 if (n & 1) else -h/2
 *                              y: -h/2 if (n & 2) else +h/2
 *                              z: -h/2 if (n & 4) else +h/2     (src/ftt.c:301-316)
 *   directions  0:+x 1:-x 2:+y 3:-y 4:+z 5:-z, opposite = d ^ 1  (src/ftt.h:78-89)
 *   neighbour of a cell = same-level cell if it exists, else the one-level
 *   coarser cell, else none                                       (src/ftt.h:491-573)
 */
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <math.h>
#include "gfsb200_internal.h"

/* ------------------------------------------------------------------ */
/* error reporting                                                      */

static __thread char g_err[512];

const char * gfsb200_last_error (void) { return g_err; }
const char * gfsb200_version (void) { return "gfsb200 0.1 (sm_100a)"; }

int gfsb200_fail (int code, const char * fmt, ...)
{
  va_list ap;
  va_start (ap, fmt);
  vsnprintf (g_err, sizeof g_err, fmt, ap);
  va_end (ap);
  return code;
}

/* ------------------------------------------------------------------ */

static int grow (gfsb200_tree * t, int32_t extra)
{
  if (t->n_cells + extra <= t->cap)
    return GFSB200_OK;
  int64_t cap = t->cap ? t->cap : 1024;
  while (cap < (int64_t) t->n_cells + extra)
    cap *= 2;
  if (cap > INT32_MAX)
    return gfsb200_fail (GFSB200_ERR_NOMEM, "flat tree exceeds 2^31 cells");
  int32_t * parent = realloc (t->parent, cap*sizeof (int32_t));
  if (parent) t->parent = parent;
  int32_t * child0 = realloc (t->child0, cap*sizeof (int32_t));
  if (child0) t->child0 = child0;
  uint8_t * level = realloc (t->level, cap);
  if (level) t->level = level;
  uint8_t * flags = realloc (t->flags, cap);
  if (flags) t->flags = flags;
  double * pos = realloc (t->pos, cap*3*sizeof (double));
  if (pos) t->pos = pos;
  if (!parent || !child0 || !level || !flags || !pos)
    return gfsb200_fail (GFSB200_ERR_NOMEM, "out of memory growing flat tree");
  t->cap = (int32_t) cap;
  return GFSB200_OK;
}

gfsb200_tree * gfsb200_tree_new (int dim)
{
  if (dim != 2 && dim != 3) {
    gfsb200_fail (GFSB200_ERR_ARG, "dim must be 2 or 3");
    return NULL;
  }
  gfsb200_tree * t = calloc (1, sizeof (gfsb200_tree));
  if (!t) {
    gfsb200_fail (GFSB200_ERR_NOMEM, "out of memory");
    return NULL;
  }
  t->dim = dim;
  t->nchild = 1 << dim;
  t->ndir = 2*dim;
  return t;
}

static void free_final (gfsb200_tree * t)
{
  free (t->neighbor); t->neighbor = NULL;
  free (t->level_start); t->level_start = NULL;
  free (t->la_slot); t->la_slot = NULL;
  free (t->vtx_off); t->vtx_off = NULL;
  free (t->vtx_cell); t->vtx_cell = NULL;
  free (t->vtx_w); t->vtx_w = NULL;
  free (t->leaf_vtx); t->leaf_vtx = NULL;
  t->n_vertices = 0;
  t->lattice_level = -1;
  t->finalized = 0;
  free (t->solid_a); t->solid_a = NULL;          /* indexed like the finalized tree */
  free (t->solid_cm); t->solid_cm = NULL;
  free (t->solid_s); t->solid_s = NULL;
}

void gfsb200_tree_free (gfsb200_tree * t)
{
  if (!t) return;
  free_final (t);
  free (t->parent); free (t->child0); free (t->level); free (t->flags); free (t->pos);
  free (t->solid_a); free (t->solid_cm); free (t->solid_s);
  free (t);
}

/* GFS_IS_MIXED cells (src/fluid.h:93): the two members of their GfsSolidVector the
 * particulate path reads -- the centre of mass enters the corner-interpolator weights
 * (distance (), src/fluid.c:2983-3003), the fluid fraction gfs_cell_volume
 * (src/domain.h:503-508), the face fractions gfs_cell_face and average_neighbor_value
 * (src/fluid.c:42-52, 64-93), i.e. the gradients.  Indices are those of the finalized tree. */
int gfsb200_tree_set_solid (gfsb200_tree * t, int cell, double a, const double cm[3], const double * s)
{
  if (!t || !t->finalized)
    return gfsb200_fail (GFSB200_ERR_STATE, "set_solid: tree not finalized");
  if (cell < 0 || cell >= t->n_cells || !cm || !(a > 0.) || !(a <= 1.))
    return gfsb200_fail (GFSB200_ERR_ARG, "set_solid: bad argument");
  if (!t->solid_a) {
    t->solid_a = malloc ((size_t) t->n_cells*sizeof (double));
    t->solid_cm = malloc ((size_t) t->n_cells*3*sizeof (double));
    t->solid_s = malloc ((size_t) t->n_cells*t->ndir*sizeof (double));
    if (!t->solid_a || !t->solid_cm || !t->solid_s) {
      free (t->solid_a); free (t->solid_cm); free (t->solid_s);
      t->solid_a = t->solid_cm = t->solid_s = NULL;
      return gfsb200_fail (GFSB200_ERR_NOMEM, "set_solid: out of memory");
    }
    for (int32_t i = 0; i < t->n_cells; i++) {
      t->solid_a[i] = 1.;
      t->solid_cm[3*i] = t->solid_cm[3*i + 1] = t->solid_cm[3*i + 2] = NAN;
      for (int d = 0; d < t->ndir; d++)
	t->solid_s[(size_t) i*t->ndir + d] = 1.;
    }
  }
  t->solid_a[cell] = a;
  for (int d = 0; d < t->ndir; d++) {
    if (s && !(s[d] >= 0. && s[d] <= 1.))
      return gfsb200_fail (GFSB200_ERR_ARG, "set_solid: face fraction %g out of [0,1]", s[d]);
    t->solid_s[(size_t) cell*t->ndir + d] = s ? s[d] : 1.;
  }
  for (int k = 0; k < 3; k++)
    t->solid_cm[3*cell + k] = k < t->dim ? cm[k] : 0.;
  /* any stencil table built before is stale */
  free (t->vtx_off); free (t->vtx_cell); free (t->vtx_w); free (t->leaf_vtx);
  t->vtx_off = NULL; t->vtx_cell = NULL; t->vtx_w = NULL; t->leaf_vtx = NULL;
  t->n_vertices = 0;
  t->lattice_level = -1;
  return GFSB200_OK;
}

int gfsb200_tree_add_root (gfsb200_tree * t, const double pos[3], int level, int is_box)
{
  if (!t || !pos || level < 0 || level > 30)
    return gfsb200_fail (GFSB200_ERR_ARG, "add_root: bad argument");
  if (t->n_cells != t->n_roots)
    return gfsb200_fail (GFSB200_ERR_STATE, "add_root: roots must be added before any split");
  if (t->n_roots == GFSB200_MAX_ROOTS)
    return gfsb200_fail (GFSB200_ERR_UNSUPPORTED, "too many root cells");
  if (t->n_roots && level != t->root_level)
    return gfsb200_fail (GFSB200_ERR_ARG, "all roots must share the domain's rootlevel");
  if (is_box && t->n_roots != t->n_box_roots)
    return gfsb200_fail (GFSB200_ERR_STATE, "GfsBox roots must precede GfsBoundary roots");
  if (grow (t, 1))
    return GFSB200_ERR_NOMEM;
  int r = t->n_roots++;
  t->root_level = level;
  if (is_box) t->n_box_roots++;
  t->root_is_box[r] = is_box != 0;
  for (int d = 0; d < 6; d++) {
    t->root_nb[r][d] = -1;
    t->periodic[r][d] = -1;
  }
  int32_t i = t->n_cells++;
  t->parent[i] = -1;
  t->child0[i] = -1;
  t->level[i] = (uint8_t) level;
  t->flags[i] = is_box ? 0 : GFSB200_CELL_BOUNDARY;
  t->pos[3*i] = pos[0]; t->pos[3*i + 1] = pos[1]; t->pos[3*i + 2] = t->dim == 3 ? pos[2] : 0.;
  free_final (t);
  return r;
}

int gfsb200_tree_link_roots (gfsb200_tree * t, int r0, int d, int r1)
{
  if (!t || r0 < 0 || r1 < 0 || r0 >= t->n_roots || r1 >= t->n_roots || d < 0 || d >= t->ndir)
    return gfsb200_fail (GFSB200_ERR_ARG, "link_roots: bad argument");
  t->root_nb[r0][d] = r1;
  t->root_nb[r1][d ^ 1] = r0;
  free_final (t);
  return GFSB200_OK;
}

int gfsb200_tree_set_periodic (gfsb200_tree * t, int box_root, int side, int matching_box_root)
{
  if (!t || box_root < 0 || box_root >= t->n_box_roots || side < 0 || side >= t->ndir ||
      matching_box_root < 0 || matching_box_root >= t->n_box_roots)
    return gfsb200_fail (GFSB200_ERR_ARG, "set_periodic: bad argument");
  if (t->root_nb[box_root][side] < 0 || t->root_is_box[t->root_nb[box_root][side]])
    return gfsb200_fail (GFSB200_ERR_STATE, "set_periodic: side %d of box %d has no boundary (ghost) tree",
			 side, box_root);
  t->periodic[box_root][side] = matching_box_root;
  return GFSB200_OK;
}

/* is child n on the positive side of its parent along `axis`? */
static inline int child_positive (int n, int axis)
{
  return axis == 0 ? (n & 1) : !((n >> axis) & 1);
}

/* ftt_cell_neighbor semantics on the growing tree (src/ftt.h:491-573) */
int32_t gfsb200_tree_neighbor (const gfsb200_tree * t, int32_t i, int d)
{
  if (t->finalized)
    return t->neighbor[(int64_t) i*t->ndir + d];
  int32_t p = t->parent[i];
  if (p < 0)
    return t->root_nb[i][d];
  int n = i - t->child0[p];
  int axis = d >> 1, positive = !(d & 1);
  int32_t c;
  if (child_positive (n, axis) != positive)
    c = t->child0[p] + (n ^ (1 << axis));       /* sibling */
  else {
    c = gfsb200_tree_neighbor (t, p, d);
    if (c >= 0 && t->child0[c] >= 0)
      c = t->child0[c] + (n ^ (1 << axis));
  }
  if (c < 0 || (t->flags[c] & GFSB200_CELL_DESTROYED))
    return -1;
  return c;
}

int gfsb200_tree_split (gfsb200_tree * t, int cell, unsigned destroyed_mask, unsigned child_flags)
{
  if (!t || cell < 0 || cell >= t->n_cells)
    return gfsb200_fail (GFSB200_ERR_ARG, "split: bad cell");
  if (t->child0[cell] >= 0)
    return gfsb200_fail (GFSB200_ERR_STATE, "split: cell %d is not a leaf", cell);
  if (t->level[cell] >= GFSB200_MAX_LEVEL)
    return gfsb200_fail (GFSB200_ERR_UNSUPPORTED, "split: level limit %d reached", GFSB200_MAX_LEVEL);
  if (grow (t, t->nchild))
    return GFSB200_ERR_NOMEM;
  free_final (t);
  int32_t c0 = t->n_cells;
  t->n_cells += t->nchild;
  t->child0[cell] = c0;
  /* ftt_cell_pos, src/ftt.c:352-366: child centre = parent centre + coords[n]*h_child/2 */
  double size = ldexp (1., -(t->level[cell] + 1))/2.;
  for (int n = 0; n < t->nchild; n++) {
    int32_t c = c0 + n;
    t->parent[c] = cell;
    t->child0[c] = -1;
    t->level[c] = t->level[cell] + 1;
    t->flags[c] = (uint8_t) ((child_flags & ~(GFSB200_CELL_DESTROYED | GFSB200_CELL_LEAF)) |
			      (t->flags[cell] & GFSB200_CELL_BOUNDARY) |
			      ((destroyed_mask >> n) & 1 ? GFSB200_CELL_DESTROYED : 0));
    t->pos[3*c]     = t->pos[3*cell]     + (child_positive (n, 0) ? 1. : -1.)*size;
    t->pos[3*c + 1] = t->pos[3*cell + 1] + (child_positive (n, 1) ? 1. : -1.)*size;
    t->pos[3*c + 2] = t->dim == 3 ? t->pos[3*cell + 2] + (child_positive (n, 2) ? 1. : -1.)*size : 0.;
  }
  return c0;
}

/* oct_new (parent, check_neighbors = TRUE), src/ftt.c:45-84: coarser face
 * neighbours are refined first so that levels across a face differ by <= 1 */
int gfsb200_tree_refine_cell (gfsb200_tree * t, int cell)
{
  if (!t || cell < 0 || cell >= t->n_cells)
    return gfsb200_fail (GFSB200_ERR_ARG, "refine_cell: bad cell");
  if (t->child0[cell] >= 0)
    return gfsb200_fail (GFSB200_ERR_STATE, "refine_cell: cell %d is not a leaf", cell);
  for (int d = 0; d < t->ndir; d++) {
    int32_t q = gfsb200_tree_neighbor (t, cell, d);
    if (q >= 0 && t->level[q] < t->level[cell]) {
      int r = gfsb200_tree_refine_cell (t, q);
      if (r < 0) return r;
    }
  }
  return gfsb200_tree_split (t, cell, 0, 0);
}

/* ftt_cell_refine, src/ftt.c:169-192 */
static int refine_rec (gfsb200_tree * t, int32_t cell, gfsb200_refine_func f, void * data)
{
  if (t->child0[cell] < 0) {
    double h = ldexp (1., -t->level[cell]);
    if (!(* f) (&t->pos[3*cell], t->level[cell], h, data))
      return GFSB200_OK;
    int r = gfsb200_tree_refine_cell (t, cell);
    if (r < 0) return r;
  }
  for (int n = 0; n < t->nchild; n++) {
    int32_t c = t->child0[cell] + n;
    if (!(t->flags[c] & GFSB200_CELL_DESTROYED)) {
      int r = refine_rec (t, c, f, data);
      if (r < 0) return r;
    }
  }
  return GFSB200_OK;
}

static int ensure_unit_box (gfsb200_tree * t)
{
  if (t->n_roots == 0) {
    double o[3] = { 0., 0., 0. };
    int r = gfsb200_tree_add_root (t, o, 0, 1);
    if (r < 0) return r;
  }
  return GFSB200_OK;
}

int gfsb200_tree_refine (gfsb200_tree * t, gfsb200_refine_func f, void * data)
{
  if (!t || !f)
    return gfsb200_fail (GFSB200_ERR_ARG, "refine: bad argument");
  if (ensure_unit_box (t)) return GFSB200_ERR_STATE;
  for (int r = 0; r < t->n_box_roots; r++) {
    int e = refine_rec (t, r, f, data);
    if (e < 0) return e;
  }
  return GFSB200_OK;
}

static int uniform_crit (const double pos[3], int level, double h, void * data)
{
  return level < *(int *) data;
}

int gfsb200_tree_refine_uniform (gfsb200_tree * t, int level)
{
  return gfsb200_tree_refine (t, uniform_crit, &level);
}

typedef struct { int dim, minlevel, maxlevel; double R, factor; } ring_crit_t;

static int ring_crit (const double pos[3], int level, double h, void * data)
{
  ring_crit_t * r = data;
  if (level < r->minlevel) return 1;
  if (level >= r->maxlevel) return 0;
  double s = sqrt (pos[0]*pos[0] + pos[1]*pos[1]) - r->R;
  double dist = r->dim == 2 ? fabs (s) : sqrt (s*s + pos[2]*pos[2]);
  return dist < r->factor*h;
}

int gfsb200_tree_refine_ring (gfsb200_tree * t, int minlevel, int maxlevel, double R, double factor)
{
  if (!t) return gfsb200_fail (GFSB200_ERR_ARG, "refine_ring: null tree");
  ring_crit_t r = { t->dim, minlevel, maxlevel, R, factor };
  return gfsb200_tree_refine (t, ring_crit, &r);
}

/* ------------------------------------------------------------------ */
/* corner balance: ftt_refine_corner, src/ftt.c:2013-2075               */

static inline int is_leaf (const gfsb200_tree * t, int32_t c) { return t->child0[c] < 0; }

static int refine_corner (const gfsb200_tree * t, int32_t cell)
{
  for (int i = 0; i < t->ndir; i++) {
    int32_t n = gfsb200_tree_neighbor (t, cell, i);
    if (n < 0 || is_leaf (t, n))
      continue;
    int axis = i >> 1, od = i ^ 1, od_pos = !(od & 1);
    /* children of n on the face towards `cell` */
    for (int k = 0; k < t->nchild; k++) {
      if (child_positive (k, axis) != od_pos)
	continue;
      int32_t c = t->child0[n] + k;
      if (t->flags[c] & GFSB200_CELL_DESTROYED)
	continue;
      /* outward neighbours of c along the other axes */
      for (int a = 0; a < t->dim; a++) {
	if (a == axis) continue;
	int dir = 2*a + (child_positive (k, a) ? 0 : 1);
	int32_t nc = gfsb200_tree_neighbor (t, c, dir);
	if (nc >= 0 && !is_leaf (t, nc))
	  return 1;
      }
      if (!is_leaf (t, c))
	for (int m = 0; m < t->nchild; m++)
	  if (child_positive (m, axis) == od_pos &&
	      !(t->flags[t->child0[c] + m] & GFSB200_CELL_DESTROYED))
	    return 1;
    }
  }
  return 0;
}

/* cell_traverse_level, src/ftt.c:814-832, applying refine_cell_corner,
 * src/simulation.c:1105-1109 */
static int sweep_level (gfsb200_tree * t, int32_t cell, int level)
{
  if (t->level[cell] == level) {
    if (is_leaf (t, cell) && refine_corner (t, cell)) {
      int r = gfsb200_tree_refine_cell (t, cell);
      if (r < 0) return r;
    }
  }
  else if (!is_leaf (t, cell)) {
    int32_t c0 = t->child0[cell];
    for (int n = 0; n < t->nchild; n++)
      if (!(t->flags[c0 + n] & GFSB200_CELL_DESTROYED)) {
	int r = sweep_level (t, c0 + n, level);
	if (r < 0) return r;
      }
  }
  return GFSB200_OK;
}

static int depth_rec (const gfsb200_tree * t, int32_t cell)
{
  int depth = t->level[cell];
  if (!is_leaf (t, cell))
    for (int n = 0; n < t->nchild; n++) {
      int32_t c = t->child0[cell] + n;
      if (!(t->flags[c] & GFSB200_CELL_DESTROYED)) {
	int d = depth_rec (t, c);
	if (d > depth) depth = d;
      }
    }
  return depth;
}

int gfsb200_tree_corner_sweep (gfsb200_tree * t)
{
  if (!t) return gfsb200_fail (GFSB200_ERR_ARG, "corner_sweep: null tree");
  if (t->n_roots != t->n_box_roots)
    return gfsb200_fail (GFSB200_ERR_STATE, "corner_sweep: add boundaries after refinement");
  int depth = 0;
  for (int r = 0; r < t->n_box_roots; r++) {
    int d = depth_rec (t, r);
    if (d > depth) depth = d;
  }
  for (int l = depth - 2; l >= 0; l--)
    for (int r = 0; r < t->n_box_roots; r++) {
      int e = sweep_level (t, r, l);
      if (e < 0) return e;
    }
  return GFSB200_OK;
}

/* ------------------------------------------------------------------ */
/* ghost trees: boundary_match / match, src/boundary.c:576-685          */

static int match_ghost (gfsb200_tree * t, int32_t g, int32_t c, int dtoward)
{
  t->flags[g] |= GFSB200_CELL_BOUNDARY;
  if (is_leaf (t, c))
    return GFSB200_OK;
  int axis = dtoward >> 1, near_pos = !(dtoward & 1);
  unsigned destroyed = 0;
  for (int n = 0; n < t->nchild; n++) {
    if (child_positive (n, axis) != near_pos)
      destroyed |= 1u << n;                     /* ftt_cell_flatten: far side */
    else if (t->flags[t->child0[c] + (n ^ (1 << axis))] & GFSB200_CELL_DESTROYED)
      destroyed |= 1u << n;                     /* match(): neighbor == NULL */
  }
  int g0 = gfsb200_tree_split (t, g, destroyed, GFSB200_CELL_BOUNDARY);
  if (g0 < 0) return g0;
  for (int n = 0; n < t->nchild; n++)
    if (!((destroyed >> n) & 1)) {
      int r = match_ghost (t, g0 + n, t->child0[c] + (n ^ (1 << axis)), dtoward);
      if (r < 0) return r;
    }
  return GFSB200_OK;
}

int gfsb200_tree_add_boundary (gfsb200_tree * t, int box_root, int side)
{
  if (!t || box_root < 0 || box_root >= t->n_box_roots || side < 0 || side >= t->ndir)
    return gfsb200_fail (GFSB200_ERR_ARG, "add_boundary: bad argument");
  if (t->root_nb[box_root][side] >= 0)
    return gfsb200_fail (GFSB200_ERR_STATE, "add_boundary: side %d already has a neighbour", side);
  /* Roots must be contiguous at the head of the array and precede all
     splits, but boundaries are created after refinement: rebuild the arrays
     with one more root slot. */
  int32_t nr = t->n_roots;
  if (nr == GFSB200_MAX_ROOTS)
    return gfsb200_fail (GFSB200_ERR_UNSUPPORTED, "too many root cells");
  if (grow (t, 1))
    return GFSB200_ERR_NOMEM;
  free_final (t);
  int32_t n = t->n_cells;
  memmove (t->parent + nr + 1, t->parent + nr, (size_t) (n - nr)*sizeof (int32_t));
  memmove (t->child0 + nr + 1, t->child0 + nr, (size_t) (n - nr)*sizeof (int32_t));
  memmove (t->level + nr + 1, t->level + nr, (size_t) (n - nr));
  memmove (t->flags + nr + 1, t->flags + nr, (size_t) (n - nr));
  memmove (t->pos + 3*(nr + 1), t->pos + 3*nr, (size_t) (n - nr)*3*sizeof (double));
  t->n_cells = n + 1;
  for (int32_t i = 0; i < t->n_cells; i++) {
    if (i == nr) continue;
    if (t->parent[i] >= nr) t->parent[i]++;
    if (t->child0[i] >= nr) t->child0[i]++;
  }
  /* the new root: one root-cell size beyond the box on `side`
     (rpos[], src/boundary.c:662-668) */
  double size = ldexp (1., -t->root_level);
  t->parent[nr] = -1;
  t->child0[nr] = -1;
  t->level[nr] = (uint8_t) t->root_level;
  t->flags[nr] = GFSB200_CELL_BOUNDARY;
  for (int a = 0; a < 3; a++)
    t->pos[3*nr + a] = t->pos[3*box_root + a];
  t->pos[3*nr + (side >> 1)] += (side & 1 ? -1. : 1.)*size;
  t->n_roots = nr + 1;
  t->root_is_box[nr] = 0;
  for (int d = 0; d < 6; d++) {
    t->root_nb[nr][d] = -1;
    t->periodic[nr][d] = -1;
  }
  t->root_nb[nr][side ^ 1] = box_root;
  t->root_nb[box_root][side] = nr;
  return match_ghost (t, nr, box_root, side ^ 1);
}

/* ------------------------------------------------------------------ */
/* finalisation                                                         */

int gfsb200_tree_finalize (gfsb200_tree * t, int32_t * perm_out)
{
  if (!t || t->n_roots == 0)
    return gfsb200_fail (GFSB200_ERR_STATE, "finalize: tree has no root");
  free_final (t);
  int32_t n = t->n_cells;
  int32_t * order = malloc ((size_t) n*sizeof (int32_t));   /* new -> old */
  int32_t * perm = malloc ((size_t) n*sizeof (int32_t));    /* old -> new */
  if (!order || !perm) {
    free (order); free (perm);
    return gfsb200_fail (GFSB200_ERR_NOMEM, "finalize: out of memory");
  }
  /* breadth-first: roots, then sibling groups in parent order */
  int32_t head = 0, tail = 0;
  for (int r = 0; r < t->n_roots; r++)
    order[tail++] = r;
  while (head < tail) {
    int32_t c = order[head++];
    if (t->child0[c] >= 0)
      for (int k = 0; k < t->nchild; k++)
	order[tail++] = t->child0[c] + k;
  }
  if (tail != n) {
    free (order); free (perm);
    return gfsb200_fail (GFSB200_ERR_STATE, "finalize: %d cells unreachable from the roots", n - tail);
  }
  for (int32_t i = 0; i < n; i++)
    perm[order[i]] = i;

  int32_t * parent = malloc ((size_t) n*sizeof (int32_t));
  int32_t * child0 = malloc ((size_t) n*sizeof (int32_t));
  uint8_t * level = malloc ((size_t) n);
  uint8_t * flags = malloc ((size_t) n);
  double * pos = malloc ((size_t) n*3*sizeof (double));
  t->neighbor = malloc ((size_t) n*t->ndir*sizeof (int32_t));
  if (!parent || !child0 || !level || !flags || !pos || !t->neighbor) {
    free (order); free (perm); free (parent); free (child0); free (level); free (flags); free (pos);
    free_final (t);
    return gfsb200_fail (GFSB200_ERR_NOMEM, "finalize: out of memory");
  }
  int min_level = t->root_level, max_level = t->root_level;
  int64_t n_leaves = 0;
  for (int32_t i = 0; i < n; i++) {
    int32_t o = order[i];
    parent[i] = t->parent[o] < 0 ? -1 : perm[t->parent[o]];
    child0[i] = t->child0[o] < 0 ? -1 : perm[t->child0[o]];
    level[i] = t->level[o];
    uint8_t f = t->flags[o] & (GFSB200_CELL_DESTROYED | GFSB200_CELL_BOUNDARY);
    if (t->child0[o] < 0 && !(f & GFSB200_CELL_DESTROYED)) {
      f |= GFSB200_CELL_LEAF;
      if (!(f & GFSB200_CELL_BOUNDARY)) n_leaves++;
    }
    flags[i] = f;
    memcpy (pos + 3*i, t->pos + 3*o, 3*sizeof (double));
    if (level[i] > max_level && !(f & GFSB200_CELL_DESTROYED)) max_level = level[i];
  }
  free (t->parent); free (t->child0); free (t->level); free (t->flags); free (t->pos);
  t->parent = parent; t->child0 = child0; t->level = level; t->flags = flags; t->pos = pos;
  t->cap = n;
  t->min_level = min_level; t->max_level = max_level; t->n_leaves = n_leaves;

  /* level starts (breadth-first order is level order) */
  int n_levels = max_level - min_level + 1;
  /* destroyed cells may sit one level deeper than max_level */
  int deepest = level[n - 1];
  if (deepest > max_level) n_levels = deepest - min_level + 1;
  t->n_levels = n_levels;
  t->level_start = malloc ((size_t) (n_levels + 1)*sizeof (int32_t));
  {
    int l = 0;
    t->level_start[0] = 0;
    for (int32_t i = 0; i < n; i++)
      while (level[i] - min_level > l)
	t->level_start[++l] = i;
    while (l < n_levels)
      t->level_start[++l] = n;
  }

  /* neighbour table, parents before children */
  for (int32_t i = 0; i < n; i++) {
    int32_t p = parent[i];
    int32_t * nb = t->neighbor + (int64_t) i*t->ndir;
    if (p < 0) {
      for (int d = 0; d < t->ndir; d++)
	nb[d] = t->root_nb[i][d];
      continue;
    }
    int k = i - child0[p];
    for (int d = 0; d < t->ndir; d++) {
      int axis = d >> 1, positive = !(d & 1);
      int32_t c;
      if (child_positive (k, axis) != positive)
	c = child0[p] + (k ^ (1 << axis));
      else {
	c = t->neighbor[(int64_t) p*t->ndir + d];
	if (c >= 0 && child0[c] >= 0)
	  c = child0[c] + (k ^ (1 << axis));
      }
      nb[d] = (c < 0 || (flags[c] & GFSB200_CELL_DESTROYED)) ? -1 : c;
    }
  }

  /* deepest level to which every GfsBox tree is complete */
  {
    int cl = min_level;
    for (int l = min_level + 1; l <= max_level; l++) {
      int64_t want = (int64_t) t->n_box_roots << (t->dim*(l - min_level));
      int32_t s = t->level_start[l - min_level], e = t->level_start[l - min_level + 1];
      if (e - s < want) break;
      int ok = 1;
      for (int64_t j = 0; j < want && ok; j++)
	if (flags[s + j] & (GFSB200_CELL_DESTROYED | GFSB200_CELL_BOUNDARY))
	  ok = 0;
      if (!ok) break;
      cl = l;
    }
    t->complete_level = cl;
  }

  /* GfsLocateArray, src/domain.c:109-131 */
  {
    double h = ldexp (1., -t->root_level);     /* ftt_level_size (rootlevel) */
    double mn[3], mx[3];
    for (int c = 0; c < 3; c++) { mn[c] = 1.7976931348623157e308; mx[c] = -mn[c]; }
    for (int r = 0; r < t->n_roots; r++)
      for (int c = 0; c < t->dim; c++) {
	double p = pos[3*r + c];
	if (p + h/2. > mx[c]) mx[c] = p + h/2.;
	if (p - h/2. < mn[c]) mn[c] = p - h/2.;
      }
    int64_t size = 1;
    for (int c = 0; c < 3; c++) {
      if (c < t->dim) {
	t->la_n[c] = (int32_t) ceil ((mx[c] - mn[c])/h - 0.5);
	t->la_min[c] = mn[c];
      }
      else {
	t->la_n[c] = 1;
	t->la_min[c] = 0.;
      }
      size *= t->la_n[c];
    }
    t->la_h = h;
    if (size <= 0 || size > (1 << 24)) {
      free (order); free (perm);
      return gfsb200_fail (GFSB200_ERR_UNSUPPORTED, "locate array of %lld slots", (long long) size);
    }
    t->la_size = (int32_t) size;
    t->la_slot = malloc ((size_t) size*sizeof (int32_t));
    char * has_boundary = calloc ((size_t) size, 1);
    for (int64_t j = 0; j < size; j++) t->la_slot[j] = -1;
    for (int r = 0; r < t->n_roots; r++) {
      int64_t index = 0;
      int ok = 1;
      for (int c = 0; c < t->dim; c++) {
	int ic = (int) floor ((pos[3*r + c] - t->la_min[c])/h);
	if (ic < 0 || ic >= t->la_n[c]) ok = 0;
	index = index*t->la_n[c] + ic;
      }
      if (!ok) continue;
      /* box_index(): the GfsBox is stored first, boundaries are PREPENDED, and
	 gfs_domain_locate only accepts a slot whose first entry is a GfsBox
	 (src/domain.c:82-98, 2632) */
      if (t->root_is_box[r]) {
	if (!has_boundary[index]) t->la_slot[index] = r;
      }
      else {
	has_boundary[index] = 1;
	t->la_slot[index] = -1;
      }
    }
    free (has_boundary);
  }

  if (perm_out)
    memcpy (perm_out, perm, (size_t) n*sizeof (int32_t));
  free (order); free (perm);
  t->finalized = 1;
  return GFSB200_OK;
}

int gfsb200_tree_get_view (const gfsb200_tree * t, gfsb200_tree_view * v)
{
  if (!t || !v)
    return gfsb200_fail (GFSB200_ERR_ARG, "get_view: null argument");
  if (!t->finalized)
    return gfsb200_fail (GFSB200_ERR_STATE, "get_view: tree not finalized");
  memset (v, 0, sizeof *v);
  v->dim = t->dim; v->n_cells = t->n_cells; v->n_roots = t->n_roots; v->n_box_roots = t->n_box_roots;
  v->min_level = t->min_level; v->max_level = t->max_level; v->complete_level = t->complete_level;
  v->n_leaves = t->n_leaves;
  v->level_start = t->level_start;
  v->parent = t->parent; v->child0 = t->child0; v->neighbor = t->neighbor;
  v->level = t->level; v->flags = t->flags; v->pos = t->pos;
  for (int c = 0; c < 3; c++) { v->la_min[c] = t->la_min[c]; v->la_n[c] = t->la_n[c]; }
  v->la_h = t->la_h;
  v->la_slot = t->la_slot;
  v->n_vertices = t->n_vertices;
  v->vtx_off = t->vtx_off; v->vtx_cell = t->vtx_cell; v->vtx_w = t->vtx_w; v->leaf_vtx = t->leaf_vtx;
  v->lattice_level = t->leaf_vtx ? t->lattice_level : -1;
  v->solid_a = t->solid_a;
  v->solid_cm = t->solid_cm;
  v->solid_s = t->solid_s;
  return GFSB200_OK;
}

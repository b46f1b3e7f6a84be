/* TEST INFRASTRUCTURE ONLY (oracle) -- never linked into, imported by or
 * executed from the product path.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this.
 *
 * CPU restatement of the Gerris Lagrangian particulate hot path, layered on
 * the REFERENCE'S OWN object code: /root/reference/src/ftt.c and src/fluid.c
 * are compiled unmodified (oracle/Makefile) and supply tree construction and
 * refinement, ftt_cell_locate, neighbours, gfs_cell_corner_interpolator,
 * gfs_interpolate and gfs_center_gradient.  What cannot be compiled out of
 * the reference tree (static functions tangled with GtsObject/GfsSimulation)
 * is restated here, each function citing the file:line it follows:
 *
 *   GfsLocateArray + gfs_domain_locate   src/domain.c:43-145, 2623-2638
 *   gfs_cell_init / variables size       src/domain.c:2932-2954, src/domain.h:147-148
 *   boundary ghost trees                 src/boundary.c:576-685, 840-858
 *   corner-balance sweep                 src/simulation.c:1105-1109, 1226-1231
 *   vorticity_vector                     modules/particulatecommon.c:142-164
 *   GfsForceLift / Drag / Buoy           modules/particulatecommon.c:423-490, 519-588, 617-655
 *   compute_forces + integrator          modules/particulatecommon.c:737-751, 768-842
 *   remove_particles_not_in_domain       modules/particulatecommon.c:955-969
 *   GfsParticulateField deposit          modules/particulatecommon.c:1929-1957
 *   force deposit (single-cell limit)    modules/particulatecommon.c:753-765, 2158-2228
 *   force deposit with smoothing kernel  modules/particulatecommon.c:2087-2228
 *   passive tracer advection             src/particle.c:31-44, src/domain.c:2764-2788
 *   particle text format                 src/particle.c:86-98, modules/particulatecommon.c:910-926
 *   particle BCs (periodic wrap / drop)  modules/particulatecommon.c:3058-3214, 3318-3395
 *
 * PARITY PIN: the reference ships no test, example or golden vector for the
 * particulates module.  This restatement is pinned to
 * OUTPUTS OF THE REFERENCE ITSELF RUN HERE: libgfsrefobj (oracle/refobj/glue.c)
 * holds modules/particulatecommon.c, src/event.c and src/particle.c compiled
 * unmodified, and tests/test_reference_objcode.py requires every function
 * restated below to reproduce them BIT FOR BIT (events, each force model,
 * inertial/added mass, tracers, cull, particle BCs, both deposits, the text
 * block).  Not covered by object code, because src/domain.c and
 * src/boundary.c do not compile here: the GfsLocateArray arithmetic, the
 * ghost-tree construction and gfs_cell_init (hand-derived known answers in
 * tests/test_oracle.py).
 *
 * Built with -ffp-contract=off to match the reference's non-FMA x86-64 build.
 */
#include <stdint.h>
#include "gfs_shadow.h"
#ifdef _OPENMP
# include <omp.h>
#endif

#ifndef M_PI
# define M_PI 3.14159265358979323846
#endif

#define ORA_MAXVAR 16
#define ORA_MAXBOX 8

typedef struct {
  GfsDomain domain;
  GfsVariable var[ORA_MAXVAR];
  int nvar;
  /* GfsBox root cells: a chain of nbox unit boxes along +x, box k centred at
     (k,0,0) and linked to box k+1 through their root neighbours (how a
     multi-box GfsDomain looks to ftt.c); `root` is box 0 */
  FttCell * root;
  FttCell * box[ORA_MAXBOX];
  int nbox;
  FttCell * broot[ORA_MAXBOX][FTT_NEIGHBORS];    /* GfsBoundary roots, by box and side */
  /* GfsLocateArray (src/domain.h:32-36) */
  double min[3], max[3], h;
  int n[3], size;
  signed char * slot;                /* 0: empty, 1 + b: GfsBox b first, -1: boundaries only */
} OraSim;

/* ------------------------------------------------------------------ */
/* cell payload: src/domain.c:2932-2954 + src/domain.h:147-148         */

static size_t variables_size (OraSim * sim)
{
  return sizeof (GfsStateVector) + sizeof (gdouble)*(MAX (sim->nvar + 1, 1) - 1);
}

static void cell_init (FttCell * cell, OraSim * sim)
{
  if (FTT_CELL_IS_LEAF (cell)) {
    g_return_if_fail (cell->data == NULL);
    cell->data = g_malloc0 (variables_size (sim));
  }
  else {
    FttCellChildren child;
    guint n;

    ftt_cell_children (cell, &child);
    for (n = 0; n < FTT_CELLS; n++) {
      g_return_if_fail (child.c[n]->data == NULL);
      child.c[n]->data = g_malloc0 (variables_size (sim));
    }
    if (GFS_CELL_IS_BOUNDARY (cell))
      for (n = 0; n < FTT_CELLS; n++)
	child.c[n]->flags |= GFS_FLAG_BOUNDARY;
  }
}

static void cell_cleanup (FttCell * cell, gpointer data)
{
  if (cell->data) {
    g_free (GFS_STATE (cell)->solid);          /* ora_set_solid */
    g_free (cell->data);
    cell->data = NULL;
  }
}

OraSim * ora_sim_new (int nvar)
{
  OraSim * sim = g_malloc0 (sizeof (OraSim));
  int i;
  g_assert (nvar >= FTT_DIMENSION && nvar <= ORA_MAXVAR);
  sim->nvar = nvar;
  for (i = 0; i < nvar; i++) {
    sim->var[i].i = i;
    sim->var[i].centered = FALSE; /* src/variable.c:114 */
    sim->var[i].domain = &sim->domain;
  }
  sim->domain.rootlevel = 0;
  sim->root = ftt_cell_new ((FttCellInitFunc) cell_init, sim);
  sim->box[0] = sim->root;
  sim->nbox = 1;
  return sim;
}

/* appends a GfsBox to the right of the last one (before any refinement):
 * what "GfsBox ... GEdge k k+1 right" builds, src/boundary.c gfs_gedge_link_boxes */
int ora_add_box (OraSim * sim)
{
  FttVector pos = { sim->nbox, 0., 0., 0. };
  FttCell * b;
  g_assert (sim->nbox < ORA_MAXBOX);
  b = ftt_cell_new ((FttCellInitFunc) cell_init, sim);
  ftt_cell_set_pos (b, &pos);
  ftt_cell_set_neighbor (sim->box[sim->nbox - 1], b, FTT_RIGHT, (FttCellInitFunc) cell_init, sim);
  sim->box[sim->nbox] = b;
  return sim->nbox++;
}

int ora_nbox (OraSim * sim) { return sim->nbox; }
uint64_t ora_box_root (OraSim * sim, int b) { return (uint64_t) (uintptr_t) sim->box[b]; }
uint64_t ora_box_boundary_root (OraSim * sim, int b, int d)
{
  return (uint64_t) (uintptr_t) sim->broot[b][d];
}

void ora_sim_destroy (OraSim * sim)
{
  FttDirection d;
  int b;
  for (b = 0; b < sim->nbox; b++)
    for (d = 0; d < FTT_NEIGHBORS; d++)
      if (sim->broot[b][d])
	ftt_cell_destroy (sim->broot[b][d], cell_cleanup, NULL);
  for (b = 0; b < sim->nbox; b++)
    ftt_cell_destroy (sim->box[b], cell_cleanup, NULL);
  free (sim->slot);
  free (sim);
}

uint64_t ora_root (OraSim * sim) { return (uint64_t) (uintptr_t) sim->root; }
uint64_t ora_boundary_root (OraSim * sim, int d)
{
  return (uint64_t) (uintptr_t) sim->broot[0][d];
}
int ora_dimension (void) { return FTT_DIMENSION; }

/* ------------------------------------------------------------------ */
/* refinement                                                           */

typedef struct {
  int minlevel, maxlevel;
  double R, factor;
} RingRefine;

static gboolean refine_uniform (FttCell * cell, gpointer data)
{
  return ftt_cell_level (cell) < *(guint *) data;
}

void ora_refine_uniform (OraSim * sim, int level)
{
  guint l = level;
  int b;
  for (b = 0; b < sim->nbox; b++)
    ftt_cell_refine (sim->box[b], refine_uniform, &l, (FttCellInitFunc) cell_init, sim);
}

/* refine while level < maxlevel and the distance
 * from the cell centre to the circle of radius R in the plane z = 0 is less
 * than factor*h_cell (in 2D: distance to the circle in the plane). */
static gboolean refine_ring (FttCell * cell, gpointer data)
{
  RingRefine * r = data;
  guint level = ftt_cell_level (cell);
  FttVector p;
  if ((gint) level < r->minlevel)
    return TRUE;
  if ((gint) level >= r->maxlevel)
    return FALSE;
  ftt_cell_pos (cell, &p);
  gdouble s = sqrt (p.x*p.x + p.y*p.y) - r->R;
#if FTT_2D
  gdouble dist = fabs (s);
#else
  gdouble dist = sqrt (s*s + p.z*p.z);
#endif
  return dist < r->factor*ftt_cell_size (cell);
}

void ora_refine_ring (OraSim * sim, int minlevel, int maxlevel, double R, double factor)
{
  RingRefine r = { minlevel, maxlevel, R, factor };
  int b;
  for (b = 0; b < sim->nbox; b++)
    ftt_cell_refine (sim->box[b], refine_ring, &r, (FttCellInitFunc) cell_init, sim);
}

/* Refines, in the order given, the leaf cell of level level[i] containing
 * (x,y,z)[i]; lets tests build the reference tree that corresponds to an
 * arbitrary flat tree.  Returns the number of cells actually refined. */
int ora_refine_points (OraSim * sim, int n, const int * level,
		       const double * x, const double * y, const double * z)
{
  int i, done = 0;
  for (i = 0; i < n; i++) {
    FttVector p = { x[i], y[i], z ? z[i] : 0., 0. };
    FttCell * cell = NULL;
    int b;
    for (b = 0; b < sim->nbox && !cell; b++)
      cell = ftt_cell_locate (sim->box[b], p, level[i]);
    if (cell && FTT_CELL_IS_LEAF (cell) && (int) ftt_cell_level (cell) == level[i]) {
      ftt_cell_refine_single (cell, (FttCellInitFunc) cell_init, sim);
      done++;
    }
  }
  return done;
}

/* src/simulation.c:1105-1109 */
static void refine_cell_corner (FttCell * cell, OraSim * sim)
{
  if (FTT_CELL_IS_LEAF (cell) && ftt_refine_corner (cell))
    ftt_cell_refine_single (cell, (FttCellInitFunc) cell_init, sim);
}

/* src/simulation.c:1226-1231 (gfs_domain_cell_traverse visits GfsBox trees only) */
void ora_corner_sweep (OraSim * sim)
{
  gint l, depth = 0;
  int b;
  for (b = 0; b < sim->nbox; b++)
    depth = MAX (depth, (gint) ftt_cell_depth (sim->box[b]));      /* gfs_domain_depth */
  for (l = depth - 2; l >= 0; l--)
    for (b = 0; b < sim->nbox; b++)
      ftt_cell_traverse (sim->box[b], FTT_PRE_ORDER, FTT_TRAVERSE_LEVEL, l,
			 (FttCellTraverseFunc) refine_cell_corner, sim);
}

/* ------------------------------------------------------------------ */
/* boundaries: src/boundary.c:576-685 (match, boundary_match), 840-858  */

typedef struct {
  OraSim * sim;
  FttCell * root;
  FttDirection d;       /* direction from the ghost tree towards the box */
  guint depth;
  gboolean changed;
} OraBoundary;

static void match (FttCell * cell, OraBoundary * boundary)
{
  FttCell * neighbor = ftt_cell_neighbor (cell, boundary->d);
  guint level = ftt_cell_level (cell);

  cell->flags |= GFS_FLAG_BOUNDARY;
  if (neighbor == NULL || ftt_cell_level (neighbor) < level) {
    if (FTT_CELL_IS_ROOT (cell)) {
      g_assert (cell == boundary->root);
      boundary->root = NULL;
    }
    ftt_cell_destroy (cell, cell_cleanup, NULL);
    boundary->changed = TRUE;
    return;
  }
  if (ftt_cell_level (neighbor) == level) {
    /* no solid fractions in the oracle's worlds */
    if (FTT_CELL_IS_LEAF (cell) && !FTT_CELL_IS_LEAF (neighbor)) {
      ftt_cell_refine_single (cell, (FttCellInitFunc) cell_init, boundary->sim);
      boundary->changed = TRUE;
    }
  }
  else
    g_assert_not_reached ();
  if (!FTT_CELL_IS_LEAF (cell))
    level++;
  if (level > boundary->depth)
    boundary->depth = level;
}

/* gfs_boundary_new + the root == NULL branch of boundary_match
 * (src/boundary.c:652-668, 840-858).  side: the side of the box the boundary
 * sits on (box->neighbor[side]).  Gerris creates its boundaries while reading
 * the .gfs file, i.e. BEFORE refinement; call ora_match_boundaries() after
 * refinement, as gfs_simulation_refine does (src/simulation.c:1233). */
void ora_add_boundary_box (OraSim * sim, int b, int side);

void ora_add_boundary (OraSim * sim, int side) { ora_add_boundary_box (sim, 0, side); }

void ora_add_boundary_box (OraSim * sim, int b, int side)
{
  static FttVector rpos[6] = {
    {1.,0.,0.}, {-1.,0.,0.}, {0.,1.,0.}, {0.,-1.,0.}, {0.,0.,1.}, {0.,0.,-1.}
  };
  FttCell * root;
  FttVector pos;
  gdouble size;
  FttDirection d, od;

  g_assert (side >= 0 && side < FTT_NEIGHBORS && b >= 0 && b < sim->nbox && sim->broot[b][side] == NULL);
  d = FTT_OPPOSITE_DIRECTION (side);
  root = ftt_cell_new ((FttCellInitFunc) cell_init, sim);
  root->flags |= GFS_FLAG_BOUNDARY;
  ftt_cell_set_level (root, ftt_cell_level (sim->box[b]));
  ftt_cell_set_neighbor_match (root, sim->box[b], d, (FttCellInitFunc) cell_init, sim);
  ftt_cell_pos (sim->box[b], &pos);
  size = ftt_cell_size (sim->box[b]);
  od = FTT_OPPOSITE_DIRECTION (d);
  pos.x += rpos[od].x*size;
  pos.y += rpos[od].y*size;
  pos.z += rpos[od].z*size;
  ftt_cell_set_pos (root, &pos);
  sim->broot[b][side] = root;
}

/* gfs_domain_match -> boundary_match for every boundary (src/boundary.c:670-685) */
void ora_match_boundaries (OraSim * sim)
{
  int side, box;
  for (box = 0; box < sim->nbox; box++)
  for (side = 0; side < FTT_NEIGHBORS; side++)
    if (sim->broot[box][side]) {
      OraBoundary b;
      guint l;
      b.sim = sim;
      b.d = FTT_OPPOSITE_DIRECTION (side);
      b.root = sim->broot[box][side];
      l = ftt_cell_level (b.root);
      b.changed = FALSE;
      b.depth = l;
      while (b.root && l <= b.depth) {
	ftt_cell_traverse_boundary (b.root, b.d, FTT_PRE_ORDER, FTT_TRAVERSE_LEVEL, l,
				    (FttCellTraverseFunc) match, &b);
	l++;
      }
      if (b.root && b.changed)
	ftt_cell_flatten (b.root, b.d, cell_cleanup, NULL);
      sim->broot[box][side] = b.root;
    }
}

/* ------------------------------------------------------------------ */
/* GfsLocateArray: src/domain.c:43-145                                  */

static void locate_index (FttVector * p, OraSim * a, gint i[FTT_DIMENSION])
{
  gint c;
  for (c = 0; c < FTT_DIMENSION; c++)
    i[c] = floor (((&p->x)[c] - a->min[c])/a->h);
}

static void root_bounds (FttCell * root, OraSim * a)
{
  FttVector p;
  ftt_cell_pos (root, &p);
  gint i;
  for (i = 0; i < FTT_DIMENSION; i++) {
    if ((&p.x)[i] + a->h/2. > a->max[i]) a->max[i] = (&p.x)[i] + a->h/2.;
    if ((&p.x)[i] - a->h/2. < a->min[i]) a->min[i] = (&p.x)[i] - a->h/2.;
  }
}

static gint locate_linear_index (FttVector * p, OraSim * a)
{
  gint i[FTT_DIMENSION], index = 0, c;
  locate_index (p, a, i);
  for (c = 0; c < FTT_DIMENSION; c++) {
    if (i[c] < 0 || i[c] >= a->n[c])
      return -1;
    index = index*a->n[c] + i[c];
  }
  return index;
}

void ora_finalize (OraSim * sim)
{
  guint i;
  FttDirection d;
  FttVector p;
  gint k;

  sim->h = ftt_level_size (sim->domain.rootlevel);
  for (i = 0; i < FTT_DIMENSION; i++) {
    sim->min[i] = G_MAXDOUBLE;
    sim->max[i] = - G_MAXDOUBLE;
  }
  {
    int bb;
    for (bb = 0; bb < sim->nbox; bb++) {
      root_bounds (sim->box[bb], sim);
      for (d = 0; d < FTT_NEIGHBORS; d++)
	if (sim->broot[bb][d])
	  root_bounds (sim->broot[bb][d], sim);
    }
  }
  sim->size = 1;
  for (i = 0; i < FTT_DIMENSION; i++) {
    g_assert (sim->max[i] > sim->min[i]);
    sim->n[i] = ceil ((sim->max[i] - sim->min[i])/sim->h - 0.5);
    sim->size *= sim->n[i];
  }
  free (sim->slot);
  sim->slot = g_malloc0 (sim->size);
  {
    int bb;
    /* box_index(), src/domain.c:82-98: the box first, its boundaries PREPENDED;
       gfs_domain_locate only accepts a slot whose first entry is a GfsBox */
    for (bb = 0; bb < sim->nbox; bb++) {
      ftt_cell_pos (sim->box[bb], &p);
      k = locate_linear_index (&p, sim);
      g_assert (k >= 0 && !sim->slot[k]);
      sim->slot[k] = 1 + bb;
    }
    for (bb = 0; bb < sim->nbox; bb++)
      for (d = 0; d < FTT_NEIGHBORS; d++)
	if (sim->broot[bb][d]) {
	  ftt_cell_pos (sim->broot[bb][d], &p);
	  k = locate_linear_index (&p, sim);
	  g_assert (k >= 0);
	  sim->slot[k] = -1;
	}
  }
}

void ora_locate_array (OraSim * sim, double * min, double * h, int * n)
{
  int c;
  for (c = 0; c < FTT_DIMENSION; c++) {
    min[c] = sim->min[c];
    n[c] = sim->n[c];
  }
  *h = sim->h;
}

/* src/domain.c:2623-2638 */
static FttCell * domain_locate (OraSim * sim, FttVector target, gint max_depth)
{
  gint i = locate_linear_index (&target, sim);
  if (i >= 0 && sim->slot[i] >= 1)
    return ftt_cell_locate (sim->box[sim->slot[i] - 1], target, max_depth);
  return NULL;
}

void ora_locate (OraSim * sim, long n, const double * x, const double * y, const double * z,
		 uint64_t * cell)
{
  long i;
#pragma omp parallel for schedule(static)
  for (i = 0; i < n; i++) {
    FttVector p = { x[i], y[i], z ? z[i] : 0., 0. };
    cell[i] = (uint64_t) (uintptr_t) domain_locate (sim, p, -1);
  }
}

/* Makes `cell` a mixed (solid-cut) cell: attaches a GfsSolidVector (src/fluid.h:54-59) with
 * fluid fraction a, centre of mass cm and face fractions s[] (NULL: all 1), the members the
 * particulate path reads (distance () src/fluid.c:2983-3003 through gfs_cell_cm;
 * gfs_cell_volume; gfs_cell_face and average_neighbor_value in the gradients).  The reference
 * computes them from a GTS surface (src/solid.c, GTS is absent here); the tests prescribe
 * them.  a <= 0 removes the solid vector again. */
void ora_set_solid (uint64_t cellp, double a, const double * cm, const double * s)
{
  FttCell * cell = (FttCell *) (uintptr_t) cellp;
  FttDirection d;
  g_assert (cell && cell->data);
  if (a <= 0.) {
    g_free (GFS_STATE (cell)->solid);
    GFS_STATE (cell)->solid = NULL;
    return;
  }
  if (!GFS_STATE (cell)->solid)
    GFS_STATE (cell)->solid = g_malloc0 (sizeof (GfsSolidVector));
  GFS_STATE (cell)->solid->a = a;
  for (d = 0; d < FTT_NEIGHBORS; d++)
    GFS_STATE (cell)->solid->s[d] = s ? s[d] : 1.;
  GFS_STATE (cell)->solid->cm.x = cm[0];
  GFS_STATE (cell)->solid->cm.y = cm[1];
  GFS_STATE (cell)->solid->cm.z = FTT_DIMENSION > 2 ? cm[2] : 0.;
  GFS_STATE (cell)->solid->ca = GFS_STATE (cell)->solid->cm;
}

/* single-point entry: the gfs_domain_locate that libgfsrefobj (the reference's
 * particulate object code, oracle/refobj/glue.c) calls back into */
FttCell * ora_locate_one (OraSim * sim, double x, double y, double z, int max_depth)
{
  FttVector p = { x, y, z, 0. };
  return domain_locate (sim, p, max_depth);
}

/* ------------------------------------------------------------------ */
/* cell access helpers for the tests                                    */

void ora_cell_info (uint64_t cellp, int * level, double * pos, unsigned * flags, int * is_leaf)
{
  FttCell * cell = (FttCell *) (uintptr_t) cellp;
  FttVector p;
  ftt_cell_pos (cell, &p);
  *level = ftt_cell_level (cell);
  pos[0] = p.x; pos[1] = p.y; pos[2] = p.z;
  *flags = cell->flags;
  *is_leaf = FTT_CELL_IS_LEAF (cell);
}

void ora_set_values (OraSim * sim, int ivar, long n, const uint64_t * cell, const double * val)
{
  long i;
  for (i = 0; i < n; i++)
    GFS_VALUEI ((FttCell *) (uintptr_t) cell[i], ivar) = val[i];
}

void ora_get_values (OraSim * sim, int ivar, long n, const uint64_t * cell, double * val)
{
  long i;
  for (i = 0; i < n; i++)
    val[i] = GFS_VALUEI ((FttCell *) (uintptr_t) cell[i], ivar);
}

uint64_t ora_neighbor (uint64_t cellp, int d)
{
  return (uint64_t) (uintptr_t) ftt_cell_neighbor ((FttCell *) (uintptr_t) cellp, d);
}

typedef struct { long n; uint64_t * cell; double * pos; int * level; int * leaf; } OraExport;

static void export_cell (FttCell * cell, OraExport * e)
{
  FttVector p;
  ftt_cell_pos (cell, &p);
  e->cell[e->n] = (uint64_t) (uintptr_t) cell;
  e->pos[3*e->n] = p.x; e->pos[3*e->n + 1] = p.y; e->pos[3*e->n + 2] = p.z;
  e->level[e->n] = ftt_cell_level (cell);
  e->leaf[e->n] = FTT_CELL_IS_LEAF (cell);
  e->n++;
}

/* every cell of the GfsBox tree (pre-order): pointer, centre, level, leaf flag;
 * arrays must hold ora_count (sim, 0) entries */
long ora_export_cells (OraSim * sim, uint64_t * cell, double * pos, int * level, int * leaf)
{
  OraExport e = { 0, cell, pos, level, leaf };
  int b;
  for (b = 0; b < sim->nbox; b++)
    ftt_cell_traverse (sim->box[b], FTT_PRE_ORDER, FTT_TRAVERSE_ALL, -1,
		       (FttCellTraverseFunc) export_cell, &e);
  return e.n;
}

static void count_cell (FttCell * cell, long * n) { (*n)++; }

long ora_count (OraSim * sim, int leaves_only)
{
  long n = 0;
  int b;
  for (b = 0; b < sim->nbox; b++)
    ftt_cell_traverse (sim->box[b], FTT_PRE_ORDER,
		       leaves_only ? FTT_TRAVERSE_LEAFS : FTT_TRAVERSE_ALL, -1,
		       (FttCellTraverseFunc) count_cell, &n);
  return n;
}

/* ------------------------------------------------------------------ */
/* interpolation / gradients straight from the reference object code    */

#if FTT_2D
# define N_CORNERS 4
static FttDirection corner[4][FTT_DIMENSION] = {   /* src/fluid.c:2588-2594 */
  { FTT_LEFT,  FTT_BOTTOM }, { FTT_RIGHT, FTT_BOTTOM },
  { FTT_RIGHT, FTT_TOP },    { FTT_LEFT,  FTT_TOP }
};
#else
# define N_CORNERS 8
static FttDirection corner[8][FTT_DIMENSION] = {   /* src/fluid.c:2596-2605 */
  { FTT_LEFT,  FTT_BOTTOM, FTT_FRONT }, { FTT_RIGHT, FTT_BOTTOM, FTT_FRONT },
  { FTT_RIGHT, FTT_TOP,    FTT_FRONT }, { FTT_LEFT,  FTT_TOP,    FTT_FRONT },
  { FTT_LEFT,  FTT_BOTTOM, FTT_BACK },  { FTT_RIGHT, FTT_BOTTOM, FTT_BACK },
  { FTT_RIGHT, FTT_TOP,    FTT_BACK },  { FTT_LEFT,  FTT_TOP,    FTT_BACK }
};
#endif

void ora_interpolate (OraSim * sim, int ivar, long n,
		      const double * x, const double * y, const double * z, double * out)
{
  long i;
#pragma omp parallel for schedule(static)
  for (i = 0; i < n; i++) {
    FttVector p = { x[i], y[i], z ? z[i] : 0., 0. };
    FttCell * cell = domain_locate (sim, p, -1);
    out[i] = cell ? gfs_interpolate (cell, p, &sim->var[ivar]) : GFS_NODATA;
  }
}

/* the (cell, weight) list of gfs_cell_corner_interpolator for corner k */
int ora_corner_interpolator (OraSim * sim, uint64_t cellp, int k, uint64_t * cells, double * w)
{
  GfsInterpolator inter;
  guint i;
  gfs_cell_corner_interpolator ((FttCell *) (uintptr_t) cellp, corner[k], -1, FALSE, &inter);
  for (i = 0; i < inter.n; i++) {
    cells[i] = (uint64_t) (uintptr_t) inter.c[i];
    w[i] = inter.w[i];
  }
  return inter.n;
}

/* out[i*N_CORNERS + k] = gfs_cell_corner_value (cell[i], corner[k], v, -1) */
void ora_corner_values (OraSim * sim, int ivar, long n, const uint64_t * cell, double * out)
{
  long i;
#pragma omp parallel for schedule(static)
  for (i = 0; i < n; i++) {
    int k;
    for (k = 0; k < N_CORNERS; k++)
      out[i*N_CORNERS + k] = gfs_cell_corner_value ((FttCell *) (uintptr_t) cell[i], corner[k],
						    &sim->var[ivar], -1);
  }
}

void ora_center_gradient (OraSim * sim, int comp, int ivar, long n, const uint64_t * cell,
			  double * out)
{
  long i;
#pragma omp parallel for schedule(static)
  for (i = 0; i < n; i++)
    out[i] = gfs_center_gradient ((FttCell *) (uintptr_t) cell[i], comp, ivar);
}

/* modules/particulatecommon.c:142-164 */
static void vorticity_vector (FttCell * cell, GfsVariable ** v, FttVector * vort)
{
  gdouble size;

  if (cell == NULL) return;
  if (v == NULL) return;

  size = ftt_cell_size (cell);
#if FTT_2D
  vort->x = 0.;
  vort->y = 0.;
  vort->z = (gfs_center_gradient (cell, FTT_X, v[1]->i) -
	     gfs_center_gradient (cell, FTT_Y, v[0]->i))/size;
#else  /* FTT_3D */
  vort->x = (gfs_center_gradient (cell, FTT_Y, v[2]->i) -
	     gfs_center_gradient (cell, FTT_Z, v[1]->i))/size;
  vort->y = (gfs_center_gradient (cell, FTT_Z, v[0]->i) -
	     gfs_center_gradient (cell, FTT_X, v[2]->i))/size;
  vort->z = (gfs_center_gradient (cell, FTT_X, v[1]->i) -
	     gfs_center_gradient (cell, FTT_Y, v[0]->i))/size;
#endif
}

/* out[3*i..] = vorticity vector of cell[i] (U,V,W are variables 0,1,2) */
void ora_vorticity (OraSim * sim, long n, const uint64_t * cell, double * out)
{
  long i;
  GfsVariable * u[3] = { &sim->var[0], &sim->var[1], &sim->var[FTT_DIMENSION > 2 ? 2 : 1] };
#pragma omp parallel for schedule(static)
  for (i = 0; i < n; i++) {
    FttVector w = { 0., 0., 0., 0. };
    vorticity_vector ((FttCell *) (uintptr_t) cell[i], u, &w);
    out[3*i] = w.x; out[3*i + 1] = w.y; out[3*i + 2] = w.z;
  }
}

/* ------------------------------------------------------------------ */
/* particles: src/particle.h:34-39, modules/particulatecommon.h:35-41    */

enum { ORA_FORCE_DRAG = 1, ORA_FORCE_LIFT = 2, ORA_FORCE_BUOY = 3,
       ORA_FORCE_INERTIAL = 4, ORA_FORCE_ADDEDMASS = 5 };

typedef struct {
  double dt;
  int n_forces;
  int force[8];        /* ORA_FORCE_*, in GfsParticleList force-list order */
  double rho;          /* fluid density 1/alpha when ivar_alpha < 0 (alpha unset => 1) */
  int ivar_alpha;      /* >= 0: per-cell alpha variable, rho = 1/alpha(cell) */
  double mu;           /* GfsSourceDiffusion constant; 0 = no diffusion source */
  int ivar_mu;         /* >= 0: per-cell viscosity variable */
  double g[3];         /* sum of GfsSource intensities on U,V,W */
  double cd_const;     /* constant drag coefficient function, NaN = built-in law */
  double cl_const;     /* constant lift coefficient function, NaN = 0.5 */
  int pattern;         /* 0: reference call pattern (one locate + interpolation set per
			  force, dead vliq evaluation), 1: fused (one locate) */
  int ivar_uold;       /* first of the FTT_DIMENSION variables Un,Vn(,Wn) of GfsForceCoeff.Uold */
  double cm_const;     /* constant GfsForceAddedMass coefficient function, NaN = 0.5 */
} OraStepParams;

typedef struct _OraParticulate OraParticulate;
struct _OraParticulate {
  /* stand-in for the GfsEvent header the GtsObject carries (src/event.h:31-44) */
  char event_header[96];
  FttVector pos, pos_old;
  guint id;
  FttVector vel;
  gdouble mass, volume;
  FttVector force;
  gpointer forces;
};

typedef struct {
  OraSim * sim;
  const OraStepParams * par;
} OraCtx;

typedef FttVector (* OraForceFunc) (OraCtx * ctx, OraParticulate * p);

static gdouble fluid_rho_at (OraCtx * ctx, FttCell * cell)
{
  /* sim->physical_params.alpha ? 1./gfs_function_value (alpha, cell) : 1. */
  if (ctx->par->ivar_alpha >= 0)
    return 1./GFS_VALUEI (cell, ctx->par->ivar_alpha);
  return ctx->par->rho;
}

static gdouble viscosity_at (OraCtx * ctx, FttCell * cell)
{
  /* d ? gfs_diffusion_cell (d->D, cell) : 0. */
  if (ctx->par->ivar_mu >= 0)
    return GFS_VALUEI (cell, ctx->par->ivar_mu);
  return ctx->par->mu;
}

static FttVector subs_fttvectors (FttVector * a, FttVector * b)
{
  FttVector result;
  FttComponent c;
  result.z = 0.; result.r = 0.;
  for (c = 0; c < FTT_DIMENSION; c++)
    (&result.x)[c] = (&a->x)[c] - (&b->x)[c];
  return result;
}

/* modules/particulatecommon.c:423-490 */
static FttVector compute_lift_force (OraCtx * ctx, OraParticulate * p)
{
  OraSim * sim = ctx->sim;
  FttVector force = { 0., 0., 0., 0. };
  FttComponent c;

  FttCell * cell = domain_locate (sim, p->pos, -1);
  if (cell == NULL) return force;

  gdouble fluid_rho = fluid_rho_at (ctx, cell);
  GfsVariable * u[3] = { &sim->var[0], &sim->var[1], &sim->var[FTT_DIMENSION > 2 ? 2 : 1] };
  gdouble viscosity = viscosity_at (ctx, cell);

  FttVector fluid_vel = { 0., 0., 0., 0. };
  for (c = 0; c < FTT_DIMENSION; c++)
    (&fluid_vel.x)[c] = gfs_interpolate (cell, p->pos, u[c]);

  FttVector relative_vel = subs_fttvectors (&fluid_vel, &p->vel);
  FttVector vorticity = { 0., 0., 0., 0. };
  vorticity_vector (cell, u, &vorticity);

  gdouble cl = 0.5;
  if (ctx->par->cl_const == ctx->par->cl_const) {
    /* a coefficient GfsFunction is attached: the reference evaluates Re etc.
       into cell variables and then calls the function; for a constant
       function only the returned value matters */
    (void) viscosity;
    cl = ctx->par->cl_const;
  }

#if FTT_2D
  force.x = fluid_rho*cl*relative_vel.y*vorticity.z;
  force.y = -fluid_rho*cl*relative_vel.x*vorticity.z;
#else
  force.x = fluid_rho*cl*(relative_vel.y*vorticity.z
			  -relative_vel.z*vorticity.y);
  force.y = fluid_rho*cl*(relative_vel.z*vorticity.x
			  -relative_vel.x*vorticity.z);
  force.z = fluid_rho*cl*(relative_vel.x*vorticity.y
			  -relative_vel.y*vorticity.x);
#endif
  return force;
}

/* modules/particulatecommon.c:519-588 */
static FttVector compute_drag_force (OraCtx * ctx, OraParticulate * p)
{
  OraSim * sim = ctx->sim;
  FttVector force = { 0., 0., 0., 0. };
  FttComponent c;

  FttCell * cell = domain_locate (sim, p->pos, -1);
  if (cell == NULL) return force;

  gdouble fluid_rho = fluid_rho_at (ctx, cell);
  GfsVariable * u[3] = { &sim->var[0], &sim->var[1], &sim->var[FTT_DIMENSION > 2 ? 2 : 1] };
  gdouble viscosity = viscosity_at (ctx, cell);

  FttVector fluid_vel = { 0., 0., 0., 0. };
  for (c = 0; c < FTT_DIMENSION; c++)
    (&fluid_vel.x)[c] = gfs_interpolate (cell, p->pos, u[c]);

  FttVector relative_vel = subs_fttvectors (&fluid_vel, &p->vel);

  gdouble dia = 2.*pow(3.0*(p->volume)/4.0/M_PI, 1./3.);
#if !FTT_2D
  gdouble norm_relative_vel = sqrt (relative_vel.x*relative_vel.x +
				    relative_vel.y*relative_vel.y +
				    relative_vel.z*relative_vel.z);
#else
  gdouble norm_relative_vel = sqrt (relative_vel.x*relative_vel.x +
				    relative_vel.y*relative_vel.y);
#endif

  gdouble cd = 0.;
  gdouble Re;
  if (viscosity == 0)
    return force;
  else
    Re = norm_relative_vel*dia*fluid_rho/viscosity;

  if (ctx->par->cd_const == ctx->par->cd_const)
    cd = ctx->par->cd_const;
  else {
    if (Re < 1e-8)
      return force;
    else if (Re < 50.0)
      cd = 16.*(1. + 0.15*pow(Re,0.5))/Re;
    else
      cd = 48.*(1. - 2.21/pow(Re,0.5))/Re;
  }
  for (c = 0; c < FTT_DIMENSION; c++)
    (&force.x)[c] += 3./(4.*dia)*cd*norm_relative_vel*(&relative_vel.x)[c]*fluid_rho;

  return force;
}

/* modules/particulatecommon.c:617-655 */
static FttVector compute_buoyancy_force (OraCtx * ctx, OraParticulate * p)
{
  OraSim * sim = ctx->sim;
  FttVector force = { 0., 0., 0., 0. };
  FttComponent c;

  FttCell * cell = domain_locate (sim, p->pos, -1);
  if (cell == NULL) return force;

  gdouble fluid_rho = fluid_rho_at (ctx, cell);
  gdouble g[3];
  for (c = 0; c < FTT_DIMENSION; c++)
    g[c] = ctx->par->g[c];

  for (c = 0; c < FTT_DIMENSION; c++)
    (&force.x)[c] += (p->mass/p->volume-fluid_rho)*g[c];

  return force;
}

/* modules/particulatecommon.c:255-303 */
static FttVector compute_inertial_force (OraCtx * ctx, OraParticulate * p)
{
  OraSim * sim = ctx->sim;
  FttVector force = { 0., 0., 0., 0. };
  FttComponent c;

  FttCell * cell = domain_locate (sim, p->pos, -1);
  if (cell == NULL) return force;

  gdouble size = ftt_cell_size(cell);

  gdouble fluid_rho = fluid_rho_at (ctx, cell);
  GfsVariable * u[3] = { &sim->var[0], &sim->var[1], &sim->var[FTT_DIMENSION > 2 ? 2 : 1] };

  FttVector fluid_vel = { 0., 0., 0., 0. };
  for (c = 0; c < FTT_DIMENSION; c++)
    (&fluid_vel.x)[c] = gfs_interpolate (cell, p->pos, u[c]);

  FttVector fluid_veln = { 0., 0., 0., 0. };
  for (c = 0; c < FTT_DIMENSION; c++)
    (&fluid_veln.x)[c] = gfs_interpolate (cell, p->pos, &sim->var[ctx->par->ivar_uold + c]);

  if(ctx->par->dt > 0.)
    for (c = 0; c < FTT_DIMENSION; c++)
      (&force.x)[c] = fluid_rho*((&fluid_vel.x)[c]-(&fluid_veln.x)[c])/ctx->par->dt;
  else
    return force;

  FttComponent c2;
  for(c = 0; c < FTT_DIMENSION; c++)
    for(c2 = 0; c2 < FTT_DIMENSION; c2++)
      (&force.x)[c] += fluid_rho*gfs_center_gradient(cell, c2, u[c]->i)*
	GFS_VALUE(cell, u[c2])/size;

  return force;
}

/* modules/particulatecommon.c:331-394, including the cumulative mass update (:391) */
static FttVector compute_addedmass_force (OraCtx * ctx, OraParticulate * p)
{
  OraSim * sim = ctx->sim;
  FttVector force = { 0., 0., 0., 0. };
  FttComponent c;

  FttCell * cell = domain_locate (sim, p->pos, -1);
  if (cell == NULL) return force;

  force = compute_inertial_force (ctx, p);

  gdouble fluid_rho = fluid_rho_at (ctx, cell);

  gdouble cm = 0.5;
  if (ctx->par->cm_const == ctx->par->cm_const)
    cm = ctx->par->cm_const;

  for (c = 0; c < FTT_DIMENSION; c++)
    (&force.x)[c] *= cm;

  p->mass += fluid_rho*p->volume*cm;

  return force;
}

static OraForceFunc force_func (int kind)
{
  switch (kind) {
  case ORA_FORCE_INERTIAL: return compute_inertial_force;
  case ORA_FORCE_ADDEDMASS: return compute_addedmass_force;
  case ORA_FORCE_DRAG: return compute_drag_force;
  case ORA_FORCE_LIFT: return compute_lift_force;
  case ORA_FORCE_BUOY: return compute_buoyancy_force;
  }
  g_assert_not_reached ();
  return NULL;
}

/* modules/particulatecommon.c:737-751 */
static void compute_forces (OraCtx * ctx, OraForceFunc f, OraParticulate * p)
{
  FttComponent c;
  FttVector new_force = (* f) (ctx, p);
  FttVector total_force = p->force;

  for ( c = 0 ; c < FTT_DIMENSION; c++)
    (&total_force.x)[c] = (&new_force.x)[c]*p->volume + (&p->force.x)[c];

#if FTT_2D
    (&total_force.x)[2] = 0.;
#endif

  p->force = total_force;
}

/* modules/particulatecommon.c:753-765 */
static void compute_forces_onfluid (OraCtx * ctx, int kind, OraParticulate * p)
{
  FttComponent c;
  if (kind != ORA_FORCE_BUOY) {
    FttVector new_force = (* force_func (kind)) (ctx, p);
    FttVector total_force = p->force;

    for ( c = 0 ; c < FTT_DIMENSION; c++)
      (&total_force.x)[c] = (&new_force.x)[c]*p->volume + (&p->force.x)[c];

    p->force = total_force;
  }
}

static volatile double ora_sink;

/* modules/particulatecommon.c:768-842 (the branch with forces attached) */
static void particulate_event (OraCtx * ctx, OraParticulate * p)
{
  OraSim * sim = ctx->sim;
  const OraStepParams * par = ctx->par;
  FttComponent c;
  int k;

  FttVector pos = p->pos;
  p->pos_old = pos;

  /* Compute forces */
  for (c = 0; c < 3; c++)
    (&p->force.x)[c] = 0.;

  for (k = 0; k < par->n_forces; k++)
    compute_forces (ctx, force_func (par->force[k]), p);

  if (par->pattern == 0) {
    /* :821-826 -- vliq is computed and never used; kept for the call pattern */
    GfsVariable * u[3] = { &sim->var[0], &sim->var[1], &sim->var[FTT_DIMENSION > 2 ? 2 : 1] };
    FttCell * cell = domain_locate (sim, p->pos, -1);
    gdouble vliq = 0.;
    if (cell)
      for (c = 0; c < FTT_DIMENSION; c++)
	vliq += pow(gfs_interpolate (cell, p->pos, u[c]),2.);
    vliq = sqrt(vliq);
    ora_sink = vliq;
  }

  for (c = 0; c < FTT_DIMENSION; c++) {
    (&pos.x)[c] +=
      (&p->vel.x)[c]*par->dt/2.;
    (&p->vel.x)[c] +=
      (&p->force.x)[c]*par->dt/p->mass;
    (&pos.x)[c] +=
      (&p->vel.x)[c]*par->dt/2.;
  }

  p->pos = pos;
}

/* The fused variant used as the "generous" CPU baseline: one locate, one set
 * of interpolations, one vorticity evaluation shared by all forces.  Same
 * arithmetic per force as above. */
static void particulate_event_fused (OraCtx * ctx, OraParticulate * p)
{
  OraSim * sim = ctx->sim;
  const OraStepParams * par = ctx->par;
  FttComponent c;
  int k;
  FttVector pos = p->pos;
  p->pos_old = pos;
  for (c = 0; c < 3; c++)
    (&p->force.x)[c] = 0.;

  FttCell * cell = domain_locate (sim, p->pos, -1);
  if (cell) {
    GfsVariable * u[3] = { &sim->var[0], &sim->var[1], &sim->var[FTT_DIMENSION > 2 ? 2 : 1] };
    gdouble fluid_rho = fluid_rho_at (ctx, cell);
    gdouble viscosity = viscosity_at (ctx, cell);
    FttVector fluid_vel = { 0., 0., 0., 0. };
    for (c = 0; c < FTT_DIMENSION; c++)
      (&fluid_vel.x)[c] = gfs_interpolate (cell, p->pos, u[c]);
    FttVector relative_vel = subs_fttvectors (&fluid_vel, &p->vel);
    for (k = 0; k < par->n_forces; k++) {
      FttVector f = { 0., 0., 0., 0. };
      if (par->force[k] == ORA_FORCE_DRAG) {
	gdouble dia = 2.*pow(3.0*(p->volume)/4.0/M_PI, 1./3.);
#if !FTT_2D
	gdouble nrm = sqrt (relative_vel.x*relative_vel.x + relative_vel.y*relative_vel.y +
			    relative_vel.z*relative_vel.z);
#else
	gdouble nrm = sqrt (relative_vel.x*relative_vel.x + relative_vel.y*relative_vel.y);
#endif
	if (viscosity != 0) {
	  gdouble Re = nrm*dia*fluid_rho/viscosity, cd = 0.;
	  gboolean zero = FALSE;
	  if (par->cd_const == par->cd_const)
	    cd = par->cd_const;
	  else if (Re < 1e-8)
	    zero = TRUE;
	  else if (Re < 50.0)
	    cd = 16.*(1. + 0.15*pow(Re,0.5))/Re;
	  else
	    cd = 48.*(1. - 2.21/pow(Re,0.5))/Re;
	  if (!zero)
	    for (c = 0; c < FTT_DIMENSION; c++)
	      (&f.x)[c] += 3./(4.*dia)*cd*nrm*(&relative_vel.x)[c]*fluid_rho;
	}
      }
      else if (par->force[k] == ORA_FORCE_LIFT) {
	FttVector w = { 0., 0., 0., 0. };
	gdouble cl = par->cl_const == par->cl_const ? par->cl_const : 0.5;
	vorticity_vector (cell, u, &w);
#if FTT_2D
	f.x = fluid_rho*cl*relative_vel.y*w.z;
	f.y = -fluid_rho*cl*relative_vel.x*w.z;
#else
	f.x = fluid_rho*cl*(relative_vel.y*w.z - relative_vel.z*w.y);
	f.y = fluid_rho*cl*(relative_vel.z*w.x - relative_vel.x*w.z);
	f.z = fluid_rho*cl*(relative_vel.x*w.y - relative_vel.y*w.x);
#endif
      }
      else if (par->force[k] == ORA_FORCE_BUOY)
	for (c = 0; c < FTT_DIMENSION; c++)
	  (&f.x)[c] += (p->mass/p->volume - fluid_rho)*par->g[c];
      for (c = 0; c < FTT_DIMENSION; c++)
	(&p->force.x)[c] = (&f.x)[c]*p->volume + (&p->force.x)[c];
    }
  }
  for (c = 0; c < FTT_DIMENSION; c++) {
    (&pos.x)[c] += (&p->vel.x)[c]*par->dt/2.;
    (&p->vel.x)[c] += (&p->force.x)[c]*par->dt/p->mass;
    (&pos.x)[c] += (&p->vel.x)[c]*par->dt/2.;
  }
  p->pos = pos;
}

/* A GfsParticleList stand-in: one heap object per particle on a singly linked
 * list (src/event.h:373-380 keeps a GSList of GtsObjects). */
typedef struct {
  long n;
  OraParticulate ** p;    /* list order */
} OraList;

OraList * ora_list_new (long n, const double * x, const double * y, const double * z,
			const double * vx, const double * vy, const double * vz,
			const double * mass, const double * volume)
{
  OraList * l = g_malloc0 (sizeof (OraList));
  long i;
  l->n = n;
  l->p = g_malloc0 (sizeof (OraParticulate *)*(n ? n : 1));
  for (i = 0; i < n; i++) {
    OraParticulate * p = g_malloc0 (sizeof (OraParticulate));
    p->id = i + 1;
    p->pos.x = x[i]; p->pos.y = y[i]; p->pos.z = z ? z[i] : 0.;
    p->vel.x = vx[i]; p->vel.y = vy[i]; p->vel.z = vz ? vz[i] : 0.;
    p->mass = mass[i]; p->volume = volume[i];
    l->p[i] = p;
  }
  return l;
}

void ora_list_destroy (OraList * l)
{
  long i;
  for (i = 0; i < l->n; i++)
    free (l->p[i]);
  free (l->p);
  free (l);
}

long ora_list_size (OraList * l) { return l->n; }

void ora_list_get_mass (OraList * l, double * mass)
{
  long i;
  for (i = 0; i < l->n; i++)
    mass[i] = l->p[i]->mass;
}

void ora_list_get (OraList * l, double * x, double * y, double * z,
		   double * vx, double * vy, double * vz,
		   double * fx, double * fy, double * fz, unsigned * id)
{
  long i;
  for (i = 0; i < l->n; i++) {
    OraParticulate * p = l->p[i];
    x[i] = p->pos.x; y[i] = p->pos.y; if (z) z[i] = p->pos.z;
    vx[i] = p->vel.x; vy[i] = p->vel.y; if (vz) vz[i] = p->vel.z;
    if (fx) { fx[i] = p->force.x; fy[i] = p->force.y; fz[i] = p->force.z; }
    if (id) id[i] = p->id;
  }
}

/* The particle block of gfs_event_list_write (src/event.c:2508-2523): per list member four
 * spaces, the object's write method -- gfs_particle_write (src/particle.c:86-98: class name,
 * " %d %g %g %g" id and position) then gfs_particulate_write
 * (modules/particulatecommon.c:910-926: " %g %g %g %g %g" mass, volume*L^dim, velocity and
 * " %g %g %g" force) -- and a newline.  No coordinate mapping (GfsMap list empty). */
int ora_list_write (OraList * l, const char * path, double L)
{
  FILE * fp = fopen (path, "w");
  long i;
  if (!fp) return -1;
  for (i = 0; i < l->n; i++) {
    OraParticulate * p = l->p[i];
    fputs ("    ", fp);
    fprintf (fp, "%s", "GfsParticulate");
    fprintf (fp, " %d %g %g %g", p->id, p->pos.x, p->pos.y, p->pos.z);
    fprintf (fp, " %g %g %g %g %g", p->mass, p->volume*pow (L, FTT_DIMENSION),
	     p->vel.x, p->vel.y, p->vel.z);
    fprintf (fp, " %g %g %g", p->force.x, p->force.y, p->force.z);
    fputc ('\n', fp);
  }
  return fclose (fp);
}

/* modules/particulatecommon.c:955-969 + 980-987: drop particles whose
 * gfs_domain_locate is NULL.  Returns the number removed. */
long ora_list_cull (OraSim * sim, OraList * l)
{
  long i, m = 0;
  for (i = 0; i < l->n; i++) {
    if (domain_locate (sim, l->p[i]->pos, -1) == NULL)
      free (l->p[i]);
    else
      l->p[m++] = l->p[i];
  }
  i = l->n - m;
  l->n = m;
  return i;
}

/* one gfs_event_list_event pass over the list (src/event.c:2430-2439);
 * nthreads <= 1 reproduces the reference's serial loop. */
void ora_list_step (OraSim * sim, OraList * l, const OraStepParams * par, int nthreads)
{
  OraCtx ctx = { sim, par };
  long i;
  if (nthreads <= 1) {
    for (i = 0; i < l->n; i++) {
      if (par->pattern == 0)
	particulate_event (&ctx, l->p[i]);
      else
	particulate_event_fused (&ctx, l->p[i]);
    }
    return;
  }
#ifdef _OPENMP
# pragma omp parallel for schedule(static) num_threads(nthreads)
#endif
  for (i = 0; i < l->n; i++) {
    if (par->pattern == 0)
      particulate_event (&ctx, l->p[i]);
    else
      particulate_event_fused (&ctx, l->p[i]);
  }
}

/* ------------------------------------------------------------------ */
/* particle boundary conditions: modules/particulatecommon.c:3049-3397   */

/* :3058-3148 */
static gboolean check_intersetion(FttVector cellpos, FttVector p0, FttVector p1,
				  FttDirection *dstore, gdouble size)
{
  gdouble t;
  FttDirection d;

  for(d = 0; d < FTT_NEIGHBORS; d++){
    gdouble normal = ((gdouble) FTT_OPPOSITE_DIRECTION(d) - (gdouble)d);
#if FTT_2D
    switch(d/2){
    case 0:
      if((p1.x - p0.x)!=0 && normal*(p1.x-p0.x) > 0){
	t = (cellpos.x + normal*size*0.5 - p0.x)/(p1.x - p0.x);
	gdouble py = p0.y + t*(p1.y - p0.y);
	if((py - cellpos.y + size*0.5)*(py - cellpos.y - size*0.5) <= 0  &&  t*(t-1)<= 0 ){
	  *dstore = d;
	  return TRUE;
	}
      }
      break;
    case 1:
      if((p1.y - p0.y)!=0 && normal*(p1.y-p0.y) > 0){
	t = (cellpos.y + normal*size*0.5- p0.y)/(p1.y - p0.y);
	gdouble px = p0.x + t*(p1.x - p0.x);
	if((px - cellpos.x + size*0.5)*(px - cellpos.x - size*0.5) <= 0  &&  t*(t-1) <= 0){
	  *dstore = d;
	  return TRUE;
	}
      }
      break;
    }
#else
    switch(d/2){
    case 0:
      if((p1.x - p0.x)!=0 && normal*(p1.x-p0.x) > 0){
	t = (cellpos.x + normal*size*0.5 - p0.x)/(p1.x - p0.x);
	gdouble py = p0.y + t*(p1.y - p0.y);
	gdouble pz = p0.z + t*(p1.z - p0.z);
	if((py - cellpos.y + size*0.5)*(py - cellpos.y - size*0.5) <= 0  &&
	   (pz - cellpos.z + size*0.5)*(pz - cellpos.z - size*0.5) <= 0
	   &&  t*(t-1)<= 0 )
	  {
	  *dstore = d;
	  return TRUE;
	}
      }
      break;
    case 1:
      if((p1.y - p0.y)!=0 && normal*(p1.y-p0.y) > 0){
	t = (cellpos.y + normal*size*0.5- p0.y)/(p1.y - p0.y);
	gdouble px = p0.x + t*(p1.x - p0.x);
	gdouble pz = p0.z + t*(p1.z - p0.z);
	if((px - cellpos.x + size*0.5)*(px - cellpos.x - size*0.5) <= 0  &&
	   (pz - cellpos.z + size*0.5)*(pz - cellpos.z - size*0.5) <= 0
	   &&  t*(t-1)<= 0 )
	  {
	  *dstore = d;
	  return TRUE;
	}
      }
      break;
    case 2:
      if((p1.z - p0.z)!=0 && normal*(p1.z-p0.z) > 0){
	t = (cellpos.z + normal*size*0.5- p0.z)/(p1.z - p0.z);
	gdouble px = p0.x + t*(p1.x - p0.x);
	gdouble py = p0.y + t*(p1.y - p0.y);
	if((px - cellpos.x + size*0.5)*(px - cellpos.x - size*0.5) <= 0  &&
	   (py - cellpos.y + size*0.5)*(py - cellpos.y - size*0.5) <= 0
	   &&  t*(t-1)<= 0 )
	  {
	  *dstore = d;
	  return TRUE;
	}
      }
      break;
    }
#endif
  }
  return FALSE;
}

/* :3151-3186.  Returns NULL when the intersection search fails (the reference
 * then reads an uninitialised direction; the oracle reports "no face"). */
static FttCell * boundarycell (OraSim * sim, OraParticulate * p, FttDirection * dstore)
{
  FttCell * cell = domain_locate (sim, p->pos_old, -1);
  if (cell == NULL) return NULL;

  FttVector cellpos;
  gdouble size;
  ftt_cell_pos(cell, &cellpos);
  size = ftt_cell_size(cell);

  if (!check_intersetion(cellpos, p->pos_old, p->pos, dstore, size))
    return NULL;
  FttCellFace face = ftt_cell_face(cell, *dstore);

  if(!face.neighbor)
    return cell;

  while(!GFS_CELL_IS_BOUNDARY(face.neighbor)) {
    cell = face.neighbor;
    ftt_cell_pos(cell, &cellpos);
    size = ftt_cell_size(cell);

    if (!check_intersetion(cellpos, p->pos_old, p->pos, dstore, size))
      return NULL;
    face = ftt_cell_face(cell, *dstore);

    if(!face.neighbor)
      return cell;
  };
  return cell;
}

/* gfs_particle_bc (:3375-3395) for the one-box domains of the oracle:
 * particles whose gfs_domain_locate is NULL leave the list; those that left
 * through a side whose GfsBoundary is periodic (bit `side` of periodic_mask;
 * the matching box is the box itself) are wrapped by periodic_bc_particle
 * (:3189-3214) and re-appended at the END of the list, the others are dropped.
 * Returns the number dropped. */
long ora_list_bc (OraSim * sim, OraList * l, unsigned periodic_mask)
{
  long i, m = 0, nw = 0, dropped = 0;
  OraParticulate ** wrapped = g_malloc0 (sizeof (OraParticulate *)*(l->n ? l->n : 1));
  for (i = 0; i < l->n; i++) {
    OraParticulate * p = l->p[i];
    if (domain_locate (sim, p->pos, -1) != NULL) {
      l->p[m++] = p;
      continue;
    }
    FttDirection d = 0;
    FttCell * cell = boundarycell (sim, p, &d);
    if (cell && sim->nbox == 1 && sim->broot[0][d] && (periodic_mask >> d & 1)) {
      FttVector box_face, box_face_nbr;
      ftt_cell_pos(sim->root, &box_face);
      ftt_cell_pos(sim->root, &box_face_nbr);
      gdouble size = ftt_cell_size(sim->root);
      gdouble normal = (gdouble)FTT_OPPOSITE_DIRECTION(d) - (gdouble) d;
      (&box_face.x)[d/2] += (gdouble)normal * size/2.;
      (&box_face_nbr.x)[d/2] -= (gdouble)normal * size/2.;
      gdouble tolerance = size/1.e8;
      gdouble distance = ((&p->pos.x)[d/2] - (&box_face.x)[d/2])*normal;
      (&p->pos.x)[d/2] = (&box_face_nbr.x)[d/2] + distance + normal*tolerance;
      (&p->pos_old.x)[d/2] = (&p->pos.x)[d/2];
      wrapped[nw++] = p;
    }
    else {
      free (p);
      dropped++;
    }
  }
  for (i = 0; i < nw; i++)
    l->p[m++] = wrapped[i];
  free (wrapped);
  l->n = m;
  return dropped;
}

/* ------------------------------------------------------------------ */
/* deposition                                                           */

/* modules/particulatecommon.c:1929-1957 with the default voidfraction_func:
 * reset on leaves is done by the caller (values arrive through ora_set_values);
 * serial scatter in list order. */
void ora_deposit_volume (OraSim * sim, OraList * l, int ivar)
{
  long i;
  for (i = 0; i < l->n; i++) {
    FttCell * cellpart = domain_locate (sim, l->p[i]->pos, -1);
    if (cellpart)
      GFS_VALUEI (cellpart, ivar) += l->p[i]->volume/ftt_cell_volume (cellpart);
  }
}

/* modules/particulatecommon.c:2177-2228 in the single-cell (nearest-cell)
 * limit of diffuse_force (:2158-2175): forces recomputed without buoyancy
 * (compute_forces_onfluid), then u_c[cell] -= F_c/rho/V_cell. */
void ora_deposit_force (OraSim * sim, OraList * l, const OraStepParams * par, int ivar0)
{
  OraCtx ctx = { sim, par };
  long i;
  int k;
  FttComponent c;
  for (i = 0; i < l->n; i++) {
    OraParticulate * p = l->p[i];
    for (c = 0; c < 3; c++)
      (&p->force.x)[c] = 0.;
    for (k = 0; k < par->n_forces; k++)
      compute_forces_onfluid (&ctx, par->force[k], p);
  }
  for (i = 0; i < l->n; i++) {
    OraParticulate * p = l->p[i];
    FttCell * cell = domain_locate (sim, p->pos, -1);
    if (cell) {
      gdouble cellvol = gfs_cell_volume (cell, &sim->domain);
      gdouble liq_rho = fluid_rho_at (&ctx, cell);
      for (c = 0; c < FTT_DIMENSION; c++)
	GFS_VALUEI (cell, ivar0 + c) -= (&p->force.x)[c]/liq_rho/cellvol;
    }
  }
}

/* ------------------------------------------------------------------ */
/* GfsSourceParticulate with its smoothing kernel:                       */
/* modules/particulatecommon.c:2087-2228                                 */

/* The user's `kernel = ...' GfsFunction, restricted to the closed forms the
 * device evaluates (gfs_function_spatial_value of a spatial function of the
 * normalised offset, src/utils.c:1476-1494, without coordinate mapping). */
enum { ORA_KERNEL_CONSTANT = 0, ORA_KERNEL_GAUSSIAN = 1, ORA_KERNEL_COMPACT = 2 };

typedef struct {
  int kind;
  double a, b;         /* CONSTANT: a;  GAUSSIAN: a*exp(-b*r2);  COMPACT: a*(1 - b*r2)^p, 0 where b*r2 >= 1 */
  int p;
  int flags;           /* bit 0: z offset taken from the cell centre (see distance_normalization below) */
} OraKernel;

static gdouble kernel_function_value (const OraKernel * k, const FttVector * q)
{
  gdouble r2 = q->x*q->x + q->y*q->y + q->z*q->z;
  switch (k->kind) {
  case ORA_KERNEL_CONSTANT: return k->a;
  case ORA_KERNEL_GAUSSIAN: return k->a*exp (- k->b*r2);
  case ORA_KERNEL_COMPACT: {
    gdouble t = 1. - k->b*r2, v = k->a;
    int i;
    if (t <= 0.) return 0.;
    for (i = 0; i < k->p; i++) v *= t;
    return v;
  }
  }
  g_assert_not_reached ();
  return 0.;
}

/* modules/particulatecommon.c:2087-2098.  NB the reference zeroes pos1->z
 * BEFORE using it in 3D, so the z offset is (0 - z_p)/rb whatever the cell;
 * kept (flags bit 0 selects the evidently intended (z_c - z_p)/rb). */
static void distance_normalization (FttVector * pos1, OraParticulate * p, int flags)
{
  gdouble rb = pow (3.*p->volume/(4.*M_PI), 1./3.);
  FttVector * pos2 = &p->pos;
  gdouble zc = pos1->z;
  pos1->x = (pos1->x - pos2->x)/rb;
  pos1->y = (pos1->y - pos2->y)/rb;
  pos1->z = 0.;
#if !FTT_2D
  if (flags & 1) pos1->z = zc;
  pos1->z = (pos1->z - pos2->z)/rb;
#endif
}

typedef struct {
  gdouble correction, volume;
  OraParticulate * p;
  const OraKernel * kernel_function;
  int ivar0;
  OraCtx * ctx;
} OraKernelData;

/* :2108-2119 */
static void kernel_volume (FttCell * cell, OraKernelData * kd)
{
  gdouble cellvol = gfs_cell_volume (cell, &kd->ctx->sim->domain);
  FttVector pos;

  kd->volume += cellvol;
  ftt_cell_pos (cell, &pos);
  distance_normalization (&pos, kd->p, kd->kernel_function->flags);
  kd->correction += kernel_function_value (kd->kernel_function, &pos)*cellvol;
}

typedef struct {
  FttVector * pos;
  gdouble distance;
} OraCondData;

/* :2126-2156 */
static gboolean cond_kernel (FttCell * cell, gpointer data)
{
  OraCondData * p = data;
  FttVector pos;
  gdouble radeq, size;

  ftt_cell_pos (cell, &pos);
  size = ftt_cell_size (cell)/2.;
#if FTT_2D
  radeq = size*sqrt (2.);
#else
  radeq = size*sqrt (3.);
#endif
  if (ftt_vector_distance (&pos, p->pos) - radeq <= p->distance)
    return TRUE;
  if (p->pos->x > pos.x + size || p->pos->x < pos.x - size ||
      p->pos->y > pos.y + size || p->pos->y < pos.y - size
#if !FTT_2D
      || p->pos->z > pos.z + size || p->pos->z < pos.z - size
#endif
      )
    return FALSE;
  return TRUE;
}

/* :2158-2175 */
static void diffuse_force (FttCell * cell, OraKernelData * kd)
{
  FttVector pos;
  FttComponent c;
  gdouble cellvol, liq_rho;

  ftt_cell_pos (cell, &pos);
  distance_normalization (&pos, kd->p, kd->kernel_function->flags);
  cellvol = gfs_cell_volume (cell, &kd->ctx->sim->domain);
  liq_rho = fluid_rho_at (kd->ctx, cell);
  if (kd->correction > 1.e-10)
    for (c = 0; c < FTT_DIMENSION; c++)
      GFS_VALUEI (cell, kd->ivar0 + c) -= (&kd->p->force.x)[c]/liq_rho/cellvol*
	kernel_function_value (kd->kernel_function, &pos)/kd->correction;
}

/* source_particulate_event, :2177-2228: forces recomputed without buoyancy for
 * every particle, then per particle two conditional traversals of every GfsBox
 * tree (gfs_domain_cell_traverse_condition, src/domain.c:1516-1574 ->
 * ftt_cell_traverse_condition, the reference's own object code): the
 * normalisation, then the deposit.  `rkernel' is used as an absolute distance
 * (:2211; influencerad :2210 is computed and never used).  The reset of the
 * target variables is done by the caller.  correction_out/volume_out (may be
 * NULL) receive the per-particle normalisation for the tests. */
void ora_deposit_force_smoothed (OraSim * sim, OraList * l, const OraStepParams * par, int ivar0,
				 double rkernel, const OraKernel * kernel,
				 double * correction_out, double * volume_out)
{
  OraCtx ctx = { sim, par };
  long i;
  int k, b;
  FttComponent c;
  for (i = 0; i < l->n; i++) {
    OraParticulate * p = l->p[i];
    for (c = 0; c < 3; c++)
      (&p->force.x)[c] = 0.;
    for (k = 0; k < par->n_forces; k++)
      compute_forces_onfluid (&ctx, par->force[k], p);
  }
  for (i = 0; i < l->n; i++) {
    OraParticulate * p = l->p[i];
    OraCondData cd = { &p->pos, rkernel };
    OraKernelData kd = { 0., 0., p, kernel, ivar0, &ctx };
    for (b = 0; b < sim->nbox; b++)
      ftt_cell_traverse_condition (sim->box[b], FTT_PRE_ORDER, FTT_TRAVERSE_LEAFS, -1,
				   (FttCellTraverseFunc) kernel_volume, &kd, cond_kernel, &cd);
    kd.correction /= kd.volume;
    if (correction_out) correction_out[i] = kd.correction;
    if (volume_out) volume_out[i] = kd.volume;
    for (b = 0; b < sim->nbox; b++)
      ftt_cell_traverse_condition (sim->box[b], FTT_PRE_ORDER, FTT_TRAVERSE_LEAFS, -1,
				   (FttCellTraverseFunc) diffuse_force, &kd, cond_kernel, &cd);
  }
}

/* ------------------------------------------------------------------ */
/* passive tracers: src/particle.c:31-44 -> src/domain.c:2764-2788      */

void ora_advect_points (OraSim * sim, long n, double * x, double * y, double * z, double dt)
{
  long i;
  GfsVariable * u[3] = { &sim->var[0], &sim->var[1], &sim->var[FTT_DIMENSION > 2 ? 2 : 1] };
#pragma omp parallel for schedule(static)
  for (i = 0; i < n; i++) {
    FttVector p = { x[i], y[i], z ? z[i] : 0., 0. };
    FttVector p0, p1;
    FttCell * cell;
    FttComponent c;
    p0 = p1 = p;
    cell = domain_locate (sim, p0, -1);
    if (cell == NULL)
      continue;
    for (c = 0; c < FTT_DIMENSION; c++)
      (&p1.x)[c] += dt*gfs_interpolate (cell, p0, u[c])/2.;
    cell = domain_locate (sim, p1, -1);
    if (cell == NULL)
      continue;
    for (c = 0; c < FTT_DIMENSION; c++)
      (&p.x)[c] += dt*gfs_interpolate (cell, p1, u[c]);
    x[i] = p.x; y[i] = p.y; if (z) z[i] = p.z;
  }
}

int ora_max_threads (void)
{
#ifdef _OPENMP
  return omp_get_max_threads ();
#else
  return 1;
#endif
}

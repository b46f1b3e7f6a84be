/* ftt_bridge.c -- the FTT flattening pass: mirrors a live Gerris FttCell
 * pointer tree (src/ftt.h:134-159) into a gfsb200 flat tree, and gathers /
 * scatters cell variables between the two.
 *
 * Compiled once per dimension (-DFTT_2D=1 or not) against the Gerris headers,
 * exactly like the reference builds libparticulates2D/3D
 * (modules/Makefile.am:154-164).  Only struct members and inline helpers of
 * ftt.h are used, so the object has no unresolved libgfs symbols and can sit
 * inside the drop-in module or be loaded on its own.
 *
 * Called after every mesh adaptation (gfs_simulation_adapt, src/adaptive.c:1445)
 * and never on the per-step path; the per-step work is the value gather only.
 */
#include <stdlib.h>
#include <string.h>
#include "ftt.h"
#include "gfsb200.h"
#include "gfsb200_ftt.h"

/* GFS_FLAG_BOUNDARY (src/fluid.h:63): ghost cells of a GfsBoundary tree */
#define BRIDGE_FLAG_BOUNDARY (1u << (FTT_FLAG_USER + 1))

/* GfsSolidVector, src/fluid.h:54-59, and where GfsStateVector (src/fluid.h:44-52) keeps the
   pointer to it: right after f[FTT_NEIGHBORS] (two doubles per face) */
struct bridge_solid {
  double s[FTT_NEIGHBORS];
  double a, fv;
  FttCell * merged;
  FttVector cm, ca, v;
};

static const struct bridge_solid * state_solid (const FttCell * c)
{
  return c->data ? *(const struct bridge_solid * const *)
    ((const char *) c->data + FTT_NEIGHBORS*2*sizeof (double)) : NULL;
}

struct gfsb200_ftt_map {
  int32_t n_cells;
  FttCell ** cell;            /* flat index -> FttCell */
  char ** data;               /* flat index -> cell->data (NULL: destroyed cell, or no data), or NULL: not cached */
  unsigned data_generation;
};

static const char * bridge_error = "";
const char * gfsb200_ftt_last_error (void) { return bridge_error; }

void gfsb200_ftt_map_free (gfsb200_ftt_map * m)
{
  if (!m) return;
  free (m->cell);
  free (m->data);
  free (m);
}

int32_t gfsb200_ftt_map_size (const gfsb200_ftt_map * m) { return m->n_cells; }
void * gfsb200_ftt_map_cell (const gfsb200_ftt_map * m, int32_t i) { return m->cell[i]; }
void * const * gfsb200_ftt_map_cells (const gfsb200_ftt_map * m) { return (void * const *) m->cell; }

int gfsb200_ftt_flatten (int n_roots, void * const * roots_, const int * is_box,
			 gfsb200_tree ** tree_out, gfsb200_ftt_map ** map_out)
{
  FttCell * const * roots = (FttCell * const *) roots_;
  if (n_roots <= 0 || !roots || !is_box || !tree_out || !map_out) {
    bridge_error = "flatten: bad argument";
    return GFSB200_ERR_ARG;
  }
  gfsb200_tree * t = gfsb200_tree_new (FTT_DIMENSION);
  if (!t) { bridge_error = gfsb200_last_error (); return GFSB200_ERR_NOMEM; }

  /* roots: GfsBox roots first (caller's order within each class is kept) */
  int * order = malloc (sizeof (int)*n_roots);
  int no = 0;
  for (int pass = 1; pass >= 0; pass--)
    for (int r = 0; r < n_roots; r++)
      if ((is_box[r] != 0) == pass)
	order[no++] = r;

  size_t cap = 1024, n = 0, n_mixed = 0;
  FttCell ** cells = malloc (cap*sizeof (FttCell *));
  int rc = GFSB200_OK;
  for (int j = 0; j < n_roots && rc >= 0; j++) {
    FttCell * root = roots[order[j]];
    if (!root || !FTT_CELL_IS_ROOT (root)) {
      bridge_error = "flatten: not a root cell";
      rc = GFSB200_ERR_ARG;
      break;
    }
    double pos[3] = { FTT_ROOT_CELL (root)->pos.x, FTT_ROOT_CELL (root)->pos.y,
		      FTT_ROOT_CELL (root)->pos.z };
    rc = gfsb200_tree_add_root (t, pos, FTT_ROOT_CELL (root)->level, is_box[order[j]]);
    cells[n++] = root;
  }
  /* root adjacency (FttRootCell.neighbors, src/ftt.h:142-149) */
  for (int j = 0; j < n_roots && rc >= 0; j++)
    for (int d = 0; d < FTT_NEIGHBORS; d++) {
      FttCell * nb = FTT_ROOT_CELL (cells[j])->neighbors.c[d];
      if (nb)
	for (int k = 0; k < n_roots; k++)
	  if (cells[k] == nb)
	    rc = gfsb200_tree_link_roots (t, j, d, k);
    }
  /* breadth-first mirror: cells[] doubles as the queue; a split appends the
     2^dim children in FTT child order, destroyed ones included, so growth
     order == flat order and the map needs no permutation */
  for (size_t head = 0; head < n && rc >= 0; head++) {
    FttCell * c = cells[head];
    if (FTT_CELL_IS_DESTROYED (c))
      continue;
    if (state_solid (c))
      n_mixed++;
    if (FTT_CELL_IS_LEAF (c))
      continue;
    struct _FttOct * oct = c->children;
    unsigned destroyed = 0, flags = 0;
    for (int k = 0; k < FTT_CELLS; k++) {
      if (FTT_CELL_IS_DESTROYED (&oct->cell[k]))
	destroyed |= 1u << k;
      if (oct->cell[k].flags & BRIDGE_FLAG_BOUNDARY)
	flags |= GFSB200_CELL_BOUNDARY;
    }
    int c0 = gfsb200_tree_split (t, (int) head, destroyed, flags);
    if (c0 < 0) { rc = c0; break; }
    if ((size_t) c0 != n) { bridge_error = "flatten: internal order error"; rc = GFSB200_ERR_STATE; break; }
    if (n + FTT_CELLS > cap) {
      cap *= 2;
      cells = realloc (cells, cap*sizeof (FttCell *));
    }
    for (int k = 0; k < FTT_CELLS; k++)
      cells[n++] = &oct->cell[k];
  }
  free (order);
  if (rc >= 0) {
    int32_t * perm = malloc (n*sizeof (int32_t));
    rc = gfsb200_tree_finalize (t, perm);
    if (rc >= 0)
      for (size_t i = 0; i < n; i++)
	if ((size_t) perm[i] != i) { bridge_error = "flatten: finalize permuted cells"; rc = GFSB200_ERR_STATE; break; }
    free (perm);
  }
  /* mixed (solid-cut) cells: fluid fraction and centre of mass go with the tree */
  if (rc >= 0 && n_mixed)
    for (size_t i = 0; i < n && rc >= 0; i++) {
      const struct bridge_solid * sv = FTT_CELL_IS_DESTROYED (cells[i]) ? NULL : state_solid (cells[i]);
      if (sv) {
	const double cm[3] = { sv->cm.x, sv->cm.y, sv->cm.z };
	rc = gfsb200_tree_set_solid (t, (int) i, sv->a, cm, sv->s);
      }
    }
  if (rc < 0) {
    if (!*bridge_error) bridge_error = gfsb200_last_error ();
    gfsb200_tree_free (t);
    free (cells);
    return rc;
  }
  gfsb200_ftt_map * m = malloc (sizeof *m);
  m->n_cells = (int32_t) n;
  m->cell = cells;
  m->data = NULL;
  m->data_generation = 0;
  *tree_out = t;
  *map_out = m;
  return GFSB200_OK;
}

/* GFS_VALUEI (cell, i) = (&GFS_STATE (cell)->place_holder)[i]  (src/fluid.h:71-72).
 * `offset` is the byte offset of place_holder inside the cell's data block
 * (offsetof (GfsStateVector, place_holder)), passed in so that this file does
 * not need fluid.h/gts.h. */
int gfsb200_ftt_gather (const gfsb200_ftt_map * m, size_t offset, int var, double nodata, double * out)
{
  return gfsb200_ftt_gather_many (m, offset, 1, &var, &nodata, &out);
}

/* one pass over the cells for all nvar variables: the cost is the pointer chase
   cell -> data (one cache miss per cell), paid once instead of once per variable,
   and spread over the host cores */
int gfsb200_ftt_gather_many (const gfsb200_ftt_map * m, size_t offset, int nvar, const int * var,
			     const double * nodata, double * const * out)
{
  return gfsb200_ftt_gather_range (m, offset, 0, m ? m->n_cells : 0, nvar, var, nodata, out);
}

/* the same for the flat cells [first, last): out[k][i] is written for first <= i < last, so that
   the caller can ship one slice to the device while the next one is being gathered */
int gfsb200_ftt_gather_range (const gfsb200_ftt_map * m, size_t offset, int32_t first, int32_t last,
			      int nvar, const int * var, const double * nodata, double * const * out)
{
  if (!m || nvar < 0 || (nvar && (!var || !nodata || !out)) || first < 0 || last > m->n_cells || first > last) {
    bridge_error = "gather: bad argument";
    return GFSB200_ERR_ARG;
  }
  if (m->data) {
    /* cached data pointers: a sequential read of the pointer array, the data blocks prefetched
       a few cells ahead -- no FttCell / FttOct line is touched */
    char * const * data = m->data;
    const int32_t n = last;
#pragma omp parallel for schedule(static)
    for (int32_t i = first; i < n; i++) {
      if (i + 32 < n && data[i + 32])
	__builtin_prefetch (data[i + 32] + offset, 0, 0);
      const char * d = data[i];
      if (!d)
	for (int k = 0; k < nvar; k++)
	  out[k][i] = nodata[k];
      else {
	const double * v = (const double *) (d + offset);
	for (int k = 0; k < nvar; k++)
	  out[k][i] = v[var[k]];
      }
    }
    return GFSB200_OK;
  }
#pragma omp parallel for schedule(static)
  for (int32_t i = first; i < last; i++) {
    FttCell * c = m->cell[i];
    if (FTT_CELL_IS_DESTROYED (c) || !c->data)
      for (int k = 0; k < nvar; k++)
	out[k][i] = nodata[k];
    else {
      const double * v = (const double *) ((const char *) c->data + offset);
      for (int k = 0; k < nvar; k++)
	out[k][i] = v[var[k]];
    }
  }
  return GFSB200_OK;
}

int gfsb200_ftt_map_cache_data (gfsb200_ftt_map * m, unsigned generation)
{
  if (!m) {
    bridge_error = "cache_data: bad argument";
    return GFSB200_ERR_ARG;
  }
  if (m->data && m->data_generation == generation)
    return GFSB200_OK;
  if (!m->data && !(m->data = malloc ((m->n_cells ? m->n_cells : 1)*sizeof (char *)))) {
    bridge_error = "cache_data: out of memory";
    return GFSB200_ERR_NOMEM;
  }
#pragma omp parallel for schedule(static)
  for (int32_t i = 0; i < m->n_cells; i++) {
    FttCell * c = m->cell[i];
    m->data[i] = FTT_CELL_IS_DESTROYED (c) ? NULL : (char *) c->data;
  }
  m->data_generation = generation;
  return GFSB200_OK;
}

int gfsb200_ftt_scatter (const gfsb200_ftt_map * m, size_t offset, int var, int leaves_only,
			 const double * in)
{
#pragma omp parallel for schedule(static)
  for (int32_t i = 0; i < m->n_cells; i++) {
    FttCell * c = m->cell[i];
    /* leaves_only: the cells gfs_domain_cell_traverse (FTT_TRAVERSE_LEAFS) visits, i.e. the
       leaves of the GfsBox trees -- ghost cells keep their value until the next gfs_domain_bc */
    if (FTT_CELL_IS_DESTROYED (c) || !c->data ||
	(leaves_only && (!FTT_CELL_IS_LEAF (c) || (c->flags & BRIDGE_FLAG_BOUNDARY))))
      continue;
    ((double *) ((char *) c->data + offset))[var] = in[i];
  }
  return GFSB200_OK;
}

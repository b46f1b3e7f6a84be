/* gfsb200_internal.h -- definitions shared by the host-side sources */
#ifndef GFSB200_INTERNAL_H
#define GFSB200_INTERNAL_H

#include <stdarg.h>
#include <stdint.h>
#include "gfsb200.h"

#ifdef __cplusplus
extern "C" {
#endif

#define GFSB200_MAX_ROOTS 64
#define GFSB200_MAX_LEVEL 20     /* vertex lattice keys are 21 bits per axis */
#define GFSB200_MAX_STENCIL 29   /* GfsInterpolator capacity, src/fluid.h:263-272 */

struct gfsb200_tree {
  int dim, nchild, ndir;
  int finalized;
  int32_t n_cells, cap;
  /* per cell */
  int32_t * parent, * child0;
  uint8_t * level, * flags;
  double * pos;                  /* [n][3] */
  /* roots: root r is cell r */
  int n_roots, n_box_roots, root_level;
  uint8_t root_is_box[GFSB200_MAX_ROOTS];
  int32_t root_nb[GFSB200_MAX_ROOTS][6];
  int32_t periodic[GFSB200_MAX_ROOTS][6];   /* box root, side -> matching box root of a
					       GfsBoundaryPeriodic, or -1 */
  /* finalized */
  int32_t * neighbor;            /* [n][ndir] */
  int32_t * level_start;
  int n_levels, min_level, max_level, complete_level;
  int64_t n_leaves;
  double la_min[3], la_h;
  int32_t la_n[3], la_size;
  int32_t * la_slot;
  /* stencils */
  int32_t n_vertices;
  int32_t * vtx_off, * vtx_cell;
  double * vtx_w;
  int32_t * leaf_vtx;
  int lattice_level;             /* >= 0: vertices are numbered row-major on the lattice of that level */
  /* mixed (solid-cut) cells, GfsSolidVector of src/fluid.h:54-59: NULL when the tree has none */
  double * solid_a;              /* [n] fluid fraction a, 1 for a cell that is not mixed */
  double * solid_cm;             /* [n][3] centre of mass of the fluid part, NaN for a cell that is not mixed */
  double * solid_s;              /* [n][ndir] fluid fraction of each face, 1 for a cell that is not mixed */
};

int gfsb200_fail (int code, const char * fmt, ...);
int32_t gfsb200_tree_neighbor (const gfsb200_tree * t, int32_t cell, int d);

#ifdef __cplusplus
}
#endif

#endif

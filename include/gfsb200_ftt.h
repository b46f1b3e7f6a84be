/* gfsb200_ftt.h -- bridge between a live Gerris FttCell tree and the flat
 * tree of gfsb200.h.  Built per dimension as libgfsb200_ftt2D.so /
 * libgfsb200_ftt3D.so (the same -DFTT_2D=1 switch the reference uses,
 * modules/Makefile.am:154-164).
 *
 * Replaces nothing in the reference by itself: it is the pass that lets
 * gfs_particle_list_event (modules/particulatecommon.c:980-1015) hand the mesh
 * of src/ftt.h:134-159 and the GFS_VALUE cell data of src/fluid.h:44-72 to the
 * device path.  Pointers are passed as void* so that callers need no Gerris
 * headers.
 */
#ifndef GFSB200_FTT_H
#define GFSB200_FTT_H

#include <stddef.h>
#include <stdint.h>
#include "gfsb200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gfsb200_ftt_map gfsb200_ftt_map;

const char * gfsb200_ftt_last_error (void);

/* roots[r]: FttCell* root of a GfsBox (is_box[r] = 1, box->root) or of a
 * GfsBoundary (is_box[r] = 0, boundary->root).  On success *tree is a
 * FINALIZED flat tree (stencils not yet built) and *map gives, for every flat
 * cell index, the FttCell it mirrors. */
int gfsb200_ftt_flatten (int n_roots, void * const * roots, const int * is_box,
			 gfsb200_tree ** tree, gfsb200_ftt_map ** map);
void gfsb200_ftt_map_free (gfsb200_ftt_map * m);
int32_t gfsb200_ftt_map_size (const gfsb200_ftt_map * m);
void * gfsb200_ftt_map_cell (const gfsb200_ftt_map * m, int32_t i);
void * const * gfsb200_ftt_map_cells (const gfsb200_ftt_map * m);

/* out[i] = GFS_VALUEI (cell_i, var) for every flat cell (nodata for destroyed
 * cells); offset = offsetof (GfsStateVector, place_holder). */
int gfsb200_ftt_gather (const gfsb200_ftt_map * m, size_t offset, int var, double nodata,
			double * out);
/* the same for nvar variables in ONE pass over the cells, parallel over the host cores:
 * out[k][i] = GFS_VALUEI (cell_i, var[k]), nodata[k] for destroyed cells */
int gfsb200_ftt_gather_many (const gfsb200_ftt_map * m, size_t offset, int nvar, const int * var,
			     const double * nodata, double * const * out);
/* ... and for the flat cells [first, last) only: lets the caller ship one slice to the device
 * (gfsb200_upload_field_part) while the next one is being gathered */
int gfsb200_ftt_gather_range (const gfsb200_ftt_map * m, size_t offset, int32_t first, int32_t last,
			      int nvar, const int * var, const double * nodata, double * const * out);
/* The per-step gather is a pointer chase cell -> data (one cache miss per cell on top of the one
 * for the cell itself).  gfsb200_ftt_map_cache_data records, once per flatten, where the data block
 * of every cell lives; the gathers above then stream through that array with software prefetch and
 * never touch the FttCell / FttOct structures.  Gerris moves the data blocks only when it grows
 * them for a new variable (gfs_domain_alloc -> box_realloc, src/domain.c:3291-3305): the caller
 * passes any value that changes when that happens (domain->allocated->len) as `generation', and the
 * cache is rebuilt when it differs from the one it was built with. */
int gfsb200_ftt_map_cache_data (gfsb200_ftt_map * m, unsigned generation);

/* GFS_VALUEI (cell_i, var) = in[i]; leaves_only: only the leaves of the GfsBox trees, the
 * cells gfs_domain_cell_traverse (FTT_TRAVERSE_LEAFS) visits (ghost cells keep their value) */
int gfsb200_ftt_scatter (const gfsb200_ftt_map * m, size_t offset, int var, int leaves_only,
			 const double * in);

#ifdef __cplusplus
}
#endif

#endif

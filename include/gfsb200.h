/* gfsb200.h -- C-ABI of the B200-native Gerris particulate hot path.
 *
 * This is the drop-in boundary: plain C, plain pointers and sizes, int status
 * codes, no GLib/GTS and no torch types.  A replacement
 * libparticulates{2D,3D}.so (see INTEGRATION.md and
 * gerris-fft-particles_b200/host/particulates_b200.c) keeps the reference's
 * GtsObject classes and .gfs syntax and calls down into these entry points
 * from the same places the reference does its per-particle CPU work.
 *
 * Reference interfaces replaced (paths relative to the reference tree):
 *   tree flattening      FttCell/FttOct pointer tree        src/ftt.h:134-159
 *   point location       gfs_domain_locate                  src/domain.c:2623-2638
 *                        ftt_cell_locate                    src/ftt.c:1535-1574
 *                        GfsLocateArray                     src/domain.c:43-145
 *   interpolation        gfs_interpolate                    src/fluid.c:2697-2710
 *                        gfs_cell_corner_interpolator       src/fluid.c:3015-3069
 *   vorticity            vorticity_vector                   modules/particulatecommon.c:142-164
 *                        gfs_center_gradient                src/fluid.c:434-475
 *   forces               compute_{drag,lift,buoyancy}_force modules/particulatecommon.c:423-655
 *   integrator           gfs_particulate_event              modules/particulatecommon.c:768-842
 *   list driver          gfs_particle_list_event            modules/particulatecommon.c:980-1015
 *   deposition           particulate_field_event            modules/particulatecommon.c:1934-1957
 *                        source_particulate_event           modules/particulatecommon.c:2177-2228
 *                          (nearest cell, and with the smoothing kernel :2087-2175)
 *   tracer advection     gfs_domain_advect_point            src/domain.c:2764-2788
 *   output at points     gfs_output_location_event          src/output.c:1153-1212
 *   particle text/dump   gfs_particle_write                 src/particle.c:86-98
 *                        gfs_particulate_write              modules/particulatecommon.c:910-926
 *
 * All functions returning int return GFSB200_OK (0) or a negative error code;
 * gfsb200_last_error() gives the message of the last failure on the calling
 * thread.  There is no CPU fallback: every compute entry point fails with
 * GFSB200_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef GFSB200_H
#define GFSB200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GFSB200_OK             0
#define GFSB200_ERR_ARG       -1
#define GFSB200_ERR_STATE     -2
#define GFSB200_ERR_NOMEM     -3
#define GFSB200_ERR_CUDA      -4
#define GFSB200_ERR_UNSUPPORTED -5

/* cell flags (flat tree) */
#define GFSB200_CELL_DESTROYED 1u   /* FTT_FLAG_DESTROYED   src/ftt.h:113-118 */
#define GFSB200_CELL_BOUNDARY  2u   /* GFS_FLAG_BOUNDARY    src/fluid.h:63   */
#define GFSB200_CELL_LEAF      4u   /* FTT_CELL_IS_LEAF and not destroyed   */

/* force kinds, in GfsParticleList force-list order */
#define GFSB200_FORCE_DRAG 1        /* GfsForceDrag */
#define GFSB200_FORCE_LIFT 2        /* GfsForceLift */
#define GFSB200_FORCE_BUOY 3        /* GfsForceBuoy */
#define GFSB200_FORCE_INERTIAL 4    /* GfsForceInertial   modules/particulatecommon.c:255-303 */
#define GFSB200_FORCE_ADDEDMASS 5   /* GfsForceAddedMass  modules/particulatecommon.c:331-394 */
#define GFSB200_MAX_FORCES 8

const char * gfsb200_last_error (void);
const char * gfsb200_version (void);

/* ------------------------------------------------------------------------
 * Flat tree (host side).  A level-ordered array of cells: all root cells
 * first (GfsBox roots before GfsBoundary roots), then every deeper level in
 * turn; the 2^dim children of a cell are contiguous, in FTT child order
 * (src/ftt.c:301-316), and sibling groups follow their parents' order, so
 * within a level cells are keyed by the Morton path from their root.
 * ------------------------------------------------------------------------ */
typedef struct gfsb200_tree gfsb200_tree;

gfsb200_tree * gfsb200_tree_new (int dim);
void gfsb200_tree_free (gfsb200_tree * t);

/* Roots must all be added before any split.  pos = centre of the root cell,
 * level = its FTT level (GfsDomain.rootlevel), is_box = 1 for a GfsBox root,
 * 0 for a GfsBoundary (ghost) root.  Returns the root index (= cell index). */
int gfsb200_tree_add_root (gfsb200_tree * t, const double pos[3], int level, int is_box);
/* root r0's neighbour in direction d is root r1 (and vice versa) */
int gfsb200_tree_link_roots (gfsb200_tree * t, int r0, int d, int r1);

/* Raw split: gives `cell` 2^dim children; bit n of destroyed_mask marks child
 * n destroyed, `child_flags` is OR-ed into every child.  No 2:1 balancing
 * (used when mirroring an existing FTT tree).  Returns index of child 0. */
int gfsb200_tree_split (gfsb200_tree * t, int cell, unsigned destroyed_mask, unsigned child_flags);
/* Split with the face 2:1 rule of oct_new(check_neighbors), src/ftt.c:45-84 */
int gfsb200_tree_refine_cell (gfsb200_tree * t, int cell);

/* Native world builders (single unit GfsBox at the origin unless roots were
 * added by hand); semantics of ftt_cell_refine, src/ftt.c:169-192. */
typedef int (* gfsb200_refine_func) (const double pos[3], int level, double h, void * data);
int gfsb200_tree_refine (gfsb200_tree * t, gfsb200_refine_func f, void * data);
int gfsb200_tree_refine_uniform (gfsb200_tree * t, int level);
/* refine while level < minlevel, or level < maxlevel and the distance from the
 * cell centre to the circle of radius R (plane z = 0) is < factor*h */
int gfsb200_tree_refine_ring (gfsb200_tree * t, int minlevel, int maxlevel, double R, double factor);
/* the corner-balance sweep of gfs_simulation_refine, src/simulation.c:1226-1231 */
int gfsb200_tree_corner_sweep (gfsb200_tree * t);
/* ghost-cell tree on side `side` of box root `box_root`, mirroring
 * boundary_match, src/boundary.c:652-685; call after all refinement */
int gfsb200_tree_add_boundary (gfsb200_tree * t, int box_root, int side);

/* the GfsBoundary on `side` of box root `box_root` is a GfsBoundaryPeriodic whose
 * matching box is `matching_box_root` (src/boundary.c:1500-1532); used by
 * gfsb200_particle_bc.  Call after add_boundary / link_roots. */
int gfsb200_tree_set_periodic (gfsb200_tree * t, int box_root, int side, int matching_box_root);

/* Mixed (solid-cut) cell of a FINALIZED tree: fluid fraction a in (0,1], centre of mass of the
 * fluid part and fluid fraction s[2*dim] of each face (GfsSolidVector.a / .cm / .s,
 * src/fluid.h:54-59; s = NULL: all faces open).  The centre of mass enters the
 * corner-interpolator weights (distance (), src/fluid.c:2983-3003), the fraction the cell
 * volume of the force deposits (gfs_cell_volume, src/domain.h:503-508), the face fractions the
 * gradients behind the vorticity and the inertial force (a closed face has no neighbour:
 * gfs_cell_face, src/fluid.c:42-52; finer neighbours are averaged with their face fractions:
 * average_neighbor_value, :64-93).  Cells that are entirely solid are destroyed cells of the
 * tree (split's destroyed_mask).  Call before build_stencils. */
int gfsb200_tree_set_solid (gfsb200_tree * t, int cell, double a, const double cm[3], const double * s);

/* Reorders to level order, builds neighbour tables, centres, locate array.
 * perm (may be NULL, else n_cells ints) receives old index -> new index. */
int gfsb200_tree_finalize (gfsb200_tree * t, int32_t * perm);
/* Corner-interpolator stencils for every (leaf, corner), deduplicated into a
 * vertex table (requires finalize). */
int gfsb200_tree_build_stencils (gfsb200_tree * t);

/* read-only views of a finalized tree (pointers owned by the tree) */
typedef struct {
  int32_t dim, n_cells, n_roots, n_box_roots;
  int32_t min_level, max_level;       /* absolute FTT levels present */
  int32_t complete_level;             /* deepest absolute level to which every GfsBox tree is
					 fully refined (>= min_level) */
  int64_t n_leaves;                   /* non-ghost leaves */
  const int32_t * level_start;        /* [max_level - min_level + 2] */
  const int32_t * parent;             /* [n_cells], -1 for roots */
  const int32_t * child0;             /* [n_cells], -1 for leaves */
  const int32_t * neighbor;           /* [n_cells][2*dim], ftt_cell_neighbor semantics, -1 = NULL */
  const uint8_t * level;              /* [n_cells] absolute level */
  const uint8_t * flags;              /* [n_cells] GFSB200_CELL_* */
  const double  * pos;                /* [n_cells][3] exact centres */
  /* GfsLocateArray */
  double la_min[3], la_h;
  int32_t la_n[3];
  const int32_t * la_slot;            /* [la_n product] GfsBox root cell or -1 */
  /* stencils (NULL / 0 until build_stencils) */
  int32_t n_vertices;
  const int32_t * vtx_off;            /* [n_vertices + 1] CSR offsets */
  const int32_t * vtx_cell;           /* stencil cells, reference slot order of the canonical corner */
  const double  * vtx_w;              /* normalised weights */
  const int32_t * leaf_vtx;           /* [n_cells][2^dim] vertex id per corner (reference corner
					 order, src/fluid.c:2588-2606), -1 for non-leaves */
  int32_t lattice_level;              /* >= 0: all leaves at this level and vertex ids are
					 row-major on its (2^L + 1)^dim lattice; else -1 */
  const double * solid_a;             /* [n_cells] fluid fraction (1 = not mixed), or NULL: no mixed cell */
  const double * solid_cm;            /* [n_cells][3] centre of mass (NaN = not mixed), or NULL */
  const double * solid_s;             /* [n_cells][2*dim] face fractions (1 = not mixed), or NULL */
} gfsb200_tree_view;

int gfsb200_tree_get_view (const gfsb200_tree * t, gfsb200_tree_view * v);
/* the (cell, weight) list of gfs_cell_corner_interpolator (cell, corner k),
 * in reference order; cells/w need 29 entries.  Returns count or <0. */
int gfsb200_tree_corner_interpolator (const gfsb200_tree * t, int cell, int k,
				      int32_t * cells, double * w);

/* ------------------------------------------------------------------------
 * Device context.  One per GPU / per process rank.
 * ------------------------------------------------------------------------ */
typedef struct gfsb200_ctx gfsb200_ctx;

/* number of CUDA devices the process sees (0: none, or no usable driver) */
int gfsb200_device_count (void);
int gfsb200_ctx_create (int device, gfsb200_ctx ** out);
void gfsb200_ctx_destroy (gfsb200_ctx * c);
/* the CUDA stream all of this context's work is issued on (cudaStream_t) */
void * gfsb200_ctx_stream (gfsb200_ctx * c);
int gfsb200_ctx_synchronize (gfsb200_ctx * c);

/* Page-locked host staging buffers for a C host that has no CUDA headers (the GModule):
 * copies from/to them run at full PCIe rate and asynchronously.  NULL on failure. */
void * gfsb200_host_alloc (size_t bytes);
void gfsb200_host_free (void * p);

/* Upload the flattened tree + stencil tables; call again after each adapt. */
int gfsb200_upload_tree (gfsb200_ctx * c, const gfsb200_tree * t);

/* Mirror the cell variables (host arrays of n_cells doubles, flat-tree order;
 * w = NULL in 2D; alpha/mu = NULL for constants) and rebuild the per-vertex
 * velocity table and the per-leaf vorticity table on the device. */
int gfsb200_upload_field (gfsb200_ctx * c, const double * u, const double * v, const double * w,
			  const double * alpha, const double * mu);
/* The same mirror in slices: copies cells [first, first + n) of every array given (the arrays are
 * indexed by flat cell, as above; the set of non-NULL arrays must be the same in every call of one
 * update) and returns at once -- page-locked host arrays make the copy overlap whatever the host
 * does next, e.g. gathering the next slice out of the FttCell tree.  The tables are NOT rebuilt:
 * finish the update with gfsb200_refresh_field. */
int gfsb200_upload_field_part (gfsb200_ctx * c, int64_t first, int64_t n,
			       const double * u, const double * v, const double * w,
			       const double * alpha, const double * mu);
/* same, from device pointers */
int gfsb200_set_field_device (gfsb200_ctx * c, const double * u, const double * v, const double * w,
			      const double * alpha, const double * mu);
/* Mirror the previous-step velocity Un,Vn(,Wn) kept by GfsForceInertial /
 * GfsForceAddedMass (GfsForceCoeff.Uold, store_domain_previous_vel,
 * modules/particulatecommon.c:99-113) and build its vertex table.  Needed only
 * when one of those forces is in the list. */
int gfsb200_upload_field_prev (gfsb200_ctx * c, const double * un, const double * vn, const double * wn);
/* recompute vertex + vorticity tables from the resident field (the per-step
 * cell pass) */
int gfsb200_refresh_field (gfsb200_ctx * c);
/* device -> host copies of derived tables, for parity tests */
int gfsb200_download_corner_values (gfsb200_ctx * c, int comp, int64_t n, const int32_t * cells,
				    double * out /* [n][2^dim] */);
int gfsb200_download_vorticity (gfsb200_ctx * c, int64_t n, const int32_t * cells,
				double * out /* [n][3] */);

/* ---- particles: SoA fp64 on the device -------------------------------- */
/* host arrays of n; z/vz = NULL in 2D; id may be NULL (ids 1..n assigned) */
int gfsb200_particles_upload (gfsb200_ctx * c, int64_t n,
			      const double * x, const double * y, const double * z,
			      const double * vx, const double * vy, const double * vz,
			      const double * mass, const double * volume, const uint32_t * id);
/* any output pointer may be NULL */
int gfsb200_particles_download (gfsb200_ctx * c,
				double * x, double * y, double * z,
				double * vx, double * vy, double * vz,
				double * fx, double * fy, double * fz,
				double * mass, double * volume, uint32_t * id, int32_t * cell);
int64_t gfsb200_particles_count (gfsb200_ctx * c);
/* reserve device storage for n particles without a host copy, and expose the
 * SoA device pointers (x,y,z,vx,vy,vz,mass,volume) so a harness can fill them
 * on the device */
int gfsb200_particles_resize (gfsb200_ctx * c, int64_t n);
int gfsb200_particles_device_ptrs (gfsb200_ctx * c, double * ptrs[8]);

typedef struct {
  double dt;                       /* sim->advection_params.dt */
  int32_t n_forces;
  int32_t force[GFSB200_MAX_FORCES]; /* GFSB200_FORCE_* in list order; n_forces = 0 => passive
				        tracer (gfs_domain_advect_point) */
  double rho;                      /* 1/alpha when no alpha array is resident (alpha unset => 1) */
  double mu;                       /* GfsSourceDiffusion constant; 0 => drag returns 0 */
  double g[3];                     /* sum of GfsSource intensities on U,V,W */
  double cd_const;                 /* constant GfsForceDrag coefficient function; NaN = built-in */
  double cl_const;                 /* constant GfsForceLift coefficient function; NaN = 0.5 */
  int32_t record_cells;            /* 1: also store each particle's containing cell index */
  int32_t record_forces;           /* 1: also store the accumulated force (particulate->force) */
  double cm_const;                 /* constant GfsForceAddedMass coefficient function; NaN = 0.5 */
  int32_t track_escapes;           /* 1: remember which particles left the domain (with their
				      previous position) for gfsb200_particle_bc */
  int32_t fuse_deposit;            /* 1: the two-way deposits of the step in the same pass: after the
				      integration each particle adds V_p/V_cell and its on-fluid force at
				      its NEW state to the deposit buffer -- the result of calling
				      gfsb200_deposit_all right after the step (the buffer is zeroed first),
				      without streaming the particles twice */
} gfsb200_step_params;

void gfsb200_step_params_default (gfsb200_step_params * p);

/* One fused locate + interpolate + forces + integrate pass over all resident
 * particles (gfs_particulate_event for every list member).  Particles whose
 * gfs_domain_locate is NULL are left untouched (cell = -1). */
int gfsb200_step (gfsb200_ctx * c, const gfsb200_step_params * p);

/* The same step for a particle list that LIVES ON THE HOST (the GtsObject list
 * of the reference, gathered into SoA arrays): positions and velocities are
 * updated in place.  The arrays are streamed through the device in chunks on
 * three streams (H2D / kernel / D2H overlap), so the cost is that of the
 * slower PCIe direction, not the sum.  Pin the arrays (cudaHostRegister /
 * cudaMallocHost) for full bandwidth.  Particles outside the domain are left
 * untouched.  chunk = 0 picks 2^20 particles. */
int gfsb200_step_host (gfsb200_ctx * c, const gfsb200_step_params * p, int64_t n,
		       double * x, double * y, double * z,
		       double * vx, double * vy, double * vz,
		       const double * mass, const double * volume, int64_t chunk);

/* gfs_particle_list_event: cull particles outside the domain
 * (remove_particles_not_in_domain), step, then gfs_particle_bc (periodic wrap /
 * drop).  *n_removed (may be NULL) counts the culled and the dropped.  With forces the
 * passes run as step -> BC -> cull -- the same list, since the step leaves a particle that
 * is outside the domain untouched -- and the BC and cull passes are skipped when the step
 * kernel counted no particle for them. */
int gfsb200_particle_list_event (gfsb200_ctx * c, const gfsb200_step_params * p,
				 int64_t * n_removed);
int gfsb200_particles_cull (gfsb200_ctx * c, int64_t * n_removed);
/* gfs_particle_bc (modules/particulatecommon.c:3375-3395) for the particles that
 * left the domain during the LAST gfsb200_step issued with track_escapes = 1
 * (gfsb200_particle_list_event always tracks): the exit face is found by the
 * reference's cell-to-cell ray walk (boundarycell / check_intersetion,
 * :3058-3186); a particle leaving through a periodic side is wrapped
 * (periodic_bc_particle :3189-3214), any other is dropped from the list.
 * Unlike the reference, wrapped particles keep their place in the list. */
int gfsb200_particle_bc (gfsb200_ctx * c, int64_t * n_wrapped, int64_t * n_dropped);
/* How many particles left the domain (gfs_domain_locate == NULL at their new position)
 * during the last gfsb200_step issued with track_escapes = 1.  A caller that keeps
 * gfs_particle_bc on the host (modules/particulatecommon.c:3375-3395: one gfs_domain_locate
 * per particle, every step) can skip it when this is 0 -- the reference function would find
 * nothing to do.  Does not consume the record: gfsb200_particle_bc may still follow. */
int gfsb200_escaped_count (gfsb200_ctx * c, int64_t * n_escaped);
/* Both counts of the last tracked step in one read-back: particles that left the domain
 * during the step, and particles that were already outside before it (the step leaves those
 * untouched; gfsb200_particles_cull removes them).  Either pointer may be NULL. */
int gfsb200_step_counts (gfsb200_ctx * c, int64_t * n_escaped, int64_t * n_outside);
/* The same record itself: list positions (idx[k], in the order of the resident list) and
 * positions BEFORE the step (old_xyz[3k..3k+2]; what the reference keeps in GfsParticle.pos_old
 * for the ray walk of gfs_particle_bc) of up to `cap` escaped particles; *n_out = how many
 * were written. */
int gfsb200_escaped_download (gfsb200_ctx * c, int64_t cap, int32_t * idx, double * old_xyz,
			      int64_t * n_out);
/* re-sort resident particles by containing cell (Morton-ordered flat index)
 * so that neighbouring threads gather neighbouring cells */
int gfsb200_particles_sort (gfsb200_ctx * c);

/* Batched gfs_domain_locate: host points -> flat cell index or -1. */
int gfsb200_locate (gfsb200_ctx * c, int64_t n, const double * x, const double * y,
		    const double * z, int32_t * cell);
/* Batched locate + gfs_interpolate of U,V,W at host points; out[comp] may be NULL */
int gfsb200_interpolate (gfsb200_ctx * c, int64_t n, const double * x, const double * y,
			 const double * z, double * u, double * v, double * w);

/* GfsOutputLocation (gfs_output_location_event, src/output.c:1153-1212) for ANY cell
 * variables, batched: every point is located once, then each of the nvar variables
 * (vars[k]: host array of n_cells doubles in flat-tree order) is evaluated there by
 * gfs_interpolate (interpolate != 0: corner stencils + trilinear, src/fluid.c:2697-2710) or
 * as the containing cell's value (interpolate == 0).  out[k]: n doubles; cell (may be NULL):
 * flat index or -1 -- the reference prints no line for a point outside the domain, and
 * out[k][i] is GFS_NODATA there.  The resident U,V,W tables are not disturbed. */
int gfsb200_output_location (gfsb200_ctx * c, int nvar, const double * const * vars, int interpolate,
			     int64_t n, const double * x, const double * y, const double * z,
			     double * const * out, int32_t * cell);

/* ---- particle data formats --------------------------------------------- */
/* The particle block of a GfsParticleList as the reference writes it into a .gfs / simulation
 * dump: one line per resident particle,
 *   "    <class> id x y z mass volume*L^dim vx vy vz fx fy fz\n"
 * with the "%d"/"%g" conversions of gfs_particle_write (src/particle.c:86-98) and
 * gfs_particulate_write (modules/particulatecommon.c:910-926), indented as gfs_event_list_write
 * does (src/event.c:2508-2523).  class_name is "GfsParticulate" for the reference's lists.
 * append != 0 appends to the file.  (%g keeps 6 digits: use the checkpoint below to restart.) */
int gfsb200_particles_write_gfs (gfsb200_ctx * c, const char * path, const char * class_name,
				 double L, int append);
/* Lossless binary checkpoint of the resident particle state (SoA fp64 columns, ids, last
 * recorded forces) and its restore; the file starts with "GFSB200P", a format version, dim
 * and the particle count.  Replaces the lossy text round trip on restart. */
int gfsb200_checkpoint_save (gfsb200_ctx * c, const char * path);
int gfsb200_checkpoint_load (gfsb200_ctx * c, const char * path);

/* ---- two-way coupling -------------------------------------------------- */
/* GfsParticulateField: field[cell] = sum V_p / V_cell over resident particles
 * (nearest cell).  The device field is zeroed first. */
int gfsb200_deposit_volume (gfsb200_ctx * c);
/* GfsSourceParticulate in the single-cell limit: forces recomputed without
 * GfsForceBuoy, field_c[cell] -= F_c / rho / V_cell. */
int gfsb200_deposit_force (gfsb200_ctx * c, const gfsb200_step_params * p);
/* both of the above in one pass over the particles (one locate, one kernel) */
int gfsb200_deposit_all (gfsb200_ctx * c, const gfsb200_step_params * p);
/* GfsSourceParticulate WITH its smoothing kernel (source_particulate_event,
 * modules/particulatecommon.c:2177-2228): the force of every particle on the
 * fluid is spread over the leaves reached by the reference's conditional
 * traversal (cond_kernel :2126-2156: cells whose circumscribed sphere comes
 * within `rkernel' -- an absolute distance, as in the reference -- of the
 * particle, or that contain it), weighted by kernel(offset/r_b)/correction
 * with the per-particle normalisation of kernel_volume :2108-2119:
 *   field_c[cell] -= F_c / rho(cell) / V_cell * K / correction.
 * The user's `kernel = <GfsFunction>' must be one of the closed forms below
 * (the offset (x,y,z) is in particle radii r_b = (3V/4pi)^(1/3)); anything else
 * stays on the host.  NB the reference's distance_normalization (:2087-2098)
 * zeroes the z offset before using it, so in 3D z = -z_particle/r_b for every
 * cell; that is reproduced unless GFSB200_KERNEL_FIX_Z is set. */
#define GFSB200_KERNEL_CONSTANT 0   /* K = a */
#define GFSB200_KERNEL_GAUSSIAN 1   /* K = a exp(-b (x^2+y^2+z^2)) */
#define GFSB200_KERNEL_COMPACT  2   /* K = a (1 - b r^2)^p where b r^2 < 1, else 0 */
#define GFSB200_KERNEL_FIX_Z    1   /* flags: z offset = (z_cell - z_particle)/r_b */
typedef struct {
  int32_t kind;                    /* GFSB200_KERNEL_* */
  int32_t p;                       /* COMPACT: integer exponent >= 0 */
  double a, b;
  int32_t flags;
  int32_t record_norm;             /* 1: keep each particle's (correction, volume) for
				      gfsb200_download_kernel_norm */
} gfsb200_kernel;
/* Recognise a user kernel by probing it: f is sampled at ~40 offsets (in
 * particle radii) and *out receives the closed form that reproduces every
 * sample to 1e-12 of the kernel's peak; GFSB200_ERR_UNSUPPORTED if none does
 * (the list then stays on the reference's CPU event).  Host-only, no device
 * needed.  The drop-in module passes a trampoline to gfs_function_spatial_value
 * of GfsSourceParticulate.kernel_function (modules/particulatecommon.c:2282-2290). */
typedef double (* gfsb200_kernel_func) (double x, double y, double z, void * data);
int gfsb200_kernel_fit (gfsb200_kernel_func f, void * data, int dim, gfsb200_kernel * out);
/* fills deposit components 1..dim (zeroed first); component 0 is untouched */
int gfsb200_deposit_force_smoothed (gfsb200_ctx * c, const gfsb200_step_params * p,
				    double rkernel, const gfsb200_kernel * kernel);
/* per-particle normalisation of the last smoothed deposit issued with record_norm = 1 */
int gfsb200_download_kernel_norm (gfsb200_ctx * c, double * correction, double * volume);
/* There are two deposition buffers; select which one the deposit calls, the
 * buffer query and the download use (default 0).  Alternating them lets the
 * all-reduce of step n run on another stream while step n+1 deposits. */
int gfsb200_deposit_select (gfsb200_ctx * c, int which);
/* device pointer / element count of the deposition buffer
 * ([1 + dim][n_cells]: void fraction, Fx, Fy(, Fz)); for the multi-GPU
 * all-reduce (NCCL) issued by the caller on gfsb200_ctx_stream() */
int gfsb200_deposit_buffer (gfsb200_ctx * c, double ** dev, int64_t * count);
int gfsb200_download_deposit (gfsb200_ctx * c, int comp, double * out /* [n_cells] */);

/* ---- multi-GPU --------------------------------------------------------- */
/* Particles shard over the GPUs of one box; the flat tree and the field are replicated.  One
 * communicator per context (= per GPU); a process may hold one rank (MPI / torchrun: one process per
 * GPU) or all of them (a serial gerris3D driving every GPU of the box).  Every collective below takes
 * the array of the CALLING PROCESS's communicators and must be called by every process of the job.
 *
 * Replaces, for a replicated field: the per-particle migration of gfs_particle_bc in a parallel
 * run (mpi_send_particle / mpi_rcv_particle, modules/particulatecommon.c:3218-3244, over
 * gfs_send_objects / gfs_receive_objects, src/domain.c:4464-4557) by gfsb200_comm_rebalance, and the
 * scalar reductions of src/utils.h:36-42 (gfs_all_reduce) by a reduction of the whole deposited
 * field.  NCCL is loaded at run time (libnccl.so.2) by the first call below: a single-GPU user never
 * needs it. */
typedef struct gfsb200_comm gfsb200_comm;
#define GFSB200_UNIQUE_ID_BYTES 128
/* rank 0 creates the id (ncclGetUniqueId); the host code hands it to every rank (MPI_Bcast, a file) */
int gfsb200_comm_unique_id (void * id /* GFSB200_UNIQUE_ID_BYTES */);
/* one process per GPU */
int gfsb200_comm_init_rank (gfsb200_ctx * c, const void * id, int rank, int nranks, gfsb200_comm ** out);
/* one process, n GPUs: rank r is ctxs[r] */
int gfsb200_comm_init_all (int n, gfsb200_ctx * const * ctxs, gfsb200_comm ** out /* [n] */);
void gfsb200_comm_destroy (gfsb200_comm * m);
int gfsb200_comm_rank (const gfsb200_comm * m);
int gfsb200_comm_size (const gfsb200_comm * m);
/* 1 when every rank reaches every other rank's memory over NVLink peer access (same box) */
int gfsb200_comm_peer_access (const gfsb200_comm * m);

/* U,V,W (+alpha, mu) of rank `root' -- host arrays as for gfsb200_upload_field, read on the root's
 * process only -- go up over PCIe ONCE and reach the other GPUs over NVLink (ncclBroadcast); every
 * rank then rebuilds its vertex and vorticity tables. */
int gfsb200_broadcast_field (gfsb200_comm * const * local, int n_local, int root,
			     const double * u, const double * v, const double * w,
			     const double * alpha, const double * mu);

/* Global cell order: the resident particles of all ranks are redistributed (NCCL send/recv over
 * NVLink) so that rank r holds, sorted by cell, the particles of the flat-tree cells
 * [split[r], split[r + 1]) -- contiguous slices of the globally Morton-sorted cloud with equal
 * particle counts up to one cell's population.  From then on a rank's deposits fall into its own
 * slice of the field (except for the particles that drift across a slice boundary before the next
 * rebalance), and gfsb200_deposit_allreduce shrinks from a reduction of the whole field to an
 * all-gather of the slices.  Particle ids travel with the particles; recorded forces do not.
 * Call it where a single GPU would call gfsb200_particles_sort. */
int gfsb200_comm_rebalance (gfsb200_comm * const * local, int n_local);
/* Target shares of the particles for the next rebalance: share[r] > 0 for every rank of the job, any
 * scale (NULL: equal shares, the default).  Equal counts are not equal work on an adaptive tree; a caller
 * that has timed its ranks (gfsb200_timer_read) hands each one a share inversely proportional to its
 * measured time per particle. */
int gfsb200_comm_set_shares (gfsb200_comm * const * local, int n_local, const double * share);
/* the slice boundaries of the last rebalance: split[0 .. size]; GFSB200_ERR_STATE before it, and on
 * adaptive trees with several ranks, where the slices are ranges of the DEPTH-FIRST leaf order (a
 * particle that crosses into a leaf of another level stays near its rank's slice; in the level-ordered
 * cell index it would land in another rank's) -- there the owner of a leaf is looked up: */
int gfsb200_comm_split (const gfsb200_comm * m, int32_t * split);
/* owner[c], c < n_cells: the rank that owns cell c after the last rebalance (255: not a leaf) */
int gfsb200_comm_owner_table (const gfsb200_comm * m, uint8_t * owner);
/* host-only helpers (exposed for tests): slice boundaries from the global per-cell particle counts;
 * owners of the leaves from the tree (child0[c]: first of the 2^dim consecutive children of cell c, < 0
 * for a leaf; roots: cells 0 .. n_roots - 1) and the counts, equal shares along the depth-first order */
int gfsb200_comm_splitters (const uint32_t * count, int32_t n_cells, int nranks, int32_t * split);
int gfsb200_comm_owner_slices (const int32_t * child0, int32_t n_cells, int32_t n_roots, int dim,
			       const uint32_t * count, int nranks, uint8_t * owner);

/* Sums the deposited field over the ranks: afterwards every rank holds sum_r deposit_r in the
 * buffer gfsb200_download_deposit reads.  Asynchronous: the exchange runs on a communication stream
 * behind the deposit; the next step may be issued at once (the library alternates the two deposit
 * buffers) and gfsb200_download_deposit / gfsb200_deposit_wait order themselves after it.
 *   - after gfsb200_comm_rebalance, on a box with peer access: the deposit kernels have already
 *     reduced every contribution into its OWNER's slice (remote fp64 reductions through NVLink peer
 *     memory for the drifters), so this is one cross-GPU barrier, the push of the own slice to every
 *     peer (copy engines) and a completion flag -- (R-1)/R of the field per GPU instead of 2(R-1)/R;
 *   - otherwise: ncclAllReduce of the whole buffer.
 * With a communicator attached each deposit call must be followed by one allreduce before the next
 * deposit of the same component. */
int gfsb200_deposit_allreduce (gfsb200_comm * const * local, int n_local);
/* orders the context's stream after the last exchange (own pushes AND the peers' pushes) */
int gfsb200_deposit_wait (gfsb200_comm * m);
#define GFSB200_EXCHANGE_AUTO 0       /* owner slices + all-gather when possible */
#define GFSB200_EXCHANGE_ALLREDUCE 1  /* always ncclAllReduce the whole buffer (required for the smoothed deposit) */
int gfsb200_comm_set_exchange (gfsb200_comm * m, int mode);
/* device time (ms, CUDA events on the communication stream) of the exchanges since the last call,
 * averaged, and the bytes one exchange sends from this rank */
int gfsb200_comm_exchange_stats (gfsb200_comm * m, double * ms, int64_t * n, int64_t * bytes_sent);

/* ---- timing ------------------------------------------------------------ */
/* average device time (ms) of the fused step kernel over the launches since
 * the last reset, measured with CUDA events on the context's stream */
int gfsb200_timer_reset (gfsb200_ctx * c);
int gfsb200_timer_read (gfsb200_ctx * c, double * step_kernel_ms, int64_t * launches);
/* time every `every`-th launch only (1, the default: all of them; 0: none).  The two events of a timed
 * launch are stream operations between the cell pass and the step kernel -- they also keep the step
 * kernel's programmatic launch from overlapping the cell pass -- and cost 1.5 % of a C2 step
 * (profiles/README.md, round 2); a sample of the launches gives the same average. */
int gfsb200_timer_sampling (gfsb200_ctx * c, int every);
/* how many of this library's own CUDA kernels have been launched by the process so far (the
 * library kernels inside cub::DeviceRadixSort / DeviceSelect are not counted) */
int64_t gfsb200_kernel_launches (void);

#ifdef __cplusplus
}
#endif

#endif /* GFSB200_H */

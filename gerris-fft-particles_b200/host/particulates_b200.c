/* particulates_b200.c -- drop-in `particulates` GModule for gerris2D / gerris3D
 * whose per-step particle work runs on a B200 through the C-ABI of
 * include/gfsb200.h.
 *
 * Build (on a machine that has Gerris' development files; GLib/GTS are absent
 * from the image this repository is developed in, so the module cannot be
 * LINKED there.  It is TYPE-CHECKED there, in 2D and 3D, against the reference's
 * own src headers and modules/particulatecommon.h with declaration-only GLib/GTS
 * stand-ins: `make -C gerris-fft-particles_b200 check-host`, run by
 * __graft_entry__.build() and tests/test_capi_symbols.py -- see INTEGRATION.md):
 *
 *   gcc -shared -fPIC [-DFTT_2D=1] `pkg-config --cflags gerris3D` \
 *       particulates_b200.c $(GERRIS)/modules/particulatecommon.c ftt_bridge.c \
 *       -L$PREFIX/lib -lgfsb200 `pkg-config --libs gerris3D` \
 *       -o libparticulates3D.so
 *
 * Design: the reference's modules/particulatecommon.c is linked UNCHANGED and
 * keeps every GtsObject class, the .gfs syntax, read/write methods and the
 * rare host-side machinery (droplet conversion, feeding, output, particle
 * BCs).  This file only replaces, in the class vtables, the three event
 * methods that make up the hot path:
 *
 *   GfsParticleList.event     gfs_particle_list_event    modules/particulatecommon.c:980-1015
 *   GfsParticulateField.event particulate_field_event    modules/particulatecommon.c:1934-1957
 *   GfsSourceParticulate.event source_particulate_event  modules/particulatecommon.c:2177-2228
 *
 * and exports the same module symbols as modules/particulates.c:24-49.
 *
 * There is NO silent CPU fallback.  A list that carries something the device does not implement
 * (user force classes, coefficient / density functions that are neither a constant nor a plain
 * cell variable, a smoothing kernel outside the closed forms of gfsb200.h), a simulation with
 * moving solids, or an MPI domain decomposition (the device path replicates the field and shards
 * the particles over the GPUs of one box) stops the run with a g_error naming the object.  Setting
 * GFSB200_ALLOW_REFERENCE_EVENT=1 opts in to the reference's own CPU event for exactly those
 * objects, with one g_warning per object.
 *
 * Environment: GFSB200_DEVICE=<n> (first device, default 0), GFSB200_DEVICES=<n> (GPUs of the box
 * one serial gerris drives, default 1),
 * GFSB200_RESIDENT=0 (write every particle back into its GfsParticulate after every event; by
 * default the device copy is authoritative between events and the objects are refreshed when
 * something on the host is about to look at them), GFSB200_MODULE_PROFILE=1.
 */
#include <stdlib.h>
#include <string.h>
#include <stddef.h>
#include <stdio.h>
#include <time.h>
#include <gfs.h>
#include "particulatecommon.h"
#include "gfsb200.h"
#include "gfsb200_ftt.h"

/* ------------------------------------------------------------------ */
/* per-list device state, hung on the GtsObject through g_object data  */

#define B200_MAX_DEV 16

typedef struct {
  gfsb200_ctx * ctx;                    /* = dev[0] */
  /* GFSB200_DEVICES=n: one serial gerris drives n GPUs of the box -- the list is sharded over them
     (gfsb200_comm_rebalance), tree and field are replicated (gfsb200_broadcast_field), deposits
     are summed (gfsb200_deposit_allreduce) before they are scattered into the cells */
  gint n_dev;
  gfsb200_ctx * dev[B200_MAX_DEV];
  gfsb200_comm * comm[B200_MAX_DEV];
  gint32 * idx_of_id;                   /* n_dev > 1: object index by particle id (the devices reorder the list) */
  guint32 idx_cap;
  gfsb200_tree * tree;
  gfsb200_ftt_map * map;
  guint mesh_epoch;                     /* b200_mesh_epoch when the tree was flattened */
  guint n_roots;                        /* box / boundary roots at that time ... */
  gsize root_hash;                      /* ... and a hash of their addresses */
  guint data_generation;                /* domain->allocated->len: the cell data blocks move when it grows */
  gboolean tree_valid;
  gdouble * dep_out;                    /* page-locked staging of one deposited component */
  gint32 dep_cap;
  gdouble * field[3];                   /* page-locked staging of U,V,W in flat order */
  gdouble * cellvar[2];                 /* page-locked staging of per-cell alpha / viscosity, when they are variables */
  gint32 field_cap;                     /* cells the five buffers above hold */
  GfsParticulate ** obj;                /* the list's objects in list order (rebuilt every event) */
  gdouble * col[10];                    /* page-locked staging: x y z vx vy vz mass volume | fx fy fz mass on the way back */
  guint32 * id;
  gint64 part_cap, obj_cap;
  gint64 n_obj;                         /* objects in obj[] */
  gboolean list_known;                  /* obj[0..n_obj) IS the GSList: no add/remove since it was walked */
  /* $GFSB200_RESIDENT=1: the device copy stays authoritative between events; the GtsObjects
     are refreshed (sync_down) only when something on the host is about to look at them */
  gboolean resident;
  gboolean mpi;                         /* parallel run: gfs_particle_bc every event (it also receives) */
  gboolean host_stale;                  /* the device holds newer particle state than the objects */
  gboolean uploaded;                    /* device_current had to upload the objects */
  gboolean syncing;                     /* inside sync_down (its own removals must not recurse) */
  GfsParticleList * plist;
  GfsVariable * alpha_var, * mu_var;    /* PhysicalParams alpha = <variable>; GfsDiffusion.mu (src/source.c:941-946) */
  GfsVariable ** uold;                  /* GfsForceCoeff.Uold of an inertial / added-mass force, or NULL */
  gboolean snapshot;                    /* a GfsForceInertial is in the list: Un,Vn,Wn are refreshed after each event */
  gint32 n_cells;
  gboolean (* reference_event) (GfsEvent *, GfsSimulation *);
  /* $GFSB200_MODULE_PROFILE: seconds spent per phase of the list event, printed on destroy */
  gdouble phase[6];
  guint n_events;
  gboolean warm;
} B200State;

enum { PH_TREE, PH_FIELD, PH_UPLOAD, PH_DEVICE, PH_DOWNLOAD, PH_BC };

static gdouble wall (void)
{
  struct timespec t;
  clock_gettime (CLOCK_MONOTONIC, &t);
  return t.tv_sec + 1e-9*t.tv_nsec;
}

static GHashTable * b200_states = NULL;   /* GfsParticleList* -> B200State* */
static gboolean (* reference_list_event) (GfsEvent *, GfsSimulation *) = NULL;
static gboolean (* reference_field_event) (GfsEvent *, GfsSimulation *) = NULL;
static gboolean (* reference_source_event) (GfsEvent *, GfsSimulation *) = NULL;

/* No silent CPU fallback: an object the device path cannot express stops the run, unless the
 * user has opted in to the reference's own CPU event for it (one warning per object). */
static void reference_or_error (gpointer object, const gchar * name, const gchar * why)
{
  static GHashTable * warned = NULL;
  const gchar * env = g_getenv ("GFSB200_ALLOW_REFERENCE_EVENT");
  if (!env || atoi (env) == 0)
    g_error ("particulates (B200): %s: %s cannot run on the device.  There is no silent CPU fallback; "
	     "set GFSB200_ALLOW_REFERENCE_EVENT=1 to run the reference's own CPU event for this object",
	     name ? name : "(unnamed)", why);
  if (!warned)
    warned = g_hash_table_new (NULL, NULL);
  if (!g_hash_table_lookup (warned, object)) {
    g_hash_table_insert (warned, object, object);
    g_warning ("particulates (B200): %s: %s cannot run on the device: the reference's CPU event runs instead "
	       "(GFSB200_ALLOW_REFERENCE_EVENT)", name ? name : "(unnamed)", why);
  }
}

static B200State * state_of (GfsParticleList * plist)
{
  B200State * s;
  if (!b200_states)
    b200_states = g_hash_table_new (NULL, NULL);
  s = g_hash_table_lookup (b200_states, plist);
  if (!s) {
    int device = 0, k;
    const gchar * env = g_getenv ("GFSB200_DEVICE");
    s = g_malloc0 (sizeof (B200State));
    if (env) device = atoi (env);
    s->n_dev = g_getenv ("GFSB200_DEVICES") ? atoi (g_getenv ("GFSB200_DEVICES")) : 1;
    if (s->n_dev < 1 || s->n_dev > B200_MAX_DEV || device + s->n_dev > MAX (gfsb200_device_count (), 1))
      g_error ("particulates (B200): GFSB200_DEVICES=%d from device %d, but the process sees %d device(s)",
	       s->n_dev, device, gfsb200_device_count ());
    for (k = 0; k < s->n_dev; k++)
      if (gfsb200_ctx_create (device + k, &s->dev[k]) != GFSB200_OK)
	g_error ("particulates (B200): %s", gfsb200_last_error ());   /* no CPU fallback */
    s->ctx = s->dev[0];
    if (s->n_dev > 1) {
      if (gfsb200_comm_init_all (s->n_dev, s->dev, s->comm) != GFSB200_OK)
	g_error ("particulates (B200): %s", gfsb200_last_error ());
      /* the smoothed deposit of GfsSourceParticulate reaches cells of every slice: whole-buffer sums */
      for (k = 0; k < s->n_dev; k++)
	if (gfsb200_comm_set_exchange (s->comm[k], GFSB200_EXCHANGE_ALLREDUCE) != GFSB200_OK)
	  g_error ("particulates (B200): %s", gfsb200_last_error ());
    }
    s->plist = plist;
    /* the device copy is authoritative between events unless GFSB200_RESIDENT=0 */
    s->resident = !(g_getenv ("GFSB200_RESIDENT") && atoi (g_getenv ("GFSB200_RESIDENT")) == 0);
    g_hash_table_insert (b200_states, plist, s);
  }
  return s;
}

static B200State * peek_state (gpointer plist)
{
  return b200_states ? g_hash_table_lookup (b200_states, plist) : NULL;
}

static void sync_down (B200State * s);

/* ------------------------------------------------------------------ */
/* Walking the GSList of a large list is a serial pointer chase (25-40 ns per node)
 * and, once the gathers are spread over the cores, the most expensive part of an
 * event.  The walk is skipped while the list is known not to have changed: the
 * GtsSListContainer holding the particles is given a subclass whose add / remove
 * methods -- the only ways the reference changes membership (gts_container_add in
 * add_particulate :1253-1255, gfs_particle_bc :3207-3210, mpi_rcv_particle;
 * gts_container_remove in the cull :965, the BCs :3335 and slist_containee_destroy)
 * -- flag the owning list first.  Order-only edits (the g_slist_reverse pairs of
 * :1253-1255) always bracket an add. */
static GHashTable * b200_watched = NULL;   /* GtsSListContainer* -> B200State* */

static GtsSListContainerClass * watched_container_class (void);

static void watched_changed (GtsContainer * c)
{
  B200State * s = b200_watched ? g_hash_table_lookup (b200_watched, c) : NULL;
  if (s) {
    /* resident mode: bring the objects up to date while obj[] still describes the list */
    if (s->host_stale && !s->syncing)
      sync_down (s);
    s->list_known = FALSE;
  }
}

static void watched_add (GtsContainer * c, GtsContainee * item)
{
  watched_changed (c);
  (* GTS_CONTAINER_CLASS (GTS_OBJECT_CLASS (watched_container_class ())->parent_class)->add) (c, item);
}

static void watched_remove (GtsContainer * c, GtsContainee * item)
{
  watched_changed (c);
  (* GTS_CONTAINER_CLASS (GTS_OBJECT_CLASS (watched_container_class ())->parent_class)->remove) (c, item);
}

static void watched_container_class_init (GtsContainerClass * klass)
{
  klass->add = watched_add;
  klass->remove = watched_remove;
}

static GtsSListContainerClass * watched_container_class (void)
{
  static GtsSListContainerClass * klass = NULL;
  if (klass == NULL) {
    GtsObjectClassInfo info = {
      "GfsB200WatchedList",
      sizeof (GtsSListContainer),
      sizeof (GtsSListContainerClass),
      (GtsObjectClassInitFunc) watched_container_class_init,
      (GtsObjectInitFunc) NULL,
      (GtsArgSetFunc) NULL,
      (GtsArgGetFunc) NULL
    };
    klass = gts_object_class_new (GTS_OBJECT_CLASS (gts_slist_container_class ()), &info);
  }
  return klass;
}

static void watch_list (B200State * s, GfsParticleList * plist)
{
  GtsSListContainer * list = GFS_EVENT_LIST (plist)->list;
  if (GTS_OBJECT (list)->klass == GTS_OBJECT_CLASS (watched_container_class ()))
    return;
  if (GTS_OBJECT (list)->klass != GTS_OBJECT_CLASS (gts_slist_container_class ()) ||
      g_getenv ("GFSB200_WALK_LIST"))
    return;                               /* someone else's subclass: keep walking every event */
  if (!b200_watched)
    b200_watched = g_hash_table_new (NULL, NULL);
  g_hash_table_insert (b200_watched, list, s);
  GTS_OBJECT (list)->klass = GTS_OBJECT_CLASS (watched_container_class ());
  s->list_known = FALSE;
}

/* the list's destroy method (gfs_particle_list_destroy, :1121-1130) with the
 * device state released first: a later list allocated at the same address must
 * not inherit this one's context and flat tree */
static void (* reference_list_destroy) (GtsObject *) = NULL;

static void b200_particle_list_destroy (GtsObject * o)
{
  B200State * s = b200_states ? g_hash_table_lookup (b200_states, o) : NULL;
  if (s) {
    FttComponent c;
    gint k;
    g_hash_table_remove (b200_states, o);
    if (b200_watched)
      g_hash_table_remove (b200_watched, GFS_EVENT_LIST (o)->list);
    if (g_getenv ("GFSB200_MODULE_PROFILE") && s->n_events)
      fprintf (stderr, "particulates (B200): %u list events after the first; ms/event: tree %.3f, field gather+upload %.3f, "
	       "particle upload %.3f, cull+step %.3f, particle download %.3f, gfs_particle_bc %.3f\n",
	       s->n_events, 1e3*s->phase[PH_TREE]/s->n_events, 1e3*s->phase[PH_FIELD]/s->n_events,
	       1e3*s->phase[PH_UPLOAD]/s->n_events, 1e3*s->phase[PH_DEVICE]/s->n_events,
	       1e3*s->phase[PH_DOWNLOAD]/s->n_events, 1e3*s->phase[PH_BC]/s->n_events);
    if (s->map) gfsb200_ftt_map_free (s->map);
    if (s->tree) gfsb200_tree_free (s->tree);
    for (k = 0; k < s->n_dev; k++)
      if (s->comm[k]) gfsb200_comm_destroy (s->comm[k]);
    for (k = 0; k < s->n_dev; k++)
      if (s->dev[k]) gfsb200_ctx_destroy (s->dev[k]);
    g_free (s->idx_of_id);
    for (c = 0; c < 3; c++)
      gfsb200_host_free (s->field[c]);
    gfsb200_host_free (s->cellvar[0]);
    gfsb200_host_free (s->cellvar[1]);
    gfsb200_host_free (s->dep_out);
    for (k = 0; k < 10; k++)
      gfsb200_host_free (s->col[k]);
    gfsb200_host_free (s->id);
    g_free (s->obj);
    g_free (s);
  }
  (* reference_list_destroy) (o);
}

/* ------------------------------------------------------------------ */
/* mesh: flatten after every adapt                                      */

static gpointer pinned (gsize bytes)
{
  gpointer p = gfsb200_host_alloc (bytes);
  if (!p)
    g_error ("particulates (B200): %s", gfsb200_last_error ());
  return p;
}

/* Mesh identity.  sim->adapts_stats is not a signature: GfsOutputAdaptStats resets it after every
 * output (src/output.c:688) and two adapts can produce the same counts.  Instead the module owns a
 * hidden (unnamed, never written) GfsVariable per domain whose callbacks Gerris runs for EVERY cell
 * it refines (coarse_fine, from gfs_cell_fine_init, src/domain.c:2979-3010) or destroys (cleanup,
 * from gfs_cell_cleanup, src/fluid.c:1965-1985 -- coarsening, box migration, domain destruction):
 * each bumps an epoch.  Boxes that ARRIVE (GfsEventBalance) are caught by the root signature. */
static guint b200_mesh_epoch = 1;
static GHashTable * b200_mesh_watch = NULL;    /* GfsDomain* -> its hidden GfsVariable */

static void mesh_refined (FttCell * parent, GfsVariable * v) { b200_mesh_epoch++; }
static void mesh_cell_destroyed (FttCell * cell, GfsVariable * v) { b200_mesh_epoch++; }
static void mesh_none (FttCell * cell, GfsVariable * v) {}

static void watch_mesh (GfsDomain * domain)
{
  GfsVariable * v;
  if (!b200_mesh_watch)
    b200_mesh_watch = g_hash_table_new (NULL, NULL);
  v = g_hash_table_lookup (b200_mesh_watch, domain);
  if (v && g_slist_find (domain->variables, v))
    return;
  v = gfs_domain_add_variable (domain, NULL, "B200 mesh watch");
  if (v == NULL)
    g_error ("particulates (B200): cannot add the mesh-watch variable");
  v->coarse_fine = mesh_refined;
  v->fine_coarse = mesh_none;
  v->cleanup = (FttCellCleanupFunc) mesh_cell_destroyed;
  g_hash_table_insert (b200_mesh_watch, domain, v);
  b200_mesh_epoch++;                    /* whatever happened before the watch started is unknown */
}

typedef struct { GPtrArray * roots; GArray * is_box; } RootList;

static void collect_roots (GfsBox * box, RootList * r)
{
  FttDirection d;
  gint one = 1, zero = 0;
  g_ptr_array_add (r->roots, box->root);
  g_array_append_val (r->is_box, one);
  for (d = 0; d < FTT_NEIGHBORS; d++)
    if (GFS_IS_BOUNDARY (box->neighbor[d]) && GFS_BOUNDARY (box->neighbor[d])->root) {
      g_ptr_array_add (r->roots, GFS_BOUNDARY (box->neighbor[d])->root);
      g_array_append_val (r->is_box, zero);
    }
}

static void refresh_tree (B200State * s, GfsSimulation * sim)
{
  GfsDomain * domain = GFS_DOMAIN (sim);
  RootList r = { g_ptr_array_new (), g_array_new (FALSE, FALSE, sizeof (gint)) };
  FttComponent c;
  gsize hash = 0;
  guint k;
  watch_mesh (domain);
  gts_container_foreach (GTS_CONTAINER (domain), (GtsFunc) collect_roots, &r);
  for (k = 0; k < r.roots->len; k++)
    hash = hash*31 + (gsize) r.roots->pdata[k];
  if (s->tree_valid && s->mesh_epoch == b200_mesh_epoch && s->n_roots == r.roots->len && s->root_hash == hash) {
    g_ptr_array_free (r.roots, TRUE);
    g_array_free (r.is_box, TRUE);
    return;
  }
  /* the flat tree's map points at FttCells of the old mesh: nothing may use it from here on */
  s->tree_valid = FALSE;
  if (s->map) gfsb200_ftt_map_free (s->map);
  if (s->tree) gfsb200_tree_free (s->tree);
  s->map = NULL; s->tree = NULL;
  if (gfsb200_ftt_flatten (r.roots->len, (void * const *) r.roots->pdata, (const int *) r.is_box->data,
			   &s->tree, &s->map) != GFSB200_OK)
    g_error ("particulates (B200): %s", gfsb200_ftt_last_error ());
  s->n_roots = r.roots->len;
  s->root_hash = hash;
  g_ptr_array_free (r.roots, TRUE);
  g_array_free (r.is_box, TRUE);
  if (gfsb200_tree_build_stencils (s->tree) != GFSB200_OK)
    g_error ("particulates (B200): %s", gfsb200_last_error ());
  for (k = 0; k < (guint) s->n_dev; k++)
    if (gfsb200_upload_tree (s->dev[k], s->tree) != GFSB200_OK)
      g_error ("particulates (B200): %s", gfsb200_last_error ());
  s->n_cells = gfsb200_ftt_map_size (s->map);
  if (s->n_cells > s->field_cap) {
    for (c = 0; c < 3; c++) {
      gfsb200_host_free (s->field[c]);
      s->field[c] = pinned (sizeof (gdouble)*s->n_cells);
    }
    for (c = 0; c < 2; c++) {
      gfsb200_host_free (s->cellvar[c]);
      s->cellvar[c] = pinned (sizeof (gdouble)*s->n_cells);
    }
    s->field_cap = s->n_cells;
  }
  s->mesh_epoch = b200_mesh_epoch;
  s->tree_valid = TRUE;
}

/* U,V,W of time t^n, ghost cells included (gfs_domain_bc has been applied by
 * the solver before events run, src/simulation.c:483) */
static void mirror_velocity (B200State * s, GfsDomain * domain)
{
  GfsVariable ** u = gfs_domain_velocity (domain);
  const size_t off = offsetof (GfsStateVector, place_holder);
  int var[5];
  double nodata[5], * out[5];
  int n = 0;
  FttComponent c;
  /* U,V,W and, when they are variables, the per-cell alpha (fluid density 1/alpha, :534-535)
     and viscosity (gfs_diffusion_cell, :540-541): ONE parallel pass over the cells */
  for (c = 0; c < FTT_DIMENSION; c++) {
    var[n] = u[c]->i; nodata[n] = GFS_NODATA; out[n++] = s->field[c];
  }
  if (s->alpha_var) {
    var[n] = s->alpha_var->i; nodata[n] = 1.; out[n++] = s->cellvar[0];
  }
  if (s->mu_var) {
    var[n] = s->mu_var->i; nodata[n] = 0.; out[n++] = s->cellvar[1];
  }
  /* the data block of every cell is looked up once per flatten (and again when Gerris has moved
     the blocks to make room for a new variable): the per-event gather never touches the tree */
  if (gfsb200_ftt_map_cache_data (s->map, 1 + (domain->allocated ? domain->allocated->len : 0)) != GFSB200_OK)
    g_error ("particulates (B200): %s", gfsb200_ftt_last_error ());
  if (s->n_dev > 1) {
    /* one gather, ONE upload (device 0), NVLink to the other GPUs, cell pass everywhere */
    gfsb200_ftt_gather_many (s->map, off, n, var, nodata, out);
    if (gfsb200_broadcast_field (s->comm, s->n_dev, 0, s->field[0], s->field[1],
				 FTT_DIMENSION > 2 ? s->field[2] : NULL,
				 s->alpha_var ? s->cellvar[0] : NULL, s->mu_var ? s->cellvar[1] : NULL) != GFSB200_OK)
      g_error ("particulates (B200): %s", gfsb200_last_error ());
  }
  else {
    /* in slices: while the copy engine ships slice k (page-locked staging, asynchronous), the host
       cores gather slice k + 1 out of the tree -- the upload hides behind the gather */
    const gint32 slices = s->n_cells > (1 << 18) ? 4 : 1;
    gint32 k;
    for (k = 0; k < slices; k++) {
      const gint32 first = (gint32) ((gint64) s->n_cells*k/slices), last = (gint32) ((gint64) s->n_cells*(k + 1)/slices);
      gfsb200_ftt_gather_range (s->map, off, first, last, n, var, nodata, out);
      if (gfsb200_upload_field_part (s->ctx, first, last - first, s->field[0], s->field[1],
				     FTT_DIMENSION > 2 ? s->field[2] : NULL,
				     s->alpha_var ? s->cellvar[0] : NULL, s->mu_var ? s->cellvar[1] : NULL) != GFSB200_OK)
	g_error ("particulates (B200): %s", gfsb200_last_error ());
    }
    if (gfsb200_refresh_field (s->ctx) != GFSB200_OK)
      g_error ("particulates (B200): %s", gfsb200_last_error ());
  }
  if (s->uold) {                          /* Un,Vn,Wn of GfsForceInertial / GfsForceAddedMass */
    for (c = 0; c < FTT_DIMENSION; c++) {
      var[c] = s->uold[c]->i; nodata[c] = GFS_NODATA; out[c] = s->field[c];
    }
    /* the staging buffers are reused: the first upload must have left them */
    gint d;
    for (d = 0; d < s->n_dev; d++)
      if (gfsb200_ctx_synchronize (s->dev[d]) != GFSB200_OK)
	g_error ("particulates (B200): %s", gfsb200_last_error ());
    gfsb200_ftt_gather_many (s->map, off, FTT_DIMENSION, var, nodata, out);
    for (d = 0; d < s->n_dev; d++)
      if (gfsb200_upload_field_prev (s->dev[d], s->field[0], s->field[1],
				     FTT_DIMENSION > 2 ? s->field[2] : NULL) != GFSB200_OK)
	g_error ("particulates (B200): %s", gfsb200_last_error ());
  }
}

/* store_domain_previous_vel, modules/particulatecommon.c:91-113 (static there) */
static void copy_cell_variable (FttCell * cell, gpointer * data)
{
  GFS_VALUE (cell, (GfsVariable *) data[0]) = GFS_VALUE (cell, (GfsVariable *) data[1]);
}

static void store_previous_velocity (GfsDomain * domain, GfsVariable ** un)
{
  GfsVariable ** u = gfs_domain_velocity (domain);
  FttComponent c;
  for (c = 0; c < FTT_DIMENSION; c++) {
    gpointer data[2] = { un[c], u[c] };
    gfs_domain_cell_traverse (domain, FTT_PRE_ORDER, FTT_TRAVERSE_LEAFS, -1,
			      (FttCellTraverseFunc) copy_cell_variable, data);
    gfs_domain_bc (domain, FTT_TRAVERSE_LEAFS, -1, un[c]);
  }
}

/* ------------------------------------------------------------------ */
/* parameters: which forces, rho, mu, g                                 */

static GfsSourceDiffusion * viscosity_source (GfsVariable * v)
{
  GSList * i = v->sources ? GTS_SLIST_CONTAINER (v->sources)->items : NULL;
  while (i) {
    if (GFS_IS_SOURCE_DIFFUSION (i->data))
      return GFS_SOURCE_DIFFUSION (i->data);
    i = i->next;
  }
  return NULL;
}

/* fills *p; returns FALSE if the list needs something only the reference
 * CPU path implements */
/* Embedded solids: the mixed cells travel with the tree (gfsb200_ftt_flatten carries each
 * GfsSolidVector: centre of mass into the interpolation weights, face fractions into the
 * gradients, fluid fraction into the deposits), which is re-flattened after an adapt only.
 * Solids that MOVE (GfsSolidMoving, modules/moving.c) change those fractions every step
 * without touching the mesh: such a simulation stays on the reference's own events. */
static gboolean moving_solids (GfsSimulation * sim)
{
  GSList * i = sim->solids ? sim->solids->items : NULL;
  for (; i; i = i->next) {
    GtsObjectClass * k = GTS_OBJECT (i->data)->klass;
    for (; k; k = k->parent_class)
      if (!strcmp (k->info.name, "GfsSolidMoving"))
	return TRUE;
  }
  return FALSE;
}

typedef struct {
  GfsVariable ** uold;                  /* -> B200State.uold */
  gboolean snapshot;
  GfsVariable * alpha_var, * mu_var;
} ListVars;

static void adopt (B200State * s, const ListVars * lv)
{
  s->uold = lv->uold;
  s->snapshot = lv->snapshot;
  s->alpha_var = lv->alpha_var;
  s->mu_var = lv->mu_var;
}

/* (reads the objects only: no device context is created for a list that ends up
   on the reference path) */
static const gchar * step_params_why (GfsParticleList * plist, GfsSimulation * sim, gfsb200_step_params * p,
				      ListVars * lv);

static gboolean step_params (GfsParticleList * plist, GfsSimulation * sim, gfsb200_step_params * p,
			     ListVars * lv, const gchar ** why)
{
  *why = step_params_why (plist, sim, p, lv);
  return *why == NULL;
}

/* NULL if the list can run on the device, else what it carries that cannot */
static const gchar * step_params_why (GfsParticleList * plist, GfsSimulation * sim, gfsb200_step_params * p,
				      ListVars * lv)
{
  GfsDomain * domain = GFS_DOMAIN (sim);
  GfsVariable ** u = gfs_domain_velocity (domain);
  GSList * i = plist->forces ? plist->forces->items : NULL;
  FttComponent c;

  gfsb200_step_params_default (p);
  memset (lv, 0, sizeof (ListVars));
  p->dt = sim->advection_params.dt;
  while (i) {
    GfsForceCoeff * coeff = GFS_IS_FORCE_COEFF (i->data) ? FORCE_COEFF (i->data) : NULL;
    gdouble k = coeff && coeff->coefficient ? gfs_function_get_constant_value (coeff->coefficient) : 0.;
    if (p->n_forces == GFSB200_MAX_FORCES)
      return "a force list longer than GFSB200_MAX_FORCES";
    if (k == G_MAXDOUBLE)
      return "a force coefficient GfsFunction that is not a constant (modules/particulatecommon.c:566-575)";
    if (GFS_IS_FORCE_DRAG (i->data)) {
      p->force[p->n_forces++] = GFSB200_FORCE_DRAG;
      if (coeff->coefficient) p->cd_const = k;
    }
    else if (GFS_IS_FORCE_ADDEDMASS (i->data)) {
      p->force[p->n_forces++] = GFSB200_FORCE_ADDEDMASS;
      if (coeff->coefficient) p->cm_const = k;
      lv->uold = coeff->Uold;
    }
    else if (GFS_IS_FORCE_INERTIAL (i->data)) {
      p->force[p->n_forces++] = GFSB200_FORCE_INERTIAL;
      lv->uold = coeff->Uold;
      lv->snapshot = TRUE;
    }
    else if (GFS_IS_FORCE_LIFT (i->data)) {
      p->force[p->n_forces++] = GFSB200_FORCE_LIFT;
      if (coeff->coefficient) p->cl_const = k;
    }
    else if (GFS_IS_FORCE_BUOY (i->data))
      p->force[p->n_forces++] = GFSB200_FORCE_BUOY;
    else
      return "a user-defined force class";
    i = i->next;
  }
  /* fluid density 1/alpha (particulatecommon.c:534-535) */
  if (sim->physical_params.alpha) {
    gdouble a = gfs_function_get_constant_value (sim->physical_params.alpha);
    if (a == G_MAXDOUBLE) {
      /* alpha = <a cell variable>: mirrored per cell; any other expression stays on the host */
      GfsVariable * v = gfs_function_get_variable (sim->physical_params.alpha);
      if (v == NULL)
	return "PhysicalParams alpha given as an expression (neither a constant nor a plain variable)";
      lv->alpha_var = v;
    }
    else
      p->rho = 1./a;
  }
  /* viscosity of the GfsSourceDiffusion on U (particulatecommon.c:540-541): the per-cell
     variable GfsDiffusion keeps when its function is not a constant, else the constant */
  {
    GfsSourceDiffusion * d = viscosity_source (u[0]);
    if (d) {
      if (d->D->mu)
	lv->mu_var = d->D->mu;
      else {
	gdouble mu = gfs_function_get_constant_value (d->D->val);
	p->mu = mu < G_MAXDOUBLE ? mu : 0.;                  /* diffusion_cell, src/source.c:941-946 */
      }
    }
  }
  /* g = sum of the GfsSource intensities on U,V,W (particulatecommon.c:634-649) */
  for (c = 0; c < FTT_DIMENSION; c++) {
    GSList * j = u[c]->sources ? GTS_SLIST_CONTAINER (u[c]->sources)->items : NULL;
    while (j) {
      if (GFS_IS_SOURCE (j->data)) {
	gdouble g = gfs_function_get_constant_value (GFS_SOURCE (j->data)->intensity);
	if (g == G_MAXDOUBLE)
	  return "a GfsSource on the velocity whose intensity is not a constant";
	p->g[c] += g;
      }
      j = j->next;
    }
  }
#ifdef HAVE_MPI
  if (domain->pid >= 0)
    return "an MPI domain decomposition (the device path replicates the field and shards the particles "
      "over the GPUs of one box: run gerris serially)";
#endif
  return NULL;
}

/* ------------------------------------------------------------------ */
/* particle objects <-> device SoA                                      */

/* the list's objects as an array, in list order (ONE walk of the GSList: it is a serial
   pointer chase, the only part of the event that cannot be spread over the cores);
   grows the page-locked columns */
static gint64 collect_particles (B200State * s, GfsParticleList * plist)
{
  GSList * i = GFS_EVENT_LIST (plist)->list->items;
  gint64 n = 0;
  gint c;
  watch_list (s, plist);
  if (s->list_known)
    return s->n_obj;
  for (; i; i = i->next) {
    if (n == s->obj_cap) {
      s->obj_cap = s->obj_cap ? 2*s->obj_cap : 4096;
      s->obj = g_realloc (s->obj, sizeof (GfsParticulate *)*s->obj_cap);
    }
    s->obj[n++] = i->data;
  }
  if (n > s->part_cap) {
    gint64 cap = n + n/8 + 1024;
    for (c = 0; c < 10; c++) {
      gfsb200_host_free (s->col[c]);
      s->col[c] = pinned (sizeof (gdouble)*cap);
    }
    gfsb200_host_free (s->id);
    s->id = pinned (sizeof (guint32)*cap);
    s->part_cap = cap;
  }
  s->n_obj = n;
  /* (stays FALSE when the container could not be watched) */
  s->list_known = GTS_OBJECT (GFS_EVENT_LIST (plist)->list)->klass == GTS_OBJECT_CLASS (watched_container_class ());
  return n;
}

static gint64 upload_particles (B200State * s, GfsParticleList * plist)
{
  gint64 n = collect_particles (s, plist), k;
  gdouble ** col = s->col;
  /* the objects are scattered over the heap (one cache miss or two each): the gather is
     spread over the host cores */
#pragma omp parallel for schedule(static)
  for (k = 0; k < n; k++) {
    GfsParticle * p = GFS_PARTICLE (s->obj[k]);
    GfsParticulate * q = s->obj[k];
    p->pos_old = p->pos;                                     /* :804-805 */
    col[0][k] = p->pos.x; col[1][k] = p->pos.y; col[2][k] = p->pos.z;
    col[3][k] = q->vel.x; col[4][k] = q->vel.y; col[5][k] = q->vel.z;
    col[6][k] = q->mass;  col[7][k] = q->volume;
    s->id[k] = p->id;
  }
  if (s->n_dev == 1) {
    if (gfsb200_particles_upload (s->ctx, n, col[0], col[1], FTT_DIMENSION > 2 ? col[2] : NULL,
				  col[3], col[4], FTT_DIMENSION > 2 ? col[5] : NULL,
				  col[6], col[7], s->id) != GFSB200_OK)
      g_error ("particulates (B200): %s", gfsb200_last_error ());
    return n;
  }
  /* several GPUs: equal slices of the list go up, then the devices sort themselves into the
     global cell order (every GPU ends with the particles of a contiguous range of cells).  The
     devices no longer hold the list in object order: objects are found again by particle id. */
  {
    guint32 max_id = 0;
    gint d;
    for (k = 0; k < n; k++)
      if (s->id[k] > max_id) max_id = s->id[k];
    if ((gint64) max_id > 16*n + (1 << 20))
      g_error ("particulates (B200): particle ids up to %u for %lld particles: too sparse for the id table "
	       "of the multi-GPU mode", max_id, (long long) n);
    if (max_id + 1 > s->idx_cap) {
      g_free (s->idx_of_id);
      s->idx_cap = max_id + 1 + max_id/8;
      s->idx_of_id = g_malloc (sizeof (gint32)*s->idx_cap);
    }
    memset (s->idx_of_id, 0xff, sizeof (gint32)*s->idx_cap);
    for (k = 0; k < n; k++) {
      if (s->idx_of_id[s->id[k]] >= 0)
	g_error ("particulates (B200): particle id %u appears twice in the list", s->id[k]);
      s->idx_of_id[s->id[k]] = (gint32) k;
    }
    for (d = 0; d < s->n_dev; d++) {
      const gint64 lo = n*d/s->n_dev, m = n*(d + 1)/s->n_dev - lo;
      if (gfsb200_particles_upload (s->dev[d], m, col[0] + lo, col[1] + lo, FTT_DIMENSION > 2 ? col[2] + lo : NULL,
				    col[3] + lo, col[4] + lo, FTT_DIMENSION > 2 ? col[5] + lo : NULL,
				    col[6] + lo, col[7] + lo, s->id + lo) != GFSB200_OK)
	g_error ("particulates (B200): %s", gfsb200_last_error ());
    }
    if (gfsb200_comm_rebalance (s->comm, s->n_dev) != GFSB200_OK)
      g_error ("particulates (B200): %s", gfsb200_last_error ());
  }
  return n;
}

/* write the device state back into the GtsObjects (s->obj holds them in upload order,
 * which the device keeps); particles the device culled (outside the domain) are removed
 * from the list as remove_particles_not_in_domain does (:955-969) */
static void download_particles_multi (B200State * s, GfsParticleList * plist, gint64 n_up);

static void download_particles (B200State * s, GfsParticleList * plist, gint64 n_up)
{
  gint64 n, k, j;
  gdouble ** col = s->col;
  if (s->n_dev > 1) {
    download_particles_multi (s, plist, n_up);
    return;
  }
  n = gfsb200_particles_count (s->ctx);
  if (gfsb200_particles_download (s->ctx, col[0], col[1], FTT_DIMENSION > 2 ? col[2] : NULL,
				  col[3], col[4], FTT_DIMENSION > 2 ? col[5] : NULL,
				  col[6], col[7], col[8], col[9], NULL, s->id, NULL) != GFSB200_OK)
    g_error ("particulates (B200): %s", gfsb200_last_error ());
  if (n != n_up) {
    /* culled on the device: drop the same objects from the list, order preserved */
    for (k = 0, j = 0; k < n_up; k++) {
      GfsParticle * p = GFS_PARTICLE (s->obj[k]);
      if (j < n && s->id[j] == p->id)
	s->obj[j++] = s->obj[k];
      else {
	gts_container_remove (GTS_CONTAINER (GFS_EVENT_LIST (plist)->list), GTS_CONTAINEE (p));
	gts_object_destroy (GTS_OBJECT (p));
      }
    }
    if (j != n)
      g_error ("particulates (B200): the device list and the GfsParticleList disagree");
    /* the removals above flagged the list; obj[0..n) is exactly what is left of it */
    s->n_obj = n;
    s->list_known = GTS_OBJECT (GFS_EVENT_LIST (plist)->list)->klass ==
      GTS_OBJECT_CLASS (watched_container_class ());
  }
#pragma omp parallel for schedule(static)
  for (k = 0; k < n; k++) {
    GfsParticle * p = GFS_PARTICLE (s->obj[k]);
    GfsParticulate * q = s->obj[k];
    p->pos.x = col[0][k]; p->pos.y = col[1][k];
    q->vel.x = col[3][k]; q->vel.y = col[4][k];
    q->force.x = col[6][k]; q->force.y = col[7][k];
    q->mass = col[9][k];                   /* GfsForceAddedMass updates it every step (:391) */
#if !FTT_2D
    p->pos.z = col[2][k]; q->vel.z = col[5][k]; q->force.z = col[8][k];
#endif
  }
}

/* the same over several GPUs: the devices' lists, one behind the other, in whatever order the
 * rebalance left them; every row finds its object through the id table */
static void download_particles_multi (B200State * s, GfsParticleList * plist, gint64 n_up)
{
  gint64 n = 0, k, j;
  gdouble ** col = s->col;
  gint d;
  for (d = 0; d < s->n_dev; d++) {
    const gint64 m = gfsb200_particles_count (s->dev[d]);
    if (n + m > s->part_cap)
      g_error ("particulates (B200): the devices hold more particles than the list");
    if (gfsb200_particles_download (s->dev[d], col[0] + n, col[1] + n, FTT_DIMENSION > 2 ? col[2] + n : NULL,
				    col[3] + n, col[4] + n, FTT_DIMENSION > 2 ? col[5] + n : NULL,
				    col[6] + n, col[7] + n, col[8] + n, col[9] + n, NULL, s->id + n, NULL) != GFSB200_OK)
      g_error ("particulates (B200): %s", gfsb200_last_error ());
    n += m;
  }
#pragma omp parallel for schedule(static)
  for (k = 0; k < n; k++) {
    const gint32 o = s->id[k] < s->idx_cap ? s->idx_of_id[s->id[k]] : -1;
    if (o >= 0) {
      GfsParticle * p = GFS_PARTICLE (s->obj[o]);
      GfsParticulate * q = s->obj[o];
      p->pos.x = col[0][k]; p->pos.y = col[1][k];
      q->vel.x = col[3][k]; q->vel.y = col[4][k];
      q->force.x = col[6][k]; q->force.y = col[7][k];
      q->mass = col[9][k];
#if !FTT_2D
      p->pos.z = col[2][k]; q->vel.z = col[5][k]; q->force.z = col[8][k];
#endif
    }
  }
  if (n != n_up) {
    /* culled on the devices: the objects whose id did not come back leave the list, order preserved */
    guint8 * seen = g_malloc0 (n_up ? n_up : 1);
    for (k = 0; k < n; k++) {
      const gint32 o = s->id[k] < s->idx_cap ? s->idx_of_id[s->id[k]] : -1;
      if (o < 0 || o >= n_up || seen[o])
	g_error ("particulates (B200): the device lists and the GfsParticleList disagree");
      seen[o] = 1;
    }
    for (k = 0, j = 0; k < n_up; k++) {
      GfsParticle * p = GFS_PARTICLE (s->obj[k]);
      if (seen[k]) {
	s->idx_of_id[p->id] = (gint32) j;
	s->obj[j++] = s->obj[k];
      }
      else {
	s->idx_of_id[p->id] = -1;
	gts_container_remove (GTS_CONTAINER (GFS_EVENT_LIST (plist)->list), GTS_CONTAINEE (p));
	gts_object_destroy (GTS_OBJECT (p));
      }
    }
    g_free (seen);
    s->n_obj = n;
    s->list_known = GTS_OBJECT (GFS_EVENT_LIST (plist)->list)->klass ==
      GTS_OBJECT_CLASS (watched_container_class ());
  }
}

/* resident mode: refresh the GtsObjects from the device (positions, velocities, forces,
 * mass; pos_old is NOT maintained between syncs -- only gfs_particle_bc reads it, and the
 * event patches it for the particles concerned) */
static void sync_down (B200State * s)
{
  if (!s->host_stale)
    return;
  s->syncing = TRUE;
  download_particles (s, s->plist, s->n_obj);
  s->syncing = FALSE;
  s->host_stale = FALSE;
}

static void sync_one (gpointer key, gpointer value, gpointer data)
{
  sync_down (value);
}

static void sync_all (void)
{
  if (b200_states)
    g_hash_table_foreach (b200_states, sync_one, NULL);
}

/* For host code outside this module that reads GfsParticulate objects of a list driven in
 * resident mode (another GModule, a user event): call this first. */
void gfsb200_module_sync (GfsParticleList * plist);
void gfsb200_module_sync (GfsParticleList * plist)
{
  B200State * s = peek_state (plist);
  if (s)
    sync_down (s);
}

/* make sure the device holds the list's current state; returns the number of particles */
static gint64 device_current (B200State * s, GfsParticleList * plist)
{
  watch_list (s, plist);
  s->uploaded = FALSE;
  if (s->resident && s->host_stale && s->list_known)
    return s->n_obj;                      /* nothing on the host has changed since the last step */
  if (s->host_stale)
    sync_down (s);
  s->uploaded = TRUE;
  return upload_particles (s, plist);
}

/* the reference's own readers and writers of particle objects (every class of
 * particulatecommon.c that looks at a list it does not own): objects first */
#define N_READERS 6
static gboolean (* reader_event[N_READERS]) (GfsEvent *, GfsSimulation *);
static GfsEventClass * reader_class[N_READERS];
static void (* reference_list_write) (GtsObject *, FILE *) = NULL;

static gboolean synced_reader_event (GfsEvent * event, GfsSimulation * sim)
{
  GtsObjectClass * klass = GTS_OBJECT (event)->klass;
  gint k;
  sync_all ();
  /* the method of the nearest wrapped ancestor */
  for (; klass; klass = klass->parent_class)
    for (k = 0; k < N_READERS; k++)
      if (klass == GTS_OBJECT_CLASS (reader_class[k]))
	return (* reader_event[k]) (event, sim);
  g_assert_not_reached ();
  return FALSE;
}

static void b200_particle_list_write (GtsObject * o, FILE * fp)
{
  B200State * s = peek_state (o);
  if (s)
    sync_down (s);
  (* reference_list_write) (o, fp);
}

/* ------------------------------------------------------------------ */
/* the replaced event methods                                           */

/* resident mode: pos_old of the particles that just left the domain, from the device's record
 * (list positions refer to the list as it was stepped, i.e. to obj[] before any compaction);
 * gfs_particle_bc walks back from there (:3151-3186) */
static void patch_pos_old (B200State * s, gint64 escaped)
{
  gint d;
  gint32 * idx = g_malloc (sizeof (gint32)*(escaped ? escaped : 1));
  gdouble * old = g_malloc (sizeof (gdouble)*3*(escaped ? escaped : 1));
  for (d = 0; d < s->n_dev; d++) {
    gint64 got = 0, e;
    guint32 * ids = NULL;
    if (gfsb200_escaped_download (s->dev[d], escaped, idx, old, &got) != GFSB200_OK)
      g_error ("particulates (B200): %s", gfsb200_last_error ());
    if (got == 0)
      continue;
    if (s->n_dev > 1) {
      /* list positions are the device's: its ids lead to the objects */
      const gint64 m = gfsb200_particles_count (s->dev[d]);
      ids = g_malloc (sizeof (guint32)*(m ? m : 1));
      if (gfsb200_particles_download (s->dev[d], NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL,
				      NULL, NULL, ids, NULL) != GFSB200_OK)
	g_error ("particulates (B200): %s", gfsb200_last_error ());
      for (e = 0; e < got; e++)
	if (idx[e] >= 0 && idx[e] < m) {
	  const guint32 id = ids[idx[e]];
	  idx[e] = id < s->idx_cap ? s->idx_of_id[id] : -1;
	}
	else
	  idx[e] = -1;
      g_free (ids);
    }
    for (e = 0; e < got; e++)
      if (idx[e] >= 0 && idx[e] < s->n_obj) {
	GfsParticle * p = GFS_PARTICLE (s->obj[idx[e]]);
	p->pos_old.x = old[3*e]; p->pos_old.y = old[3*e + 1]; p->pos_old.z = old[3*e + 2];
      }
  }
  g_free (idx);
  g_free (old);
}

static gboolean b200_particle_list_event (GfsEvent * event, GfsSimulation * sim)
{
  GfsParticleList * plist = GFS_PARTICLE_LIST (event);
  gfsb200_step_params par;
  ListVars lv;
  B200State * s;
  gint64 removed = 0, escaped = 0, n_up;
  gboolean carried;
  gdouble t[7];
  gint k;

  const gchar * why = NULL;
  if (moving_solids (sim))
    why = "a simulation with moving solids (GfsSolidMoving changes the cell fractions every step)";
  if (why || !step_params (plist, sim, &par, &lv, &why)) {
    reference_or_error (plist, "GfsParticleList", why);
    gfsb200_module_sync (plist);
    return (* reference_list_event) (event, sim);            /* opted in: GFSB200_ALLOW_REFERENCE_EVENT */
  }

  /* the timing gate of gfs_event_list_event (src/event.c:2430-2439) */
  if (!(* GFS_EVENT_CLASS (gfs_event_class ())->event) (event, sim))
    return FALSE;

  s = state_of (plist);
  adopt (s, &lv);
  t[0] = wall ();
  refresh_tree (s, sim);
  t[1] = wall ();
  mirror_velocity (s, GFS_DOMAIN (sim));
  t[2] = wall ();
  /* By default the host objects are authoritative between events (FeedParticle,
     DropletToParticle, outputs and BCs all mutate them): gather + upload here, download +
     scatter below.  In resident mode (GFSB200_RESIDENT=1) device_current skips the upload
     when nothing on the host has touched the list since the last event. */
  carried = s->resident && s->host_stale && s->list_known;   /* what device_current is about to decide */
  n_up = device_current (s, plist);
  carried = carried && !s->uploaded;
  t[3] = wall ();
  par.record_forces = 1;
  /* cull + step on the device; the BCs stay on the host in this binding because
     gfs_particle_bc also ships particles to other MPI ranks (:3218-3244).  A
     single-process periodic run can instead declare its periodic sides with
     gfsb200_tree_set_periodic and call gfsb200_particle_list_event, which wraps
     and drops on the device (gfsb200_particle_bc). */
  par.track_escapes = par.n_forces > 0;
  /* remove_particles_not_in_domain (:955-969) is a whole pass (locate + flag + select).  It is
     skipped when no particle can be outside: the device state was carried over from the last
     event (resident mode), in which the step kernel counted no particle leaving, and nothing on
     the host has touched the list since.  (It cannot simply run after the step instead: the
     particles that leave during this step must reach gfs_particle_bc, not the cull.) */
  if (!(carried && par.track_escapes))
    for (k = 0; k < s->n_dev; k++) {
      gint64 r = 0;
      if (gfsb200_particles_cull (s->dev[k], &r) != GFSB200_OK)
	g_error ("particulates (B200): %s", gfsb200_last_error ());
      removed += r;
    }
  for (k = 0; k < s->n_dev; k++)
    if (gfsb200_step (s->dev[k], &par) != GFSB200_OK)
      g_error ("particulates (B200): %s", gfsb200_last_error ());
  if (par.track_escapes) {
    gint64 outside = 0;
    for (k = 0; k < s->n_dev; k++) {
      gint64 e = 0, o = 0;
      if (gfsb200_step_counts (s->dev[k], &e, &o) != GFSB200_OK)
	g_error ("particulates (B200): %s", gfsb200_last_error ());
      escaped += e;
      outside += o;
    }
    /* (before any compaction below: the record holds positions in the list as it was stepped) */
    if (s->resident && escaped > 0)
      patch_pos_old (s, escaped);
    if (outside > 0) {
      /* A carried-over list holds particles that are outside the domain (the escape count of the
	 last event tests the box hull; a particle can also end up in a destroyed, i.e. solid,
	 cell): the step has left them untouched; cull them now, as remove_particles_not_in_domain
	 would have before the step, and let the objects follow */
      for (k = 0; k < s->n_dev; k++) {
	gint64 r = 0;
	if (gfsb200_particles_cull (s->dev[k], &r) != GFSB200_OK)
	  g_error ("particulates (B200): %s", gfsb200_last_error ());
	removed += r;
      }
    }
  }
  t[4] = wall ();
  if (s->resident && par.track_escapes && removed == 0 && escaped == 0 && s->list_known) {
    /* nothing the host has to act on: the objects are refreshed when somebody asks */
    s->host_stale = TRUE;
    t[5] = wall ();
  }
  else {
    s->host_stale = FALSE;
    download_particles (s, plist, n_up);
    t[5] = wall ();
  }

  /* :993, host side as in the reference.  gfs_particle_bc spends one gfs_domain_locate per
     particle to find those that left the domain; the step kernel has already counted them,
     and when there are none the reference function has nothing to do. */
  if (!par.track_escapes || escaped > 0)
    gfs_particle_bc (plist);
  t[6] = wall ();
  if (s->warm) {                /* the first event pays for the flatten and the page-locked buffers */
    for (k = 0; k < 6; k++)
      s->phase[k] += t[k + 1] - t[k];
    s->n_events++;
  }
  s->warm = TRUE;
  /* :1003-1012: the reference refreshes Un,Vn,Wn only when the list holds a
     GfsForceInertial; a GfsForceAddedMass alone keeps the snapshot its read method
     took (verified against the reference's object code, tests/test_reference_objcode.py) */
  if (s->uold && s->snapshot)
    store_previous_velocity (GFS_DOMAIN (sim), s->uold);
  return TRUE;
}

/* page-locked buffer for one component of the deposited field (kept: an event per step must not
   allocate) */
static gdouble * deposit_staging (B200State * s)
{
  if (s->n_cells > s->dep_cap) {
    gfsb200_host_free (s->dep_out);
    s->dep_out = pinned (sizeof (gdouble)*s->n_cells);
    s->dep_cap = s->n_cells;
  }
  return s->dep_out;
}

static gboolean b200_particulate_field_event (GfsEvent * event, GfsSimulation * sim)
{
  GfsVariable * v = GFS_VARIABLE (event);
  GfsParticulateField * pfield = GFS_PARTICULATE_FIELD (v);
  gfsb200_step_params par;
  ListVars lv;
  B200State * s;
  gdouble * out;
  gint k;

  const gchar * why = NULL;
  if (moving_solids (sim))
    why = "a simulation with moving solids";
  if (why || !step_params (pfield->plist, sim, &par, &lv, &why)) {
    reference_or_error (event, "GfsParticulateField", why);
    gfsb200_module_sync (pfield->plist);
    return (* reference_field_event) (event, sim);
  }
  if (!(* GFS_EVENT_CLASS (gfs_variable_class ())->event) (event, sim))
    return FALSE;
  s = state_of (pfield->plist);
  adopt (s, &lv);
  refresh_tree (s, sim);
  device_current (s, pfield->plist);
  for (k = 0; k < s->n_dev; k++)
    if (gfsb200_deposit_volume (s->dev[k]) != GFSB200_OK)
      g_error ("particulates (B200): %s", gfsb200_last_error ());
  if (s->n_dev > 1 && gfsb200_deposit_allreduce (s->comm, s->n_dev) != GFSB200_OK)
    g_error ("particulates (B200): %s", gfsb200_last_error ());
  out = deposit_staging (s);
  if (gfsb200_download_deposit (s->ctx, 0, out) != GFSB200_OK)
    g_error ("particulates (B200): %s", gfsb200_last_error ());
  /* gfs_cell_reset on the leaves + the scatter of :1945-1953 */
  gfsb200_ftt_scatter (s->map, offsetof (GfsStateVector, place_holder), v->i, TRUE, out);
  return TRUE;
}

/* GfsSourceParticulate (source_particulate_event, :2177-2228): the force of the
 * particles on the fluid, spread by the user's kernel.  The kernel GfsFunction
 * is a compiled expression; it is recognised by probing (gfsb200_kernel_fit)
 * and the event stays on the reference path when it is none of the closed
 * forms the device evaluates. */
static gdouble kernel_trampoline (gdouble x, gdouble y, gdouble z, gpointer data)
{
  FttVector q = { x, y, z };
  return gfs_function_spatial_value ((GfsFunction *) data, &q);
}

static gboolean b200_source_particulate_event (GfsEvent * event, GfsSimulation * sim)
{
  /* (the GFS_SOURCE_PARTICULATE macro of particulatecommon.h:234-236 names a
     class getter that does not exist; a plain cast is what it expands to) */
  GfsSourceParticulate * sp = (GfsSourceParticulate *) event;
  gfsb200_step_params par;
  ListVars lv;
  gfsb200_kernel kernel;
  B200State * s;
  gdouble * out, * force[3];
  gint64 n, k = 0;
  GSList * i;
  FttComponent c;
  gint d;

  const gchar * why = NULL;
  if (moving_solids (sim))
    why = "a simulation with moving solids";
  if (!why && step_params (sp->plist, sim, &par, &lv, &why) &&
      gfsb200_kernel_fit (kernel_trampoline, sp->kernel_function, FTT_DIMENSION, &kernel) != GFSB200_OK)
    why = "a smoothing kernel that is none of the closed forms the device evaluates (constant, Gaussian, "
      "compact polynomial: include/gfsb200.h)";
  if (why) {
    reference_or_error (event, "GfsSourceParticulate", why);
    gfsb200_module_sync (sp->plist);
    return (* reference_source_event) (event, sim);
  }
  /* the timing gate of the parent class (:2180-2181) */
  if (!(* GFS_EVENT_CLASS (GTS_OBJECT_CLASS (gfs_source_particulate_class ())->parent_class)->event)
      (event, sim))
    return FALSE;

  s = state_of (sp->plist);
  adopt (s, &lv);
  refresh_tree (s, sim);
  mirror_velocity (s, GFS_DOMAIN (sim));
  n = device_current (s, sp->plist);
  kernel.record_norm = 0;
  for (d = 0; d < s->n_dev; d++)
    if (gfsb200_deposit_force_smoothed (s->dev[d], &par, sp->rkernel, &kernel) != GFSB200_OK)
      g_error ("particulates (B200): %s", gfsb200_last_error ());
  if (s->n_dev > 1 && gfsb200_deposit_allreduce (s->comm, s->n_dev) != GFSB200_OK)
    g_error ("particulates (B200): %s", gfsb200_last_error ());
  /* <plist>_Fx,_Fy,_Fz: gfs_cell_reset on the leaves + the scatter of diffuse_force */
  out = deposit_staging (s);
  for (c = 0; c < FTT_DIMENSION; c++) {
    if (gfsb200_download_deposit (s->ctx, 1 + c, out) != GFSB200_OK)
      g_error ("particulates (B200): %s", gfsb200_last_error ());
    gfsb200_ftt_scatter (s->map, offsetof (GfsStateVector, place_holder), sp->u[c]->i, TRUE, out);
  }
  /* the reference leaves the on-fluid force in particulate->force (:2195-2201); in resident
     mode with stale objects the device keeps it and the next sync_down delivers it */
  if (!state_of (sp->plist)->host_stale) {
    gdouble * mass = g_malloc (sizeof (gdouble)*(n ? n : 1));
    for (c = 0; c < 3; c++)
      force[c] = g_malloc (sizeof (gdouble)*(n ? n : 1));
    /* the on-fluid pass of a GfsForceAddedMass updates the particle mass too (compute_forces_onfluid
       -> :391): it comes back with the forces, or the next upload would undo it */
    if (s->n_dev == 1) {
      if (gfsb200_particles_download (s->ctx, NULL, NULL, NULL, NULL, NULL, NULL, force[0], force[1], force[2],
				      mass, NULL, NULL, NULL) != GFSB200_OK)
	g_error ("particulates (B200): %s", gfsb200_last_error ());
      for (i = GFS_EVENT_LIST (sp->plist)->list->items; i && k < n; i = i->next, k++) {
	GfsParticulate * q = GFS_PARTICULATE (i->data);
	q->force.x = force[0][k]; q->force.y = force[1][k];
	q->force.z = FTT_DIMENSION > 2 ? force[2][k] : 0.;
	q->mass = mass[k];
      }
    }
    else {
      guint32 * ids = g_malloc (sizeof (guint32)*(n ? n : 1));
      gint64 off = 0, j;
      for (d = 0; d < s->n_dev; d++) {
	const gint64 m = gfsb200_particles_count (s->dev[d]);
	if (off + m > n)
	  g_error ("particulates (B200): the device lists and the GfsParticleList disagree");
	if (gfsb200_particles_download (s->dev[d], NULL, NULL, NULL, NULL, NULL, NULL, force[0] + off, force[1] + off,
					force[2] + off, mass + off, NULL, ids + off, NULL) != GFSB200_OK)
	  g_error ("particulates (B200): %s", gfsb200_last_error ());
	off += m;
      }
      for (j = 0; j < off; j++) {
	const gint32 o = ids[j] < s->idx_cap ? s->idx_of_id[ids[j]] : -1;
	if (o >= 0 && o < s->n_obj) {
	  GfsParticulate * q = s->obj[o];
	  q->force.x = force[0][j]; q->force.y = force[1][j];
	  q->force.z = FTT_DIMENSION > 2 ? force[2][j] : 0.;
	  q->mass = mass[j];
	}
      }
      g_free (ids);
    }
    for (c = 0; c < 3; c++)
      g_free (force[c]);
    g_free (mass);
  }
  return TRUE;
}

/* ------------------------------------------------------------------ */
/* module symbols, as modules/particulates.c:24-49                      */

const gchar gfs_module_name[] = "particulates";
const gchar * g_module_check_init (void);

const gchar * g_module_check_init (void)
{
  /* instantiate the reference classes in the reference's order ... */
  gfs_particulate_class ();
  gfs_particle_list_class ();
  gfs_force_inertial_class ();
  gfs_force_addedmass_class ();
  gfs_force_lift_class ();
  gfs_force_drag_class ();
  gfs_force_buoy_class ();
  gfs_particle_force_class ();
  gfs_source_particulate_class ();
  gfs_source_particulatevol_class ();
  gfs_source_particulatemass_class ();
  gfs_droplet_to_particle_class ();
  gfs_particle_to_droplet_class ();
  gfs_feed_particle_class ();
  gfs_output_particle_list_class ();
  gfs_particulate_field_class ();

  /* ... then route the hot-path events to the device */
  reference_list_event = GFS_EVENT_CLASS (gfs_particle_list_class ())->event;
  GFS_EVENT_CLASS (gfs_particle_list_class ())->event = b200_particle_list_event;
  /* ... and, for resident mode, bring the objects up to date before the reference's own
     readers / writers of a list look at them */
  {
    GfsEventClass * readers[N_READERS] = {
      GFS_EVENT_CLASS (gfs_droplet_to_particle_class ()), GFS_EVENT_CLASS (gfs_particle_to_droplet_class ()),
      GFS_EVENT_CLASS (gfs_feed_particle_class ()), GFS_EVENT_CLASS (gfs_output_particle_list_class ()),
      GFS_EVENT_CLASS (gfs_source_particulatevol_class ()), GFS_EVENT_CLASS (gfs_source_particulatemass_class ())
    };
    gint k;
    for (k = 0; k < N_READERS; k++) {
      reader_class[k] = readers[k];
      reader_event[k] = readers[k]->event;
      readers[k]->event = synced_reader_event;
    }
  }
  reference_list_write = GTS_OBJECT_CLASS (gfs_particle_list_class ())->write;
  GTS_OBJECT_CLASS (gfs_particle_list_class ())->write = b200_particle_list_write;
  reference_list_destroy = GTS_OBJECT_CLASS (gfs_particle_list_class ())->destroy;
  GTS_OBJECT_CLASS (gfs_particle_list_class ())->destroy = b200_particle_list_destroy;
  reference_field_event = GFS_EVENT_CLASS (gfs_particulate_field_class ())->event;
  GFS_EVENT_CLASS (gfs_particulate_field_class ())->event = b200_particulate_field_event;
  reference_source_event = GFS_EVENT_CLASS (gfs_source_particulate_class ())->event;
  GFS_EVENT_CLASS (gfs_source_particulate_class ())->event = b200_source_particulate_event;
  return NULL;
}

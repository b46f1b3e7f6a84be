/* TEST INFRASTRUCTURE ONLY (oracle) -- never linked into, imported by or
 * executed from the product path.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this.
 *
 * libgfsrefobj{2D,3D}.so = the REFERENCE'S OWN OBJECT CODE for the whole
 * particulate layer, not a restatement of it:
 *
 *   modules/particulatecommon.c   GfsParticulate, GfsParticleList, GfsForce*,
 *                                 GfsParticulateField, GfsSourceParticulate,
 *                                 gfs_particle_bc  (every function of SURVEY 8a)
 *   src/event.c                   GfsEvent gating, gfs_event_do, GfsEventList
 *   src/particle.c                GfsParticle (passive tracer event, %g writer)
 *   src/domain.c                  GfsLocateArray, gfs_domain_locate, traversals,
 *                                 gfs_domain_advect_point, gfs_domain_timer_*
 *   src/fluid.c, src/ftt.c        gfs_interpolate, gfs_center_gradient, trees
 *
 * are compiled UNMODIFIED from where they lie under /root/reference
 * (oracle/Makefile), against the reference's own headers and the
 * declaration-level GLib/GTS stand-in of gerris-fft-particles_b200/host/check.
 * This file is the run-time those objects need and the reference tree does not
 * contain: GLib and GTS are third-party dependencies absent from this image
 * (GTS 0.7.6, GLib 2.x: configure.ac:162-200), so the handful of their
 * functions the path calls are restated from their published behaviour
 * (object.c / container.c of GTS 0.7.6, gslist.c of GLib), and the few Gerris
 * functions that live in files which are not linked here (src/utils.c,
 * src/variable.c, src/source.c, src/boundary.c, src/simulation.c: each drags in
 * the solver) are restated with their file:line.  Everything the path does NOT call resolves to an aborting
 * stub (mkstubs.sh), so a call that strays off the restated set fails loudly.
 *
 * The FttCell trees and their cell data come from
 * libgfsoracle (oracle/particulate_port.c: trees built by the reference's
 * ftt.c object code); they are shared by pointer -- FttCell and GfsStateVector
 * have the same layout in both libraries because both use the reference's
 * ftt.h / fluid.h.
 */
#include <gts.h>
#include <sys/times.h>
#include <unistd.h>
#include "simulation.h"
#include "event.h"
#include "particle.h"
#include "source.h"
#include "boundary.h"
#include "particulatecommon.h"

#define REF_EXPORT __attribute__((visibility("default")))

/* ------------------------------------------------------------------ */
/* GLib (restated: gmem.c, gslist.c, garray.c, gmessages.c)            */

gpointer g_malloc (gsize n)
{
  gpointer p;
  if (n == 0) return NULL;
  p = malloc (n);
  if (!p) abort ();
  return p;
}

gpointer g_malloc0 (gsize n)
{
  gpointer p;
  if (n == 0) return NULL;
  p = calloc (1, n);
  if (!p) abort ();
  return p;
}

gpointer g_realloc (gpointer p, gsize n)
{
  if (n == 0) { free (p); return NULL; }
  p = realloc (p, n);
  if (!p) abort ();
  return p;
}

void g_free (gpointer p) { free (p); }

gchar * g_strdup (const gchar * s)
{
  gchar * d;
  if (!s) return NULL;
  d = g_malloc (strlen (s) + 1);
  strcpy (d, s);
  return d;
}

gchar * g_strconcat (const gchar * s, ...)
{
  va_list ap;
  gsize len;
  const gchar * t;
  gchar * d;
  if (!s) return NULL;
  len = strlen (s) + 1;
  va_start (ap, s);
  while ((t = va_arg (ap, const gchar *))) len += strlen (t);
  va_end (ap);
  d = g_malloc (len);
  strcpy (d, s);
  va_start (ap, s);
  while ((t = va_arg (ap, const gchar *))) strcat (d, t);
  va_end (ap);
  return d;
}

gchar * g_strdup_printf (const gchar * format, ...)
{
  va_list ap;
  gchar * d = NULL;
  va_start (ap, format);
  if (vasprintf (&d, format, ap) < 0) abort ();
  va_end (ap);
  return d;
}

const gchar * g_getenv (const gchar * variable) { return getenv (variable); }

static long ref_warnings = 0;

void g_log (const gchar * domain, GLogLevelFlags level, const gchar * format, ...)
{
  va_list ap;
  if (level & G_LOG_LEVEL_WARNING) {
    /* g_warning: log and continue (the added-mass/lift code warns per call) */
    if (ref_warnings++ > 8) return;
  }
  va_start (ap, format);
  fputs ("refobj: ", stderr);
  vfprintf (stderr, format, ap);
  fputc ('\n', stderr);
  va_end (ap);
  if (level & G_LOG_LEVEL_ERROR)
    abort ();
}

void g_assertion_message_expr (const char * domain, const char * file, int line, const char * func,
			       const char * expr)
{
  fprintf (stderr, "refobj: %s:%d (%s): assertion failed: (%s)\n", file, line, func, expr ? expr : "not reached");
  abort ();
}

GSList * g_slist_prepend (GSList * l, gpointer data)
{
  GSList * n = g_malloc (sizeof (GSList));
  n->data = data;
  n->next = l;
  return n;
}

GSList * g_slist_append (GSList * l, gpointer data)
{
  GSList * n = g_malloc (sizeof (GSList)), * i = l;
  n->data = data;
  n->next = NULL;
  if (!l) return n;
  while (i->next) i = i->next;
  i->next = n;
  return l;
}

GSList * g_slist_remove (GSList * l, gconstpointer data)
{
  GSList ** pp = &l;
  while (*pp) {
    if ((*pp)->data == data) {
      GSList * dead = *pp;
      *pp = dead->next;
      free (dead);
      break;
    }
    pp = &(*pp)->next;
  }
  return l;
}

GSList * g_slist_find (GSList * l, gconstpointer data)
{
  while (l && l->data != data) l = l->next;
  return l;
}

GSList * g_slist_reverse (GSList * l)
{
  GSList * prev = NULL;
  while (l) {
    GSList * next = l->next;
    l->next = prev;
    prev = l;
    l = next;
  }
  return prev;
}

void g_slist_free (GSList * l)
{
  while (l) {
    GSList * next = l->next;
    free (l);
    l = next;
  }
}

guint g_slist_length (GSList * l)
{
  guint n = 0;
  while (l) { n++; l = l->next; }
  return n;
}

typedef struct { GArray a; guint elt, alloc; } RealArray;

GArray * g_array_new (gboolean zero_terminated, gboolean clear, guint element_size)
{
  RealArray * a = g_malloc0 (sizeof (RealArray));
  a->elt = element_size;
  return &a->a;
}

GArray * g_array_append_vals (GArray * array, gconstpointer data, guint len)
{
  RealArray * a = (RealArray *) array;
  if (array->len + len > a->alloc) {
    a->alloc = MAX (2*a->alloc, array->len + len + 16);
    array->data = g_realloc (array->data, (gsize) a->alloc*a->elt);
  }
  memcpy (array->data + (gsize) array->len*a->elt, data, (gsize) len*a->elt);
  array->len += len;
  return array;
}

gchar * g_array_free (GArray * array, gboolean free_segment)
{
  gchar * d = array->data;
  if (free_segment) { free (d); d = NULL; }
  free (array);
  return d;
}

typedef struct { GPtrArray a; guint alloc; } RealPtrArray;

GPtrArray * g_ptr_array_new (void) { return g_malloc0 (sizeof (RealPtrArray)); }

void g_ptr_array_add (GPtrArray * array, gpointer p)
{
  RealPtrArray * a = (RealPtrArray *) array;
  if (array->len + 1 > a->alloc) {
    a->alloc = MAX (2*a->alloc, 16);
    array->pdata = g_realloc (array->pdata, sizeof (gpointer)*a->alloc);
  }
  array->pdata[array->len++] = p;
}

gpointer * g_ptr_array_free (GPtrArray * array, gboolean free_seg)
{
  gpointer * d = array->pdata;
  if (free_seg) { free (d); d = NULL; }
  free (array);
  return d;
}

/* GHashTable (ghash.c): chained buckets, direct hash/equality when none is given */
typedef struct _HNode HNode;
struct _HNode { gpointer key, value; HNode * next; };
struct _GHashTable { GHashFunc hash; GEqualFunc equal; guint nb; HNode ** bucket; };

guint g_str_hash (gconstpointer v)
{
  const signed char * p = v;
  guint h = 5381;
  for (; *p; p++) h = (h << 5) + h + *p;
  return h;
}

gboolean g_str_equal (gconstpointer a, gconstpointer b) { return strcmp (a, b) == 0; }

GHashTable * g_hash_table_new (GHashFunc hash, GEqualFunc equal)
{
  GHashTable * h = g_malloc0 (sizeof (GHashTable));
  h->hash = hash; h->equal = equal;
  h->nb = 257;
  h->bucket = g_malloc0 (sizeof (HNode *)*h->nb);
  return h;
}

static HNode ** hash_find (GHashTable * h, gconstpointer key)
{
  guint k = h->hash ? (* h->hash) (key) : (guint) ((gulong) key >> 4);
  HNode ** n = &h->bucket[k % h->nb];
  while (*n && !(h->equal ? (* h->equal) ((*n)->key, key) : (*n)->key == key))
    n = &(*n)->next;
  return n;
}

void g_hash_table_insert (GHashTable * h, gpointer key, gpointer value)
{
  HNode ** n = hash_find (h, key);
  if (!*n) {
    *n = g_malloc0 (sizeof (HNode));
    (*n)->key = key;
  }
  (*n)->value = value;
}

gpointer g_hash_table_lookup (GHashTable * h, gconstpointer key)
{
  HNode ** n = hash_find (h, key);
  return *n ? (*n)->value : NULL;
}

void g_hash_table_foreach (GHashTable * h, GHFunc func, gpointer data)
{
  guint b;
  for (b = 0; b < h->nb; b++) {
    HNode * n = h->bucket[b];
    while (n) {
      HNode * next = n->next;
      (* func) (n->key, n->value, data);
      n = next;
    }
  }
}

gboolean g_hash_table_remove (GHashTable * h, gconstpointer key)
{
  HNode ** n = hash_find (h, key);
  if (*n) {
    HNode * dead = *n;
    *n = dead->next;
    free (dead);
    return TRUE;
  }
  return FALSE;
}

/* ------------------------------------------------------------------ */
/* GTS 0.7.6 object system (restated: src/object.c)                     */

static void object_destroy (GtsObject * object)
{
  object->klass = NULL;
  g_free (object);
}

static void object_class_init (GtsObjectClass * klass)
{
  klass->clone = NULL;
  klass->destroy = object_destroy;
  klass->read = NULL;
  klass->write = NULL;
  klass->color = NULL;
  klass->attributes = NULL;
}

static void object_init (GtsObject * object)
{
  object->reserved = NULL;
  object->flags = 0;
}

/* ancestors' class_init first, the class's own last (object.c: gts_object_class_init) */
static void run_class_init (GtsObjectClass * klass, GtsObjectClass * parent_class)
{
  if (parent_class) {
    run_class_init (klass, parent_class->parent_class);
    if (parent_class->info.class_init_func)
      (* parent_class->info.class_init_func) (klass);
  }
}

gpointer gts_object_class_new (GtsObjectClass * parent_class, GtsObjectClassInfo * info)
{
  GtsObjectClass * klass;

  g_assert (info != NULL);
  g_assert (parent_class == NULL || info->object_size >= parent_class->info.object_size);
  g_assert (parent_class == NULL || info->class_size >= parent_class->info.class_size);
  klass = g_malloc0 (info->class_size);
  klass->info = *info;
  klass->parent_class = parent_class;
  run_class_init (klass, klass);
  return klass;
}

gpointer gts_object_class (void)
{
  static GtsObjectClass * klass = NULL;
  if (klass == NULL) {
    GtsObjectClassInfo info = {
      "GtsObject", sizeof (GtsObject), sizeof (GtsObjectClass),
      (GtsObjectClassInitFunc) object_class_init, (GtsObjectInitFunc) object_init, NULL, NULL
    };
    klass = gts_object_class_new (NULL, &info);
  }
  return klass;
}

static void run_object_init (GtsObject * object, GtsObjectClass * klass)
{
  if (klass) {
    run_object_init (object, klass->parent_class);
    if (klass->info.object_init_func)
      (* klass->info.object_init_func) (object);
  }
}

GtsObject * gts_object_new (GtsObjectClass * klass)
{
  GtsObject * object;
  g_assert (klass != NULL);
  object = g_malloc0 (klass->info.object_size);
  object->klass = klass;
  run_object_init (object, klass);
  return object;
}

void gts_object_destroy (GtsObject * object)
{
  g_assert (object != NULL && object->klass->destroy);
  (* object->klass->destroy) (object);
}

gpointer gts_object_class_is_from_class (gpointer klass, gpointer from)
{
  GtsObjectClass * c = klass;
  while (c) {
    if (c == from) return klass;
    c = c->parent_class;
  }
  return NULL;
}

gpointer gts_object_is_from_class (gpointer object, gpointer klass)
{
  if (object == NULL || klass == NULL) return NULL;
  return gts_object_class_is_from_class (((GtsObject *) object)->klass, klass) ? object : NULL;
}

/* GTS 0.7.6 containers (restated: src/container.c) */

GtsContaineeClass * gts_containee_class (void)
{
  static GtsContaineeClass * klass = NULL;
  if (klass == NULL) {
    GtsObjectClassInfo info = {
      "GtsContainee", sizeof (GtsContainee), sizeof (GtsContaineeClass), NULL, NULL, NULL, NULL
    };
    klass = gts_object_class_new (gts_object_class (), &info);
  }
  return klass;
}

static void slist_containee_destroy (GtsObject * object)
{
  GtsSListContainee * item = (GtsSListContainee *) object;
  GSList * i = item->containers;
  while (i) {
    GSList * next = i->next;
    gts_container_remove (i->data, GTS_CONTAINEE (item));
    i = next;
  }
  g_assert (item->containers == NULL);
  (* GTS_OBJECT_CLASS (gts_slist_containee_class ())->parent_class->destroy) (object);
}

static void slist_containee_add_container (GtsContainee * i, gpointer c)
{
  GtsSListContainee * item = (GtsSListContainee *) i;
  if (!g_slist_find (item->containers, c))
    item->containers = g_slist_prepend (item->containers, c);
}

static void slist_containee_remove_container (GtsContainee * i, gpointer c)
{
  GtsSListContainee * item = (GtsSListContainee *) i;
  item->containers = g_slist_remove (item->containers, c);
}

static gboolean slist_containee_is_contained (GtsContainee * i, gpointer c)
{
  return g_slist_find (((GtsSListContainee *) i)->containers, c) != NULL;
}

static void slist_containee_class_init (GtsContaineeClass * klass)
{
  klass->add_container = slist_containee_add_container;
  klass->remove_container = slist_containee_remove_container;
  klass->is_contained = slist_containee_is_contained;
  GTS_OBJECT_CLASS (klass)->destroy = slist_containee_destroy;
}

GtsSListContaineeClass * gts_slist_containee_class (void)
{
  static GtsSListContaineeClass * klass = NULL;
  if (klass == NULL) {
    GtsObjectClassInfo info = {
      "GtsSListContainee", sizeof (GtsSListContainee), sizeof (GtsSListContaineeClass),
      (GtsObjectClassInitFunc) slist_containee_class_init, NULL, NULL, NULL
    };
    klass = gts_object_class_new (GTS_OBJECT_CLASS (gts_containee_class ()), &info);
  }
  return klass;
}

static void container_add (GtsContainer * c, GtsContainee * item)
{
  if (GTS_CONTAINEE_CLASS (GTS_OBJECT (item)->klass)->add_container)
    (* GTS_CONTAINEE_CLASS (GTS_OBJECT (item)->klass)->add_container) (item, c);
}

static void container_remove (GtsContainer * c, GtsContainee * item)
{
  if (GTS_CONTAINEE_CLASS (GTS_OBJECT (item)->klass)->remove_container)
    (* GTS_CONTAINEE_CLASS (GTS_OBJECT (item)->klass)->remove_container) (item, c);
}

static void container_class_init (GtsContainerClass * klass)
{
  klass->add = container_add;
  klass->remove = container_remove;
  klass->foreach = NULL;
  klass->size = NULL;
}

GtsContainerClass * gts_container_class (void)
{
  static GtsContainerClass * klass = NULL;
  if (klass == NULL) {
    GtsObjectClassInfo info = {
      "GtsContainer", sizeof (GtsContainer), sizeof (GtsContainerClass),
      (GtsObjectClassInitFunc) container_class_init, NULL, NULL, NULL
    };
    klass = gts_object_class_new (GTS_OBJECT_CLASS (gts_slist_containee_class ()), &info);
  }
  return klass;
}

GtsContainer * gts_container_new (GtsContainerClass * klass)
{
  return GTS_CONTAINER (gts_object_new (GTS_OBJECT_CLASS (klass)));
}

void gts_container_add (GtsContainer * c, GtsContainee * item)
{
  g_assert (c != NULL && item != NULL);
  g_assert (GTS_CONTAINER_CLASS (GTS_OBJECT (c)->klass)->add);
  (* GTS_CONTAINER_CLASS (GTS_OBJECT (c)->klass)->add) (c, item);
}

void gts_container_remove (GtsContainer * c, GtsContainee * item)
{
  g_assert (c != NULL && item != NULL);
  g_assert (GTS_CONTAINER_CLASS (GTS_OBJECT (c)->klass)->remove);
  (* GTS_CONTAINER_CLASS (GTS_OBJECT (c)->klass)->remove) (c, item);
}

void gts_container_foreach (GtsContainer * c, GtsFunc func, gpointer data)
{
  g_assert (c != NULL && func != NULL);
  if (GTS_CONTAINER_CLASS (GTS_OBJECT (c)->klass)->foreach)
    (* GTS_CONTAINER_CLASS (GTS_OBJECT (c)->klass)->foreach) (c, func, data);
}

guint gts_container_size (GtsContainer * c)
{
  g_assert (c != NULL);
  return GTS_CONTAINER_CLASS (GTS_OBJECT (c)->klass)->size ?
    (* GTS_CONTAINER_CLASS (GTS_OBJECT (c)->klass)->size) (c) : 0;
}

/* GtsSListContainer: add PREPENDS (container.c: slist_container_add); foreach
   keeps the next link so that the callback may remove the current item */
static void slist_container_add (GtsContainer * c, GtsContainee * item)
{
  GtsSListContainer * s = GTS_SLIST_CONTAINER (c);
  g_return_if_fail (s->frozen == FALSE);
  if (!g_slist_find (s->items, item)) {
    s->items = g_slist_prepend (s->items, item);
    (* GTS_CONTAINER_CLASS (GTS_OBJECT_CLASS (gts_slist_container_class ())->parent_class)->add) (c, item);
  }
}

static void slist_container_remove (GtsContainer * c, GtsContainee * item)
{
  GtsSListContainer * s = GTS_SLIST_CONTAINER (c);
  g_return_if_fail (s->frozen == FALSE);
  if (g_slist_find (s->items, item)) {
    s->items = g_slist_remove (s->items, item);
    (* GTS_CONTAINER_CLASS (GTS_OBJECT_CLASS (gts_slist_container_class ())->parent_class)->remove) (c, item);
  }
}

static void slist_container_foreach (GtsContainer * c, GtsFunc func, gpointer data)
{
  GSList * i = GTS_SLIST_CONTAINER (c)->items;
  while (i) {
    GSList * next = i->next;
    (* func) (i->data, data);
    i = next;
  }
}

static guint slist_container_size (GtsContainer * c)
{
  return g_slist_length (GTS_SLIST_CONTAINER (c)->items);
}

static void slist_container_destroy (GtsObject * object)
{
  GtsSListContainer * s = (GtsSListContainer *) object;
  GSList * i = s->items;
  while (i) {
    GSList * next = i->next;
    container_remove (GTS_CONTAINER (s), i->data);
    i = next;
  }
  g_slist_free (s->items);
  s->items = NULL;
  (* GTS_OBJECT_CLASS (gts_slist_container_class ())->parent_class->destroy) (object);
}

static void slist_container_class_init (GtsContainerClass * klass)
{
  klass->add = slist_container_add;
  klass->remove = slist_container_remove;
  klass->foreach = slist_container_foreach;
  klass->size = slist_container_size;
  GTS_OBJECT_CLASS (klass)->destroy = slist_container_destroy;
}

GtsSListContainerClass * gts_slist_container_class (void)
{
  static GtsSListContainerClass * klass = NULL;
  if (klass == NULL) {
    GtsObjectClassInfo info = {
      "GtsSListContainer", sizeof (GtsSListContainer), sizeof (GtsSListContainerClass),
      (GtsObjectClassInitFunc) slist_container_class_init, NULL, NULL, NULL
    };
    klass = gts_object_class_new (GTS_OBJECT_CLASS (gts_container_class ()), &info);
  }
  return klass;
}

/* ------------------------------------------------------------------ */
/* Gerris classes whose source files do not compile here: only the class
 * objects (name, sizes, parent) -- the path uses them for GFS_IS_*() tests
 * and inherited GfsEvent methods                                        */

#define SIMPLE_CLASS(func, ctype, cname, otype, parent)			\
  ctype * func (void)							\
  {									\
    static ctype * klass = NULL;					\
    if (klass == NULL) {						\
      GtsObjectClassInfo info = { cname, sizeof (otype), sizeof (ctype), NULL, NULL, NULL, NULL }; \
      klass = gts_object_class_new (GTS_OBJECT_CLASS (parent), &info);	\
    }									\
    return klass;							\
  }

/* src/utils.c: GfsFunction is opaque to the path; here a constant, a cell
   variable, or a C function of (x,y,z) -- what a .gfs expression compiles to */
struct _GfsFunction {
  GtsObject parent;
  gdouble val;
  GfsVariable * v;
  gdouble (* spatial) (gdouble x, gdouble y, gdouble z, gpointer data);
  gpointer data;
};

SIMPLE_CLASS (gfs_function_class, GfsFunctionClass, "GfsFunction", struct _GfsFunction, gts_object_class ())
SIMPLE_CLASS (gfs_function_spatial_class, GfsFunctionClass, "GfsFunctionSpatial", struct _GfsFunction,
	      gfs_function_class ())

/* src/utils.c:1231-1259 for the constant and the variable case, L = 1 */
gdouble gfs_function_value (GfsFunction * f, FttCell * cell)
{
  g_return_val_if_fail (f != NULL, 0.);
  return f->v ? GFS_VALUE (cell, f->v) : f->val;
}

/* src/utils.c:1476-1494 */
gdouble gfs_function_spatial_value (GfsFunction * f, const FttVector * p)
{
  g_return_val_if_fail (f != NULL, 0.);
  g_return_val_if_fail (p != NULL, 0.);
  return f->spatial ? (* f->spatial) (p->x, p->y, p->z, f->data) : f->val;
}

/* src/utils.c:1394-1399 */
GfsVariable * gfs_function_get_variable (GfsFunction * f)
{
  g_return_val_if_fail (f != NULL, NULL);
  return f->v;
}

gdouble gfs_function_get_constant_value (GfsFunction * f)
{
  return f->v || f->spatial ? G_MAXDOUBLE : f->val;
}

GfsFunction * gfs_function_new (GfsFunctionClass * klass, gdouble val)
{
  GfsFunction * f = (GfsFunction *) gts_object_new (GTS_OBJECT_CLASS (klass));
  f->val = val;
  return f;
}

/* src/variable.c:104-138 (class), :111-118 (init) */
static void ref_variable_init (GfsVariable * v)
{
  GFS_EVENT (v)->istep = 1;
  v->centered = FALSE;
  v->component = FTT_DIMENSION;
}

GfsVariableClass * gfs_variable_class (void)
{
  static GfsVariableClass * klass = NULL;
  if (klass == NULL) {
    GtsObjectClassInfo info = {
      "GfsVariable", sizeof (GfsVariable), sizeof (GfsVariableClass),
      NULL, (GtsObjectInitFunc) ref_variable_init, NULL, NULL
    };
    klass = gts_object_class_new (GTS_OBJECT_CLASS (gfs_event_class ()), &info);
  }
  return klass;
}

/* gfs_variable_new, src/variable.c:140-181, for the harness: the cells of a RefSim carry a fixed
 * number of value slots (the oracle allocated them), so a variable added at run time -- the
 * drop-in module's unnamed mesh-watch variable, which never reads or writes its slot -- shares the
 * last one instead of growing every cell (gfs_domain_alloc).  The callbacks start out unset: the
 * harness' refinement copies the parent's values (refobj_sim_refine below). */
GfsVariable * gfs_variable_new (GfsVariableClass * klass, GfsDomain * domain, const gchar * name,
				const gchar * description)
{
  GfsVariable * v = GFS_VARIABLE (gts_object_new (GTS_OBJECT_CLASS (klass)));
  GSList * last = domain->variables;
  g_return_val_if_fail (last != NULL, NULL);
  while (last->next) last = last->next;
  gfs_object_simulation_set (v, domain);
  v->i = GFS_VARIABLE (last->data)->i;
  v->domain = domain;
  v->name = name ? g_strdup (name) : NULL;
  return v;
}

/* src/variable.c:183-191 */
GfsVariable * gfs_variable_from_name (GSList * i, const gchar * name)
{
  g_return_val_if_fail (name != NULL, NULL);
  while (i && (!GFS_VARIABLE (i->data)->name || strcmp (name, GFS_VARIABLE (i->data)->name)))
    i = i->next;
  return i ? GFS_VARIABLE (i->data) : NULL;
}

/* src/source.c: class skeletons */
SIMPLE_CLASS (gfs_source_generic_class, GfsSourceGenericClass, "GfsSourceGeneric", GfsSourceGeneric,
	      gfs_event_class ())
SIMPLE_CLASS (gfs_source_scalar_class, GfsSourceGenericClass, "GfsSourceScalar", GfsSourceScalar,
	      gfs_source_generic_class ())
SIMPLE_CLASS (gfs_source_velocity_class, GfsSourceGenericClass, "GfsSourceVelocity", GfsSourceVelocity,
	      gfs_source_generic_class ())
SIMPLE_CLASS (gfs_source_class, GfsSourceGenericClass, "GfsSource", GfsSource, gfs_source_scalar_class ())
SIMPLE_CLASS (gfs_source_diffusion_class, GfsSourceGenericClass, "GfsSourceDiffusion", GfsSourceDiffusion,
	      gfs_source_scalar_class ())
SIMPLE_CLASS (gfs_diffusion_class, GfsDiffusionClass, "GfsDiffusion", GfsDiffusion, gfs_event_class ())

/* src/source.c:941-946 (diffusion_cell) and :1002-1005 */
static gdouble ref_diffusion_cell (GfsDiffusion * d, FttCell * cell)
{
  gdouble val;
  if (d->mu) return GFS_VALUE (cell, d->mu);
  val = gfs_function_get_constant_value (d->val);
  return val < G_MAXDOUBLE ? val : 0.;
}

gdouble gfs_diffusion_cell (GfsDiffusion * d, FttCell * cell)
{
  return (* d->cell) (d, cell);
}

/* src/output.c, src/surface.c: parents of classes the module instantiates at load
   (GfsOutputParticleList, GfsDropletToParticle ...); never used by the path */
SIMPLE_CLASS (gfs_output_class, GfsOutputClass, "GfsOutput", GfsOutput, gfs_event_class ())
SIMPLE_CLASS (gfs_output_particle_class, GfsOutputClass, "GfsOutputParticle", GfsOutput, gfs_output_class ())
SIMPLE_CLASS (gfs_surface_class, GfsGenericSurfaceClass, "GfsSurface", GfsSurface, gts_object_class ())

/* src/boundary.c: class skeletons */
SIMPLE_CLASS (gfs_box_class, GfsBoxClass, "GfsBox", GfsBox, gts_slist_container_class ())
SIMPLE_CLASS (gfs_boundary_class, GfsBoundaryClass, "GfsBoundary", GfsBoundary, gts_object_class ())
SIMPLE_CLASS (gfs_boundary_periodic_class, GfsBoundaryClass, "GfsBoundaryPeriodic", GfsBoundaryPeriodic,
	      gfs_boundary_class ())

/* ------------------------------------------------------------------ */
/* the simulation: a GfsSimulation whose boxes wrap trees built elsewhere */

typedef FttCell * (* RefLocateFunc) (gpointer handle, gdouble x, gdouble y, gdouble z, gint max_depth);
typedef void (* RefBcFunc) (gpointer handle, gint ivar);

typedef struct {
  GfsSimulation sim;
  gint nbox;
  GfsBox ** box;
  RefLocateFunc locate;
  RefBcFunc bc;
  gpointer handle;
  gint nvar;
  GfsVariable ** var;
  GfsDiffusion * D;
} RefSim;

static void sim_foreach (GtsContainer * c, GtsFunc func, gpointer data)
{
  RefSim * s = (RefSim *) c;
  gint b;
  for (b = 0; b < s->nbox; b++)
    (* func) (s->box[b], data);
}

static void sim_class_init (GtsContainerClass * klass)
{
  klass->foreach = sim_foreach;
}

GfsSimulationClass * gfs_simulation_class (void)
{
  static GfsSimulationClass * klass = NULL;
  if (klass == NULL) {
    GtsObjectClassInfo info = {
      "GfsSimulation", sizeof (RefSim), sizeof (GfsSimulationClass),
      (GtsObjectClassInitFunc) sim_class_init, NULL, NULL, NULL
    };
    klass = gts_object_class_new (GTS_OBJECT_CLASS (gts_container_class ()), &info);
  }
  return klass;
}

/* gfs_domain_bc fills the ghost cells of one variable.  src/domain.c's own (weakened at link
   time, oracle/Makefile) walks the GfsBc objects of src/boundary.c, which the ghost trees
   built by libgfsoracle do not carry; this one calls back into libgfsoracle when it is given a
   filler (the tests set the ghost values themselves) */
void gfs_domain_bc (GfsDomain * domain, FttTraverseFlags flags, gint max_depth, GfsVariable * v)
{
  RefSim * s = (RefSim *) domain;
  if (s->bc)
    (* s->bc) (s->handle, v->i);
}

/* src/simulation.c:1893-1932 for a simulation without GfsMap objects (none of
   the configurations declares one): the L/lambda scaling only */
void gfs_simulation_map (GfsSimulation * sim, FttVector * p)
{
  FttComponent c;
  for (c = 0; c < 3; c++)
    (&p->x)[c] *= (&GFS_DOMAIN (sim)->lambda.x)[c]/sim->physical_params.L;
}

void gfs_simulation_map_inverse (GfsSimulation * sim, FttVector * p)
{
  FttComponent c;
  for (c = 0; c < 3; c++)
    (&p->x)[c] *= sim->physical_params.L/(&GFS_DOMAIN (sim)->lambda.x)[c];
}

/* src/utils.c:2282-2300: debug messages are off unless gfs_debug_enabled (TRUE) */
void gfs_debug (const gchar * format, ...)
{
}

/* src/utils.c:1870-1936: GfsClock, the user-time clock behind gfs_domain_timer_start/stop
   (src/domain.c:4137-4180, now the reference's own object code): one times() call per reading */
struct _GfsClock {
  clock_t start, stop;
  gboolean started;
};

GfsClock * gfs_clock_new (void)
{
  GfsClock * t = g_malloc (sizeof (GfsClock));
  t->start = -1;
  t->started = FALSE;
  return t;
}

void gfs_clock_start (GfsClock * t)
{
  struct tms tm;
  g_return_if_fail (t != NULL);
  g_return_if_fail (!t->started);
  if (times (&tm) == (clock_t) -1)
    g_warning ("cannot read clock");
  t->start = tm.tms_utime;
  t->started = TRUE;
}

gdouble gfs_clock_elapsed (GfsClock * t)
{
  g_return_val_if_fail (t != NULL, 0.);
  g_return_val_if_fail (t->start >= 0, 0.);
  if (t->started == FALSE)
    return (t->stop - t->start)/(gdouble) sysconf (_SC_CLK_TCK);
  else {
    struct tms tm;
    if (times (&tm) == (clock_t) -1)
      g_warning ("cannot read clock");
    return (tm.tms_utime - t->start)/(gdouble) sysconf (_SC_CLK_TCK);
  }
}

void gfs_clock_destroy (GfsClock * t)
{
  g_free (t);
}

/* fluid.c externals: mixed (solid) cells and statistics */
/* src/solid.c:1264-1273 */
void gfs_cell_cm (const FttCell * cell, FttVector * cm)
{
  if (GFS_IS_MIXED (cell))
    *cm = GFS_STATE (cell)->solid->cm;
  else
    ftt_cell_pos (cell, cm);
}
void gts_range_init (GtsRange * r) { memset (r, 0, sizeof (GtsRange)); r->min = G_MAXDOUBLE; r->max = - G_MAXDOUBLE; }
void gts_range_add_value (GtsRange * r, gdouble val)
{
  if (val < r->min) r->min = val;
  if (val > r->max) r->max = val;
  r->sum += val; r->sum2 += val*val; r->n++;
}
void gts_range_update (GtsRange * r)
{
  if (r->n > 0) { r->mean = r->sum/r->n; r->stddev = sqrt (fabs (r->sum2/r->n - r->mean*r->mean)); }
}

/* ------------------------------------------------------------------ */
/* exported driver                                                      */

enum { REF_FORCE_DRAG = 1, REF_FORCE_LIFT = 2, REF_FORCE_BUOY = 3,
       REF_FORCE_INERTIAL = 4, REF_FORCE_ADDEDMASS = 5 };

/* same layout as OraStepParams (oracle/particulate_port.c) */
typedef struct {
  double dt;
  int n_forces;
  int force[8];
  double rho;
  int ivar_alpha;
  double mu;
  int ivar_mu;
  double g[3];
  double cd_const, cl_const;
  int pattern;
  int ivar_uold;
  double cm_const;
} RefStepParams;

REF_EXPORT int refobj_dimension (void) { return FTT_DIMENSION; }

/* nbox GfsBox roots; broot[b*FTT_NEIGHBORS + d] = root of the GfsBoundary ghost
 * tree on side d of box b (or NULL); bit d of periodic_mask makes the
 * boundaries on side d GfsBoundaryPeriodic, matched to the box at the far end
 * of the chain (the box itself for single-box domains). */
REF_EXPORT RefSim * refobj_sim_new (int nbox, FttCell ** root, FttCell ** broot, unsigned periodic_mask,
				    RefLocateFunc locate, RefBcFunc bc, void * handle, int nvar)
{
  RefSim * s = (RefSim *) gts_object_new (GTS_OBJECT_CLASS (gfs_simulation_class ()));
  GfsDomain * domain = GFS_DOMAIN (s);
  gint b, i;
  FttDirection d;

  s->locate = locate;
  s->bc = bc;
  s->handle = handle;
  s->nbox = nbox;
  s->box = g_malloc0 (sizeof (GfsBox *)*nbox);
  for (b = 0; b < nbox; b++) {
    GfsBox * box = GFS_BOX (gts_object_new (GTS_OBJECT_CLASS (gfs_box_class ())));
    gfs_object_simulation_set (box, s);
    box->root = root[b];
    box->id = b + 1;
    FTT_ROOT_CELL (root[b])->parent = box;
    s->box[b] = box;
  }
  for (b = 0; b < nbox; b++)
    for (d = 0; d < FTT_NEIGHBORS; d++) {
      FttCell * r = broot ? broot[b*FTT_NEIGHBORS + d] : NULL;
      if (r) {
	gboolean periodic = (periodic_mask >> d) & 1;
	GfsBoundary * bo = GFS_BOUNDARY (gts_object_new (GTS_OBJECT_CLASS (periodic ?
									   gfs_boundary_periodic_class () :
									   gfs_boundary_class ())));
	bo->root = r;
	bo->box = s->box[b];
	bo->d = FTT_OPPOSITE_DIRECTION (d);
	if (periodic) {
	  /* the chain runs along x: the far end for x sides, the box itself otherwise */
	  GFS_BOUNDARY_PERIODIC (bo)->matching = d/2 == 0 ? s->box[d == FTT_RIGHT ? 0 : nbox - 1] : s->box[b];
	  GFS_BOUNDARY_PERIODIC (bo)->d = d;
	}
	s->box[b]->neighbor[d] = GTS_OBJECT (bo);
      }
      else if (d == FTT_RIGHT && b + 1 < nbox)
	s->box[b]->neighbor[d] = GTS_OBJECT (s->box[b + 1]);
      else if (d == FTT_LEFT && b > 0)
	s->box[b]->neighbor[d] = GTS_OBJECT (s->box[b - 1]);
    }

  s->nvar = nvar;
  s->var = g_malloc0 (sizeof (GfsVariable *)*nvar);
  for (i = 0; i < nvar; i++) {
    static const gchar * uvw[3] = { "U", "V", "W" };
    GfsVariable * v = GFS_VARIABLE (gts_object_new (GTS_OBJECT_CLASS (gfs_variable_class ())));
    gfs_object_simulation_set (v, s);
    v->i = i;
    v->domain = domain;
    v->name = i < FTT_DIMENSION ? g_strdup (uvw[i]) : g_strdup_printf ("A%d", i);
    if (i < FTT_DIMENSION)
      v->component = i;
    s->var[i] = v;
    domain->variables = g_slist_append (domain->variables, v);
  }
  s->sim.physical_params.L = 1.;
  s->sim.solids = GTS_SLIST_CONTAINER (gts_container_new (GTS_CONTAINER_CLASS (gts_slist_container_class ())));
  s->sim.events = GTS_SLIST_CONTAINER (gts_container_new (GTS_CONTAINER_CLASS (gts_slist_container_class ())));
  s->sim.maps = GTS_SLIST_CONTAINER (gts_container_new (GTS_CONTAINER_CLASS (gts_slist_container_class ())));
  domain->objects = g_hash_table_new (g_str_hash, g_str_equal);
  domain->lambda.x = domain->lambda.y = domain->lambda.z = 1.;
  s->sim.time.end = G_MAXDOUBLE;
  s->sim.time.iend = G_MAXINT;
  /* what gfs_domain_init / gfs_simulation_init set up (src/domain.c:630-690) and the
     GfsLocateArray gfs_domain_locate needs (built after the boxes are linked, as
     gfs_domain_read does, src/domain.c:820-823) */
  domain->rootlevel = FTT_ROOT_CELL (root[0])->level;
  domain->timers = g_hash_table_new (g_str_hash, g_str_equal);
  domain->timer = gfs_clock_new ();
  gfs_clock_start (domain->timer);
  domain->array = gfs_locate_array_new (domain);
  return s;
}

static void clear_sources (RefSim * s)
{
  gint c;
  for (c = 0; c < FTT_DIMENSION; c++)
    if (s->var[c]->sources) {
      gts_container_foreach (s->var[c]->sources, (GtsFunc) gts_object_destroy, NULL);
      gts_object_destroy (GTS_OBJECT (s->var[c]->sources));
      s->var[c]->sources = NULL;
    }
  if (s->D) {
    gts_object_destroy (GTS_OBJECT (s->D->val));
    gts_object_destroy (GTS_OBJECT (s->D));
    s->D = NULL;
  }
  if (s->sim.physical_params.alpha) {
    gts_object_destroy (GTS_OBJECT (s->sim.physical_params.alpha));
    s->sim.physical_params.alpha = NULL;
  }
}

static GtsContainer * sources_of (GfsVariable * v)
{
  if (!v->sources)
    v->sources = gts_container_new (GTS_CONTAINER_CLASS (gts_slist_container_class ()));
  return v->sources;
}

/* what the .gfs file of a run declares around the particle list:
 *   PhysicalParams { alpha = ... }, SourceViscosity mu, Source {} V g ...,
 *   Time { dtmax = dt } */
REF_EXPORT void refobj_sim_configure (RefSim * s, const RefStepParams * par, int timers_on)
{
  gint c;

  clear_sources (s);
  s->sim.advection_params.dt = par->dt;
  /* (timers_on is ignored: gfs_domain_timer_start/stop are the reference's own and always run) */
  if (par->ivar_alpha >= 0) {
    GfsFunction * f = gfs_function_new (gfs_function_class (), 0.);
    f->v = s->var[par->ivar_alpha];
    s->sim.physical_params.alpha = f;
  }
  else if (par->rho != 1.)
    s->sim.physical_params.alpha = gfs_function_new (gfs_function_class (), 1./par->rho);
  if (par->ivar_mu >= 0 || par->mu != 0.) {
    GfsSourceDiffusion * sd = (GfsSourceDiffusion *)
      gts_object_new (GTS_OBJECT_CLASS (gfs_source_diffusion_class ()));
    s->D = (GfsDiffusion *) gts_object_new (GTS_OBJECT_CLASS (gfs_diffusion_class ()));
    s->D->val = gfs_function_new (gfs_function_class (), par->mu);
    s->D->mu = par->ivar_mu >= 0 ? s->var[par->ivar_mu] : NULL;
    s->D->cell = ref_diffusion_cell;
    sd->D = s->D;
    gts_container_add (sources_of (s->var[0]), GTS_CONTAINEE (sd));
  }
  for (c = 0; c < FTT_DIMENSION; c++)
    if (par->g[c] != 0.) {
      GfsSource * src = (GfsSource *) gts_object_new (GTS_OBJECT_CLASS (gfs_source_class ()));
      src->intensity = gfs_function_new (gfs_function_class (), par->g[c]);
      gts_container_add (sources_of (s->var[c]), GTS_CONTAINEE (src));
    }
}

REF_EXPORT void refobj_sim_destroy (RefSim * s)
{
  gint b, i;
  FttDirection d;
  clear_sources (s);
  for (b = 0; b < s->nbox; b++)
    for (d = 0; d < FTT_NEIGHBORS; d++)
      if (s->box[b]->neighbor[d] && GFS_IS_BOUNDARY (s->box[b]->neighbor[d]))
	gts_object_destroy (s->box[b]->neighbor[d]);
  for (b = 0; b < s->nbox; b++) {
    FTT_ROOT_CELL (s->box[b]->root)->parent = NULL;
    gts_object_destroy (GTS_OBJECT (s->box[b]));
  }
  g_free (s->box);
  for (i = 0; i < s->nvar; i++) {
    g_free (s->var[i]->name);
    gts_object_destroy (GTS_OBJECT (s->var[i]));
  }
  g_free (s->var);
  g_slist_free (GFS_DOMAIN (s)->variables);
  gfs_locate_array_destroy (GFS_DOMAIN (s)->array);
  gfs_clock_destroy (GFS_DOMAIN (s)->timer);
  gts_container_foreach (GTS_CONTAINER (s->sim.solids), (GtsFunc) gts_object_destroy, NULL);
  gts_object_destroy (GTS_OBJECT (s->sim.solids));
  gts_object_destroy (GTS_OBJECT (s->sim.events));
  gts_object_destroy (GTS_OBJECT (s->sim.maps));
  gts_object_destroy (GTS_OBJECT (s));
}

/* declares a solid in the simulation: a GfsSolid (moving == 0), or a GfsSolidMoving
   (modules/moving.c), which the drop-in module hands back to the reference's events.
   Only the class name and the entry in sim->solids matter to the particulate path. */
SIMPLE_CLASS (ref_solid_class, GtsSListContaineeClass, "GfsSolid", GtsSListContainee, gts_slist_containee_class ())
SIMPLE_CLASS (ref_solid_moving_class, GtsSListContaineeClass, "GfsSolidMoving", GtsSListContainee, ref_solid_class ())

REF_EXPORT void refobj_sim_add_solid (RefSim * s, int moving)
{
  gts_container_add (GTS_CONTAINER (s->sim.solids),
		     GTS_CONTAINEE (gts_object_new (GTS_OBJECT_CLASS (moving ? ref_solid_moving_class () :
									ref_solid_class ()))));
}

/* gfs_domain_locate (src/domain.c:2623-2638 over the GfsLocateArray of :43-145), the
   reference's own object code, for n points */
/* --- mesh adaptation, as gfs_simulation_adapt leaves it (src/adaptive.c) ----------------------
 * refine: ftt_cell_refine_single (face 2:1 balance included) with the cell initialisation of
 * gfs_cell_fine_init (src/domain.c:2979-3010) for the harness' fixed-size payload -- the children
 * get the parent's values, then every variable's coarse_fine method runs;
 * coarsen: ftt_cell_coarsen with gfs_cell_cleanup (src/fluid.c:1965-1985): every variable's
 * cleanup method, then the payload is freed. */
static void ref_cell_fine_init (FttCell * parent, RefSim * s)
{
  FttCellChildren child;
  guint n;
  GSList * i;
  const size_t bytes = sizeof (GfsStateVector) + (s->nvar - 1)*sizeof (gdouble);
  ftt_cell_children (parent, &child);
  for (n = 0; n < FTT_CELLS; n++)
    if (child.c[n]) {
      child.c[n]->data = g_malloc0 (bytes);
      memcpy (child.c[n]->data, parent->data, bytes);
      GFS_STATE (child.c[n])->solid = NULL;
    }
  for (i = GFS_DOMAIN (s)->variables; i; i = i->next) {
    GfsVariable * v = i->data;
    if (v->coarse_fine)
      (* v->coarse_fine) (parent, v);
  }
}

static void ref_cell_cleanup (FttCell * cell, RefSim * s)
{
  GSList * i;
  if (cell->data)
    for (i = GFS_DOMAIN (s)->variables; i; i = i->next) {
      GfsVariable * v = i->data;
      if (v->cleanup)
	(* v->cleanup) (cell, v);
    }
  g_free (cell->data);
  cell->data = NULL;
}

static gboolean coarsen_always (FttCell * cell, gpointer data) { return TRUE; }

REF_EXPORT int refobj_sim_refine (RefSim * s, FttCell * cell)
{
  if (!cell || !FTT_CELL_IS_LEAF (cell))
    return 0;
  ftt_cell_refine_single (cell, (FttCellInitFunc) ref_cell_fine_init, s);
  return 1;
}

/* an entirely solid cell, as gfs_init_solid_fractions leaves it: destroyed (src/solid.c:833) */
REF_EXPORT int refobj_sim_destroy_cell (RefSim * s, FttCell * cell)
{
  if (!cell || FTT_CELL_IS_ROOT (cell))
    return 0;
  ftt_cell_destroy (cell, (FttCellCleanupFunc) ref_cell_cleanup, s);
  return 1;
}

REF_EXPORT int refobj_sim_coarsen (RefSim * s, FttCell * cell)
{
  if (!cell || FTT_CELL_IS_LEAF (cell))
    return 0;
  return ftt_cell_coarsen (cell, coarsen_always, NULL, (FttCellCleanupFunc) ref_cell_cleanup, s);
}

REF_EXPORT void refobj_locate (RefSim * s, long n, const double * x, const double * y, const double * z,
			       void ** cell)
{
  long i;
  for (i = 0; i < n; i++) {
    FttVector p = { x[i], y[i], z ? z[i] : 0., 0. };
    cell[i] = gfs_domain_locate (GFS_DOMAIN (s), p, -1, NULL);
  }
}

REF_EXPORT void refobj_locate_array (RefSim * s, double * min, double * h, int * n)
{
  GfsLocateArray * a = GFS_DOMAIN (s)->array;
  gint c;
  for (c = 0; c < FTT_DIMENSION; c++) {
    min[c] = a->min[c];
    n[c] = a->n[c];
  }
  *h = a->h;
}

REF_EXPORT void refobj_sim_time (RefSim * s, double * t, int * i)
{
  *t = s->sim.time.t;
  *i = s->sim.time.i;
}

/* --- GfsParticleList ------------------------------------------------- */

static GtsObjectClass * force_class (int kind)
{
  switch (kind) {
  case REF_FORCE_DRAG:      return GTS_OBJECT_CLASS (gfs_force_drag_class ());
  case REF_FORCE_LIFT:      return GTS_OBJECT_CLASS (gfs_force_lift_class ());
  case REF_FORCE_BUOY:      return GTS_OBJECT_CLASS (gfs_force_buoy_class ());
  case REF_FORCE_INERTIAL:  return GTS_OBJECT_CLASS (gfs_force_inertial_class ());
  case REF_FORCE_ADDEDMASS: return GTS_OBJECT_CLASS (gfs_force_addedmass_class ());
  }
  g_assert_not_reached ();
  return NULL;
}

/* Builds the object graph gfs_particle_list_read (:1022-1094) and
 * gfs_event_list_read (src/event.c:2446-2506) leave behind for
 *   GfsParticleList { istep = 1 } GfsParticulate { id x y z mass volume vx vy vz ... }
 *                   { GfsForce... }
 * : particles added one by one (the container prepends) then reversed, every
 * particle given the list's event parameters (copy_event), the force list
 * reversed and shared by pointer.  n_forces == 0 gives plain GfsParticle
 * tracers when `tracers' is set (GfsParticulate with forces == NULL behaves
 * the same, :811-813). */
REF_EXPORT GfsParticleList * refobj_list_new (RefSim * s, long n, const double * x, const double * y,
					      const double * z, const double * vx, const double * vy,
					      const double * vz, const double * mass, const double * volume,
					      const RefStepParams * par)
{
  GfsParticleList * plist = GFS_PARTICLE_LIST (gts_object_new (GTS_OBJECT_CLASS (gfs_particle_list_class ())));
  GfsEventList * l = GFS_EVENT_LIST (plist);
  GfsEvent * le = GFS_EVENT (plist);
  GSList * it;
  long i;
  gint k, c;

  gfs_object_simulation_set (plist, s);
  le->name = g_strdup ("P");
  g_hash_table_insert (GFS_DOMAIN (s)->objects, le->name, plist);     /* src/event.c:199-202 */
  gfs_event_set (le, -1., -1., -1., -1, -1, 1);
  l->klass = GTS_OBJECT_CLASS (gfs_particulate_class ());
  for (i = 0; i < n; i++) {
    GtsObject * o = gts_object_new (l->klass);
    GfsParticle * p = GFS_PARTICLE (o);
    GfsParticulate * pa = GFS_PARTICULATE (o);
    gfs_object_simulation_set (o, s);
    p->id = i + 1;
    p->pos.x = x[i]; p->pos.y = y[i]; p->pos.z = z ? z[i] : 0.;
    p->pos_old = p->pos;
    pa->vel.x = vx[i]; pa->vel.y = vy[i]; pa->vel.z = vz ? vz[i] : 0.;
    pa->mass = mass[i];
    pa->volume = volume[i];
    /* gts_container_add (l->list, o) without its O(n) "already there?" search (the
       object is new): prepend + the containee's add_container, as slist_container_add does */
    l->list->items = g_slist_prepend (l->list->items, o);
    container_add (GTS_CONTAINER (l->list), GTS_CONTAINEE (o));
  }
  l->list->items = g_slist_reverse (l->list->items);
  for (it = l->list->items; it; it = it->next)
    gfs_event_set (GFS_EVENT (it->data), le->start, le->end, le->step, le->istart, le->iend, le->istep);

  for (k = 0; k < par->n_forces; k++) {
    GtsObject * o = gts_object_new (force_class (par->force[k]));
    GfsForceCoeff * coeff = GFS_IS_FORCE_COEFF (o) ? FORCE_COEFF (o) : NULL;
    gdouble cst = G_MAXDOUBLE;
    gfs_object_simulation_set (o, s);
    if (coeff) {
      for (c = 0; c < FTT_DIMENSION; c++)
	coeff->Uold[c] = par->ivar_uold >= 0 ? s->var[par->ivar_uold + c] : NULL;
      if (par->force[k] == REF_FORCE_DRAG && par->cd_const == par->cd_const) cst = par->cd_const;
      if (par->force[k] == REF_FORCE_LIFT && par->cl_const == par->cl_const) cst = par->cl_const;
      if (par->force[k] == REF_FORCE_ADDEDMASS && par->cm_const == par->cm_const) cst = par->cm_const;
      if (cst < G_MAXDOUBLE) {
	/* a coefficient function: the reference stores Rep, Urelp ... in cell
	   variables before evaluating it (:566-575); they need a home */
	g_assert (s->nvar >= 5);
	coeff->re_p = s->var[s->nvar - 1];
	coeff->u_rel = s->var[s->nvar - 2];
	coeff->v_rel = s->var[s->nvar - 3];
	coeff->w_rel = s->var[s->nvar - 4];
	coeff->pdia = s->var[s->nvar - 5];
	coeff->coefficient = gfs_function_new (gfs_function_class (), cst);
      }
    }
    gts_container_add (GTS_CONTAINER (plist->forces), GTS_CONTAINEE (o));
  }
  if (plist->forces->items != NULL) {
    plist->forces->items = g_slist_reverse (plist->forces->items);
    gts_container_foreach (GTS_CONTAINER (l->list), (GtsFunc) assign_forces, plist->forces);
  }
  plist->idlast = n;
  plist->first_call = TRUE;
  return plist;
}

REF_EXPORT void refobj_list_destroy (GfsParticleList * plist)
{
  gts_object_destroy (GTS_OBJECT (plist));
}

REF_EXPORT long refobj_list_size (GfsParticleList * plist)
{
  return g_slist_length (GFS_EVENT_LIST (plist)->list->items);
}

REF_EXPORT void refobj_list_get (GfsParticleList * plist, double * x, double * y, double * z,
				 double * vx, double * vy, double * vz, double * fx, double * fy, double * fz,
				 double * mass, unsigned * id)
{
  GSList * i = GFS_EVENT_LIST (plist)->list->items;
  long k = 0;
  for (; i; i = i->next, k++) {
    GfsParticle * p = GFS_PARTICLE (i->data);
    GfsParticulate * pa = GFS_PARTICULATE (i->data);
    if (x) x[k] = p->pos.x;
    if (y) y[k] = p->pos.y;
    if (z) z[k] = p->pos.z;
    if (vx) vx[k] = pa->vel.x;
    if (vy) vy[k] = pa->vel.y;
    if (vz) vz[k] = pa->vel.z;
    if (fx) fx[k] = pa->force.x;
    if (fy) fy[k] = pa->force.y;
    if (fz) fz[k] = pa->force.z;
    if (mass) mass[k] = pa->mass;
    if (id) id[k] = p->id;
  }
}

/* one pass of simulation_run's event loop over this list (src/simulation.c:
 * 483: gts_container_foreach (sim->events, gfs_event_do, sim)), then the
 * time level advances as at the bottom of the loop (:541-543) */
REF_EXPORT int refobj_list_event (RefSim * s, GfsParticleList * plist, int steps)
{
  int k, fired = 0;
  for (k = 0; k < steps; k++) {
    gfs_event_do (GFS_EVENT (plist), GFS_SIMULATION (s));
    fired += GFS_EVENT (plist)->realised;
    s->sim.time.t += s->sim.advection_params.dt;
    s->sim.time.i++;
  }
  return fired;
}

/* the list's cull alone (remove_particles_not_in_domain is static: run through
   the public event with a time step of 0 and no forces would move tracers, so
   instead expose the same test the cull applies) */
REF_EXPORT long refobj_list_outside (RefSim * s, GfsParticleList * plist)
{
  GSList * i = GFS_EVENT_LIST (plist)->list->items;
  long n = 0;
  for (; i; i = i->next)
    if (gfs_domain_locate (GFS_DOMAIN (s), GFS_PARTICLE (i->data)->pos, -1, NULL) == NULL)
      n++;
  return n;
}

/* a single force model on a single particle, through the vtable the class
   init installed (GfsParticleForce.force): f per unit volume, :423-655 */
REF_EXPORT void refobj_force (GfsParticleList * plist, long index, int k, double f[3])
{
  GSList * i = GFS_EVENT_LIST (plist)->list->items, * j = plist->forces->items;
  FttVector v;
  while (index-- > 0) i = i->next;
  while (k-- > 0) j = j->next;
  v = (* GFS_PARTICLE_FORCE (j->data)->force) (GFS_PARTICLE (i->data), GFS_PARTICLE_FORCE (j->data));
  f[0] = v.x; f[1] = v.y; f[2] = v.z;
}

/* gfs_particle_bc (:3375-3395) alone */
REF_EXPORT void refobj_list_bc (GfsParticleList * plist)
{
  gfs_particle_bc (plist);
}

/* the reference's writer of the particle block (gfs_particulate_write :910-926
   over gfs_particle_write src/particle.c:86-98), one line per particle */
REF_EXPORT int refobj_list_write (GfsParticleList * plist, const char * path)
{
  FILE * fp = fopen (path, "w");
  GSList * i = GFS_EVENT_LIST (plist)->list->items;
  if (!fp) return -1;
  for (; i; i = i->next) {
    fputs ("    ", fp);                 /* as gfs_event_list_write does, src/event.c:2516-2520 */
    (* GTS_OBJECT (i->data)->klass->write) (i->data, fp);
    fputc ('\n', fp);
  }
  fclose (fp);
  return 0;
}

/* --- GfsParticulateField --------------------------------------------- */

/* GfsParticulateField <var ivar> P : one event (reset + scatter, :1934-1957) */
REF_EXPORT void refobj_field_event (RefSim * s, GfsParticleList * plist, int ivar)
{
  GfsVariable * v = GFS_VARIABLE (gts_object_new (GTS_OBJECT_CLASS (gfs_particulate_field_class ())));
  gfs_object_simulation_set (v, s);
  v->i = ivar;
  v->domain = GFS_DOMAIN (s);
  GFS_PARTICULATE_FIELD (v)->plist = plist;
  gfs_event_set (GFS_EVENT (v), -1., -1., -1., -1, -1, 1);
  GFS_EVENT (v)->i = s->sim.time.i;
  gfs_event_do (GFS_EVENT (v), GFS_SIMULATION (s));
  g_assert (GFS_EVENT (v)->realised);
  gts_object_destroy (GTS_OBJECT (v));
}

/* --- GfsSourceParticulate -------------------------------------------- */

/* the closed forms of OraKernel, as the C a user's `kernel = ...` compiles to */
typedef struct { int kind; double a, b; int p; int flags; } RefKernel;

static gdouble kernel_spatial (gdouble x, gdouble y, gdouble z, gpointer data)
{
  const RefKernel * k = data;
  gdouble r2 = x*x + y*y + z*z;
  switch (k->kind) {
  case 0: return k->a;
  case 1: return k->a*exp (- k->b*r2);
  case 2: {
    gdouble t = 1. - k->b*r2, v = k->a;
    gint i;
    if (t <= 0.) return 0.;
    for (i = 0; i < k->p; i++) v *= t;
    return v;
  }
  case 3: return k->a/(1. + k->b*r2);     /* none of the closed forms of gfsb200.h: stays on the host */
  }
  g_assert_not_reached ();
  return 0.;
}

/* GfsSourceParticulate {} U V W P { rkernel = r kernel = f } : one event
 * (:2177-2228) depositing into variables ivar0..ivar0+dim-1 */
REF_EXPORT void refobj_source_event (RefSim * s, GfsParticleList * plist, int ivar0, double rkernel,
				     const RefKernel * kernel)
{
  GfsSourceParticulate * sp = GFS_SOURCE_PARTICULATE (gts_object_new (GTS_OBJECT_CLASS (gfs_source_particulate_class ())));
  RefKernel k = *kernel;
  gint c;

  gfs_object_simulation_set (sp, s);
  sp->plist = plist;
  sp->rkernel = rkernel;
  if (!sp->kernel_function)
    sp->kernel_function = gfs_function_new (gfs_function_spatial_class (), 0.);
  gfs_object_simulation_set (sp->kernel_function, s);
  sp->kernel_function->spatial = kernel_spatial;
  sp->kernel_function->data = &k;
  for (c = 0; c < FTT_DIMENSION; c++)
    sp->u[c] = s->var[ivar0 + c];
  gfs_event_set (GFS_EVENT (sp), -1., -1., -1., -1, -1, 1);
  GFS_EVENT (sp)->i = s->sim.time.i;
  gfs_event_do (GFS_EVENT (sp), GFS_SIMULATION (s));
  g_assert (GFS_EVENT (sp)->realised);
  gts_object_destroy (GTS_OBJECT (sp));
}

REF_EXPORT long refobj_warnings (void) { return ref_warnings; }

/* libgfsrefmod only: the drop-in module (gerris-fft-particles_b200/host/
 * particulates_b200.c) is linked in; run the entry point GLib calls when
 * Gerris g_module_open()s it (src/simulation.c:199-225).  Returns 1 when the
 * module is present and its init reported success, 0 when this is the plain
 * reference library. */
const gchar * g_module_check_init (void) __attribute__((weak));
extern const gchar gfs_module_name[] __attribute__((weak));

REF_EXPORT int refobj_module_init (void)
{
  static int done = 0;
  if (!g_module_check_init)
    return 0;
  if (!done) {
    const gchar * err = g_module_check_init ();
    if (err) {
      fprintf (stderr, "refobj: module %s failed to initialise: %s\n", gfs_module_name, err);
      return -1;
    }
    done = 1;
  }
  return 1;
}

/* the GfsParticleList's own write method (gfs_particle_list_write :1096-1113 over
   gfs_event_list_write src/event.c:2508-2523): the whole list as it appears in a dump */
REF_EXPORT int refobj_list_class_write (GfsParticleList * plist, const char * path)
{
  FILE * fp = fopen (path, "w");
  if (!fp) return -1;
  (* GTS_OBJECT (plist)->klass->write) (GTS_OBJECT (plist), fp);
  fputc ('\n', fp);
  fclose (fp);
  return 0;
}

/* libgfsrefmod: what host code outside the module calls before it reads the objects of a
   list driven in resident mode */
void gfsb200_module_sync (GfsParticleList * plist) __attribute__((weak));

REF_EXPORT int refobj_list_sync (GfsParticleList * plist)
{
  if (!gfsb200_module_sync)
    return 0;
  gfsb200_module_sync (plist);
  return 1;
}

REF_EXPORT const char * refobj_module_name (void)
{
  return g_module_check_init ? gfs_module_name : NULL;
}

/* called by the generated stubs */
void refobj_unimplemented (const char * name)
{
  fprintf (stderr, "refobj: `%s' is not part of the particulate path and has no run-time here\n", name);
  abort ();
}

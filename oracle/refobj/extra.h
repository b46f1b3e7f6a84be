/* TEST INFRASTRUCTURE ONLY (oracle).
 *
 * Force-included in front of the reference sources that oracle/Makefile
 * compiles into oracle/_ref/libgfsrefobj{2D,3D}.so: the few GLib/GTS names
 * modules/particulatecommon.c and src/domain.c use that the declaration-level stand-in
 * (gerris-fft-particles_b200/host/check/gts.h) does not carry.  None of them
 * is on the particulate path; those that are functions resolve to aborting
 * stubs generated at link time (oracle/refobj/mkstubs.sh). */
#ifndef REFOBJ_EXTRA_H
#define REFOBJ_EXTRA_H

#include <gts.h>

#ifndef G_LIKELY
# define G_LIKELY(x) (x)
#endif

guint g_str_hash (gconstpointer v);
gboolean g_str_equal (gconstpointer a, gconstpointer b);
void g_hash_table_foreach (GHashTable * h, GHFunc func, gpointer data);
#define g_ptr_array_index(array, i) ((array)->pdata)[i]
void g_ptr_array_set_size (GPtrArray * a, gint n);
typedef struct _GtsFifo GtsFifo;
GtsFifo * gts_fifo_new (void);
void gts_fifo_push (GtsFifo * f, gpointer data);
gpointer gts_fifo_pop (GtsFifo * f);
void gts_fifo_destroy (GtsFifo * f);
gboolean gts_fifo_is_empty (GtsFifo * f);
GSList * g_slist_reverse (GSList * l);
GSList * g_slist_find (GSList * l, gconstpointer data);
gpointer gts_object_class_is_from_class (gpointer klass, gpointer from);
void gts_point_transform (GtsPoint * p, GtsMatrix * m);
void gts_triangle_revert (GtsTriangle * t);
GtsMatrix * gts_matrix_inverse (GtsMatrix * m);
GtsMatrix * gts_matrix_product (GtsMatrix * m1, GtsMatrix * m2);
GtsMatrix * gts_matrix_scale (GtsMatrix * m, GtsVector s);
GtsMatrix * gts_matrix_translate (GtsMatrix * m, GtsVector t);
void gts_matrix_destroy (GtsMatrix * m);
void gts_surface_foreach_vertex (GtsSurface * s, GtsFunc func, gpointer data);
void gts_surface_foreach_face (GtsSurface * s, GtsFunc func, gpointer data);

#endif

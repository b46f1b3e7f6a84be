/* COMPILE-CHECK ONLY -- declaration-level stand-in for <gts.h> (GTS 0.7.x) and
 * the parts of <glib.h> that the Gerris headers and particulates_b200.c name.
 *
 * GLib and GTS are not installed in the image this repository is developed
 * in, so the drop-in module (../particulates_b200.c) cannot be linked there.
 * `make -C gerris-fft-particles_b200 check-host` still type-checks it
 * (gcc -fsyntax-only) against the REFERENCE'S OWN headers (src headers,
 * modules/particulatecommon.h, used in place) with these declarations
 * standing in for the two libraries.  Nothing here is linked or shipped; on a
 * machine with Gerris installed the real headers are used (see the build line
 * at the top of ../particulates_b200.c).
 */
#ifndef GFSB200_CHECK_GTS_H
#define GFSB200_CHECK_GTS_H

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdarg.h>
#include <float.h>
#include <limits.h>
#include <math.h>
#include <stdint.h>

/* ---- glib ------------------------------------------------------------ */
typedef double gdouble;
typedef float gfloat;
typedef unsigned int guint;
typedef int gint;
typedef int gboolean;
typedef void * gpointer;
typedef const void * gconstpointer;
typedef char gchar;
typedef unsigned char guchar;
typedef unsigned long gulong;
typedef long glong;
typedef short gshort;
typedef unsigned short gushort;
typedef unsigned short guint16;
typedef int gint32;
typedef unsigned int guint32;
typedef int64_t gint64;
typedef uint64_t guint64;
typedef size_t gsize;
typedef unsigned char guint8;

#ifndef TRUE
# define TRUE 1
#endif
#ifndef FALSE
# define FALSE 0
#endif
#define G_MAXDOUBLE DBL_MAX
#define G_MAXINT INT_MAX
#define G_MININT INT_MIN
#define G_MAXUINT UINT_MAX
#define G_STMT_START do
#define G_STMT_END while (0)
#define G_GNUC_UNUSED __attribute__((unused))
#define G_GNUC_PRINTF(a, b) __attribute__((format (printf, a, b)))
#define G_BEGIN_DECLS
#define G_END_DECLS
#define G_LOG_DOMAIN ((gchar *) 0)
#ifndef MAX
# define MAX(a, b) (((a) > (b)) ? (a) : (b))
#endif
#ifndef MIN
# define MIN(a, b) (((a) < (b)) ? (a) : (b))
#endif
#ifndef ABS
# define ABS(a) (((a) < 0) ? -(a) : (a))
#endif
#ifndef CLAMP
# define CLAMP(x, lo, hi) (((x) > (hi)) ? (hi) : (((x) < (lo)) ? (lo) : (x)))
#endif
#define GPOINTER_TO_INT(p) ((gint) (glong) (p))
#define GPOINTER_TO_UINT(p) ((guint) (gulong) (p))
#define GINT_TO_POINTER(i) ((gpointer) (glong) (i))
#define GUINT_TO_POINTER(u) ((gpointer) (gulong) (u))

typedef enum { G_LOG_LEVEL_ERROR = 1 << 2, G_LOG_LEVEL_CRITICAL = 1 << 3, G_LOG_LEVEL_WARNING = 1 << 4,
	       G_LOG_LEVEL_MESSAGE = 1 << 5, G_LOG_LEVEL_INFO = 1 << 6, G_LOG_LEVEL_DEBUG = 1 << 7 } GLogLevelFlags;
void g_log (const gchar * domain, GLogLevelFlags level, const gchar * format, ...) G_GNUC_PRINTF (3, 4);
#define g_error(...) do { g_log (G_LOG_DOMAIN, G_LOG_LEVEL_ERROR, __VA_ARGS__); for (;;) ; } while (0)
#define g_warning(...) g_log (G_LOG_DOMAIN, G_LOG_LEVEL_WARNING, __VA_ARGS__)
#define g_message(...) g_log (G_LOG_DOMAIN, G_LOG_LEVEL_MESSAGE, __VA_ARGS__)
void g_assertion_message_expr (const char * domain, const char * file, int line, const char * func, const char * expr);
#define g_assert(expr) do { if (!(expr)) g_assertion_message_expr (G_LOG_DOMAIN, __FILE__, __LINE__, __func__, #expr); } while (0)
#define g_assert_not_reached() g_assertion_message_expr (G_LOG_DOMAIN, __FILE__, __LINE__, __func__, NULL)
#define g_return_if_fail(expr) do { if (!(expr)) return; } while (0)
#define g_return_val_if_fail(expr, val) do { if (!(expr)) return (val); } while (0)

gpointer g_malloc (gsize n);
gpointer g_malloc0 (gsize n);
gpointer g_realloc (gpointer p, gsize n);
void g_free (gpointer p);
gchar * g_strdup (const gchar * s);
gchar * g_strconcat (const gchar * s, ...);
gchar * g_strdup_printf (const gchar * format, ...) G_GNUC_PRINTF (1, 2);
const gchar * g_getenv (const gchar * variable);

typedef struct _GSList GSList;
struct _GSList { gpointer data; GSList * next; };
GSList * g_slist_prepend (GSList * l, gpointer data);
GSList * g_slist_append (GSList * l, gpointer data);
GSList * g_slist_remove (GSList * l, gconstpointer data);
void g_slist_free (GSList * l);
GSList * g_slist_find (GSList * l, gconstpointer data);
guint g_slist_length (GSList * l);
typedef struct _GList GList;
struct _GList { gpointer data; GList * next, * prev; };
typedef struct _GHashTable GHashTable;
typedef guint (* GHashFunc) (gconstpointer key);
typedef gboolean (* GEqualFunc) (gconstpointer a, gconstpointer b);
typedef void (* GFunc) (gpointer data, gpointer user_data);
typedef void (* GHFunc) (gpointer key, gpointer value, gpointer user_data);
GHashTable * g_hash_table_new (GHashFunc hash, GEqualFunc equal);
void g_hash_table_insert (GHashTable * h, gpointer key, gpointer value);
gpointer g_hash_table_lookup (GHashTable * h, gconstpointer key);
gboolean g_hash_table_remove (GHashTable * h, gconstpointer key);
void g_hash_table_foreach (GHashTable * h, GHFunc func, gpointer data);
typedef struct _GNode GNode;
struct _GNode { gpointer data; GNode * next, * prev, * parent, * children; };
typedef struct _GTimer GTimer;
typedef struct { gchar * str; gsize len; gsize allocated_len; } GString;
typedef struct { gpointer * pdata; guint len; } GPtrArray;
GPtrArray * g_ptr_array_new (void);
void g_ptr_array_add (GPtrArray * a, gpointer p);
gpointer * g_ptr_array_free (GPtrArray * a, gboolean free_seg);
typedef struct { gchar * data; guint len; } GArray;
GArray * g_array_new (gboolean zero_terminated, gboolean clear, guint element_size);
GArray * g_array_append_vals (GArray * a, gconstpointer data, guint len);
gchar * g_array_free (GArray * a, gboolean free_segment);
#define g_array_append_val(a, v) g_array_append_vals (a, &(v), 1)
#define g_array_index(a, t, i) (((t *) (void *) (a)->data)[i])

/* ---- gts -------------------------------------------------------------- */
#define GTS_C_VAR extern
typedef struct _GtsObjectClassInfo GtsObjectClassInfo;
typedef struct _GtsObject GtsObject;
typedef struct _GtsObjectClass GtsObjectClass;
typedef struct _GtsFile GtsFile;
typedef void (* GtsObjectClassInitFunc) (GtsObjectClass * objclass);
typedef void (* GtsObjectInitFunc) (GtsObject * obj);
typedef void (* GtsArgSetFunc) (GtsObject * obj);
typedef void (* GtsArgGetFunc) (GtsObject * obj);
typedef gint (* GtsFunc) (gpointer item, gpointer data);
typedef gpointer (* GtsKeyFunc) (gpointer item, gpointer data);
typedef gdouble GtsVector[3];
typedef gdouble GtsVector4[4];
typedef GtsVector4 GtsMatrix;

struct _GtsObjectClassInfo {
  gchar name[40];
  guint object_size, class_size;
  GtsObjectClassInitFunc class_init_func;
  GtsObjectInitFunc object_init_func;
  GtsArgSetFunc arg_set_func;
  GtsArgGetFunc arg_get_func;
};
struct _GtsObject { GtsObjectClass * klass; gpointer reserved; guint32 flags; };
struct _GtsObjectClass {
  GtsObjectClassInfo info;
  GtsObjectClass * parent_class;
  void (* clone) (GtsObject *, GtsObject *);
  void (* destroy) (GtsObject *);
  void (* read) (GtsObject **, GtsFile *);
  void (* write) (GtsObject *, FILE *);
  void (* attributes) (GtsObject *, GtsObject *);
  GtsObject * (* color) (GtsObject *);
};
#define GTS_OBJECT_CAST(obj, type, klass) ((type *) (obj))
#define GTS_OBJECT_CLASS_CAST(objklass, type, klass) ((type *) (objklass))
#define GTS_OBJECT(obj) GTS_OBJECT_CAST (obj, GtsObject, gts_object_class ())
#define GTS_OBJECT_CLASS(klass) GTS_OBJECT_CLASS_CAST (klass, GtsObjectClass, gts_object_class ())
#define GTS_IS_OBJECT(obj) (gts_object_is_from_class (obj, gts_object_class ()))
gpointer gts_object_class (void);
gpointer gts_object_class_new (GtsObjectClass * parent_class, GtsObjectClassInfo * info);
gpointer gts_object_is_from_class (gpointer object, gpointer klass);
GtsObject * gts_object_new (GtsObjectClass * klass);
void gts_object_destroy (GtsObject * object);
GtsObjectClass * gts_object_class_from_name (const gchar * name);

typedef enum {
  GTS_NONE = 1 << 8, GTS_INT = 1 << 9, GTS_UINT = 1 << 10, GTS_FLOAT = 1 << 11,
  GTS_DOUBLE = 1 << 12, GTS_STRING = 1 << 13, GTS_FILE = 1 << 14, GTS_ERROR = 1 << 15, GTS_OBJ = 1 << 16
} GtsTokenType;
struct _GtsFile {
  FILE * fp; gchar * s, * s1;
  guint line, pos;
  GString * token;
  GtsTokenType type;
  gchar * error;
  guint curline, curpos, scope, scope_max;
  gint next_token;
  gchar * delimiters, * comments, * tokens;
};
typedef struct { GtsTokenType type; gchar name[30]; gboolean unique; gpointer data; gboolean set; guint line, pos; } GtsFileVariable;
void gts_file_error (GtsFile * f, const gchar * format, ...) G_GNUC_PRINTF (2, 3);
void gts_file_next_token (GtsFile * f);
void gts_file_first_token_after (GtsFile * f, GtsTokenType type);
guint gts_file_read (GtsFile * f, gpointer ptr, guint size, guint nmemb);
void gts_file_assign_variables (GtsFile * f, GtsFileVariable * vars);

typedef struct { GtsObject object; } GtsContainee;
typedef struct { GtsObjectClass parent_class;
  void (* remove_container) (GtsContainee *, gpointer);
  void (* add_container) (GtsContainee *, gpointer);
  void (* foreach) (GtsContainee *, GtsFunc, gpointer);
  gboolean (* is_contained) (GtsContainee *, gpointer);
  void (* replace) (GtsContainee *, GtsContainee *); } GtsContaineeClass;
typedef struct { GtsContainee containee; GSList * containers; } GtsSListContainee;
typedef struct { GtsContaineeClass parent_class; } GtsSListContaineeClass;
typedef struct { GtsSListContainee object; } GtsContainer;
typedef struct { GtsSListContaineeClass parent_class;
  void (* add) (GtsContainer *, GtsContainee *);
  void (* remove) (GtsContainer *, GtsContainee *);
  void (* foreach) (GtsContainer *, GtsFunc, gpointer);
  guint (* size) (GtsContainer *); } GtsContainerClass;
typedef struct { GtsContainer c; GHashTable * items; gboolean frozen; } GtsHashContainer;
typedef struct { GtsContainerClass parent_class; } GtsHashContainerClass;
typedef struct { GtsContainer c; GSList * items; gboolean frozen; } GtsSListContainer;
typedef struct { GtsContainerClass parent_class; } GtsSListContainerClass;
#define GTS_CONTAINEE_CLASS(klass) GTS_OBJECT_CLASS_CAST (klass, GtsContaineeClass, gts_containee_class ())
#define GTS_CONTAINER_CLASS(klass) GTS_OBJECT_CLASS_CAST (klass, GtsContainerClass, gts_container_class ())
#define GTS_CONTAINEE(obj) GTS_OBJECT_CAST (obj, GtsContainee, gts_containee_class ())
#define GTS_CONTAINER(obj) GTS_OBJECT_CAST (obj, GtsContainer, gts_container_class ())
#define GTS_SLIST_CONTAINER(obj) GTS_OBJECT_CAST (obj, GtsSListContainer, gts_slist_container_class ())
#define GTS_SLIST_CONTAINEE(obj) GTS_OBJECT_CAST (obj, GtsSListContainee, gts_slist_containee_class ())
GtsContaineeClass * gts_containee_class (void);
GtsSListContaineeClass * gts_slist_containee_class (void);
GtsContainerClass * gts_container_class (void);
GtsSListContainerClass * gts_slist_container_class (void);
GtsContainer * gts_container_new (GtsContainerClass * klass);
void gts_container_add (GtsContainer * c, GtsContainee * item);
void gts_container_remove (GtsContainer * c, GtsContainee * item);
void gts_container_foreach (GtsContainer * c, GtsFunc func, gpointer data);
guint gts_container_size (GtsContainer * c);

typedef struct { GtsObject object; gdouble x, y, z; } GtsPoint;
typedef struct { GtsObjectClass parent_class; gboolean binary; } GtsPointClass;
typedef struct { GtsPoint p; GSList * segments; } GtsVertex;
typedef struct { GtsPointClass parent_class; void (* intersection_attributes) (GtsVertex *, GtsObject *, GtsObject *); } GtsVertexClass;
typedef struct { GtsObject object; GtsVertex * v1, * v2; } GtsSegment;
typedef struct { GtsObjectClass parent_class; } GtsSegmentClass;
typedef struct { GtsSegment segment; GSList * triangles; } GtsEdge;
typedef struct { GtsSegmentClass parent_class; } GtsEdgeClass;
typedef struct { GtsObject object; GtsEdge * e1, * e2, * e3; } GtsTriangle;
typedef struct { GtsObjectClass parent_class; } GtsTriangleClass;
typedef struct { GtsTriangle triangle; GSList * surfaces; } GtsFace;
typedef struct { GtsTriangleClass parent_class; } GtsFaceClass;
typedef struct _GtsSurface GtsSurface;
typedef struct { GtsObjectClass parent_class; void (* add_face) (GtsSurface *, GtsFace *); void (* remove_face) (GtsSurface *, GtsFace *); } GtsSurfaceClass;
typedef struct { GtsObject object; gpointer bounded; gdouble x1, y1, z1, x2, y2, z2; } GtsBBox;
typedef struct { gdouble min, max, sum, sum2, mean, stddev; guint n; } GtsRange;
void gts_range_init (GtsRange * r);
void gts_range_add_value (GtsRange * r, gdouble val);
void gts_range_update (GtsRange * r);
gboolean gts_bboxes_are_overlapping (GtsBBox * bb1, GtsBBox * bb2);
GtsPointClass * gts_point_class (void);
GtsVertexClass * gts_vertex_class (void);
GtsEdgeClass * gts_edge_class (void);
GtsFaceClass * gts_face_class (void);

typedef struct { GtsHashContainer c; gpointer graph_class, node_class, edge_class; } GtsGraph;
typedef struct { GtsHashContainerClass parent_class; gfloat (* weight) (GtsGraph *); } GtsGraphClass;
typedef struct { GtsGraph graph; gfloat weight; } GtsWGraph;
typedef struct { GtsGraphClass parent_class; } GtsWGraphClass;
typedef struct { GtsSListContainer container; guint level; } GtsGNode;
typedef struct { GtsSListContainerClass parent_class; gfloat (* weight) (GtsGNode *); void (* write) (GtsGNode *, FILE *); } GtsGNodeClass;
typedef struct { GtsContainee containee; GtsGNode * n1, * n2; } GtsGEdge;
typedef struct { GtsContaineeClass parent_class; GtsGEdge * (* link) (GtsGEdge *, GtsGNode *, GtsGNode *); gfloat (* weight) (GtsGEdge *); void (* write) (GtsGEdge *, FILE *); } GtsGEdgeClass;
GtsGraphClass * gts_graph_class (void);
GtsWGraphClass * gts_wgraph_class (void);
GtsGNodeClass * gts_gnode_class (void);
GtsGEdgeClass * gts_gedge_class (void);
#define GTS_GRAPH(obj) GTS_OBJECT_CAST (obj, GtsGraph, gts_graph_class ())
#define GTS_GNODE(obj) GTS_OBJECT_CAST (obj, GtsGNode, gts_gnode_class ())

#endif /* GFSB200_CHECK_GTS_H */

/* TEST INFRASTRUCTURE ONLY -- minimal stand-in for <gts.h>/<glib.h>.
 *
 * GLib and GTS are not installed in this image.  The reference's arithmetic
 * core (src/ftt.c, src/fluid.c) only uses them for typedefs, assertions,
 * malloc wrappers and a handful of containers; this header supplies just
 * enough of that surface for the two files to compile UNMODIFIED from
 * /root/reference/src (see oracle/Makefile).  Nothing here is product code.
 */
#ifndef GFSB200_ORACLE_SHIM_GTS_H
#define GFSB200_ORACLE_SHIM_GTS_H

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#include <limits.h>
#include <math.h>

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
typedef unsigned short guint16;
typedef unsigned int guint32;
typedef size_t gsize;

#ifndef TRUE
# define TRUE 1
#endif
#ifndef FALSE
# define FALSE 0
#endif

#define G_MAXDOUBLE DBL_MAX
#define G_MAXINT INT_MAX
#define G_MININT INT_MIN
#define GTS_C_VAR extern
#define G_LOG_DOMAIN "gfs-oracle"
#define G_LOG_LEVEL_ERROR 4
#define G_STMT_START do
#define G_STMT_END while (0)
#define G_GNUC_UNUSED __attribute__((unused))

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

void gfs_shim_log (int level, const char * fmt, ...);
#define g_log(domain, level, ...) gfs_shim_log (level, __VA_ARGS__)
#define g_warning(...) gfs_shim_log (2, __VA_ARGS__)
#define g_error(...) do { gfs_shim_log (4, __VA_ARGS__); abort (); } while (0)

#define g_assert(expr) do { if (!(expr)) { \
  fprintf (stderr, "%s:%d: assertion failed: %s\n", __FILE__, __LINE__, #expr); \
  abort (); } } while (0)
#define g_assert_not_reached() do { \
  fprintf (stderr, "%s:%d: should not be reached\n", __FILE__, __LINE__); \
  abort (); } while (0)
#define g_return_if_fail(expr) do { if (!(expr)) { \
  fprintf (stderr, "%s:%d: check failed: %s\n", __FILE__, __LINE__, #expr); \
  return; } } while (0)
#define g_return_val_if_fail(expr, val) do { if (!(expr)) { \
  fprintf (stderr, "%s:%d: check failed: %s\n", __FILE__, __LINE__, #expr); \
  return (val); } } while (0)

static inline gpointer g_malloc (gsize n) {
  gpointer p = malloc (n ? n : 1);
  if (!p) abort ();
  return p;
}
static inline gpointer g_malloc0 (gsize n) {
  gpointer p = calloc (1, n ? n : 1);
  if (!p) abort ();
  return p;
}
static inline gpointer g_realloc (gpointer q, gsize n) {
  gpointer p = realloc (q, n ? n : 1);
  if (!p) abort ();
  return p;
}
#define g_free(p) free (p)

/* containers (only what ftt.c / fluid.c touch) */
typedef struct _GSList GSList;
struct _GSList { gpointer data; GSList * next; };

typedef struct { gpointer * pdata; guint len; guint alloc; } GPtrArray;
GPtrArray * g_ptr_array_new (void);
void g_ptr_array_add (GPtrArray * a, gpointer p);
gpointer * g_ptr_array_free (GPtrArray * a, gboolean free_seg);

typedef struct { gchar * data; guint len; guint alloc; guint esize; } GArray;
GArray * g_array_new (gboolean zero_terminated, gboolean clear, guint element_size);
GArray * g_array_append_vals (GArray * a, gconstpointer data, guint len);
gchar * g_array_free (GArray * a, gboolean free_segment);
#define g_array_append_val(a, v) g_array_append_vals (a, &(v), 1)
#define g_array_index(a, t, i) (((t *) (void *) (a)->data)[i])

/* GTS bits */
typedef struct { gpointer klass; guint flags; gpointer reserved; } GtsObject;
#define GTS_OBJECT(o) ((GtsObject *) (o))
typedef struct { GtsObject object; gdouble x, y, z; } GtsPoint;
typedef struct { GtsObject object; gpointer bounded;
                 gdouble x1, y1, z1, x2, y2, z2; } GtsBBox;
typedef struct { gdouble min, max, sum, sum2, mean, stddev; guint n; } GtsRange;
void gts_range_init (GtsRange * r);
void gts_range_add_value (GtsRange * r, gdouble val);
void gts_range_update (GtsRange * r);
gboolean gts_bboxes_are_overlapping (GtsBBox * bb1, GtsBBox * bb2);

typedef enum {
  GTS_NONE = 1 << 8, GTS_INT = 1 << 9, GTS_UINT = 1 << 10, GTS_FLOAT = 1 << 11,
  GTS_DOUBLE = 1 << 12, GTS_STRING = 1 << 13, GTS_FILE = 1 << 14, GTS_ERROR = 1 << 15
} GtsTokenType;
typedef struct { gchar * str; gsize len; } GString;
typedef struct {
  FILE * fp; gchar * s, * s1;
  guint line, pos;
  GString * token;
  GtsTokenType type;
  gchar * error;
  guint curline, curpos, scope, scope_max;
  gint next_token;
  gchar * delimiters, * comments, * tokens;
} GtsFile;
void gts_file_error (GtsFile * f, const gchar * format, ...);
void gts_file_next_token (GtsFile * f);
void gts_file_first_token_after (GtsFile * f, GtsTokenType type);
guint gts_file_read (GtsFile * f, gpointer ptr, guint size, guint nmemb);

#endif /* GFSB200_ORACLE_SHIM_GTS_H */

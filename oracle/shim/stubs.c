/* TEST INFRASTRUCTURE ONLY -- link-time stand-ins for the GLib/GTS/libgfs
 * symbols that the reference's src/ftt.c and src/fluid.c reference but that
 * the particulate hot path never reaches with our solid-free, metric-free
 * synthetic domains.  Where a function IS reachable its reference semantics
 * are restated with a file:line citation.
 */
#include <stdarg.h>
#include "gfs_shadow.h"

void gfs_shim_log (int level, const char * fmt, ...)
{
  va_list ap;
  va_start (ap, fmt);
  fprintf (stderr, "[gfs-oracle:%d] ", level);
  vfprintf (stderr, fmt, ap);
  fputc ('\n', stderr);
  va_end (ap);
  if (level >= 4)
    abort ();
}

/* --- GPtrArray / GArray: growable arrays (used by ftt traversal caches and
 *     the linear-problem stencils; not on the hot path) --- */
GPtrArray * g_ptr_array_new (void)
{
  return g_malloc0 (sizeof (GPtrArray));
}

void g_ptr_array_add (GPtrArray * a, gpointer p)
{
  if (a->len == a->alloc) {
    a->alloc = a->alloc ? 2*a->alloc : 16;
    a->pdata = g_realloc (a->pdata, a->alloc*sizeof (gpointer));
  }
  a->pdata[a->len++] = p;
}

gpointer * g_ptr_array_free (GPtrArray * a, gboolean free_seg)
{
  gpointer * seg = a->pdata;
  if (free_seg) {
    free (seg);
    seg = NULL;
  }
  free (a);
  return seg;
}

GArray * g_array_new (gboolean zero_terminated, gboolean clear, guint element_size)
{
  GArray * a = g_malloc0 (sizeof (GArray));
  a->esize = element_size;
  return a;
}

GArray * g_array_append_vals (GArray * a, gconstpointer data, guint len)
{
  if (a->len + len > a->alloc) {
    a->alloc = MAX (2*a->alloc, a->len + len + 16);
    a->data = g_realloc (a->data, (gsize) a->alloc*a->esize);
  }
  memcpy (a->data + (gsize) a->len*a->esize, data, (gsize) len*a->esize);
  a->len += len;
  return a;
}

gchar * g_array_free (GArray * a, gboolean free_segment)
{
  gchar * seg = a->data;
  if (free_segment) {
    free (seg);
    seg = NULL;
  }
  free (a);
  return seg;
}

/* --- GtsRange (statistics helpers, off the hot path) --- */
void gts_range_init (GtsRange * r)
{
  r->max = - G_MAXDOUBLE;
  r->min = G_MAXDOUBLE;
  r->sum = r->sum2 = r->mean = r->stddev = 0.;
  r->n = 0;
}

void gts_range_add_value (GtsRange * r, gdouble val)
{
  if (val < r->min) r->min = val;
  if (val > r->max) r->max = val;
  r->sum += val;
  r->sum2 += val*val;
  r->n++;
}

void gts_range_update (GtsRange * r)
{
  if (r->n > 0) {
    r->mean = r->sum/r->n;
    r->stddev = sqrt (fabs (r->sum2 - r->sum*r->sum/(gdouble) r->n)/(gdouble) r->n);
  }
}

gboolean gts_bboxes_are_overlapping (GtsBBox * bb1, GtsBBox * bb2)
{
  if (bb1 == bb2) return TRUE;
  return !(bb1->x1 > bb2->x2 || bb2->x1 > bb1->x2 ||
           bb1->y1 > bb2->y2 || bb2->y1 > bb1->y2 ||
           bb1->z1 > bb2->z2 || bb2->z1 > bb1->z2);
}

/* --- GtsFile: the tree readers in ftt.c are never called by the oracle --- */
void gts_file_error (GtsFile * f, const gchar * format, ...)
{
  f->type = GTS_ERROR;
}
void gts_file_next_token (GtsFile * f) { f->type = GTS_ERROR; }
void gts_file_first_token_after (GtsFile * f, GtsTokenType type) { f->type = GTS_ERROR; }
guint gts_file_read (GtsFile * f, gpointer ptr, guint size, guint nmemb) { return 0; }

/* --- libgfs symbols referenced by fluid.c --- */
GfsSurfaceGenericBcClass gfs_shim_surface_bc_class = { NULL };

/* reference: src/domain.h:504-508 */
gdouble gfs_cell_volume (const FttCell * cell, const GfsDomain * domain)
{
  double v = ftt_cell_volume (cell)*(GFS_IS_MIXED (cell) ? GFS_STATE (cell)->solid->a : 1.);
  return domain->cell_metric ? (* domain->cell_metric) (domain, cell)*v : v;
}

/* reference: src/domain.c:453-459 */
gdouble gfs_domain_cell_fraction (const GfsDomain * domain, const FttCell * cell)
{
  gdouble a = GFS_IS_MIXED (cell) ? GFS_STATE (cell)->solid->a : 1.;
  if (domain->cell_metric)
    a *= (* domain->cell_metric) (domain, cell);
  return a;
}

/* reference: src/domain.c:415-421 */
gdouble gfs_domain_face_fraction (const GfsDomain * domain, const FttCellFace * face)
{
  gdouble f = GFS_FACE_FRACTION (face);
  if (domain->face_metric)
    f *= (* domain->face_metric) (domain, face);
  return f;
}

/* reference: src/solid.c:1264 -- only meaningful for mixed cells, which the
 * oracle's worlds do not contain; non-mixed cells return the cell centre. */
void gfs_cell_cm (const FttCell * cell, FttVector * cm)
{
  if (GFS_IS_MIXED (cell))
    *cm = GFS_STATE (cell)->solid->cm;
  else
    ftt_cell_pos (cell, cm);
}

void gfs_face_ca (const FttCellFace * face, FttVector * ca)
{
  ftt_face_pos (face, ca);
}

void gfs_solid_normal (const FttCell * cell, FttVector * n)
{
  n->x = n->y = n->z = 0.;
}

void gfs_eigenvalues (gdouble a[FTT_DIMENSION][FTT_DIMENSION],
                      gdouble d[FTT_DIMENSION],
                      gdouble v[FTT_DIMENSION][FTT_DIMENSION])
{
  fprintf (stderr, "gfs_eigenvalues: not available in the oracle shim\n");
  abort ();
}

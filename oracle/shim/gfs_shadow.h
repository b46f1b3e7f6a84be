/* TEST INFRASTRUCTURE ONLY -- shadow of the parts of Gerris' variable.h,
 * domain.h and solid.h that src/fluid.c touches.
 *
 * fluid.c does `#include "variable.h"` etc., and a quote-include resolves in
 * the includer's own directory first, so the real headers (which drag in all
 * of GTS) would be found.  oracle/Makefile therefore pre-defines their include
 * guards (__VARIABLE_H__, __DOMAIN_H__, __SOLID_H__) and force-includes this
 * file, which supplies the few types/macros/prototypes fluid.c really needs.
 * The reference sources themselves are compiled unmodified, where they lie.
 *
 * Layout note: struct _GfsVariable / _GfsDomain here only need to agree with
 * oracle/particulate_port.c (which creates them), not with libgfs.
 */
#ifndef GFSB200_ORACLE_SHIM_SHADOW_H
#define GFSB200_ORACLE_SHIM_SHADOW_H

#include "fluid.h"   /* the real one, from /root/reference/src */

/* --- variable.h (reference: src/variable.h:36-58,76) --- */
typedef struct _GfsSurfaceGenericBc GfsSurfaceGenericBc;
typedef struct _GfsSurfaceGenericBcClass GfsSurfaceGenericBcClass;
struct _GfsSurfaceGenericBc { GtsObject parent; };
struct _GfsSurfaceGenericBcClass {
  void (* bc) (FttCell *, GfsSurfaceGenericBc *);
};
extern GfsSurfaceGenericBcClass gfs_shim_surface_bc_class;
#define GFS_SURFACE_GENERIC_BC_CLASS(k) (&gfs_shim_surface_bc_class)

struct _GfsVariable {
  GtsObject parent;
  guint i;
  FttComponent component;
  GfsVariable * vector[FTT_DIMENSION];
  gchar * name, * description;
  gboolean centered;
  void (* fine_coarse) (FttCell *, GfsVariable *);
  void (* coarse_fine) (FttCell *, GfsVariable *);
  gpointer sources;
  GfsSurfaceGenericBc * surface_bc;
  gpointer default_bc;
  GfsDomain * domain;
  FttCellCleanupFunc cleanup;
  gdouble units;
  gdouble orientation;
  gboolean even;
};
#define GFS_VARIABLE(obj) ((GfsVariable *) (obj))
#define GFS_VALUE(cell, v) ((&GFS_STATE (cell)->place_holder)[(v)->i])

/* --- utils.h (reference: src/utils.h:80-84) --- */
#define GFS_NODATA G_MAXDOUBLE
#define GFS_HAS_DATA(cell, v) (GFS_VALUE (cell, v) != GFS_NODATA)
#define GFS_DOUBLE_TO_POINTER(d) (*((gpointer *) &(d)))

/* --- domain.h (reference: src/domain.h:55-129; only members fluid.c reads) --- */
struct _GfsDomain {
  GSList * variables;
  gdouble (* face_metric)  (const GfsDomain *, const FttCellFace *);
  gdouble (* cell_metric)  (const GfsDomain *, const FttCell *);
  void    (* solid_metric) (const GfsDomain *, const FttCell *, FttVector *);
  gdouble (* scale_metric) (const GfsDomain *, const FttCell *, FttComponent);
  guint rootlevel;
};
gdouble gfs_cell_volume (const FttCell * cell, const GfsDomain * domain);
gdouble gfs_domain_cell_fraction (const GfsDomain * domain, const FttCell * cell);
gdouble gfs_domain_face_fraction (const GfsDomain * domain, const FttCellFace * face);
gdouble gfs_domain_face_fraction_right (const GfsDomain * domain, const FttCellFace * face);
gdouble gfs_domain_solid_metric (const GfsDomain * domain, const FttCell * cell, FttVector * m);

/* --- simulation.h:158 -- the oracle's domains are never axisymmetric --- */
#define GFS_IS_AXI(obj) (FALSE)

/* --- poisson.h bits (linear problem / stencils, untouched by the hot path) --- */
struct _GfsLinearProblem {
  GPtrArray * LP;
  GArray * rhs, * lhs;
  GfsVariable * id, * neighbor, * neighborw;
  gint istart;
  guint maxsize;
};

/* --- solid.h --- */
void gfs_cell_cm (const FttCell * cell, FttVector * cm);
void gfs_face_ca (const FttCellFace * face, FttVector * ca);
void gfs_solid_normal (const FttCell * cell, FttVector * n);
void gfs_eigenvalues (gdouble a[FTT_DIMENSION][FTT_DIMENSION],
                      gdouble d[FTT_DIMENSION],
                      gdouble v[FTT_DIMENSION][FTT_DIMENSION]);

#endif /* GFSB200_ORACLE_SHIM_SHADOW_H */

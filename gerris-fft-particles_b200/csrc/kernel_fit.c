/* kernel_fit.c -- recognise the smoothing kernel of a GfsSourceParticulate.
 *
 * The reference lets the user write `kernel = <GfsFunction>' (an arbitrary
 * compiled expression of the normalised offset x,y,z; read at
 * modules/particulatecommon.c:2282-2290, evaluated by
 * gfs_function_spatial_value in kernel_volume/diffuse_force :2108-2175).  The
 * device evaluates three closed forms (include/gfsb200.h: constant, Gaussian,
 * compactly supported polynomial).  A compiled expression cannot be inspected,
 * but it can be PROBED: the host module samples it and asks here whether one
 * of the closed forms reproduces every sample to 1e-12 of its peak; if none
 * does, the list stays on the reference's CPU event.
 */
#include <math.h>
#include <string.h>
#include "gfsb200_internal.h"

#define N_PROBE 40
#define FIT_TOL 1e-12

static void probe_point (int k, int dim, double q[3])
{
  /* radii from 0.05 to ~4 particle radii, directions off the axes; a fixed
     low-discrepancy pattern so that the verdict is reproducible */
  const double r = 0.05 + 4.*(k + 0.5)/N_PROBE*(k + 0.5)/N_PROBE;
  const double phi = 2.399963229728653*k;           /* golden angle */
  const double ct = dim == 3 ? 1. - 2.*((k*7 + 3) % N_PROBE + 0.5)/N_PROBE : 0.;
  const double st = sqrt (1. - ct*ct);
  q[0] = r*st*cos (phi); q[1] = r*st*sin (phi); q[2] = dim == 3 ? r*ct : 0.;
}

static double closed_form (const gfsb200_kernel * k, const double q[3])
{
  const double r2 = q[0]*q[0] + q[1]*q[1] + q[2]*q[2];
  if (k->kind == GFSB200_KERNEL_GAUSSIAN)
    return k->a*exp (-k->b*r2);
  if (k->kind == GFSB200_KERNEL_COMPACT) {
    const double t = 1. - k->b*r2;
    double v = k->a;
    int i;
    if (t <= 0.) return 0.;
    for (i = 0; i < k->p; i++) v *= t;
    return v;
  }
  return k->a;
}

static int reproduces (const gfsb200_kernel * k, int dim, const double q[][3], const double * f, int n)
{
  int i;
  (void) dim;
  for (i = 0; i < n; i++)
    if (!(fabs (closed_form (k, q[i]) - f[i]) <= FIT_TOL*fabs (k->a)))
      return 0;
  return 1;
}

int gfsb200_kernel_fit (gfsb200_kernel_func f, void * data, int dim, gfsb200_kernel * out)
{
  double q[N_PROBE + 2][3], v[N_PROBE + 2];
  gfsb200_kernel k;
  int i, p;

  if (!f || !out || (dim != 2 && dim != 3))
    return gfsb200_fail (GFSB200_ERR_ARG, "kernel_fit: bad argument");
  memset (&k, 0, sizeof k);
  /* samples: the origin, a point at unit radius, then the pattern */
  q[0][0] = q[0][1] = q[0][2] = 0.;
  q[1][0] = 0.6; q[1][1] = 0.8; q[1][2] = 0.;
  for (i = 0; i < N_PROBE; i++)
    probe_point (i, dim, q[i + 2]);
  for (i = 0; i < N_PROBE + 2; i++) {
    v[i] = (* f) (q[i][0], q[i][1], q[i][2], data);
    if (!(v[i] == v[i]) || fabs (v[i]) > 1e300)
      return gfsb200_fail (GFSB200_ERR_UNSUPPORTED, "kernel_fit: kernel is not finite at (%g,%g,%g)",
			   q[i][0], q[i][1], q[i][2]);
  }
  k.a = v[0];
  k.p = 1;
  if (k.a == 0.)
    return gfsb200_fail (GFSB200_ERR_UNSUPPORTED, "kernel_fit: kernel vanishes at the particle centre");

  k.kind = GFSB200_KERNEL_CONSTANT;
  for (i = 1; i < N_PROBE + 2 && v[i] == k.a; i++)
    ;
  if (i == N_PROBE + 2) { *out = k; return GFSB200_OK; }

  /* a exp(-b r^2): b from the unit-radius sample */
  if (v[1]/k.a > 0.) {
    k.kind = GFSB200_KERNEL_GAUSSIAN;
    k.b = -log (v[1]/k.a);
    if (reproduces (&k, dim, q, v, N_PROBE + 2)) { *out = k; return GFSB200_OK; }
  }

  /* a (1 - b r^2)^p: b from the innermost pattern sample (r = 0.05..), p = 1..8 */
  {
    const double * q0 = q[2];
    const double r2 = q0[0]*q0[0] + q0[1]*q0[1] + q0[2]*q0[2];
    for (p = 1; p <= 8; p++) {
      const double ratio = v[2]/k.a;
      if (ratio <= 0.) break;
      k.kind = GFSB200_KERNEL_COMPACT;
      k.p = p;
      k.b = (1. - pow (ratio, 1./p))/r2;
      if (reproduces (&k, dim, q, v, N_PROBE + 2)) {
	/* b is known to ~1e-13/r2 from one sample; polish it on the sample with the
	   largest lever arm still inside the support */
	int best = -1;
	double best_r2 = 0.;
	for (i = 1; i < N_PROBE + 2; i++) {
	  const double s = q[i][0]*q[i][0] + q[i][1]*q[i][1] + q[i][2]*q[i][2];
	  if (v[i]/k.a > 0.05 && s > best_r2) { best = i; best_r2 = s; }
	}
	if (best >= 0) {
	  gfsb200_kernel k2 = k;
	  k2.b = (1. - pow (v[best]/k.a, 1./p))/best_r2;
	  if (reproduces (&k2, dim, q, v, N_PROBE + 2)) k = k2;
	}
	*out = k;
	return GFSB200_OK;
      }
    }
  }
  return gfsb200_fail (GFSB200_ERR_UNSUPPORTED,
		       "kernel_fit: kernel is none of constant / a exp(-b r^2) / a (1 - b r^2)^p");
}

/* Brent bracketing solver, Newton polishing solver and the two convergence tests, with the
 * state machine GSL documents.  Test infrastructure, see gsl_errno.h. */
#ifndef QB_GSL_ROOTS_H
#define QB_GSL_ROOTS_H
#include <math.h>
#include <stdlib.h>
#include "gsl_errno.h"
#include "gsl_math.h"

static inline int gsl_root_test_delta (double x1, double x0, double epsabs, double epsrel) {
  const double tolerance = epsabs + epsrel * fabs (x1);
  if (epsabs < 0.0 || epsrel < 0.0) GSL_ERROR ("negative tolerance", GSL_EINVAL);
  if (fabs (x1 - x0) < tolerance || x1 == x0) return GSL_SUCCESS;
  return GSL_CONTINUE;
}
static inline int gsl_root_test_interval (double x_lower, double x_upper, double epsabs, double epsrel) {
  const double abs_lower = fabs (x_lower), abs_upper = fabs (x_upper);
  double min_abs;
  if (epsabs < 0.0 || epsrel < 0.0) GSL_ERROR ("negative tolerance", GSL_EINVAL);
  if (x_lower > x_upper) GSL_ERROR ("lower bound larger than upper bound", GSL_EINVAL);
  if ((x_lower > 0.0 && x_upper > 0.0) || (x_lower < 0.0 && x_upper < 0.0))
    min_abs = abs_lower < abs_upper ? abs_lower : abs_upper;
  else
    min_abs = 0;
  if (fabs (x_upper - x_lower) < epsabs + epsrel * min_abs) return GSL_SUCCESS;
  return GSL_CONTINUE;
}

/* ---- bracketing (Brent) ---- */
typedef struct { const char* name; } gsl_root_fsolver_type;
static const gsl_root_fsolver_type qb_brent_type = { "brent" };
static const gsl_root_fsolver_type* const gsl_root_fsolver_brent = &qb_brent_type;
typedef struct {
  const gsl_root_fsolver_type* type;
  gsl_function* function;
  double root, x_lower, x_upper;
  double a, b, c, d, e, fa, fb, fc;
} gsl_root_fsolver;
static inline gsl_root_fsolver* gsl_root_fsolver_alloc (const gsl_root_fsolver_type* T) {
  gsl_root_fsolver* s = (gsl_root_fsolver*) calloc (1, sizeof (gsl_root_fsolver));
  s->type = T;
  return s;
}
static inline void gsl_root_fsolver_free (gsl_root_fsolver* s) { free (s); }
static inline const char* gsl_root_fsolver_name (const gsl_root_fsolver* s) { return s->type->name; }
static inline double gsl_root_fsolver_root (const gsl_root_fsolver* s) { return s->root; }
static inline double gsl_root_fsolver_x_lower (const gsl_root_fsolver* s) { return s->x_lower; }
static inline double gsl_root_fsolver_x_upper (const gsl_root_fsolver* s) { return s->x_upper; }
static inline int gsl_root_fsolver_set (gsl_root_fsolver* s, gsl_function* f, double x_lower, double x_upper) {
  if (x_lower > x_upper) GSL_ERROR ("invalid interval (lower > upper)", GSL_EINVAL);
  s->function = f;
  s->x_lower = x_lower;
  s->x_upper = x_upper;
  s->root = 0.5 * (x_lower + x_upper);
  const double f_lower = GSL_FN_EVAL (f, x_lower), f_upper = GSL_FN_EVAL (f, x_upper);
  if (!isfinite (f_lower) || !isfinite (f_upper)) GSL_ERROR ("function value is not finite", GSL_EBADFUNC);
  s->a = x_lower; s->fa = f_lower;
  s->b = x_upper; s->fb = f_upper;
  s->c = x_upper; s->fc = f_upper;
  s->d = x_upper - x_lower;
  s->e = x_upper - x_lower;
  if ((f_lower < 0.0 && f_upper < 0.0) || (f_lower > 0.0 && f_upper > 0.0))
    GSL_ERROR ("endpoints do not straddle y=0", GSL_EINVAL);
  return GSL_SUCCESS;
}
static inline int gsl_root_fsolver_iterate (gsl_root_fsolver* s) {
  double tol, m;
  int ac_equal = 0;
  double a = s->a, b = s->b, c = s->c, fa = s->fa, fb = s->fb, fc = s->fc, d = s->d, e = s->e;
  if ((fb < 0 && fc < 0) || (fb > 0 && fc > 0)) { ac_equal = 1; c = a; fc = fa; d = b - a; e = b - a; }
  if (fabs (fc) < fabs (fb)) { ac_equal = 1; a = b; b = c; c = a; fa = fb; fb = fc; fc = fa; }
  tol = 0.5 * GSL_DBL_EPSILON * fabs (b);
  m = 0.5 * (c - b);
  if (fb == 0) {
    s->root = b; s->x_lower = b; s->x_upper = b;
    return GSL_SUCCESS;
  }
  if (fabs (m) <= tol) {
    s->root = b;
    if (b < c) { s->x_lower = b; s->x_upper = c; } else { s->x_lower = c; s->x_upper = b; }
    return GSL_SUCCESS;
  }
  if (fabs (e) < tol || fabs (fa) <= fabs (fb)) {
    d = m; e = m;                      /* bisection */
  } else {
    double p, q, r;
    double sv = fb / fa;
    if (ac_equal) { p = 2 * m * sv; q = 1 - sv; }            /* secant */
    else {                                                   /* inverse quadratic */
      q = fa / fc; r = fb / fc;
      p = sv * (2 * m * q * (q - r) - (b - a) * (r - 1));
      q = (q - 1) * (r - 1) * (sv - 1);
    }
    if (p > 0) q = -q; else p = -p;
    {
      const double t1 = 3 * m * q - fabs (tol * q), t2 = fabs (e * q);
      if (2 * p < (t1 < t2 ? t1 : t2)) { e = d; d = p / q; }
      else { d = m; e = m; }
    }
  }
  a = b; fa = fb;
  if (fabs (d) > tol) b += d; else b += (m > 0 ? +tol : -tol);
  fb = GSL_FN_EVAL (s->function, b);
  if (!isfinite (fb)) GSL_ERROR ("function value is not finite", GSL_EBADFUNC);
  s->a = a; s->b = b; s->c = c; s->d = d; s->e = e; s->fa = fa; s->fb = fb; s->fc = fc;
  s->root = b;
  if ((fb < 0 && fc < 0) || (fb > 0 && fc > 0)) c = a;
  if (b < c) { s->x_lower = b; s->x_upper = c; } else { s->x_lower = c; s->x_upper = b; }
  return GSL_SUCCESS;
}

/* ---- polishing (Newton) ---- */
typedef struct { const char* name; } gsl_root_fdfsolver_type;
static const gsl_root_fdfsolver_type qb_newton_type = { "newton" };
static const gsl_root_fdfsolver_type* const gsl_root_fdfsolver_newton = &qb_newton_type;
typedef struct {
  const gsl_root_fdfsolver_type* type;
  gsl_function_fdf* fdf;
  double root, f, df;
} gsl_root_fdfsolver;
static inline gsl_root_fdfsolver* gsl_root_fdfsolver_alloc (const gsl_root_fdfsolver_type* T) {
  gsl_root_fdfsolver* s = (gsl_root_fdfsolver*) calloc (1, sizeof (gsl_root_fdfsolver));
  s->type = T;
  return s;
}
static inline void gsl_root_fdfsolver_free (gsl_root_fdfsolver* s) { free (s); }
static inline const char* gsl_root_fdfsolver_name (const gsl_root_fdfsolver* s) { return s->type->name; }
static inline double gsl_root_fdfsolver_root (const gsl_root_fdfsolver* s) { return s->root; }
static inline int gsl_root_fdfsolver_set (gsl_root_fdfsolver* s, gsl_function_fdf* fdf, double root) {
  s->fdf = fdf;
  s->root = root;
  GSL_FN_FDF_EVAL_F_DF (fdf, root, &s->f, &s->df);
  return GSL_SUCCESS;
}
static inline int gsl_root_fdfsolver_iterate (gsl_root_fdfsolver* s) {
  double root_new, f_new, df_new;
  if (s->df == 0.0) GSL_ERROR ("derivative is zero", GSL_EZERODIV);
  root_new = s->root - (s->f / s->df);
  s->root = root_new;
  GSL_FN_FDF_EVAL_F_DF (s->fdf, root_new, &f_new, &df_new);
  s->f = f_new; s->df = df_new;
  if (!isfinite (f_new)) GSL_ERROR ("function value is not finite", GSL_EBADFUNC);
  if (!isfinite (df_new)) GSL_ERROR ("derivative value is not finite", GSL_EBADFUNC);
  return GSL_SUCCESS;
}
#endif

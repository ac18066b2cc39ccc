/* The three densities quaff evaluates (test infrastructure, see gsl_errno.h). */
#ifndef QB_GSL_RANDIST_H
#define QB_GSL_RANDIST_H
#include <math.h>
#include <stddef.h>
/* Pr(k) = Gamma(n+k)/(Gamma(k+1) Gamma(n)) p^n (1-p)^k, written the way GSL writes it */
static inline double gsl_ran_negative_binomial_pdf (const unsigned int k, const double p, double n) {
  const double f = lgamma (k + n), a = lgamma (n), b = lgamma (k + 1.0);
  return exp (f - a - b) * pow (p, n) * pow (1 - p, (double) k);
}
static inline double gsl_ran_beta_pdf (const double x, const double a, const double b) {
  if (x < 0 || x > 1) return 0;
  const double gab = lgamma (a + b), ga = lgamma (a), gb = lgamma (b);
  if (x == 0.0 || x == 1.0) {
    if (a > 1.0 && b > 1.0) return 0.0;
    return exp (gab - ga - gb) * pow (x, a - 1) * pow (1 - x, b - 1);
  }
  return exp (gab - ga - gb + log (x) * (a - 1) + log1p (-x) * (b - 1));
}
static inline double gsl_ran_dirichlet_pdf (const size_t K, const double alpha[], const double theta[]) {
  double lp = 0, sum_alpha = 0;
  for (size_t i = 0; i < K; ++i) lp += (alpha[i] - 1.0) * log (theta[i]);
  for (size_t i = 0; i < K; ++i) sum_alpha += alpha[i];
  lp += lgamma (sum_alpha);
  for (size_t i = 0; i < K; ++i) lp -= lgamma (alpha[i]);
  return exp (lp);
}
#endif

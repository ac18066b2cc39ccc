/* digamma / trigamma for positive arguments: upward recurrence to x >= 10, then the
 * asymptotic (Bernoulli) series.  Test infrastructure, see gsl_errno.h. */
#ifndef QB_GSL_SF_PSI_H
#define QB_GSL_SF_PSI_H
#include <math.h>
static inline double gsl_sf_psi (double x) {
  double r = 0;
  while (x < 10) { r -= 1 / x; x += 1; }
  const double i = 1 / x, i2 = i * i;
  /* ln x - 1/2x - sum B_2n / (2n x^2n) */
  const double s = i2 * (1.0/12 - i2 * (1.0/120 - i2 * (1.0/252 - i2 * (1.0/240 - i2 * (1.0/132 - i2 * (691.0/32760 - i2 * (1.0/12)))))));
  return r + log (x) - 0.5 * i - s;
}
static inline double gsl_sf_psi_1 (double x) {
  double r = 0;
  while (x < 10) { r += 1 / (x * x); x += 1; }
  const double i = 1 / x, i2 = i * i;
  /* 1/x + 1/2x^2 + sum B_2n / x^(2n+1) */
  const double s = i * i2 * (1.0/6 - i2 * (1.0/30 - i2 * (1.0/42 - i2 * (1.0/30 - i2 * (5.0/66 - i2 * (691.0/2730 - i2 * (7.0/6)))))));
  return r + i + 0.5 * i2 + s;
}
#endif

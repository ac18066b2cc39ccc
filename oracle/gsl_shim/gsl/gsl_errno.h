/* Header-only stand-in for the parts of GSL that ihh/quaff touches.
 * TEST INFRASTRUCTURE ONLY: used to build the reference oracle (oracle/_ref); GSL is not
 * installed in this image.  Spec: SURVEY.md section 9.6.  Error codes follow GSL's numbering. */
#ifndef QB_GSL_ERRNO_H
#define QB_GSL_ERRNO_H
#include <stdio.h>
enum {
  GSL_SUCCESS = 0, GSL_FAILURE = -1, GSL_CONTINUE = -2,
  GSL_EDOM = 1, GSL_ERANGE = 2, GSL_EFAULT = 3, GSL_EINVAL = 4,
  GSL_EBADFUNC = 9, GSL_ERUNAWAY = 11, GSL_EZERODIV = 12
};
static inline const char* gsl_strerror (int code) {
  switch (code) {
  case GSL_SUCCESS: return "success";
  case GSL_FAILURE: return "failure";
  case GSL_CONTINUE: return "the iteration has not converged yet";
  case GSL_EDOM: return "input domain error";
  case GSL_ERANGE: return "output range error";
  case GSL_EINVAL: return "invalid argument supplied by user";
  case GSL_EBADFUNC: return "problem with user-supplied function";
  case GSL_ERUNAWAY: return "iterative process is out of control";
  case GSL_EZERODIV: return "tried to divide by zero";
  default: return "unknown error code";
  }
}
/* real GSL calls the installed handler (default: abort); the reference pre-checks the
 * inputs so these sites are not reached on valid data -- report and return the code. */
#define GSL_ERROR(msg, code) \
  do { fprintf (stderr, "gsl-shim: %s:%d: %s (%s)\n", __FILE__, __LINE__, msg, gsl_strerror (code)); return code; } while (0)
#endif

/* gsl_function / gsl_function_fdf stand-ins (test infrastructure, see gsl_errno.h). */
#ifndef QB_GSL_MATH_H
#define QB_GSL_MATH_H
#include <math.h>
#include <float.h>
#define GSL_DBL_EPSILON 2.2204460492503131e-16
typedef struct { double (*function) (double x, void* params); void* params; } gsl_function;
#define GSL_FN_EVAL(F,x) (*((F)->function))(x,(F)->params)
typedef struct {
  double (*f) (double x, void* params);
  double (*df) (double x, void* params);
  void (*fdf) (double x, void* params, double* f, double* df);
  void* params;
} gsl_function_fdf;
#define GSL_FN_FDF_EVAL_F(FDF,x) (*((FDF)->f))(x,(FDF)->params)
#define GSL_FN_FDF_EVAL_DF(FDF,x) (*((FDF)->df))(x,(FDF)->params)
#define GSL_FN_FDF_EVAL_F_DF(FDF,x,y,dy) (*((FDF)->fdf))(x,(FDF)->params,(y),(dy))
#endif

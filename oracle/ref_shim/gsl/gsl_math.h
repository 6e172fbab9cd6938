#ifndef CMCB200_SHIM_GSL_MATH_H
#define CMCB200_SHIM_GSL_MATH_H
#include <math.h>
#include <float.h>
typedef struct { double (*function)(double x, void *params); void *params; } gsl_function;
#define GSL_FN_EVAL(F, x) (*((F)->function))(x, (F)->params)
#define GSL_SIGN(x) ((x) >= 0.0 ? 1 : -1)
#define GSL_MAX(a, b) ((a) > (b) ? (a) : (b))
#define GSL_MIN(a, b) ((a) < (b) ? (a) : (b))
#define GSL_DBL_EPSILON 2.2204460492503131e-16
#endif

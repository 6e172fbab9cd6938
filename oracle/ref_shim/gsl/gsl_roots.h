#ifndef CMCB200_SHIM_GSL_ROOTS_H
#define CMCB200_SHIM_GSL_ROOTS_H
#include <stddef.h>
#include "gsl_math.h"
typedef struct { const char *name; size_t size; } gsl_root_fsolver_type;
typedef struct { const gsl_root_fsolver_type *type; gsl_function *function; double root, x_lower, x_upper; void *state; } gsl_root_fsolver;
extern const gsl_root_fsolver_type *gsl_root_fsolver_brent;
extern const gsl_root_fsolver_type *gsl_root_fsolver_bisection;
gsl_root_fsolver *gsl_root_fsolver_alloc(const gsl_root_fsolver_type *T);
void gsl_root_fsolver_free(gsl_root_fsolver *s);
int gsl_root_fsolver_set(gsl_root_fsolver *s, gsl_function *f, double x_lower, double x_upper);
int gsl_root_fsolver_iterate(gsl_root_fsolver *s);
double gsl_root_fsolver_root(const gsl_root_fsolver *s);
double gsl_root_fsolver_x_lower(const gsl_root_fsolver *s);
double gsl_root_fsolver_x_upper(const gsl_root_fsolver *s);
int gsl_root_test_interval(double x_lower, double x_upper, double epsabs, double epsrel);
int gsl_root_test_delta(double x1, double x0, double epsabs, double epsrel);
#endif

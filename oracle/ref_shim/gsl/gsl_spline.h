#ifndef CMCB200_SHIM_GSL_SPLINE_H
#define CMCB200_SHIM_GSL_SPLINE_H
#include <stddef.h>
#include "gsl_math.h"
typedef struct { size_t cache; } gsl_interp_accel;
typedef struct { const char *name; } gsl_interp_type;
typedef struct { const gsl_interp_type *type; double *x, *y; size_t size; } gsl_spline;
extern const gsl_interp_type *gsl_interp_linear;
extern const gsl_interp_type *gsl_interp_cspline;
gsl_interp_accel *gsl_interp_accel_alloc(void);
void gsl_interp_accel_free(gsl_interp_accel *a);
gsl_spline *gsl_spline_alloc(const gsl_interp_type *T, size_t size);
int gsl_spline_init(gsl_spline *s, const double *x, const double *y, size_t size);
double gsl_spline_eval(const gsl_spline *s, double x, gsl_interp_accel *a);
void gsl_spline_free(gsl_spline *s);
#endif

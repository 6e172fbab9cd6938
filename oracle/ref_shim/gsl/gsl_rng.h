#ifndef CMCB200_SHIM_GSL_RNG_H
#define CMCB200_SHIM_GSL_RNG_H
#include <stdio.h>
typedef struct { const char *name; unsigned long max, min; size_t size; } gsl_rng_type;
typedef struct { const gsl_rng_type *type; void *state; } gsl_rng;
extern const gsl_rng_type *gsl_rng_mt19937;
extern const gsl_rng_type *gsl_rng_taus2;
gsl_rng *gsl_rng_alloc(const gsl_rng_type *T);
void gsl_rng_free(gsl_rng *r);
void gsl_rng_set(const gsl_rng *r, unsigned long seed);
double gsl_rng_uniform(const gsl_rng *r);
unsigned long gsl_rng_get(const gsl_rng *r);
unsigned long gsl_rng_uniform_int(const gsl_rng *r, unsigned long n);
int gsl_rng_fwrite(FILE *f, const gsl_rng *r);
int gsl_rng_fread(FILE *f, gsl_rng *r);
#endif

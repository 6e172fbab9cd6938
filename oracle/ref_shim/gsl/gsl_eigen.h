#ifndef CMCB200_SHIM_GSL_EIGEN_H
#define CMCB200_SHIM_GSL_EIGEN_H
#include "gsl_math.h"
/* TEST INFRASTRUCTURE ONLY: typedef-level stand-ins so that src/cmc/cmc_io.c compiles (its tidal-tensor reader uses them; that function is
 * never called by the harness and is dropped by --gc-sections) */
typedef struct { int n; } gsl_eigen_symm_workspace;
typedef struct { int n; double *data; } gsl_vector;
typedef struct { int n1, n2; double *data; } gsl_matrix;
typedef struct { gsl_matrix matrix; } gsl_matrix_view;
gsl_matrix_view gsl_matrix_view_array(double *base, size_t n1, size_t n2);
#endif

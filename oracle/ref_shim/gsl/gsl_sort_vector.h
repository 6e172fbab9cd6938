#ifndef CMCB200_SHIM_GSL_SORT_VECTOR_H
#define CMCB200_SHIM_GSL_SORT_VECTOR_H
#include "gsl_math.h"
#endif

#ifndef CMCB200_SHIM_GSL_EIGEN_H
#define CMCB200_SHIM_GSL_EIGEN_H
#include "gsl_math.h"
#endif

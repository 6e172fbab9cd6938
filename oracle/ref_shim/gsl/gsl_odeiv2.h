#ifndef CMCB200_SHIM_GSL_ODEIV2_H
#define CMCB200_SHIM_GSL_ODEIV2_H
#include "gsl_math.h"
#endif

#ifndef CMCB200_SHIM_GSL_INTEGRATION_H
#define CMCB200_SHIM_GSL_INTEGRATION_H
#include "gsl_math.h"
#endif

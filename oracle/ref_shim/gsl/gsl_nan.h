#ifndef CMCB200_SHIM_GSL_NAN_H
#define CMCB200_SHIM_GSL_NAN_H
#include <math.h>
#define GSL_POSINF INFINITY
#define GSL_NEGINF (-INFINITY)
#define GSL_NAN NAN
#endif

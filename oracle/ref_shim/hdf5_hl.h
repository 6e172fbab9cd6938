/* TEST INFRASTRUCTURE ONLY: stand-in for <hdf5_hl.h>. */
#ifndef CMCB200_SHIM_HDF5_HL_H
#define CMCB200_SHIM_HDF5_HL_H
#include "hdf5.h"
#endif

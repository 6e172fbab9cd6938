/* TEST INFRASTRUCTURE ONLY: typedef-level stand-in for <hdf5.h>; no HDF5 I/O is on the hot path. */
#ifndef CMCB200_SHIM_HDF5_H
#define CMCB200_SHIM_HDF5_H
#include <stddef.h>
typedef long hid_t;
typedef int herr_t;
typedef unsigned long long hsize_t;
typedef int hbool_t;
#define H5T_NATIVE_DOUBLE 1
#define H5T_NATIVE_LONG 2
#define H5T_NATIVE_INT 3
#define H5T_C_S1 4
#define H5F_ACC_TRUNC 0
#define H5F_ACC_RDWR 1
#define H5P_DEFAULT 0
#endif

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
/* what write_snapshot (src/cmc/cmc_io.c:2960) needs to compile: the H5 / H5TB calls themselves are recorders in ref_harness.c */
typedef int htri_t;
typedef int H5Z_filter_t;
#define H5Z_FILTER_DEFLATE 1
#define H5Z_FILTER_CONFIG_ENCODE_ENABLED 1u
#define H5Z_FILTER_CONFIG_DECODE_ENABLED 2u
#define H5F_ACC_EXCL 4
#define HOFFSET(S, M) (offsetof(S, M))
#define H5E_BEGIN_TRY {
#define H5E_END_TRY }
#endif

#ifndef CMCB200_SHIM_FITSIO_H
#define CMCB200_SHIM_FITSIO_H
typedef struct { int dummy; } fitsfile;
#endif

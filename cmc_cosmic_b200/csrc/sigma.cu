#include "ctx.cuh"
void sigma_free(cmcb200_ctx *c) { (void) c; }

// Shared device helpers for the cmcb200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "fp64_exact.h"

#define CMCB_CUDA_OK(call)                                                                              \
	do {                                                                                             \
		cudaError_t _e = (call);                                                                 \
		if (_e != cudaSuccess) {                                                                 \
			cmcb::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e)); \
			return -1;                                                                       \
		}                                                                                        \
	} while (0)

namespace cmcb {

void set_error(const char *fmt, ...);
const char *get_error();

constexpr int NUM_SMS_B200 = 148;

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ unsigned warp_id() { return threadIdx.x >> 5; }

// acquire/release accesses for the decoupled look-back descriptors
__device__ __forceinline__ int ld_acquire(const int *p)
{
	int v;
	asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ void st_release(int *p, int v)
{
	asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned *p)
{
	unsigned v;
	asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ void st_release_u32(unsigned *p, unsigned v)
{
	asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ double ld_relaxed_f64(const double *p)
{
	double v;
	asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
	return v;
}

} // namespace cmcb

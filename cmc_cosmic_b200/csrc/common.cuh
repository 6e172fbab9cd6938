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
// relaxed device-scope accesses (L2, no L1 invalidate / fence): enough when the polled word carries the data itself
__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned *p)
{
	unsigned v;
	asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ void st_relaxed_u32(unsigned *p, unsigned v)
{
	asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ double ld_relaxed_f64(const double *p)
{
	double v;
	asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
	return v;
}

// ---- 1-D bulk copy global -> shared through the TMA unit (cp.async.bulk), completion counted on an mbarrier ------------------------
// One thread issues the copy; every thread of the CTA then waits on the barrier's phase.  bytes: a multiple of 16, both addresses 16-byte
// aligned.  Used to stage the coarsest ladder level of the orbit kernels (<= 40 KB) once per persistent CTA.
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); // make the init visible to the async proxy
}
__device__ __forceinline__ void bulk_copy_g2s(void *smem_dst, const void *gsrc, unsigned bytes, unsigned long long *bar)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)), "l"(gsrc),
		     "r"(bytes), "r"(smem_u32(bar))
		     : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
	unsigned done = 0, spins = 0;
	while (!done) {
		asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
			     : "=r"(done)
			     : "r"(smem_u32(bar)), "r"(parity)
			     : "memory");
		if (!done && ++spins > (1u << 26)) __trap(); // a copy that never lands must not hang the GPU
	}
}

} // namespace cmcb

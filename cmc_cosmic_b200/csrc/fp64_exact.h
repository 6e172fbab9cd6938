// IEEE-754 binary64 operations with exactly one rounding each, usable from host and device code.
//
// The reference is built for baseline x86-64 (SSE2, no FMA contraction), so every `double` operation in
// its hot path rounds once.  nvcc contracts a*b+c into FMA by default; these wrappers pin the rounding on
// the device (and are plain operators on the host, where the translation units are built with
// -ffp-contract=off).  All arithmetic on the parity-critical path goes through them.
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define CMCB_HD __host__ __device__ __forceinline__
#else
#define CMCB_HD inline
#endif

namespace cmcb {

CMCB_HD double dadd(double a, double b)
{
#if defined(__CUDA_ARCH__)
	return __dadd_rn(a, b);
#else
	return a + b;
#endif
}
CMCB_HD double dsub(double a, double b)
{
#if defined(__CUDA_ARCH__)
	return __dsub_rn(a, b);
#else
	return a - b;
#endif
}
CMCB_HD double dmul(double a, double b)
{
#if defined(__CUDA_ARCH__)
	return __dmul_rn(a, b);
#else
	return a * b;
#endif
}
CMCB_HD double ddiv(double a, double b)
{
#if defined(__CUDA_ARCH__)
	return __ddiv_rn(a, b);
#else
	return a / b;
#endif
}
CMCB_HD double dsqrt(double a)
{
#if defined(__CUDA_ARCH__)
	return __dsqrt_rn(a);
#else
	return __builtin_sqrt(a);
#endif
}
CMCB_HD double dfma(double a, double b, double c)
{
#if defined(__CUDA_ARCH__)
	return __fma_rn(a, b, c);
#else
	return __builtin_fma(a, b, c);
#endif
}
CMCB_HD uint64_t dbits(double a)
{
#if defined(__CUDA_ARCH__)
	return (uint64_t) __double_as_longlong(a);
#else
	uint64_t u;
	memcpy(&u, &a, 8);
	return u;
#endif
}
CMCB_HD double dfrombits(uint64_t u)
{
#if defined(__CUDA_ARCH__)
	return __longlong_as_double((long long) u);
#else
	double a;
	memcpy(&a, &u, 8);
	return a;
#endif
}
CMCB_HD double dabs(double a) { return dfrombits(dbits(a) & 0x7fffffffffffffffull); }
CMCB_HD bool dfinite(double a) { return ((dbits(a) >> 52) & 0x7ff) != 0x7ff; }
CMCB_HD double dmax(double a, double b) { return a >= b ? a : b; }
CMCB_HD double dmin(double a, double b) { return a <= b ? a : b; }

} // namespace cmcb

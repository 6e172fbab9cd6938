// Counter-based uniform draws for the kernels that sample (orbit.cu: new radii; dynamics.cu: encounter angles).
#pragma once

namespace cmcb {

// Philox4x32-10 (Salmon et al., SC'11), counter = (star index lo, hi, tcount lo, block of four draws), key = seed.
struct PhiloxDraws {
	unsigned key0, key1, c0, c1, c2, c3;
	unsigned o0, o1, o2, o3;
	int have;
	__device__ void init(unsigned long long seed, unsigned long long tcount, unsigned long long star)
	{
		key0 = (unsigned) seed; key1 = (unsigned) (seed >> 32);
		c0 = (unsigned) star; c1 = (unsigned) (star >> 32); c2 = (unsigned) tcount; c3 = 0;
		have = 0;
	}
	__device__ void refill()
	{
		unsigned k0 = key0, k1 = key1, x0 = c0, x1 = c1, x2 = c2, x3 = c3;
#pragma unroll
		for (int r = 0; r < 10; r++) {
			unsigned hi0 = __umulhi(0xD2511F53u, x0), lo0 = 0xD2511F53u * x0;
			unsigned hi1 = __umulhi(0xCD9E8D57u, x2), lo1 = 0xCD9E8D57u * x2;
			unsigned y0 = hi1 ^ x1 ^ k0, y1 = lo1, y2 = hi0 ^ x3 ^ k1, y3 = lo0;
			x0 = y0; x1 = y1; x2 = y2; x3 = y3;
			k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
		}
		o0 = x0; o1 = x1; o2 = x2; o3 = x3;
		c3++;
		have = 4;
	}
	__device__ bool has(int) const { return true; }
	__device__ double next()
	{
		if (!have) refill();
		unsigned v = have == 4 ? o0 : (have == 3 ? o1 : (have == 2 ? o2 : o3));
		have--;
		return (double) v / 4294967296.0; // same 2^-32 grid as taus113's rng_t113_dbl_new (taus113-v2.c:191-193)
	}
};

} // namespace cmcb

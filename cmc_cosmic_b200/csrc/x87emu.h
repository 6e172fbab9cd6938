// Software model of the x87 80-bit extended format (64-bit significand, round-to-nearest-even), host + device.
//
// The reference evaluates Q(r) = 2(E - (Phi + phi_s)) - (J/r)^2 in `long double`
// (function_q, reference src/cmc/cmc_utils.c:1654-1662; caller find_zero_Q, src/cmc/cmc_nr.c:316-383), which on
// x86-64/GCC is this format.  The index bisections only consume the SIGN of the result (and one qmin<qmax
// comparison), so the CUDA path evaluates Q in FP64 with a rigorous error bound (orbit_core.h) and comes here
// only when |Q| is inside that bound -- then the x87 arithmetic is replayed exactly, operation by operation,
// so kmin/kmax stay bit-identical to the reference.  No GPU has an 80-bit type; this is integer arithmetic.
#pragma once
#include "fp64_exact.h"

namespace cmcb {

// value = (-1)^s * m * 2^(e - 63); m has bit 63 set unless the value is zero (m == 0)
struct X87 {
	uint64_t m;
	int32_t e;
	uint32_t s;
};

CMCB_HD int x87_clz64(uint64_t x)
{
#if defined(__CUDA_ARCH__)
	return __clzll((long long) x);
#else
	return x ? __builtin_clzll(x) : 64;
#endif
}

CMCB_HD void x87_mul64(uint64_t a, uint64_t b, uint64_t &hi, uint64_t &lo)
{
#if defined(__CUDA_ARCH__)
	hi = __umul64hi(a, b);
	lo = a * b;
#else
	unsigned __int128 p = (unsigned __int128) a * b;
	hi = (uint64_t) (p >> 64);
	lo = (uint64_t) p;
#endif
}

// finite doubles only (the callers never pass inf/nan: E, J, r, Phi, m of live stars)
CMCB_HD X87 x87_from_double(double d)
{
	uint64_t b = dbits(d);
	X87 x;
	x.s = (uint32_t) (b >> 63);
	int ef = (int) ((b >> 52) & 0x7ff);
	uint64_t frac = b & 0xfffffffffffffull;
	if (ef == 0) {
		if (frac == 0) { x.m = 0; x.e = 0; return x; }
		int z = x87_clz64(frac);
		x.m = frac << z;
		x.e = -1022 - 52 + (63 - z);
		return x;
	}
	x.m = (frac | 0x10000000000000ull) << 11;
	x.e = ef - 1023;
	return x;
}

// round a 128-bit significand hi:lo (hi normalised, bit 63 set; `sticky` = bits already shifted out) to 64 bits, RNE
CMCB_HD X87 x87_round(uint32_t s, int32_t e, uint64_t hi, uint64_t lo, int sticky)
{
	X87 x;
	x.s = s;
	const uint64_t HALF = 0x8000000000000000ull;
	bool up = (lo > HALF) || (lo == HALF && (sticky || (hi & 1)));
	if (up) {
		hi++;
		if (hi == 0) { hi = HALF; e++; }
	}
	x.m = hi;
	x.e = e;
	return x;
}

CMCB_HD X87 x87_neg(X87 a) { a.s ^= 1; return a; }

CMCB_HD X87 x87_add(X87 a, X87 b)
{
	if (a.m == 0) { if (b.m == 0) { b.s = a.s & b.s; } return b; }
	if (b.m == 0) return a;
	// make |a| >= |b|
	if (b.e > a.e || (b.e == a.e && b.m > a.m)) { X87 t = a; a = b; b = t; }
	int sh = a.e - b.e;
	uint64_t bhi, blo;
	int sticky = 0;
	if (sh == 0) { bhi = b.m; blo = 0; }
	else if (sh < 64) { bhi = b.m >> sh; blo = b.m << (64 - sh); }
	else if (sh == 64) { bhi = 0; blo = b.m; }
	else if (sh < 128) { bhi = 0; blo = b.m >> (sh - 64); sticky = (b.m << (128 - sh)) != 0; }
	else { bhi = 0; blo = 0; sticky = 1; }
	uint64_t hi, lo;
	int32_t e = a.e;
	if (a.s == b.s) {
		lo = blo; // a's low word is 0
		hi = a.m + bhi;
		if (hi < a.m) { // carry out of bit 63
			sticky |= (int) (lo & 1);
			lo = (lo >> 1) | (hi << 63);
			hi = (hi >> 1) | 0x8000000000000000ull;
			e++;
		}
		return x87_round(a.s, e, hi, lo, sticky);
	}
	// subtraction of magnitudes: (a.m:0) - (bhi:blo) - (sticky ? tiny : 0)
	lo = 0 - blo;
	uint64_t borrow = blo != 0;
	hi = a.m - bhi - borrow;
	if (sticky) { // the discarded tail of b makes the true difference slightly smaller: borrow one ulp of lo, keep sticky
		if (lo == 0) hi--;
		lo--;
	}
	if (hi == 0 && lo == 0) { X87 z; z.m = 0; z.e = 0; z.s = 0; return z; } // exact cancellation -> +0 under RNE
	int z = hi ? x87_clz64(hi) : 64 + x87_clz64(lo);
	if (z) { // normalise (z >= 2 implies sh <= 1, so nothing was lost and sticky == 0)
		if (z < 64) { hi = (hi << z) | (lo >> (64 - z)); lo <<= z; }
		else { hi = lo << (z - 64); lo = 0; }
		e -= z;
	}
	return x87_round(a.s, e, hi, lo, sticky);
}

CMCB_HD X87 x87_sub(X87 a, X87 b) { return x87_add(a, x87_neg(b)); }

CMCB_HD X87 x87_mul(X87 a, X87 b)
{
	X87 x;
	x.s = a.s ^ b.s;
	if (a.m == 0 || b.m == 0) { x.m = 0; x.e = 0; return x; }
	uint64_t hi, lo;
	x87_mul64(a.m, b.m, hi, lo);
	int32_t e = a.e + b.e + 1;
	if (!(hi >> 63)) { hi = (hi << 1) | (lo >> 63); lo <<= 1; e--; }
	return x87_round(x.s, e, hi, lo, 0);
}

CMCB_HD X87 x87_mul2(X87 a) { if (a.m) a.e++; return a; }

// 128-by-64 division for a normalised divisor d (bit 63 set) and n_hi < d: classic shift-subtract.
// Slow (64 steps) on purpose: it only runs on the marginal-sign path.
CMCB_HD void x87_div128(uint64_t n_hi, uint64_t n_lo, uint64_t d, uint64_t &q, uint64_t &rem)
{
#if !defined(__CUDA_ARCH__)
	unsigned __int128 n = ((unsigned __int128) n_hi << 64) | n_lo;
	q = (uint64_t) (n / d);
	rem = (uint64_t) (n % d);
#else
	uint64_t r = n_hi, quo = 0;
	for (int i = 63; i >= 0; i--) {
		uint64_t top = r >> 63;
		r = (r << 1) | ((n_lo >> i) & 1);
		quo <<= 1;
		if (top || r >= d) { r -= d; quo |= 1; }
	}
	q = quo;
	rem = r;
#endif
}

CMCB_HD X87 x87_div(X87 a, X87 b)
{
	X87 x;
	x.s = a.s ^ b.s;
	if (a.m == 0) { x.m = 0; x.e = 0; return x; } // b != 0 by contract
	uint64_t q, rem;
	int32_t e;
	if (a.m < b.m) { x87_div128(a.m, 0, b.m, q, rem); e = a.e - b.e - 1; }       // a.m*2^64 / b.m in [2^63, 2^64)
	else { x87_div128(a.m >> 1, a.m << 63, b.m, q, rem); e = a.e - b.e; }          // a.m*2^63 / b.m in [2^63, 2^64)
	// RNE from the remainder: compare 2*rem with b.m without overflow
	bool up = rem > b.m - rem || (rem == b.m - rem && (q & 1));
	if (up) { q++; if (q == 0) { q = 0x8000000000000000ull; e++; } }
	x.m = q;
	x.e = e;
	return x;
}

// (double) conversion, RNE to 53 bits.  Normal range only is exact; results below 2^-1022 go through a scaled
// multiply (may double-round there; |Q| never gets near it).
CMCB_HD double x87_to_double(X87 a)
{
	if (a.m == 0) return a.s ? -0.0 : 0.0;
	uint64_t keep = a.m >> 11, rest = a.m & 0x7ff;
	int32_t e = a.e;
	if (rest > 0x400 || (rest == 0x400 && (keep & 1))) { keep++; if (keep >> 53) { keep >>= 1; e++; } }
	if (e > 1023) return a.s ? -__builtin_huge_val() : __builtin_huge_val();
	if (e >= -1022) return dfrombits(((uint64_t) a.s << 63) | ((uint64_t) (e + 1023) << 52) | (keep & 0xfffffffffffffull));
	double v = (double) keep; // exact, 53 bits
	int sh = e - 52;
	while (sh < -1000) { v *= 0x1p-1000; sh += 1000; }
	v *= dfrombits((uint64_t) (sh + 1023) << 52);
	return a.s ? -v : v;
}

CMCB_HD int x87_sign(X87 a) { return a.m == 0 ? 0 : (a.s ? -1 : 1); }

// function_q (reference cmc_utils.c:1654-1662) replayed operation by operation.
//   r, pot : radius / potential at the probe (doubles widened exactly);  E, J : the star's energy (incl. self term) and J
//   m_j, r_j : star_m[j], star_r[j] of the star itself;  invN : 1.0/clus.N_STAR evaluated in double
// Returns the double the reference's function_q returns.
CMCB_HD double x87_function_q(double r, double pot, double E, double J, double m_j, double r_j, double invN)
{
	X87 rx = x87_from_double(r);
	// Jr = SQR(J/r): the quotient is rounded to 64 bits, then to double (SQR takes a double), squared in double
	double jr = x87_to_double(x87_div(x87_from_double(J), rx));
	X87 Jr = x87_from_double(dmul(jr, jr));
	// phis = MPI_PHI_S(r, j): (r >= star_r[j] ? star_m[j]/r : star_m[j]/star_r[j]) * (1.0/N_STAR)
	X87 ratio = (r >= r_j) ? x87_div(x87_from_double(m_j), rx) : x87_from_double(ddiv(m_j, r_j));
	X87 phis = x87_mul(ratio, x87_from_double(invN));
	// res = 2.0 * (E - (pot + phis)) - Jr, rounded to double on return
	X87 inner = x87_add(x87_from_double(pot), phis);
	X87 diff = x87_sub(x87_from_double(E), inner);
	return x87_to_double(x87_sub(x87_mul2(diff), Jr));
}

} // namespace cmcb

// Bit-exact parallel evaluation of a sequentially rounded FP64 recurrence
//        y_i = fl(y_{i-1} + s*a_i),   s = +1 or -1,  a_i >= 0,  y > 0
// which is what the reference's potential rebuild is made of (reference src/cmc/cmc_utils.c:685-693 the forward
// mass sum, :717-720 the backward `mprev -= star_m[k]/N_STAR` and `phi[k] = phi[k+1] - ...` chains).
//
// Why: the reference evaluates these chains strictly in order, and for equal-mass clusters the rounding error of
// `mprev -= const` has constant sign per binade and accumulates LINEARLY (measured with the oracle: 2e-11 at
// N=1e5 ... 3e-9 at N=1e7 relative to an exactly rounded sum).  A conventional parallel scan is more accurate than
// the reference and therefore misses the 1e-12 parity bound.  Instead of matching approximately we reproduce the
// sequential roundings exactly:
//
//   While y_{i-1} and y_i lie in the same binade [2^e, 2^(e+1)) both are multiples of u = 2^(e-52), and
//   fl(y_{i-1} +- a_i) = y_{i-1} +- R_e(a_i) where R_e rounds a_i to the nearest multiple of u -- unless a_i lies
//   exactly half-way (tie -> the chain's round-half-even depends on y).  R_e(a_i) does not depend on y, so inside
//   such a run the chain is a sum of multiples of u that stays below 2^(e+1): every partial sum is exactly
//   representable and the summation ORDER no longer matters -- a parallel scan gives the sequential bits.
//
//   The few elements where that argument does not apply ("events": the binade of y changes, it cannot be decided
//   from an approximate prefix with a rigorous error bound, a tie, a negative/non-finite term, y <= 0) are
//   resolved one after the other with the hardware's own rounding by a single thread, using the exact run sums
//   in between.  There are O(log2(range of y)) + O(n * 2^-20) of them.
//
// This header holds the per-element classification shared by the CUDA kernels (potential.cu) and the host-side
// emulation used by the CPU tests (tests/host_emul/).
#pragma once
#include "fp64_exact.h"

namespace cmcb {

struct SeqTerm {
	double v;   // R_e(a) for a regular element, 0 for an event
	int event;  // 1: must be applied sequentially with the hardware rounding
};

// a        : the term (its sign is applied by the chain direction, a itself must be >= 0 to be regular)
// y_before : approximate y_{i-1};  y_after : approximate y_i  (any parallel prefix, each within `i` roundings of exact)
// y0_abs   : |y_0|;   i : 1-based position in the chain (bounds the accumulated rounding of chain and prefix)
CMCB_HD SeqTerm seq_classify(double a, double y_before, double y_after, double y0_abs, long i)
{
	SeqTerm t;
	t.v = 0.0;
	t.event = 1;
	// fl(y + 0) = y for every y: a zero term is regular whatever the chain's value is (a chain of zeros -- Eint of a cluster without stellar
	// evolution, the binding energies of a cluster without binaries -- has no events at all)
	if (a == 0.0) { t.event = 0; return t; }
	// rigorous bound on |approximate - true sequential value|: the chain commits <= i roundings of <= 2^-53 * max|y|,
	// the parallel prefix at most as many; (i + 64) * 2^-51 leaves a factor 2 of slack.
	double S = dmax(y0_abs, dmax(dabs(y_before), dabs(y_after)));
	double delta = dmul(dmul((double) (i + 64), 4.440892098500626e-16 /* 2^-51 */), S);
	double lo = dsub(dmin(y_before, y_after), delta);
	double hi = dadd(dmax(y_before, y_after), delta);
	if (!(a >= 0.0) || !dfinite(a) || !(lo > 0.0) || !dfinite(hi)) return t;
	uint64_t elo = dbits(lo) >> 52, ehi = dbits(hi) >> 52; // sign bit is 0 here
	if (elo != ehi || elo < 64) return t;                   // binade undecided or changing (or too close to subnormals)
	double C = dfrombits(elo << 52);                        // 2^e
	if (a > C) return t;                                    // cannot happen inside one binade; keep the add below exact
	double R = dsub(dadd(a, C), C);                         // a rounded to a multiple of u = 2^(e-52)
	double rho = dsub(a, R);                                // exact (error term of the addition)
	double half_u = dfrombits((elo - 53) << 52);
	if (dabs(rho) == half_u) return t;                      // exact tie: round-half-even depends on y
	t.v = R;
	t.event = 0;
	return t;
}

// The same for a chain whose terms carry their own sign, y_i = fl(y_{i-1} + a_i) with a_i of either sign and y > 0 -- a sliding-window sum
// kept up to date by subtracting the element that leaves and adding the one that enters (calc_sigma_r, reference
// src/cmc/cmc_dynamics_helper.c:2935-2968).  fl(y - a) = y - R_e(a) by the same argument.  Such a chain is not monotone: it crosses a
// binade boundary every time its noise carries it over one (every crossing is an event), and the accumulated rounding is bounded by the
// largest value the chain has EVER had, which the neighbouring values do not reveal: the caller supplies y_max >= |y_j| for all j.
CMCB_HD SeqTerm seq_classify_signed(double a, double y_before, double y_after, double y_max, long i)
{
	SeqTerm t;
	t.v = 0.0;
	t.event = 1;
	if (a == 0.0) { t.event = 0; return t; } // fl(y + 0) = y
	double S = dmax(y_max, dmax(dabs(y_before), dabs(y_after)));
	double delta = dmul(dmul((double) (i + 64), 4.440892098500626e-16 /* 2^-51 */), S);
	double lo = dsub(dmin(y_before, y_after), delta);
	double hi = dadd(dmax(y_before, y_after), delta);
	double aa = dabs(a);
	if (!dfinite(a) || !(lo > 0.0) || !dfinite(hi)) return t;
	uint64_t elo = dbits(lo) >> 52, ehi = dbits(hi) >> 52;
	if (elo != ehi || elo < 64) return t;
	double C = dfrombits(elo << 52);
	if (aa > C) return t;
	double R = dsub(dadd(aa, C), C);
	double rho = dsub(aa, R);
	double half_u = dfrombits((elo - 53) << 52);
	if (dabs(rho) == half_u) return t;
	t.v = a < 0.0 ? -R : R;
	t.event = 0;
	return t;
}

} // namespace cmcb

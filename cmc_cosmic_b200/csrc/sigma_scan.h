// calc_sigma_r's two running window sums (reference src/cmc/cmc_dynamics_helper.c:2821-2975) evaluated in parallel with the reference's
// own bits.  Shared by sigma.cu and its host build (tests/host_emul/sigma_host.cpp).
//
// The reference keeps  y <- fl(fl(y - a_leave) + a_enter)  per star: a chain of 2N separately rounded operations whose value hovers around
// 2p terms, so it crosses a binade boundary on ~5 % of its operations -- far too often for the event list of exactsum.h / seqscan.cuh, and
// its rounding drift (up to N 2^-53 relative against a freshly summed window) rules out anything but an exact replay.
//
// What makes a replay parallel: a rounding to nearest-even onto the grid u commutes with a translation by a multiple of 2u.  So if along a
// stretch of operations the binade of every intermediate value is KNOWN (from an approximate value with a rigorous error band that does
// not straddle a power of two), the stretch's map F: y_in -> y_out satisfies F(y + m P) = F(y) + m P for P = 2 ulp(highest binade visited):
// F is fixed by its values on the K' = P / ulp(y_in) residues of y_in's mantissa, i.e. by a TABLE indexed with the low log2 K' mantissa
// bits of y_in.  Going down in binade costs nothing; every binade the stretch climbs above y_in's doubles the table.
//   segment  S = 8 stars (16 operations), one thread: approximate trajectory from a freshly summed window, classification, and the table
//            by running the <= 8 candidates through the operations with the hardware's own adds;
//   group    16 segments: a table of <= 16 entries by walking the candidates through the segment tables (one thread per candidate);
//   tile     16 groups, one CTA: a table of <= 128 entries by walking the candidates through the group tables;
//   chain    one thread walks the tile tables in order (one dependent lookup per 2048 stars) and leaves every tile's exact y_in;
//   finish   each tile walks its segment tables from the exact y_in, each thread replays its 16 operations and stores the outputs.
// Where the band straddles a power of two ("hard": ~1 operation in 10^8 at N = 10^7) the segment is replayed operation by operation by
// whoever walks over it, and the tile has no table: the chain walker goes through its segments.  A stretch that climbs more binades than
// the table sizes allow is replayed the same way.
//
// The band: the reference's value differs from the exactly summed window by its accumulated rounding, <= sum over all earlier operations
// of 2^-53 |y|.  Every term sits in at most 2p windows, so the values at the stars so far sum to <= 2p x (the prefix sum of the terms up
// to the window's end), the values after the subtractions are smaller still, and the 2p additions that build the first window add at most
// 2p x that prefix: sum |y| <= 6p x prefix.  dband = 2^-52 x 12p x prefix carries a factor 4 on top, which also pays for using the
// approximate values in the induction (valid while the true value stays within a factor 2 of the approximate one; delta <= y~/4 keeps it
// there -- otherwise: `broken`, the whole chain is replayed by one thread).
#pragma once
#include "fp64_exact.h"
#include "sigma_core.h"

namespace cmcb {

constexpr int SIG_S = 8;
constexpr int SIG_THREADS = 256;
constexpr int SIG_TILE = SIG_S * SIG_THREADS;
constexpr int SIG_G = 16;                      // segments per group
constexpr int SIG_NGRP = SIG_THREADS / SIG_G;  // groups per tile
constexpr int SIG_KS_LOG_MAX = 3;
constexpr int SIG_KG_LOG_MAX = 4;
constexpr int SIG_KT_LOG_MAX = 7;
constexpr int SIG_KS_MAX = 1 << SIG_KS_LOG_MAX;
constexpr int SIG_KT_MAX = 1 << SIG_KT_LOG_MAX;
constexpr int SIG_KG_MAX = 1 << SIG_KG_LOG_MAX;
static_assert(SIG_NGRP * SIG_KG_MAX <= SIG_THREADS, "one thread per (group, candidate)");

enum { SIG_TABLE = 0, SIG_SEQ = 1, SIG_HARD = 2 };

struct SigSeg {
	double base;              // candidate 0: the approximate y_in with its low logk mantissa bits cleared
	double yout[SIG_KS_MAX];
	int e_in, logk, flag, emax;
};

// a group of SIG_G consecutive segments: the same kind of table one level up, so that a walk over a tile is SIG_NGRP + SIG_G dependent
// lookups instead of SIG_THREADS.  SIG_SEQ: the group climbs too many binades for its table -- it is walked segment by segment.
struct SigGrp {
	double base;
	double yout[SIG_KG_MAX];
	int e_in, logk, flag, emax;
};

struct SigTile { // 16-byte multiple: batches are staged with 128-bit copies
	double base;
	double yout[SIG_KT_MAX];
	int e_in, logk, flag, pad;
	double pad2;
};
static_assert(sizeof(SigTile) % 16 == 0, "SigTile is copied in 16-byte words");

// terms of one chain as the loop sees them: local index k = 1 - p .. n + p at [k + p - 1]
struct SigTerms {
	const double *t;
	long p;
	CMCB_HD double at(long k) const { return t[k + p - 1]; }
};

CMCB_HD int sig_expo(double y) { return (int) (dbits(y) >> 52); } // sign included: a negative value never matches a positive one's

// the operations of stars si0 .. si0 + cnt - 1 (si0 >= 2): sub[i] / add[i] = the terms leaving / entering at star si0 + i (0 where the
// window does not move, src/cmc/cmc_dynamics_helper.c:2923-2931)
CMCB_HD void sig_segment_terms(const SigmaWindow &w, const SigTerms &T, long si0, int cnt, double *sub, double *add)
{
	long prev = w.simin(si0 - 1);
	for (int i = 0; i < SIG_S; i++) {
		sub[i] = 0.0;
		add[i] = 0.0;
		if (i < cnt) {
			const long cur = w.simin(si0 + i);
			if (cur != prev) { sub[i] = T.at(prev); add[i] = T.at(cur + (2 * w.p - 1)); }
			prev = cur;
		}
	}
}

// the window of star si summed afresh in index order: the approximate value the bands are centred on
CMCB_HD double sig_fresh_window(const SigmaWindow &w, const SigTerms &T, long si)
{
	const long a = w.simin(si);
	double s = 0.0;
	for (long k = a; k <= a + 2 * w.p - 1; k++) s = dadd(s, T.at(k));
	return s;
}

CMCB_HD double sig_replay(const double *sub, const double *add, int cnt, double y)
{
	for (int i = 0; i < cnt; i++) { y = dsub(y, sub[i]); y = dadd(y, add[i]); }
	return y;
}

// One segment: classification and table.  y: approximate value entering the segment; dband: 2^-52 D (the drift bound, absolute);
// *broken is raised when the band is no longer small against the value.
CMCB_HD void sig_build_segment(const double *sub, const double *add, int cnt, double y, double dband, long p, SigSeg *o, int *broken)
{
	o->e_in = sig_expo(y);
	o->flag = SIG_TABLE;
	double t = y, ymax = y, ymin = y;
	for (int i = 0; i < cnt; i++) {
		const double t1 = dsub(t, sub[i]), t2 = dadd(t1, add[i]);
		ymax = dmax(ymax, dmax(t1, t2));
		ymin = dmin(ymin, dmin(t1, t2));
		t = t2;
	}
	// the local trajectory and the fresh sum commit <= 2p + 2S roundings, the candidates sit within 2^KT_LOG_MAX ulps of the approximate value
	// twice the drift bound: a candidate sits up to one band away from the true value, whose trajectory sits up to one band away from y~
	const double delta = dadd(dmul(2.0, dband), dmul(dmul((double) (2 * p + 2 * SIG_S + 2 * SIG_KT_MAX + 16), 2.220446049250313e-16 /* 2^-52 */), ymax));
	if (!(ymin > 0.0) || !dfinite(ymax) || !(delta <= dmul(0.25, ymin))) {
		o->flag = SIG_HARD;
		if (broken && !(delta <= dmul(0.25, ymin))) *broken = 1;
	}
	int emax = o->e_in;
	if (o->flag != SIG_HARD) {
		t = y;
		if (sig_expo(dsub(t, delta)) != o->e_in || sig_expo(dadd(t, delta)) != o->e_in || o->e_in < 64) o->flag = SIG_HARD;
		for (int i = 0; i < cnt; i++) {
			const double t1 = dsub(t, sub[i]), t2 = dadd(t1, add[i]);
			const int e1 = sig_expo(t1), e2 = sig_expo(t2);
			if (sig_expo(dsub(t1, delta)) != e1 || sig_expo(dadd(t1, delta)) != e1 || sig_expo(dsub(t2, delta)) != e2 || sig_expo(dadd(t2, delta)) != e2 ||
			    e1 < 64 || e2 < 64)
				o->flag = SIG_HARD;
			emax = e1 > emax ? e1 : emax;
			emax = e2 > emax ? e2 : emax;
			t = t2;
		}
	}
	o->emax = emax;
	o->logk = emax - o->e_in + 1;
	o->base = y;
	if (o->flag == SIG_HARD) return;
	if (o->logk > SIG_KS_LOG_MAX) { o->flag = SIG_SEQ; return; }
	const uint64_t K = 1ull << o->logk, b0 = dbits(y) & ~(K - 1);
	o->base = dfrombits(b0);
	// the candidates in three batches (0-1, 2-3, 4-7): most segments stay within one binade of y_in and need the first only
	double c[SIG_KS_MAX];
	for (int r = 0; r < SIG_KS_MAX; r++) c[r] = dfrombits(b0 | (uint64_t) (r & (int) (K - 1)));
	int mism = 0;
	t = y;
	for (int i = 0; i < cnt; i++) {
		const double t1 = dsub(t, sub[i]), t2 = dadd(t1, add[i]);
		const int e1 = sig_expo(t1), e2 = sig_expo(t2);
		for (int r = 0; r < 2; r++) {
			const double c1 = dsub(c[r], sub[i]);
			c[r] = dadd(c1, add[i]);
			mism |= (sig_expo(c1) ^ e1) | (sig_expo(c[r]) ^ e2);
		}
		if (K > 2) {
			for (int r = 2; r < 4; r++) {
				const double c1 = dsub(c[r], sub[i]);
				c[r] = dadd(c1, add[i]);
				mism |= (sig_expo(c1) ^ e1) | (sig_expo(c[r]) ^ e2);
			}
		}
		if (K > 4) {
			for (int r = 4; r < SIG_KS_MAX; r++) {
				const double c1 = dsub(c[r], sub[i]);
				c[r] = dadd(c1, add[i]);
				mism |= (sig_expo(c1) ^ e1) | (sig_expo(c[r]) ^ e2);
			}
		}
		t = t2;
	}
	for (int r = 2; r < SIG_KS_MAX; r++)
		if ((uint64_t) r >= K) c[r] = c[r & (int) (K - 1)];
	if (mism) { o->flag = SIG_HARD; return; } // cannot happen while the band is rigorous; the replay is always right
	for (int r = 0; r < SIG_KS_MAX; r++) o->yout[r] = c[r];
}

// the segment applied to the exact (or a candidate) value y
CMCB_HD double sig_apply_segment(const SigSeg &s, const double *sub, const double *add, int cnt, double y, int *inconsistent)
{
	if (s.flag == SIG_TABLE) {
		if (sig_expo(y) == s.e_in) {
			const uint64_t idx = dbits(y) & ((1ull << s.logk) - 1);
			const double yc = dfrombits(dbits(s.base) | idx);
			return dadd(s.yout[idx], dsub(y, yc));
		}
		if (inconsistent) *inconsistent = 1;
	}
	return sig_replay(sub, add, cnt, y);
}

CMCB_HD double sig_apply_tile_table(const SigTile &t, double y, int *inconsistent)
{
	if (sig_expo(y) != t.e_in) { if (inconsistent) *inconsistent = 1; }
	const uint64_t idx = dbits(y) & ((1ull << t.logk) - 1);
	const double yc = dfrombits(dbits(t.base) | idx);
	return dadd(t.yout[idx], dsub(y, yc));
}

// tile header from its segments (the entries are filled by the candidates' walks): HARD when a segment is, or when the tile climbs more
// binades above its y_in than the table holds
CMCB_HD void sig_tile_header(const SigSeg *seg, int nseg, SigTile *t)
{
	int emax = seg[0].e_in, hard = 0;
	for (int s = 0; s < nseg; s++) {
		emax = seg[s].emax > emax ? seg[s].emax : emax;
		hard |= seg[s].flag == SIG_HARD;
	}
	t->e_in = seg[0].e_in;
	t->logk = emax - seg[0].e_in + 1;
	t->flag = (hard || t->logk > SIG_KT_LOG_MAX) ? SIG_HARD : SIG_TABLE;
	t->pad = 0;
	const uint64_t K = 1ull << (t->logk > SIG_KT_LOG_MAX ? SIG_KT_LOG_MAX : t->logk);
	// seg[0].base is the approximate y_in with fewer low bits cleared (or none): clear the tile's
	t->base = dfrombits(dbits(seg[0].base) & ~(K - 1));
}

// group header from its segments; the entries are filled by the candidates' walks
CMCB_HD void sig_group_header(const SigSeg *seg, int nseg, SigGrp *g)
{
	int emax = seg[0].e_in, hard = 0;
	for (int s = 0; s < nseg; s++) {
		emax = seg[s].emax > emax ? seg[s].emax : emax;
		hard |= seg[s].flag == SIG_HARD;
	}
	g->e_in = seg[0].e_in;
	g->emax = emax;
	g->logk = emax - seg[0].e_in + 1;
	g->flag = hard ? SIG_HARD : (g->logk > SIG_KG_LOG_MAX ? SIG_SEQ : SIG_TABLE);
	const uint64_t K = 1ull << (g->logk > SIG_KG_LOG_MAX ? SIG_KG_LOG_MAX : g->logk);
	g->base = dfrombits(dbits(seg[0].base) & ~(K - 1));
}
CMCB_HD double sig_group_candidate(const SigGrp &g, int r) { return dfrombits(dbits(g.base) | (uint64_t) r); }

// tile header from its groups
CMCB_HD void sig_tile_header_from_groups(const SigGrp *grp, int ngrp, SigTile *t)
{
	int emax = grp[0].e_in, hard = 0;
	for (int g = 0; g < ngrp; g++) {
		emax = grp[g].emax > emax ? grp[g].emax : emax;
		hard |= grp[g].flag == SIG_HARD;
	}
	t->e_in = grp[0].e_in;
	t->logk = emax - grp[0].e_in + 1;
	t->flag = (hard || t->logk > SIG_KT_LOG_MAX) ? SIG_HARD : SIG_TABLE;
	t->pad = 0;
	t->pad2 = 0.0;
	const uint64_t K = 1ull << (t->logk > SIG_KT_LOG_MAX ? SIG_KT_LOG_MAX : t->logk);
	t->base = dfrombits(dbits(grp[0].base) & ~(K - 1));
}

CMCB_HD double sig_tile_candidate(const SigTile &t, int r) { return dfrombits(dbits(t.base) | (uint64_t) r); }

// the drift bound of tile t from the inclusive prefix sums of the terms over blocks of SIG_TILE entries (index q = k + p - 1): the tile's
// last window ends inside block t + 1
CMCB_HD double sig_dband(const double *block_prefix_incl, long n_blocks, long tile, long p)
{
	const long b = tile + 1 < n_blocks ? tile + 1 : n_blocks - 1;
	return dmul(dmul(2.220446049250313e-16 /* 2^-52 */, (double) (12 * p)), block_prefix_incl[b]);
}

// number of tiles of a rank with n local stars: stars 2 .. n in tiles of SIG_TILE
CMCB_HD long sig_num_tiles(long n) { return n >= 2 ? (n - 1 + SIG_TILE - 1) / SIG_TILE : 0; }

} // namespace cmcb

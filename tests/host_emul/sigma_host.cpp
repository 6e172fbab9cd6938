// TEST-ONLY host build of cmc_cosmic_b200/csrc/sigma_scan.h: the four device passes of sigma.cu run as plain loops over tiles and segments,
// calling the same per-segment / per-tile code, so that the decomposition can be checked against the oracle without a GPU.
#include <math.h>
#include <vector>
#include "../../cmc_cosmic_b200/csrc/sigma_scan.h"

using namespace cmcb;

namespace {

struct Chain {
	std::vector<double> term; // [n + 2p]
	std::vector<double> dband;
	std::vector<SigTile> tile;
	std::vector<double> yin;  // exact value entering each tile
	std::vector<double> y;    // exact value at every star [n + 1]
};

void build_segments(const SigmaWindow &w, const SigTerms &T, long tile, double dband, std::vector<SigSeg> &seg, int *nseg, int *broken)
{
	const long s_first = 2 + tile * SIG_TILE, s_end = s_first + SIG_TILE < w.n + 1 ? s_first + SIG_TILE : w.n + 1;
	*nseg = 0;
	for (long s0 = s_first; s0 < s_end; s0 += SIG_S) {
		const int cnt = (int) (s0 + SIG_S <= s_end ? SIG_S : s_end - s0);
		double sub[SIG_S], add[SIG_S];
		sig_segment_terms(w, T, s0, cnt, sub, add);
		const double y = sig_fresh_window(w, T, s0 - 1);
		sig_build_segment(sub, add, cnt, y, dband, w.p, &seg[*nseg], broken);
		(*nseg)++;
	}
}

// segments [s_lo, s_hi) of the tile from y
double walk_segments(const SigmaWindow &w, const SigTerms &T, long tile, const std::vector<SigSeg> &seg, int s_lo, int s_hi, double y, double *seg_yin, int *inconsistent)
{
	const long s_first = 2 + tile * SIG_TILE, s_end = s_first + SIG_TILE < w.n + 1 ? s_first + SIG_TILE : w.n + 1;
	for (int s = s_lo; s < s_hi; s++) {
		const long s0 = s_first + (long) s * SIG_S;
		const int cnt = (int) (s0 + SIG_S <= s_end ? SIG_S : s_end - s0);
		if (seg_yin) seg_yin[s] = y;
		double sub[SIG_S], add[SIG_S];
		sig_segment_terms(w, T, s0, cnt, sub, add);
		y = sig_apply_segment(seg[s], sub, add, cnt, y, inconsistent);
	}
	return y;
}

// group tables of a tile (after its segments): what the (group, candidate) threads of the device do
void build_groups(const SigmaWindow &w, const SigTerms &T, long tile, const std::vector<SigSeg> &seg, int nseg, std::vector<SigGrp> &grp, int *ngrp, int *inconsistent)
{
	*ngrp = (nseg + SIG_G - 1) / SIG_G;
	for (int g = 0; g < *ngrp; g++) {
		const int s_lo = g * SIG_G, s_hi = s_lo + SIG_G < nseg ? s_lo + SIG_G : nseg;
		sig_group_header(&seg[s_lo], s_hi - s_lo, &grp[g]);
		if (grp[g].flag == SIG_TABLE)
			for (int r = 0; r < (1 << grp[g].logk); r++)
				grp[g].yout[r] = walk_segments(w, T, tile, seg, s_lo, s_hi, sig_group_candidate(grp[g], r), nullptr, inconsistent);
	}
}

// the tile's groups from y: a table lookup per group, or its segments one by one
double walk_groups(const SigmaWindow &w, const SigTerms &T, long tile, const std::vector<SigSeg> &seg, int nseg, const std::vector<SigGrp> &grp, int ngrp, double y,
		   double *grp_yin, int *inconsistent)
{
	for (int g = 0; g < ngrp; g++) {
		if (grp_yin) grp_yin[g] = y;
		const int s_lo = g * SIG_G, s_hi = s_lo + SIG_G < nseg ? s_lo + SIG_G : nseg;
		if (grp[g].flag == SIG_TABLE && sig_expo(y) == grp[g].e_in) {
			const uint64_t idx = dbits(y) & ((1ull << grp[g].logk) - 1);
			y = dadd(grp[g].yout[idx], dsub(y, dfrombits(dbits(grp[g].base) | idx)));
		} else {
			if (grp[g].flag == SIG_TABLE) *inconsistent = 1;
			y = walk_segments(w, T, tile, seg, s_lo, s_hi, y, nullptr, inconsistent);
		}
	}
	return y;
}

} // namespace

// vr, vt: local stars and ghosts, local index k = 1 - p .. n + p at [k + p - 1].  counters: [0] hard segments, [1] wide (replayed) segments,
// [2] hard tiles, [3] broken, [4] inconsistent, [5] tiles.  band_scale inflates the drift bound (test hook: forces hard segments).
extern "C" int emul_sigma_scan(long p, long n, long first_global, long N_MAX, const double *star_r, const double *star_m, double madhoc, const double *vr,
			       const double *vt, int mave, double band_scale, double *out_a, double *out_sigma, long *counters)
{
	SigmaWindow w;
	w.p = p; w.n = n; w.first_global = first_global; w.N_MAX = N_MAX; w.first_rank = first_global == 1;
	for (int i = 0; i < 6; i++) counters[i] = 0;
	if (!w.valid()) return -1;
	Chain ch[2];
	for (int c = 0; c < 2; c++) ch[c].term.assign(n + 2 * p, 0.0);
	for (long k = 1 - p; k <= n + p; k++) {
		const long g = first_global + k - 1;
		if (g < 1 || g > N_MAX) continue;
		sigma_terms(star_m[g], madhoc, vr[k + p - 1], vt[k + p - 1], &ch[0].term[k + p - 1], &ch[1].term[k + p - 1]);
	}
	const long ntiles = sig_num_tiles(n);
	counters[5] = ntiles;
	std::vector<SigSeg> seg(SIG_THREADS);
	std::vector<SigGrp> grp(SIG_NGRP);
	int broken = 0, inconsistent = 0;
	for (int c = 0; c < 2; c++) {
		Chain &C = ch[c];
		SigTerms T{C.term.data(), p};
		C.dband.assign(ntiles, 0.0); C.tile.resize(ntiles); C.yin.assign(ntiles + 1, 0.0); C.y.assign(n + 1, 0.0);
		// pass 1: prefix sums of the terms over blocks of SIG_TILE entries -> the tiles' drift bounds
		const long nq = n + 2 * p, nblk = (nq + SIG_TILE - 1) / SIG_TILE;
		std::vector<double> pre(nblk);
		double acc = 0.0;
		for (long b = 0; b < nblk; b++) {
			for (long q = b * SIG_TILE; q < nq && q < (b + 1) * SIG_TILE; q++) acc += C.term[q];
			pre[b] = acc;
		}
		for (long t = 0; t < ntiles; t++) C.dband[t] = sig_dband(pre.data(), nblk, t, p) * band_scale;
		// pass 2: segment tables -> tile tables
		for (long t = 0; t < ntiles; t++) {
			int nseg;
			build_segments(w, T, t, C.dband[t], seg, &nseg, &broken);
			for (int s = 0; s < nseg; s++) { counters[0] += seg[s].flag == SIG_HARD; counters[1] += seg[s].flag == SIG_SEQ; }
			int ngrp;
			build_groups(w, T, t, seg, nseg, grp, &ngrp, &inconsistent);
			sig_tile_header_from_groups(grp.data(), ngrp, &C.tile[t]);
			if (C.tile[t].flag == SIG_TABLE)
				for (int r = 0; r < (1 << C.tile[t].logk); r++)
					C.tile[t].yout[r] = walk_groups(w, T, t, seg, nseg, grp, ngrp, sig_tile_candidate(C.tile[t], r), nullptr, &inconsistent);
			else counters[2]++;
		}
	}
	counters[3] = broken;
	for (int c = 0; c < 2; c++) {
		Chain &C = ch[c];
		SigTerms T{C.term.data(), p};
		// pass 3: the chain walker
		double y = 0.0;
		for (long j = 0; j < 2 * p; j++) { long k; int sg; w.op(j, &k, &sg); y = dadd(y, T.at(k)); }
		C.y[1] = y;
		if (broken) { // one thread replays everything
			for (long si = 2; si <= n; si++) {
				const long a = w.simin(si - 1), b = w.simin(si);
				if (a != b) { y = dsub(y, T.at(a)); y = dadd(y, T.at(b + 2 * p - 1)); }
				C.y[si] = y;
			}
			continue;
		}
		for (long t = 0; t < ntiles; t++) {
			C.yin[t] = y;
			if (C.tile[t].flag == SIG_TABLE) y = sig_apply_tile_table(C.tile[t], y, &inconsistent);
			else {
				int nseg;
				build_segments(w, T, t, C.dband[t], seg, &nseg, nullptr);
				y = walk_segments(w, T, t, seg, 0, nseg, y, nullptr, &inconsistent);
			}
		}
		// pass 4: every tile from its exact y_in
		std::vector<double> seg_yin(SIG_THREADS);
		for (long t = 0; t < ntiles; t++) {
			int nseg;
			build_segments(w, T, t, C.dband[t], seg, &nseg, nullptr);
			int ngrp;
			double grp_yin[SIG_NGRP];
			build_groups(w, T, t, seg, nseg, grp, &ngrp, &inconsistent);
			walk_groups(w, T, t, seg, nseg, grp, ngrp, C.yin[t], grp_yin, &inconsistent);
			for (int g = 0; g < ngrp; g++) {
				const int s_lo = g * SIG_G, s_hi = s_lo + SIG_G < nseg ? s_lo + SIG_G : nseg;
				walk_segments(w, T, t, seg, s_lo, s_hi, grp_yin[g], seg_yin.data(), &inconsistent);
			}
			const long s_first = 2 + t * SIG_TILE, s_end = s_first + SIG_TILE < n + 1 ? s_first + SIG_TILE : n + 1;
			for (int s = 0; s < nseg; s++) {
				const long s0 = s_first + (long) s * SIG_S;
				const int cnt = (int) (s0 + SIG_S <= s_end ? SIG_S : s_end - s0);
				double sub[SIG_S], add[SIG_S], yy = seg_yin[s];
				sig_segment_terms(w, T, s0, cnt, sub, add);
				for (int i = 0; i < cnt; i++) { yy = dsub(yy, sub[i]); yy = dadd(yy, add[i]); C.y[s0 + i] = yy; }
			}
		}
	}
	counters[4] = inconsistent;
	for (long si = 1; si <= n; si++) {
		out_a[si] = mave ? ch[1].y[si] / 2. / p : star_r[first_global + si - 1];
		out_sigma[si] = sqrt(ch[0].y[si] / ch[1].y[si]);
	}
	return 0;
}

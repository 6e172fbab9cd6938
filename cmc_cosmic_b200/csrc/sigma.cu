// calc_sigma_r on the resident arrays (reference src/cmc/cmc_dynamics_helper.c:2821-2975): the sliding-window velocity
// dispersion around every local star with the reference's own bits -- its two running sums are replayed exactly through the segment and
// tile tables of sigma_scan.h.
//
//   sigma_terms_kernel    the two terms of every star of the rank and of its ghost stars (one strided pass over the records), and their
//                         sums over blocks of 2048 entries; sigma_prefix_kernel turns those into the tiles' drift bounds
//   sigma_tile_kernel     one CTA per (tile of 2048 stars, chain): the tile's terms staged in shared memory, a segment table per thread, a
//                         table per group of 16 segments (one thread per group and candidate), the tile table by walking its <= 128
//                         candidates through the 16 group tables
//   sigma_chain_kernel    one CTA per chain: one thread walks the tile tables in order (batches staged in shared memory) and leaves the
//                         exact value entering every tile; a tile without a table is rebuilt by the CTA and walked segment by segment
//   sigma_finish_kernel   one CTA per tile: segment tables again, one walk from the exact value, every thread replays its 16 operations
//                         for both chains and stores sigma and r / the mean mass
// Bound: HBM for the streaming part (32-byte sector of each record + 8 B of star_m in, 16 B of terms out and in twice, 16 B of results
// out = ~100 B/star), the serial walks (16 + 16 dependent shared-memory lookups per tile, one per tile in the chain kernel) for the rest.
#include <string.h>
#include "ctx.cuh"
#include "sigma_scan.h"

namespace cmcb {

constexpr int SIG_BATCH = 48;             // tile tables staged per step of the chain walker
constexpr int SIG_TERM_SMEM_MAX = 3072;   // terms of a tile + its halo: SIG_TILE + 2p + 1 <= this (p <= 511)

struct SigBuffers {
	double *term[2];       // [n + 2p] per chain
	double *blk[2];        // inclusive prefix sums of the terms over blocks of SIG_TILE entries
	SigTile *tile[2];      // [ntiles]
	double *yin[2];        // [ntiles] exact value entering the tile
	double *y1;            // [2] exact value at star 1
	double *ghost;         // [4][p]: prev vr, prev vt, next vr, next vt
	int *flags;            // [0] broken, [1] inconsistent, [2] hard segments, [3] wide segments, [4] hard tiles, [5] force sequential
	double *out_a, *out_sigma; // [n + 1]
};

struct SigArgs {
	const char *star1;
	const double *r, *m;
	SigmaWindow w;
	double madhoc;
	long n_tiles, n_blocks;
	int mave;
	int has_ghost_prev, has_ghost_next;
	SigBuffers b;
};

// ---- pass 1 -------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SIG_THREADS) sigma_terms_kernel(SigArgs a)
{
	__shared__ double red[2][SIG_THREADS / 32];
	const long nq = a.w.n + 2 * a.w.p;
	double sa = 0.0, sb = 0.0;
	for (int j = 0; j < SIG_S; j++) {
		const long q = (long) blockIdx.x * SIG_TILE + (long) j * SIG_THREADS + threadIdx.x; // coalesced over the block's 2048 entries
		if (q >= nq) break;
		const long k = q - a.w.p + 1, g = a.w.first_global + k - 1;
		double ta = 0.0, tb = 0.0;
		if (g >= 1 && g <= a.w.N_MAX) {
			double vr = 0.0, vt = 0.0;
			bool have = true;
			if (k < 1) { have = a.has_ghost_prev; if (have) { vr = a.b.ghost[k + a.w.p - 1]; vt = a.b.ghost[a.w.p + k + a.w.p - 1]; } }
			else if (k > a.w.n) { have = a.has_ghost_next; if (have) { vr = a.b.ghost[2 * a.w.p + k - a.w.n - 1]; vt = a.b.ghost[3 * a.w.p + k - a.w.n - 1]; } }
			else {
				const char *rec = a.star1 + (k - 1) * CMCB_SIZEOF_STAR;
				vr = *(const double *) (rec + CMCB_OFF_VR);
				vt = *(const double *) (rec + CMCB_OFF_VT);
			}
			if (have) sigma_terms(a.m[g], a.madhoc, vr, vt, &ta, &tb);
		}
		a.b.term[0][q] = ta;
		a.b.term[1][q] = tb;
		sa += ta;
		sb += tb;
	}
	for (int d = 16; d > 0; d >>= 1) { sa += __shfl_down_sync(0xffffffffu, sa, d); sb += __shfl_down_sync(0xffffffffu, sb, d); }
	if (lane_id() == 0) { red[0][warp_id()] = sa; red[1][warp_id()] = sb; }
	__syncthreads();
	if (threadIdx.x < 2) {
		double s = 0.0;
		for (int w = 0; w < SIG_THREADS / 32; w++) s += red[threadIdx.x][w];
		a.b.blk[threadIdx.x][blockIdx.x] = s;
	}
}

// inclusive prefix sums of the block sums, one warp per chain (a few thousand blocks): every lane sums a contiguous chunk (independent
// loads), the lanes' sums are scanned by shuffles, every lane writes its chunk's prefixes
__global__ void sigma_prefix_kernel(SigArgs a)
{
	const int c = threadIdx.x >> 5;
	if (c > 1) return;
	const long per = (a.n_blocks + 31) / 32, lo = (long) lane_id() * per, hi = lo + per < a.n_blocks ? lo + per : a.n_blocks;
	double s = 0.0;
	for (long i = lo; i < hi; i++) s += a.b.blk[c][i];
	double inc = s;
	for (int d = 1; d < 32; d <<= 1) {
		const double o = __shfl_up_sync(0xffffffffu, inc, d);
		if (lane_id() >= (unsigned) d) inc += o;
	}
	double run = inc - s;
	for (long i = lo; i < hi; i++) { run += a.b.blk[c][i]; a.b.blk[c][i] = run; }
}

// ---- per-tile machinery shared by the three kernels behind ------------------------------------------------------------------------
struct alignas(16) TileSmem {
	double term[SIG_TERM_SMEM_MAX];
	SigSeg seg[SIG_THREADS];
	SigGrp grp[SIG_NGRP];
	SigTile tile;
	double seg_yin[SIG_THREADS];
	double grp_yin[SIG_NGRP];
};

// stars [s_first, s_end) of a tile and the local indices its operations touch
struct TileRange {
	long s_first, s_end, k_lo;
	int nseg, nterm;
};
__device__ __forceinline__ TileRange tile_range(const SigmaWindow &w, long tile)
{
	TileRange t;
	t.s_first = 2 + tile * SIG_TILE;
	t.s_end = t.s_first + SIG_TILE < w.n + 1 ? t.s_first + SIG_TILE : w.n + 1;
	t.nseg = (int) ((t.s_end - t.s_first + SIG_S - 1) / SIG_S);
	// simin() never decreases along the stars: the first window of the tile starts at the lowest index its operations touch, the last one ends
	// at the highest (where the window has stopped sliding at the end of the cluster it lies further back than si - p)
	t.k_lo = w.simin(t.s_first - 1);
	long k_hi = w.simin(t.s_end - 1) + (2 * w.p - 1);
	t.nterm = (int) (k_hi - t.k_lo + 1);
	return t;
}

// the tile's terms into shared memory; afterwards a SigTerms over them is indexed with the same local indices k
__device__ __forceinline__ SigTerms tile_stage_terms(const SigArgs &a, int chain, const TileRange &t, double *sm)
{
	const double *src = a.b.term[chain] + (t.k_lo + a.w.p - 1);
	for (int i = threadIdx.x; i < t.nterm; i += SIG_THREADS) sm[i] = src[i];
	SigTerms T;
	T.t = sm - (t.k_lo + a.w.p - 1); // at(k) = t[k + p - 1] = sm[k - k_lo]
	T.p = a.w.p;
	return T;
}

__device__ __forceinline__ int seg_count(const TileRange &t, int s)
{
	const long s0 = t.s_first + (long) s * SIG_S;
	return (int) (s0 + SIG_S <= t.s_end ? SIG_S : t.s_end - s0);
}

// thread s builds segment s (after the terms are staged and the CTA has synchronised)
__device__ __forceinline__ void tile_build_segments(const SigArgs &a, const SigTerms &T, const TileRange &t, double dband, TileSmem *sm, bool count)
{
	const int s = threadIdx.x;
	if (s < t.nseg) {
		const long s0 = t.s_first + (long) s * SIG_S;
		const int cnt = seg_count(t, s);
		double sub[SIG_S], add[SIG_S];
		sig_segment_terms(a.w, T, s0, cnt, sub, add);
		const double y = sig_fresh_window(a.w, T, s0 - 1);
		int broken = 0;
		SigSeg sg;
		sig_build_segment(sub, add, cnt, y, dband, a.w.p, &sg, &broken);
		sm->seg[s] = sg;
		if (count) {
			if (broken) atomicExch(&a.b.flags[0], 1);
			if (sg.flag == SIG_HARD) atomicAdd(&a.b.flags[2], 1);
			if (sg.flag == SIG_SEQ) atomicAdd(&a.b.flags[3], 1);
		}
	}
}

// one thread walks segments [s_lo, s_hi) of the tile from y; seg_yin (may be null) receives the value entering every segment
__device__ __forceinline__ double tile_walk_segments(const SigArgs &a, const SigTerms &T, const TileRange &t, const TileSmem *sm, int s_lo, int s_hi, double y,
						     double *seg_yin, int *inconsistent)
{
	for (int s = s_lo; s < s_hi; s++) {
		if (seg_yin) seg_yin[s] = y;
		const SigSeg &sg = sm->seg[s];
		if (sg.flag == SIG_TABLE && sig_expo(y) == sg.e_in) {
			const uint64_t idx = dbits(y) & ((1ull << sg.logk) - 1);
			y = dadd(sg.yout[idx], dsub(y, dfrombits(dbits(sg.base) | idx)));
		} else {
			if (sg.flag == SIG_TABLE) *inconsistent = 1;
			double sub[SIG_S], add[SIG_S];
			const int cnt = seg_count(t, s);
			sig_segment_terms(a.w, T, t.s_first + (long) s * SIG_S, cnt, sub, add);
			y = sig_replay(sub, add, cnt, y);
		}
	}
	return y;
}

// the group tables of the tile, one thread per (group, candidate); the CTA has synchronised behind tile_build_segments, and is synchronised
// again when this returns
__device__ __forceinline__ void tile_build_groups(const SigArgs &a, const SigTerms &T, const TileRange &t, TileSmem *sm)
{
	const int g = threadIdx.x / SIG_KG_MAX, r = threadIdx.x % SIG_KG_MAX, ngrp = (t.nseg + SIG_G - 1) / SIG_G;
	const int s_lo = g * SIG_G, s_hi = s_lo + SIG_G < t.nseg ? s_lo + SIG_G : t.nseg;
	if (g < ngrp && r == 0) sig_group_header(&sm->seg[s_lo], s_hi - s_lo, &sm->grp[g]);
	__syncthreads();
	if (g < ngrp && sm->grp[g].flag == SIG_TABLE && r < (1 << sm->grp[g].logk)) {
		int inc = 0;
		sm->grp[g].yout[r] = tile_walk_segments(a, T, t, sm, s_lo, s_hi, sig_group_candidate(sm->grp[g], r), nullptr, &inc);
		if (inc) atomicExch(&a.b.flags[1], 1);
	}
	__syncthreads();
}

// one thread walks the tile's groups from y: a lookup per group, or the group's segments one by one
__device__ __forceinline__ double tile_walk_groups(const SigArgs &a, const SigTerms &T, const TileRange &t, const TileSmem *sm, double y, double *grp_yin, int *inconsistent)
{
	const int ngrp = (t.nseg + SIG_G - 1) / SIG_G;
	for (int g = 0; g < ngrp; g++) {
		if (grp_yin) grp_yin[g] = y;
		const SigGrp &gr = sm->grp[g];
		if (gr.flag == SIG_TABLE && sig_expo(y) == gr.e_in) {
			const uint64_t idx = dbits(y) & ((1ull << gr.logk) - 1);
			y = dadd(gr.yout[idx], dsub(y, dfrombits(dbits(gr.base) | idx)));
		} else {
			if (gr.flag == SIG_TABLE) *inconsistent = 1;
			const int s_lo = g * SIG_G, s_hi = s_lo + SIG_G < t.nseg ? s_lo + SIG_G : t.nseg;
			y = tile_walk_segments(a, T, t, sm, s_lo, s_hi, y, nullptr, inconsistent);
		}
	}
	return y;
}

// ---- pass 2: tile tables ----------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SIG_THREADS) sigma_tile_kernel(SigArgs a)
{
	extern __shared__ __align__(16) unsigned char smem_raw[];
	TileSmem *sm = (TileSmem *) smem_raw;
	const long tile = blockIdx.x;
	const int chain = blockIdx.y;
	const TileRange t = tile_range(a.w, tile);
	const SigTerms T = tile_stage_terms(a, chain, t, sm->term);
	__syncthreads();
	tile_build_segments(a, T, t, sig_dband(a.b.blk[chain], a.n_blocks, tile, a.w.p), sm, true);
	__syncthreads();
	tile_build_groups(a, T, t, sm);
	if (threadIdx.x == 0) {
		sig_tile_header_from_groups(sm->grp, (t.nseg + SIG_G - 1) / SIG_G, &sm->tile);
		if (sm->tile.flag == SIG_HARD) atomicAdd(&a.b.flags[4], 1);
	}
	__syncthreads();
	SigTile *out = &a.b.tile[chain][tile];
	if (sm->tile.flag == SIG_TABLE && threadIdx.x < (1u << sm->tile.logk)) {
		int inc = 0;
		out->yout[threadIdx.x] = tile_walk_groups(a, T, t, sm, sig_tile_candidate(sm->tile, threadIdx.x), nullptr, &inc);
		if (inc) atomicExch(&a.b.flags[1], 1);
	}
	if (threadIdx.x == 0) { out->base = sm->tile.base; out->e_in = sm->tile.e_in; out->logk = sm->tile.logk; out->flag = sm->tile.flag; out->pad = 0; }
}

// ---- pass 3: the chain ------------------------------------------------------------------------------------------------------------
// the whole loop by one thread, both chains, outputs included: windows that do not fit the shard, or a broken band
__device__ void sigma_sequential(const SigArgs &a)
{
	SigTerms TA{a.b.term[0], a.w.p}, TB{a.b.term[1], a.w.p};
	const long p = a.w.p, n = a.w.n;
	long siminlast = 1, simaxlast = a.w.first_rank ? 0 : -p;
	double ya = 0.0, yb = 0.0;
	for (long si = 1; si <= n; si++) {
		const long g_si = a.w.first_global + si - 1;
		long simin = si - p, simax = simin + (2 * p - 1);
		const long g_simin = g_si - p, g_simax = g_simin + (2 * p - 1);
		if (g_simin < 1) { simin = 1; simax = simin + (2 * p - 1); }
		else if (g_simax > a.w.N_MAX) { simax = n; simin = simax - (2 * p - 1); }
		for (long k = siminlast; k < simin; k++) { ya = dsub(ya, TA.at(k)); yb = dsub(yb, TB.at(k)); }
		for (long k = simaxlast + 1; k <= simax; k++) { ya = dadd(ya, TA.at(k)); yb = dadd(yb, TB.at(k)); }
		a.b.out_a[si] = a.mave ? ddiv(ddiv(yb, 2.), (double) p) : a.r[g_si];
		a.b.out_sigma[si] = dsqrt(ddiv(ya, yb));
		siminlast = simin;
		simaxlast = simax;
	}
}

__global__ void __launch_bounds__(SIG_THREADS) sigma_chain_kernel(SigArgs a)
{
	extern __shared__ __align__(16) unsigned char smem_raw[];
	TileSmem *sm = (TileSmem *) smem_raw;
	SigTile *batch = (SigTile *) (smem_raw + sizeof(TileSmem));
	__shared__ double s_y;
	__shared__ long s_pos;
	__shared__ int s_hard, s_seq;
	const int chain = blockIdx.x;
	if (threadIdx.x == 0) s_seq = a.b.flags[0] | a.b.flags[5];
	__syncthreads();
	if (s_seq) {
		if (chain == 0 && threadIdx.x == 0) sigma_sequential(a);
		return;
	}
	if (threadIdx.x == 0) {
		// the 2p additions that build the first window (reference :2951-2968 at si = 1)
		const SigTerms T{a.b.term[chain], a.w.p};
		double y = 0.0;
		const long k0 = (a.w.first_rank ? 0 : -a.w.p) + 1;
		for (long j = 0; j < 2 * a.w.p; j++) y = dadd(y, T.at(k0 + j));
		a.b.y1[chain] = y;
		s_y = y;
		s_pos = 0;
	}
	__syncthreads();
	// the batches of tile tables are double-buffered: the walker asks the TMA unit for the next batch (one bulk copy, completion counted on
	// an mbarrier) before it walks the current one
	__shared__ unsigned long long bar[2];
	unsigned phase[2] = {0, 0};
	auto stage = [&](long b0, int buf) { // thread 0 only
		const long nb = a.n_tiles - b0 < SIG_BATCH ? a.n_tiles - b0 : SIG_BATCH;
		if (nb <= 0) return;
		asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // the buffer's earlier reads are done before the async proxy overwrites it
		bulk_copy_g2s(batch + buf * SIG_BATCH, a.b.tile[chain] + b0, (unsigned) (nb * sizeof(SigTile)), &bar[buf]);
	};
	if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); }
	__syncthreads();
	if (threadIdx.x == 0) stage(0, 0);
	int cur = 0;
	for (long b0 = 0; b0 < a.n_tiles; b0 += SIG_BATCH, cur ^= 1) {
		const long nb = a.n_tiles - b0 < SIG_BATCH ? a.n_tiles - b0 : SIG_BATCH;
		const SigTile *bt = batch + cur * SIG_BATCH;
		bool first = true;
		for (;;) {
			if (threadIdx.x == 0) {
				if (first) {
					stage(b0 + SIG_BATCH, cur ^ 1);
					mbar_wait(&bar[cur], phase[cur]);
					phase[cur] ^= 1;
				}
				double y = s_y;
				long pos = s_pos;
				int hard = 0, inc = 0;
				// the header of the next tile is loaded while the current one's entry is on its way: the dependent chain per tile is
				// index -> one shared-memory load -> two additions
				int e_in = 0, logk = 0, flag = SIG_HARD;
				double base = 0.0;
				if (pos < b0 + nb) { const SigTile &tt = bt[pos - b0]; e_in = tt.e_in; logk = tt.logk; flag = tt.flag; base = tt.base; }
				while (pos < b0 + nb) {
					const SigTile &tt = bt[pos - b0];
					if (flag != SIG_TABLE || sig_expo(y) != e_in) { hard = 1; if (flag == SIG_TABLE) inc = 1; break; }
					const uint64_t idx = dbits(y) & ((1ull << logk) - 1);
					const double yo = tt.yout[idx], yc = dfrombits(dbits(base) | idx);
					a.b.yin[chain][pos] = y;
					pos++;
					if (pos < b0 + nb) { const SigTile &nx = bt[pos - b0]; e_in = nx.e_in; logk = nx.logk; flag = nx.flag; base = nx.base; }
					y = dadd(yo, dsub(y, yc));
				}
				if (inc) atomicExch(&a.b.flags[1], 1);
				s_y = y; s_pos = pos; s_hard = hard;
			}
			first = false;
			__syncthreads();
			if (!s_hard) break;
			// a tile without a table: rebuild its segments and groups with the whole CTA, then one thread walks them
			const long tile = s_pos;
			const TileRange t = tile_range(a.w, tile);
			const SigTerms T = tile_stage_terms(a, chain, t, sm->term);
			__syncthreads();
			tile_build_segments(a, T, t, sig_dband(a.b.blk[chain], a.n_blocks, tile, a.w.p), sm, false);
			__syncthreads();
			tile_build_groups(a, T, t, sm);
			if (threadIdx.x == 0) {
				int inc = 0;
				a.b.yin[chain][tile] = s_y;
				s_y = tile_walk_groups(a, T, t, sm, s_y, nullptr, &inc);
				s_pos = tile + 1;
				if (inc) atomicExch(&a.b.flags[1], 1);
			}
			__syncthreads();
		}
	}
}

// ---- pass 4: outputs --------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SIG_THREADS, 3) sigma_finish_kernel(SigArgs a)
{
	extern __shared__ __align__(16) unsigned char smem_raw[];
	TileSmem *sm = (TileSmem *) smem_raw;
	if (a.b.flags[0] | a.b.flags[5]) return; // sigma_sequential wrote everything
	const long tile = blockIdx.x;
	const TileRange t = tile_range(a.w, tile);
	double ya[SIG_S];
	const int s = threadIdx.x;
	const long s0 = t.s_first + (long) s * SIG_S;
	const int cnt = s < t.nseg ? seg_count(t, s) : 0;
	for (int chain = 0; chain < 2; chain++) {
		const SigTerms T = tile_stage_terms(a, chain, t, sm->term);
		__syncthreads();
		tile_build_segments(a, T, t, sig_dband(a.b.blk[chain], a.n_blocks, tile, a.w.p), sm, false);
		__syncthreads();
		tile_build_groups(a, T, t, sm);
		if (threadIdx.x == 0) {
			int inc = 0;
			tile_walk_groups(a, T, t, sm, a.b.yin[chain][tile], sm->grp_yin, &inc);
			if (inc) atomicExch(&a.b.flags[1], 1);
		}
		__syncthreads();
		if (threadIdx.x < (t.nseg + SIG_G - 1) / SIG_G) { // one thread per group: the values entering its segments
			int inc = 0;
			const int g = threadIdx.x, s_lo = g * SIG_G, s_hi = s_lo + SIG_G < t.nseg ? s_lo + SIG_G : t.nseg;
			tile_walk_segments(a, T, t, sm, s_lo, s_hi, sm->grp_yin[g], sm->seg_yin, &inc);
			if (inc) atomicExch(&a.b.flags[1], 1);
		}
		__syncthreads();
		if (s < t.nseg) {
			double sub[SIG_S], add[SIG_S], y = sm->seg_yin[s];
			sig_segment_terms(a.w, T, s0, cnt, sub, add);
#pragma unroll
			for (int i = 0; i < SIG_S; i++) {
				if (i < cnt) {
					y = dadd(dsub(y, sub[i]), add[i]);
					if (chain == 0) ya[i] = y;
					else {
						const long si = s0 + i;
						a.b.out_a[si] = a.mave ? ddiv(ddiv(y, 2.), (double) a.w.p) : a.r[a.w.first_global + si - 1];
						a.b.out_sigma[si] = dsqrt(ddiv(ya[i], y));
					}
				}
			}
		}
		__syncthreads();
	}
	if (tile == 0 && threadIdx.x == 0) { // star 1: the first window as it was built
		const double y1a = a.b.y1[0], y1b = a.b.y1[1];
		a.b.out_a[1] = a.mave ? ddiv(ddiv(y1b, 2.), (double) a.w.p) : a.r[a.w.first_global];
		a.b.out_sigma[1] = dsqrt(ddiv(y1a, y1b));
	}
}

// n == 1 .. : without tiles star 1 still needs its output
__global__ void sigma_star1_kernel(SigArgs a)
{
	if (a.b.flags[0] | a.b.flags[5]) return;
	const double y1a = a.b.y1[0], y1b = a.b.y1[1];
	a.b.out_a[1] = a.mave ? ddiv(ddiv(y1b, 2.), (double) a.w.p) : a.r[a.w.first_global];
	a.b.out_sigma[1] = dsqrt(ddiv(y1a, y1b));
}

} // namespace cmcb

struct SigScratch {
	SigBuffers b;
	long cap_n, cap_p;
};

void sigma_free(cmcb200_ctx *c)
{
	SigScratch *s = (SigScratch *) c->d_sigma;
	if (!s) return;
	for (int i = 0; i < 2; i++) { cudaFree(s->b.term[i]); cudaFree(s->b.blk[i]); cudaFree(s->b.tile[i]); cudaFree(s->b.yin[i]); }
	cudaFree(s->b.y1); cudaFree(s->b.ghost); cudaFree(s->b.flags); cudaFree(s->b.out_a); cudaFree(s->b.out_sigma);
	delete s;
	c->d_sigma = nullptr;
}

extern "C" int cmcb200_dev_calc_sigma_r(cmcb200_ctx *c, const cmcb200_sigma_params *p, double *sig_r_or_mave, double *sig_sigma, long stats[6])
{
	CMCB_CUDA_OK(cudaSetDevice(c->device));
	if (join_uploads(c)) return -1;
	const long n = p->n_local, P = p->p;
	if (n < 1 || P < 1 || n + 2 > c->n_star_dim) { set_error("calc_sigma_r: n_local %ld, p %ld do not fit (capacity %ld)", n, P, c->n_star_dim); return -1; }
	if (SIG_TILE + 2 * P + 1 > SIG_TERM_SMEM_MAX) { set_error("calc_sigma_r: window 2 x %ld too wide for the staged tiles", P); return -1; }
	SigScratch *s = (SigScratch *) c->d_sigma;
	if (s && (s->cap_p < P)) { sigma_free(c); s = nullptr; }
	if (!s) {
		s = new SigScratch();
		memset(s, 0, sizeof *s);
		const long cap = c->n_star_dim, nq = cap + 2 * P, nt = sig_num_tiles(cap) + 2, nblk = (nq + SIG_TILE - 1) / SIG_TILE + 1;
		for (int i = 0; i < 2; i++) {
			CMCB_CUDA_OK(cudaMalloc(&s->b.term[i], (size_t) nq * sizeof(double)));
			CMCB_CUDA_OK(cudaMalloc(&s->b.blk[i], (size_t) nblk * sizeof(double)));
			CMCB_CUDA_OK(cudaMalloc(&s->b.tile[i], (size_t) nt * sizeof(SigTile)));
			CMCB_CUDA_OK(cudaMalloc(&s->b.yin[i], (size_t) nt * sizeof(double)));
		}
		CMCB_CUDA_OK(cudaMalloc(&s->b.y1, 2 * sizeof(double)));
		CMCB_CUDA_OK(cudaMalloc(&s->b.ghost, (size_t) 4 * P * sizeof(double)));
		CMCB_CUDA_OK(cudaMalloc(&s->b.flags, 8 * sizeof(int)));
		CMCB_CUDA_OK(cudaMalloc(&s->b.out_a, (size_t) (cap + 1) * sizeof(double)));
		CMCB_CUDA_OK(cudaMalloc(&s->b.out_sigma, (size_t) (cap + 1) * sizeof(double)));
		s->cap_n = cap; s->cap_p = P;
		c->d_sigma = s;
		const int smem_tile = (int) sizeof(TileSmem), smem_chain = smem_tile + 2 * SIG_BATCH * (int) sizeof(SigTile);
		CMCB_CUDA_OK(cudaFuncSetAttribute(sigma_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_tile));
		CMCB_CUDA_OK(cudaFuncSetAttribute(sigma_finish_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_tile));
		CMCB_CUDA_OK(cudaFuncSetAttribute(sigma_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_chain));
	}
	SigArgs a;
	memset(&a, 0, sizeof a);
	a.star1 = star1(c);
	a.r = c->d_r; a.m = c->d_m;
	a.w.p = P; a.w.n = n; a.w.first_global = p->first_global; a.w.N_MAX = p->N_MAX; a.w.first_rank = p->first_global == 1;
	a.madhoc = 1.0 / ((double) p->N_STAR);
	a.mave = p->r_0_mave_1 != 0;
	a.b = s->b;
	a.n_tiles = a.w.valid() ? sig_num_tiles(n) : 0;
	a.n_blocks = (n + 2 * P + SIG_TILE - 1) / SIG_TILE;
	int hflags[8] = {0, 0, 0, 0, 0, a.w.valid() ? 0 : 1, 0, 0};
	if (p->force_sequential) hflags[5] = 1;
	// the ghost stars of the neighbouring ranks (what the reference exchanges with MPI_Send / MPI_Recv, :2862-2900)
	a.has_ghost_prev = p->ghost_prev_vr && p->ghost_prev_vt;
	a.has_ghost_next = p->ghost_next_vr && p->ghost_next_vt;
	double *hg = (double *) ((char *) c->h_pinned + 1024); // room for 4 x 96 doubles; larger windows go through a pageable copy
	const bool small = 4 * P * sizeof(double) <= 3072;
	std::vector<double> big;
	if (!small) { big.resize(4 * P); hg = big.data(); }
	const double *src[4] = {p->ghost_prev_vr, p->ghost_prev_vt, p->ghost_next_vr, p->ghost_next_vt};
	for (int i = 0; i < 4; i++)
		for (long j = 0; j < P; j++) hg[i * P + j] = src[i] ? src[i][j] : 0.0;
	CMCB_CUDA_OK(cudaMemcpyAsync(s->b.ghost, hg, (size_t) 4 * P * sizeof(double), cudaMemcpyHostToDevice, c->stream));
	int *hf = (int *) c->h_pinned;
	memcpy(hf, hflags, sizeof hflags);
	CMCB_CUDA_OK(cudaMemcpyAsync(s->b.flags, hf, sizeof hflags, cudaMemcpyHostToDevice, c->stream));
	if (!small) CMCB_CUDA_OK(cudaStreamSynchronize(c->stream));
	{
		StageTimer t(c, ST_SIGMA);
		const int smem_tile = (int) sizeof(TileSmem), smem_chain = smem_tile + 2 * SIG_BATCH * (int) sizeof(SigTile);
		sigma_terms_kernel<<<(unsigned) a.n_blocks, SIG_THREADS, 0, c->stream>>>(a);
		sigma_prefix_kernel<<<1, 64, 0, c->stream>>>(a);
		c->launches += 2;
		if (a.n_tiles > 0) {
			sigma_tile_kernel<<<dim3((unsigned) a.n_tiles, 2), SIG_THREADS, smem_tile, c->stream>>>(a);
			c->launches += 1;
		}
		sigma_chain_kernel<<<2, SIG_THREADS, smem_chain, c->stream>>>(a);
		c->launches += 1;
		if (a.n_tiles > 0) sigma_finish_kernel<<<(unsigned) a.n_tiles, SIG_THREADS, smem_tile, c->stream>>>(a);
		else sigma_star1_kernel<<<1, 1, 0, c->stream>>>(a);
		c->launches += 1;
	}
	CMCB_CUDA_OK(cudaGetLastError());
	CMCB_CUDA_OK(cudaMemcpyAsync(hf, s->b.flags, 8 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
	if (sig_r_or_mave) { StageTimer t(c, ST_D2H); CMCB_CUDA_OK(cudaMemcpyAsync(sig_r_or_mave + 1, s->b.out_a + 1, (size_t) n * sizeof(double), cudaMemcpyDeviceToHost, c->stream)); c->bytes_d2h += n * 8; }
	if (sig_sigma) { StageTimer t(c, ST_D2H); CMCB_CUDA_OK(cudaMemcpyAsync(sig_sigma + 1, s->b.out_sigma + 1, (size_t) n * sizeof(double), cudaMemcpyDeviceToHost, c->stream)); c->bytes_d2h += n * 8; }
	CMCB_CUDA_OK(cudaStreamSynchronize(c->stream));
	if (stats) { stats[0] = hf[2]; stats[1] = hf[3]; stats[2] = hf[4]; stats[3] = hf[0]; stats[4] = hf[1]; stats[5] = a.n_tiles; }
	if (hf[1] && !p->force_sequential) { // a table was entered from an unexpected binade: the band was not rigorous -- cannot happen; replay everything
		cmcb200_sigma_params q = *p;
		q.force_sequential = 1;
		const int rc = cmcb200_dev_calc_sigma_r(c, &q, sig_r_or_mave, sig_sigma, nullptr);
		if (stats) stats[4] = 1;
		return rc;
	}
	return 0;
}

// device pointers of the last calc_sigma_r's outputs ([n_local + 1], entry 0 unused) for kernels that consume them in place
extern "C" int cmcb200_dev_sigma_arrays(cmcb200_ctx *c, const double **sig_r_or_mave, const double **sig_sigma)
{
	SigScratch *s = (SigScratch *) c->d_sigma;
	if (!s) { set_error("sigma_arrays: no calc_sigma_r ran on this context"); return -1; }
	*sig_r_or_mave = s->b.out_a;
	*sig_sigma = s->b.out_sigma;
	return 0;
}

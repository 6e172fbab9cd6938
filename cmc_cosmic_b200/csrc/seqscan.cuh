// Device machinery for exactsum.h: a sequentially rounded FP64 recurrence y_i = fl(y_{i-1} +- a_i) evaluated by
// parallel scans with bit-identical results (see exactsum.h for the argument).
//
// Three launches per chain, all on one stream, no host synchronisation, chain length read from device memory:
//   seq_scan_classify   one pass over the terms, 2048 per CTA.  Decoupled look-back (Merrill & Garland) over
//                       tile aggregates gives every element an APPROXIMATE prefix; with it each element is
//                       classified (regular / event) and the tile's exact run sums are reduced in shared memory.
//                       Writes per tile: approximate exclusive prefix, event count, exact tail sum, event records.
//   seq_scan_resolve    one CTA.  Scans the per-tile summaries (event offsets, exact carry-in run sums), then a
//                       single thread applies the events in order with the hardware rounding (tens of events).
//   seq_scan_finalize   second pass over the terms: recomputes the same classification (same code, same inputs,
//                       same saved tile prefix -> identical), adds the exact carry-in, writes y through the
//                       chain's store().
// If a tile holds more than SEQ_EV_PER_TILE events or the chain more than SEQ_EV_MAX, `fallback` is set and
// seq_scan_resolve evaluates the whole chain with one thread (correct, slow, counted in stats).
//
// Bound: HBM.  Algorithmic traffic per element: terms read twice (8 B each pass per input array) + 8 B written.
#pragma once
#include "common.cuh"
#include "exactsum.h"

namespace cmcb {

constexpr int SEQ_THREADS = 256;
constexpr int SEQ_ITEMS = 8;
constexpr int SEQ_TILE = SEQ_THREADS * SEQ_ITEMS; // 2048
constexpr int SEQ_EV_PER_TILE = 64;
constexpr int SEQ_EV_MAX = 16384;
constexpr int SEQ_RESOLVE_THREADS = 1024;
constexpr int SEQ_EV_SMEM = 2048;

struct SeqEvent {
	double a;     // the term of the event element
	double gprev; // exact run sum inside the tile before it (since the tile start or the previous event of the tile)
};

// Scratch shared by the chains of one context (they run one after the other on one stream).
struct SeqWork {
	unsigned *tile_counter; // dynamic tile id
	// look-back descriptor of a tile: {value, state} in ONE 16-byte word (state 0 none, 1 value = approximate aggregate, 2 value =
	// approximate inclusive prefix), written and polled with single 128-bit relaxed device-scope accesses -- the pair is consistent by
	// construction, so the poll needs no acquire (ld.acquire costs an L1 invalidate + fence per poll) and the publish no fence
	double2 *tile_desc;
	double *tile_prefix;    // approximate exclusive prefix (saved for finalize)
	int *tile_evcount;      // events in the tile
	double *tile_tail;      // exact run sum after the tile's last event (whole tile if none)
	int *tile_evoff;        // exclusive scan of tile_evcount
	double *tile_carry;     // exact run sum entering the tile (since the last event before it)
	SeqEvent *tile_events;  // [ntiles][SEQ_EV_PER_TILE]
	double *ev_value;       // [SEQ_EV_MAX] exact chain value right after each event, in order
	double *ev_a, *ev_seg;  // [SEQ_EV_MAX] staging for the walker
	int *fallback;          // set when the event capacity is exceeded
	long *stats;            // [0] events, [1] fallbacks (accumulated)
	long max_tiles;
};

__device__ __forceinline__ void desc_publish(double2 *d, double value, int state)
{
	// ONE .b128 access: the PTX memory model guarantees single-copy atomicity for it (a .v2.b64 access is formally two scalar accesses;
	// both compile to STG.E.128.STRONG.GPU / LDG.E.128.STRONG.GPU)
	asm volatile("{ .reg .b128 q; mov.b128 q, {%1, %2}; st.relaxed.gpu.global.b128 [%0], q; }" ::"l"(d), "l"(__double_as_longlong(value)), "l"((long long) state) : "memory");
}
__device__ __forceinline__ int desc_poll(const double2 *d, double *value)
{
	long long v, s;
	asm volatile("{ .reg .b128 q; ld.relaxed.gpu.global.b128 q, [%2]; mov.b128 {%0, %1}, q; }" : "=l"(v), "=l"(s) : "l"(d) : "memory");
	*value = __longlong_as_double(v);
	return (int) s;
}

// ---- block-wide scans (256 threads) ----------------------------------------------------------------------------

// exclusive prefix of one double per thread, plus the block total; deterministic order
__device__ __forceinline__ double block_excl_sum(double v, double *smem_w /*[8]*/, double *total)
{
	double inc = v;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		double o = __shfl_up_sync(0xffffffffu, inc, d);
		if (lane_id() >= d) inc = dadd(o, inc);
	}
	if (lane_id() == 31) smem_w[warp_id()] = inc;
	__syncthreads();
	double wpre = 0.0, tot = 0.0;
#pragma unroll
	for (int w = 0; w < SEQ_THREADS / 32; w++) {
		double x = smem_w[w];
		if (w < warp_id()) wpre = dadd(wpre, x);
		tot = dadd(tot, x);
	}
	__syncthreads();
	*total = tot;
	double excl = __shfl_up_sync(0xffffffffu, inc, 1);
	if (lane_id() == 0) excl = 0.0;
	return dadd(wpre, excl);
}

// segmented-sum state: f = an event was seen, v = exact run sum since the last event (or since the start), c = events seen
struct SegState {
	double v;
	int f;
	int c;
};
__device__ __forceinline__ SegState seg_combine(SegState left, SegState right)
{
	SegState o;
	o.v = right.f ? right.v : left.v + right.v; // exact (multiples of one ulp within one binade)
	o.f = left.f | right.f;
	o.c = left.c + right.c;
	return o;
}
__device__ __forceinline__ SegState seg_shfl_up(SegState s, int d)
{
	SegState o;
	o.v = __shfl_up_sync(0xffffffffu, s.v, d);
	o.f = __shfl_up_sync(0xffffffffu, s.f, d);
	o.c = __shfl_up_sync(0xffffffffu, s.c, d);
	return o;
}
// exclusive segmented prefix over the threads of the block + block aggregate
__device__ __forceinline__ SegState block_excl_seg(SegState mine, SegState *smem_s /*[8]*/, SegState *total)
{
	SegState inc = mine;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		SegState o = seg_shfl_up(inc, d);
		if (lane_id() >= d) inc = seg_combine(o, inc);
	}
	if (lane_id() == 31) smem_s[warp_id()] = inc;
	__syncthreads();
	SegState wpre = {0.0, 0, 0}, tot = {0.0, 0, 0};
#pragma unroll
	for (int w = 0; w < SEQ_THREADS / 32; w++) {
		SegState x = smem_s[w];
		if (w < warp_id()) wpre = seg_combine(wpre, x);
		tot = seg_combine(tot, x);
	}
	__syncthreads();
	*total = tot;
	SegState excl = seg_shfl_up(inc, 1);
	if (lane_id() == 0) { excl.v = 0.0; excl.f = 0; excl.c = 0; }
	return seg_combine(wpre, excl);
}

// ---- per-tile work shared by classify and finalize --------------------------------------------------------------

template <class Chain>
struct TileData {
	double a[SEQ_ITEMS];
	SeqTerm t[SEQ_ITEMS];
};

// loads the tile's terms (chain order, blocked: thread t owns elements [t*ITEMS, (t+1)*ITEMS)) and returns the
// thread's approximate exclusive prefix inside the tile and the tile aggregate
template <class Chain>
__device__ __forceinline__ double tile_load(const Chain &ch, long n, long tile_start, TileData<Chain> &td, double *smem_w, double *tile_agg)
{
	long base = tile_start + (long) threadIdx.x * SEQ_ITEMS;
	double tsum = 0.0;
#pragma unroll
	for (int j = 0; j < SEQ_ITEMS; j++) {
		long i = base + j;
		td.a[j] = i < n ? ch.term(i) : 0.0;
		tsum = dadd(tsum, td.a[j]);
	}
	return block_excl_sum(tsum, smem_w, tile_agg);
}

// classification of the thread's elements from the approximate prefixes
template <class Chain>
__device__ __forceinline__ void tile_classify(const Chain &ch, long n, long tile_start, double y0, double tile_prefix, double thread_excl, TileData<Chain> &td)
{
	long base = tile_start + (long) threadIdx.x * SEQ_ITEMS;
	double run = thread_excl;
	double y0a = dabs(y0);
	double S_before = dadd(tile_prefix, run);
	double y_before = Chain::SIGN > 0 ? dadd(y0, S_before) : dsub(y0, S_before);
#pragma unroll
	for (int j = 0; j < SEQ_ITEMS; j++) {
		long i = base + j;
		run = dadd(run, td.a[j]);
		double S_after = dadd(tile_prefix, run);
		double y_after = Chain::SIGN > 0 ? dadd(y0, S_after) : dsub(y0, S_after);
		if (i < n) td.t[j] = seq_classify(td.a[j], y_before, y_after, y0a, i + 1);
		else { td.t[j].v = 0.0; td.t[j].event = 0; }
		y_before = y_after;
	}
}

template <class Chain>
__device__ __forceinline__ SegState thread_seg_state(const TileData<Chain> &td)
{
	SegState s = {0.0, 0, 0};
#pragma unroll
	for (int j = 0; j < SEQ_ITEMS; j++) {
		if (td.t[j].event) { s.v = 0.0; s.f = 1; s.c++; }
		else s.v += td.t[j].v;
	}
	return s;
}

// ---- kernel 1: approximate look-back scan + classification --------------------------------------------------

template <class Chain>
__global__ void __launch_bounds__(SEQ_THREADS) seq_scan_classify(Chain ch, SeqWork w)
{
	__shared__ double smem_w[SEQ_THREADS / 32];
	__shared__ SegState smem_s[SEQ_THREADS / 32];
	__shared__ unsigned s_tile;
	__shared__ double s_prefix;
	const long n = ch.length();
	const double y0 = ch.y0();
	const long ntiles = (n + SEQ_TILE - 1) / SEQ_TILE;
	for (;;) {
		if (threadIdx.x == 0) s_tile = atomicAdd(w.tile_counter, 1u);
		__syncthreads();
		const long tile = s_tile;
		__syncthreads();
		if (tile >= ntiles) return;
		const long tile_start = tile * SEQ_TILE;
		TileData<Chain> td;
		double agg;
		double thread_excl = tile_load(ch, n, tile_start, td, smem_w, &agg);
		// publish the aggregate, then look back for the exclusive prefix (warp 0)
		if (threadIdx.x == 0) {
			if (tile == 0) { desc_publish(&w.tile_desc[0], agg, 2); s_prefix = 0.0; }
			else desc_publish(&w.tile_desc[tile], agg, 1);
		}
		if (tile > 0 && warp_id() == 0) {
			double prefix = 0.0;
			long look = tile - 1;
			for (;;) {
				long t = look - lane_id();
				int st = 0;
				double val = 0.0;
				if (t >= 0) {
					do { st = desc_poll(&w.tile_desc[t], &val); } while (st == 0);
				} else st = 2; // virtual predecessor with inclusive prefix 0
				unsigned incl_mask = __ballot_sync(0xffffffffu, st == 2);
				int first_incl = __ffs(incl_mask) - 1; // nearest predecessor holding an inclusive prefix
				double contrib = (!incl_mask || (int) lane_id() <= first_incl) ? val : 0.0;
#pragma unroll
				for (int d = 16; d > 0; d >>= 1) contrib = dadd(contrib, __shfl_xor_sync(0xffffffffu, contrib, d));
				prefix = dadd(prefix, contrib);
				if (incl_mask) break;
				look -= 32;
			}
			if (lane_id() == 0) {
				desc_publish(&w.tile_desc[tile], dadd(prefix, agg), 2);
				s_prefix = prefix;
			}
		}
		__syncthreads();
		const double tile_prefix = s_prefix;
		if (threadIdx.x == 0) w.tile_prefix[tile] = tile_prefix;
		tile_classify(ch, n, tile_start, y0, tile_prefix, thread_excl, td);
		SegState mine = thread_seg_state(td), tot;
		SegState pre = block_excl_seg(mine, smem_s, &tot);
		// event records in element order: ordinal = events before it in the tile
		SegState cur = pre;
#pragma unroll
		for (int j = 0; j < SEQ_ITEMS; j++) {
			if (td.t[j].event) {
				if (cur.c < SEQ_EV_PER_TILE) {
					SeqEvent e;
					e.a = td.a[j];
					e.gprev = cur.v;
					w.tile_events[tile * SEQ_EV_PER_TILE + cur.c] = e;
				}
				cur.v = 0.0; cur.f = 1; cur.c++;
			} else cur.v += td.t[j].v;
		}
		if (threadIdx.x == 0) {
			w.tile_evcount[tile] = tot.c;
			w.tile_tail[tile] = tot.v;
			if (tot.c > SEQ_EV_PER_TILE) atomicExch(w.fallback, 1);
		}
		__syncthreads();
	}
}

// ---- kernel 2: tile summaries + sequential event resolution ---------------------------------------------------

template <class Chain>
__global__ void __launch_bounds__(SEQ_RESOLVE_THREADS) seq_scan_resolve(Chain ch, SeqWork w, double *result)
{
	__shared__ double s_a[SEQ_EV_SMEM], s_seg[SEQ_EV_SMEM];
	__shared__ SegState s_warp[SEQ_RESOLVE_THREADS / 32];
	__shared__ SegState s_carry; // running (count, run sum) across chunks of tiles
	__shared__ int s_fallback;
	const long n = ch.length();
	const double y0 = ch.y0();
	const long ntiles = (n + SEQ_TILE - 1) / SEQ_TILE;
	const int tid = threadIdx.x;
	if (tid == 0) { s_carry.v = 0.0; s_carry.f = 0; s_carry.c = 0; s_fallback = *w.fallback; }
	__syncthreads();
	// (a) exclusive scans over tiles: event offsets and exact carry-in run sums
	for (long chunk = 0; chunk < ntiles && !s_fallback; chunk += SEQ_RESOLVE_THREADS) {
		long t = chunk + tid;
		SegState mine = {0.0, 0, 0};
		if (t < ntiles) { mine.c = w.tile_evcount[t]; mine.f = mine.c > 0; mine.v = w.tile_tail[t]; }
		SegState inc = mine;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			SegState o = seg_shfl_up(inc, d);
			if ((tid & 31) >= d) inc = seg_combine(o, inc);
		}
		if ((tid & 31) == 31) s_warp[tid >> 5] = inc;
		__syncthreads();
		SegState pre = s_carry;
		for (int q = 0; q < (tid >> 5); q++) pre = seg_combine(pre, s_warp[q]);
		SegState excl = seg_shfl_up(inc, 1);
		if ((tid & 31) == 0) { excl.v = 0.0; excl.f = 0; excl.c = 0; }
		excl = seg_combine(pre, excl);
		if (t < ntiles) { w.tile_evoff[t] = excl.c; w.tile_carry[t] = excl.v; }
		__syncthreads();
		if (tid == SEQ_RESOLVE_THREADS - 1) s_carry = seg_combine(excl, mine);
		__syncthreads();
	}
	const int n_events = s_carry.c;
	const double final_run = s_carry.v; // exact run sum after the last event through the end of the chain
	if (tid == 0 && n_events > SEQ_EV_MAX) { s_fallback = 1; *w.fallback = 1; }
	__syncthreads();
	if (s_fallback) { // capacity exceeded: evaluate the chain in order with one thread
		if (tid == 0) {
			double y = y0;
			for (long i = 0; i < n; i++) {
				double a = ch.term(i);
				y = Chain::SIGN > 0 ? dadd(y, a) : dsub(y, a);
				ch.store(i, y);
			}
			*result = y;
			w.stats[1] += 1;
		}
		return;
	}
	// (b) gather the event records in order
	for (long t = tid; t < ntiles; t += SEQ_RESOLVE_THREADS) {
		int c = w.tile_evcount[t];
		if (c) {
			int off = w.tile_evoff[t];
			double carry = w.tile_carry[t];
			for (int l = 0; l < c; l++) {
				SeqEvent e = w.tile_events[t * SEQ_EV_PER_TILE + l];
				double seg = l == 0 ? carry + e.gprev : e.gprev; // exact
				int q = off + l;
				if (q < SEQ_EV_SMEM) { s_a[q] = e.a; s_seg[q] = seg; }
				else { w.ev_a[q] = e.a; w.ev_seg[q] = seg; }
			}
		}
	}
	__syncthreads();
	// (c) the sequential part: each event is one exact add of the preceding run and one hardware-rounded add of the term
	if (tid == 0) {
		double y = y0;
		for (int q = 0; q < n_events; q++) {
			double seg = q < SEQ_EV_SMEM ? s_seg[q] : w.ev_seg[q];
			double a = q < SEQ_EV_SMEM ? s_a[q] : w.ev_a[q];
			double before = Chain::SIGN > 0 ? dadd(y, seg) : dsub(y, seg);
			y = Chain::SIGN > 0 ? dadd(before, a) : dsub(before, a);
			w.ev_value[q] = y;
		}
		*result = Chain::SIGN > 0 ? dadd(y, final_run) : dsub(y, final_run);
		w.stats[0] += n_events;
	}
}

// ---- kernel 3: exact values ---------------------------------------------------------------------------------

template <class Chain>
__global__ void __launch_bounds__(SEQ_THREADS) seq_scan_finalize(Chain ch, SeqWork w)
{
	__shared__ double smem_w[SEQ_THREADS / 32];
	__shared__ SegState smem_s[SEQ_THREADS / 32];
	if (*w.fallback) return;
	const long n = ch.length();
	const double y0 = ch.y0();
	const long ntiles = (n + SEQ_TILE - 1) / SEQ_TILE;
	for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
		const long tile_start = tile * SEQ_TILE;
		TileData<Chain> td;
		double agg;
		double thread_excl = tile_load(ch, n, tile_start, td, smem_w, &agg);
		tile_classify(ch, n, tile_start, y0, w.tile_prefix[tile], thread_excl, td);
		SegState mine = thread_seg_state(td), tot;
		SegState cur = block_excl_seg(mine, smem_s, &tot);
		const int evoff = w.tile_evoff[tile];
		const double carry = w.tile_carry[tile];
		const double base_tile = evoff > 0 ? w.ev_value[evoff - 1] : y0;
		long base = tile_start + (long) threadIdx.x * SEQ_ITEMS;
#pragma unroll
		for (int j = 0; j < SEQ_ITEMS; j++) {
			long i = base + j;
			double y;
			if (td.t[j].event) {
				y = w.ev_value[evoff + cur.c];
				cur.v = 0.0; cur.f = 1; cur.c++;
			} else {
				cur.v += td.t[j].v;
				double b = cur.c == 0 ? base_tile : w.ev_value[evoff + cur.c - 1];
				double run = cur.c == 0 ? carry + cur.v : cur.v; // exact
				y = Chain::SIGN > 0 ? dadd(b, run) : dsub(b, run);
			}
			if (i < n) ch.store(i, y);
		}
		__syncthreads();
	}
}

static __global__ void seq_scan_reset(SeqWork w, long ntiles)
{
	long i = blockIdx.x * (long) blockDim.x + threadIdx.x;
	if (i == 0) { *w.tile_counter = 0; *w.fallback = 0; }
	if (i < ntiles) w.tile_desc[i] = make_double2(0.0, 0.0); // state 0 (the state travels as the bit pattern of the second double)
}

// Launch the three kernels for one chain.  n_upper bounds ch.length() (which lives on the device).
// need_values = false: only the chain's final value is wanted (d_result, written by seq_scan_resolve) -- the second pass over the terms is skipped
template <class Chain>
inline void seq_scan_launch(const Chain &ch, const SeqWork &w, long n_upper, double *d_result, int grid_classify, int grid_finalize, cudaStream_t st, bool need_values = true)
{
	long ntiles = (n_upper + SEQ_TILE - 1) / SEQ_TILE;
	if (ntiles < 1) ntiles = 1;
	seq_scan_reset<<<(unsigned) ((ntiles + 255) / 256), 256, 0, st>>>(w, ntiles);
	int g1 = (int) (ntiles < grid_classify ? ntiles : grid_classify);
	seq_scan_classify<Chain><<<g1, SEQ_THREADS, 0, st>>>(ch, w);
	seq_scan_resolve<Chain><<<1, SEQ_RESOLVE_THREADS, 0, st>>>(ch, w, d_result);
	if (!need_values) return;
	int g3 = (int) (ntiles < grid_finalize ? ntiles : grid_finalize);
	seq_scan_finalize<Chain><<<g3, SEQ_THREADS, 0, st>>>(ch, w);
}

} // namespace cmcb

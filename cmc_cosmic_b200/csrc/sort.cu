// Radius sort of the star array on one GPU: qsorts_new -> sample_sort (reference src/cmc/cmc_sort.c:375-629 with
// procs == 1, wrapper src/cmc/cmc_utils.c:2232-2254), plus the AoS->SoA extraction of post_sort_comm
// (src/cmc/cmc_utils.c:2259-2288) fused into the record gather.
//
//   sort_extract_keys   key_i = order-preserving u64 image of star[i].r           (compare_type, cmc_sort.c:245-250)
//   sort_histogram      one pass over the keys: the 8 x 256 digit histograms of an LSD radix sort, and the count of
//                       keys == SF_INFINITY / > SF_INFINITY                        (remove_stripped_stars, :343-354)
//   sort_scan_hist      digit bases; passes whose digit is constant are skipped
//   sort_onesweep_pass  one kernel per 8-bit digit, single pass over (key, index) pairs: warp-level multisplit
//                       ranking, per-digit decoupled look-back across tiles (onesweep), scatter staged through
//                       shared memory so global writes are digit-contiguous.  Stable => ties keep input order.
//   sort_gather_records star_out[i] = star_in[perm[i]] (424-byte records moved as 8-byte words, whole warp per
//                       record), emits star_r / star_m and the per-star "has binary" flags
//   sort_binary_*       binaries follow their star: compacted in star order, binind renumbered 1,2,..  (:512-587, 716-771)
//
// Bound: HBM.  Algorithmic bytes per star: keys 8 read + 4 written (permutation); records 424 + 424.
#include "sort.cuh"
#include <cstdlib>

namespace cmcb {

// ---- keys ---------------------------------------------------------------------------------------------------

__host__ __device__ constexpr uint64_t key_sf_infinity() { return 0x4202A05F20000000ull | 0x8000000000000000ull; } // 1.0e10

__global__ void sort_extract_keys(const char *__restrict__ star, long n, long key_off, uint64_t *__restrict__ keys)
{
	long i = blockIdx.x * (long) blockDim.x + threadIdx.x;
	long stride = (long) gridDim.x * blockDim.x;
	for (; i < n; i += stride) keys[i] = key_from_double(*(const double *) (star + i * CMCB_SIZEOF_STAR + key_off));
}

// ---- histogram ----------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(512) sort_histogram(const uint64_t *__restrict__ keys, long n, SortState *st)
{
	__shared__ unsigned h[8][256];
	__shared__ unsigned s_eq, s_gt;
	for (int i = threadIdx.x; i < 8 * 256; i += blockDim.x) (&h[0][0])[i] = 0;
	if (threadIdx.x == 0) { s_eq = 0; s_gt = 0; }
	__syncthreads();
	long i = blockIdx.x * (long) blockDim.x + threadIdx.x;
	long stride = (long) gridDim.x * blockDim.x;
	unsigned eq = 0, gt = 0;
	const uint64_t kinf = key_sf_infinity();
	for (; i < n; i += stride) {
		uint64_t k = keys[i];
#pragma unroll
		for (int p = 0; p < 8; p++) atomicAdd(&h[p][(k >> (8 * p)) & 255], 1u);
		eq += k == kinf;
		gt += k > kinf;
	}
	if (eq) atomicAdd(&s_eq, eq);
	if (gt) atomicAdd(&s_gt, gt);
	__syncthreads();
	for (int j = threadIdx.x; j < 8 * 256; j += blockDim.x) {
		unsigned v = (&h[0][0])[j];
		if (v) atomicAdd(&st->hist[0][0] + j, v);
	}
	if (threadIdx.x == 0) {
		if (s_eq) atomicAdd((unsigned long long *) &st->n_eq_inf, (unsigned long long) s_eq);
		if (s_gt) atomicAdd((unsigned long long *) &st->n_gt_inf, (unsigned long long) s_gt);
	}
}

// exclusive scan of each digit histogram; decide which passes move data and where each pass reads from
__global__ void sort_scan_hist(SortState *st, long n)
{
	__shared__ unsigned s_w[8];
	const int p = blockIdx.x, d = threadIdx.x; // 8 blocks x 256 threads
	unsigned c = st->hist[p][d];
	unsigned inc = c;
	for (int o = 1; o < 32; o <<= 1) {
		unsigned v = __shfl_up_sync(0xffffffffu, inc, o);
		if ((d & 31) >= o) inc += v;
	}
	if ((d & 31) == 31) s_w[d >> 5] = inc;
	__syncthreads();
	unsigned pre = 0;
	for (int w = 0; w < (d >> 5); w++) pre += s_w[w];
	st->base[p][d] = pre + inc - c;
	if (c == (unsigned) n && n > 0) st->skip[p] = 1; // every key has this digit: the pass would be the identity
}

__global__ void sort_plan(SortState *st, long n)
{
	// which buffer (0 = A, 1 = B) each executed pass reads; the sorted pairs end up in buffer `final_buf`
	int cur = 0;
	for (int p = 0; p < 8; p++) {
		st->src_buf[p] = cur;
		if (!st->skip[p]) cur ^= 1;
	}
	st->final_buf = cur;
	// remove_stripped_stars (cmc_sort.c:343-354): pop trailing keys == SF_INFINITY
	st->n_new = st->n_gt_inf == 0 ? n - st->n_eq_inf : n;
	st->n_total = n;
}

__global__ void sort_reset(SortState *st, unsigned *tile_hist, long n_words)
{
	long i = blockIdx.x * (long) blockDim.x + threadIdx.x;
	if (i == 0) {
		st->n_eq_inf = 0; st->n_gt_inf = 0; st->n_bin = 0;
		for (int p = 0; p < 8; p++) { st->skip[p] = 0; st->tile_counter[p] = 0; }
	}
	if (i < 8 * 256) (&st->hist[0][0])[i] = 0;
	long stride = (long) gridDim.x * blockDim.x;
	for (; i < n_words; i += stride) tile_hist[i] = 0;
}

// ---- onesweep pass ------------------------------------------------------------------------------------------

constexpr unsigned LB_AGG = 1u << 30, LB_INCL = 2u << 30, LB_MASK = (1u << 30) - 1;

// lanes of the warp whose 8-bit digit equals this lane's: one vote per digit bit.  The eight votes of an item depend on nothing but its
// key, so the votes of a whole batch of items are issued back to back -- unlike MATCH.ANY, whose result each item's shared-memory counter
// update then waited for (48 % of the pass's stall samples in profiles/ncu_summary_r02b.txt were short-scoreboard waits behind it)
__device__ __forceinline__ unsigned digit_peers(unsigned d)
{
	unsigned m = 0xffffffffu;
#pragma unroll
	for (int b = 0; b < 8; b++) {
		const bool bit = (d >> b) & 1u;
		const unsigned v = __ballot_sync(0xffffffffu, bit);
		m &= bit ? v : ~v;
	}
	return m;
}

constexpr int SORT_RANK_BATCH = 8; // items whose peer masks are computed together before their counters are updated in order

template <bool FIRST, bool LAST, bool BALLOT>
__global__ void __launch_bounds__(SORT_THREADS, SORT_CTAS_PER_SM) sort_onesweep_pass(SortBuffers b, SortState *st, unsigned *tile_hist /*[ntiles][256]*/, long n, int pass)
{
	extern __shared__ __align__(16) unsigned char smem_raw[];
	uint64_t *s_keys = (uint64_t *) smem_raw;                       // [TILE]
	unsigned *s_vals = (unsigned *) (s_keys + SORT_TILE);            // [TILE]
	unsigned *s_cnt = s_vals + SORT_TILE;                            // [WARPS][256]
	unsigned *s_local_base = s_cnt + SORT_WARPS * 256;               // [256] start of each digit inside the tile
	unsigned *s_glob = s_local_base + 256;                           // [256] global start of this tile's run of each digit
	__shared__ unsigned s_tile, s_wsum[SORT_WARPS];

	if (st->skip[pass]) return;
	const int src = st->src_buf[pass];
	const uint64_t *__restrict__ kin = src ? b.keys_b : b.keys_a;
	const unsigned *__restrict__ vin = src ? b.vals_b : b.vals_a;
	uint64_t *__restrict__ kout = src ? b.keys_a : b.keys_b;
	unsigned *__restrict__ vout = src ? b.vals_a : b.vals_b;
	const int shift = 8 * pass;
	const long ntiles = (n + SORT_TILE - 1) / SORT_TILE;
	const unsigned lane = lane_id(), warp = warp_id();
	const unsigned lt_mask = (1u << lane) - 1;

	for (;;) {
		if (threadIdx.x == 0) s_tile = atomicAdd(&st->tile_counter[pass], 1u);
		for (int i = threadIdx.x; i < SORT_WARPS * 256; i += SORT_THREADS) s_cnt[i] = 0;
		__syncthreads();
		const long tile = s_tile;
		if (tile >= ntiles) return;
		const long tile_start = tile * SORT_TILE;
		const long tile_n = (n - tile_start) < SORT_TILE ? (n - tile_start) : SORT_TILE;

		// load: warp w owns elements [w*32*ITEMS, (w+1)*32*ITEMS) of the tile; item s of lane l is element s*32+l of it
		uint64_t key[SORT_ITEMS];
		unsigned val[SORT_ITEMS];
		unsigned short rank[SORT_ITEMS];
		const long wbase = tile_start + (long) warp * 32 * SORT_ITEMS;
#pragma unroll
		for (int s = 0; s < SORT_ITEMS; s++) {
			long i = wbase + s * 32 + lane;
			bool ok = i < n;
			key[s] = ok ? kin[i] : ~0ull; // padding sorts last inside digit 255 and is never written
			val[s] = FIRST ? (unsigned) i : (ok ? vin[i] : 0u);
		}
		// rank inside the warp's segment, stable: groups of equal digit via match.any, one counter update per group
		unsigned *cnt = s_cnt + warp * 256;
		if (BALLOT) {
#pragma unroll
			for (int s0 = 0; s0 < SORT_ITEMS; s0 += SORT_RANK_BATCH) {
				unsigned peers[SORT_RANK_BATCH];
#pragma unroll
				for (int t = 0; t < SORT_RANK_BATCH; t++) peers[t] = digit_peers((unsigned) (key[s0 + t] >> shift) & 255u);
#pragma unroll
				for (int t = 0; t < SORT_RANK_BATCH; t++) { // in item order: the ranking is stable
					const unsigned d = (unsigned) (key[s0 + t] >> shift) & 255u;
					const int leader = __ffs(peers[t]) - 1;
					unsigned old = 0;
					if ((int) lane == leader) { old = cnt[d]; cnt[d] = old + __popc(peers[t]); }
					old = __shfl_sync(0xffffffffu, old, leader);
					rank[s0 + t] = (unsigned short) (old + __popc(peers[t] & lt_mask));
					__syncwarp();
				}
			}
		} else {
#pragma unroll
			for (int s = 0; s < SORT_ITEMS; s++) {
				unsigned d = (unsigned) (key[s] >> shift) & 255u;
				unsigned peers = __match_any_sync(0xffffffffu, d);
				int leader = __ffs(peers) - 1;
				unsigned old = 0;
				if ((int) lane == leader) { old = cnt[d]; cnt[d] = old + __popc(peers); }
				old = __shfl_sync(0xffffffffu, old, leader);
				rank[s] = (unsigned short) (old + __popc(peers & lt_mask));
				__syncwarp();
			}
		}
		__syncthreads();
		// per digit (threads 0..255, thread d = digit d): turn the per-warp counts into warp offsets, get the tile count
		const unsigned d = threadIdx.x;
		const bool digit_thread = d < 256;
		unsigned tcount = 0, inc = 0, pad = 0;
		if (digit_thread) {
#pragma unroll
			for (int w = 0; w < SORT_WARPS; w++) {
				unsigned c = s_cnt[w * 256 + d];
				s_cnt[w * 256 + d] = tcount;
				tcount += c;
			}
			if (d == 255) tcount -= (unsigned) (SORT_TILE - tile_n); // padding lives at the end of digit 255
			// publish, then look back over earlier tiles for this digit.  Each look-back word holds its state AND its count, so
			// relaxed device-scope accesses are enough (ld.acquire compiles to a load + CCTL.IVALL + fence: the spin invalidated
			// the SM's whole L1 ~10 times per tile and digit -- 12 % of the pass's stall samples in ncu).
			unsigned *my = tile_hist + tile * 256 + d;
			unsigned excl = 0;
			if (tile == 0) st_relaxed_u32(my, LB_INCL | tcount);
			else {
				st_relaxed_u32(my, LB_AGG | tcount);
				for (long t = tile - 1; t >= 0; t--) {
					unsigned v;
					do { v = ld_relaxed_u32(tile_hist + t * 256 + d); } while ((v >> 30) == 0);
					excl += v & LB_MASK;
					if (v & LB_INCL) break;
				}
				st_relaxed_u32(my, LB_INCL | (excl + tcount));
			}
			s_glob[d] = st->base[pass][d] + excl;
			// position of each digit inside the tile (exclusive scan of tcount over the 256 digits)
			pad = (d == 255) ? (unsigned) (SORT_TILE - tile_n) : 0u;
			inc = tcount + pad;
#pragma unroll
			for (int o = 1; o < 32; o <<= 1) {
				unsigned v = __shfl_up_sync(0xffffffffu, inc, o);
				if (lane >= (unsigned) o) inc += v;
			}
			if (lane == 31) s_wsum[warp] = inc;
		}
		__syncthreads();
		if (digit_thread) {
			unsigned pre = 0;
#pragma unroll
			for (int w = 0; w < 256 / 32; w++) pre += (w < (int) warp) ? s_wsum[w] : 0u;
			s_local_base[d] = pre + inc - (tcount + pad);
		}
		__syncthreads();
		// scatter into shared memory in sorted-by-digit order
#pragma unroll
		for (int s = 0; s < SORT_ITEMS; s++) {
			unsigned dd = (unsigned) (key[s] >> shift) & 255u;
			unsigned pos = s_local_base[dd] + cnt[dd] + rank[s];
			s_keys[pos] = key[s];
			s_vals[pos] = val[s];
		}
		__syncthreads();
		// write out: consecutive threads -> consecutive positions of a digit run
		for (int p = threadIdx.x; p < (int) tile_n; p += SORT_THREADS) {
			uint64_t k = s_keys[p];
			unsigned dd = (unsigned) (k >> shift) & 255u;
			long dst = (long) s_glob[dd] + (p - (int) s_local_base[dd]);
			if (!LAST) kout[dst] = k;
			vout[dst] = s_vals[p];
		}
		__syncthreads();
	}
}

// ---- record gather -------------------------------------------------------------------------------------------

// star_out[i] = star_in[perm[i]]; also star_r[1+i], star_m[1+i] (post_sort_comm) and flag[i] = binind > 0 for i < n_new.
// One warp per record and iteration, 4 records in flight per warp; 8-byte words (424 = 53 * 8).
constexpr int GATHER_UNROLL = 4;
struct GatherFuse { // see SortArgs::fuse_transfer
	const double *old_r, *old_m;
	long n_old_local, first_global, N_MAX;
};

template <bool FUSE>
__global__ void __launch_bounds__(256) sort_gather_records(const char *__restrict__ in, char *__restrict__ out, SortBuffers b, const SortState *st,
						       double *__restrict__ star_r, double *__restrict__ star_m, unsigned char *__restrict__ binflag, GatherFuse f)
{
	const unsigned *__restrict__ perm = st->final_buf ? b.vals_b : b.vals_a;
	const long n = st->n_total, n_new = st->n_new;
	const unsigned lane = lane_id();
	long warp_global = (blockIdx.x * (long) blockDim.x + threadIdx.x) >> 5;
	long nwarps = ((long) gridDim.x * blockDim.x) >> 5;
	for (long i0 = warp_global * GATHER_UNROLL; i0 < n; i0 += nwarps * GATHER_UNROLL) {
		unsigned long long w0[GATHER_UNROLL], w1[GATHER_UNROLL], tab[GATHER_UNROLL];
		// all loads of the GATHER_UNROLL records first (independent, in flight together) ...
#pragma unroll
		for (int u = 0; u < GATHER_UNROLL; u++) {
			long i = i0 + u;
			if (i < n) {
				const long p = perm[i];
				const unsigned long long *src = (const unsigned long long *) (in + p * CMCB_SIZEOF_STAR);
				w0[u] = __ldg(src + lane);
				w1[u] = lane < CMCB_STAR_WORDS - 32 ? __ldg(src + 32 + lane) : 0ull;
				if (FUSE) { // the record's global index BEFORE the sort: get_global_idx (cmc_utils.c:2384-2407)
					const long j = p + 1;
					const long g = j > f.n_old_local ? 1 + f.N_MAX + (j - f.n_old_local) : f.first_global + j - 1;
					tab[u] = 0ull;
					if (lane == CMCB_OFF_M / 8) tab[u] = (unsigned long long) __double_as_longlong(__ldg(f.old_m + g));
					else if (lane == CMCB_OFF_ROLD / 8) tab[u] = (unsigned long long) __double_as_longlong(__ldg(f.old_r + g));
				}
			}
		}
		// ... then the transfer: words r 0, vr 1, vt 2 <- rnew 8, vrnew 9, vtnew 10; m 3 <- star_m[g]; rOld 11 <- star_r[g]
		if (FUSE) {
#pragma unroll
			for (int u = 0; u < GATHER_UNROLL; u++) {
				if (i0 + u < n) {
					const unsigned long long moved = __shfl_sync(0xffffffffu, w0[u], (lane + 8) & 31);
					if (lane < 3) w0[u] = moved;
					else if (lane == CMCB_OFF_M / 8 || lane == CMCB_OFF_ROLD / 8) w0[u] = tab[u];
				}
			}
		}
#pragma unroll
		for (int u = 0; u < GATHER_UNROLL; u++) {
			long i = i0 + u;
			if (i < n) {
				unsigned long long *dst = (unsigned long long *) (out + i * CMCB_SIZEOF_STAR);
				dst[lane] = w0[u];
				if (lane < CMCB_STAR_WORDS - 32) dst[32 + lane] = w1[u];
				if (i < n_new) {
					if (lane == CMCB_OFF_R / 8) star_r[1 + i] = __longlong_as_double((long long) w0[u]);
					if (lane == CMCB_OFF_M / 8) star_m[1 + i] = __longlong_as_double((long long) w0[u]);
					if (lane == CMCB_OFF_BININD / 8) binflag[i] = (long long) w0[u] > 0;
				}
			}
		}
	}
}

// ---- binaries --------------------------------------------------------------------------------------------------

constexpr int BIN_TILE = 2048; // flags per block (256 threads x 8)

__global__ void __launch_bounds__(256) sort_binary_count(const unsigned char *__restrict__ binflag, const SortState *st, unsigned *__restrict__ block_cnt)
{
	__shared__ unsigned s_w[8];
	const long n_new = st->n_new;
	long base = blockIdx.x * (long) BIN_TILE + threadIdx.x * 8;
	unsigned c = 0;
	for (int j = 0; j < 8; j++) c += (base + j < n_new) ? binflag[base + j] : 0;
	for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
	if (lane_id() == 0) s_w[warp_id()] = c;
	__syncthreads();
	if (threadIdx.x == 0) { unsigned t = 0; for (int w = 0; w < 8; w++) t += s_w[w]; block_cnt[blockIdx.x] = t; }
}

__global__ void __launch_bounds__(1024) sort_binary_scan(unsigned *block_cnt, long nblocks_max, SortState *st)
{
	__shared__ unsigned s_w[32];
	__shared__ unsigned s_carry;
	const long nblocks = (st->n_new + BIN_TILE - 1) / BIN_TILE;
	if (threadIdx.x == 0) s_carry = 0;
	__syncthreads();
	for (long c0 = 0; c0 < nblocks; c0 += 1024) {
		long t = c0 + threadIdx.x;
		unsigned v = t < nblocks ? block_cnt[t] : 0, inc = v;
		for (int o = 1; o < 32; o <<= 1) { unsigned x = __shfl_up_sync(0xffffffffu, inc, o); if (lane_id() >= (unsigned) o) inc += x; }
		if (lane_id() == 31) s_w[warp_id()] = inc;
		__syncthreads();
		unsigned pre = s_carry;
		for (unsigned w = 0; w < warp_id(); w++) pre += s_w[w];
		if (t < nblocks) block_cnt[t] = pre + inc - v;
		__syncthreads();
		if (threadIdx.x == 1023) s_carry = pre + inc;
		__syncthreads();
	}
	if (threadIdx.x == 0) st->n_bin = s_carry;
	(void) nblocks_max;
}

// renumber binind = 1,2,.. in star order and move the binary records (one warp per binary, 464 = 58 * 8 bytes)
__global__ void __launch_bounds__(256) sort_binary_move(char *__restrict__ star, const unsigned char *__restrict__ binflag, const unsigned *__restrict__ block_off,
						    const SortState *st, const char *__restrict__ bin_in, char *__restrict__ bin_out)
{
	__shared__ unsigned s_w[8];
	__shared__ unsigned s_idx[BIN_TILE]; // star index (within tile) of the b-th binary of this tile
	const long n_new = st->n_new;
	const long tile0 = blockIdx.x * (long) BIN_TILE;
	if (tile0 >= n_new) return;
	long base = tile0 + threadIdx.x * 8;
	unsigned f[8], c = 0;
	for (int j = 0; j < 8; j++) { f[j] = (base + j < n_new) ? binflag[base + j] : 0; c += f[j]; }
	unsigned inc = c;
	for (int o = 1; o < 32; o <<= 1) { unsigned x = __shfl_up_sync(0xffffffffu, inc, o); if (lane_id() >= (unsigned) o) inc += x; }
	if (lane_id() == 31) s_w[warp_id()] = inc;
	__syncthreads();
	unsigned pre = 0, tot = 0;
	for (unsigned w = 0; w < 8; w++) { if (w < warp_id()) pre += s_w[w]; tot += s_w[w]; }
	unsigned r = pre + inc - c;
	for (int j = 0; j < 8; j++) if (f[j]) s_idx[r++] = threadIdx.x * 8 + j;
	__syncthreads();
	const unsigned off = block_off[blockIdx.x];
	for (unsigned bidx = warp_id(); bidx < tot; bidx += 8) {
		long si = tile0 + s_idx[bidx];
		long long *pb = (long long *) (star + si * CMCB_SIZEOF_STAR + CMCB_OFF_BININD);
		long long old = *pb; // 1-based index into the pre-sort binary array
		long long nw = (long long) off + bidx + 1;
		const unsigned long long *src = (const unsigned long long *) (bin_in + (old - 1) * CMCB_SIZEOF_BINARY);
		unsigned long long *dst = (unsigned long long *) (bin_out + (nw - 1) * CMCB_SIZEOF_BINARY);
		unsigned long long a = src[lane_id()], c2 = lane_id() < CMCB_BINARY_WORDS - 32 ? src[32 + lane_id()] : 0ull;
		dst[lane_id()] = a;
		if (lane_id() < CMCB_BINARY_WORDS - 32) dst[32 + lane_id()] = c2;
		__syncwarp();
		if (lane_id() == 0) *pb = nw;
	}
}

// binary[n_bin+1 .. n_clear] = 0  (the reference memsets binary[1..N_BIN_DIM_OPT-1], cmc_sort.c:541)
__global__ void sort_binary_clear(char *bin_out, const SortState *st, long n_clear)
{
	long w0 = (long) st->n_bin * CMCB_BINARY_WORDS, w1 = n_clear * CMCB_BINARY_WORDS;
	long i = w0 + blockIdx.x * (long) blockDim.x + threadIdx.x;
	long stride = (long) gridDim.x * blockDim.x;
	for (; i < w1; i += stride) ((unsigned long long *) bin_out)[i] = 0ull;
}

// ---- host side ----------------------------------------------------------------------------------------------

static bool g_sort_ballot = true; // ranking by per-bit votes (default) or by MATCH.ANY

size_t sort_smem_bytes() { return SORT_TILE * 8 + SORT_TILE * 4 + SORT_WARPS * 256 * 4 + 256 * 4 * 2; }

int sort_setup()
{
	size_t sm = sort_smem_bytes();
	cudaError_t e;
	e = cudaFuncSetAttribute(sort_onesweep_pass<true, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sm);
	if (e == cudaSuccess) e = cudaFuncSetAttribute(sort_onesweep_pass<false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sm);
	if (e == cudaSuccess) e = cudaFuncSetAttribute(sort_onesweep_pass<false, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sm);
	if (e == cudaSuccess) e = cudaFuncSetAttribute(sort_onesweep_pass<true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sm);
	if (e == cudaSuccess) e = cudaFuncSetAttribute(sort_onesweep_pass<false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sm);
	if (e == cudaSuccess) e = cudaFuncSetAttribute(sort_onesweep_pass<false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sm);
	const char *k = getenv("CMCB200_SORT_RANK"); // "0": MATCH.ANY ranking (A/B measurement)
	g_sort_ballot = !(k && atoi(k) == 0);
	return e == cudaSuccess ? 0 : -1;
}

// identity permutation into buffer A, only needed when pass 0 was skipped (all low bytes equal)
__global__ void sort_iota_if_skipped(SortBuffers b, const SortState *st, long n)
{
	if (!st->skip[0]) return;
	long i = blockIdx.x * (long) blockDim.x + threadIdx.x;
	long stride = (long) gridDim.x * blockDim.x;
	for (; i < n; i += stride) b.vals_a[i] = (unsigned) i;
}

// stage 0: keys + digit histograms + plan
int sort_stage_keys(const SortArgs &a, cudaStream_t s)
{
	const long n = a.n;
	const long ntiles = (n + SORT_TILE - 1) / SORT_TILE;
	if (n > 0 && !a.keys_ready) sort_extract_keys<<<NUM_SMS_B200 * 4, 256, 0, s>>>(a.star_in, n, a.key_off, a.buf.keys_a);
	sort_reset<<<NUM_SMS_B200 * 2, 512, 0, s>>>(a.st, a.tile_hist, ntiles * 256 * 8);
	if (n > 0) sort_histogram<<<NUM_SMS_B200 * 2, 512, 0, s>>>(a.buf.keys_a, n, a.st);
	sort_scan_hist<<<8, 256, 0, s>>>(a.st, n);
	sort_plan<<<1, 1, 0, s>>>(a.st, n);
	return n > 0 ? (a.keys_ready ? 4 : 5) : 3;
}

// stage 1: the LSD passes (pass 0 generates the indices itself; when it is skipped they come from sort_iota_if_skipped)
int sort_stage_passes(const SortArgs &a, cudaStream_t s)
{
	const long n = a.n;
	if (n <= 0) return 0;
	const long ntiles = (n + SORT_TILE - 1) / SORT_TILE;
	sort_iota_if_skipped<<<NUM_SMS_B200 * 2, 256, 0, s>>>(a.buf, a.st, n);
	const int grid = (int) (ntiles < NUM_SMS_B200 * SORT_CTAS_PER_SM ? ntiles : NUM_SMS_B200 * SORT_CTAS_PER_SM);
	const size_t sm = sort_smem_bytes();
	for (int p = 0; p < 8; p++) {
		unsigned *th = a.tile_hist + (long) p * ntiles * 256;
		if (g_sort_ballot) {
			if (p == 0) sort_onesweep_pass<true, false, true><<<grid, SORT_THREADS, sm, s>>>(a.buf, a.st, th, n, p);
			else if (p == 7 && !a.keep_keys) sort_onesweep_pass<false, true, true><<<grid, SORT_THREADS, sm, s>>>(a.buf, a.st, th, n, p);
			else sort_onesweep_pass<false, false, true><<<grid, SORT_THREADS, sm, s>>>(a.buf, a.st, th, n, p);
		} else {
			if (p == 0) sort_onesweep_pass<true, false, false><<<grid, SORT_THREADS, sm, s>>>(a.buf, a.st, th, n, p);
			else if (p == 7 && !a.keep_keys) sort_onesweep_pass<false, true, false><<<grid, SORT_THREADS, sm, s>>>(a.buf, a.st, th, n, p);
			else sort_onesweep_pass<false, false, false><<<grid, SORT_THREADS, sm, s>>>(a.buf, a.st, th, n, p);
		}
	}
	return 9;
}

// stage 2: move the 424-byte records, emit star_r / star_m / binary flags
int sort_stage_gather(const SortArgs &a, cudaStream_t s)
{
	GatherFuse f{a.old_r, a.old_m, a.n_old_local, a.first_global, a.N_MAX};
	if (a.n > 0) {
		if (a.fuse_transfer) sort_gather_records<true><<<NUM_SMS_B200 * 8, 256, 0, s>>>(a.star_in, a.star_out, a.buf, a.st, a.star_r, a.star_m, a.binflag, f);
		else sort_gather_records<false><<<NUM_SMS_B200 * 8, 256, 0, s>>>(a.star_in, a.star_out, a.buf, a.st, a.star_r, a.star_m, a.binflag, f);
	}
	return a.n > 0 ? 1 : 0;
}

// count + scan of the binary flags of `n` stars (used by the sharded path, whose shard size is not what the local sort left in st->n_new)
__global__ void sort_set_n_new(SortState *st, long n) { st->n_new = n; st->n_total = n; st->n_bin = 0; }
int sort_binary_count_scan(const unsigned char *binflag, SortState *st, unsigned *bin_block_cnt, long n, cudaStream_t s)
{
	sort_set_n_new<<<1, 1, 0, s>>>(st, n);
	if (n <= 0) return 1;
	const long nbt = (n + BIN_TILE - 1) / BIN_TILE;
	sort_binary_count<<<(unsigned) nbt, 256, 0, s>>>(binflag, st, bin_block_cnt);
	sort_binary_scan<<<1, 1024, 0, s>>>(bin_block_cnt, nbt, st);
	return 3;
}
// the same when st->n_new already holds the shard size (set on the device) and the host only knows an upper bound for the grids
int sort_binary_count_scan_planned(const unsigned char *binflag, SortState *st, unsigned *bin_block_cnt, long n_upper, cudaStream_t s)
{
	if (n_upper <= 0) return 0;
	const long nbt = (n_upper + BIN_TILE - 1) / BIN_TILE;
	sort_binary_count<<<(unsigned) nbt, 256, 0, s>>>(binflag, st, bin_block_cnt);
	sort_binary_scan<<<1, 1024, 0, s>>>(bin_block_cnt, nbt, st);
	return 2;
}
int sort_binary_clear_device(char *bin_out, const SortState *st, long n_clear, cudaStream_t s)
{
	if (!bin_out || n_clear <= 0) return 0;
	sort_binary_clear<<<NUM_SMS_B200 * 2, 256, 0, s>>>(bin_out, st, n_clear);
	return 1;
}

// stage 3: binaries follow their star
int sort_stage_binaries(const SortArgs &a, cudaStream_t s)
{
	int launches = 0;
	if (a.n > 0) {
		launches += a.bin_in ? 3 : 2;
		const long nbt = (a.n + BIN_TILE - 1) / BIN_TILE;
		sort_binary_count<<<(unsigned) nbt, 256, 0, s>>>(a.binflag, a.st, a.bin_block_cnt);
		sort_binary_scan<<<1, 1024, 0, s>>>(a.bin_block_cnt, nbt, a.st);
		if (a.bin_in) sort_binary_move<<<(unsigned) nbt, 256, 0, s>>>(a.star_out, a.binflag, a.bin_block_cnt, a.st, a.bin_in, a.bin_out);
	}
	if (a.bin_out && a.n_bin_clear > 0) { sort_binary_clear<<<NUM_SMS_B200 * 2, 256, 0, s>>>(a.bin_out, a.st, a.n_bin_clear); launches++; }
	return launches;
}

} // namespace cmcb

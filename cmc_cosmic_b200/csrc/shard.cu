// Radius-sharded operation over the GPUs of one NVLink/NVSwitch node: one process, one context, one GPU per rank; rank g owns the
// sorted stars Start[g]..End[g] of findLimits -- the reference's own MPI partition (src/cmc/cmc_utils.c:2299-2333) -- and every rank
// holds the full star_r / star_m / star_phi tables, as in the reference (include/cmc/cmc_vars.h:216-226).
//
// What replaces sample_sort + load_balance + post_sort_comm (src/cmc/cmc_sort.c:375-791, src/cmc/cmc_utils.c:2259-2288):
//   1. every rank radix-sorts its own (key = rnew, index) pairs                                   [sort.cu]
//   2. shard_push_keys      the sorted keys go into every peer's key table by plain stores over NVLink (peer memory mapped with
//                           CUDA IPC), 8 B per star per peer
//   3. shard_rank_kernel    each local element computes its GLOBAL rank = own position + (number of smaller keys in every other rank's
//                           run; ties go to the lower rank, then to the lower local index: the stable order of the concatenated shards),
//                           hence its destination rank and its final position there -- the partition is realised exactly, so the
//                           reference's second exchange (load_balance) has nothing left to do
//   4. shard_push_records   the fused gather + all-to-all: each 424-byte record is read once from the local array and written once,
//                           straight to its final slot in the destination GPU's array (r <- rnew .. applied on the way): a warp
//                           per record, 8-byte lanes (the records are only 8-byte aligned: 424 = 53 * 8), two contiguous runs of 256 + 168
//                           bytes over NVLink; no staging buffer, no receive-side merge, no NCCL call on the data path
//   5. shard_publish_tables every rank writes the r / m of the shard it now holds into every peer's tables (the MPI_Allgatherv pair)
// Between the steps the ranks meet at a barrier: in the asynchronous form (cmcb200_shard_step_async, the default of the host mirror) flags in
// each other's device memory, posted and awaited by one-warp kernels on the same stream -- no host round trip; in the older form a barrier
// supplied by the HOST layer (torch.distributed / MPI).  Kernel completion makes the peer stores visible, the barrier orders them against the
// next step's reads.  The potential is then rebuilt on every rank from the
// full tables (replicated, like potential_calculate in the reference, cmc_utils.c:660,716).
#include <cstring>
#include <utility>
#include "ctx.cuh"
#include "potential.cuh"

namespace cmcb {

__global__ void shard_push_keys(const SortBuffers b, const SortState *st, uint64_t *const *peer_keys, int nranks, long slot_offset)
{
	const uint64_t *__restrict__ keys = st->final_buf ? b.keys_b : b.keys_a;
	const long n = st->n_new;
	for (long i = blockIdx.x * (long) blockDim.x + threadIdx.x; i < n; i += (long) gridDim.x * blockDim.x) {
		const uint64_t k = keys[i];
		for (int q = 0; q < nranks; q++) peer_keys[q][slot_offset + i] = k;
	}
}

struct ShardMeta { // per step, uploaded by the host
	long n_keep[SHARD_MAX_RANKS]; // stars each rank contributes (stripped stars already trimmed)
	long start0[SHARD_MAX_RANKS]; // 0-based global rank of the first star of each destination rank
	long count[SHARD_MAX_RANKS];  // stars of each destination rank
	long n_total;                 // their sum: the new clus.N_MAX
	int nranks, rank;
};

// ---- barriers between the ranks without the host ---------------------------------------------------------------------------------------
// Every rank owns a ShardSync block that all ranks map (CUDA IPC).  "post": a one-thread kernel, queued behind the kernels whose peer stores
// it publishes, writes this rank's values into every peer's block, fences at system scope and then raises its flag there to the current
// epoch (st.release.sys).  "wait": a one-warp kernel spins (ld.acquire.sys) until every rank's flag in the OWN block has reached the epoch;
// the kernels queued behind it then read what the peers stored.  Epochs only grow, so a flag is never reset.  A wait gives up after
// SHARD_WAIT_TIMEOUT_NS (a peer died): it sets `error` and lets the stream drain -- the host sees the flag at its next synchronisation.
struct ShardSync {
	unsigned long long flag[SHARD_MAX_RANKS][8]; // [source rank][phase]: epoch of that rank's last post
	long n_keep[SHARD_MAX_RANKS];                // stars each rank contributes to this step's exchange (posted in phase 0)
	int error;
};
constexpr unsigned long long SHARD_WAIT_TIMEOUT_NS = 4000000000ull;

__device__ __forceinline__ void st_release_sys_u64(unsigned long long *p, unsigned long long v)
{
	asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long *p)
{
	unsigned long long v;
	asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns()
{
	unsigned long long t;
	asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
	return t;
}

struct SyncPeers { ShardSync *p[SHARD_MAX_RANKS]; };

__global__ void shard_post_kernel(SyncPeers peers, int nranks, int rank, int phase, unsigned long long epoch, const SortState *st)
{
	if (threadIdx.x != 0 || blockIdx.x != 0) return;
	if (st) { const long n = st->n_new; for (int q = 0; q < nranks; q++) peers.p[q]->n_keep[rank] = n; }
	__threadfence_system();
	for (int q = 0; q < nranks; q++) st_release_sys_u64(&peers.p[q]->flag[rank][phase], epoch);
}

__global__ void shard_wait_kernel(ShardSync *mine, int nranks, int phase, unsigned long long epoch)
{
	const int q = threadIdx.x;
	if (q < nranks) {
		const unsigned long long t0 = globaltimer_ns();
		while (ld_acquire_sys_u64(&mine->flag[q][phase]) < epoch) {
			if (globaltimer_ns() - t0 > SHARD_WAIT_TIMEOUT_NS) { mine->error = 1 + phase; break; }
			__nanosleep(200);
		}
	}
}

// after the phase-0 wait: every rank's count is here; the new partition is findLimits over their sum (cmc_sort.c:416-421), computed by
// every rank for itself -- integer arithmetic, identical everywhere
__global__ void shard_plan_kernel(const ShardSync *mine, ShardMeta *meta, int nranks, int rank, int blk)
{
	if (threadIdx.x != 0 || blockIdx.x != 0) return;
	long total = 0;
	for (int q = 0; q < nranks; q++) { meta->n_keep[q] = mine->n_keep[q]; total += mine->n_keep[q]; }
	const long blocks = total / blk, remain = total % blk;
	const long chunk = (blocks / nranks) * blk, extra = blocks % nranks;
	for (int i = 0; i < nranks; i++) { // findIndices / findLimits, src/cmc/cmc_utils.c:2299-2333
		long b, e;
		if (i < extra) { b = i * chunk + (long) i * blk + 1; e = b + chunk + blk - 1; }
		else { b = i * chunk + extra * blk + 1; e = b + chunk - 1; }
		if (i == nranks - 1) e += remain;
		meta->start0[i] = b - 1;
		meta->count[i] = e - b + 1;
	}
	meta->nranks = nranks;
	meta->rank = rank;
	meta->n_total = total;
}

// the sort state describes the local sorted run until the records have left; from then on the shard this rank received
__global__ void shard_adopt_shard_kernel(SortState *st, const ShardMeta *meta)
{
	if (threadIdx.x != 0 || blockIdx.x != 0) return;
	st->n_new = meta->count[meta->rank];
	st->n_total = st->n_new;
	st->n_bin = 0;
}

__device__ __forceinline__ long lower_bound_u64(const uint64_t *__restrict__ a, long n, uint64_t key) // #elements < key
{
	long lo = 0, hi = n;
	while (lo < hi) {
		const long mid = (lo + hi) >> 1;
		if (a[mid] < key) lo = mid + 1; else hi = mid;
	}
	return lo;
}
__device__ __forceinline__ long upper_bound_u64(const uint64_t *__restrict__ a, long n, uint64_t key) // #elements <= key
{
	long lo = 0, hi = n;
	while (lo < hi) {
		const long mid = (lo + hi) >> 1;
		if (a[mid] <= key) lo = mid + 1; else hi = mid;
	}
	return lo;
}

__global__ void __launch_bounds__(256) shard_rank_kernel(const uint64_t *__restrict__ keys_all, long stride, const ShardMeta *__restrict__ meta,
							  unsigned char *__restrict__ dest_rank, unsigned *__restrict__ dest_pos)
{
	// A CTA ranks 256 CONSECUTIVE elements of the local sorted run: in every other rank's run their answers lie between the lower bound of the
	// first key and the upper bound of the last one.  2 (P - 1) threads find those two positions with full-length searches; the per-element
	// searches then run over the few hundred entries in between (cache-resident) instead of over the whole run.
	__shared__ long s_lo[SHARD_MAX_RANKS], s_hi[SHARD_MAX_RANKS];
	const int P = meta->nranks, g = meta->rank;
	const long n = meta->n_keep[g];
	const uint64_t *__restrict__ mine = keys_all + (long) g * stride;
	for (long base = blockIdx.x * 256L; base < n; base += gridDim.x * 256L) {
		const long last = base + 255 < n ? base + 255 : n - 1;
		if (threadIdx.x < 2 * P) {
			const int q = threadIdx.x >> 1;
			if (q != g) {
				const uint64_t *run = keys_all + (long) q * stride;
				if (threadIdx.x & 1) s_hi[q] = upper_bound_u64(run, meta->n_keep[q], mine[last]);
				else s_lo[q] = lower_bound_u64(run, meta->n_keep[q], mine[base]);
			}
		}
		__syncthreads();
		const long i = base + threadIdx.x;
		if (i < n) {
			const uint64_t key = mine[i];
			long rank = i;
			for (int q = 0; q < P; q++) {
				if (q == g) continue;
				const uint64_t *run = keys_all + (long) q * stride + s_lo[q];
				const long len = s_hi[q] - s_lo[q];
				rank += s_lo[q] + (q < g ? upper_bound_u64(run, len, key) : lower_bound_u64(run, len, key));
			}
			int d = 0;
			while (d + 1 < P && rank >= meta->start0[d + 1]) d++;
			dest_rank[i] = (unsigned char) d;
			dest_pos[i] = (unsigned) (rank - meta->start0[d]);
		}
		__syncthreads();
	}
}

struct PushArgs {
	const char *in;             // &star[1] of this rank before the sort
	char *out[SHARD_MAX_RANKS]; // &star[1] of every rank's output array
	unsigned char *src_out[SHARD_MAX_RANKS]; // every rank's "which rank did this record come from" array (binaries are pulled from there)
	int rank;
	const double *old_r, *old_m;
	long n_old_local, first_global, N_MAX;
};

// star_out[dest][pos] = transfer(star_in[perm[i]]): one warp per record, 8-byte words (424 = 53 * 8), GATHER_UNROLL records in flight
// FUSE: the ComputeIntermediateEnergy transfer is applied while the record moves (the key was rnew); otherwise the record moves as it is
// (the key was r: qsorts_new on records whose transfer the caller already made)
constexpr int PUSH_UNROLL = 4;
template <bool FUSE>
__global__ void __launch_bounds__(256) shard_push_records(PushArgs a, SortBuffers b, const SortState *st, const unsigned char *__restrict__ dest_rank,
							   const unsigned *__restrict__ dest_pos)
{
	const unsigned *__restrict__ perm = st->final_buf ? b.vals_b : b.vals_a;
	const long n = st->n_new; // stripped stars (r == SF_INFINITY, the tail of the local order) are not sent anywhere
	const unsigned lane = lane_id();
	const long warp_global = (blockIdx.x * (long) blockDim.x + threadIdx.x) >> 5;
	const long nwarps = ((long) gridDim.x * blockDim.x) >> 5;
	for (long i0 = warp_global * PUSH_UNROLL; i0 < n; i0 += nwarps * PUSH_UNROLL) {
		unsigned long long w0[PUSH_UNROLL], w1[PUSH_UNROLL], tab[PUSH_UNROLL];
		unsigned long long *dst[PUSH_UNROLL];
#pragma unroll
		for (int u = 0; u < PUSH_UNROLL; u++) {
			const long i = i0 + u;
			if (i < n) {
				const long p = perm[i];
				const unsigned long long *src = (const unsigned long long *) (a.in + p * CMCB_SIZEOF_STAR);
				w0[u] = __ldg(src + lane);
				w1[u] = lane < CMCB_STAR_WORDS - 32 ? __ldg(src + 32 + lane) : 0ull;
				tab[u] = 0ull;
				if (FUSE) {
					const long j = p + 1; // get_global_idx (cmc_utils.c:2384-2407) of the record before the sort
					const long g = j > a.n_old_local ? 1 + a.N_MAX + (j - a.n_old_local) : a.first_global + j - 1;
					if (lane == CMCB_OFF_M / 8) tab[u] = (unsigned long long) __double_as_longlong(__ldg(a.old_m + g));
					else if (lane == CMCB_OFF_ROLD / 8) tab[u] = (unsigned long long) __double_as_longlong(__ldg(a.old_r + g));
				}
				dst[u] = (unsigned long long *) (a.out[dest_rank[i]] + (long) dest_pos[i] * CMCB_SIZEOF_STAR);
			}
		}
		// ComputeIntermediateEnergy's transfer (cmc_utils.c:380-390): r, vr, vt <- rnew, vrnew, vtnew; m <- star_m[g]; rOld <- star_r[g]
#pragma unroll
		for (int u = 0; u < PUSH_UNROLL; u++) {
			if (i0 + u < n) {
				if (FUSE) {
					const unsigned long long moved = __shfl_sync(0xffffffffu, w0[u], (lane + 8) & 31);
					if (lane < 3) w0[u] = moved;
					else if (lane == CMCB_OFF_M / 8 || lane == CMCB_OFF_ROLD / 8) w0[u] = tab[u];
				}
				dst[u][lane] = w0[u];
				if (lane < CMCB_STAR_WORDS - 32) dst[u][32 + lane] = w1[u];
				if (lane == 0) a.src_out[dest_rank[i0 + u]][dest_pos[i0 + u]] = (unsigned char) a.rank;
			}
		}
	}
}

// the shard this rank holds after the exchange: r, m of its records into every rank's (alternate) tables; binary flags for the local re-pack
__global__ void shard_publish_tables(const char *__restrict__ star1, long n_local, long start0, double *const *peer_r, double *const *peer_m, int nranks,
				     unsigned char *__restrict__ binflag)
{
	for (long i = blockIdx.x * (long) blockDim.x + threadIdx.x; i < n_local; i += (long) gridDim.x * blockDim.x) {
		const char *rec = star1 + i * CMCB_SIZEOF_STAR;
		const double r = *(const double *) (rec + CMCB_OFF_R), m = *(const double *) (rec + CMCB_OFF_M);
		binflag[i] = *(const long long *) (rec + CMCB_OFF_BININD) > 0;
		for (int q = 0; q < nranks; q++) {
			peer_r[q][1 + start0 + i] = r;
			peer_m[q][1 + start0 + i] = m;
		}
	}
}

// the same with the shard's extent taken from the device-side plan (asynchronous step: the host does not know it yet); one coalesced pass
// per destination rank -- consecutive threads store consecutive table entries of ONE peer
__global__ void __launch_bounds__(256) shard_publish_tables_planned(const char *__restrict__ star1, const ShardMeta *__restrict__ meta, double *const *peer_r,
									    double *const *peer_m, unsigned char *__restrict__ binflag)
{
	__shared__ double s_r[256], s_m[256];
	const int P = meta->nranks;
	const long n_local = meta->count[meta->rank], start0 = meta->start0[meta->rank];
	for (long base = blockIdx.x * 256L; base < n_local; base += gridDim.x * 256L) {
		const long i = base + threadIdx.x;
		if (i < n_local) {
			const char *rec = star1 + i * CMCB_SIZEOF_STAR;
			s_r[threadIdx.x] = *(const double *) (rec + CMCB_OFF_R);
			s_m[threadIdx.x] = *(const double *) (rec + CMCB_OFF_M);
			binflag[i] = *(const long long *) (rec + CMCB_OFF_BININD) > 0;
		}
		__syncthreads();
		const long cnt = n_local - base < 256 ? n_local - base : 256;
		// 2 P runs of `cnt` doubles: thread t stores element (t mod 256) of run (t / 256), so a warp writes 256 contiguous bytes of one peer
		for (long t = threadIdx.x; t < 2L * P * 256; t += 256) {
			const int run = (int) (t >> 8), e = (int) (t & 255);
			if (e < cnt) {
				const int q = run >> 1;
				double *dst = (run & 1) ? peer_m[q] : peer_r[q];
				dst[1 + start0 + base + e] = (run & 1) ? s_m[e] : s_r[e];
			}
		}
		__syncthreads();
	}
}

// ---- the new tables straight from the sorted keys -----------------------------------------------------------------------------------------
// After the global ranks are known every element of a rank's sorted run has its place in the new star_r / star_m tables: the run's owner sends
// (global rank, star_m) of its elements -- two contiguous arrays -- to every rank, and every rank scatters (r from the key it already holds, m)
// into its own tables locally.  Nothing of this needs the 424-byte records, so it runs on a second stream WHILE they travel, followed by the
// potential rebuild: at 8 GPUs the table exchange and the (replicated) potential disappear behind the record push.
struct RankPeers { unsigned *rank[SHARD_MAX_RANKS]; double *m[SHARD_MAX_RANKS]; };
template <bool FUSE>
__global__ void __launch_bounds__(256) shard_publish_ranks(PushArgs a, SortBuffers b, const SortState *st, const ShardMeta *__restrict__ meta,
							    const unsigned char *__restrict__ dest_rank, const unsigned *__restrict__ dest_pos, RankPeers rp, long slot_offset)
{
	const unsigned *__restrict__ perm = st->final_buf ? b.vals_b : b.vals_a;
	const long n = st->n_new;
	const int P = meta->nranks;
	for (long i = blockIdx.x * (long) blockDim.x + threadIdx.x; i < n; i += (long) gridDim.x * blockDim.x) {
		const unsigned grank = (unsigned) (meta->start0[dest_rank[i]] + dest_pos[i]);
		const long p = perm[i];
		double m;
		if (FUSE) { // star_m[g] of the record before the sort (ComputeIntermediateEnergy's m <- star_m[g], cmc_utils.c:387)
			const long j = p + 1;
			const long g = j > a.n_old_local ? 1 + a.N_MAX + (j - a.n_old_local) : a.first_global + j - 1;
			m = a.old_m[g];
		} else {
			m = *(const double *) (a.in + p * CMCB_SIZEOF_STAR + CMCB_OFF_M);
		}
		for (int q = 0; q < P; q++) { rp.rank[q][slot_offset + i] = grank; rp.m[q][slot_offset + i] = m; }
	}
}

__global__ void __launch_bounds__(256) shard_scatter_tables(const uint64_t *__restrict__ keys_all, const unsigned *__restrict__ rank_all, const double *__restrict__ m_all,
							     long stride, const ShardMeta *__restrict__ meta, double *__restrict__ star_r, double *__restrict__ star_m)
{
	const int q = blockIdx.y;
	const long n = meta->n_keep[q];
	const uint64_t *__restrict__ keys = keys_all + (long) q * stride;
	const unsigned *__restrict__ rk = rank_all + (long) q * stride;
	const double *__restrict__ mm = m_all + (long) q * stride;
	for (long i = blockIdx.x * (long) blockDim.x + threadIdx.x; i < n; i += (long) gridDim.x * blockDim.x) {
		const uint64_t k = keys[i];
		const uint64_t bits = (k >> 63) ? (k ^ 0x8000000000000000ull) : ~k; // inverse of key_from_double (sort.cu)
		const long pos = 1 + (long) rk[i];
		star_r[pos] = __longlong_as_double((long long) bits);
		star_m[pos] = mm[i];
	}
	if (q == 0 && blockIdx.x == 0 && threadIdx.x == 0) { // the entry behind the last star ends potential_calculate's count of N_MAX (cmc_utils.c:685)
		star_r[1 + meta->n_total] = CMCB200_SF_INFINITY;
		star_m[1 + meta->n_total] = 0.0;
	}
}

// flag[i] = star i of the shard this rank received has a binary
__global__ void shard_binflag_kernel(const char *__restrict__ star1, const ShardMeta *__restrict__ meta, unsigned char *__restrict__ binflag)
{
	const long n_local = meta->count[meta->rank];
	for (long i = blockIdx.x * (long) blockDim.x + threadIdx.x; i < n_local; i += (long) gridDim.x * blockDim.x)
		binflag[i] = *(const long long *) (star1 + i * CMCB_SIZEOF_STAR + CMCB_OFF_BININD) > 0;
}

// entries [first, first + n) of this rank's r / m / phi tables into the same entries of every peer's tables (coalesced runs per peer)
struct TablePeers { double *r[SHARD_MAX_RANKS], *m[SHARD_MAX_RANKS], *phi[SHARD_MAX_RANKS]; };
__global__ void shard_share_tables_kernel(const double *__restrict__ r, const double *__restrict__ m, const double *__restrict__ phi, long first, long n,
					  TablePeers tp, int nranks, int rank)
{
	const long stride = (long) gridDim.x * blockDim.x;
	for (long i = blockIdx.x * (long) blockDim.x + threadIdx.x; i < n; i += stride) {
		const double a = r[first + i], b = m[first + i], c = phi[first + i];
		for (int q = 0; q < nranks; q++) {
			if (q == rank) continue;
			tp.r[q][first + i] = a; tp.m[q][first + i] = b; tp.phi[q][first + i] = c;
		}
	}
}

// Binaries follow their star (cmc_sort.c:512-587, 716-771): the records of this rank's new shard carry the index their binary had in the
// SOURCE rank's binary array; in star order each binary_t is pulled from that rank's array over NVLink into the local array and binind
// becomes 1, 2, ...  (block_off = exclusive scan of the per-2048-star binary counts, as in sort_binary_move).
constexpr int SHARD_BIN_TILE = 2048;
struct BinPull {
	const char *bin_in[SHARD_MAX_RANKS]; // &binary[1] of every rank before the exchange
};
__global__ void __launch_bounds__(256) shard_binary_pull(char *__restrict__ star, const unsigned char *__restrict__ binflag, const unsigned char *__restrict__ src_rank,
							  const unsigned *__restrict__ block_off, long n_local_in, const SortState *st, BinPull bp, char *__restrict__ bin_out)
{
	__shared__ unsigned s_w[8];
	__shared__ unsigned s_idx[SHARD_BIN_TILE];
	const long n_local = st ? st->n_new : n_local_in; // asynchronous step: the shard size lives on the device
	const long tile0 = blockIdx.x * (long) SHARD_BIN_TILE;
	if (tile0 >= n_local) return;
	const long base = tile0 + threadIdx.x * 8;
	unsigned f[8], c = 0;
	for (int j = 0; j < 8; j++) { f[j] = (base + j < n_local) ? binflag[base + j] : 0; c += f[j]; }
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
		const long si = tile0 + s_idx[bidx];
		long long *pb = (long long *) (star + si * CMCB_SIZEOF_STAR + CMCB_OFF_BININD);
		const long long old = *pb; // 1-based index in the source rank's binary array
		const long long nw = (long long) off + bidx + 1;
		const unsigned long long *src = (const unsigned long long *) (bp.bin_in[src_rank[si]] + (old - 1) * CMCB_SIZEOF_BINARY);
		unsigned long long *dst = (unsigned long long *) (bin_out + (nw - 1) * CMCB_SIZEOF_BINARY);
		const unsigned long long x = src[lane_id()], y = lane_id() < CMCB_BINARY_WORDS - 32 ? src[32 + lane_id()] : 0ull;
		dst[lane_id()] = x;
		if (lane_id() < CMCB_BINARY_WORDS - 32) dst[32 + lane_id()] = y;
		__syncwarp();
		if (lane_id() == 0) *pb = nw;
	}
}

} // namespace cmcb

using namespace cmcb;

static void *own_buffer(cmcb200_ctx *c, int which)
{
	switch (which) {
	case SH_STAR0: return c->d_star[0];
	case SH_STAR1: return c->d_star[1];
	case SH_BIN0: return c->d_bin[0];
	case SH_BIN1: return c->d_bin[1];
	case SH_R: return c->r_parity ? (void *) c->d_r_alt : (void *) c->d_r;
	case SH_M: return c->r_parity ? (void *) c->d_m_alt : (void *) c->d_m;
	case SH_R_ALT: return c->r_parity ? (void *) c->d_r : (void *) c->d_r_alt;
	case SH_M_ALT: return c->r_parity ? (void *) c->d_m : (void *) c->d_m_alt;
	case SH_KEYS_ALL: return c->d_keys_all;
	case SH_SRC_ALL: return c->d_src_all;
	case SH_SYNC: return c->d_shard_sync;
	case SH_PHI: return c->d_phi;
	case SH_RANK_ALL: return c->d_rank_all;
	case SH_M_ALL: return c->d_m_all;
	}
	return nullptr;
}

extern "C" int cmcb200_shard_handle_bytes(void) { return SH_N_BUFFERS * (int) sizeof(cudaIpcMemHandle_t); }

// Allocates the exchange buffers for `nranks` ranks and writes this rank's CUDA IPC handles (cmcb200_shard_handle_bytes() bytes) to
// `handles`; the host layer allgathers them (MPI_Allgather / torch.distributed) and hands the concatenation to cmcb200_shard_attach.
extern "C" int cmcb200_shard_export(cmcb200_ctx *c, int rank, int nranks, void *handles)
{
	if (nranks < 1 || nranks > SHARD_MAX_RANKS || rank < 0 || rank >= nranks) { set_error("shard_export: rank %d of %d (max %d ranks)", rank, nranks, SHARD_MAX_RANKS); return -1; }
	CMCB_CUDA_OK(cudaSetDevice(c->device));
	if (join_uploads(c)) return -1;
	c->rank = rank;
	c->nranks = nranks;
	c->keys_stride = c->n_star_dim;
	if (!c->d_keys_all) {
		CMCB_CUDA_OK(cudaMalloc((void **) &c->d_keys_all, (size_t) nranks * c->keys_stride * sizeof(uint64_t)));
		CMCB_CUDA_OK(cudaMalloc((void **) &c->d_src_all, (size_t) c->n_star_dim));
		CMCB_CUDA_OK(cudaMalloc((void **) &c->d_dest_pos, (size_t) c->n_star_dim * sizeof(unsigned)));
		CMCB_CUDA_OK(cudaMalloc((void **) &c->d_dest_rank, (size_t) c->n_star_dim));
		CMCB_CUDA_OK(cudaMalloc((void **) &c->d_shard_meta, sizeof(ShardMeta) + 4 * SHARD_MAX_RANKS * sizeof(void *)));
		CMCB_CUDA_OK(cudaMalloc((void **) &c->d_rank_all, (size_t) nranks * c->keys_stride * sizeof(unsigned)));
		CMCB_CUDA_OK(cudaMalloc((void **) &c->d_m_all, (size_t) nranks * c->keys_stride * sizeof(double)));
		CMCB_CUDA_OK(cudaStreamCreateWithFlags(&c->s_aux, cudaStreamNonBlocking));
		CMCB_CUDA_OK(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
		CMCB_CUDA_OK(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
		CMCB_CUDA_OK(cudaMalloc(&c->d_shard_sync, 1 << 21)); // an allocation of its own (2 MB granule): its IPC handle maps nothing else
		CMCB_CUDA_OK(cudaMemset(c->d_shard_sync, 0, sizeof(ShardSync)));
		c->shard_epoch = 0;
	}
	cudaIpcMemHandle_t *h = (cudaIpcMemHandle_t *) handles;
	for (int b = 0; b < SH_N_BUFFERS; b++) {
		c->own_buf[b] = own_buffer(c, b);
		CMCB_CUDA_OK(cudaIpcGetMemHandle(&h[b], c->own_buf[b]));
	}
	return 0;
}

extern "C" int cmcb200_shard_attach(cmcb200_ctx *c, const void *all_handles)
{
	CMCB_CUDA_OK(cudaSetDevice(c->device));
	if (join_uploads(c)) return -1;
	const cudaIpcMemHandle_t *h = (const cudaIpcMemHandle_t *) all_handles;
	for (int q = 0; q < c->nranks; q++)
		for (int b = 0; b < SH_N_BUFFERS; b++) {
			if (q == c->rank) { c->peer_buf[q][b] = c->own_buf[b]; continue; }
			CMCB_CUDA_OK(cudaIpcOpenMemHandle(&c->peer_buf[q][b], h[q * SH_N_BUFFERS + b], cudaIpcMemLazyEnablePeerAccess));
		}
	return 0;
}

extern "C" int cmcb200_shard_detach(cmcb200_ctx *c)
{
	CMCB_CUDA_OK(cudaSetDevice(c->device));
	if (join_uploads(c)) return -1;
	CMCB_CUDA_OK(cudaStreamSynchronize(c->stream));
	for (int q = 0; q < c->nranks; q++)
		for (int b = 0; b < SH_N_BUFFERS; b++)
			if (q != c->rank && c->peer_buf[q][b]) { cudaIpcCloseMemHandle(c->peer_buf[q][b]); c->peer_buf[q][b] = nullptr; }
	c->nranks = 1;
	c->rank = 0;
	return 0;
}

// the physical buffer of rank q that currently plays `role` (the r / m tables and the star arrays ping-pong in lock step on all ranks)
static void *peer_role(cmcb200_ctx *c, int q, int role)
{
	switch (role) {
	case SH_R: return c->peer_buf[q][c->r_parity ? SH_R_ALT : SH_R];
	case SH_M: return c->peer_buf[q][c->r_parity ? SH_M_ALT : SH_M];
	case SH_R_ALT: return c->peer_buf[q][c->r_parity ? SH_R : SH_R_ALT];
	case SH_M_ALT: return c->peer_buf[q][c->r_parity ? SH_M : SH_M_ALT];
	default: return c->peer_buf[q][role];
	}
}

// fused = 1 (default): the sort key is rnew and the ComputeIntermediateEnergy transfer rides on the exchange (resident pipeline);
// fused = 0: the key is r and the records move unchanged -- qsorts_new as the reference calls it, after its own transfer
extern "C" int cmcb200_shard_fuse_transfer(cmcb200_ctx *c, int fused)
{
	c->shard_unfused = fused ? 0 : 1;
	return 0;
}

// Step 1 + 2: local radix sort on rnew (sorted keys kept), trailing stripped stars trimmed (*n_keep = stars this rank contributes),
// then the sorted keys are stored into every rank's key table.  The host layer follows with a barrier.
extern "C" int cmcb200_shard_sort_local(cmcb200_ctx *c, long n_local, long *n_keep)
{
	if (n_local + 2 > c->n_star_dim) { set_error("shard_sort_local: n_local + 2 = %ld exceeds n_star_dim = %ld", n_local + 2, c->n_star_dim); return -1; }
	CMCB_CUDA_OK(cudaSetDevice(c->device));
	if (join_uploads(c)) return -1;
	SortArgs a;
	memset(&a, 0, sizeof(a));
	a.star_in = star1(c);
	a.n = n_local;
	a.key_off = c->shard_unfused ? CMCB_OFF_R : CMCB_OFF_RNEW;
	a.buf = c->sbuf;
	a.st = c->d_sort;
	a.tile_hist = c->d_tile_hist;
	a.keep_keys = true;
	a.keys_ready = false;
	{ StageTimer t(c, ST_KEYS); c->launches += sort_stage_keys(a, c->stream); }
	{ StageTimer t(c, ST_PASSES); c->launches += sort_stage_passes(a, c->stream); }
	{
		StageTimer t(c, ST_SHARD_KEYS);
		uint64_t *peers[SHARD_MAX_RANKS];
		for (int q = 0; q < c->nranks; q++) peers[q] = (uint64_t *) c->peer_buf[q][SH_KEYS_ALL];
		uint64_t **d_peers = (uint64_t **) ((char *) c->d_shard_meta + sizeof(ShardMeta));
		CMCB_CUDA_OK(cudaMemcpyAsync(d_peers, peers, c->nranks * sizeof(void *), cudaMemcpyHostToDevice, c->stream));
		if (n_local > 0) {
			shard_push_keys<<<NUM_SMS_B200 * 4, 256, 0, c->stream>>>(c->sbuf, c->d_sort, d_peers, c->nranks, (long) c->rank * c->keys_stride);
			c->launches++;
		}
	}
	CMCB_CUDA_OK(cudaGetLastError());
	long *h = (long *) c->h_pinned; // n_eq_inf, n_gt_inf, n_total, n_new, n_bin
	CMCB_CUDA_OK(cudaMemcpyAsync(h, &c->d_sort->n_eq_inf, 5 * sizeof(long), cudaMemcpyDeviceToHost, c->stream));
	CMCB_CUDA_OK(cudaStreamSynchronize(c->stream));
	*n_keep = h[3];
	return 0;
}

// Step 3 + 4 (after the barrier): global ranks, then every record goes to its final slot on its destination rank.
//   n_keep_all[q]   what every rank's cmcb200_shard_sort_local returned
//   Start / End     findLimits(sum n_keep_all, MIN_CHUNK_SIZE, nranks): the NEW partition (1-based, inclusive)
//   n_old_local, first_global, N_MAX   the partition the orbit step ran on (for the fused r <- rnew transfer)
// The host layer follows with a barrier, then calls cmcb200_shard_publish.
extern "C" int cmcb200_shard_exchange(cmcb200_ctx *c, const long *n_keep_all, const int *Start, const int *End, long n_old_local, long first_global, long N_MAX)
{
	CMCB_CUDA_OK(cudaSetDevice(c->device));
	if (join_uploads(c)) return -1;
	ShardMeta m;
	memset(&m, 0, sizeof(m));
	m.nranks = c->nranks;
	m.rank = c->rank;
	for (int q = 0; q < c->nranks; q++) {
		m.n_keep[q] = n_keep_all[q];
		m.start0[q] = Start[q] - 1;
		m.count[q] = End[q] - Start[q] + 1;
		if (m.count[q] + 2 > c->n_star_dim) { set_error("shard_exchange: rank %d would hold %ld stars, capacity %ld", q, m.count[q], c->n_star_dim); return -1; }
	}
	CMCB_CUDA_OK(cudaMemcpyAsync(c->d_shard_meta, &m, sizeof(m), cudaMemcpyHostToDevice, c->stream));
	const long n = n_keep_all[c->rank];
	{
		StageTimer t(c, ST_SHARD_RANK);
		if (n > 0) {
			long blocks = (n + 255) / 256;
			if (blocks > NUM_SMS_B200 * 16) blocks = NUM_SMS_B200 * 16;
			shard_rank_kernel<<<(unsigned) blocks, 256, 0, c->stream>>>(c->d_keys_all, c->keys_stride, (const ShardMeta *) c->d_shard_meta, c->d_dest_rank, c->d_dest_pos);
			c->launches++;
		}
	}
	{
		StageTimer t(c, ST_GATHER);
		PushArgs a;
		a.in = star1(c);
		for (int q = 0; q < c->nranks; q++) {
			a.out[q] = (char *) c->peer_buf[q][SH_STAR0 + (c->cur_star ^ 1)] + CMCB_SIZEOF_STAR;
			a.src_out[q] = (unsigned char *) c->peer_buf[q][SH_SRC_ALL];
		}
		a.rank = c->rank;
		a.old_r = c->d_r; a.old_m = c->d_m;
		a.n_old_local = n_old_local; a.first_global = first_global; a.N_MAX = N_MAX;
		// the sentinel record star[0] stays with its rank
		CMCB_CUDA_OK(cudaMemcpyAsync(c->d_star[c->cur_star ^ 1], c->d_star[c->cur_star], CMCB_SIZEOF_STAR, cudaMemcpyDeviceToDevice, c->stream));
		if (n > 0) {
			if (c->shard_unfused) shard_push_records<false><<<NUM_SMS_B200 * 8, 256, 0, c->stream>>>(a, c->sbuf, c->d_sort, c->d_dest_rank, c->d_dest_pos);
			else shard_push_records<true><<<NUM_SMS_B200 * 8, 256, 0, c->stream>>>(a, c->sbuf, c->d_sort, c->d_dest_rank, c->d_dest_pos);
			c->launches++;
		}
	}
	CMCB_CUDA_OK(cudaGetLastError());
	return 0;
}

// Step 5 (after the barrier): this rank now holds `n_new_local` sorted stars starting at global index start_global (1-based); their
// r, m go into every rank's alternate tables.  The host layer follows with a barrier, then calls cmcb200_shard_finish.
extern "C" int cmcb200_shard_publish(cmcb200_ctx *c, long n_new_local, long start_global, long n_bin_clear, long *n_bin)
{
	CMCB_CUDA_OK(cudaSetDevice(c->device));
	if (join_uploads(c)) return -1;
	c->cur_star ^= 1;
	StageTimer t(c, ST_SHARD_TABLES);
	void *ptrs[2 * SHARD_MAX_RANKS];
	for (int q = 0; q < c->nranks; q++) { ptrs[q] = peer_role(c, q, SH_R_ALT); ptrs[SHARD_MAX_RANKS + q] = peer_role(c, q, SH_M_ALT); }
	void **d_ptrs = (void **) ((char *) c->d_shard_meta + sizeof(ShardMeta) + SHARD_MAX_RANKS * sizeof(void *));
	CMCB_CUDA_OK(cudaMemcpyAsync(d_ptrs, ptrs, sizeof(ptrs), cudaMemcpyHostToDevice, c->stream));
	// the table sentinel [0] stays
	CMCB_CUDA_OK(cudaMemcpyAsync(c->d_r_alt, c->d_r, sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
	CMCB_CUDA_OK(cudaMemcpyAsync(c->d_m_alt, c->d_m, sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
	if (n_new_local > 0) {
		long blocks = (n_new_local + 255) / 256;
		if (blocks > NUM_SMS_B200 * 8) blocks = NUM_SMS_B200 * 8;
		shard_publish_tables<<<(unsigned) blocks, 256, 0, c->stream>>>(star1(c), n_new_local, start_global - 1, (double *const *) d_ptrs,
									       (double *const *) (d_ptrs + SHARD_MAX_RANKS), c->nranks, c->d_binflag);
		c->launches++;
	}
	// binaries: count / scan as on one GPU (sort.cu), then pull each record from the rank its star came from
	*n_bin = 0;
	{
		StageTimer tb(c, ST_BINARIES);
		if (n_bin_clear > c->n_bin_dim - 1) n_bin_clear = c->n_bin_dim - 1;
		char *bin_out = c->d_bin[c->cur_bin ^ 1] + CMCB_SIZEOF_BINARY;
		c->launches += sort_binary_count_scan(c->d_binflag, c->d_sort, c->d_bin_block_cnt, n_new_local, c->stream);
		if (n_new_local > 0) {
			BinPull bp;
			for (int q = 0; q < c->nranks; q++) bp.bin_in[q] = (const char *) c->peer_buf[q][SH_BIN0 + c->cur_bin] + CMCB_SIZEOF_BINARY;
			const long nbt = (n_new_local + SHARD_BIN_TILE - 1) / SHARD_BIN_TILE;
			shard_binary_pull<<<(unsigned) nbt, 256, 0, c->stream>>>(star1(c), c->d_binflag, c->d_src_all, c->d_bin_block_cnt, n_new_local, nullptr, bp, bin_out);
			c->launches++;
		}
		c->launches += sort_binary_clear_device(bin_out, c->d_sort, n_bin_clear, c->stream);
		c->cur_bin ^= 1;
	}
	CMCB_CUDA_OK(cudaGetLastError());
	long *h = (long *) c->h_pinned;
	CMCB_CUDA_OK(cudaMemcpyAsync(h, &c->d_sort->n_bin, sizeof(long), cudaMemcpyDeviceToHost, c->stream));
	CMCB_CUDA_OK(cudaStreamSynchronize(c->stream));
	*n_bin = h[0];
	return 0;
}

// after the last barrier: the alternate tables become THE tables (on all ranks in lock step)
extern "C" int cmcb200_shard_finish(cmcb200_ctx *c)
{
	std::swap(c->d_r, c->d_r_alt);
	std::swap(c->d_m, c->d_m_alt);
	c->r_parity ^= 1;
	return 0;
}


// ---- the whole sharded sort as ONE asynchronous sequence ---------------------------------------------------------------------------------
// cmcb200_shard_sort_local + _exchange + _publish + _finish without a single host round trip: the ranks meet at three device-side barriers
// (ShardSync above), the counts every rank contributes travel with the first one, the new partition is planned on every device, and the
// kernels behind it take their extents from that plan.  The host only enqueues; cmcb200_shard_step_result waits and returns what the
// reference's globals need (clus.N_MAX, Start / End, N_b_local).  One host synchronisation per sort instead of three plus an allgather.
static int shard_post(cmcb200_ctx *c, int phase, const SortState *st)
{
	SyncPeers sp;
	for (int q = 0; q < c->nranks; q++) sp.p[q] = (ShardSync *) c->peer_buf[q][SH_SYNC];
	shard_post_kernel<<<1, 32, 0, c->stream>>>(sp, c->nranks, c->rank, phase, c->shard_epoch, st);
	c->launches++;
	return 0;
}
static int shard_wait(cmcb200_ctx *c, int phase)
{
	shard_wait_kernel<<<1, 32, 0, c->stream>>>((ShardSync *) c->d_shard_sync, c->nranks, phase, c->shard_epoch);
	c->launches++;
	return 0;
}

extern "C" int cmcb200_shard_step_async(cmcb200_ctx *c, long n_local, long n_old_local, long first_global, long N_MAX, int min_chunk, long n_bin_clear,
					      long pot_N_STAR, double pot_cenma_m, double pot_cenma_m_new)
{
	if (n_local + 2 > c->n_star_dim) { set_error("shard_step: n_local + 2 = %ld exceeds n_star_dim = %ld", n_local + 2, c->n_star_dim); return -1; }
	if (!c->d_shard_sync || c->nranks < 1) { set_error("shard_step: call cmcb200_shard_export / _attach first"); return -1; }
	if (N_MAX + 2 > c->n_table_dim) { set_error("shard_step: N_MAX + 2 exceeds the table capacity"); return -1; }
	CMCB_CUDA_OK(cudaSetDevice(c->device));
	if (join_uploads(c)) return -1;
	c->shard_epoch++;
	ShardMeta *d_meta = (ShardMeta *) c->d_shard_meta;
	// 1. local sort, sorted keys into every rank's key table, this rank's count with the first barrier
	SortArgs a;
	memset(&a, 0, sizeof(a));
	a.star_in = star1(c);
	a.n = n_local;
	a.key_off = c->shard_unfused ? CMCB_OFF_R : CMCB_OFF_RNEW;
	a.buf = c->sbuf;
	a.st = c->d_sort;
	a.tile_hist = c->d_tile_hist;
	a.keep_keys = true;
	a.keys_ready = false;
	{ StageTimer t(c, ST_KEYS); c->launches += sort_stage_keys(a, c->stream); }
	{ StageTimer t(c, ST_PASSES); c->launches += sort_stage_passes(a, c->stream); }
	{
		StageTimer t(c, ST_SHARD_KEYS);
		uint64_t *peers[SHARD_MAX_RANKS];
		for (int q = 0; q < c->nranks; q++) peers[q] = (uint64_t *) c->peer_buf[q][SH_KEYS_ALL];
		uint64_t **d_peers = (uint64_t **) ((char *) c->d_shard_meta + sizeof(ShardMeta));
		CMCB_CUDA_OK(cudaMemcpyAsync(d_peers, peers, c->nranks * sizeof(void *), cudaMemcpyHostToDevice, c->stream));
		if (n_local > 0) {
			shard_push_keys<<<NUM_SMS_B200 * 4, 256, 0, c->stream>>>(c->sbuf, c->d_sort, d_peers, c->nranks, (long) c->rank * c->keys_stride);
			c->launches++;
		}
		shard_post(c, 0, c->d_sort);
		shard_wait(c, 0);
		shard_plan_kernel<<<1, 32, 0, c->stream>>>((const ShardSync *) c->d_shard_sync, d_meta, c->nranks, c->rank, min_chunk);
		c->launches++;
	}
	PushArgs pa;
	pa.in = star1(c);
	for (int q = 0; q < c->nranks; q++) {
		pa.out[q] = (char *) c->peer_buf[q][SH_STAR0 + (c->cur_star ^ 1)] + CMCB_SIZEOF_STAR;
		pa.src_out[q] = (unsigned char *) c->peer_buf[q][SH_SRC_ALL];
	}
	pa.rank = c->rank;
	pa.old_r = c->d_r; pa.old_m = c->d_m;
	pa.n_old_local = n_old_local; pa.first_global = first_global; pa.N_MAX = N_MAX;
	// 2. global ranks; (global rank, star_m) of this run to every rank
	{
		StageTimer t(c, ST_SHARD_RANK);
		if (n_local > 0) {
			long blocks = (n_local + 255) / 256;
			if (blocks > NUM_SMS_B200 * 16) blocks = NUM_SMS_B200 * 16;
			shard_rank_kernel<<<(unsigned) blocks, 256, 0, c->stream>>>(c->d_keys_all, c->keys_stride, d_meta, c->d_dest_rank, c->d_dest_pos);
			RankPeers rp;
			for (int q = 0; q < c->nranks; q++) { rp.rank[q] = (unsigned *) c->peer_buf[q][SH_RANK_ALL]; rp.m[q] = (double *) c->peer_buf[q][SH_M_ALL]; }
			if (c->shard_unfused) shard_publish_ranks<false><<<(unsigned) blocks, 256, 0, c->stream>>>(pa, c->sbuf, c->d_sort, d_meta, c->d_dest_rank, c->d_dest_pos, rp, (long) c->rank * c->keys_stride);
			else shard_publish_ranks<true><<<(unsigned) blocks, 256, 0, c->stream>>>(pa, c->sbuf, c->d_sort, d_meta, c->d_dest_rank, c->d_dest_pos, rp, (long) c->rank * c->keys_stride);
			c->launches += 2;
		}
		shard_post(c, 4, nullptr);
	}
	// 3. second stream: once every rank's (rank, m) arrays are here, the new tables are scattered locally and the potential is rebuilt --
	//    while the main stream moves the records
	CMCB_CUDA_OK(cudaEventRecord(c->ev_fork, c->stream));
	CMCB_CUDA_OK(cudaStreamWaitEvent(c->s_aux, c->ev_fork, 0));
	{
		cudaStream_t keep = c->stream;
		c->stream = c->s_aux;
		{
			StageTimer t(c, ST_SHARD_TABLES);
			shard_wait(c, 4);
			CMCB_CUDA_OK(cudaMemcpyAsync(c->d_r_alt, c->d_r, sizeof(double), cudaMemcpyDeviceToDevice, c->s_aux)); // the table sentinel [0] stays
			CMCB_CUDA_OK(cudaMemcpyAsync(c->d_m_alt, c->d_m, sizeof(double), cudaMemcpyDeviceToDevice, c->s_aux));
			long blocks = (c->keys_stride + 255) / 256;
			if (blocks > NUM_SMS_B200 * 2) blocks = NUM_SMS_B200 * 2;
			shard_scatter_tables<<<dim3((unsigned) blocks, (unsigned) c->nranks), 256, 0, c->s_aux>>>(c->d_keys_all, c->d_rank_all, c->d_m_all, c->keys_stride, d_meta, c->d_r_alt, c->d_m_alt);
			c->launches++;
		}
		c->stream = keep;
	}
	// the alternate tables become THE tables (on all ranks in lock step); the fused push below still reads the OLD ones (pa.old_r / old_m)
	std::swap(c->d_r, c->d_r_alt);
	std::swap(c->d_m, c->d_m_alt);
	c->r_parity ^= 1;
	c->pot_queued = pot_N_STAR > 0;
	if (pot_N_STAR > 0 && potential_enqueue(c, N_MAX, pot_N_STAR, pot_cenma_m, pot_cenma_m_new, c->s_aux)) return -1;
	CMCB_CUDA_OK(cudaEventRecord(c->ev_join, c->s_aux));
	// 4. main stream: every record to its final slot on its destination rank
	{
		StageTimer t(c, ST_GATHER);
		CMCB_CUDA_OK(cudaMemcpyAsync(c->d_star[c->cur_star ^ 1], c->d_star[c->cur_star], CMCB_SIZEOF_STAR, cudaMemcpyDeviceToDevice, c->stream));
		if (n_local > 0) {
			if (c->shard_unfused) shard_push_records<false><<<NUM_SMS_B200 * 8, 256, 0, c->stream>>>(pa, c->sbuf, c->d_sort, c->d_dest_rank, c->d_dest_pos);
			else shard_push_records<true><<<NUM_SMS_B200 * 8, 256, 0, c->stream>>>(pa, c->sbuf, c->d_sort, c->d_dest_rank, c->d_dest_pos);
			c->launches++;
		}
		shard_post(c, 1, nullptr);
		shard_wait(c, 1);
	}
	// 5. the shard this rank now holds: its binaries pulled from the ranks their stars came from
	c->cur_star ^= 1;
	{
		StageTimer tb(c, ST_BINARIES);
		if (n_bin_clear > c->n_bin_dim - 1) n_bin_clear = c->n_bin_dim - 1;
		char *bin_out = c->d_bin[c->cur_bin ^ 1] + CMCB_SIZEOF_BINARY;
		shard_adopt_shard_kernel<<<1, 32, 0, c->stream>>>(c->d_sort, d_meta);
		shard_binflag_kernel<<<NUM_SMS_B200 * 4, 256, 0, c->stream>>>(star1(c), d_meta, c->d_binflag);
		c->launches += 2;
		const long n_upper = c->n_star_dim; // the shard may be larger than what this rank contributed: size the grids for the capacity
		const long nbt = (n_upper + SHARD_BIN_TILE - 1) / SHARD_BIN_TILE;
		c->launches += sort_binary_count_scan_planned(c->d_binflag, c->d_sort, c->d_bin_block_cnt, n_upper, c->stream);
		BinPull bp;
		for (int q = 0; q < c->nranks; q++) bp.bin_in[q] = (const char *) c->peer_buf[q][SH_BIN0 + c->cur_bin] + CMCB_SIZEOF_BINARY;
		shard_binary_pull<<<(unsigned) nbt, 256, 0, c->stream>>>(star1(c), c->d_binflag, c->d_src_all, c->d_bin_block_cnt, 0, c->d_sort, bp, bin_out);
		c->launches++;
		c->launches += sort_binary_clear_device(bin_out, c->d_sort, n_bin_clear, c->stream);
		c->cur_bin ^= 1;
		shard_post(c, 2, nullptr);
		shard_wait(c, 2);
	}
	// 6. join: whatever is queued next on the context's stream finds the new tables and the new potential
	CMCB_CUDA_OK(cudaStreamWaitEvent(c->stream, c->ev_join, 0));
	CMCB_CUDA_OK(cudaGetLastError());
	return 0;
}

extern "C" int cmcb200_shard_step_result(cmcb200_ctx *c, long *n_keep_all, int *Start, int *End, long *n_bin, double pot_out[4])
{
	CMCB_CUDA_OK(cudaSetDevice(c->device));
	char *h = (char *) c->h_pinned;
	if (c->pot_queued) CMCB_CUDA_OK(cudaMemcpyAsync(h + 2048, c->d_pot, sizeof(PotScalars), cudaMemcpyDeviceToHost, c->stream));
	CMCB_CUDA_OK(cudaMemcpyAsync(h, c->d_shard_meta, sizeof(ShardMeta), cudaMemcpyDeviceToHost, c->stream));
	CMCB_CUDA_OK(cudaMemcpyAsync(h + sizeof(ShardMeta), &c->d_sort->n_bin, sizeof(long), cudaMemcpyDeviceToHost, c->stream));
	CMCB_CUDA_OK(cudaMemcpyAsync(h + sizeof(ShardMeta) + 8, &((ShardSync *) c->d_shard_sync)->error, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
	CMCB_CUDA_OK(cudaStreamSynchronize(c->stream));
	const ShardMeta *m = (const ShardMeta *) h;
	const int err = *(const int *) (h + sizeof(ShardMeta) + 8);
	if (err) { set_error("shard_step: barrier %d timed out -- a peer rank did not arrive", err - 1); return -2; }
	for (int q = 0; q < c->nranks; q++) {
		if (n_keep_all) n_keep_all[q] = m->n_keep[q];
		if (Start) Start[q] = (int) (m->start0[q] + 1);
		if (End) End[q] = (int) (m->start0[q] + m->count[q]);
		if (m->count[q] + 2 > c->n_star_dim) { set_error("shard_step: rank %d holds %ld stars, capacity %ld", q, m->count[q], c->n_star_dim); return -1; }
	}
	if (n_bin) *n_bin = *(const long *) (h + sizeof(ShardMeta));
	if (pot_out && c->pot_queued) { // N_MAX, cenma.m, cenma.m_new, Mtotal of the rebuild queued with the step (the caller: Rtidal = orbit_r Mtotal^(1/3), cmc_utils.c:707)
		const PotScalars *ps = (const PotScalars *) (h + 2048);
		pot_out[0] = (double) ps->n_max; pot_out[1] = ps->cenma_m; pot_out[2] = ps->cenma_m_new; pot_out[3] = ps->Mtotal;
	}
	return 0;
}


int shard_share_table_slices(cmcb200_ctx *c, long first, long n)
{
	if (c->nranks <= 1) return 0;
	TablePeers tp;
	for (int q = 0; q < c->nranks; q++) {
		tp.r[q] = (double *) peer_role(c, q, SH_R);
		tp.m[q] = (double *) peer_role(c, q, SH_M);
		tp.phi[q] = (double *) c->peer_buf[q][SH_PHI];
	}
	if (n > 0) {
		long blocks = (n + 255) / 256;
		if (blocks > NUM_SMS_B200 * 8) blocks = NUM_SMS_B200 * 8;
		shard_share_tables_kernel<<<(unsigned) blocks, 256, 0, c->stream>>>(c->d_r, c->d_m, c->d_phi, first, n, tp, c->nranks, c->rank);
		c->launches++;
	}
	c->table_epoch++;
	SyncPeers sp;
	for (int q = 0; q < c->nranks; q++) sp.p[q] = (ShardSync *) c->peer_buf[q][SH_SYNC];
	shard_post_kernel<<<1, 32, 0, c->stream>>>(sp, c->nranks, c->rank, 3, c->table_epoch, nullptr);
	shard_wait_kernel<<<1, 32, 0, c->stream>>>((ShardSync *) c->d_shard_sync, c->nranks, 3, c->table_epoch);
	c->launches += 2;
	CMCB_CUDA_OK(cudaGetLastError());
	return 0;
}

// The per-star new-orbit step on the device: get_positions / get_positions_loop
// (reference src/cmc/cmc_evolution_thr.c:784-1023).  The per-star arithmetic is orbit_core.h (shared with the
// host-side emulation the CPU tests run); this file is the data movement and the work distribution around it.
//
//   orbit_pack_tables       (r_k, Phi_k) -> one 16-byte entry: every bisection probe needs both, so a probe costs one
//                           32-byte sector instead of two
//   orbit_step_kernel       one thread per star, consecutive threads = consecutive stars in radius order (the first probes
//                           of the index bisections of neighbours fall into the same sectors).  Probes are division-free
//                           FP64 sign tests with a rigorous bound, exact x87 replay inside the bound (cold, not inlined).
//                           A star whose own Q is <= 0 would run get_positive_Q_index's O(N) linear scan
//                           (cmc_orbit.c:46-57) in ONE thread while the rest of the GPU waits: it is queued instead.
//   orbit_scan_kernel       the queued linear scans as a grid-wide find-first: work items (chunk of 4096 table entries,
//                           queued star) are handed out chunk-major so all SMs work on the lowest unresolved chunks; a hit
//                           retires the later chunks of that star.
//   orbit_deferred_kernel   finishes the queued stars with the index the scan found.
// The r/Phi table is L2-resident up to N ~ 7e6 (16 B per star, 126 MB L2); beyond that the bottom levels of every bisection
// are DRAM sector reads.  Bound: dependent-load latency x occupancy, then L2/DRAM sector bandwidth -- not HBM streaming.
#include "orbit.cuh"

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

struct CsrDraws {
	const double *d;
	long pos, end;
	__device__ bool has(int n) const { return pos + n <= end; }
	__device__ double next() { return d[pos++]; }
};

__device__ __forceinline__ double ld_field(const char *rec, int off) { return *(const double *) (rec + off); }
__device__ __forceinline__ void st_field(char *rec, int off, double v) { *(double *) (rec + off) = v; }

__global__ void orbit_pack_tables(const double *__restrict__ r, const double *__restrict__ phi, RPhi *__restrict__ rp, long n)
{
	long i = blockIdx.x * (long) blockDim.x + threadIdx.x;
	const long stride = (long) gridDim.x * blockDim.x;
	for (; i < n; i += stride) {
		double2 v;
		v.x = r[i];
		v.y = phi[i];
		*reinterpret_cast<double2 *>(rp + i) = v;
	}
}

struct StarIn {
	long g;
	bool is_new;
	double E, J, vr, vt, rnew_of_new;
};

__device__ __forceinline__ StarIn load_star(const OrbitArgs &a, long j0)
{
	StarIn in;
	const long j = j0 + 1; // the reference's 1-based local index
	in.is_new = j > a.n_old_local;
	in.g = in.is_new ? 1 + a.t.N_MAX + (j - a.n_old_local) : a.first_global + j - 1; // get_global_idx, cmc_utils.c:2384-2407
	const char *rec = a.star1 + j0 * CMCB_SIZEOF_STAR;
	in.E = ld_field(rec, CMCB_OFF_E); in.J = ld_field(rec, CMCB_OFF_J);
	in.vr = ld_field(rec, CMCB_OFF_VR); in.vt = ld_field(rec, CMCB_OFF_VT);
	in.rnew_of_new = in.is_new ? ld_field(rec, CMCB_OFF_RNEW) : 0.0;
	return in;
}

// what the loop body writes into star[j] (cmc_evolution_thr.c:835-839, 852-853, 918-931; remove_star :662-664)
__device__ __forceinline__ void store_star(const OrbitArgs &a, long j0, const OrbitOut &o)
{
	char *rec = a.star1 + j0 * CMCB_SIZEOF_STAR;
	if (o.status != ORBIT_NTRY) {
		st_field(rec, CMCB_OFF_RNEW, o.rnew);
		st_field(rec, CMCB_OFF_VRNEW, o.vrnew);
		st_field(rec, CMCB_OFF_VTNEW, o.vtnew);
	}
	if (o.status == ORBIT_OK) st_field(rec, CMCB_OFF_X, o.X);
	if (o.status == ORBIT_OK || o.status == ORBIT_TIDAL) {
		st_field(rec, CMCB_OFF_R_PERI, o.r_peri);
		st_field(rec, CMCB_OFF_R_APO, o.r_apo);
	}
	if (a.status) a.status[j0] = o.status;
	if (a.kmin) a.kmin[j0] = o.kmin;
	if (a.kmax) a.kmax[j0] = o.kmax;
	if (a.n_draws) a.n_draws[j0] = o.n_draws;
}

template <bool RECORDED>
__device__ __forceinline__ void run_star(const OrbitArgs &a, long j0, const StarIn &in, long ktemp_in, bool allow_defer, OrbitOut &o)
{
	if (RECORDED) {
		CsrDraws dr;
		dr.d = a.draws; dr.pos = a.draw_off[j0]; dr.end = a.draw_off[j0 + 1];
		orbit_step_star(a.t, in.g, in.is_new, in.rnew_of_new, in.E, in.J, in.vr, in.vt, a.Rtidal, dr, o, ktemp_in, allow_defer);
	} else {
		PhiloxDraws dr;
		dr.init(a.seed, a.tcount, (unsigned long long) in.g);
		orbit_step_star(a.t, in.g, in.is_new, in.rnew_of_new, in.E, in.J, in.vr, in.vt, a.Rtidal, dr, o, ktemp_in, allow_defer);
	}
}

// block-level accumulation of the per-launch statistics: one atomic per warp / block and counter
__device__ __forceinline__ void accumulate_stats(OrbitStats *st, const OrbitOut &o, bool valid)
{
	unsigned long long my_max = (valid && o.status == ORBIT_OK) ? dbits(o.rnew) : 0ull; // positive doubles order like their bit patterns
	unsigned long long my_draws = valid ? o.n_draws : 0, my_exact = valid ? o.n_exact : 0, my_brent = valid ? o.n_brent : 0;
#pragma unroll
	for (int d = 16; d > 0; d >>= 1) {
		unsigned long long m = __shfl_xor_sync(0xffffffffu, my_max, d);
		my_max = m > my_max ? m : my_max;
		my_draws += __shfl_xor_sync(0xffffffffu, my_draws, d);
		my_exact += __shfl_xor_sync(0xffffffffu, my_exact, d);
		my_brent += __shfl_xor_sync(0xffffffffu, my_brent, d);
	}
	if (lane_id() == 0) {
		if (my_max) atomicMax(&st->max_rad_bits, my_max);
		if (my_draws) atomicAdd(&st->n_draws, my_draws);
		if (my_exact) atomicAdd(&st->n_exact, my_exact);
		if (my_brent) atomicAdd(&st->n_brent, my_brent);
	}
	if (valid && o.status > 0 && o.status < 6) atomicAdd(&st->n_status[o.status], 1ull); // the rare outcomes; OK count = n - sum
}

// MINB = resident CTAs per SM the register allocation is capped for (4: 128 registers .. 8: 64 registers, with spills)
template <bool RECORDED, int MINB>
__global__ void __launch_bounds__(ORBIT_THREADS, MINB) orbit_step_kernel(OrbitArgs a)
{
	const long j0 = blockIdx.x * (long) blockDim.x + threadIdx.x;
	const bool valid = j0 < a.n_local;
	OrbitOut o;
	o.status = -1; o.rnew = 0.0; o.n_draws = o.n_exact = o.n_brent = 0;
	if (valid) {
		const StarIn in = load_star(a, j0);
		run_star<RECORDED>(a, j0, in, KTEMP_COMPUTE, true, o);
		if (o.status == ORBIT_DEFERRED) {
			const unsigned slot = atomicAdd(&a.stats->n_defer, 1u);
			if (slot < ORBIT_DEFER_CAP) {
				a.defer_star[slot] = j0;
				a.defer_found[slot] = ORBIT_NOT_FOUND;
			} else { // queue full (never seen: ~4e-6 of the stars defer): last resort, the scan in this thread
				run_star<RECORDED>(a, j0, in, KTEMP_COMPUTE, false, o);
			}
		}
		if (o.status != ORBIT_DEFERRED) store_star(a, j0, o);
	}
	accumulate_stats(a.stats, o, valid && o.status != ORBIT_DEFERRED);
}

// ---- the queued linear scans ------------------------------------------------------------------------------------------

// first k in [0, N_MAX) with function_q(k) > 0 for every queued star (get_positive_Q_index, cmc_orbit.c:46-57: the loop runs
// k = 0, 1, .. while Q <= 0 and k <= N_MAX, and a result >= N_MAX means "none")
__global__ void __launch_bounds__(ORBIT_SCAN_THREADS) orbit_scan_kernel(OrbitArgs a)
{
	__shared__ unsigned s_item;
	__shared__ long s_best;
	const unsigned n_defer = min(a.stats->n_defer, (unsigned) ORBIT_DEFER_CAP);
	if (n_defer == 0) return;
	const long n_scan = a.t.N_MAX; // candidates 0 .. N_MAX-1
	const long n_chunks = (n_scan + ORBIT_SCAN_CHUNK - 1) / ORBIT_SCAN_CHUNK;
	const unsigned long long n_items = (unsigned long long) n_chunks * n_defer;
	for (;;) {
		__syncthreads();
		if (threadIdx.x == 0) {
			// One thread draws tickets until it holds an item that still matters: a chunk lying entirely above a hit that already
			// exists for its star cannot contribute (hits only ever decrease), and deciding that here keeps the block convergent.
			unsigned long long it;
			do {
				it = atomicAdd(&a.stats->scan_ticket, 1u);
			} while (it < n_items && *(volatile long *) &a.defer_found[it % n_defer] < (long) (it / n_defer) * ORBIT_SCAN_CHUNK);
			s_item = it < n_items ? (unsigned) it : 0xffffffffu;
			s_best = ORBIT_NOT_FOUND;
		}
		__syncthreads();
		if (s_item == 0xffffffffu) return;
		const unsigned long long item = s_item;
		const long chunk = (long) (item / n_defer);
		const unsigned slot = (unsigned) (item % n_defer);
		const long k0 = chunk * ORBIT_SCAN_CHUNK;
		const long j0 = a.defer_star[slot];
		const StarIn in = load_star(a, j0);
		const StarConst s = star_const(a.t, in.g, in.E, in.J);
		long best = ORBIT_NOT_FOUND;
		int n_exact = 0;
#pragma unroll 4
		for (int i = 0; i < ORBIT_SCAN_CHUNK / ORBIT_SCAN_THREADS; i++) {
			const long k = k0 + i * ORBIT_SCAN_THREADS + threadIdx.x;
			if (k < n_scan && k < best) {
				if (sign_Q(a.t, s, k, &n_exact) > 0) best = k;
			}
		}
#pragma unroll
		for (int d = 16; d > 0; d >>= 1) {
			long o = __shfl_xor_sync(0xffffffffu, best, d);
			best = o < best ? o : best;
			n_exact += __shfl_xor_sync(0xffffffffu, n_exact, d);
		}
		if (lane_id() == 0) {
			if (best != ORBIT_NOT_FOUND) atomicMin((long long *) &s_best, (long long) best);
			if (n_exact) atomicAdd(&a.stats->n_exact, (unsigned long long) n_exact);
		}
		__syncthreads();
		if (threadIdx.x == 0 && s_best != ORBIT_NOT_FOUND) atomicMin((long long *) &a.defer_found[slot], (long long) s_best);
	}
}

template <bool RECORDED>
__global__ void __launch_bounds__(ORBIT_THREADS) orbit_deferred_kernel(OrbitArgs a)
{
	const unsigned n_defer = min(a.stats->n_defer, (unsigned) ORBIT_DEFER_CAP);
	for (unsigned base = blockIdx.x * blockDim.x; base < n_defer; base += gridDim.x * blockDim.x) { // block-uniform trip count
		const unsigned slot = base + threadIdx.x;
		const bool valid = slot < n_defer;
		OrbitOut o;
		o.status = -1; o.rnew = 0.0; o.n_draws = o.n_exact = o.n_brent = 0;
		if (valid) {
			const long j0 = a.defer_star[slot];
			const long found = a.defer_found[slot];
			const StarIn in = load_star(a, j0);
			run_star<RECORDED>(a, j0, in, found == ORBIT_NOT_FOUND ? -1 : found, false, o);
			store_star(a, j0, o);
		}
		accumulate_stats(a.stats, o, valid);
	}
}

__global__ void orbit_stats_reset(OrbitStats *s)
{
	s->max_rad_bits = 0; s->n_draws = 0; s->n_exact = 0; s->n_brent = 0;
	for (int i = 0; i < 6; i++) s->n_status[i] = 0;
	s->n_defer = 0; s->scan_ticket = 0;
}

int orbit_pack_tables_device(const double *r, const double *phi, RPhi *rp, long n, cudaStream_t s)
{
	if (n <= 0) return 0;
	long blocks = (n + 255) / 256;
	if (blocks > NUM_SMS_B200 * 16) blocks = NUM_SMS_B200 * 16;
	orbit_pack_tables<<<(unsigned) blocks, 256, 0, s>>>(r, phi, rp, n);
	return 1;
}

int orbit_step_device(const OrbitArgs &a, cudaStream_t s)
{
	orbit_stats_reset<<<1, 1, 0, s>>>(a.stats);
	if (a.n_local <= 0) return 1;
	const unsigned blocks = (unsigned) ((a.n_local + ORBIT_THREADS - 1) / ORBIT_THREADS);
	const int minb = a.min_blocks >= 8 ? 8 : (a.min_blocks >= 6 ? 6 : 4);
#define CMCB_LAUNCH_STEP(REC, MB) orbit_step_kernel<REC, MB><<<blocks, ORBIT_THREADS, 0, s>>>(a)
	if (a.draws) { if (minb == 8) CMCB_LAUNCH_STEP(true, 8); else if (minb == 6) CMCB_LAUNCH_STEP(true, 6); else CMCB_LAUNCH_STEP(true, 4); }
	else { if (minb == 8) CMCB_LAUNCH_STEP(false, 8); else if (minb == 6) CMCB_LAUNCH_STEP(false, 6); else CMCB_LAUNCH_STEP(false, 4); }
#undef CMCB_LAUNCH_STEP
	orbit_scan_kernel<<<NUM_SMS_B200 * 4, ORBIT_SCAN_THREADS, 0, s>>>(a);
	if (a.draws) orbit_deferred_kernel<true><<<NUM_SMS_B200, ORBIT_THREADS, 0, s>>>(a);
	else orbit_deferred_kernel<false><<<NUM_SMS_B200, ORBIT_THREADS, 0, s>>>(a);
	return 4;
}

} // namespace cmcb

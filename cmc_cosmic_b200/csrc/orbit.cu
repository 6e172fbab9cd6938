// The per-star new-orbit step on the device: get_positions / get_positions_loop
// (reference src/cmc/cmc_evolution_thr.c:784-1023).  The per-star arithmetic is orbit_core.h (shared with the
// host-side emulation the CPU tests run); this file is the data movement and the work distribution around it.
//
//   orbit_pack_tables       (r_k, Phi_k) -> one 16-byte entry: every bisection probe needs both, so a probe costs one
//                           32-byte sector instead of two; the same pass emits the finest level of the Q ladder and
//                           verifies that star_r is sorted.  qlad_coarsen / rlad_build: the other ladder levels (orbit_core.h)
//   orbit_prepare_kernel    persistent CTAs, the coarsest Q-ladder level staged in shared memory by one TMA bulk copy per CTA;
//                           one thread per star, consecutive threads = consecutive stars in radius order.  Probes of the index
//                           bisections are answered from the ladder where a block bound proves the sign, else by the
//                           division-free FP64 sign test with a rigorous bound on the fine table, exact x87 replay inside
//                           the bound (cold, not inlined).
//                           A star whose own Q is <= 0 would run get_positive_Q_index's O(N) linear scan
//                           (cmc_orbit.c:46-57) in ONE thread while the rest of the GPU waits: it is queued instead.
//   orbit_scan_kernel       the queued linear scans as a grid-wide find-first: work items (chunk of 4096 table entries,
//                           queued star) are handed out chunk-major so all SMs work on the lowest unresolved chunks; a hit
//                           retires the later chunks of that star.
//   orbit_deferred_kernel   finishes the queued stars with the index the scan found.
// The fine r/Phi table (16 B per star) leaves L2 beyond N ~ 4e6; the ladders keep all but the last ~6 probes of a bisection and
// all but the last line of a potential() lookup off it.  Bound: dependent-load latency x occupancy, then DRAM sector
// bandwidth -- not HBM streaming.
#include "orbit.cuh"
#include "philox.cuh"

namespace cmcb {

struct CsrDraws {
	const double *d;
	long pos, end;
	__device__ bool has(int n) const { return pos + n <= end; }
	__device__ double next() { return d[pos++]; }
};

static_assert(QLAD_MIN_SHIFT == 4, "orbit_pack_tables reduces the finest Q-ladder blocks over 16-lane groups");

// rp[i] = (r[i], phi[i]) for i < n; with q_fine != nullptr also block b of the finest Q-ladder level (entries [16 b, min(16 b + 16, M)]) and
// the sortedness check of r[0 .. M] (any violation, or a NaN in phi, raises *unsorted: the kernels then ignore the ladders)
__global__ void __launch_bounds__(256) orbit_pack_tables(const double *__restrict__ r, const double *__restrict__ phi, RPhi *__restrict__ rp, RPhiF *__restrict__ rpf,
							 long n, long M, LBlock *__restrict__ q_fine, int *unsorted)
{
	const long stride = (long) gridDim.x * blockDim.x; // a multiple of 32: the 16-lane groups stay aligned to the blocks
	const long n_round = ((n > M + 1 ? n : M + 1) + 31) / 32 * 32;
	bool bad = false;
	for (long i = blockIdx.x * (long) blockDim.x + threadIdx.x; i < n_round; i += stride) {
		double2 v = make_double2(0.0, 0.0);
		if (i < n) {
			v.x = r[i];
			v.y = phi[i];
			*reinterpret_cast<double2 *>(rp + i) = v;
			if (rpf) *reinterpret_cast<float2 *>(rpf + i) = make_float2(__double2float_rn(v.x), __double2float_rn(v.y)); // the screening copy (orbit_core.h: sign_Q_screened)
		}
		if (q_fine) {
			const bool in = i <= M;
			if (i < M && !(v.x <= r[i + 1])) bad = true;
			if (in && !(v.y == v.y)) bad = true;
			double pmin = in ? v.y : __builtin_huge_val(), pmax = in ? v.y : -__builtin_huge_val();
#pragma unroll
			for (int d = 8; d > 0; d >>= 1) {
				pmin = fmin(pmin, __shfl_xor_sync(0xffffffffu, pmin, d));
				pmax = fmax(pmax, __shfl_xor_sync(0xffffffffu, pmax, d));
			}
			if ((lane_id() & 15) == 0 && in) {
				const long k1 = i + 16 < M ? i + 16 : M; // the block's last entry is the next block's first
				const double p1 = phi[k1];
				LBlock B;
				B.r_lo = v.x; B.r_hi = r[k1]; B.phi_min = fmin(pmin, p1); B.phi_max = fmax(pmax, p1);
				double2 *o = reinterpret_cast<double2 *>(q_fine + (i >> 4));
				o[0] = make_double2(B.r_lo, B.r_hi);
				o[1] = make_double2(B.phi_min, B.phi_max);
			}
		}
	}
	if (bad) *unsorted = 1;
}

// one coarser level of the Q ladder: block b = the summary of child blocks [8 b, 8 b + 8)
__global__ void qlad_coarsen(const LBlock *__restrict__ child, long n_child, LBlock *__restrict__ parent, long n_parent)
{
	const long b = blockIdx.x * (long) blockDim.x + threadIdx.x;
	if (b >= n_parent) return;
	const long c0 = b << QLAD_STEP;
	const long nc = n_child - c0 < (1 << QLAD_STEP) ? n_child - c0 : (1 << QLAD_STEP);
	parent[b] = lblock_merge(child + c0, nc);
}

struct RLadderPlan {
	double *lvl[LADDER_MAX_LEVELS];
	int shift[LADDER_MAX_LEVELS];
	long count[LADDER_MAX_LEVELS], alloc[LADDER_MAX_LEVELS];
};
// the r ladder: level blockIdx.y holds star_r[b << shift], +inf in the padding
__global__ void rlad_build(const double *__restrict__ r, RLadderPlan p)
{
	const int l = blockIdx.y;
	const long stride = (long) gridDim.x * blockDim.x;
	for (long i = blockIdx.x * (long) blockDim.x + threadIdx.x; i < p.alloc[l]; i += stride)
		p.lvl[l][i] = i < p.count[l] ? r[i << p.shift[l]] : __builtin_huge_val();
}

// the sample ladder: level blockIdx.y holds the table entries (r, Phi) at k = b << shift
struct SLadderPlan {
	RPhi *lvl[LADDER_MAX_LEVELS];
	RPhiF *topf; // level 0 rounded to FP32 (nullptr: none)
	int shift[LADDER_MAX_LEVELS];
	long count[LADDER_MAX_LEVELS];
};
__global__ void slad_build(const RPhi *__restrict__ rp, SLadderPlan p)
{
	const int l = blockIdx.y;
	const long stride = (long) gridDim.x * blockDim.x;
	for (long i = blockIdx.x * (long) blockDim.x + threadIdx.x; i < p.count[l]; i += stride) {
		const double2 v = *reinterpret_cast<const double2 *>(rp + (i << p.shift[l]));
		*reinterpret_cast<double2 *>(p.lvl[l] + i) = v;
		if (l == 0 && p.topf) *reinterpret_cast<float2 *>(p.topf + i) = make_float2((float) v.x, (float) v.y);
	}
}

struct StarIn {
	kint g;
	bool is_new;
	double E, J, rnew_of_new;
};

__device__ __forceinline__ StarIn load_star(const OrbitArgs &a, long j0)
{
	StarIn in;
	const long j = j0 + 1; // the reference's 1-based local index
	in.is_new = j > a.n_old_local;
	in.g = (kint) (in.is_new ? 1 + a.t.N_MAX + (j - a.n_old_local) : a.first_global + j - 1); // get_global_idx, cmc_utils.c:2384-2407
	const char *rec = a.star1 + j0 * CMCB_SIZEOF_STAR;
	in.E = ld_field(rec, CMCB_OFF_E); in.J = ld_field(rec, CMCB_OFF_J);
	in.rnew_of_new = in.is_new ? ld_field(rec, CMCB_OFF_RNEW) : 0.0;
	return in;
}

// what the loop body writes into star[j] (cmc_evolution_thr.c:835-839, 852-853, 918-931; remove_star :662-664)
__device__ __forceinline__ void store_star(const OrbitArgs &a, long j0, const OrbitOut &o)
{
	char *rec = a.star1 + j0 * CMCB_SIZEOF_STAR;
	if (o.status == ORBIT_CIRCULAR) { // rnew / vrnew / vtnew = r / vr / vt (:836-838)
		st_field(rec, CMCB_OFF_RNEW, o.rnew);
		st_field(rec, CMCB_OFF_VRNEW, ld_field(rec, CMCB_OFF_VR));
		st_field(rec, CMCB_OFF_VTNEW, ld_field(rec, CMCB_OFF_VT));
	} else if (o.status != ORBIT_NTRY) {
		st_field(rec, CMCB_OFF_RNEW, o.rnew);
		st_field(rec, CMCB_OFF_VRNEW, o.vrnew);
		st_field(rec, CMCB_OFF_VTNEW, o.vtnew);
	}
	if (o.status == ORBIT_OK) st_field(rec, CMCB_OFF_X, o.X);
	if (o.status == ORBIT_OK || o.status == ORBIT_TIDAL) {
		st_field(rec, CMCB_OFF_R_PERI, o.r_peri);
		st_field(rec, CMCB_OFF_R_APO, o.r_apo);
	}
	if (a.sort_keys) a.sort_keys[j0] = key_from_double(o.status != ORBIT_NTRY ? o.rnew : ld_field(rec, CMCB_OFF_RNEW)); // what the record now holds
	if (a.status) a.status[j0] = o.status;
	if (a.kmin) a.kmin[j0] = o.kmin;
	if (a.kmax) a.kmax[j0] = o.kmax;
	if (a.n_draws) a.n_draws[j0] = o.n_draws;
}

// ---- stage A: the orbit of every star (index bisections, r_min / r_max, envelope) ------------------------------------------

// block-level accumulation of the per-launch statistics: one atomic per warp and counter
__device__ __forceinline__ void accumulate_stats(OrbitStats *st, int status, double rnew, int n_draws, int n_exact, int n_brent, bool valid)
{
	unsigned long long my_max = (valid && status == ORBIT_OK) ? dbits(rnew) : 0ull; // positive doubles order like their bit patterns
	unsigned my_draws = valid ? n_draws : 0, my_exact = valid ? n_exact : 0, my_brent = valid ? n_brent : 0;
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
		if (my_draws) atomicAdd(&st->n_draws, (unsigned long long) my_draws);
		if (my_exact) atomicAdd(&st->n_exact, (unsigned long long) my_exact);
		if (my_brent) atomicAdd(&st->n_brent, (unsigned long long) my_brent);
	}
	if (valid && status > 0 && status < 6) atomicAdd(&st->n_status[status], 1ull); // the rare outcomes; OK count = n - sum
}

// A star whose orbit is known but whose new radius is not drawn yet parks its job -- the envelope F, r_min, r_max and its E, J -- in five
// compact arrays (OrbitArgs::job).
__device__ __forceinline__ void store_prepared(const OrbitArgs &a, long j0, const OrbitOut &o, const OrbitJob &job, double E, double J)
{
	if (o.status == ORBIT_PENDING) {
		// r_min / r_max (:919-920) reach the record when a try is accepted (store_star, next to X in the same 32-byte sector): storing them here
		// as well cost stage A a read-modify-write of that sector for every star.  The job goes to the compact arrays the rejection rounds read.
		a.job[0 * a.job_stride + j0] = job.F;
		a.job[1 * a.job_stride + j0] = job.rmin;
		a.job[2 * a.job_stride + j0] = job.rmax;
		a.job[3 * a.job_stride + j0] = E;
		a.job[4 * a.job_stride + j0] = J;
		a.flag[j0] = 1;
		if (a.kmin) a.kmin[j0] = o.kmin;
		if (a.kmax) a.kmax[j0] = o.kmax;
	} else {
		a.flag[j0] = 0;
		store_star(a, j0, o);
	}
}

// CTAS = resident CTAs per SM the register allocation is capped for (4: 64 registers, 3: 80)
template <int CTAS>
__global__ void __launch_bounds__(ORBIT_THREADS, CTAS) orbit_prepare_kernel(OrbitArgs a)
{
	extern __shared__ __align__(128) unsigned char orbit_smem[];
	__shared__ __align__(8) unsigned long long bar;
	__shared__ QLadder s_qlad; // the level table, indexed dynamically by the bisections
	__shared__ SLadder s_slad;
	if (threadIdx.x == 0) {
		s_qlad = a.t.qlad;
		s_slad = a.t.slad;
		// max |Phi| of a sorted table sits at its innermost entries (Phi_0 is the central-mass slot)
		s_slad.phi_abs_max = fmax(fabs(tbl_rp(a.t, 0).phi), fabs(tbl_rp(a.t, 1).phi));
	}
	LadderTops top; // (stage_top_level's barrier also publishes s_qlad / s_slad; without a ladder nobody reads them)
	const bool lad_on = a.t.qlad.n_levels > 0 && *a.lad_unsorted == 0;
	const bool slad_on = a.t.slad.n_levels > 0 && *a.lad_unsorted == 0; // (never both: the two ladders share one buffer)
	top.q = (const LBlock *) stage_top_level(a.t.qlad.blk[0], a.q_top_bytes, lad_on && a.q_top_bytes > 0, orbit_smem, &bar);
	if (lad_on && a.q_top_bytes == 0) { top.q = a.t.qlad.blk[0]; __syncthreads(); } // coarsest level too large to stage (tuning knobs): read in place
	if (CMCB_SLAD_F32_TOP && slad_on && a.s_topf_bytes > 0) { // the FP32 copy of the coarsest level in shared memory; its FP64 entries stay in global memory for the undecided probes
		top.sf = (const RPhiF *) stage_top_level(a.t.slad.topf, a.s_topf_bytes, true, orbit_smem, &bar);
		top.s = a.t.slad.lvl[0];
	} else top.s = (const RPhi *) stage_top_level(a.t.slad.lvl[0], a.s_top_bytes, slad_on, orbit_smem, &bar);
	if (!lad_on && !slad_on) __syncthreads();
	top.r = nullptr; // potential() is not called in this stage
	top.qdesc = &s_qlad;
	top.sdesc = &s_slad;
	int n_exact = 0, n_brent = 0;
	for (long base = blockIdx.x * (long) ORBIT_THREADS; base < a.n_local; base += (long) gridDim.x * ORBIT_THREADS) {
		const long j0 = base + threadIdx.x;
		if (j0 >= a.n_local) break;
		OrbitOut o;
		StarIn in = load_star(a, j0);
		if (a.fuse_energy && !in.is_new) { // ComputeEnergy's first loop (reference src/cmc/cmc_utils.c:612-616) for this star, then stored
			char *rec = a.star1 + j0 * CMCB_SIZEOF_STAR;
			const RPhi own = tbl_rp(a.t, in.g);
			const double vr = ld_field(rec, CMCB_OFF_VR), vt = ld_field(rec, CMCB_OFF_VT);
			in.E = dadd(own.phi, dmul(0.5, dadd(dmul(vr, vr), dmul(vt, vt))));
			in.J = dmul(own.r, vt);
			st_field(rec, CMCB_OFF_E, in.E);
			st_field(rec, CMCB_OFF_J, in.J);
		}
		const StarConst s = star_const(a.t, in.g, in.E, in.J);
		OrbitJob job;
		orbit_prepare(a.t, top, s, in.is_new, in.rnew_of_new, a.Rtidal, KTEMP_COMPUTE, true, o, job);
		if (o.status == ORBIT_DEFERRED) {
			const unsigned slot = atomicAdd(&a.stats->n_defer, 1u);
			if (slot < ORBIT_DEFER_CAP) {
				a.defer_star[slot] = j0;
				a.defer_found[slot] = ORBIT_NOT_FOUND;
				a.defer_first[slot] = 0xffffffffu;
			} else { // queue full (never seen: ~1e-5 of the stars defer): last resort, the scan in this thread
				orbit_prepare(a.t, top, s, in.is_new, in.rnew_of_new, a.Rtidal, KTEMP_COMPUTE, false, o, job);
			}
		}
		if (o.status != ORBIT_DEFERRED) {
			store_prepared(a, j0, o, job, in.E, in.J);
			n_exact += o.n_exact; n_brent += o.n_brent;
			if (o.status > 0 && o.status < 6) atomicAdd(&a.stats->n_status[o.status], 1ull); // the rare outcomes; OK count = n - sum
		}
	}
	accumulate_stats(a.stats, -1, 0.0, 0, n_exact, n_brent, true); // one atomic per warp and counter, once per persistent CTA
}

// ---- the queued linear scans ------------------------------------------------------------------------------------------
//
// first k in [0, N_MAX) with function_q(k) > 0 for every queued star (get_positive_Q_index, cmc_orbit.c:46-57: the loop runs
// k = 0, 1, .. while Q <= 0 and k <= N_MAX, and a result >= N_MAX means "none").
//   orbit_scan_filter   one thread per (queued star, chunk of 4096 table entries): a two-load bound (q_maybe_positive_on) proves
//                       Q <= 0 on the whole chunk for every chunk well below the sign change; a chunk it cannot dismiss is tested again in
//                       16 sub-chunks of 256 entries, and the lowest sub-chunk that still cannot be dismissed is kept per star
//   orbit_scan_chunks   one CTA per queued star walks the sub-chunks from that one upwards, one table entry per thread, evaluating every
//                       sign exactly as the reference would, until a sub-chunk holds an index with Q > 0 (almost always within a few steps)
// Every entry below the answer is either in a range proven all-non-positive or evaluated, so the answer is the reference's.
constexpr int ORBIT_SCAN_SUB = ORBIT_SCAN_THREADS;                    // 256 entries
constexpr int ORBIT_SCAN_SUBS = ORBIT_SCAN_CHUNK / ORBIT_SCAN_SUB;     // 16 sub-chunks per chunk

__device__ __forceinline__ StarConst deferred_const(const OrbitArgs &a, unsigned slot)
{
	const long j0 = a.defer_star[slot];
	const StarIn in = load_star(a, j0);
	return star_const(a.t, in.g, in.E, in.J);
}

__global__ void __launch_bounds__(256) orbit_scan_filter(OrbitArgs a)
{
	const unsigned n_defer = min(a.stats->n_defer, (unsigned) ORBIT_DEFER_CAP);
	if (n_defer == 0) return;
	const long n_scan = a.t.N_MAX; // candidates 0 .. N_MAX-1
	const long n_chunks = (n_scan + ORBIT_SCAN_CHUNK - 1) / ORBIT_SCAN_CHUNK;
	const unsigned long long n_items = (unsigned long long) n_chunks * n_defer;
	// consecutive threads = consecutive chunks of one star: the star's constants are fetched from cache, the bound loads coalesce
	for (unsigned long long item = blockIdx.x * (unsigned long long) blockDim.x + threadIdx.x; item < n_items; item += (unsigned long long) gridDim.x * blockDim.x) {
		const unsigned slot = (unsigned) (item / n_chunks);
		const long chunk = (long) (item % n_chunks);
		if ((unsigned) (chunk * ORBIT_SCAN_SUBS) >= *(volatile unsigned *) &a.defer_first[slot]) continue; // a lower candidate is already known
		const long k0 = chunk * ORBIT_SCAN_CHUNK;
		const long k1 = min(k0 + (long) ORBIT_SCAN_CHUNK, n_scan) - 1;
		const StarConst s = deferred_const(a, slot);
		if (!q_maybe_positive_on(s, tbl_rp(a.t, (kint) k0), tbl_rp(a.t, (kint) k1))) continue;
		for (int sub = 0; sub < ORBIT_SCAN_SUBS; sub++) {
			const long s0 = k0 + (long) sub * ORBIT_SCAN_SUB;
			if (s0 > k1) break;
			const long s1 = min(s0 + (long) ORBIT_SCAN_SUB - 1, k1);
			if (q_maybe_positive_on(s, tbl_rp(a.t, (kint) s0), tbl_rp(a.t, (kint) s1))) {
				atomicMin(&a.defer_first[slot], (unsigned) (chunk * ORBIT_SCAN_SUBS + sub));
				break;
			}
		}
	}
}

__global__ void __launch_bounds__(ORBIT_SCAN_THREADS) orbit_scan_chunks(OrbitArgs a)
{
	__shared__ long s_best[ORBIT_SCAN_THREADS / 32];
	const unsigned n_defer = min(a.stats->n_defer, (unsigned) ORBIT_DEFER_CAP);
	const long n_scan = a.t.N_MAX;
	const long n_subs = (n_scan + ORBIT_SCAN_SUB - 1) / ORBIT_SCAN_SUB;
	for (unsigned slot = blockIdx.x; slot < n_defer; slot += gridDim.x) {
		const StarConst s = deferred_const(a, slot);
		int n_exact = 0;
		long found = ORBIT_NOT_FOUND;
		for (long sub = a.defer_first[slot]; sub < n_subs && found == ORBIT_NOT_FOUND; sub++) { // 0xffffffff: no candidate at all
			const long k = sub * ORBIT_SCAN_SUB + threadIdx.x;
			long best = ORBIT_NOT_FOUND;
			if (k < n_scan && sign_Q(a.t, s, (kint) k, &n_exact) > 0) best = k;
#pragma unroll
			for (int d = 16; d > 0; d >>= 1) {
				const long o = __shfl_xor_sync(0xffffffffu, best, d);
				best = o < best ? o : best;
			}
			if (lane_id() == 0) s_best[warp_id()] = best;
			__syncthreads();
#pragma unroll
			for (int w = 0; w < ORBIT_SCAN_THREADS / 32; w++) found = s_best[w] < found ? s_best[w] : found;
			__syncthreads();
		}
#pragma unroll
		for (int d = 16; d > 0; d >>= 1) n_exact += __shfl_xor_sync(0xffffffffu, n_exact, d);
		if (lane_id() == 0 && n_exact) atomicAdd(&a.stats->n_exact, (unsigned long long) n_exact);
		if (threadIdx.x == 0) a.defer_found[slot] = found;
	}
}

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
			const StarConst s = star_const(a.t, in.g, in.E, in.J);
			OrbitJob job;
			LadderTops top; // the few queued stars bisect on the fine table
			top.q = nullptr; top.r = nullptr; top.qdesc = nullptr; top.s = nullptr; top.sdesc = nullptr;
			orbit_prepare(a.t, top, s, in.is_new, in.rnew_of_new, a.Rtidal, found == ORBIT_NOT_FOUND ? -1 : (kint) found, false, o, job);
			store_prepared(a, j0, o, job, in.E, in.J);
		}
		accumulate_stats(a.stats, o.status, 0.0, 0, o.n_exact, o.n_brent, valid);
	}
}

// ---- stage B: rejection sampling of the new radius, one ROUND of tries per launch -----------------------------------------
//
// Try number t of a star consumes uniforms 2t and 2t+1 of that star's stream and, when accepted, 2t+2 for the sign of v_r
// (:863,867,923) -- so tries are independent of each other, and "the first accepted try in order" can be found round by round
// instead of by one thread looping: every launch works on a COMPACTED list of the stars still pending, each lane evaluates one
// try (a full-depth FindZero_r bisection), accepted stars are finished, the others are appended to the next round's list.
// No lane waits for a neighbour's longer rejection streak (the old per-thread loop ran at 60 % lane utilisation).
// Late rounds hold few stars and are latency-bound, so they evaluate W consecutive tries of each star at once (W lanes per
// star, first accepted in order wins: identical result, fewer dependent rounds).

struct TryDraws { // random access into a star's uniform stream
	double X, u;
};

template <bool RECORDED>
__device__ __forceinline__ bool draws_for_try(const OrbitArgs &a, long j0, kint g, unsigned t, TryDraws &d)
{
	if (RECORDED) {
		const long pos = a.draw_off[j0] + 2l * t;
		if (pos + 3 > a.draw_off[j0 + 1]) return false; // try + sign draw must exist
		d.X = a.draws[pos];
		d.u = a.draws[pos + 1];
		return true;
	} else {
		PhiloxDraws p;
		p.init(a.seed, a.tcount, (unsigned long long) g);
		p.c3 = t >> 1; // uniforms 4b .. 4b+3 live in block b
		p.refill();
		d.X = (double) ((t & 1) ? p.o2 : p.o0) / 4294967296.0; // same 2^-32 grid as taus113's rng_t113_dbl_new (taus113-v2.c:191-193)
		d.u = (double) ((t & 1) ? p.o3 : p.o1) / 4294967296.0;
		return true;
	}
}

template <bool RECORDED>
__device__ __forceinline__ double sign_draw_for_try(const OrbitArgs &a, long j0, kint g, unsigned t)
{
	if (RECORDED) return a.draws[a.draw_off[j0] + 2l * t + 2];
	PhiloxDraws p;
	p.init(a.seed, a.tcount, (unsigned long long) g);
	p.c3 = (2 * t + 2) >> 2;
	p.refill();
	return (double) ((t & 1) ? p.o0 : p.o2) / 4294967296.0;
}

// The rejection rounds are bound by the latency of potential()'s dependent loads (72 % of the try kernel's stall samples are long-scoreboard
// waits) at the 32 resident warps per SM its 64 registers allow.  So the lookup is a kernel of its own: it needs the proposed radius and the
// ladder search only, fits 42 registers, and runs with 48 resident warps per SM (512-thread CTAs, 3 per SM: the staged top level is 40 KB per
// CTA); it leaves potential(r) of every pending star's try `t` in the job array's sixth component (NaN: the bracket check failed).  The try
// kernel proper (POT_READY) then has no dependent random access left.
constexpr int ORBIT_LOOKUP_THREADS = 512;
template <bool RECORDED, bool FIRST>
__global__ void __launch_bounds__(ORBIT_LOOKUP_THREADS, 3) orbit_lookup_kernel(OrbitArgs a, int round, unsigned t)
{
	extern __shared__ __align__(128) unsigned char orbit_smem[];
	__shared__ __align__(8) unsigned long long bar;
	const unsigned n_work = FIRST ? (unsigned) a.n_local : a.stats->round_count[round];
	if ((unsigned long long) blockIdx.x * ORBIT_LOOKUP_THREADS >= n_work) return;
	LadderTops top;
	top.q = nullptr; top.qdesc = nullptr; top.s = nullptr; top.sdesc = nullptr;
	top.r = (const double *) stage_top_level(a.t.rlad.lvl[0], a.r_top_bytes, a.t.rlad.n_levels > 0 && *a.lad_unsorted == 0, orbit_smem, &bar);
	const int *__restrict__ list_in = a.pend_list[round & 1];
	for (unsigned w = blockIdx.x * ORBIT_LOOKUP_THREADS + threadIdx.x; w < n_work; w += gridDim.x * ORBIT_LOOKUP_THREADS) {
		const long j0 = FIRST ? (long) w : (long) list_in[w];
		if (FIRST && a.flag[j0] == 0) continue;
		const kint g = (kint) (j0 + 1 > a.n_old_local ? 1 + a.t.N_MAX + (j0 + 1 - a.n_old_local) : a.first_global + j0);
		TryDraws d;
		double pot = __longlong_as_double(0x7ff8000000000000ll);
		if (t < CMCB_N_TRY && draws_for_try<RECORDED>(a, j0, g, t, d)) {
			const double r = orbit_try_radius(__ldg(a.job + 1 * a.job_stride + j0), __ldg(a.job + 2 * a.job_stride + j0), d.X);
			int fail = 0;
			const double p = potential_at(a.t, top, r, &fail);
			if (!fail) pot = p;
		}
		a.job[5 * a.job_stride + j0] = pot;
	}
}

// W lanes per star evaluate tries t0 .. t0+W-1.  FIRST: the work list is "every star whose flag is set" (round 0, no list yet).
// LOOP: keep going (t0 += W) until the star is accepted or N_TRY is exceeded (the last round).
// POT_READY (W == 1 only): potential(r) of this try was left in the job array by orbit_lookup_kernel.
template <bool RECORDED, int W, bool FIRST, bool LOOP, bool POT_READY = false>
__global__ void __launch_bounds__(ORBIT_TRY_THREADS, 4) orbit_try_kernel(OrbitArgs a, int round, unsigned t0_base)
{
	extern __shared__ __align__(128) unsigned char orbit_smem[];
	__shared__ __align__(8) unsigned long long bar;
	const unsigned n_work = FIRST ? (unsigned) a.n_local : a.stats->round_count[round];
	if ((unsigned long long) blockIdx.x * (ORBIT_TRY_THREADS / W) >= n_work) return; // nothing for this CTA: skip the staging too
	LadderTops top;
	top.q = nullptr; top.qdesc = nullptr; top.s = nullptr; top.sdesc = nullptr;
	top.r = POT_READY ? nullptr : (const double *) stage_top_level(a.t.rlad.lvl[0], a.r_top_bytes, a.t.rlad.n_levels > 0 && *a.lad_unsorted == 0, orbit_smem, &bar);
	const int *__restrict__ list_in = a.pend_list[round & 1];
	int *__restrict__ list_out = a.pend_list[(round + 1) & 1];
	const unsigned sub = threadIdx.x % W;
	const unsigned groups_per_block = ORBIT_TRY_THREADS / W;
	const unsigned n_iter = (n_work + groups_per_block * gridDim.x - 1) / (groups_per_block * gridDim.x); // grid-uniform trip count
	for (unsigned it = 0; it < n_iter; it++) {
		const unsigned w = (it * gridDim.x + blockIdx.x) * groups_per_block + threadIdx.x / W;
		bool have = w < n_work;
		long j0 = 0;
		if (have) {
			j0 = FIRST ? (long) w : (long) list_in[w];
			if (FIRST) have = a.flag[j0] != 0;
		}
		int status = -1, n_draws = 0;
		double rnew = 0.0;
		bool requeue = false;
		if (have) {
			// the star's global index (get_global_idx) and its job: nothing of the 424-byte record is read in the rounds
			StarIn in;
			in.is_new = j0 + 1 > a.n_old_local;
			in.g = (kint) (in.is_new ? 1 + a.t.N_MAX + (j0 + 1 - a.n_old_local) : a.first_global + j0);
			OrbitJob job;
			job.F = __ldg(a.job + 0 * a.job_stride + j0); job.rmin = __ldg(a.job + 1 * a.job_stride + j0); job.rmax = __ldg(a.job + 2 * a.job_stride + j0);
			in.E = __ldg(a.job + 3 * a.job_stride + j0); in.J = __ldg(a.job + 4 * a.job_stride + j0);
			StarConst s;
			s.g_j = in.g; s.r_j = tbl_r(a.t, in.g); s.m_j = tbl_ld(a.t.m, in.g); s.J = in.J;
			s.E = dadd(in.E, phi_self(s.r_j, s.r_j, s.m_j, a.t.invN)); // :806
			s.JJ = dmul(s.J, s.J);
			unsigned t0 = t0_base;
			for (;;) {
				const unsigned t = t0 + sub;
				TryDraws d;
				int rc = -1; // -1: fatal (out of draws / bracket failure / beyond N_TRY), 0: rejected, 1: accepted
				double r = 0.0, vr = 0.0;
				if (t < CMCB_N_TRY && draws_for_try<RECORDED>(a, j0, in.g, t, d)) {
					if (POT_READY) {
						const double pot_raw = __ldg(a.job + 5 * a.job_stride + j0);
						r = orbit_try_radius(job.rmin, job.rmax, d.X);
						rc = orbit_try_finish(a.t, s, job, d.X, d.u, r, pot_raw, pot_raw != pot_raw, &vr);
					} else rc = orbit_try(a.t, top, s, job, d.X, d.u, &r, &vr);
				}
				// first lane of the group (lowest try index) that did not reject decides
				int first_rc = rc;
				unsigned first_sub = rc != 0 ? sub : W;
				if (W > 1) {
					const unsigned gmask = W == 32 ? 0xffffffffu : (((1u << W) - 1u) << (lane_id() & ~(W - 1)));
					const unsigned decided = __ballot_sync(gmask, rc != 0) & gmask;
					if (decided) {
						const int src = __ffs(decided) - 1;
						first_rc = __shfl_sync(gmask, rc, src);
						first_sub = (unsigned) src - (lane_id() & ~(W - 1));
					} else { first_rc = 0; first_sub = W; }
				}
				if (first_rc == 0) { // all W tries rejected
					if (LOOP) { t0 += W; continue; }
					requeue = sub == 0;
					break;
				}
				if (sub == first_sub) {
					OrbitOut o;
					o.kmin = o.kmax = -1; o.n_exact = o.n_brent = 0;
					if (first_rc > 0) orbit_finish(s, job, d.X, r, vr, sign_draw_for_try<RECORDED>(a, j0, in.g, t), (int) t + 1, o);
					else { // :890-893, fatal in the reference
						o.status = ORBIT_NTRY; o.rnew = s.r_j; o.vrnew = o.vtnew = 0.0; o.X = o.r_peri = o.r_apo = 0.0;
						o.n_draws = 2 * (int) (t < CMCB_N_TRY ? t + 1 : CMCB_N_TRY);
					}
					char *wrec = a.star1 + j0 * CMCB_SIZEOF_STAR;
					if (o.status == ORBIT_OK) {
						st_field(wrec, CMCB_OFF_RNEW, o.rnew); st_field(wrec, CMCB_OFF_VRNEW, o.vrnew); st_field(wrec, CMCB_OFF_VTNEW, o.vtnew);
						st_field(wrec, CMCB_OFF_X, o.X);
						st_field(wrec, CMCB_OFF_R_PERI, o.r_peri); // :919-920, next to X in the same 32-byte sector
						st_field(wrec, CMCB_OFF_R_APO, o.r_apo);
					}
					if (a.sort_keys) a.sort_keys[j0] = key_from_double(o.status == ORBIT_OK ? o.rnew : ld_field(wrec, CMCB_OFF_RNEW));
					if (a.status) a.status[j0] = o.status;
					if (a.n_draws) a.n_draws[j0] = o.n_draws;
					status = o.status; n_draws = o.n_draws; rnew = o.rnew;
				}
				break;
			}
		}
		// compact the rejected stars into the next round's list: one atomic per warp
		if (!LOOP) {
			const unsigned m = __ballot_sync(0xffffffffu, requeue);
			if (m) {
				unsigned base = 0;
				if (lane_id() == 0) base = atomicAdd(&a.stats->round_count[round + 1], (unsigned) __popc(m));
				base = __shfl_sync(0xffffffffu, base, 0);
				if (requeue) list_out[base + __popc(m & ((1u << lane_id()) - 1u))] = (int) j0;
			}
		}
		accumulate_stats(a.stats, status, rnew, n_draws, 0, 0, status >= 0);
	}
}

__global__ void orbit_stats_reset(OrbitStats *s);

// ---- the energy scheme's lookups at the new radii -----------------------------------------------------------------------
//
//   UOLD: energy_conservation2 (reference src/cmc/cmc_utils.c:2217-2226)   star[i].Uoldrnew = potential(rnew) + MPI_PHI_S(rnew, g_i), every star
//   EI:   ComputeIntermediateEnergy's first loop (:371-378)                 star[j].EI = sqr(vr) + sqr(vt) + star_phi[g_j] - potential(rnew)
//                                                                           for the stars with rnew < 1e6
// One potential() lookup per star through the r ladder (as in the rejection rounds), persistent CTAs.  A lookup whose bracket check
// (cmc_utils.c:144-152) fails is fatal in the reference: it is counted in stats->n_status[ORBIT_NTRY] and reported by the caller.
template <bool UOLD, bool EI>
__global__ void __launch_bounds__(ORBIT_TRY_THREADS, ORBIT_CTAS_PER_SM) energy_rnew_kernel(OrbitArgs a)
{
	extern __shared__ __align__(128) unsigned char orbit_smem[];
	__shared__ __align__(8) unsigned long long bar;
	LadderTops top;
	top.q = nullptr; top.qdesc = nullptr; top.s = nullptr; top.sdesc = nullptr;
	top.r = (const double *) stage_top_level(a.t.rlad.lvl[0], a.r_top_bytes, a.t.rlad.n_levels > 0 && *a.lad_unsorted == 0, orbit_smem, &bar);
	for (long j0 = blockIdx.x * (long) ORBIT_TRY_THREADS + threadIdx.x; j0 < a.n_local; j0 += (long) gridDim.x * ORBIT_TRY_THREADS) {
		const long j = j0 + 1;
		const kint g = (kint) (j > a.n_old_local ? 1 + a.t.N_MAX + (j - a.n_old_local) : a.first_global + j - 1); // get_global_idx
		char *rec = a.star1 + j0 * CMCB_SIZEOF_STAR;
		const double rn = ld_field(rec, CMCB_OFF_RNEW);
		if (!UOLD && !(rn < 1.0e6)) continue;
		int fail = 0;
		const double pot = potential_at(a.t, top, rn, &fail);
		if (fail) { atomicAdd(&a.stats->n_status[ORBIT_NTRY], 1ull); continue; }
		const RPhi own = tbl_rp(a.t, g);
		if (UOLD) st_field(rec, CMCB_OFF_UOLDRNEW, dadd(pot, phi_self(rn, own.r, tbl_ld(a.t.m, g), a.t.invN)));
		if (EI && rn < 1.0e6) {
			const double vr = ld_field(rec, CMCB_OFF_VR), vt = ld_field(rec, CMCB_OFF_VT);
			st_field(rec, CMCB_OFF_EI, dsub(dadd(dadd(dmul(vr, vr), dmul(vt, vt)), own.phi), pot));
		}
	}
}

int orbit_energy_rnew(const OrbitArgs &a, bool uold, bool ei, cudaStream_t s)
{
	orbit_stats_reset<<<1, 1, 0, s>>>(a.stats);
	if (a.n_local <= 0) return 1;
	long blocks = (a.n_local + ORBIT_TRY_THREADS - 1) / ORBIT_TRY_THREADS;
	if (blocks > (long) NUM_SMS_B200 * ORBIT_CTAS_PER_SM) blocks = (long) NUM_SMS_B200 * ORBIT_CTAS_PER_SM;
	if (uold && ei) energy_rnew_kernel<true, true><<<(unsigned) blocks, ORBIT_TRY_THREADS, a.r_top_bytes, s>>>(a);
	else if (uold) energy_rnew_kernel<true, false><<<(unsigned) blocks, ORBIT_TRY_THREADS, a.r_top_bytes, s>>>(a);
	else energy_rnew_kernel<false, true><<<(unsigned) blocks, ORBIT_TRY_THREADS, a.r_top_bytes, s>>>(a);
	return 2;
}

int orbit_reset_stats(const OrbitArgs &a, cudaStream_t s)
{
	orbit_stats_reset<<<1, 1, 0, s>>>(a.stats);
	return 1;
}

__global__ void orbit_stats_reset(OrbitStats *s)
{
	s->max_rad_bits = 0; s->n_draws = 0; s->n_exact = 0; s->n_brent = 0;
	for (int i = 0; i < 6; i++) s->n_status[i] = 0;
	s->n_defer = 0;
	for (int i = 0; i < ORBIT_MAX_ROUNDS + 2; i++) s->round_count[i] = 0;
}

long ladder_q_capacity(long n_table_dim)
{
	int sh[LADDER_MAX_LEVELS];
	long cnt[LADDER_MAX_LEVELS], total = 0;
	const int n = qlad_plan(n_table_dim, sh, cnt);
	for (int l = 0; l < n; l++) total += cnt[l] + 8;
	return total + 8;
}
long ladder_r_capacity(long n_table_dim)
{
	int sh[LADDER_MAX_LEVELS];
	long cnt[LADDER_MAX_LEVELS], al[LADDER_MAX_LEVELS], total = 0;
	const int n = rlad_plan(n_table_dim, sh, cnt, al);
	for (int l = 0; l < n; l++) total += al[l] + 16;
	return total + 16;
}

int orbit_build_tables_device(const double *r, const double *phi, RPhi *rp, RPhiF *rpf, long n, long M, const LadderBuffers &lb, OrbitArgs &a, cudaStream_t s)
{
	a.t.rpf = rpf;
	a.t.qlad.n_levels = 0; a.t.rlad.n_levels = 0; a.t.rlad.n_top = 0; a.t.slad.n_levels = 0;
	a.q_top_bytes = a.r_top_bytes = a.s_top_bytes = a.s_topf_bytes = 0;
	a.t.slad.topf = nullptr;
	a.lad_unsorted = lb.unsorted;
	if (n <= 0) return 0;
	int launches = 0;
	long qcnt[LADDER_MAX_LEVELS], rcnt[LADDER_MAX_LEVELS], ral[LADDER_MAX_LEVELS];
	int nq = (lb.q && !lb.sampled) ? qlad_plan(M, a.t.qlad.shift, qcnt) : 0;
	{ // tuning knob: drop the levels coarser than q_max_shift (brackets that wide then probe the fine table, see QLadder::max_width)
		int drop = 0;
		while (drop < nq && a.t.qlad.shift[drop] > lb.q_max_shift) drop++;
		for (int l = drop; l < nq; l++) { a.t.qlad.shift[l - drop] = a.t.qlad.shift[l]; qcnt[l - drop] = qcnt[l]; }
		nq -= drop;
	}
	a.t.qlad.factor_log2 = lb.q_factor_log2;
	a.t.qlad.max_width = lb.q_max_width_log2 >= 31 ? 0x7fffffff : (1 << lb.q_max_width_log2);
	int nr = lb.r ? rlad_plan(M, a.t.rlad.shift, rcnt, ral) : 0;
	// levels are stored finest first, each on a 128-byte boundary
	LBlock *qlvl[LADDER_MAX_LEVELS];
	{
		long off = 0;
		for (int l = nq - 1; l >= 0; l--) { qlvl[l] = lb.q + off; off += (qcnt[l] + 3) / 4 * 4; }
		if (off > lb.q_cap) nq = 0; // cannot happen: the capacity was planned for the largest table
	}
	RLadderPlan rplan;
	{
		long off = 0;
		for (int l = nr - 1; l >= 0; l--) { rplan.lvl[l] = lb.r + off; rplan.shift[l] = a.t.rlad.shift[l]; rplan.count[l] = rcnt[l]; rplan.alloc[l] = ral[l]; off += ral[l]; }
		if (off > lb.r_cap) nr = 0;
	}
	cudaMemsetAsync(lb.unsorted, 0, sizeof(int), s);
	long blocks = ((n > M + 1 ? n : M + 1) + 255) / 256;
	if (blocks > NUM_SMS_B200 * 16) blocks = NUM_SMS_B200 * 16;
	orbit_pack_tables<<<(unsigned) blocks, 256, 0, s>>>(r, phi, rp, rpf, n, M, nq ? qlvl[nq - 1] : nullptr, lb.unsorted);
	launches++;
	for (int l = nq - 2; l >= 0; l--) {
		qlad_coarsen<<<(unsigned) ((qcnt[l] + 255) / 256), 256, 0, s>>>(qlvl[l + 1], qcnt[l + 1], qlvl[l], qcnt[l]);
		launches++;
	}
	if (nr) {
		long rb = (ral[nr - 1] + 255) / 256;
		if (rb > NUM_SMS_B200 * 8) rb = NUM_SMS_B200 * 8;
		rlad_build<<<dim3((unsigned) rb, (unsigned) nr), 256, 0, s>>>(r, rplan);
		launches++;
	}
	if (lb.q && lb.sampled) { // the sample ladder of find_zero_Q_sampled (orbit_core.h), in the Q ladder's buffer: 16-byte entries, 32-byte blocks
		long scnt[LADDER_MAX_LEVELS];
		const int ns = slad_plan(M, a.t.slad.shift, scnt);
		SLadderPlan sp;
		long off = 0; // in RPhi entries; every level on a 128-byte boundary
		for (int l = ns - 1; l >= 0; l--) { sp.lvl[l] = (RPhi *) lb.q + off; sp.shift[l] = a.t.slad.shift[l]; sp.count[l] = scnt[l]; off += (scnt[l] + 7) / 8 * 8; }
		sp.topf = nullptr;
		if (ns && lb.f32_top && off + (scnt[0] / 2 + 8) <= 2 * lb.q_cap) { sp.topf = (RPhiF *) ((RPhi *) lb.q + off); off += (scnt[0] / 2 + 8) / 8 * 8; }
		if (ns && off <= 2 * lb.q_cap) {
			long sb = (scnt[ns - 1] + 255) / 256;
			if (sb > NUM_SMS_B200 * 8) sb = NUM_SMS_B200 * 8;
			slad_build<<<dim3((unsigned) sb, (unsigned) ns), 256, 0, s>>>(rp, sp);
			launches++;
			a.t.slad.n_levels = ns;
			for (int l = 0; l < ns; l++) a.t.slad.lvl[l] = sp.lvl[l];
			a.t.slad.margin = (double) M * 7.105427357601002e-15 /* 2^-47 */ + 9.094947017729282e-13 /* 2^-40 */;
			a.t.slad.phi_abs_max = 0.0; // filled by the kernels from the innermost entries
			a.s_top_bytes = (unsigned) (scnt[0] * sizeof(RPhi));
			a.t.slad.topf = sp.topf;
			a.s_topf_bytes = sp.topf ? (unsigned) ((scnt[0] + 1) / 2 * 2 * sizeof(RPhiF)) : 0u; // a multiple of 16 bytes for the bulk copy
		}
	}
	a.t.qlad.n_levels = nq;
	for (int l = 0; l < nq; l++) a.t.qlad.blk[l] = qlvl[l];
	if (nq && qcnt[0] <= QLAD_TOP_BLOCKS) a.q_top_bytes = (unsigned) (qcnt[0] * sizeof(LBlock));
	a.t.rlad.n_levels = nr;
	a.t.rlad.whole_lines = lb.r_whole_lines ? 1 : 0;
	for (int l = 0; l < nr; l++) a.t.rlad.lvl[l] = rplan.lvl[l];
	if (nr) { a.t.rlad.n_top = (int) rcnt[0]; a.r_top_bytes = (unsigned) ((rcnt[0] + 1) / 2 * 2 * sizeof(double)); }
	return launches;
}

template <bool RECORDED>
static int launch_try_rounds(const OrbitArgs &a, cudaStream_t s)
{
	// Rounds 0..4: one try per pending star (about 100, 48, 23, 11, 5.3 % of the stars); round 5: four tries at once for the
	// ~2.5 % left; round 6: a warp per star, 32 tries at a time until accepted.  The grids only bound the parallelism (the
	// kernels run grid-stride over the device-side count), sized for the expected survivors with a wide margin.
	const long n = a.n_local;
	const unsigned smem = a.r_top_bytes;
	auto grid_for = [&](double frac, int w) { // persistent CTAs: never more than fit on the device at once
		long g = (long) (n * frac * w / ORBIT_TRY_THREADS) + 1;
		const long cap = (long) NUM_SMS_B200 * ORBIT_CTAS_PER_SM;
		return (unsigned) (g < 1 ? 1 : (g > cap ? cap : g));
	};
	if (a.split_lookup) { // rounds 0..4 as lookup + try: see orbit_lookup_kernel
		auto lgrid = [&](double frac) {
			long g = (long) (n * frac / ORBIT_LOOKUP_THREADS) + 1;
			const long cap = (long) NUM_SMS_B200 * 3;
			return (unsigned) (g < 1 ? 1 : (g > cap ? cap : g));
		};
		const double frac[5] = {1.0, 0.60, 0.35, 0.20, 0.12};
		for (int rd = 0; rd < 5; rd++) {
			if (rd == 0) {
				orbit_lookup_kernel<RECORDED, true><<<lgrid(frac[rd]), ORBIT_LOOKUP_THREADS, smem, s>>>(a, rd, (unsigned) rd);
				orbit_try_kernel<RECORDED, 1, true, false, true><<<grid_for(frac[rd], 1), ORBIT_TRY_THREADS, 0, s>>>(a, rd, (unsigned) rd);
			} else {
				orbit_lookup_kernel<RECORDED, false><<<lgrid(frac[rd]), ORBIT_LOOKUP_THREADS, smem, s>>>(a, rd, (unsigned) rd);
				orbit_try_kernel<RECORDED, 1, false, false, true><<<grid_for(frac[rd], 1), ORBIT_TRY_THREADS, 0, s>>>(a, rd, (unsigned) rd);
			}
		}
	} else {
	orbit_try_kernel<RECORDED, 1, true, false><<<grid_for(1.0, 1), ORBIT_TRY_THREADS, smem, s>>>(a, 0, 0u);
	orbit_try_kernel<RECORDED, 1, false, false><<<grid_for(0.60, 1), ORBIT_TRY_THREADS, smem, s>>>(a, 1, 1u);
	orbit_try_kernel<RECORDED, 1, false, false><<<grid_for(0.35, 1), ORBIT_TRY_THREADS, smem, s>>>(a, 2, 2u);
	orbit_try_kernel<RECORDED, 1, false, false><<<grid_for(0.20, 1), ORBIT_TRY_THREADS, smem, s>>>(a, 3, 3u);
	orbit_try_kernel<RECORDED, 1, false, false><<<grid_for(0.12, 1), ORBIT_TRY_THREADS, smem, s>>>(a, 4, 4u);
	}
	orbit_try_kernel<RECORDED, 4, false, false><<<grid_for(0.07, 4), ORBIT_TRY_THREADS, smem, s>>>(a, 5, 5u);
	orbit_try_kernel<RECORDED, 32, false, true><<<grid_for(0.01, 32), ORBIT_TRY_THREADS, smem, s>>>(a, 6, 9u);
	return a.split_lookup ? 12 : 7;
}

// stage A: reset + the orbit of every star
int orbit_stage_prepare(const OrbitArgs &a, cudaStream_t s)
{
	orbit_stats_reset<<<1, 1, 0, s>>>(a.stats);
	if (a.n_local <= 0) return 1;
	long blocks = (a.n_local + ORBIT_THREADS - 1) / ORBIT_THREADS;
	int ctas = a.prepare_ctas >= 3 && a.prepare_ctas <= 6 ? a.prepare_ctas : ORBIT_CTAS_PER_SM;
	const unsigned s_bytes = a.s_topf_bytes ? a.s_topf_bytes : a.s_top_bytes;
	const unsigned top_bytes = a.q_top_bytes > s_bytes ? a.q_top_bytes : s_bytes;
	while (ctas > 4 && (long) ctas * (top_bytes + 1024) > 227 * 1024) ctas--; // the staged levels of all resident CTAs must fit
	if (blocks > (long) NUM_SMS_B200 * ctas) blocks = (long) NUM_SMS_B200 * ctas;
	if (ctas == 3) orbit_prepare_kernel<3><<<(unsigned) blocks, ORBIT_THREADS, top_bytes, s>>>(a);
	else if (ctas == 5) orbit_prepare_kernel<5><<<(unsigned) blocks, ORBIT_THREADS, top_bytes, s>>>(a);
	else if (ctas == 6) orbit_prepare_kernel<6><<<(unsigned) blocks, ORBIT_THREADS, top_bytes, s>>>(a);
	else orbit_prepare_kernel<ORBIT_CTAS_PER_SM><<<(unsigned) blocks, ORBIT_THREADS, top_bytes, s>>>(a);
	return 2;
}

// the queued linear scans and the stars that waited for them
int orbit_stage_scan(const OrbitArgs &a, cudaStream_t s)
{
	if (a.n_local <= 0) return 0;
	orbit_scan_filter<<<NUM_SMS_B200 * 4, 256, 0, s>>>(a);
	orbit_scan_chunks<<<NUM_SMS_B200 * 2, ORBIT_SCAN_THREADS, 0, s>>>(a);
	orbit_deferred_kernel<<<NUM_SMS_B200, ORBIT_THREADS, 0, s>>>(a);
	return 3;
}

// stage B: the rejection rounds
int orbit_stage_tries(const OrbitArgs &a, cudaStream_t s)
{
	if (a.n_local <= 0) return 0;
	return a.draws ? launch_try_rounds<true>(a, s) : launch_try_rounds<false>(a, s);
}

} // namespace cmcb

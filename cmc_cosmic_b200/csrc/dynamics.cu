// dynamics_apply's pair loop on the resident arrays (reference src/cmc/cmc_dynamics.c:206-542): every pair of neighbouring
// stars gets its kinematics, local density, cross section and encounter probability; pairs whose draw falls below it are LISTED for the
// caller (fewbody, collisions and BSE stay host code of the reference), all others are relaxed in place -- new vr, vt, E, J.  One thread
// per pair; the pair's two 424-byte records are neighbours, so a warp touches 64 consecutive records.
//
// Bound: DRAM sectors.  Per star the loop reads the sectors of (r, vr, vt, m), (E, J), Y, binind (+ rad, se_k with SS_COLLISION) and
// writes the first three: <= 6 + 3 sectors of 32 bytes against 40 bytes of fields.  The four transcendental pairs (sincos x3, log) are
// CUDA's double-precision routines: results agree with the reference's glibc to a few ulp, not bit for bit (dynamics_core.h).
#include <algorithm>
#include <string.h>
#include "ctx.cuh"
#include "dynamics_core.h"
#include "philox.cuh"

namespace cmcb {

constexpr int DYN_THREADS = 256;
constexpr int DYN_SCAN_BLOCK = 1024;

struct DynBlockStats { // one block's share of the scattering-angle statistics (cmc_dynamics.c:473-482)
	unsigned Nrel, Nrelbeta[4];
	double q[4], m[4], r[4];
};

struct DynEnc { // an encounter pair, appended in any order; the host sorts by pair
	long pair;
	int type;
	double rperi, W, rcm;
};

struct DynWork {
	unsigned *list;          // [n_local] non-flagged stars in star order (threebb compaction), 0-based local indices
	unsigned *blk_cnt;       // per scan block
	unsigned *n_list;        // total
	DynBlockStats *bstats;   // per pair block
	DynEnc *enc;
	unsigned long long *n_enc;
	double *totals;          // 5 counts (as doubles) + 12 sums
	long enc_cap;
	double *draws;
};

struct DynArgs {
	char *star1;
	const char *bin0; // binary[0]
	const double *r, *m, *phi;
	long n_local, first_global, n_pairs;
	DynConfig cfg;
	DynWork w;
	int use_list;
	unsigned long long seed, tcount;
};

__device__ __forceinline__ double ldf(const char *rec, int off) { return *(const double *) (rec + off); }
__device__ __forceinline__ long ldl(const char *rec, int off) { return *(const long *) (rec + off); }
__device__ __forceinline__ void stf(char *rec, int off, double v) { *(double *) (rec + off) = v; }

// ---- pairing when stars flagged by the three-body loop are skipped (cmc_dynamics.c:226-240) ------------------------------------------
__global__ void __launch_bounds__(DYN_SCAN_BLOCK) dyn_flag_count_kernel(DynArgs a)
{
	__shared__ unsigned wsum[DYN_SCAN_BLOCK / 32];
	const long i = (long) blockIdx.x * DYN_SCAN_BLOCK + threadIdx.x;
	const unsigned keep = i < a.n_local && ldl(a.star1 + i * CMCB_SIZEOF_STAR, CMCB_OFF_THREEBB_INTERACTED) != 1;
	const unsigned b = __ballot_sync(0xffffffffu, keep);
	if (lane_id() == 0) wsum[warp_id()] = __popc(b);
	__syncthreads();
	if (threadIdx.x == 0) {
		unsigned s = 0;
		for (int w = 0; w < DYN_SCAN_BLOCK / 32; w++) s += wsum[w];
		a.w.blk_cnt[blockIdx.x] = s;
	}
}
__global__ void __launch_bounds__(1024) dyn_flag_scan_kernel(DynArgs a, long n_blocks)
{
	// one block: exclusive scan of the block counts, chunk by chunk
	__shared__ unsigned wsum[32];
	__shared__ unsigned carry;
	if (threadIdx.x == 0) carry = 0;
	__syncthreads();
	for (long base = 0; base < n_blocks; base += 1024) {
		const long i = base + threadIdx.x;
		const unsigned v = i < n_blocks ? a.w.blk_cnt[i] : 0;
		unsigned inc = v;
		for (int d = 1; d < 32; d <<= 1) {
			const unsigned o = __shfl_up_sync(0xffffffffu, inc, d);
			if (lane_id() >= (unsigned) d) inc += o;
		}
		if (lane_id() == 31) wsum[warp_id()] = inc;
		__syncthreads();
		unsigned wpre = 0;
		for (unsigned w = 0; w < warp_id(); w++) wpre += wsum[w];
		const unsigned c0 = carry;
		if (i < n_blocks) a.w.blk_cnt[i] = c0 + wpre + inc - v;
		__syncthreads();
		if (threadIdx.x == 1023) carry = c0 + wpre + inc;
		__syncthreads();
	}
	if (threadIdx.x == 0) *a.w.n_list = carry;
}
__global__ void __launch_bounds__(DYN_SCAN_BLOCK) dyn_flag_fill_kernel(DynArgs a)
{
	__shared__ unsigned wsum[DYN_SCAN_BLOCK / 32];
	const long i = (long) blockIdx.x * DYN_SCAN_BLOCK + threadIdx.x;
	const unsigned keep = i < a.n_local && ldl(a.star1 + i * CMCB_SIZEOF_STAR, CMCB_OFF_THREEBB_INTERACTED) != 1;
	const unsigned b = __ballot_sync(0xffffffffu, keep);
	if (lane_id() == 0) wsum[warp_id()] = __popc(b);
	__syncthreads();
	unsigned wpre = 0;
	for (unsigned w = 0; w < warp_id(); w++) wpre += wsum[w];
	if (keep) a.w.list[a.w.blk_cnt[blockIdx.x] + wpre + __popc(b & ((1u << lane_id()) - 1))] = (unsigned) i;
}

// ---- the pair loop ----------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void dyn_load_star(const DynArgs &a, long k /* 1-based local */, DynStar *s)
{
	const char *rec = a.star1 + (k - 1) * CMCB_SIZEOF_STAR;
	const long g = a.first_global + k - 1;
	s->vr = ldf(rec, CMCB_OFF_VR);
	s->vt = ldf(rec, CMCB_OFF_VT);
	s->binind = ldl(rec, CMCB_OFF_BININD);
	s->m = a.m[g];
	s->r = a.r[g];
	s->phi = a.phi[g];
	s->rad = 0.0;
	s->se_k = 0;
	s->bin_a = 0.0;
	if (s->binind > 0) s->bin_a = ldf(a.bin0 + s->binind * CMCB_SIZEOF_BINARY, CMCB_BIN_OFF_A);
	else if (a.cfg.ss_collision) {
		s->rad = ldf(rec, CMCB_OFF_RAD);
		s->se_k = *(const int *) (rec + CMCB_OFF_SE_K);
	}
}

#ifndef CMCB_DYN_MIN_BLOCKS
#define CMCB_DYN_MIN_BLOCKS 1
#endif
__global__ void __launch_bounds__(DYN_THREADS, CMCB_DYN_MIN_BLOCKS) dyn_pair_kernel(DynArgs a)
{
	__shared__ DynBlockStats sh[DYN_THREADS / 32];
	const long i = (long) blockIdx.x * DYN_THREADS + threadIdx.x;
	unsigned nrel = 0, nb[4] = {0, 0, 0, 0};
	double sq[4] = {0, 0, 0, 0}, sm[4] = {0, 0, 0, 0}, sr[4] = {0, 0, 0, 0};
	bool live = i < a.n_pairs;
	long k = 2 * i + 1, kp = 2 * i + 2;
	if (live && a.use_list) {
		// pair i exists while the reference's loop condition holds when it is reached: si = (previous kp) + 1 <= n - n % 2 - 1 (:216)
		const unsigned nl = *a.w.n_list;
		const long lim = a.n_local - a.n_local % 2 - 1;
		if (2 * i + 1 >= (long) nl) live = false;
		else {
			k = (long) a.w.list[2 * i] + 1;
			kp = (long) a.w.list[2 * i + 1] + 1;
			if (i > 0 && (long) a.w.list[2 * i - 1] + 2 > lim) live = false;
		}
	}
	if (live) {
		DynStar s0, s1;
		dyn_load_star(a, k, &s0);
		dyn_load_star(a, kp, &s1);
		long kmin, kmax;
		dyn_n_local_range(a.first_global + k - 1, a.cfg.p, a.cfg.N_LIMIT, &kmin, &kmax);
		double u[3];
		if (a.w.draws) { u[0] = a.w.draws[3 * i]; u[1] = a.w.draws[3 * i + 1]; u[2] = a.w.draws[3 * i + 2]; }
		else {
			PhiloxDraws ph;
			ph.init(a.seed ^ 0xD1B54A32D192ED03ull, a.tcount, (unsigned long long) (a.first_global + k - 1));
			u[0] = ph.next(); u[1] = ph.next(); u[2] = ph.next();
		}
		DynPair o;
		dyn_pair(a.cfg, s0, s1, a.r[kmin], a.r[kmax], u, &o);
		char *r0 = a.star1 + (k - 1) * CMCB_SIZEOF_STAR, *r1 = a.star1 + (kp - 1) * CMCB_SIZEOF_STAR;
		stf(r0, CMCB_OFF_Y, o.Y);
		stf(r1, CMCB_OFF_Y, o.Y);
		if (o.status == DYN_RELAXED) {
			stf(r0, CMCB_OFF_VR, o.vr[0]); stf(r0, CMCB_OFF_VT, o.vt[0]); stf(r0, CMCB_OFF_E, o.E[0]); stf(r0, CMCB_OFF_J, o.J[0]);
			stf(r1, CMCB_OFF_VR, o.vr[1]); stf(r1, CMCB_OFF_VT, o.vt[1]); stf(r1, CMCB_OFF_E, o.E[1]); stf(r1, CMCB_OFF_J, o.J[1]);
			nrel = 1;
			const double relbeta[4] = {DYN_PI / 2.0, DYN_PI / 4.0, DYN_PI / 8.0, DYN_PI / 16.0};
#pragma unroll
			for (int j = 0; j < 4; j++) {
				if (o.beta > relbeta[j]) {
					nb[j] = 1;
					sq[j] = (s0.m > s1.m ? s0.m : s1.m) / (s0.m < s1.m ? s0.m : s1.m);
					sm[j] = (s0.m + s1.m) / 2.0 * a.cfg.units_mstar / DYN_MSUN;
					sr[j] = o.rcm;
				}
			}
		} else if (o.status == DYN_ENC_SS || o.status == DYN_ENC_BIN) {
			const unsigned long long slot = atomicAdd(a.w.n_enc, 1ull);
			if ((long) slot < a.w.enc_cap) {
				DynEnc e;
				e.pair = i; e.type = o.status; e.rperi = o.rperi; e.W = o.W; e.rcm = o.rcm;
				a.w.enc[slot] = e;
			}
		}
	}
	// block totals in a fixed order: lanes by shuffle, warps by one thread
#pragma unroll
	for (int d = 16; d > 0; d >>= 1) {
		nrel += __shfl_down_sync(0xffffffffu, nrel, d);
#pragma unroll
		for (int j = 0; j < 4; j++) {
			nb[j] += __shfl_down_sync(0xffffffffu, nb[j], d);
			sq[j] += __shfl_down_sync(0xffffffffu, sq[j], d);
			sm[j] += __shfl_down_sync(0xffffffffu, sm[j], d);
			sr[j] += __shfl_down_sync(0xffffffffu, sr[j], d);
		}
	}
	if (lane_id() == 0) {
		DynBlockStats &t = sh[warp_id()];
		t.Nrel = nrel;
		for (int j = 0; j < 4; j++) { t.Nrelbeta[j] = nb[j]; t.q[j] = sq[j]; t.m[j] = sm[j]; t.r[j] = sr[j]; }
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		DynBlockStats t = sh[0];
		for (int w = 1; w < DYN_THREADS / 32; w++) {
			t.Nrel += sh[w].Nrel;
			for (int j = 0; j < 4; j++) { t.Nrelbeta[j] += sh[w].Nrelbeta[j]; t.q[j] += sh[w].q[j]; t.m[j] += sh[w].m[j]; t.r[j] += sh[w].r[j]; }
		}
		a.w.bstats[blockIdx.x] = t;
	}
}

// totals[0..4] = Nrel, Nrelbeta[0..3]; totals[5 + 4 c + j] = sum c (q, m, r) of threshold j.  Fixed summation order: thread t adds blocks
// t, t + 256, ... in one pass over the block results, then the threads are added in order.
__global__ void __launch_bounds__(256) dyn_stats_kernel(DynArgs a, long n_blocks)
{
	__shared__ double part[17][256];
	double acc[17];
	for (int q = 0; q < 17; q++) acc[q] = 0.0;
	for (long b = threadIdx.x; b < n_blocks; b += 256) {
		const DynBlockStats t = a.w.bstats[b];
		acc[0] += (double) t.Nrel;
		for (int j = 0; j < 4; j++) { acc[1 + j] += (double) t.Nrelbeta[j]; acc[5 + j] += t.q[j]; acc[9 + j] += t.m[j]; acc[13 + j] += t.r[j]; }
	}
	for (int q = 0; q < 17; q++) part[q][threadIdx.x] = acc[q];
	__syncthreads();
	if (threadIdx.x < 17) {
		double tot = 0.0;
		for (int t = 0; t < 256; t++) tot += part[threadIdx.x][t];
		a.w.totals[threadIdx.x] = tot;
	}
}

// ---- simul_relax_new (reference src/cmc/cmc_dynamics_helper.c:2526-2580): the relaxation time step -----------------------------------------
// One thread per block of 2p consecutive stars: the three sums freshly added in star order, the local density at the block's centre, dt;
// the minimum over the blocks through an atomic on the bit pattern (positive doubles order like their bits).  Every operation is an IEEE
// add / mul / div / sqrt in the reference's order: the result is the reference's double.
__global__ void __launch_bounds__(128) simul_relax_kernel(const char *star1, const double *r, const double *m, long n_local, long first_global, long N_MAX,
							    double N_STAR, double madhoc, long p, double thetasemax, unsigned long long *dtmin_bits)
{
	const long nblk = n_local - p > 1 + p ? (n_local - p - (1 + p) + 2 * p - 1) / (2 * p) : 0; // si = 1 + p, 1 + 3p, ... while si < n_local - p
	const long b = (long) blockIdx.x * blockDim.x + threadIdx.x;
	if (b >= nblk) return;
	const long si = 1 + p + 2 * p * b, simin = si - p, simax = simin + (2 * p - 1);
	double Mv2ave = 0.0, Mave = 0.0, M2ave = 0.0;
	for (long k = simin; k <= simax; k++) {
		const char *rec = star1 + (k - 1) * CMCB_SIZEOF_STAR;
		const double vr = ldf(rec, CMCB_OFF_VR), vt = ldf(rec, CMCB_OFF_VT);
		const double tmp = m[first_global + k - 1] * madhoc;
		Mv2ave += tmp * (dyn_sqr(vr) + dyn_sqr(vt));
		Mave += tmp;
		M2ave += dyn_sqr(tmp);
	}
	Mv2ave /= (double) (2 * p);
	Mave /= (double) (2 * p);
	M2ave /= (double) (2 * p);
	const double sigma = sqrt(Mv2ave / Mave);
	const double W = 4.0 * sigma / sqrt(3.0 * DYN_PI);
	long kmin, kmax;
	dyn_n_local_range(first_global + si - 1, p, N_MAX, &kmin, &kmax);
	const double n_local_dens = (2.0 * ((double) p)) * 3.0 / (4.0 * DYN_PI * (dyn_cub(r[kmax]) - dyn_cub(r[kmin])));
	const double dt = dyn_sqr(2.0 * thetasemax / DYN_PI) * (DYN_PI / 32.0) * dyn_cub(W) / (N_STAR * n_local_dens * (4.0 * M2ave));
	if (dt < __longlong_as_double((long long) *dtmin_bits)) { // MIN(dtmin, dt): a NaN never replaces the minimum
		if (dt >= 0.0) atomicMin(dtmin_bits, (unsigned long long) __double_as_longlong(dt));
		else atomicExch(dtmin_bits, (unsigned long long) __double_as_longlong(dt)); // a negative step: cannot happen with sane inputs; keep it visible
	}
}

} // namespace cmcb

struct DynScratch {
	DynWork w;
	long cap_stars, cap_draws;
	bool last_use_list;
};

void dynamics_free(cmcb200_ctx *c)
{
	if (c->d_dtmin) { cudaFree(c->d_dtmin); c->d_dtmin = nullptr; }
	DynScratch *s = (DynScratch *) c->d_dynamics;
	if (!s) return;
	cudaFree(s->w.list); cudaFree(s->w.blk_cnt); cudaFree(s->w.n_list); cudaFree(s->w.bstats); cudaFree(s->w.enc); cudaFree(s->w.n_enc); cudaFree(s->w.totals);
	if (s->w.draws) cudaFree(s->w.draws);
	delete s;
	c->d_dynamics = nullptr;
}

extern "C" int cmcb200_dev_dynamics_apply(cmcb200_ctx *c, const cmcb200_dynamics_params *p, cmcb200_dynamics_result *res)
{
	CMCB_CUDA_OK(cudaSetDevice(c->device));
	if (join_uploads(c)) return -1;
	if (p->n_local + 2 > c->n_star_dim) { set_error("dynamics_apply: %ld stars exceed the capacity %ld", p->n_local, c->n_star_dim); return -1; }
	if (p->tc_polytrope) { set_error("dynamics_apply: TC_POLYTROPE is not handled on the device"); return -4; }
	if (p->bh_loss_cone && p->cenma_m > 0.0) { set_error("dynamics_apply: BH_LOSS_CONE with a central mass (bh_rand_walk) is not handled on the device"); return -4; }
	if (p->avekernel < 1 || p->N_MAX < 2 * p->avekernel + 1) { set_error("dynamics_apply: AVEKERNEL %ld does not fit N_MAX %ld", p->avekernel, p->N_MAX); return -1; }
	DynScratch *s = (DynScratch *) c->d_dynamics;
	if (!s) {
		s = new DynScratch();
		memset(s, 0, sizeof *s);
		const long n = c->n_star_dim, n_pairs = n / 2 + 1;
		s->w.enc_cap = n_pairs; // every pair may be an encounter (40 bytes per pair)
		CMCB_CUDA_OK(cudaMalloc(&s->w.list, (size_t) n * sizeof(unsigned)));
		CMCB_CUDA_OK(cudaMalloc(&s->w.blk_cnt, (size_t) (n / DYN_SCAN_BLOCK + 2) * sizeof(unsigned)));
		CMCB_CUDA_OK(cudaMalloc(&s->w.n_list, sizeof(unsigned)));
		CMCB_CUDA_OK(cudaMalloc(&s->w.bstats, (size_t) (n_pairs / DYN_THREADS + 2) * sizeof(DynBlockStats)));
		CMCB_CUDA_OK(cudaMalloc(&s->w.enc, (size_t) s->w.enc_cap * sizeof(DynEnc)));
		CMCB_CUDA_OK(cudaMalloc(&s->w.n_enc, sizeof(unsigned long long)));
		CMCB_CUDA_OK(cudaMalloc(&s->w.totals, 17 * sizeof(double)));
		s->cap_stars = n;
		c->d_dynamics = s;
	}
	DynArgs a;
	memset(&a, 0, sizeof a);
	a.star1 = star1(c);
	a.bin0 = c->d_bin[c->cur_bin];
	a.r = c->d_r; a.m = c->d_m; a.phi = c->d_phi;
	a.n_local = p->n_local;
	a.first_global = p->first_global;
	a.n_pairs = p->n_local / 2; // the last star of an odd count stays unpaired (:216)
	a.use_list = p->honor_threebb != 0;
	a.seed = p->seed; a.tcount = p->tcount;
	DynConfig &g = a.cfg;
	g.relaxation = p->relaxation; g.ss_collision = p->ss_collision; g.binsingle = p->binsingle; g.binbin = p->binbin;
	g.tidal_capture = p->tidal_capture; g.bh_capture = p->bh_capture; g.p = p->avekernel;
	g.gamma = p->gamma; g.coll_factor = p->coll_factor; g.tc_factor = p->tc_factor; g.wd_tc = p->wd_tc;
	g.units_mstar = p->units_mstar; g.units_l = p->units_l; g.units_t = p->units_t;
	g.dt = p->dt; g.madhoc = 1.0 / ((double) p->N_STAR); g.N_STAR = (double) p->N_STAR; g.N_LIMIT = p->N_MAX;
	a.w = s->w;
	a.w.draws = nullptr;
	if (p->draws && a.n_pairs > 0) {
		if (ensure((void **) &s->w.draws, &s->cap_draws, 3 * a.n_pairs, sizeof(double))) return -1;
		StageTimer t(c, ST_H2D);
		CMCB_CUDA_OK(cudaMemcpyAsync(s->w.draws, p->draws, (size_t) (3 * a.n_pairs) * sizeof(double), cudaMemcpyHostToDevice, c->stream));
		a.w.draws = s->w.draws;
	}
	CMCB_CUDA_OK(cudaMemsetAsync(s->w.n_enc, 0, sizeof(unsigned long long), c->stream));
	const long n_blocks = (a.n_pairs + DYN_THREADS - 1) / DYN_THREADS;
	{
		StageTimer t(c, ST_DYNAMICS);
		if (a.use_list && p->n_local > 0) {
			const long nb = (p->n_local + DYN_SCAN_BLOCK - 1) / DYN_SCAN_BLOCK;
			dyn_flag_count_kernel<<<(unsigned) nb, DYN_SCAN_BLOCK, 0, c->stream>>>(a);
			dyn_flag_scan_kernel<<<1, 1024, 0, c->stream>>>(a, nb);
			dyn_flag_fill_kernel<<<(unsigned) nb, DYN_SCAN_BLOCK, 0, c->stream>>>(a);
			c->launches += 3;
		}
		if (n_blocks > 0) {
			dyn_pair_kernel<<<(unsigned) n_blocks, DYN_THREADS, 0, c->stream>>>(a);
			c->launches += 1;
		}
		dyn_stats_kernel<<<1, 256, 0, c->stream>>>(a, n_blocks);
		c->launches += 1;
	}
	CMCB_CUDA_OK(cudaGetLastError());
	double *h = (double *) c->h_pinned;
	unsigned long long *hn = (unsigned long long *) (h + 17);
	unsigned *hl = (unsigned *) (h + 18);
	CMCB_CUDA_OK(cudaMemcpyAsync(h, s->w.totals, 17 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
	CMCB_CUDA_OK(cudaMemcpyAsync(hn, s->w.n_enc, sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
	if (a.use_list) CMCB_CUDA_OK(cudaMemcpyAsync(hl, s->w.n_list, sizeof(unsigned), cudaMemcpyDeviceToHost, c->stream));
	CMCB_CUDA_OK(cudaStreamSynchronize(c->stream));
	if (res) {
		res->n_encounters = (long) *hn;
		res->Nrel = (long) h[0];
		for (int j = 0; j < 4; j++) {
			res->Nrelbeta[j] = (long) h[1 + j];
			res->qaverelbeta[j] = h[5 + j]; res->maverelbeta[j] = h[9 + j]; res->raverelbeta[j] = h[13 + j];
		}
		res->n_pairs = a.n_pairs; // with skipped stars: an upper bound; Nrel + n_encounters (+ untouched pairs) is the number processed
	}
	s->last_use_list = a.use_list != 0;
	if ((long) *hn > s->w.enc_cap) { set_error("dynamics_apply: %llu encounter pairs exceed the list's capacity %ld", *hn, s->w.enc_cap); return -3; }
	return 0;
}

// the encounter pairs of the last cmcb200_dev_dynamics_apply in pair order: 1-based local indices of both stars, the type
// (1 single-single, 2 with a binary) and what the reference hands to sscollision_do / binint_do (rperi, W, rcm)
extern "C" int cmcb200_dev_dynamics_encounters(cmcb200_ctx *c, long cap, long *k, long *kp, int *type, double *rperi, double *W, double *rcm, long *count)
{
	CMCB_CUDA_OK(cudaSetDevice(c->device));
	DynScratch *s = (DynScratch *) c->d_dynamics;
	if (!s) { set_error("dynamics_encounters: no dynamics_apply ran on this context"); return -1; }
	unsigned long long n = 0;
	CMCB_CUDA_OK(cudaMemcpyAsync(&n, s->w.n_enc, sizeof n, cudaMemcpyDeviceToHost, c->stream));
	CMCB_CUDA_OK(cudaStreamSynchronize(c->stream));
	*count = (long) n;
	if ((long) n > cap || (long) n > s->w.enc_cap) { set_error("dynamics_encounters: %llu pairs, room for %ld", n, cap); return -3; }
	if (n == 0) return 0;
	std::vector<DynEnc> e(n);
	CMCB_CUDA_OK(cudaMemcpyAsync(e.data(), s->w.enc, n * sizeof(DynEnc), cudaMemcpyDeviceToHost, c->stream));
	std::vector<unsigned> lst;
	unsigned n_list = 0;
	CMCB_CUDA_OK(cudaStreamSynchronize(c->stream));
	std::sort(e.begin(), e.end(), [](const DynEnc &x, const DynEnc &y) { return x.pair < y.pair; });
	// with skipped stars the pair's members come from the compaction list
	const bool use_list = s->last_use_list;
	for (unsigned long long i = 0; i < n; i++) {
		long a0 = 2 * e[i].pair + 1, a1 = 2 * e[i].pair + 2;
		if (use_list) {
			if (lst.empty()) {
				CMCB_CUDA_OK(cudaMemcpy(&n_list, s->w.n_list, sizeof n_list, cudaMemcpyDeviceToHost));
				lst.resize(n_list);
				CMCB_CUDA_OK(cudaMemcpy(lst.data(), s->w.list, (size_t) n_list * sizeof(unsigned), cudaMemcpyDeviceToHost));
			}
			a0 = (long) lst[2 * e[i].pair] + 1; a1 = (long) lst[2 * e[i].pair + 1] + 1;
		}
		k[i] = a0; kp[i] = a1; type[i] = e[i].type; rperi[i] = e[i].rperi; W[i] = e[i].W; rcm[i] = e[i].rcm;
	}
	return 0;
}

// simul_relax_new (src/cmc/cmc_dynamics_helper.c:2526-2580) on the resident arrays: *DTrel = this rank's minimum (+inf without a complete block)
extern "C" int cmcb200_dev_simul_relax(cmcb200_ctx *c, long n_local, long first_global, long N_MAX, long N_STAR, long avekernel, double thetasemax, double *DTrel)
{
	CMCB_CUDA_OK(cudaSetDevice(c->device));
	if (join_uploads(c)) return -1;
	if (avekernel < 1 || n_local + 2 > c->n_star_dim) { set_error("simul_relax: AVEKERNEL %ld, n_local %ld", avekernel, n_local); return -1; }
	if (!c->d_dtmin) CMCB_CUDA_OK(cudaMalloc(&c->d_dtmin, sizeof(unsigned long long)));
	unsigned long long *d = c->d_dtmin;
	double *h = (double *) c->h_pinned;
	const double inf = __builtin_inf();
	memcpy(h, &inf, sizeof inf);
	CMCB_CUDA_OK(cudaMemcpyAsync(d, h, sizeof(double), cudaMemcpyHostToDevice, c->stream));
	const long p = avekernel, nblk = n_local - p > 1 + p ? (n_local - p - (1 + p) + 2 * p - 1) / (2 * p) : 0;
	if (nblk > 0) {
		StageTimer t(c, ST_DYNAMICS);
		simul_relax_kernel<<<(unsigned) ((nblk + 127) / 128), 128, 0, c->stream>>>(star1(c), c->d_r, c->d_m, n_local, first_global, N_MAX, (double) N_STAR,
											   1.0 / ((double) N_STAR), p, thetasemax, d);
		c->launches += 1;
	}
	CMCB_CUDA_OK(cudaGetLastError());
	CMCB_CUDA_OK(cudaMemcpyAsync(h + 1, d, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
	CMCB_CUDA_OK(cudaStreamSynchronize(c->stream));
	*DTrel = h[1];
	return 0;
}

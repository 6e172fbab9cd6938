// energy_conservation1 and set_velocities3 (energy_conservation3) on the resident arrays -- SURVEY 8f N1, the last potential() loop of
// the energy scheme.  Reference: src/cmc/cmc_utils.c:2153-2174, 254-363.  The per-star arithmetic and the decomposition of the serial
// excess-energy carry are in velocities_core.h (shared with the host emulation the CPU tests check against the oracle).
//
// Kernels of the first loop (:268-311), all on the context's stream:
//   sv3_prepare_kernel  persistent CTAs, one thread per star: potential(rOld) through the r ladder (coarsest level TMA-staged in shared
//                       memory, as in the rejection rounds), vnew2, the rescaled (vr, vt) -> .vr/.vt/.vrnew/.vtnew, kind byte + operand.
//                       Streaming bytes per star: 64 read (8 fields) + 32 written + 9 of scratch; one ladder lookup.
//   sv3_block_kernel    one warp per block of 1024 stars: scans the kind bytes with 128-bit loads; only a block holding a source is walked
//                       (lane 0, from its first source) for the carry it emits when 0 enters.
//   sv3_carry_kernel    one warp: the carry entering every block, 32 block results per step; a step with no carry in flight and no
//                       emitting block is one coalesced load and a ballot.  A block a positive carry enters is replayed from it.
//   sv3_apply_kernel    one warp per block: the real walk of the blocks that hold a source or are entered by a positive carry.
// Second loop (:323-343), multi-rank only: sv3_second_kernel, one warp; each lane prepares one star's capacity (independent of the
// carry), the carry then runs over the 32 lanes with shuffles.
#include "ctx.cuh"
#include "velocities_core.h"

namespace cmcb {

struct Sv3Buffers {
	unsigned char *kind; // [n_pad] SV3_* bits per local star, n_pad = n_local rounded up to SV3_BLOCK (the padding is zeroed)
	double *val;         // [n_local] chain operand: a source's deficit, a sink's v^2 after the rescaling
	double *eb;          // [n_blocks] carry leaving the block when 0 enters
	double *ein;         // [n_blocks] carry entering the block
	unsigned char *has;  // [n_blocks] the block holds a source
	double *result;      // [0] leftover Eexcess of the last loop run
};

struct Sv3Noop {
	CMCB_HD void operator()(long, double) const {}
};
struct Sv3Scale { // star[i].vr *= exc_ratio; star[i].vt *= exc_ratio (:305-306)
	char *star1;
	CMCB_HD void operator()(long i, double ratio) const
	{
		double *vr = (double *) (star1 + i * CMCB_SIZEOF_STAR + CMCB_OFF_VR), *vt = (double *) (star1 + i * CMCB_SIZEOF_STAR + CMCB_OFF_VT);
		*vr = dmul(*vr, ratio);
		*vt = dmul(*vt, ratio);
	}
};

__device__ __forceinline__ Sv3Index sv3_index(const OrbitArgs &a)
{
	Sv3Index gi;
	gi.first_global = a.first_global; gi.n_old_local = a.n_old_local; gi.N_MAX = a.t.N_MAX;
	return gi;
}

// energy_conservation1, cmc_utils.c:2158-2173
__global__ void __launch_bounds__(256) sv3_save_kernel(OrbitArgs a)
{
	const Sv3Index gi = sv3_index(a);
	for (long j0 = blockIdx.x * 256L + threadIdx.x; j0 < a.n_local; j0 += gridDim.x * 256L) {
		char *rec = a.star1 + j0 * CMCB_SIZEOF_STAR;
		st_field(rec, CMCB_OFF_VTOLD, ld_field(rec, CMCB_OFF_VT));
		st_field(rec, CMCB_OFF_VROLD, ld_field(rec, CMCB_OFF_VR));
		st_field(rec, CMCB_OFF_UOLDROLD, sv3_U_own(a.t, gi(j0)));
	}
}

__global__ void __launch_bounds__(ORBIT_TRY_THREADS, ORBIT_CTAS_PER_SM) sv3_prepare_kernel(OrbitArgs a, Sv3Buffers b, long n_pad)
{
	extern __shared__ __align__(128) unsigned char orbit_smem[];
	__shared__ __align__(8) unsigned long long bar;
	LadderTops top;
	top.q = nullptr; top.qdesc = nullptr; top.s = nullptr; top.sdesc = nullptr;
	top.r = (const double *) stage_top_level(a.t.rlad.lvl[0], a.r_top_bytes, a.t.rlad.n_levels > 0 && *a.lad_unsorted == 0, orbit_smem, &bar);
	const Sv3Index gi = sv3_index(a);
	for (long j0 = blockIdx.x * (long) ORBIT_TRY_THREADS + threadIdx.x; j0 < n_pad; j0 += (long) gridDim.x * ORBIT_TRY_THREADS) {
		if (j0 >= a.n_local) { b.kind[j0] = 0; continue; }
		char *rec = a.star1 + j0 * CMCB_SIZEOF_STAR;
		Sv3In s;
		s.interacted = *(const long *) (rec + CMCB_OFF_INTERACTED);
		unsigned kind = 0;
		double val = 0.0;
		if (s.interacted == 0) {
			s.rOld = ld_field(rec, CMCB_OFF_ROLD);
			s.vrold = ld_field(rec, CMCB_OFF_VROLD); s.vtold = ld_field(rec, CMCB_OFF_VTOLD);
			s.Uoldrold = ld_field(rec, CMCB_OFF_UOLDROLD); s.Uoldrnew = ld_field(rec, CMCB_OFF_UOLDRNEW);
			s.vr = ld_field(rec, CMCB_OFF_VR); s.vt = ld_field(rec, CMCB_OFF_VT);
			int fail = 0;
			double vr = s.vr, vt = s.vt;
			kind = sv3_prepare_star(a.t, top, gi(j0), s, &vr, &vt, &val, &fail);
			if (fail) { atomicAdd(&a.stats->n_status[ORBIT_NTRY], 1ull); kind = 0; } // fatal in the reference (cmc_utils.c:144-152); reported by the caller
			else if (!(kind & SV3_SOURCE)) {
				st_field(rec, CMCB_OFF_VR, vr); st_field(rec, CMCB_OFF_VT, vt);       // :292-293
				st_field(rec, CMCB_OFF_VRNEW, vr); st_field(rec, CMCB_OFF_VTNEW, vt); // :295-296
			}
		}
		b.kind[j0] = (unsigned char) kind;
		b.val[j0] = val;
	}
}

// position of the first source among the SV3_BLOCK kind bytes at p (16-byte aligned), 0xffffffff if none; warp-uniform result
__device__ __forceinline__ unsigned sv3_first_source(const unsigned char *p)
{
	static_assert(SV3_BLOCK == 1024, "one lane scans 32 kind bytes");
	const uint4 *q = (const uint4 *) (p + lane_id() * 32);
	const uint4 v0 = q[0], v1 = q[1];
	const unsigned w[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
	const unsigned pattern = SV3_SOURCE * 0x01010101u;
	unsigned first = 0xffffffffu;
#pragma unroll
	for (int i = 7; i >= 0; i--) {
		const unsigned hit = w[i] & pattern;
		if (hit) first = lane_id() * 32 + i * 4 + ((__ffs(hit) - 1) >> 3);
	}
	return __reduce_min_sync(0xffffffffu, first);
}

__global__ void __launch_bounds__(256) sv3_block_kernel(OrbitArgs a, Sv3Buffers b, long n_blocks)
{
	const long blk = blockIdx.x * 8L + warp_id();
	if (blk >= n_blocks) return;
	const long lo = blk * SV3_BLOCK, hi = lo + SV3_BLOCK < a.n_local ? lo + SV3_BLOCK : a.n_local;
	const unsigned first = sv3_first_source(b.kind + lo);
	if (lane_id() != 0) return;
	if (first == 0xffffffffu) { b.eb[blk] = 0.0; b.has[blk] = 0; return; }
	b.has[blk] = 1;
	b.eb[blk] = sv3_walk(b.kind, b.val, a.t.m, sv3_index(a), lo + first, hi, 0.0, Sv3Noop());
}

__global__ void __launch_bounds__(32) sv3_carry_kernel(OrbitArgs a, Sv3Buffers b, long n_blocks)
{
	constexpr int AHEAD = 8; // block results of AHEAD steps are loaded together: the walk is a chain of dependent steps, the loads are not
	const unsigned lane = lane_id();
	const Sv3Index gi = sv3_index(a);
	double E = 0.0; // warp-uniform
	for (long c0 = 0; c0 < n_blocks; c0 += 32 * AHEAD) {
		double ek[AHEAD];
#pragma unroll
		for (int k = 0; k < AHEAD; k++) {
			const long blk = c0 + 32L * k + lane;
			ek[k] = blk < n_blocks ? b.eb[blk] : 0.0;
		}
#pragma unroll
		for (int k = 0; k < AHEAD; k++) {
			const long b0 = c0 + 32L * k, blk = b0 + lane;
			if (b0 >= n_blocks) break;
			const double e = ek[k];
			if (E == 0.0 && __ballot_sync(0xffffffffu, e != 0.0) == 0) { // nothing in flight, nothing emitted in these 32 blocks
				if (blk < n_blocks) b.ein[blk] = 0.0;
				continue;
			}
			for (int l = 0; l < 32 && b0 + l < n_blocks; l++) {
				const double el = __shfl_sync(0xffffffffu, e, l);
				if (lane == 0) b.ein[b0 + l] = E;
				if (E == 0.0) E = el; // 0 enters: the block kernel's result stands
				else {            // a chain crosses into this block: replay it from the carry
					double En = 0.0;
					if (lane == 0) {
						const long lo = (b0 + l) * SV3_BLOCK, hi = lo + SV3_BLOCK < a.n_local ? lo + SV3_BLOCK : a.n_local;
						En = sv3_walk(b.kind, b.val, a.t.m, gi, lo, hi, E, Sv3Noop());
					}
					E = __shfl_sync(0xffffffffu, En, 0);
				}
			}
		}
	}
	if (lane == 0) b.result[0] = E;
}

__global__ void __launch_bounds__(256) sv3_apply_kernel(OrbitArgs a, Sv3Buffers b, long n_blocks)
{
	const long blk = (blockIdx.x * 256L + threadIdx.x) >> 5;
	if (blk >= n_blocks || lane_id() != 0) return;
	const double E = b.ein[blk];
	if (E == 0.0 && !b.has[blk]) return;
	const long lo = blk * SV3_BLOCK, hi = lo + SV3_BLOCK < a.n_local ? lo + SV3_BLOCK : a.n_local;
	Sv3Scale scale;
	scale.star1 = a.star1;
	sv3_walk(b.kind, b.val, a.t.m, sv3_index(a), lo, hi, E, scale);
}

// one pass of the do-while loop's body (:323-343) from the received carry
__global__ void __launch_bounds__(32) sv3_second_kernel(OrbitArgs a, Sv3Buffers b, double E_in)
{
	const unsigned lane = lane_id();
	const Sv3Index gi = sv3_index(a);
	double E = E_in; // warp-uniform
	for (long i0 = 0; i0 < a.n_local && E > 0; i0 += 32) { // Eexcess > 0 false: no later star changes, nothing adds to it in this loop
		const long i = i0 + lane;
		unsigned kind = 0;
		double vr = 0, vt = 0, m = 1, v2 = 0, cap = 0;
		char *rec = a.star1 + i * CMCB_SIZEOF_STAR;
		if (i < a.n_local) kind = b.kind[i];
		const bool sink = (kind & (SV3_ACTIVE | SV3_POSITIVE)) == (SV3_ACTIVE | SV3_POSITIVE);
		if (sink) { // the capacity does not depend on the carry (:333)
			vr = ld_field(rec, CMCB_OFF_VR); vt = ld_field(rec, CMCB_OFF_VT);
			const double vrnew = ld_field(rec, CMCB_OFF_VRNEW), vtnew = ld_field(rec, CMCB_OFF_VTNEW);
			m = tbl_ld(a.t.m, gi(i));
			const double v2_new = dadd(dmul(vrnew, vrnew), dmul(vtnew, vtnew));
			v2 = dadd(dmul(vr, vr), dmul(vt, vt));
			cap = dsub(dmul(dmul(v2, m), 0.5), dmul(dmul(dmul(0.8, 0.5), m), v2_new));
		}
		const unsigned sinks = __ballot_sync(0xffffffffu, sink);
		double my_dump = 0.0;
		bool mine = false;
		for (unsigned rest = sinks; rest; rest &= rest - 1) {
			const int l = __ffs(rest) - 1;
			const double cl = __shfl_sync(0xffffffffu, cap, l);
			const double dump = cl <= E ? cl : E;
			if (E > 0) {
				if ((int) lane == l) { my_dump = dump; mine = true; }
				E = dsub(E, dump);
			}
		}
		if (mine) {
			const double ratio = dsqrt(ddiv(dsub(v2, ddiv(dmul(2.0, my_dump), m)), v2));
			st_field(rec, CMCB_OFF_VR, dmul(vr, ratio));
			st_field(rec, CMCB_OFF_VT, dmul(vt, ratio));
		}
	}
	if (lane == 0) b.result[0] = E;
}

// ---- tidally_strip_stars (reference src/cmc/cmc_evolution_thr.c:443-645): the per-star criterion of one pass -------------------------------
// A star still in the cluster (rnew < 1e6) is stripped when r_apo > Rtidal (TIDAL_TREATMENT 0, :479) or E > gierszalpha * phi_rtidal
// (TIDAL_TREATMENT 1, :555): rnew = SF_INFINITY, vrnew = vtnew = 0 (:481-483).  The hits are rare; their local indices are queued and
// sorted by the caller, which does the reference's per-star bookkeeping in star order on the host.
__global__ void __launch_bounds__(256) tidal_select_kernel(OrbitArgs a, int treatment, double threshold, long *list, unsigned *count, unsigned cap)
{
	// phase 1: the criterion only -- nothing is stored, so a list that does not fit can be retried with a larger one
	for (long j0 = blockIdx.x * 256L + threadIdx.x; j0 < a.n_local; j0 += gridDim.x * 256L) {
		const char *rec = a.star1 + j0 * CMCB_SIZEOF_STAR;
		if (!(ld_field(rec, CMCB_OFF_RNEW) < 1000000)) continue;
		const double v = ld_field(rec, treatment == 0 ? CMCB_OFF_R_APO : CMCB_OFF_E);
		if (!(v > threshold)) continue;
		const unsigned pos = atomicAdd(count, 1u);
		if (pos < cap) list[pos] = j0 + 1;
	}
}

// phase 2, once the list is known to fit: the stores of the reference's loop body (src/cmc/cmc_evolution_thr.c:497-499 / :573-575) for the
// listed stars (1-based local indices, star order) and star_m[g] of each (:471), the mass the caller's sums use
__global__ void tidal_apply_kernel(OrbitArgs a, const long *__restrict__ list, long n, long N_MAX, const double *__restrict__ star_m, double *__restrict__ m_out)
{
	const long k = blockIdx.x * (long) blockDim.x + threadIdx.x;
	if (k >= n) return;
	const long j = list[k];
	char *rec = a.star1 + (j - 1) * CMCB_SIZEOF_STAR;
	st_field(rec, CMCB_OFF_RNEW, CMCB_SF_INFINITY);
	st_field(rec, CMCB_OFF_VRNEW, 0.0);
	st_field(rec, CMCB_OFF_VTNEW, 0.0);
	const long g = j > a.n_old_local ? 1 + N_MAX + (j - a.n_old_local) : a.first_global + j - 1; // get_global_idx, src/cmc/cmc_utils.c:2384-2407
	m_out[k] = star_m[g];
}

// potential() (reference src/cmc/cmc_utils.c:112-165) at n radii through the r ladder; a failed bracket check is counted
__global__ void __launch_bounds__(ORBIT_TRY_THREADS) potential_at_kernel(OrbitArgs a, const double *__restrict__ r, long n, double *__restrict__ out)
{
	extern __shared__ __align__(128) unsigned char orbit_smem[];
	__shared__ __align__(8) unsigned long long bar;
	LadderTops top;
	top.q = nullptr; top.qdesc = nullptr; top.s = nullptr; top.sdesc = nullptr;
	top.r = (const double *) stage_top_level(a.t.rlad.lvl[0], a.r_top_bytes, a.t.rlad.n_levels > 0 && *a.lad_unsorted == 0, orbit_smem, &bar);
	for (long i = blockIdx.x * (long) ORBIT_TRY_THREADS + threadIdx.x; i < n; i += (long) gridDim.x * ORBIT_TRY_THREADS) {
		int fail = 0;
		out[i] = potential_at(a.t, top, r[i], &fail);
		if (fail) atomicAdd(&a.stats->n_status[ORBIT_NTRY], 1ull);
	}
}

} // namespace cmcb

static int sv3_buffers(cmcb200_ctx *c, Sv3Buffers &b)
{
	const long n_pad = (c->n_star_dim + SV3_BLOCK - 1) / SV3_BLOCK * SV3_BLOCK, nb = n_pad / SV3_BLOCK;
	const size_t off_val = (size_t) n_pad, off_eb = off_val + 8 * (size_t) n_pad, off_ein = off_eb + 8 * (size_t) nb, off_res = off_ein + 8 * (size_t) nb,
		     off_has = off_res + 64, total = off_has + (size_t) nb;
	if (!c->d_sv3) CMCB_CUDA_OK(cudaMalloc(&c->d_sv3, total));
	unsigned char *base = (unsigned char *) c->d_sv3;
	b.kind = base; b.val = (double *) (base + off_val); b.eb = (double *) (base + off_eb); b.ein = (double *) (base + off_ein);
	b.result = (double *) (base + off_res); b.has = base + off_has;
	return 0;
}

static long persistent_blocks(long n)
{
	long blocks = (n + ORBIT_TRY_THREADS - 1) / ORBIT_TRY_THREADS;
	const long cap = (long) NUM_SMS_B200 * ORBIT_CTAS_PER_SM;
	return blocks > cap ? cap : (blocks < 1 ? 1 : blocks);
}

extern "C" int cmcb200_dev_energy_conservation1(cmcb200_ctx *c, const cmcb200_orbit_params *p)
{
	CMCB_CUDA_OK(cudaSetDevice(c->device));
	if (join_uploads(c)) return -1;
	OrbitArgs a;
	if (orbit_begin(c, p, false, false, a)) return -1;
	if (p->n_local > 0) {
		StageTimer t(c, ST_ENERGY);
		sv3_save_kernel<<<(unsigned) persistent_blocks(p->n_local), 256, 0, c->stream>>>(a);
		c->launches += 1;
	}
	CMCB_CUDA_OK(cudaGetLastError());
	return 0;
}

// reads the loop's leftover and the count of failed potential() brackets
static int sv3_finish(cmcb200_ctx *c, const Sv3Buffers &b, double *Eexcess_left, const char *what)
{
	CMCB_CUDA_OK(cudaGetLastError());
	OrbitStats *h = (OrbitStats *) c->h_pinned;
	double *hr = (double *) ((char *) c->h_pinned + sizeof(OrbitStats));
	CMCB_CUDA_OK(cudaMemcpyAsync(h, c->d_ostats, sizeof(OrbitStats), cudaMemcpyDeviceToHost, c->stream));
	CMCB_CUDA_OK(cudaMemcpyAsync(hr, b.result, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
	CMCB_CUDA_OK(cudaStreamSynchronize(c->stream));
	if (h->n_status[ORBIT_NTRY]) { // eprintf + exit_cleanly(-2) in the reference (cmc_utils.c:144-152)
		set_error("%s: binary search (FindZero_r) failed for %llu stars", what, h->n_status[ORBIT_NTRY]);
		return -2;
	}
	if (Eexcess_left) *Eexcess_left = *hr;
	return 0;
}

extern "C" int cmcb200_dev_set_velocities3(cmcb200_ctx *c, const cmcb200_orbit_params *p, double *Eexcess_left)
{
	CMCB_CUDA_OK(cudaSetDevice(c->device));
	if (join_uploads(c)) return -1;
	OrbitArgs a;
	if (orbit_begin(c, p, false, false, a)) return -1;
	Sv3Buffers b;
	if (sv3_buffers(c, b)) return -1;
	{
		StageTimer t(c, ST_ENERGY);
		const long n_pad = (p->n_local + SV3_BLOCK - 1) / SV3_BLOCK * SV3_BLOCK, nb = n_pad / SV3_BLOCK;
		c->launches += orbit_reset_stats(a, c->stream);
		if (nb > 0) {
			sv3_prepare_kernel<<<(unsigned) persistent_blocks(n_pad), ORBIT_TRY_THREADS, a.r_top_bytes, c->stream>>>(a, b, n_pad);
			sv3_block_kernel<<<(unsigned) ((nb + 7) / 8), 256, 0, c->stream>>>(a, b, nb);
		}
		sv3_carry_kernel<<<1, 32, 0, c->stream>>>(a, b, nb);
		if (nb > 0) sv3_apply_kernel<<<(unsigned) ((nb + 7) / 8), 256, 0, c->stream>>>(a, b, nb);
		c->launches += nb > 0 ? 4 : 1;
	}
	return sv3_finish(c, b, Eexcess_left, "set_velocities3");
}

extern "C" int cmcb200_dev_set_velocities3_pass(cmcb200_ctx *c, const cmcb200_orbit_params *p, double Eexcess_in, double *Eexcess_left)
{
	CMCB_CUDA_OK(cudaSetDevice(c->device));
	if (join_uploads(c)) return -1;
	if (!c->d_sv3) { set_error("set_velocities3_pass: call cmcb200_dev_set_velocities3 first (it records which stars the second loop corrects)"); return -1; }
	OrbitArgs a;
	if (orbit_begin(c, p, false, false, a)) return -1;
	Sv3Buffers b;
	if (sv3_buffers(c, b)) return -1;
	{
		StageTimer t(c, ST_ENERGY);
		c->launches += orbit_reset_stats(a, c->stream);
		sv3_second_kernel<<<1, 32, 0, c->stream>>>(a, b, Eexcess_in);
		c->launches += 1;
	}
	return sv3_finish(c, b, Eexcess_left, "set_velocities3");
}

#include <algorithm>

extern "C" int cmcb200_dev_potential(cmcb200_ctx *c, const cmcb200_orbit_params *p, const double *r, long n, double *out)
{
	CMCB_CUDA_OK(cudaSetDevice(c->device));
	if (join_uploads(c)) return -1;
	if (n <= 0) return 0;
	OrbitArgs a;
	if (orbit_begin(c, p, false, false, a)) return -1;
	if (ensure((void **) &c->d_draws, &c->draws_cap, 2 * n, sizeof(double))) return -1;
	double *d_in = c->d_draws, *d_out = c->d_draws + n;
	CMCB_CUDA_OK(cudaMemcpyAsync(d_in, r, (size_t) n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
	{
		StageTimer t(c, ST_ENERGY);
		c->launches += orbit_reset_stats(a, c->stream);
		potential_at_kernel<<<(unsigned) persistent_blocks(n), ORBIT_TRY_THREADS, a.r_top_bytes, c->stream>>>(a, d_in, n, d_out);
		c->launches += 1;
	}
	CMCB_CUDA_OK(cudaGetLastError());
	OrbitStats *h = (OrbitStats *) c->h_pinned;
	CMCB_CUDA_OK(cudaMemcpyAsync(h, c->d_ostats, sizeof(OrbitStats), cudaMemcpyDeviceToHost, c->stream));
	CMCB_CUDA_OK(cudaMemcpyAsync(out, d_out, (size_t) n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
	CMCB_CUDA_OK(cudaStreamSynchronize(c->stream));
	if (h->n_status[ORBIT_NTRY]) { set_error("potential: binary search (FindZero_r) failed for %llu radii", h->n_status[ORBIT_NTRY]); return -2; }
	return 0;
}

extern "C" int cmcb200_dev_tidally_strip_select(cmcb200_ctx *c, const cmcb200_orbit_params *p, int tidal_treatment, double threshold, long *idx_out,
						 long cap, long *count, void *star_out, void *binary_out, double *m_out)
{
	CMCB_CUDA_OK(cudaSetDevice(c->device));
	if (join_uploads(c)) return -1;
	*count = 0;
	if (tidal_treatment != 0 && tidal_treatment != 1) return 0; // the reference's loop body has no other branch (:476, :551)
	if (p->n_local <= 0 || cap <= 0) return 0;
	if (p->n_local + 2 > c->n_star_dim) { set_error("tidally_strip_stars: n_local exceeds the capacity"); return -1; }
	OrbitArgs a;
	memset(&a, 0, sizeof(a));
	a.star1 = star1(c);
	a.n_local = p->n_local;
	a.n_old_local = p->n_old_local;
	a.first_global = p->first_global;
	// queue: the deferred-scan list (ORBIT_DEFER_CAP longs) is free between orbit steps; the masses go to the list of found indices next to it
	const long qcap = cap < ORBIT_DEFER_CAP ? cap : ORBIT_DEFER_CAP;
	unsigned *d_count = c->d_defer_first;
	CMCB_CUDA_OK(cudaMemsetAsync(d_count, 0, sizeof(unsigned), c->stream));
	{
		StageTimer t(c, ST_ENERGY);
		tidal_select_kernel<<<(unsigned) persistent_blocks(p->n_local), 256, 0, c->stream>>>(a, tidal_treatment, threshold, c->d_defer_star, d_count, (unsigned) qcap);
		c->launches += 1;
	}
	CMCB_CUDA_OK(cudaGetLastError());
	unsigned *h_count = (unsigned *) c->h_pinned;
	CMCB_CUDA_OK(cudaMemcpyAsync(h_count, d_count, sizeof(unsigned), cudaMemcpyDeviceToHost, c->stream));
	CMCB_CUDA_OK(cudaStreamSynchronize(c->stream));
	const long n = *h_count;
	// nothing has been stored yet: the caller may come back with a longer list
	if (n > qcap) { set_error("tidally_strip_stars: %ld stars meet the criterion in one pass, the caller's list holds %ld (no record was changed)", n, qcap); *count = n; return -3; }
	if (n == 0) return 0;
	CMCB_CUDA_OK(cudaMemcpyAsync(idx_out, c->d_defer_star, (size_t) n * sizeof(long), cudaMemcpyDeviceToHost, c->stream));
	CMCB_CUDA_OK(cudaStreamSynchronize(c->stream));
	std::sort(idx_out, idx_out + n); // star order
	*count = n;
	double *d_mass = (double *) c->d_defer_found;
	CMCB_CUDA_OK(cudaMemcpyAsync(c->d_defer_star, idx_out, (size_t) n * sizeof(long), cudaMemcpyHostToDevice, c->stream));
	{
		StageTimer t(c, ST_ENERGY);
		tidal_apply_kernel<<<(unsigned) ((n + 127) / 128), 128, 0, c->stream>>>(a, c->d_defer_star, n, p->N_MAX, c->d_m, d_mass);
		c->launches += 1;
	}
	CMCB_CUDA_OK(cudaGetLastError());
	if (m_out) CMCB_CUDA_OK(cudaMemcpyAsync(m_out, d_mass, (size_t) n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
	if (star_out) {
		StageTimer t(c, ST_D2H);
		for (long k = 0; k < n; k++)
			CMCB_CUDA_OK(cudaMemcpyAsync((char *) star_out + k * CMCB_SIZEOF_STAR, star1(c) + (idx_out[k] - 1) * CMCB_SIZEOF_STAR, CMCB_SIZEOF_STAR, cudaMemcpyDeviceToHost, c->stream));
		CMCB_CUDA_OK(cudaStreamSynchronize(c->stream));
		if (binary_out) {
			for (long k = 0; k < n; k++) {
				const long b = *(const long *) ((const char *) star_out + k * CMCB_SIZEOF_STAR + CMCB_OFF_BININD);
				if (b > 0 && b < c->n_bin_dim)
					CMCB_CUDA_OK(cudaMemcpyAsync((char *) binary_out + k * CMCB_SIZEOF_BINARY, bin1(c) + (b - 1) * CMCB_SIZEOF_BINARY, CMCB_SIZEOF_BINARY, cudaMemcpyDeviceToHost, c->stream));
			}
		}
	}
	CMCB_CUDA_OK(cudaStreamSynchronize(c->stream));
	return 0;
}

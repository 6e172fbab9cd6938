// The per-star new-orbit step on the device: get_positions / get_positions_loop
// (reference src/cmc/cmc_evolution_thr.c:784-1023).  The per-star arithmetic is orbit_core.h (shared with the
// host-side emulation the CPU tests run); this file is the data movement around it.
//
// v1 mapping: one thread per star, consecutive threads = consecutive stars in radius order, so the first probes of
// the two index bisections (brackets [0, k] and [k, N_MAX+1] for neighbouring k) fall into the same cache lines.
// The r/Phi tables are read through the read-only path; they fit the 126 MB L2 up to N ~ 8e6 (16 B per star).
// Bound: L2 / dependent-load latency and the FP64 pipe -- not HBM.
#include "orbit.cuh"

namespace cmcb {

// Philox4x32-10 (Salmon et al., SC'11), counter = (star index lo, hi, tcount lo, block of four draws), key = seed.
struct PhiloxDraws {
	unsigned key0, key1, c0, c1, c2, c3;
	unsigned out[4];
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
		out[0] = x0; out[1] = x1; out[2] = x2; out[3] = x3;
		c3++;
		have = 4;
	}
	__device__ bool has(int) const { return true; }
	__device__ double next()
	{
		if (!have) refill();
		unsigned v = out[4 - have];
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

template <bool RECORDED>
__global__ void __launch_bounds__(128) orbit_step_kernel(OrbitArgs a)
{
	__shared__ unsigned long long s_max[4];
	const long j0 = blockIdx.x * (long) blockDim.x + threadIdx.x;
	unsigned long long my_max = 0, my_draws = 0, my_exact = 0, my_brent = 0;
	int my_status = -1;
	if (j0 < a.n_local) {
		const long j = j0 + 1; // the reference's 1-based local index
		const bool is_new = j > a.n_old_local;
		const long g = is_new ? 1 + a.t.N_MAX + (j - a.n_old_local) : a.first_global + j - 1;
		char *rec = a.star1 + j0 * CMCB_SIZEOF_STAR;
		const double E = ld_field(rec, CMCB_OFF_E), J = ld_field(rec, CMCB_OFF_J);
		const double vr = ld_field(rec, CMCB_OFF_VR), vt = ld_field(rec, CMCB_OFF_VT);
		const double rnew_of_new = is_new ? ld_field(rec, CMCB_OFF_RNEW) : 0.0;
		OrbitOut o;
		if (RECORDED) {
			CsrDraws dr;
			dr.d = a.draws; dr.pos = a.draw_off[j0]; dr.end = a.draw_off[j0 + 1];
			orbit_step_star(a.t, g, is_new, rnew_of_new, E, J, vr, vt, a.Rtidal, dr, o);
		} else {
			PhiloxDraws dr;
			dr.init(a.seed, a.tcount, (unsigned long long) g);
			orbit_step_star(a.t, g, is_new, rnew_of_new, E, J, vr, vt, a.Rtidal, dr, o);
		}
		// what the loop body writes into star[j] (cmc_evolution_thr.c:835-839, 852-853, 918-931; remove_star :662-664)
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
		if (o.status == ORBIT_OK) my_max = dbits(o.rnew);
		my_draws = o.n_draws; my_exact = o.n_exact; my_brent = o.n_brent;
		my_status = o.status;
	}
	// block reductions -> one atomic per block and counter
	for (int d = 16; d > 0; d >>= 1) {
		unsigned long long m = __shfl_xor_sync(0xffffffffu, my_max, d);
		my_max = m > my_max ? m : my_max;
		my_draws += __shfl_xor_sync(0xffffffffu, my_draws, d);
		my_exact += __shfl_xor_sync(0xffffffffu, my_exact, d);
		my_brent += __shfl_xor_sync(0xffffffffu, my_brent, d);
	}
	if (lane_id() == 0) {
		s_max[warp_id()] = my_max;
		if (my_draws) atomicAdd(&a.stats->n_draws, my_draws);
		if (my_exact) atomicAdd(&a.stats->n_exact, my_exact);
		if (my_brent) atomicAdd(&a.stats->n_brent, my_brent);
	}
	if (my_status > 0) atomicAdd(&a.stats->n_status[my_status], 1ull); // the rare outcomes; OK count = n - sum
	__syncthreads();
	if (threadIdx.x == 0) {
		unsigned long long m = s_max[0];
		for (int w = 1; w < 4; w++) m = s_max[w] > m ? s_max[w] : m;
		if (m) atomicMax(&a.stats->max_rad_bits, m);
	}
}

__global__ void orbit_stats_reset(OrbitStats *s)
{
	s->max_rad_bits = 0; s->n_draws = 0; s->n_exact = 0; s->n_brent = 0;
	for (int i = 0; i < 6; i++) s->n_status[i] = 0;
}

int orbit_step_device(const OrbitArgs &a, cudaStream_t s)
{
	orbit_stats_reset<<<1, 1, 0, s>>>(a.stats);
	if (a.n_local <= 0) return 1;
	unsigned blocks = (unsigned) ((a.n_local + 127) / 128);
	if (a.draws) orbit_step_kernel<true><<<blocks, 128, 0, s>>>(a);
	else orbit_step_kernel<false><<<blocks, 128, 0, s>>>(a);
	return 2;
}

} // namespace cmcb

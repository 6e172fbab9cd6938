// Potential rebuild on the device: potential_calculate (reference src/cmc/cmc_utils.c:675-736).
//
//   N_MAX      = number of leading stars with r < SF_INFINITY                       (:685-696)
//   mass sum   = sequential sum of star_m[1..N_MAX]                                 (:686)      chain F (+)
//   Mtotal     = sum*madhoc + cenma.m*madhoc, after cenma.m += cenma.m_new          (:699-703)
//   mprev_k    = Mtotal - sum_{j>k} star_m[j]/N_STAR, sequentially from k = N_MAX    (:716,720)  chain M (-)
//   phi_k      = phi_{k+1} - mprev_k (1/r_k - 1/r_{k+1}), phi_{N_MAX+1} = 0          (:717-719)  chain P (+, on -phi)
//   phi_0      = phi_1 + cenma.m*madhoc/r_1                                          (:729)
// All three recurrences are evaluated with the reference's sequential roundings by seqscan.cuh.
#include "seqscan.cuh"
#include "potential.cuh"

namespace cmcb {

// ---- chains -------------------------------------------------------------------------------------------------

struct ChainMass { // y_i = y_{i-1} + star_m[i+1], i = 0..N_MAX-1
	static constexpr int SIGN = +1;
	const double *m;
	const PotScalars *sc;
	__device__ long length() const { return sc->n_max; }
	__device__ double y0() const { return 0.0; }
	__device__ double term(long i) const { return m[i + 1]; }
	__device__ void store(long, double) const {}
};

struct ChainMprev { // element i <-> k = N_MAX - i (k = N_MAX..2): mprev_{k-1} = mprev_k - star_m[k]/N_STAR
	static constexpr int SIGN = -1;
	const double *m;
	double *mprev; // mprev[k] = value used at k
	const PotScalars *sc;
	double n_star;
	__device__ long length() const { return sc->n_max > 0 ? sc->n_max - 1 : 0; }
	__device__ double y0() const { return sc->Mtotal; }
	__device__ double term(long i) const { return ddiv(m[sc->n_max - i], n_star); }
	__device__ void store(long i, double y) const { mprev[sc->n_max - i - 1] = y; }
};

struct ChainPhi { // element i <-> k = N_MAX - i (k = N_MAX..1): -phi_k = -phi_{k+1} + mprev_k (1/r_k - 1/r_{k+1})
	static constexpr int SIGN = +1;
	const double *r, *mprev;
	double *phi;
	const PotScalars *sc;
	__device__ long length() const { return sc->n_max; }
	__device__ double y0() const { return 0.0; }
	__device__ double term(long i) const
	{
		long k = sc->n_max - i;
		return dmul(mprev[k], dsub(ddiv(1.0, r[k]), ddiv(1.0, r[k + 1])));
	}
	__device__ void store(long i, double y) const { phi[sc->n_max - i] = dsub(0.0, y); } // 0.0 - y keeps +0.0 like the reference
};

// ---- small kernels ------------------------------------------------------------------------------------------

__global__ void pot_init_scalars(PotScalars *sc, long n_max_in, double cenma_m, double cenma_m_new)
{
	sc->n_max = n_max_in; // lowered by pot_find_nmax
	sc->first_inf = n_max_in + 1;
	sc->cenma_m = cenma_m;
	sc->cenma_m_new = cenma_m_new;
	sc->mass_sum = 0.0;
	sc->Mtotal = 0.0;
	sc->phi_sum = 0.0;
}

// first k in [1, n_max_in] with !(star_r[k] < SF_INFINITY)   (:685)
__global__ void pot_find_nmax(const double *__restrict__ r, long n_max_in, PotScalars *sc)
{
	long k = 1 + blockIdx.x * (long) blockDim.x + threadIdx.x;
	long stride = (long) gridDim.x * blockDim.x;
	long best = n_max_in + 1;
	for (; k <= n_max_in; k += stride)
		if (!(r[k] < 1.0e10)) { best = k; break; } // ascending per thread: the first hit is this thread's minimum
	for (int d = 16; d > 0; d >>= 1) {
		long o = __shfl_xor_sync(0xffffffffu, best, d);
		best = o < best ? o : best;
	}
	if (lane_id() == 0 && best <= n_max_in) atomicMin((unsigned long long *) &sc->first_inf, (unsigned long long) best);
}

__global__ void pot_set_nmax(PotScalars *sc) { sc->n_max = sc->first_inf - 1; }

// :696-714 and the start value of the mprev chain
__global__ void pot_totals(PotScalars *sc, double *r, double *phi, double *mprev, double madhoc)
{
	long n = sc->n_max;
	sc->cenma_m = dadd(sc->cenma_m, sc->cenma_m_new);
	sc->cenma_m_new = 0.0;
	sc->Mtotal = dadd(dmul(sc->mass_sum, madhoc), dmul(sc->cenma_m, madhoc));
	r[n + 1] = 1.0e10;
	phi[n + 1] = 0.0;
	mprev[n] = sc->Mtotal; // value used at k = N_MAX (index 0 is a harmless slot when N_MAX == 0)
}

__global__ void pot_phi0(const PotScalars *sc, const double *r, double *phi, double madhoc)
{
	phi[0] = dadd(phi[1], ddiv(dmul(sc->cenma_m, madhoc), r[1])); // :729
}

// ---- host side ----------------------------------------------------------------------------------------------

size_t pot_work_bytes(long n_dim)
{
	long ntiles = (n_dim + SEQ_TILE - 1) / SEQ_TILE + 1;
	size_t b = 0;
	b += 256;                                         // counter, fallback, stats
	b += ntiles * (sizeof(int) * 3 + sizeof(double) * 5);
	b += ntiles * (size_t) SEQ_EV_PER_TILE * sizeof(SeqEvent);
	b += (size_t) SEQ_EV_MAX * 3 * sizeof(double);
	return b + 4096;
}

// carve the SeqWork arrays out of one allocation
void pot_work_bind(SeqWork &w, void *base, long n_dim)
{
	long ntiles = (n_dim + SEQ_TILE - 1) / SEQ_TILE + 1;
	char *p = (char *) base;
	auto take = [&](size_t bytes) { char *q = p; p += (bytes + 255) & ~(size_t) 255; return (void *) q; };
	w.tile_counter = (unsigned *) take(sizeof(unsigned));
	w.fallback = (int *) take(sizeof(int));
	w.stats = (long *) take(2 * sizeof(long));
	w.tile_evcount = (int *) take(ntiles * sizeof(int));
	w.tile_evoff = (int *) take(ntiles * sizeof(int));
	w.tile_desc = (double2 *) take(ntiles * sizeof(double2));
	w.tile_prefix = (double *) take(ntiles * sizeof(double));
	w.tile_tail = (double *) take(ntiles * sizeof(double));
	w.tile_carry = (double *) take(ntiles * sizeof(double));
	w.tile_events = (SeqEvent *) take(ntiles * (size_t) SEQ_EV_PER_TILE * sizeof(SeqEvent));
	w.ev_value = (double *) take(SEQ_EV_MAX * sizeof(double));
	w.ev_a = (double *) take(SEQ_EV_MAX * sizeof(double));
	w.ev_seg = (double *) take(SEQ_EV_MAX * sizeof(double));
	w.max_tiles = ntiles;
}

int potential_calculate_device(const PotArgs &a, cudaStream_t st)
{
	const int grid_persist = NUM_SMS_B200 * 4;
	pot_init_scalars<<<1, 1, 0, st>>>(a.sc, a.n_max_in, a.cenma_m, a.cenma_m_new);
	if (a.n_max_in > 0) {
		int blocks = (int) ((a.n_max_in + 255) / 256);
		if (blocks > NUM_SMS_B200 * 8) blocks = NUM_SMS_B200 * 8;
		pot_find_nmax<<<blocks, 256, 0, st>>>(a.r, a.n_max_in, a.sc);
	}
	pot_set_nmax<<<1, 1, 0, st>>>(a.sc);
	ChainMass cf{a.m, a.sc};
	seq_scan_launch(cf, a.work, a.n_max_in, &a.sc->mass_sum, grid_persist, grid_persist, st, false); // only the sum is used (Mtotal): no second pass
	pot_totals<<<1, 1, 0, st>>>(a.sc, a.r, a.phi, a.mprev, a.madhoc);
	ChainMprev cm{a.m, a.mprev, a.sc, a.n_star};
	seq_scan_launch(cm, a.work, a.n_max_in, &a.sc->mprev_end, grid_persist, grid_persist, st);
	ChainPhi cp{a.r, a.mprev, a.phi, a.sc};
	seq_scan_launch(cp, a.work, a.n_max_in, &a.sc->phi_sum, grid_persist, grid_persist, st);
	pot_phi0<<<1, 1, 0, st>>>(a.sc, a.r, a.phi, a.madhoc);
	return (a.n_max_in > 0 ? 5 : 4) + 3 * 4 - 1; // scalars, find, set, totals, phi0 + 3 chains x (reset, classify, resolve, finalize), the mass sum without its finalize
}

} // namespace cmcb

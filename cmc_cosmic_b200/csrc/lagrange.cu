// Lagrange radii on the resident arrays: comp_mass_percent (reference src/cmc/cmc_utils.c:872-985) -- SURVEY.md 8f N2, the first
// consumer of the freshly sorted tables after calc_potential_new and the observable of the "statistically agree" check.
//
//   mprev_k    = mprev_0 + sum_{j<=k} star_m[j]/N_STAR, sequentially from k = 1 (:934)               chain E (+), replicated tables
//   ke_rad_l   = sum_{j<=l} 0.5 star_m[g_j] madhoc vr_j^2, rank-local, sequentially from l = 1 (:907-913) and the same for
//   ke_tan, v2_rad, v2_tan                                                                            chains K0..K3 (+), local records
//   entry mcount is recorded at the first k (one threshold per k) with mprev_k > mass_pc[mcount] Mtotal (:936-977)
// All five recurrences carry the reference's sequential roundings (seqscan.cuh), so the radii, star counts and sums are the
// reference's bit for bit.  The sums are returned rank-local; the caller adds the MPI_Exscan prefix for indices a rank other than 0
// owns, exactly like the reference (:945-951).  pow() for the densities runs on the host (glibc, as in the reference).
#include <cmath>
#include <cstring>
#include <vector>
#include "ctx.cuh"
#include "seqscan.cuh"

namespace cmcb {

constexpr int MASS_PC_MAX = 256;

struct LagrangeDev { // device-resident result of one call
	long n_filled;
	long k[MASS_PC_MAX];
	double mprev[MASS_PC_MAX], r[MASS_PC_MAX];
	double sums[4][MASS_PC_MAX];
	int owner[MASS_PC_MAX];
	double totals[4];
	double thr[MASS_PC_MAX]; // mass_pc (in)
	double last;             // seq_scan_resolve's slot for a chain's final value (not used here)
};

struct ChainEnclosed { // y_i = y_{i-1} + star_m[i+1] / N_STAR, i = 0..N_MAX-1; stored at mcum[i+1]
	static constexpr int SIGN = +1;
	const double *m;
	double *mcum;
	long n;
	double n_star, start;
	__device__ long length() const { return n; }
	__device__ double y0() const { return start; }
	__device__ double term(long i) const { return ddiv(m[i + 1], n_star); }
	__device__ void store(long i, double y) const { mcum[i + 1] = y; }
};

// WHICH: 0 ke_rad, 1 ke_tan, 2 v2_rad, 3 v2_tan.  Element i <-> local star i + 1 (record i of star1)
template <int WHICH>
struct ChainKinetic {
	static constexpr int SIGN = +1;
	const char *star1;
	const double *m;
	double *acc; // acc[l] = sum over local stars 1..l
	long n, first_global;
	double madhoc;
	__device__ long length() const { return n; }
	__device__ double y0() const { return 0.0; }
	__device__ double term(long i) const
	{
		const double v = *(const double *) (star1 + i * CMCB_SIZEOF_STAR + ((WHICH & 1) ? CMCB_OFF_VT : CMCB_OFF_VR));
		if (WHICH >= 2) return dmul(v, v);
		return dmul(dmul(dmul(dmul(0.5, m[first_global + i]), madhoc), v), v); // 0.5 * star_m[g_k] * madhoc * v * v, left to right
	}
	__device__ void store(long i, double y) const { acc[i + 1] = y; }
};

// one thread: the threshold loop (:933-977) on the finished mprev array
__global__ void lagrange_select(const double *__restrict__ mcum, const double *__restrict__ star_r, long N_MAX, long count, long mcount0, double Mtotal,
				LagrangeDev *out)
{
	long mcount = mcount0, kprev = 0;
	for (; mcount < count; mcount++) {
		const double thr = dmul(out->thr[mcount], Mtotal);
		// mcum is non-decreasing (sequentially rounded sums of non-negative terms): first k in (kprev, N_MAX] with mcum[k] > thr
		long lo = kprev + 1, hi = N_MAX + 1;
		while (lo < hi) {
			const long mid = (lo + hi) >> 1;
			if (mcum[mid] > thr) hi = mid; else lo = mid + 1;
		}
		if (lo > N_MAX) break;
		out->k[mcount] = lo; out->mprev[mcount] = mcum[lo]; out->r[mcount] = star_r[lo];
		out->owner[mcount] = 0;
		kprev = lo;
	}
	out->n_filled = mcount;
}

// clusdyn_calculate (:1613-1624): `while (m < 0.5 Mtotal) { m += star_m[k] / N_STAR; k++; }  rh = star_r[k]` on the finished sums
__global__ void lagrange_half_mass(const double *__restrict__ mcum, const double *__restrict__ star_r, long N_MAX, double Mtotal, double *rh)
{
	const double half = dmul(0.5, Mtotal);
	long lo = 1, hi = N_MAX + 1; // first k with !(mcum[k] < half); mcum is non-decreasing
	if (!(0.0 < half)) { *rh = star_r[1]; return; } // the loop body never runs: k stays 1
	while (lo < hi) {
		const long mid = (lo + hi) >> 1;
		if (mcum[mid] < half) lo = mid + 1; else hi = mid;
	}
	*rh = star_r[(lo <= N_MAX ? lo : N_MAX) + 1]; // k was incremented once more after the add that reached the half
}

// the cumulative sum `which` at the recorded indices this rank owns, and the rank's total
__global__ void lagrange_pick(const double *__restrict__ acc, long n_local, long first_global, long mcount0, int which, LagrangeDev *out)
{
	const long i = mcount0 + threadIdx.x;
	if (i < out->n_filled) {
		const long k = out->k[i];
		if (k >= first_global && k < first_global + n_local) {
			out->sums[which][i] = acc[k - first_global + 1];
			out->owner[i] = 1;
		}
	}
	if (threadIdx.x == 0) out->totals[which] = n_local > 0 ? acc[n_local] : 0.0;
}

} // namespace cmcb

extern "C" int cmcb200_dev_clusdyn_calculate(cmcb200_ctx *c, long N_MAX, long N_STAR, double Mtotal, double *rh)
{
	if (N_MAX + 2 > c->n_table_dim) { set_error("clusdyn_calculate: N_MAX + 2 exceeds the table capacity"); return -1; }
	CMCB_CUDA_OK(cudaSetDevice(c->device));
	if (join_uploads(c)) return -1;
	if (!c->d_lagrange) CMCB_CUDA_OK(cudaMalloc(&c->d_lagrange, sizeof(LagrangeDev)));
	LagrangeDev *d = (LagrangeDev *) c->d_lagrange;
	{
		StageTimer t(c, ST_LAGRANGE);
		ChainEnclosed ce{c->d_m, c->d_mprev, N_MAX, (double) N_STAR, 0.0};
		seq_scan_launch(ce, c->seqwork, N_MAX, &d->last, NUM_SMS_B200 * 4, NUM_SMS_B200 * 4, c->stream);
		lagrange_half_mass<<<1, 1, 0, c->stream>>>(c->d_mprev, c->d_r, N_MAX, Mtotal, &d->totals[0]);
		c->launches += 5;
	}
	CMCB_CUDA_OK(cudaGetLastError());
	CMCB_CUDA_OK(cudaMemcpyAsync(c->h_pinned, &d->totals[0], sizeof(double), cudaMemcpyDeviceToHost, c->stream));
	CMCB_CUDA_OK(cudaStreamSynchronize(c->stream));
	*rh = *(double *) c->h_pinned;
	return 0;
}

extern "C" int cmcb200_dev_comp_mass_percent(cmcb200_ctx *c, const cmcb200_mass_pc_params *p, cmcb200_mass_pc_result *res)
{
	if (p->count < 0 || p->count > MASS_PC_MAX) { set_error("comp_mass_percent: MASS_PC_COUNT = %ld exceeds %d", p->count, MASS_PC_MAX); return -1; }
	if (p->N_MAX + 2 > c->n_table_dim || p->n_local + 2 > c->n_star_dim) { set_error("comp_mass_percent: sizes exceed the context capacity"); return -1; }
	CMCB_CUDA_OK(cudaSetDevice(c->device));
	if (join_uploads(c)) return -1;
	if (!c->d_lagrange) CMCB_CUDA_OK(cudaMalloc(&c->d_lagrange, sizeof(LagrangeDev)));
	LagrangeDev *d = (LagrangeDev *) c->d_lagrange;
	const double madhoc = 1.0 / ((double) p->N_STAR);
	// :876-895 the thresholds the central mass alone exceeds
	double mprev0 = 0.0;
	long mcount0 = 0;
	if (p->bh_include) {
		mprev0 = p->cenma_m * madhoc;
		for (mcount0 = 0; mcount0 < p->count; mcount0++) {
			if (mprev0 / p->Mtotal > p->mass_pc[mcount0]) {
				res->mass_r[mcount0] = p->minimum_r; res->ave_mass_r[mcount0] = 0.0; res->no_star_r[mcount0] = 0; res->densities_r[mcount0] = 0.0;
				res->ke_rad_r[mcount0] = 0.0; res->ke_tan_r[mcount0] = 0.0; res->v2_rad_r[mcount0] = 0.0; res->v2_tan_r[mcount0] = 0.0;
				if (res->owner) res->owner[mcount0] = 0;
			} else break;
		}
	}
	CMCB_CUDA_OK(cudaMemsetAsync(d, 0, sizeof(LagrangeDev), c->stream));
	if (p->count > 0) CMCB_CUDA_OK(cudaMemcpyAsync(d->thr, p->mass_pc, (size_t) p->count * sizeof(double), cudaMemcpyHostToDevice, c->stream));
	const int grid = NUM_SMS_B200 * 4;
	double *mcum = c->d_mprev;                 // scratch of the potential rebuild, free between its calls
	double *acc = (double *) c->sbuf.keys_a;   // scratch of the sort, free between its calls (n_star_dim doubles)
	double *dummy = &d->last;                  // seq_scan_launch wants a slot for the chain's last value
	{
		StageTimer t(c, ST_LAGRANGE);
		ChainEnclosed ce{c->d_m, mcum, p->N_MAX, (double) p->N_STAR, mprev0};
		seq_scan_launch(ce, c->seqwork, p->N_MAX, dummy, grid, grid, c->stream);
		lagrange_select<<<1, 1, 0, c->stream>>>(mcum, c->d_r, p->N_MAX, p->count, mcount0, p->Mtotal, d);
		c->launches += 5;
		const char *s1 = star1(c);
		auto run = [&](auto chain, int which) {
			seq_scan_launch(chain, c->seqwork, p->n_local, dummy, grid, grid, c->stream);
			lagrange_pick<<<1, MASS_PC_MAX, 0, c->stream>>>(acc, p->n_local, p->first_global, mcount0, which, d);
			c->launches += 5;
		};
		run(ChainKinetic<0>{s1, c->d_m, acc, p->n_local, p->first_global, madhoc}, 0);
		run(ChainKinetic<1>{s1, c->d_m, acc, p->n_local, p->first_global, madhoc}, 1);
		run(ChainKinetic<2>{s1, c->d_m, acc, p->n_local, p->first_global, madhoc}, 2);
		run(ChainKinetic<3>{s1, c->d_m, acc, p->n_local, p->first_global, madhoc}, 3);
	}
	CMCB_CUDA_OK(cudaGetLastError());
	static thread_local LagrangeDev h;
	CMCB_CUDA_OK(cudaMemcpyAsync(&h, d, sizeof(LagrangeDev), cudaMemcpyDeviceToHost, c->stream));
	CMCB_CUDA_OK(cudaStreamSynchronize(c->stream));
	for (long i = mcount0; i < h.n_filled; i++) { // :937-940, with the host's libm like the reference
		const long k = h.k[i];
		res->mass_r[i] = h.r[i];
		res->ave_mass_r[i] = h.mprev[i] / p->Mtotal / k * p->initial_total_mass;
		res->no_star_r[i] = (double) k;
		res->densities_r[i] = h.mprev[i] / (4.0 / 3.0 * 3.14159265358979323 /* PI, include/cmc/cmc.h:62 */ * pow(h.r[i], 3));
		if (h.owner[i]) { res->ke_rad_r[i] = h.sums[0][i]; res->ke_tan_r[i] = h.sums[1][i]; res->v2_rad_r[i] = h.sums[2][i]; res->v2_tan_r[i] = h.sums[3][i]; }
		if (res->owner) res->owner[i] = h.owner[i];
	}
	for (int q = 0; q < 4; q++) res->totals[q] = h.totals[q];
	res->filled = h.n_filled;
	return 0;
}

// ---- ComputeEnergy's second loop (reference src/cmc/cmc_utils.c:619-636): the rank's energy sums ---------------------------------------------
// Four sequentially summed scalars over the local stars: kinetic, potential (two additions per star), internal, binary binding.  Each is a
// chain of same-signed terms, summed on its magnitudes (fl(-x - a) = -fl(x + a)) through the exact scan; an element of the other sign
// is an event and takes the hardware rounding.  Only the final values are wanted: store() is empty.
namespace cmcb {

struct EtotArgs {
	const char *star1, *bin1;
	const double *m, *phi;
	long n, first_global;
	double madhoc, cenma_m, n_max;
};
// WHICH: 0 kinetic, 1 potential (2 elements per star), 2 internal, 3 binding, 4 potential without a central mass (its share of every star is
// +-0: one element per star)
template <int WHICH>
struct ChainEtot {
	static constexpr int SIGN = +1;
	EtotArgs a;
	__device__ long length() const { return WHICH == 1 ? 2 * a.n : a.n; }
	__device__ double y0() const { return 0.0; }
	__device__ void store(long, double) const {}
	__device__ double term(long e) const
	{
		const long i = WHICH == 1 ? e >> 1 : e;
		if (WHICH == 4) return -dmul(dmul(a.phi[a.first_global + i], a.m[a.first_global + i]), a.madhoc);                        // :622
		const char *rec = a.star1 + i * CMCB_SIZEOF_STAR;
		const long j = a.first_global + i;
		if (WHICH == 0) {
			const double vr = *(const double *) (rec + CMCB_OFF_VR), vt = *(const double *) (rec + CMCB_OFF_VT);
			return dmul(dmul(dmul(0.5, dadd(dmul(vr, vr), dmul(vt, vt))), a.m[j]), a.madhoc);                 // :621
		}
		if (WHICH == 1) {
			if (!(e & 1)) return -dmul(dmul(a.phi[j], a.m[j]), a.madhoc);                                     // :622
			return -ddiv(dmul(dmul(a.phi[0], a.cenma_m), a.madhoc), a.n_max);                                  // :623
		}
		const long b = *(const long *) (rec + CMCB_OFF_BININD);
		if (b == 0) return WHICH == 2 ? *(const double *) (rec + CMCB_OFF_EINT) : 0.0;                            // :625-626
		const char *bin = a.bin1 + (b - 1) * CMCB_SIZEOF_BINARY;
		if (!*(const int *) (bin + 80 /* inuse */)) return 0.0;                                                    // :627
		if (WHICH == 2) return dadd(*(const double *) (bin + 48 /* Eint1 */), *(const double *) (bin + 56 /* Eint2 */)); // :630
		return ddiv(dmul(dmul(*(const double *) (bin + 32 /* m1 */), a.madhoc), dmul(*(const double *) (bin + 40 /* m2 */), a.madhoc)),
			    dmul(2.0, *(const double *) (bin + 64 /* a */)));                                               // -(:628-629)
	}
};

} // namespace cmcb

extern "C" int cmcb200_dev_compute_energy_totals(cmcb200_ctx *c, long n_local, long first_global, long N_MAX, long N_STAR, double cenma_m, double out[4])
{
	if (n_local + 2 > c->n_star_dim || first_global + n_local > c->n_table_dim) { set_error("ComputeEnergy: sizes exceed the context capacity"); return -1; }
	CMCB_CUDA_OK(cudaSetDevice(c->device));
	if (join_uploads(c)) return -1;
	if (!c->d_lagrange) CMCB_CUDA_OK(cudaMalloc(&c->d_lagrange, sizeof(LagrangeDev)));
	LagrangeDev *d = (LagrangeDev *) c->d_lagrange;
	CMCB_CUDA_OK(cudaMemsetAsync(d->totals, 0, sizeof(d->totals), c->stream));
	CMCB_CUDA_OK(cudaMemsetAsync(c->seqwork.stats, 0, 2 * sizeof(long), c->stream)); // cmcb200_potential_stats then reports this call's chains
	EtotArgs a{star1(c), bin1(c), c->d_m, c->d_phi, n_local, first_global, 1.0 / ((double) N_STAR), cenma_m, (double) N_MAX};
	const int grid = NUM_SMS_B200 * 4;
	if (n_local > 0) {
		StageTimer t(c, ST_ENERGY);
		seq_scan_launch(ChainEtot<0>{a}, c->seqwork, n_local, &d->totals[0], grid, grid, c->stream, false);
		if (cenma_m == 0.0) seq_scan_launch(ChainEtot<4>{a}, c->seqwork, n_local, &d->totals[1], grid, grid, c->stream, false);
		else { // two additions per star: a scan twice as long as the tables the shared scratch was sized for
			if (!c->d_seqwork2) {
				const size_t wb = pot_work_bytes(2 * c->n_star_dim);
				CMCB_CUDA_OK(cudaMalloc(&c->d_seqwork2, wb));
				CMCB_CUDA_OK(cudaMemsetAsync(c->d_seqwork2, 0, wb, c->stream));
			}
			SeqWork w2;
			pot_work_bind(w2, c->d_seqwork2, 2 * c->n_star_dim);
			seq_scan_launch(ChainEtot<1>{a}, w2, 2 * n_local, &d->totals[1], grid, grid, c->stream, false);
		}
		seq_scan_launch(ChainEtot<2>{a}, c->seqwork, n_local, &d->totals[2], grid, grid, c->stream, false);
		seq_scan_launch(ChainEtot<3>{a}, c->seqwork, n_local, &d->totals[3], grid, grid, c->stream, false);
		c->launches += 12; // only the sums are wanted: no second pass over the terms
	}
	CMCB_CUDA_OK(cudaGetLastError());
	double *h = (double *) c->h_pinned;
	CMCB_CUDA_OK(cudaMemcpyAsync(h, d->totals, 4 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
	CMCB_CUDA_OK(cudaStreamSynchronize(c->stream));
	out[0] = h[0];
	out[1] = h[1] == 0.0 ? 0.0 : -h[1]; // the potential and binding sums were taken on their magnitudes
	out[2] = h[2];
	out[3] = h[3] == 0.0 ? 0.0 : -h[3];
	return 0;
}


// ---- comp_multi_mass_percent (reference src/cmc/cmc_utils.c:769-861): Lagrange radii per mass bin -----------------------------------------
// For every mass bin the reference sums the masses of the bin's stars in star order (twice: the total, then the running fraction), and
// interpolates radius over mass fraction linearly (gsl_interp_linear: un-vendored GSL, restated like the oracle restates it).  Here: one exact
// sequential-sum chain per bin over the whole table whose terms are star_m for the bin's stars and 0 for the others (a zero term is regular:
// fl(y + 0) = y), then per threshold two binary searches on the finished running sums.
namespace cmcb {

constexpr int MASS_BINS_MAX = 64;
struct MultiMassArgs {
	double bins[MASS_BINS_MAX]; // mass_bins[], ascending
	double thr[MASS_PC_MAX];    // mass_pc[]
	long n_bins, count;
	double solar_mass_dyn;
};

// find_stars_mass_bin (:756-764) -> m_binsearch (include/common/gensearch.h): largest index with mass_bins[index] < smass, -1 outside the bins
__device__ __forceinline__ long mass_bin_of(const double *bins, long n_bins, double smass)
{
	if (smass < bins[0] || smass > bins[n_bins - 1]) return -1;
	long lo = 0, hi = n_bins - 1;
	do {
		const long k = (lo + hi + 1) / 2;
		if (bins[k] < smass) lo = k; else hi = k - 1;
	} while (hi != lo);
	return lo;
}

struct ChainBinMass { // y_i = y_{i-1} + (star i+1 in bin b ? star_m[i+1] : 0); stored at ysum[i+1]
	static constexpr int SIGN = +1;
	const double *m;
	double *ysum;
	const MultiMassArgs *a;
	long n, b;
	__device__ long length() const { return n; }
	__device__ double y0() const { return 0.0; }
	__device__ double term(long i) const
	{
		const double mi = m[i + 1];
		return mass_bin_of(a->bins, a->n_bins, ddiv(mi, a->solar_mass_dyn)) == b ? mi : 0.0;
	}
	__device__ void store(long i, double y) const { ysum[i + 1] = y; }
};

// thread c: multi_mass_r[b][c] from the bin's running sums ysum[1..N_MAX] (non-decreasing; it steps exactly at the bin's stars)
__global__ void multi_mass_eval(const double *__restrict__ ysum, const double *__restrict__ star_r, long N_MAX, const MultiMassArgs *a, long b, double *out, int *filled)
{
	const long c = threadIdx.x;
	if (c >= a->count) return;
	const double total = N_MAX > 0 ? ysum[N_MAX] : 0.0;
	auto first_with = [&](auto pred) { // first k in [1, N_MAX] with pred(ysum[k]); N_MAX + 1 if none (pred is monotone along the sums)
		long lo = 1, hi = N_MAX + 1;
		while (lo < hi) {
			const long mid = (lo + hi) >> 1;
			if (pred(ysum[mid])) hi = mid; else lo = mid + 1;
		}
		return lo;
	};
	const long k_first = first_with([&](double y) { return y > 0.0; });
	if (k_first > N_MAX) { if (c == 0) filled[b] = 0; return; } // number_inbin == 0: the row is left untouched (:845-846)
	if (c == 0) filled[b] = 1;
	const long k_last = first_with([&](double y) { return y >= total; });
	if (k_first == k_last) { out[b * a->count + c] = star_r[k_last]; return; } // number_inbin == 1 (:847-850)
	const double x = a->thr[c];
	// gsl_interp_bsearch over percents[]: the first star of the bin whose fraction exceeds x closes the interval, the last one if none does
	long k_hi = first_with([&](double y) { return ddiv(y, total) > x; });
	if (k_hi > N_MAX) k_hi = k_last;
	const double y_before = ysum[k_hi - 1]; // the running sum up to the previous star of the bin (ysum[0] = 0 is never written: guard below)
	double x_lo = 0.0, r_lo = 0.0;
	if (k_hi > k_first) {
		const double yb = y_before;
		const long k_lo = first_with([&](double y) { return y >= yb; }); // where the sum reached that value: the previous star of the bin
		x_lo = ddiv(yb, total);
		r_lo = star_r[k_lo];
	}
	const double x_hi = ddiv(ysum[k_hi], total), r_hi = star_r[k_hi];
	const double dx = dsub(x_hi, x_lo);
	double val = dx > 0.0 ? dadd(r_lo, dmul(ddiv(dsub(x, x_lo), dx), dsub(r_hi, r_lo))) : 0.0;
	if (x < 0.0 || x > 1.0) val = __longlong_as_double(0x7ff8000000000000ll);
	out[b * a->count + c] = val;
}

} // namespace cmcb

extern "C" int cmcb200_dev_comp_multi_mass_percent(cmcb200_ctx *c, long N_MAX, const double *mass_bins, long n_bins, double solar_mass_dyn,
						    const double *mass_pc, long count, double *multi_mass_r)
{
	if (n_bins <= 1) return 0; // :786
	if (n_bins > MASS_BINS_MAX || count < 0 || count > MASS_PC_MAX) { set_error("comp_multi_mass_percent: at most %d bins and %d fractions", MASS_BINS_MAX, MASS_PC_MAX); return -1; }
	if (N_MAX + 2 > c->n_table_dim) { set_error("comp_multi_mass_percent: N_MAX + 2 exceeds the table capacity"); return -1; }
	CMCB_CUDA_OK(cudaSetDevice(c->device));
	if (join_uploads(c)) return -1;
	if (!c->d_lagrange) CMCB_CUDA_OK(cudaMalloc(&c->d_lagrange, sizeof(LagrangeDev)));
	if (!c->d_multimass) CMCB_CUDA_OK(cudaMalloc(&c->d_multimass, sizeof(MultiMassArgs) + (size_t) MASS_BINS_MAX * MASS_PC_MAX * sizeof(double) + MASS_BINS_MAX * sizeof(int)));
	LagrangeDev *d = (LagrangeDev *) c->d_lagrange;
	MultiMassArgs *d_a = (MultiMassArgs *) c->d_multimass;
	double *d_out = (double *) ((char *) c->d_multimass + sizeof(MultiMassArgs));
	int *d_filled = (int *) (d_out + (size_t) MASS_BINS_MAX * MASS_PC_MAX);
	MultiMassArgs h;
	memset(&h, 0, sizeof(h));
	for (long i = 0; i < n_bins; i++) h.bins[i] = mass_bins[i];
	for (long i = 0; i < count; i++) h.thr[i] = mass_pc[i];
	h.n_bins = n_bins; h.count = count; h.solar_mass_dyn = solar_mass_dyn;
	CMCB_CUDA_OK(cudaMemcpyAsync(d_a, &h, sizeof(h), cudaMemcpyHostToDevice, c->stream));
	CMCB_CUDA_OK(cudaStreamSynchronize(c->stream)); // h lives on this stack frame
	{
		StageTimer t(c, ST_LAGRANGE);
		for (long b = 0; b < n_bins; b++) {
			ChainBinMass ch{c->d_m, c->d_mprev, d_a, N_MAX, b}; // d_mprev: the potential's scratch, free between two rebuilds
			seq_scan_launch(ch, c->seqwork, N_MAX, &d->last, NUM_SMS_B200 * 4, NUM_SMS_B200 * 4, c->stream);
			if (count > 0) multi_mass_eval<<<1, MASS_PC_MAX, 0, c->stream>>>(c->d_mprev, c->d_r, N_MAX, d_a, b, d_out, d_filled);
			c->launches += 5;
		}
	}
	CMCB_CUDA_OK(cudaGetLastError());
	std::vector<double> out((size_t) n_bins * count);
	std::vector<int> filled(n_bins);
	if (count > 0) CMCB_CUDA_OK(cudaMemcpyAsync(out.data(), d_out, out.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
	CMCB_CUDA_OK(cudaMemcpyAsync(filled.data(), d_filled, n_bins * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
	CMCB_CUDA_OK(cudaStreamSynchronize(c->stream));
	for (long b = 0; b < n_bins; b++)
		if (filled[b]) for (long k = 0; k < count; k++) multi_mass_r[b * count + k] = out[b * count + k]; // rows of empty bins stay as they are
	return 0;
}

// Host build of cmc_cosmic_b200/csrc/orbit_core.h (the exact per-star code the CUDA kernel runs), so the CPU test
// suite can check it bit-for-bit against the oracle / the compiled reference without a GPU.
// TEST-ONLY: the product library never runs this on the host.  Build: g++ -O2 -ffp-contract=off -shared -fPIC.
#include <vector>
#include "../../cmc_cosmic_b200/csrc/orbit_core.h"
using namespace cmcb;

// the packed (r, phi) table the kernels read (built on the device by pack_tables_kernel in the product)
static std::vector<RPhi> pack(const double *star_r, const double *star_phi, long n)
{
	std::vector<RPhi> v(n);
	for (long i = 0; i < n; i++) { v[i].r = star_r[i]; v[i].phi = star_phi[i]; }
	return v;
}

struct ArrayDraws {
	const double *d; long pos, end;
	bool has(int n) const { return pos + n <= end; }
	double next() { return d[pos++]; }
};

// the 16-ary index over star_r (built on the device by orbit_build_index in the product)
struct HostIndex {
	std::vector<double> r, lvl;
	RIndex x;
	HostIndex(const double *star_r, kint M)
	{
		r.assign(star_r, star_r + M + 1);
		r.resize(M + 1 + 32, 0.0); // readable past M; the values there must not matter (the leaf step masks them)
		long total;
		ridx_plan(M, x.off, &x.H, &total);
		lvl.assign(total, __builtin_huge_val());
		for (int h = 1; h <= x.H; h++) {
			long n = (h < x.H ? x.off[h + 1] : total) - x.off[h];
			for (long i = 0; i < n; i++) lvl[x.off[h] + i] = ridx_sample(star_r, M, h, i);
		}
		x.r = r.data(); x.lvl = lvl.data(); x.M = M;
	}
};

extern "C" double emul_get_positions(const double *star_r, const double *star_m, const double *star_phi, long N_MAX, long N_STAR,
				     double cenma_m, long n_local, const long *g_idx, const unsigned char *is_new,
				     const double *E_in, const double *J_in, const double *vr_in, const double *vt_in, const double *rnew_in,
				     double Rtidal, const double *draws, long n_draws, const long *csr_off, int force_exact,
				     int *status, double *rnew, double *vrnew, double *vtnew, double *X, double *r_peri, double *r_apo,
				     long *kmin, long *kmax, long *counters /* [0]=draws used, [1]=exact Q evals, [2]=brent */)
{
	OrbitTables t;
	long n_tab = N_MAX + 2;
	for (long j = 0; j < n_local; j++) if (g_idx[j] + 1 > n_tab) n_tab = g_idx[j] + 1;
	std::vector<RPhi> rp = pack(star_r, star_phi, n_tab);
	t.rp = rp.data(); t.m = star_m; t.N_MAX = N_MAX; t.invN = 1.0 / N_STAR; t.madhoc = 1.0 / ((double) N_STAR); t.cenma_m = cenma_m;
	HostIndex hx(star_r, (kint) (N_MAX + 1));
	t.ridx = hx.x;
	if (force_exact) t.ridx.r = nullptr; // force_exact = 1: plain bisection instead of the 16-ary index
	double max_rad = 0.0;
	long pos = 0;
	counters[0] = counters[1] = counters[2] = 0;
	for (long j = 0; j < n_local; j++) {
		ArrayDraws dr;
		dr.d = draws;
		if (csr_off) { dr.pos = csr_off[j]; dr.end = csr_off[j + 1]; } else { dr.pos = pos; dr.end = n_draws; }
		OrbitOut o;
		orbit_step_star(t, g_idx[j], is_new ? is_new[j] != 0 : false, rnew_in ? rnew_in[j] : 0.0, E_in[j], J_in[j], vr_in[j], vt_in[j], Rtidal, dr, o);
		pos = dr.pos;
		status[j] = o.status; rnew[j] = o.rnew; vrnew[j] = o.vrnew; vtnew[j] = o.vtnew; X[j] = o.X; r_peri[j] = o.r_peri; r_apo[j] = o.r_apo;
		kmin[j] = o.kmin; kmax[j] = o.kmax;
		counters[0] += o.n_draws; counters[1] += o.n_exact; counters[2] += o.n_brent;
		if (o.status == ORBIT_OK && o.rnew > max_rad) max_rad = o.rnew;
	}
	return max_rad;
}

// unit access for targeted tests
extern "C" long emul_find_zero_Q(const double *star_r, const double *star_m, const double *star_phi, long N_MAX, long N_STAR, long g_j,
				 double E, double J, long kmin, long kmax, long *n_exact)
{
	std::vector<RPhi> rp = pack(star_r, star_phi, N_MAX + 2);
	OrbitTables t; t.rp = rp.data(); t.m = star_m; t.N_MAX = N_MAX; t.invN = 1.0 / N_STAR; t.madhoc = t.invN; t.cenma_m = 0;
	t.ridx.r = nullptr;
	StarConst s; s.g_j = g_j; s.r_j = star_r[g_j]; s.m_j = star_m[g_j]; s.E = E; s.J = J;
	star_const_derive(t, s);
	int ne = 0;
	long k = find_zero_Q(t, s, kmin, kmax, &ne);
	*n_exact = ne;
	return k;
}

// FindZero_r(1, N_MAX + 1, r): the reference's bisection and the 16-ary index, side by side
extern "C" void emul_find_zero_r_both(const double *star_r, long N_MAX, const double *radii, long n, long *k_bisect, long *k_indexed)
{
	std::vector<RPhi> rp = pack(star_r, star_r, N_MAX + 2);
	OrbitTables t; t.rp = rp.data(); t.m = nullptr; t.N_MAX = N_MAX; t.invN = 0; t.madhoc = 0; t.cenma_m = 0;
	HostIndex hx(star_r, (kint) (N_MAX + 1));
	for (long i = 0; i < n; i++) {
		k_bisect[i] = find_zero_r(t, 1, (kint) (N_MAX + 1), radii[i]);
		k_indexed[i] = find_zero_r_indexed(hx.x, radii[i]);
	}
}

// TEST-ONLY host build of cmc_cosmic_b200/csrc/dynamics_core.h: the per-pair code of dynamics.cu's kernel run in a plain loop with glibc's
// sin / cos / log / pow, where it must agree with the oracle's restatement bit for bit (tests/test_host.py).
#include "../../cmc_cosmic_b200/csrc/dynamics_core.h"

using namespace cmcb;

extern "C" long emul_dynamics_pairs(const double *star_r, const double *star_m, const double *star_phi, long N_MAX, long N_STAR, long n_local,
				    long first_global, double dt, const long *sw /* relaxation ss_collision binsingle binbin tidal_capture bh_capture p */,
				    const double *par /* gamma coll_factor tc_factor wd_tc units_mstar units_l units_t */, const double *draws, double *vr,
				    double *vt, double *E, double *J, double *Y, const long *binind, const int *se_k, const double *rad, const double *bin_a,
				    int *status, double *rperi, double *W, double *rcm, double *beta)
{
	DynConfig c;
	c.relaxation = sw[0]; c.ss_collision = sw[1]; c.binsingle = sw[2]; c.binbin = sw[3]; c.tidal_capture = sw[4]; c.bh_capture = sw[5]; c.p = sw[6];
	c.gamma = par[0]; c.coll_factor = par[1]; c.tc_factor = par[2]; c.wd_tc = par[3]; c.units_mstar = par[4]; c.units_l = par[5]; c.units_t = par[6];
	c.dt = dt; c.madhoc = 1.0 / ((double) N_STAR); c.N_STAR = (double) N_STAR; c.N_LIMIT = N_MAX;
	const long n_pairs = n_local / 2;
	for (long i = 0; i < n_pairs; i++) {
		DynStar s[2];
		for (int q = 0; q < 2; q++) {
			const long k = 2 * i + 1 + q, g = first_global + k - 1;
			s[q].vr = vr[k - 1]; s[q].vt = vt[k - 1]; s[q].binind = binind[k - 1];
			s[q].m = star_m[g]; s[q].r = star_r[g]; s[q].phi = star_phi[g];
			s[q].rad = rad[k - 1]; s[q].se_k = se_k[k - 1];
			s[q].bin_a = s[q].binind > 0 ? bin_a[s[q].binind] : 0.0;
		}
		long kmin, kmax;
		dyn_n_local_range(first_global + 2 * i, c.p, c.N_LIMIT, &kmin, &kmax);
		DynPair o;
		dyn_pair(c, s[0], s[1], star_r[kmin], star_r[kmax], draws + 3 * i, &o);
		Y[2 * i] = Y[2 * i + 1] = o.Y;
		status[i] = o.status; rperi[i] = o.rperi; W[i] = o.W; rcm[i] = o.rcm; beta[i] = o.beta;
		if (o.status == DYN_RELAXED)
			for (int q = 0; q < 2; q++) { vr[2 * i + q] = o.vr[q]; vt[2 * i + q] = o.vt[q]; E[2 * i + q] = o.E[q]; J[2 * i + q] = o.J[q]; }
	}
	return n_pairs;
}

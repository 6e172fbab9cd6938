// dynamics_apply's pair loop without the strong encounters (reference src/cmc/cmc_dynamics.c:206-500) for ONE pair of neighbouring stars:
// the pair's kinematics (calc_encounter_dyns, src/cmc/cmc_dynamics_helper.c:831-870), the local density (calc_n_local, :768-786), the
// cross section of the pair's type (cmc_dynamics.c:250-432), the encounter probability (:436), the decision (:444) and the two-body
// relaxation branch (:468-540).  Shared by the kernel (dynamics.cu) and its host build (tests/host_emul), which is bit-identical to the
// CPU restatement the tests use; on the device sin / cos / log / pow are CUDA's (<= 2 ulp), hence the tolerance of the GPU parity tests.
// Every expression keeps the reference's operation order (the library is built with -fmad=false).
#pragma once
#include <math.h>
#include "fp64_exact.h"

namespace cmcb {

constexpr double DYN_PI = 3.14159265358979323;          // include/cmc/cmc.h:62
constexpr double DYN_FB_PI = 3.141592653589793238462643; // include/fewbody-0.24/fewbody.h:35
constexpr double DYN_MSUN = 1.989e+33;                   // include/cmc/cmc.h:66
constexpr double DYN_XBS = 2.0, DYN_XBB = 2.0;           // include/cmc/cmc.h:83-85

enum { DYN_RELAXED = 0, DYN_ENC_SS = 1, DYN_ENC_BIN = 2, DYN_UNTOUCHED = 3 };

struct DynConfig { // the reference's switches (include/cmc/cmc_vars.h) and what the loop takes from the cluster state
	long relaxation, ss_collision, binsingle, binbin, tidal_capture, bh_capture, p /* AVEKERNEL */;
	double gamma, coll_factor, tc_factor, wd_tc, units_mstar, units_l, units_t;
	double dt, madhoc, N_STAR;
	long N_LIMIT; // clus.N_MAX
};

struct DynStar { // what the loop reads of one star
	double vr, vt, m /* star_m[g] */, r /* star_r[g] */, phi /* star_phi[g] */, rad, bin_a /* binary[binind].a, 0 for a single */;
	long binind;
	int se_k;
};

struct DynPair {
	int status;
	double Y, rperi, W, rcm, beta /* before the clamp */;
	double vr[2], vt[2], E[2], J[2]; // new values of k and kp (status == DYN_RELAXED)
};

CMCB_HD double dyn_sqr(double x) { return x * x; }
CMCB_HD double dyn_cub(double x) { return x * x * x; }
CMCB_HD double dyn_gravfocus(double rperi, double madhoc, double msum, double W) // PI rperi^2 (1 + 2 madhoc (m_k + m_kp) / (rperi W^2))
{
	return DYN_PI * dyn_sqr(rperi) * (1.0 + 2.0 * madhoc * msum / (rperi * dyn_sqr(W)));
}

// r_lo = star_r[kmin], r_hi = star_r[kmax] of calc_n_local (the caller resolves the indices with dyn_n_local_range)
CMCB_HD void dyn_n_local_range(long k, long p, long N_LIMIT, long *kmin, long *kmax)
{
	*kmin = k - p;
	*kmax = k + p;
	if (*kmin < 0) { *kmin = 0; *kmax = 2 * p + 1; }
	else if (*kmax > N_LIMIT) { *kmax = N_LIMIT; *kmin = N_LIMIT - 2 * p - 1; }
}

// u[0] = Y, u[1] = the encounter draw, u[2] = the draw of psi
CMCB_HD void dyn_pair(const DynConfig &c, const DynStar &a, const DynStar &b, double r_lo, double r_hi, const double u[3], DynPair *o)
{
	double v[4], vp[4], w[4];
	o->Y = u[0];
	const double phi = u[0] * 2.0 * DYN_PI;
	v[1] = a.vt; v[2] = 0.0; v[3] = a.vr;
	vp[1] = b.vt * cos(phi); vp[2] = b.vt * sin(phi); vp[3] = b.vr;
	for (int j = 1; j <= 3; j++) w[j] = vp[j] - v[j];
	const double W = sqrt(dyn_sqr(w[1]) + dyn_sqr(w[2]) + dyn_sqr(w[3]));
	const double rcm = (a.m * a.r + b.m * b.r) / (a.m + b.m);
	const double n_local = (2.0 * ((double) c.p)) * 3.0 / (4.0 * DYN_PI * (dyn_cub(r_hi) - dyn_cub(r_lo)));
	const double mass_k = a.m, mass_kp = b.m, madhoc = c.madhoc;
	double rperi = 0.0, S;
	if (a.binind > 0 && b.binind > 0) {
		rperi = DYN_XBB * (a.bin_a + b.bin_a);
		S = c.binbin ? dyn_gravfocus(rperi, madhoc, mass_k + mass_kp, W) : 0.0;
	} else if (a.binind > 0 || b.binind > 0) {
		rperi = DYN_XBS * (a.binind > 0 ? a.bin_a : b.bin_a);
		S = c.binsingle ? dyn_gravfocus(rperi, madhoc, mass_k + mass_kp, W) : 0.0;
	} else if (c.ss_collision) {
		const int kk = a.se_k, kkp = b.se_k;
		double S_tmp = 0.0, collisions_multiple;
		// cmc_dynamics.c:305-313 overwrites what :285-303 (CO_TDE) chose: WD_TC for two white dwarfs when WD_TC > 0, else COLL_FACTOR
		if (c.wd_tc > 0 && (kkp >= 10 && kkp <= 12) && (kk >= 10 && kk <= 12)) collisions_multiple = c.wd_tc;
		else collisions_multiple = c.coll_factor;
		if (c.tidal_capture) {
			double S_lombardi;
			if ((kk <= 1 || kk >= 10 || kk == 7) && (kkp >= 2 && kkp <= 9 && kkp != 7)) {
				rperi = 1.3 * b.rad;
				S_lombardi = dyn_gravfocus(rperi, madhoc, mass_k + mass_kp, W);
			} else if ((kkp <= 1 || kkp >= 10 || kkp == 7) && (kk >= 2 && kk <= 9 && kk != 7)) {
				rperi = 1.3 * a.rad;
				S_lombardi = dyn_gravfocus(rperi, madhoc, mass_k + mass_kp, W);
			} else S_lombardi = 0.0;
			S_tmp = S_tmp > S_lombardi ? S_tmp : S_lombardi;
		}
		if (c.bh_capture && kk == 14 && kkp == 14) {
			const double clight10o7 = pow(2.9979e10 / (c.units_l / c.units_t), 1.428571);
			rperi = 2.6789966 * madhoc * pow(mass_k * mass_kp, 0.28571428) * pow(mass_k + mass_kp, 0.42857143) / pow(W, 0.57142857) / clight10o7;
			const double S_brem = dyn_gravfocus(rperi, madhoc, mass_k + mass_kp, W);
			S_tmp = S_tmp > S_brem ? S_tmp : S_brem;
		}
		if (c.tc_factor > 1 && (kk <= 1 || kk == 7 || kk >= 10) && (kkp <= 1 || kkp == 7 || kkp >= 10)) {
			rperi = c.tc_factor * (a.rad + b.rad);
			const double S_tc_simple = dyn_gravfocus(rperi, madhoc, mass_k + mass_kp, W);
			S_tmp = S_tmp > S_tc_simple ? S_tmp : S_tc_simple;
		}
		rperi = collisions_multiple * (a.rad + b.rad);
		const double S_coll = dyn_gravfocus(rperi, madhoc, mass_k + mass_kp, W);
		S = S_coll > S_tmp ? S_coll : S_tmp;
		rperi = madhoc * (mass_k + mass_kp) / dyn_sqr(W) * (-1.0 + sqrt(1.0 + S / DYN_FB_PI * dyn_sqr(W * W / (madhoc * mass_k + madhoc * mass_kp))));
	} else S = 0.0;
	const double P_enc = n_local * W * S * (c.dt * c.N_STAR / log(c.gamma * c.N_STAR));
	o->rperi = rperi; o->W = W; o->rcm = rcm; o->beta = 0.0;
	if (u[1] < P_enc) { o->status = (a.binind > 0 || b.binind > 0) ? DYN_ENC_BIN : DYN_ENC_SS; return; }
	if (!c.relaxation) { o->status = DYN_UNTOUCHED; return; }
	double w1[4], w2[4], w_new[4], v_new[4], vp_new[4];
	const double Trel12 = (DYN_PI / 32.0) * dyn_cub(W) / (c.N_STAR * n_local * dyn_sqr((mass_k + mass_kp) * madhoc));
	double beta = (DYN_PI / 2.0) * sqrt(c.dt / Trel12);
	o->beta = beta;
	if (beta > DYN_PI / 2.0) beta = DYN_PI / 2.0;
	const double wp = sqrt(dyn_sqr(w[1]) + dyn_sqr(w[2]));
	w1[1] = -w[2] * W / wp; w1[2] = w[1] * W / wp; w1[3] = 0.0;
	w2[1] = -w[1] * w[3] / wp; w2[2] = -w[2] * w[3] / wp; w2[3] = wp;
	const double psi = u[2] * 2 * DYN_PI;
	const double cb = cos(beta), sb = sin(beta), cp = cos(psi), sp = sin(psi);
	for (int j = 1; j <= 3; j++) w_new[j] = w[j] * cb + w1[j] * sb * cp + w2[j] * sb * sp;
	for (int j = 1; j <= 3; j++) {
		v_new[j] = v[j] - mass_kp / (mass_k + mass_kp) * (w_new[j] - w[j]);
		vp_new[j] = vp[j] + mass_k / (mass_k + mass_kp) * (w_new[j] - w[j]);
	}
	o->vr[0] = v_new[3];
	o->vt[0] = sqrt(dyn_sqr(v_new[1]) + dyn_sqr(v_new[2]));
	o->vr[1] = vp_new[3];
	o->vt[1] = sqrt(dyn_sqr(vp_new[1]) + dyn_sqr(vp_new[2]));
	o->E[0] = a.phi + 0.5 * (dyn_sqr(o->vr[0]) + dyn_sqr(o->vt[0])); // set_star_EJ, cmc_dynamics_helper.c:878-882
	o->J[0] = a.r * o->vt[0];
	o->E[1] = b.phi + 0.5 * (dyn_sqr(o->vr[1]) + dyn_sqr(o->vt[1]));
	o->J[1] = b.r * o->vt[1];
	o->status = DYN_RELAXED;
}

} // namespace cmcb

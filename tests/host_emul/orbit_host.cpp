// Host build of cmc_cosmic_b200/csrc/orbit_core.h (the exact per-star code the CUDA kernel runs), so the CPU test
// suite can check it bit-for-bit against the oracle / the compiled reference without a GPU.
// TEST-ONLY: the product library never runs this on the host.  Build: g++ -O2 -ffp-contract=off -shared -fPIC.
#include <algorithm>
#include <stdlib.h>
#include <vector>
#include "../../cmc_cosmic_b200/csrc/orbit_core.h"
using namespace cmcb;

// the packed (r, phi) table the kernels read (built on the device by pack_tables_kernel in the product)
static std::vector<RPhi> pack(const double *star_r, const double *star_phi, long n)
{
	std::vector<RPhi> v(n + 16); // the last step of the r ladder reads a whole 8-entry line (masked beyond the table)
	for (long i = 0; i < n; i++) { v[i].r = star_r[i]; v[i].phi = star_phi[i]; }
	for (long i = n; i < n + 16; i++) { v[i].r = 0.0; v[i].phi = 0.0; }
	return v;
}

struct ArrayDraws {
	const double *d; long pos, end;
	bool has(int n) const { return pos + n <= end; }
	double next() { return d[pos++]; }
};

// the ladders (built on the device by orbit_build_ladders in the product): every level straight from the fine table,
// independently of the device builder's level-by-level merge
struct HostLadders {
	std::vector<std::vector<LBlock>> qb;
	std::vector<std::vector<double>> rl;
	QLadder q;
	RLadder r;
	bool unsorted = false;
	HostLadders(const std::vector<RPhi> &rp, kint M)
	{
		long cnt[LADDER_MAX_LEVELS], alloc[LADDER_MAX_LEVELS];
		q.n_levels = qlad_plan(M, q.shift, cnt);
		q.factor_log2 = QLAD_FACTOR_LOG2; q.max_width = 0x7fffffff;
		qb.resize(q.n_levels);
		for (int l = 0; l < q.n_levels; l++) {
			qb[l].resize(cnt[l]);
			for (long b = 0; b < cnt[l]; b++) {
				long k0 = b << q.shift[l], k1 = (b + 1) << q.shift[l];
				if (k1 > M) k1 = M;
				qb[l][b] = lblock_of(rp.data(), k0, k1, &unsorted);
			}
			q.blk[l] = qb[l].data();
		}
		r.n_levels = rlad_plan(M, r.shift, cnt, alloc);
		rl.resize(r.n_levels);
		for (int l = 0; l < r.n_levels; l++) {
			rl[l].assign(alloc[l], __builtin_huge_val());
			for (long b = 0; b < cnt[l]; b++) rl[l][b] = rp[b << r.shift[l]].r;
			r.lvl[l] = rl[l].data();
		}
		r.n_top = r.n_levels ? (int) cnt[0] : 0;
		r.whole_lines = 0;
		if (unsorted) { q.n_levels = 0; r.n_levels = 0; }
	}
};

extern "C" double emul_get_positions(const double *star_r, const double *star_m, const double *star_phi, long N_MAX, long N_STAR,
				     double cenma_m, long n_local, const long *g_idx, const unsigned char *is_new,
				     const double *E_in, const double *J_in, const double *vr_in, const double *vt_in, const double *rnew_in,
				     double Rtidal, const double *draws, long n_draws, const long *csr_off, int force_exact,
				     int *status, double *rnew, double *vrnew, double *vtnew, double *X, double *r_peri, double *r_apo,
				     long *kmin, long *kmax, long *counters /* [0]=draws used, [1]=exact Q evals, [2]=brent, [3]/[4]=ladder levels in use (Q / r) */)
{
	OrbitTables t;
	long n_tab = N_MAX + 2;
	for (long j = 0; j < n_local; j++) if (g_idx[j] + 1 > n_tab) n_tab = g_idx[j] + 1;
	std::vector<RPhi> rp = pack(star_r, star_phi, n_tab);
	t.rp = rp.data(); t.m = star_m; t.N_MAX = N_MAX; t.invN = 1.0 / N_STAR; t.madhoc = 1.0 / ((double) N_STAR); t.cenma_m = cenma_m;
	std::vector<RPhiF> rpf(rp.size()); // the FP32 screening copy, as orbit_pack_tables rounds it
	for (size_t i = 0; i < rp.size(); i++) { rpf[i].r = (float) rp[i].r; rpf[i].phi = (float) rp[i].phi; }
	t.rpf = rpf.data();
	HostLadders lad(rp, (kint) (N_MAX + 1));
	t.qlad = lad.q; t.rlad = lad.r;
	if (force_exact == 1) { t.qlad.n_levels = 0; t.rlad.n_levels = 0; t.rpf = nullptr; } // 1: every probe reads the fine FP64 table
	if (force_exact == 2) t.qlad.n_levels = 0; // 2: r ladder + FP32 screening of the index bisections, no Q ladder
	// 3: the device's default configuration -- r ladder + sample ladder (find_zero_Q_sampled), no Q ladder, no screening
	std::vector<std::vector<RPhi>> slv;
	std::vector<RPhiF> stopf; // the coarsest sample level rounded to FP32, as slad_build writes it
	t.slad.n_levels = 0;
	if (force_exact == 3 && lad.r.n_levels > 0) {
		t.qlad.n_levels = 0; t.rpf = nullptr;
		long scnt[LADDER_MAX_LEVELS];
		const int ns = slad_plan(N_MAX + 1, t.slad.shift, scnt);
		slv.resize(ns);
		for (int l = 0; l < ns; l++) {
			slv[l].resize(scnt[l]);
			for (long b = 0; b < scnt[l]; b++) slv[l][b] = rp[b << t.slad.shift[l]];
			t.slad.lvl[l] = slv[l].data();
		}
		stopf.resize(ns ? scnt[0] : 0);
		for (size_t b = 0; b < stopf.size(); b++) { stopf[b].r = (float) slv[0][b].r; stopf[b].phi = (float) slv[0][b].phi; }
		t.slad.topf = (ns && !getenv("CMCB_EMUL_NO_F32_TOP")) ? stopf.data() : nullptr;
		t.slad.n_levels = ns;
		t.slad.margin = (double) (N_MAX + 1) * 7.105427357601002e-15 + 9.094947017729282e-13;
		t.slad.phi_abs_max = __builtin_fmax(__builtin_fabs(rp[0].phi), __builtin_fabs(rp[1].phi));
	}
	counters[3] = t.qlad.n_levels; counters[4] = t.rlad.n_levels;
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

// unit access for targeted tests: find_zero_Q with (use_ladder = 1) and without the ladder
extern "C" long emul_find_zero_Q(const double *star_r, const double *star_m, const double *star_phi, long N_MAX, long N_STAR, long g_j,
				 double E, double J, long kmin, long kmax, long *n_exact, int use_ladder)
{
	std::vector<RPhi> rp = pack(star_r, star_phi, N_MAX + 2);
	OrbitTables t; t.rpf = nullptr; t.slad.n_levels = 0; t.rp = rp.data(); t.m = star_m; t.N_MAX = N_MAX; t.invN = 1.0 / N_STAR; t.madhoc = t.invN; t.cenma_m = 0;
	HostLadders lad(rp, (kint) (N_MAX + 1));
	t.qlad = lad.q; t.rlad = lad.r;
	if (!use_ladder) { t.qlad.n_levels = 0; t.rlad.n_levels = 0; }
	StarConst s; s.g_j = g_j; s.r_j = star_r[g_j]; s.m_j = star_m[g_j]; s.E = E; s.J = J;
	star_const_derive(t, s);
	int ne = 0;
	LadderTops top; top.q = t.qlad.n_levels ? t.qlad.blk[0] : nullptr; top.r = t.rlad.n_levels ? t.rlad.lvl[0] : nullptr; top.qdesc = &t.qlad; top.s = nullptr; top.sdesc = nullptr;
	long k = find_zero_Q(t, top, s, kmin, kmax, &ne);
	*n_exact = ne;
	return k;
}

// FindZero_r(1, N_MAX + 1, r): the reference's bisection and the r ladder, side by side
extern "C" void emul_find_zero_r_both(const double *star_r, long N_MAX, const double *radii, long n, long *k_bisect, long *k_ladder)
{
	std::vector<RPhi> rp = pack(star_r, star_r, N_MAX + 2);
	OrbitTables t; t.rpf = nullptr; t.slad.n_levels = 0; t.rp = rp.data(); t.m = nullptr; t.N_MAX = N_MAX; t.invN = 0; t.madhoc = 0; t.cenma_m = 0;
	HostLadders lad(rp, (kint) (N_MAX + 1));
	t.qlad = lad.q; t.rlad = lad.r;
	for (long i = 0; i < n; i++) {
		k_bisect[i] = find_zero_r(t, 1, (kint) (N_MAX + 1), radii[i]);
		k_ladder[i] = t.rlad.n_levels ? find_zero_r_ladder(t, t.rlad.lvl[0], radii[i]) : -1;
	}
}

// the ladder plans (levels, shifts, sizes), for the tests of the layout arithmetic
extern "C" int emul_ladder_plan(long M, int *qshift, long *qcount, int *rshift, long *rcount, long *ralloc, int *n_r)
{
	*n_r = rlad_plan(M, rshift, rcount, ralloc);
	return qlad_plan(M, qshift, qcount);
}

// set_velocities3's first loop (reference cmc_utils.c:268-311) the way velocities.cu runs it: prepare per star, per-block carries from 0,
// the serial pass over the block results, the real walk of the blocks that need one -- here one after the other, with the block length
// as a parameter so that small clusters have many blocks.  vr/vt/vrnew/vtnew in/out.  Returns the leftover Eexcess; counters[0] = blocks
// replayed by the carry pass, [1] = blocks walked by apply, [2] = failed potential() brackets.
#include "../../cmc_cosmic_b200/csrc/velocities_core.h"
extern "C" double emul_set_velocities3(const double *star_r, const double *star_m, const double *star_phi, long N_MAX, long N_STAR, long n_local,
				       long first_global, long blk, const long *interacted, const double *rOld, const double *vrold, const double *vtold,
				       const double *Uoldrold, const double *Uoldrnew, double *vr, double *vt, double *vrnew, double *vtnew,
				       unsigned char *kind, long *counters)
{
	std::vector<RPhi> rp = pack(star_r, star_phi, N_MAX + 2);
	OrbitTables t; t.rpf = nullptr; t.slad.n_levels = 0; t.rp = rp.data(); t.m = star_m; t.N_MAX = N_MAX; t.invN = 1.0 / N_STAR; t.madhoc = 1.0 / ((double) N_STAR); t.cenma_m = 0;
	HostLadders lad(rp, (kint) (N_MAX + 1));
	t.qlad = lad.q; t.rlad = lad.r;
	LadderTops top; top.q = nullptr; top.qdesc = nullptr; top.s = nullptr; top.sdesc = nullptr; top.r = t.rlad.n_levels ? t.rlad.lvl[0] : nullptr;
	const Sv3Index gi{first_global, n_local, N_MAX};
	std::vector<double> val(n_local);
	counters[0] = counters[1] = counters[2] = 0;
	for (long j = 0; j < n_local; j++) { // prepare
		Sv3In s{interacted[j], rOld[j], vrold[j], vtold[j], Uoldrold[j], Uoldrnew[j], vr[j], vt[j]};
		int fail = 0;
		double a = vr[j], b = vt[j];
		kind[j] = (unsigned char) sv3_prepare_star(t, top, gi(j), s, &a, &b, &val[j], &fail);
		counters[2] += fail;
		if ((kind[j] & SV3_ACTIVE) && !(kind[j] & SV3_SOURCE)) { vr[j] = vrnew[j] = a; vt[j] = vtnew[j] = b; }
	}
	const long nb = (n_local + blk - 1) / blk;
	std::vector<double> eb(nb), ein(nb);
	std::vector<char> has(nb);
	auto none = [](long, double) {};
	for (long b = 0; b < nb; b++) { // blocks
		const long lo = b * blk, hi = std::min(lo + blk, n_local);
		has[b] = 0;
		for (long i = lo; i < hi; i++) if (kind[i] & SV3_SOURCE) has[b] = 1;
		eb[b] = has[b] ? sv3_walk(kind, val.data(), star_m, gi, lo, hi, 0.0, none) : 0.0;
	}
	double E = 0.0;
	for (long b = 0; b < nb; b++) { // carry
		ein[b] = E;
		if (E == 0.0) E = eb[b];
		else { E = sv3_walk(kind, val.data(), star_m, gi, b * blk, std::min((b + 1) * blk, n_local), E, none); counters[0]++; }
	}
	for (long b = 0; b < nb; b++) { // apply
		if (ein[b] == 0.0 && !has[b]) continue;
		counters[1]++;
		sv3_walk(kind, val.data(), star_m, gi, b * blk, std::min((b + 1) * blk, n_local), ein[b], [&](long i, double ratio) { vr[i] = dmul(vr[i], ratio); vt[i] = dmul(vt[i], ratio); });
	}
	return E;
}

// one pass of the second loop (:323-343) from the received Eexcess
extern "C" double emul_set_velocities3_second(const double *star_m, long N_MAX, long n_local, long first_global, const unsigned char *kind, double E,
					      double *vr, double *vt, const double *vrnew, const double *vtnew)
{
	const Sv3Index gi{first_global, n_local, N_MAX};
	for (long i = 0; i < n_local; i++) {
		double ratio;
		if (sv3_second_step(kind[i], vr[i], vt[i], vrnew[i], vtnew[i], star_m[gi(i)], E, &ratio)) { vr[i] = dmul(vr[i], ratio); vt[i] = dmul(vt[i], ratio); }
	}
	return E;
}

// calc_sigma_r through the operation list of sigma_core.h.  mode 0: the list applied in order (checks the mapping against the oracle);
// mode 1: the list through the exact-scan decomposition (approximate prefix in a different order, seq_classify_signed with the largest
// freshly summed window as the bound, events applied in order) -- what the device path will run.  vr / vt: local stars 1..n at [p .. p+n-1],
// the ghost stars of the previous rank at [0 .. p-1], of the next rank at [p+n .. 2p+n-1].  Outputs [n + 1], entry 0 unused; counters[0] = events.
#include "../../cmc_cosmic_b200/csrc/sigma_core.h"
#include "../../cmc_cosmic_b200/csrc/exactsum.h"
extern "C" int emul_calc_sigma_r(long p, long n, long first_global, long N_MAX, const double *star_r, const double *star_m, double madhoc,
				 const double *vr, const double *vt, int r_0_mave_1, int mode, double *sig_a, double *sig_sigma, long *counters)
{
	SigmaWindow w{p, n, first_global, N_MAX, first_global == 1};
	if (!w.valid()) return -1;
	const long L = w.n_ops();
	std::vector<double> t2(n + 2 * p), t1(n + 2 * p); // index k + p - 1 for local star k in [1 - p, n + p]
	for (long k = 1 - p; k <= n + p; k++) {
		const long g = first_global + k - 1;
		const double m = (g >= 0 && g <= N_MAX + 1) ? star_m[g] : 0.0;
		sigma_terms(m, madhoc, vr[k + p - 1], vt[k + p - 1], &t2[k + p - 1], &t1[k + p - 1]);
	}
	counters[0] = 0;
	for (int which = 0; which < 2; which++) {
		const std::vector<double> &t = which == 0 ? t2 : t1;
		std::vector<double> a(L), y(L);
		for (long j = 0; j < L; j++) { long k; int s; w.op(j, &k, &s); a[j] = s == 0 ? 0.0 : (s > 0 ? t[k + p - 1] : -t[k + p - 1]); }
		if (mode == 0) { double acc = 0.0; for (long j = 0; j < L; j++) { acc = dadd(acc, a[j]); y[j] = acc; } }
		else {
			double ymax = 0.0; // the largest freshly summed window, with a margin
			for (long si = 1; si <= n; si++) { double s = 0.0; const long lo = w.simin(si); for (long k = lo; k < lo + 2 * p; k++) s += t[k + p - 1]; if (s > ymax) ymax = s; }
			ymax *= 1.01;
			std::vector<double> S(L); // approximate prefix: blocks of 37, block sums accumulated separately
			{ double base = 0.0; for (long b = 0; b < L; b += 37) { double s = 0.0; for (long j = b; j < L && j < b + 37; j++) { s += a[j]; S[j] = base + s; } base += s; } }
			double base = 0.0, G = 0.0;
			for (long j = 0; j < L; j++) {
				SeqTerm c = seq_classify_signed(a[j], j ? S[j - 1] : 0.0, S[j], ymax, j + 1);
				if (c.event) { base = dadd(dadd(base, G), a[j]); G = 0.0; y[j] = base; counters[0]++; }
				else { G += c.v; y[j] = base + G; }
			}
		}
		for (long j = 0; j < L; j++) {
			const long si = w.star_of(j);
			if (!si) continue;
			if (which == 0) sig_sigma[si] = y[j];                   // Mv2ave, finished below
			else {
				const double Mave = y[j];
				sig_sigma[si] = dsqrt(ddiv(sig_sigma[si], Mave));  // :2977
				sig_a[si] = r_0_mave_1 == 0 ? star_r[first_global + si - 1] : ddiv(ddiv(Mave, 2.), (double) p); // :2971-2974
			}
		}
	}
	return 0;
}

#ifdef CMCB_SCREEN_STATS
extern "C" void emul_screen_stats(long *out, int reset) { out[0] = g_screen_stats[0]; out[1] = g_screen_stats[1]; if (reset) g_screen_stats[0] = g_screen_stats[1] = 0; }
#endif

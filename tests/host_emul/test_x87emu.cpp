// x87emu.h against the host's real long double (x86-64: 80-bit extended).  Build: g++ -O2 -ffp-contract=off.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cfloat>
#include <random>
#include "../../cmc_cosmic_b200/csrc/x87emu.h"
#include "../../cmc_cosmic_b200/csrc/orbit_core.h"
using namespace cmcb;

static_assert(LDBL_MANT_DIG == 64, "needs x87 long double");

static X87 from_ld(long double v)
{
	X87 x; x.s = std::signbit(v); if (v == 0) { x.m = 0; x.e = 0; return x; }
	int ex; long double f = frexpl(fabsl(v), &ex); // f in [0.5,1)
	x.m = (uint64_t) ldexpl(f, 64); x.e = ex - 1; return x;
}
static long double to_ld(X87 x) { long double v = ldexpl((long double) x.m, x.e - 63); return x.s ? -v : v; }
static bool same(X87 x, long double v) { long double w = to_ld(x); return memcmp(&w, &v, 10) == 0 || (w == 0 && v == 0); }

static inline double SQRd(double a) { return a * a; }
// the reference expression, restated with the same types (cmc_utils.c:1654-1662, cmc.h:2175)
static double ref_function_q(long double r, long double pot, long double E, long double J, double m_j, double r_j, long N_STAR)
{
	double res; long double Jr, phis;
	Jr = SQRd((J) / (r));
	phis = ((r) >= r_j ? m_j / (r) : m_j / r_j) * (1.0 / N_STAR);
	res = (2.0 * ((E) - (pot + phis)) - Jr);
	return res;
}

int main(int argc, char **argv)
{
	long iters = argc > 1 ? atol(argv[1]) : 2000000;
	std::mt19937_64 rng(987);
	std::uniform_real_distribution<double> U(0.0, 1.0);
	long bad = 0;
	auto rnd = [&](int spread) { double v = ldexp(1.0 + U(rng), (int) (rng() % (2 * spread + 1)) - spread); if ((rng() & 3) == 0) { int ex; double f = frexp(v, &ex); v = ldexp(floor(ldexp(f, 20)), ex - 20); } return (rng() & 1) ? v : -v; };
	for (long it = 0; it < iters; it++) {
		double a = rnd(40), b = rnd(40);
		if ((rng() & 7) == 0) b = a * (1 + ((int) (rng() % 9) - 4) * ldexp(1.0, -(int) (rng() % 64)));
		if (b == 0) b = a;
		long double la = a, lb = b;
		X87 xa = x87_from_double(a), xb = x87_from_double(b);
		// chained extended operands too: c = a*b (64-bit), then mix
		X87 xc = x87_mul(xa, xb); long double lc = la * lb;
		X87 xd = x87_div(xa, xb); long double ld = la / lb;
		bool ok = same(xa, la) && same(xc, lc) && same(xd, ld) && same(x87_add(xa, xb), la + lb) && same(x87_sub(xa, xb), la - lb)
			&& same(x87_add(xc, xd), lc + ld) && same(x87_sub(xc, xd), lc - ld) && same(x87_sub(xd, xc), ld - lc)
			&& same(x87_mul(xc, xd), lc * ld) && same(x87_div(xc, xd), lc / ld) && same(x87_add(xc, xa), lc + la) && same(x87_sub(xc, xa), lc - la);
		double dc = (double) lc, dd = (double) ld, ds = (double) (lc - ld);
		double ec = x87_to_double(xc), ed = x87_to_double(xd), es = x87_to_double(x87_sub(xc, xd));
		ok = ok && memcmp(&dc, &ec, 8) == 0 && memcmp(&dd, &ed, 8) == 0 && (memcmp(&ds, &es, 8) == 0 || (ds == 0 && es == 0));
		if (!ok && bad++ < 5) printf("op mismatch a=%a b=%a\n", a, b);
	}
	printf("basic ops: %ld/%ld mismatches\n", bad, iters);
	// function_q on cluster-like numbers, including forced near-cancellation so the rounding pattern matters
	long badq = 0, sign_only_diff = 0;
	for (long it = 0; it < iters; it++) {
		long N = 1000 + rng() % 10000000;
		double r = exp(-7 + 14 * U(rng)), r_j = exp(-7 + 14 * U(rng)), m_j = 0.1 + 5 * U(rng);
		double pot = -(0.05 + 2 * U(rng)), J = U(rng) * r * sqrt(2 * fabs(pot));
		double E = pot * U(rng);
		if (it & 1) { // put E right at the root: Q ~ few ulps
			long double Jr = SQRd((double) ((long double) J / r));
			long double ph = ((long double) r >= r_j ? m_j / (long double) r : (long double) (m_j / r_j)) * (1.0 / N);
			E = (double) ((long double) pot + ph + Jr / 2) * (1 + ((int) (rng() % 7) - 3) * 0x1p-52);
		}
		if ((it & 7) == 3) r = 1e-20;
		double q0 = ref_function_q(r, pot, E, J, m_j, r_j, N);
		double q1 = x87_function_q(r, pot, E, J, m_j, r_j, 1.0 / N);
		if (memcmp(&q0, &q1, 8) != 0 && !(q0 == 0 && q1 == 0)) { if (badq++ < 5) printf("Q mismatch: %a vs %a\n", q0, q1); }
		double qd = 2.0 * (E - (pot + (r >= r_j ? m_j / r : m_j / r_j) * (1.0 / N))) - SQRd(J / r);
		if ((qd < 0) != (q0 < 0) || (qd > 0) != (q0 > 0)) sign_only_diff++;
	}
	printf("function_q: %ld/%ld mismatches (plain-FP64 evaluation had the wrong sign in %ld of them)\n", badq, iters, sign_only_diff);
	// the division-free sign test of the bisection probes (orbit_core.h sign_Q_at): whenever its FP64 fast path decides, the sign
	// must be the sign of the reference's function_q.  E is placed at the root and moved away in steps of 2^s ulps so that the
	// decision straddles the error bound; radial (J = 0) and r < r_j / r >= r_j cases included.
	long bads = 0, fast = 0, slow = 0;
	for (long it = 0; it < iters; it++) {
		long N = 1000 + rng() % 10000000;
		double r = exp(-9 + 18 * U(rng)), m_j = 0.1 + 5 * U(rng);
		double r_j = (it % 3 == 0) ? r : exp(-9 + 18 * U(rng));
		double pot = -(0.05 + 2 * U(rng)) * ((it & 15) == 7 ? 1e-6 : 1.0);
		double J = ((it & 7) == 5) ? 0.0 : U(rng) * r * sqrt(2 * fabs(pot));
		if ((it & 31) == 9) r = 1e-20;
		long double Jr = SQRd((double) ((long double) J / r));
		long double ph = ((long double) r >= r_j ? m_j / (long double) r : (long double) (m_j / r_j)) * (1.0 / N);
		double E0 = (double) ((long double) pot + ph + Jr / 2);
		int sh = (int) (rng() % 12);
		long k = (long) (rng() % 9) - 4;
		double E = E0 * (1 + (double) (k << sh) * 0x1p-52);
		if (!(E < 0) || !std::isfinite(E)) continue;
		RPhi tab[1];
		tab[0].r = r; tab[0].phi = pot;
		OrbitTables t; t.rpf = nullptr; t.slad.n_levels = 0; t.rp = tab; t.m = nullptr; t.N_MAX = 0; t.invN = 1.0 / N; t.madhoc = t.invN; t.cenma_m = 0;
		StarConst sc; sc.g_j = 0; sc.r_j = r_j; sc.m_j = m_j; sc.E = E; sc.J = J;
		star_const_derive(t, sc);
		int ne = 0;
		int sg = sign_Q_at(t, sc, tab[0], &ne);
		double q0 = ref_function_q(r, pot, E, J, m_j, r_j, N);
		int want = q0 > 0 ? 1 : (q0 < 0 ? -1 : 0);
		if (ne) slow++; else fast++;
		if (sg != want) { if (bads++ < 5) printf("sign mismatch: got %d want %d (q=%a, exact path %d)\n", sg, want, q0, ne); }
	}
	printf("sign_Q_at: %ld mismatches; fast path decided %ld, exact replay %ld\n", bads, fast, slow);
	printf("%s\n", (bad || badq || bads) ? "FAIL" : "PASS");
	return (bad || badq || bads) != 0;
}

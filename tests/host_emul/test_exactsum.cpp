// Host-side emulation of the tile/segment/event decomposition used by cmc_cosmic_b200/csrc/potential.cu,
// checked bit-for-bit against the plain sequential loop.  Build: g++ -O2 -ffp-contract=off.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <random>
#include "../../cmc_cosmic_b200/csrc/exactsum.h"
using namespace cmcb;

static long g_events_total = 0;

// approximate inclusive prefix: blocked + pairwise, deliberately NOT the sequential order
static void approx_prefix(const std::vector<double> &a, std::vector<double> &S)
{
	size_t n = a.size();
	S.resize(n);
	const size_t B = 37;
	std::vector<double> blk;
	for (size_t b = 0; b < n; b += B) { double s = 0; for (size_t i = b; i < n && i < b + B; i++) s += a[i]; blk.push_back(s); }
	std::vector<double> bp(blk.size() + 1, 0.0);
	for (size_t b = blk.size(); b-- > 0;) bp[b] = 0; // placeholder
	double acc = 0; // Kahan over block sums to get a *different* rounding than sequential
	double c = 0;
	for (size_t b = 0; b < blk.size(); b++) { bp[b] = acc; double y = blk[b] - c, t = acc + y; c = (t - acc) - y; acc = t; }
	for (size_t b = 0, k = 0; b < n; b += B, k++) { double s = 0; for (size_t i = b; i < n && i < b + B; i++) { s += a[i]; S[i] = bp[k] + s; } }
}

// the emulated algorithm: classification -> exact run sums -> sequential event resolution -> outputs
static void srs_emulated(const std::vector<double> &a, double y0, int sgn, std::vector<double> &y)
{
	size_t n = a.size();
	std::vector<double> S;
	approx_prefix(a, S);
	y.resize(n);
	double base = y0;      // exact value right after the last event (or y0)
	double G = 0.0;        // exact run sum since then
	for (size_t i = 0; i < n; i++) {
		double Sb = i ? S[i - 1] : 0.0;
		double yb = sgn > 0 ? y0 + Sb : y0 - Sb, ya = sgn > 0 ? y0 + S[i] : y0 - S[i];
		SeqTerm t = seq_classify(a[i], yb, ya, fabs(y0), (long) i + 1);
		if (t.event) {
			double before = sgn > 0 ? base + G : base - G; // exact
			base = sgn > 0 ? before + a[i] : before - a[i]; // the hardware rounding
			G = 0.0;
			y[i] = base;
			g_events_total++;
		} else {
			G += t.v; // exact in any order
			y[i] = sgn > 0 ? base + G : base - G;
		}
	}
}

// signed chain (terms carry their sign, y0 = 0), with the caller's bound on the largest chain value
static int check_signed(const char *name, const std::vector<double> &a, double y_max)
{
	size_t n = a.size();
	std::vector<double> S, y(n);
	approx_prefix(a, S);
	double base = 0.0, G = 0.0;
	long events = 0;
	for (size_t i = 0; i < n; i++) {
		SeqTerm t = seq_classify_signed(a[i], i ? S[i - 1] : 0.0, S[i], y_max, (long) i + 1);
		if (t.event) { base = (base + G) + a[i]; G = 0.0; y[i] = base; events++; }
		else { G += t.v; y[i] = base + G; }
	}
	double s = 0.0;
	size_t bad = 0, first = 0;
	for (size_t i = 0; i < n; i++) { s = s + a[i]; if (memcmp(&s, &y[i], 8) != 0) { if (!bad) first = i; bad++; } }
	printf("%-34s n=%-9zu events=%-6ld %s", name, n, events, bad ? "MISMATCH" : "bit-exact\n");
	if (bad) printf(" (%zu, first at %zu)\n", bad, first);
	return bad != 0;
}

// the op sequence of a window of 2p terms sliding over t: 2p adds, then (subtract the oldest, add the next) per step
static std::vector<double> window_ops(const std::vector<double> &t, size_t p, double *y_max)
{
	std::vector<double> a;
	double w = 0, wmax = 0;
	for (size_t k = 0; k < 2 * p && k < t.size(); k++) { a.push_back(t[k]); w += t[k]; }
	wmax = w;
	for (size_t k = 2 * p; k < t.size(); k++) {
		a.push_back(-t[k - 2 * p]); a.push_back(t[k]);
		double d = 0; for (size_t q = k - 2 * p + 1; q <= k; q++) d += t[q]; // the window summed afresh
		if (d > wmax) wmax = d;
	}
	*y_max = 1.01 * wmax;
	return a;
}

static int check(const char *name, const std::vector<double> &a, double y0, int sgn)
{
	std::vector<double> y;
	long ev0 = g_events_total;
	srs_emulated(a, y0, sgn, y);
	double s = y0;
	size_t bad = 0, first = 0;
	for (size_t i = 0; i < a.size(); i++) {
		s = sgn > 0 ? s + a[i] : s - a[i];
		if (memcmp(&s, &y[i], 8) != 0 && !(std::isnan(s) && std::isnan(y[i]))) { if (!bad) first = i; bad++; }
	}
	printf("%-34s n=%-9zu events=%-6ld %s", name, a.size(), g_events_total - ev0, bad ? "MISMATCH" : "bit-exact\n");
	if (bad) printf(" (%zu, first at %zu)\n", bad, first);
	return bad != 0;
}

int main(int argc, char **argv)
{
	size_t N = argc > 1 ? strtoul(argv[1], 0, 10) : 1000000;
	std::mt19937_64 rng(12345);
	std::uniform_real_distribution<double> U(0.0, 1.0);
	int fails = 0;
	{ // equal masses: the linear-drift case (mprev -= 1/N)
		for (size_t n : {size_t(1000), size_t(100000), N, size_t(1048576)}) {
			std::vector<double> a(n - 1, 1.0 / (double) n);
			fails += check("equal-mass mprev chain (-)", a, 1.0, -1);
			std::vector<double> ones(n, 1.0);
			fails += check("unit masses forward sum (+)", ones, 0.0, +1);
		}
	}
	{ // IMF-like masses
		std::vector<double> m(N);
		double tot = 0;
		for (auto &x : m) { x = 0.08 * pow(1.0 - U(rng) * (1 - pow(0.08 / 150.0, 1.35)), -1.0 / 1.35); tot += x; }
		std::vector<double> a(N);
		for (size_t i = 0; i < N; i++) a[i] = (m[i] / tot * N) / (double) N;
		double y0 = 0; for (auto x : a) y0 += x;
		fails += check("IMF masses mprev chain (-)", a, y0, -1);
		fails += check("IMF masses forward sum (+)", m, 0.0, +1);
	}
	{ // phi-like increasing chain: terms spanning many decades, starting from 0
		std::vector<double> a(N);
		for (size_t i = 0; i < N; i++) a[i] = ((double) (N - i) / N) * exp(-12.0 * U(rng)) * 1e-3 / N * (1 + 1e3 * (double) i / N);
		fails += check("phi-like chain from 0 (+)", a, 0.0, +1);
	}
	{ // adversarial: many exact ties, zeros, a negative and a huge term, a NaN-free inf-free mix
		std::vector<double> a(200000);
		for (size_t i = 0; i < a.size(); i++) {
			int k = rng() % 10;
			a[i] = k == 0 ? 0.0 : k == 1 ? ldexp(1.0, -53) : k == 2 ? ldexp(3.0, -54) : k == 3 ? ldexp(1.0, -52) * (1 + (rng() % 7)) + ldexp(1.0, -53) : U(rng) * 1e-5;
		}
		a[777] = -1e-3; a[9999] = 3.0; a[10000] = 0.5;
		fails += check("adversarial ties (+) from 1", a, 1.0, +1);
		fails += check("adversarial ties (-) from 8", a, 8.0, -1);
	}
	{ // chain that runs through zero and goes negative (all events there)
		std::vector<double> a(5000, 1e-3);
		fails += check("runs below zero (-)", a, 1.0, -1);
	}
	{ // random lengths / scales
		for (int rep = 0; rep < 200; rep++) {
			size_t n = 1 + rng() % 5000;
			double scale = ldexp(1.0, (int) (rng() % 80) - 40);
			std::vector<double> a(n);
			for (auto &x : a) x = U(rng) * scale * ((rng() & 7) ? 1.0 : ldexp(1.0, -(int) (rng() % 40)));
			double y0 = (rng() & 1) ? 0.0 : U(rng) * scale * n;
			std::vector<double> y; srs_emulated(a, y0, +1, y);
			double s = y0; for (size_t i = 0; i < n; i++) { s += a[i]; if (memcmp(&s, &y[i], 8)) { fails++; printf("random(+) rep %d mismatch at %zu\n", rep, i); break; } }
			double y1 = 0; for (auto x : a) y1 += x; y1 *= 1.0 + U(rng);
			srs_emulated(a, y1, -1, y);
			s = y1; for (size_t i = 0; i < n; i++) { s -= a[i]; if (memcmp(&s, &y[i], 8)) { fails++; printf("random(-) rep %d mismatch at %zu\n", rep, i); break; } }
		}
		printf("random chains done\n");
	}
	{ // sliding-window sums (calc_sigma_r): m v^2 falling by orders of magnitude outwards with star-to-star noise, equal masses, IMF masses
		for (size_t p : {size_t(20), size_t(3)}) {
			std::vector<double> t(N / 4);
			double ymax;
			for (size_t i = 0; i < t.size(); i++) t[i] = (1.0 / t.size()) * U(rng) * 2.0 / pow(1.0 + 30.0 * (double) i / t.size(), 2.0);
			{ std::vector<double> ops = window_ops(t, p, &ymax); fails += check_signed("window sum of m v^2 (+-)", ops, ymax); }
			for (auto &x : t) x = 1.0 / (double) t.size();
			{ std::vector<double> ops = window_ops(t, p, &ymax); fails += check_signed("window sum of equal masses (+-)", ops, ymax); }
			for (auto &x : t) x = 1.0 / 39.5 / 1048576.0; // 39 c < 2^-20 < 40 c: every operation crosses a binade boundary
			{ std::vector<double> ops = window_ops(t, 20, &ymax); fails += check_signed("window sum straddling 2^-20 (+-)", ops, ymax); }
			for (auto &x : t) x = 0.08 * pow(1.0 - U(rng) * (1 - pow(0.08 / 150.0, 1.35)), -1.0 / 1.35) / t.size();
			{ std::vector<double> ops = window_ops(t, p, &ymax); fails += check_signed("window sum of IMF masses (+-)", ops, ymax); }
		}
	}
	printf("%s (%d failures)\n", fails ? "FAIL" : "PASS", fails);
	return fails != 0;
}

/* TEST INFRASTRUCTURE ONLY -- see cmc_oracle.h.  CPU restatement of the reference's
 * sort / potential / new-orbit path with explicit arguments.  Compiled with -ffp-contract=off:
 * the reference is built for baseline x86-64 (no FMA), so every double operation rounds once,
 * and `long double` is the x87 80-bit format exactly as in the reference build. */
#include "cmc_oracle.h"
#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ partition ---- */

/* cmc_utils.c:2299-2311 (findIndices): N blocks of blkSize over procs ranks, 1-based inclusive */
static void orc_find_indices(long nblocks, int blk, int procs, int i, int *begin, int *end)
{
	long chunk = (nblocks / procs) * blk;
	long extra = nblocks % procs;
	if (i < extra) {
		*begin = (int) (i * chunk + (long) i * blk + 1);
		*end = (int) (*begin + chunk + blk - 1);
	} else {
		*begin = (int) (i * chunk + extra * blk + 1);
		*end = (int) (*begin + chunk - 1);
	}
}

/* cmc_utils.c:2319-2333 (findLimits): remainder N % blk goes to the last rank */
void orc_find_limits(long N, int blk, int procs, int *Start, int *End)
{
	int nblocks = (int) (N / blk), rem = (int) (N % blk);
	for (int i = 0; i < procs; i++)
		orc_find_indices(nblocks, blk, procs, i, &Start[i], &End[i]);
	End[procs - 1] += rem;
}

/* ------------------------------------------------------------------ potential ---- */

/* cmc_utils.c:675-736 */
long orc_potential_calculate(double *star_r, const double *star_m, double *star_phi, long N_MAX_in, long N_STAR,
			     double *cenma_m, double *cenma_m_new, double madhoc, double orbit_r, double *Mtotal, double *Rtidal)
{
	long k = 1;
	double mprev = 0.0;
	while (star_r[k] < ORC_SF_INFINITY && k <= N_MAX_in) { /* :685-693 forward sequential sum */
		mprev += star_m[k];
		k++;
	}
	long N_MAX = k - 1;                                     /* :696 */
	*cenma_m += *cenma_m_new;                               /* :699-700 */
	*cenma_m_new = 0.0;
	*Mtotal = mprev * madhoc + *cenma_m * madhoc;           /* :703 */
	*Rtidal = orbit_r * pow(*Mtotal, 1.0 / 3.0);            /* :707 */
	star_r[N_MAX + 1] = ORC_SF_INFINITY;                    /* :713-714 */
	star_phi[N_MAX + 1] = 0.0;
	mprev = *Mtotal;
	for (k = N_MAX; k >= 1; k--) {                          /* :717-727 backward sequential recurrence */
		star_phi[k] = star_phi[k + 1] - mprev * (1.0 / star_r[k] - 1.0 / star_r[k + 1]);
		mprev -= star_m[k] / N_STAR;
	}
	star_phi[0] = star_phi[1] + *cenma_m * madhoc / star_r[1]; /* :729 */
	return N_MAX;
}

/* ------------------------------------------------------------------ lookups ---- */

/* cmc_nr.c:62-82: largest k in [kmin,kmax] with star_r[k] < r, probing ktry=(kmin+kmax+1)/2 */
long orc_FindZero_r(const double *star_r, long kmin, long kmax, double r)
{
	do {
		long ktry = (kmin + kmax + 1) / 2;
		if (star_r[ktry] < r) kmin = ktry; else kmax = ktry - 1;
	} while (kmax != kmin);
	return kmin;
}

/* cmc_utils.c:112-165 with SEARCH_GRID=0 (cmc_io.c:1738 default; grid init is commented out, cmc_io.c:3390-3396) */
double orc_potential(const orc_tables_t *t, double r)
{
	const double *sr = t->r, *sp = t->phi;
	if (r == 0.0) return sp[0];                                              /* :117-119 */
	if (r < sr[1]) return sp[0] - (sp[0] - sp[1]) * sr[1] / r;               /* :121-123 */
	long kmax = t->N_MAX + 1, kmin = 1, i;
	i = (kmin == kmax - 1) ? kmin : orc_FindZero_r(sr, kmin, kmax, r);       /* :133-140 */
	if (sr[i] > r || sr[i + 1] < r) {                                        /* :144-152 fatal in the reference */
		fprintf(stderr, "orc_potential: bracket failed i=%ld r=%g\n", i, r);
		return NAN;
	}
	if (i == 0) return sp[1];
	return sp[i] + (sp[i + 1] - sp[i]) * (1.0 / sr[i] - 1.0 / r) / (1.0 / sr[i] - 1.0 / sr[i + 1]); /* :158-161 */
}

/* MPI_PHI_S, include/cmc/cmc.h:2175, for a double radius */
static double phi_s(const orc_tables_t *t, double rad, long j)
{
	return (rad >= t->r[j] ? t->m[j] / rad : t->m[j] / t->r[j]) * (1.0 / t->N_STAR);
}

static inline double sqr_d(double a) { return a * a; } /* SQR under GCC: include/cmc/cmc.h:2026-2028 */

/* cmc_utils.c:1654-1662.  x87 long double; SQR() takes and returns double, so (J/r) is rounded to
 * double before squaring; the self term is long double only when r >= star_r[j] (the division has a
 * long double operand) and plain double otherwise; the result is rounded to double. */
double orc_function_q(const orc_tables_t *t, long j, long double r, long double pot, long double E, long double J)
{
	long double Jr = sqr_d((double) (J / r));
	long double phis = (r >= t->r[j] ? t->m[j] / r : (long double) (t->m[j] / t->r[j])) * (1.0 / t->N_STAR);
	double res = (double) (2.0 * (E - (pot + phis)) - Jr);
	return res;
}

/* cmc_nr.c:316-383 */
long orc_find_zero_Q(const orc_tables_t *t, long j, long kmin, long kmax, long double E, long double J)
{
	long double qmin = orc_function_q(t, j, t->r[kmin], t->phi[kmin], E, J);
	long double qmax = orc_function_q(t, j, t->r[kmax], t->phi[kmax], E, J);
	int increasing = qmin < qmax;                                            /* :342 */
	do {
		long ktry = (kmin + kmax + 1) / 2;
		long double q = orc_function_q(t, j, t->r[ktry], t->phi[ktry], E, J);
		int go_up = increasing ? (q < 0.e0) : (q > 0.e0);                /* :349 / :362 */
		if (go_up) kmin = ktry; else kmax = ktry - 1;
	} while (kmax != kmin);
	return kmin;
}

/* cmc_orbit.c:722-740 */
static void set_a_b(const orc_tables_t *t, long index, long k, double *a, double *b)
{
	if (k == 0) {
		*a = -t->cenma_m * t->madhoc;
		*b = t->phi[1] + t->cenma_m * t->madhoc / t->r[1] + t->m[index] * t->madhoc / t->r[index];
	} else {
		double rk = t->r[k], rk1 = t->r[k + 1];
		double Uk = t->phi[k] + phi_s(t, rk, index);
		double Uk1 = t->phi[k + 1] + phi_s(t, rk1, index);
		*a = (Uk1 - Uk) / (1. / rk1 - 1. / rk);
		*b = (Uk / rk1 - Uk1 / rk) / (1. / rk1 - 1. / rk);
	}
}

/* cmc_nr.c:279-302 (calc_pot_within_interval) + :393-405 (calc_vr_within_interval) */
static double vr2_within_interval(const orc_tables_t *t, long index, long k, double E, double J, double r)
{
	double pot;
	if (fabs(r - t->r[k]) < DBL_EPSILON) pot = t->phi[k];
	else if (fabs(r - t->r[k + 1]) < DBL_EPSILON) pot = t->phi[k + 1];
	else if (r < t->r[1]) pot = t->phi[0] - t->cenma_m * t->madhoc / r;
	else pot = t->phi[k] + (t->phi[k + 1] - t->phi[k]) * (1.0 / t->r[k] - 1.0 / r) / (1.0 / t->r[k] - 1.0 / t->r[k + 1]);
	return orc_function_q(t, index, r, pot, E, J);
}

/* cmc_nr.c:500-603: Brent-Dekker on [r_k, r_k+1] (GSL gsl_root_fsolver_brent, un-vendored dependency, version
 * unpinned -- PARITY UNPINNED for this branch), stop on |hi-lo| < eps + eps*min(|lo|,|hi|) with
 * eps = APSIDES_PRECISION = 1e-11 (cmc_io.c:1752), at most APSIDES_MAX_ITER = 100 iterations; in the last 10 the
 * APSIDES_CONVERGENCE = 5e-13 delta test also applies (:556-569); then pick the end with vr^2 >= 0 (:590-597). */
double orc_find_root_vr(const orc_tables_t *t, long index, long k, double E, double J)
{
	const double PREC = 1.0e-11, CONV = 5.e-13;
	long iter = 100;
	double a = t->r[k], b = t->r[k + 1], c = b;
	double fa = vr2_within_interval(t, index, k, E, J, a), fb = vr2_within_interval(t, index, k, E, J, b), fc = fb;
	double d = b - a, e = b - a, root = 0.5 * (a + b), lo = a, hi = b, prev_apsis = -1.0;
	int not_converged = 1;
	while (not_converged && iter) {
		int ac_equal = 0;
		double tol, m;
		if ((fb < 0 && fc < 0) || (fb > 0 && fc > 0)) { ac_equal = 1; c = a; fc = fa; d = b - a; e = b - a; }
		if (fabs(fc) < fabs(fb)) { ac_equal = 1; a = b; b = c; c = a; fa = fb; fb = fc; fc = fa; }
		tol = 0.5 * DBL_EPSILON * fabs(b);
		m = 0.5 * (c - b);
		if (fb == 0) { root = lo = hi = b; }
		else if (fabs(m) <= tol) { root = b; if (b < c) { lo = b; hi = c; } else { lo = c; hi = b; } }
		else {
			if (fabs(e) < tol || fabs(fa) <= fabs(fb)) { d = m; e = m; }
			else {
				double p, q, r, s = fb / fa;
				if (ac_equal) { p = 2 * m * s; q = 1 - s; }
				else { q = fa / fc; r = fb / fc; p = s * (2 * m * q * (q - r) - (b - a) * (r - 1)); q = (q - 1) * (r - 1) * (s - 1); }
				if (p > 0) q = -q; else p = -p;
				double lim1 = 3 * m * q - fabs(tol * q), lim2 = fabs(e * q);
				if (2 * p < (lim1 < lim2 ? lim1 : lim2)) { e = d; d = p / q; } else { d = m; e = m; }
			}
			a = b; fa = fb;
			if (fabs(d) > tol) b += d; else b += (m > 0 ? +tol : -tol);
			fb = vr2_within_interval(t, index, k, E, J, b);
			root = b;
			if ((fb < 0 && fc < 0) || (fb > 0 && fc > 0)) c = a;
			if (b < c) { lo = b; hi = c; } else { lo = c; hi = b; }
		}
		{
			double al = fabs(lo), ah = fabs(hi);
			double min_abs = ((lo > 0.0 && hi > 0.0) || (lo < 0.0 && hi < 0.0)) ? (al < ah ? al : ah) : 0;
			not_converged = !(fabs(hi - lo) < PREC + PREC * min_abs);
		}
		iter--;
		if (iter < 10) {
			if (prev_apsis < 0.) prev_apsis = root;
			else {
				int delta_continue = !(fabs(root - prev_apsis) < CONV + CONV * fabs(root) || root == prev_apsis);
				not_converged = not_converged && delta_continue;
				prev_apsis = root;
			}
		}
	}
	double apsis = root;
	if (vr2_within_interval(t, index, k, E, J, apsis) < 0.) {
		if (vr2_within_interval(t, index, k, E, J, lo) < 0.) apsis = hi;
		if (vr2_within_interval(t, index, k, E, J, hi) < 0.) apsis = lo;
	}
	return apsis;
}

/* cmc_orbit.c:30-59 */
static long get_positive_Q_index(const orc_tables_t *t, long g_si, int is_new, double rnew, double E, double J)
{
	long ktemp = g_si;
	if (is_new) ktemp = orc_FindZero_r(t->r, 0, t->N_MAX, rnew) + 1;         /* :39-41 */
	double Q = orc_function_q(t, g_si, t->r[ktemp], t->phi[ktemp], E, J);
	if (Q <= 0.0) {                                                           /* :46-57 linear scan from 0 */
		ktemp = -1;
		do {
			ktemp++;
			Q = orc_function_q(t, g_si, t->r[ktemp], t->phi[ktemp], E, J);
		} while (Q <= 0.0 && ktemp <= t->N_MAX);
		if (ktemp >= t->N_MAX) return -1;
	}
	return ktemp;
}

/* cmc_orbit.c:413-540 */
void orc_calc_orbit_new(const orc_tables_t *t, long g_si, int is_new, double rnew_of_new, double E, double J, orc_orbit_t *o)
{
	o->used_brent = 0;
	long ktemp = get_positive_Q_index(t, g_si, is_new, rnew_of_new, E, J);
	int circular = 0;
	double rmin = 0, rmax = 0, dQdr_min = 0, dQdr_max = 0, a, b;
	long kmin = 0, kmax = 0;
	if (ktemp <= -1) {
		circular = 1;                                                     /* :424-438 */
	} else {
		kmin = orc_find_zero_Q(t, g_si, 0, ktemp, E, J);                  /* :441 */
		kmax = orc_find_zero_Q(t, g_si, ktemp, t->N_MAX + 1, E, J);       /* :444 */
		double rkmin = t->r[kmin], rk1min = t->r[kmin + 1], rkmax = t->r[kmax], rk1max = t->r[kmax + 1];
		set_a_b(t, g_si, kmin, &a, &b);                                   /* :456 */
		double inside = a * a - 2.0 * J * J * (b - E);
		if (inside < 0.0) {                                               /* :459-474 */
			rmin = -1.0 * J * J / a;
			if (rmin < rkmin || rmin > rk1min) { rmin = orc_find_root_vr(t, g_si, kmin, E, J); o->used_brent++; }
		} else {
			rmin = J * J / (-a + sqrt(inside));                       /* :476 */
		}
		dQdr_min = 2.0 * J * J / (rmin * rmin * rmin) + 2.0 * a / (rmin * rmin); /* :479 */
		set_a_b(t, g_si, kmax, &a, &b);                                   /* :482 */
		inside = a * a - 2.0 * J * J * (b - E);
		if (inside < 0.0) {                                               /* :485-500 */
			rmax = -a / (2.0 * (b - E));
			if (rmax < rkmax || rmax > rk1max) { rmax = orc_find_root_vr(t, g_si, kmax, E, J); o->used_brent++; }
		} else {
			rmax = (-a + sqrt(inside)) / (2.0 * (b - E));             /* :502 */
		}
		dQdr_max = 2.0 * J * J / (rmax * rmax * rmax) + 2.0 * a / (rmax * rmax); /* :505 */
		if (rmin > rmax || dQdr_min <= 0. || dQdr_max >= 0.) circular = 1; /* :510 */
	}
	if (circular) {                                                           /* :520-528 */
		o->rp = o->ra = t->r[g_si];
		o->kmin = o->kmax = g_si;
		o->dQdrp = o->dQdra = 0.0;
		o->circular_flag = 1;
	} else {
		o->rp = rmin; o->ra = rmax; o->kmin = kmin; o->kmax = kmax;
		o->dQdrp = dQdr_min; o->dQdra = dQdr_max; o->circular_flag = 0;
	}
}

/* ------------------------------------------------------------------ orbit step ---- */

/* cmc_evolution_thr.c:784-937 */
double orc_get_positions(const orc_tables_t *t, long n_local, const long *g_idx, const unsigned char *is_new,
			 const double *E_in, const double *J_in, const double *vr_in, const double *vt_in, const double *rnew_in,
			 double Rtidal, const double *draws, long n_draws, const long *csr_off_in, long *draw_off_out,
			 int *status, double *rnew, double *vrnew, double *vtnew, double *X_out, double *r_peri, double *r_apo,
			 long *kmin_out, long *kmax_out, long *n_brent)
{
	double max_rad = 0.0;
	long pos = 0, brent = 0;
	for (long j = 0; j < n_local; j++) {
		long g_j = g_idx[j];
		if (csr_off_in) pos = csr_off_in[j];
		if (draw_off_out) draw_off_out[j] = pos;
		double E = E_in[j] + phi_s(t, t->r[g_j], g_j);                    /* :806 */
		double J = J_in[j];
		if (kmin_out) kmin_out[j] = -1;
		if (kmax_out) kmax_out[j] = -1;
		X_out[j] = r_peri[j] = r_apo[j] = NAN;                            /* "not written" marker for the caller */
		if (t->m[g_j] < ORC_ZERO) {                                       /* :816-820 */
			status[j] = ORC_ST_MASSLESS;
			rnew[j] = ORC_SF_INFINITY; vrnew[j] = 0.0; vtnew[j] = 0.0;  /* remove_star_center :760-765 */
			continue;
		}
		if (E >= 0.0) {                                                   /* :823-829 */
			status[j] = ORC_ST_UNBOUND;
			rnew[j] = ORC_SF_INFINITY; vrnew[j] = 0.0; vtnew[j] = 0.0;  /* remove_star :662-664 */
			continue;
		}
		orc_orbit_t o;
		orc_calc_orbit_new(t, g_j, is_new ? is_new[j] : 0, rnew_in ? rnew_in[j] : 0.0, E, J, &o); /* :832 */
		brent += o.used_brent;
		if (kmin_out) kmin_out[j] = o.kmin;
		if (kmax_out) kmax_out[j] = o.kmax;
		if (o.circular_flag == 1) {                                       /* :835-839 */
			status[j] = ORC_ST_CIRCULAR;
			rnew[j] = t->r[g_j]; vrnew[j] = vr_in[j]; vtnew[j] = vt_in[j];
			continue;
		}
		double rmin = o.rp, rmax = o.ra, dQdr_min = o.dQdrp, dQdr_max = o.dQdra;
		if (rmax >= Rtidal) {                                             /* :848-855 */
			status[j] = ORC_ST_TIDAL;
			r_apo[j] = rmax; r_peri[j] = rmin;
			rnew[j] = ORC_SF_INFINITY; vrnew[j] = 0.0; vtnew[j] = 0.0;
			continue;
		}
		double g1 = sqrt(3.0 * (rmax - rmin) / dQdr_min);                 /* :857-860 */
		double g2 = sqrt(-3.0 * (rmax - rmin) / dQdr_max);
		double F = 1.2 * (g1 >= g2 ? g1 : g2);
		double r = 0.0, vr = 0.0, X = 0.0;
		long k;
		for (k = 1; k <= ORC_N_TRY; k++) {                                /* :862-888 */
			if (pos + 2 > (csr_off_in ? csr_off_in[j + 1] : n_draws)) { fprintf(stderr, "orc_get_positions: out of draws at star %ld\n", j); status[j] = ORC_ST_NTRY; break; }
			X = draws[pos++];
			double s0 = 2.0 * X - 1.0;
			double g0 = F * draws[pos++];
			r = 0.5 * (rmin + rmax) + 0.25 * (rmax - rmin) * (3.0 * s0 - s0 * s0 * s0);
			double pot = orc_potential(t, r) + phi_s(t, r, g_j);
			double drds = 0.25 * (rmax - rmin) * (3.0 - 3.0 * s0 * s0);
			double Q = 2.0 * E - 2.0 * pot - J * J / r / r;
			if (Q >= 0.0) vr = sqrt(Q); else vr = 0;
			if (g0 < 1.0 / vr * drds) break;
		}
		if (k == ORC_N_TRY + 1 || status[j] == ORC_ST_NTRY) { status[j] = ORC_ST_NTRY; rnew[j] = NAN; vrnew[j] = vtnew[j] = NAN; continue; }
		X_out[j] = X; r_peri[j] = rmin; r_apo[j] = rmax;                  /* :918-920 */
		if (draws[pos++] < 0.5) vr = -vr;                                 /* :923-924 */
		status[j] = ORC_ST_OK;
		rnew[j] = r; vrnew[j] = vr; vtnew[j] = J / r;                     /* :926-931 */
		if (r > max_rad) max_rad = r;
	}
	if (draw_off_out) draw_off_out[n_local] = pos;
	if (n_brent) *n_brent = brent;
	return max_rad;
}

/* ------------------------------------------------------------------ sort ---- */

typedef struct { double key; long idx; } keyidx_t;
static int cmp_keyidx(const void *a, const void *b)
{
	const keyidx_t *x = a, *y = b;
	if (x->key < y->key) return -1;       /* compare_type, cmc_sort.c:245-250 */
	if (x->key > y->key) return 1;
	return (x->idx > y->idx) - (x->idx < y->idx); /* ties: original order (libc qsort leaves this unspecified; cmc_utils.c:670 assumes none) */
}

long orc_sort_stars(unsigned char *star1, long n, long sizeof_star, long key_off, long binind_off,
		    unsigned char *binary1, long n_bin_dim, long sizeof_binary, long *n_bin_out, long *perm_out)
{
	keyidx_t *ki = malloc((n > 0 ? n : 1) * sizeof(*ki));
	for (long i = 0; i < n; i++) { memcpy(&ki[i].key, star1 + i * sizeof_star + key_off, 8); ki[i].idx = i; }
	qsort(ki, n, sizeof(*ki), cmp_keyidx);                                    /* cmc_sort.c:407 */
	unsigned char *tmp = malloc((n > 0 ? n : 1) * sizeof_star);
	for (long i = 0; i < n; i++) memcpy(tmp + i * sizeof_star, star1 + ki[i].idx * sizeof_star, sizeof_star);
	long n_new = n;
	while (n_new > 0 && ki[n_new - 1].key == ORC_SF_INFINITY) n_new--;       /* remove_stripped_stars, :343-354 */
	/* binaries follow their star: packed in star order (:521-538, 716-729), numbered from 1 (:571-579, 757-766),
	 * the rest of binary[1..n_bin_dim-1] zeroed (:541) */
	long nb = 0;
	if (binary1) {
		unsigned char *btmp = malloc((size_t) (n_bin_dim > 0 ? n_bin_dim : 1) * sizeof_binary);
		for (long i = 0; i < n_new; i++) {
			long bi; memcpy(&bi, tmp + i * sizeof_star + binind_off, 8);
			if (bi > 0) {
				memcpy(btmp + nb * sizeof_binary, binary1 + (bi - 1) * sizeof_binary, sizeof_binary);
				nb++;
				memcpy(tmp + i * sizeof_star + binind_off, &nb, 8);
			}
		}
		memset(binary1, 0, (size_t) (n_bin_dim - 1) * sizeof_binary);
		memcpy(binary1, btmp, (size_t) nb * sizeof_binary);
		free(btmp);
	}
	memcpy(star1, tmp, (size_t) n * sizeof_star);
	if (perm_out) for (long i = 0; i < n; i++) perm_out[i] = ki[i].idx;
	if (n_bin_out) *n_bin_out = nb;
	free(tmp); free(ki);
	return n_new;
}

/* ------------------------------------------------------------------ taus113 ---- */

const unsigned int orc_JPoly_2_20[4] = {0x0c382e31, 0x1b040425, 0x0b49a509, 0x0173f6b0}; /* taus113-v2.c:70 */
const unsigned int orc_JPoly_2_80[4] = {0x487cf69c, 0x00be6310, 0x04bfe2bb, 0x000824f9}; /* taus113-v2.c:74 */

/* one step of the four Tausworthe components (L'Ecuyer 1999 taus113; taus113-v2.c:166-181):
 * component c: b = ((z << q) ^ z) >> (k - s);  z = ((z & mask) << s) ^ b */
static void t113_step(unsigned int z[4])
{
	static const int q[4] = {6, 2, 13, 3}, rs[4] = {13, 27, 21, 12}, s[4] = {18, 2, 7, 13};
	static const unsigned int mask[4] = {4294967294u, 4294967288u, 4294967280u, 4294967168u};
	for (int c = 0; c < 4; c++) {
		unsigned int b = ((z[c] << q[c]) ^ z[c]) >> rs[c];
		z[c] = ((z[c] & mask[c]) << s[c]) ^ b;
	}
}

double orc_t113_dbl(unsigned int z[4])
{
	t113_step(z);
	return (double) (z[0] ^ z[1] ^ z[2] ^ z[3]) / 4294967296.0;              /* :191-193 */
}

void orc_t113_seed(unsigned long s, unsigned int z[4])                        /* :197-216 */
{
	static const unsigned int lo[4] = {2, 8, 16, 128};
	if (s == 0) s = 1;
	unsigned long v = s;
	for (int c = 0; c < 4; c++) {
		v = (69069UL * v) & 0xffffffffUL;
		if (v < lo[c]) v += lo[c];
		z[c] = (unsigned int) v;
	}
	for (int i = 0; i < 10; i++) t113_step(z);
}

void orc_t113_jump(unsigned int z[4], const unsigned int jpoly[4])            /* :253-270 */
{
	unsigned int acc[4] = {0, 0, 0, 0}, st[4] = {z[0], z[1], z[2], z[3]};
	for (int bit = 0; bit < 32; bit++) {
		for (int c = 0; c < 4; c++)
			if (jpoly[c] & (1u << bit)) acc[c] ^= st[c];
		t113_step(st);
	}
	memcpy(z, acc, sizeof(acc));
}

/* n consecutive rng_t113_dbl_new() values (convenience for the tests) */
void orc_t113_fill(unsigned int z[4], double *out, long n)
{
	for (long i = 0; i < n; i++) out[i] = orc_t113_dbl(z);
}

/* ---- comp_mass_percent, src/cmc/cmc_utils.c:872-985 ------------------------------------------------------------------------------ */

#define ORC_PI 3.14159265358979323 /* include/cmc/cmc.h:62 */

long orc_comp_mass_percent(const double *star_r, const double *star_m, long N_MAX, long N_STAR, const double *vr, const double *vt, long n_local,
			   long first_global, const double prefix[4], int is_rank0, double cenma_m, double madhoc, double Mtotal,
			   double initial_total_mass, int bh_include, double minimum_r, const double *mass_pc, long count, double *mass_r,
			   double *ave_mass_r, double *no_star_r, double *densities_r, double *ke_rad_r, double *ke_tan_r, double *v2_rad_r,
			   double *v2_tan_r, int *owner, double totals[4])
{
	double mprev;
	long k, mcount = 0;
	/* :876-895 */
	if (bh_include) {
		mprev = cenma_m * madhoc;
		for (mcount = 0; mcount < count; mcount++) {
			if (mprev / Mtotal > mass_pc[mcount]) {
				mass_r[mcount] = minimum_r; ave_mass_r[mcount] = 0.0; no_star_r[mcount] = 0; densities_r[mcount] = 0.0;
				ke_rad_r[mcount] = 0.0; ke_tan_r[mcount] = 0.0; v2_rad_r[mcount] = 0.0; v2_tan_r[mcount] = 0.0;
			} else break;
		}
	} else {
		mprev = 0.;
		mcount = 0;
	}
	if (owner) for (k = 0; k < count; k++) owner[k] = 0;
	/* :899-913 the rank-local cumulative sums, every intermediate value kept */
	double *a0 = (double *) calloc((size_t) n_local + 1, sizeof(double)), *a1 = (double *) calloc((size_t) n_local + 1, sizeof(double));
	double *a2 = (double *) calloc((size_t) n_local + 1, sizeof(double)), *a3 = (double *) calloc((size_t) n_local + 1, sizeof(double));
	for (k = 1; k <= n_local; k++) {
		long g_k = first_global + k - 1;
		a0[k] = a0[k - 1] + 0.5 * star_m[g_k] * madhoc * vr[k - 1] * vr[k - 1];
		a1[k] = a1[k - 1] + 0.5 * star_m[g_k] * madhoc * vt[k - 1] * vt[k - 1];
		a2[k] = a2[k - 1] + vr[k - 1] * vr[k - 1];
		a3[k] = a3[k - 1] + vt[k - 1] * vt[k - 1];
	}
	totals[0] = a0[n_local]; totals[1] = a1[n_local]; totals[2] = a2[n_local]; totals[3] = a3[n_local];
	/* :933-977 */
	if (mcount < count) {
		for (k = 1; k <= N_MAX; k++) {
			mprev += star_m[k] / N_STAR;
			if (mprev > mass_pc[mcount] * Mtotal) {
				mass_r[mcount] = star_r[k];
				ave_mass_r[mcount] = mprev / Mtotal / k * initial_total_mass;
				no_star_r[mcount] = k;
				densities_r[mcount] = mprev / (4.0 / 3.0 * ORC_PI * pow(star_r[k], 3));
				if (k >= first_global && k < first_global + n_local) { /* this rank owns k (findProcForIndex) */
					long l = k - first_global + 1;                       /* get_local_idx */
					if (is_rank0) { ke_rad_r[mcount] = a0[l]; ke_tan_r[mcount] = a1[l]; v2_rad_r[mcount] = a2[l]; v2_tan_r[mcount] = a3[l]; }
					else {
						ke_rad_r[mcount] = prefix[0] + a0[l]; ke_tan_r[mcount] = prefix[1] + a1[l];
						v2_rad_r[mcount] = prefix[2] + a2[l]; v2_tan_r[mcount] = prefix[3] + a3[l];
					}
					if (owner) owner[mcount] = 1;
				}
				mcount++;
				if (mcount == count) break;
			}
		}
	}
	free(a0); free(a1); free(a2); free(a3);
	return mcount;
}

/* ---- energy_conservation2 / ComputeIntermediateEnergy, src/cmc/cmc_utils.c:2217-2226, 368-378 ---------------------------------- */

static double orc_mpi_phi_s(const orc_tables_t *t, double rad, long j) /* MPI_PHI_S, include/cmc/cmc.h:2175 */
{
	return (rad >= t->r[j] ? t->m[j] / rad : t->m[j] / t->r[j]) * (1.0 / t->N_STAR);
}

void orc_energy_conservation2(const orc_tables_t *t, long n_local, const long *g_idx, const double *rnew, double *Uoldrnew)
{
	for (long i = 0; i < n_local; i++) Uoldrnew[i] = orc_potential(t, rnew[i]) + orc_mpi_phi_s(t, rnew[i], g_idx[i]);
}

void orc_intermediate_energy(const orc_tables_t *t, long n_local, const long *g_idx, const double *rnew, const double *vr, const double *vt, double *EI)
{
	for (long j = 0; j < n_local; j++)
		if (rnew[j] < 1.0e6) EI[j] = vr[j] * vr[j] + vt[j] * vt[j] + t->phi[g_idx[j]] - orc_potential(t, rnew[j]);
}

/* ---- energy_conservation1 / set_velocities3 (energy_conservation3), src/cmc/cmc_utils.c:2153-2174, 254-352 -------------------------- */

void orc_energy_conservation1(const orc_tables_t *t, long n_local, const long *g_idx, const double *vr, const double *vt, double *vrold,
			      double *vtold, double *Uoldrold)
{
	for (long i = 0; i < n_local; i++) {
		vtold[i] = vt[i];
		vrold[i] = vr[i];
		Uoldrold[i] = t->phi[g_idx[i]] + orc_mpi_phi_s(t, t->r[g_idx[i]], g_idx[i]); /* :2168 */
	}
}

#define ORC_MIN(a, b) ((a) <= (b) ? (a) : (b)) /* include/cmc/cmc.h:2242 */

double orc_set_velocities3_first(const orc_tables_t *t, long n_local, const long *g_idx, const long *interacted, const double *rOld,
				 const double *vrold, const double *vtold, const double *Uoldrold, const double *Uoldrnew, double *vr, double *vt,
				 double *vrnew, double *vtnew, char *vnew2_arr)
{
	const double q = 0.5, E_dump_factor = 0.8; /* :258, :264 */
	double Eexcess = 0.0;
	for (long i = 0; i < n_local; i++) {
		const long g = g_idx[i];
		const double m = t->m[g];
		if (interacted[i] != 0) continue; /* :273 */
		const double Unewrold = orc_potential(t, rOld[i]) + orc_mpi_phi_s(t, rOld[i], g);
		const double Unewrnew = t->phi[g] + orc_mpi_phi_s(t, t->r[g], g);
		const double vold2 = vtold[i] * vtold[i] + vrold[i] * vrold[i];
		const double vnew2 = vold2 + 2.0 * (1.0 - q) * (Uoldrold[i] - Uoldrnew[i]) + 2.0 * q * (Unewrold - Unewrnew); /* :281-282 */
		vnew2_arr[i] = (vnew2 > 0.0) ? 1 : 0;
		if (vnew2 <= 0.0) {
			Eexcess += 0.5 * (vr[i] * vr[i] + vt[i] * vt[i] - vnew2) * m; /* :288 */
		} else {
			const double alpha = sqrt(vnew2 / (vr[i] * vr[i] + vt[i] * vt[i])); /* :291 */
			vr[i] *= alpha;
			vt[i] *= alpha;
			const double v2 = vr[i] * vr[i] + vt[i] * vt[i];
			vrnew[i] = vr[i];
			vtnew[i] = vt[i];
			const double E_dump_capacity = E_dump_factor * 0.5 * m * v2; /* :300 */
			const double E_dump = ORC_MIN(E_dump_capacity, Eexcess);
			if (Eexcess > 0) {
				const double exc_ratio = sqrt((v2 - 2 * E_dump / m) / v2);
				vr[i] *= exc_ratio;
				vt[i] *= exc_ratio;
				Eexcess -= E_dump;
			}
		}
	}
	return Eexcess;
}

double orc_set_velocities3_second(const orc_tables_t *t, long n_local, const long *g_idx, const long *interacted, const char *vnew2_arr,
				  double Eexcess, double *vr, double *vt, const double *vrnew, const double *vtnew)
{
	const double E_dump_factor = 0.8;
	for (long i = 0; i < n_local; i++) {
		const double m = t->m[g_idx[i]];
		const double v2_new = vrnew[i] * vrnew[i] + vtnew[i] * vtnew[i];
		const double v2 = vr[i] * vr[i] + vt[i] * vt[i];
		if (interacted[i] == 0 && vnew2_arr[i] == 1) {
			const double E_dump_capacity = v2 * m * 0.5 - E_dump_factor * 0.5 * m * v2_new; /* :333 */
			const double E_dump = ORC_MIN(E_dump_capacity, Eexcess);
			if (Eexcess > 0) {
				const double exc_ratio = sqrt((v2 - 2 * E_dump / m) / v2);
				vr[i] *= exc_ratio;
				vt[i] *= exc_ratio;
				Eexcess -= E_dump;
			}
		}
	}
	return Eexcess;
}

/* ---- calc_sigma_r, src/cmc/cmc_dynamics_helper.c:2821-2975 ----------------------------------------------------------------------- */

void orc_calc_sigma_r(long p, long n_limit, long first_global, long N_MAX, int is_first_rank, const double *star_r, const double *star_m,
		      double madhoc, const double *vr, const double *vt, const double *ghost_prev_vr, const double *ghost_prev_vt,
		      const double *ghost_next_vr, const double *ghost_next_vt, int r_0_mave_1, double *sig_r_or_mave, double *sig_sigma)
{
	long siminlast = 1, simaxlast = is_first_rank ? 0 : -p; /* :2906-2910 */
	double Mv2ave = 0.0, Mave = 0.0;
	for (long si = 1; si <= n_limit; si++) {
		const long g_si = first_global + si - 1;
		long simin = si - p, simax = simin + (2 * p - 1);
		const long g_simin = g_si - p, g_simax = g_simin + (2 * p - 1);
		if (g_simin < 1) { simin = 1; simax = simin + (2 * p - 1); }                 /* :2923-2926 */
		else if (g_simax > N_MAX) { simax = n_limit; simin = simax - (2 * p - 1); } /* :2927-2931 */
		for (long k = siminlast; k < simin; k++) { /* the stars leaving the window, :2935-2949 */
			const double a = k < 1 ? ghost_prev_vr[k + p - 1] : vr[k - 1], b = k < 1 ? ghost_prev_vt[k + p - 1] : vt[k - 1];
			const long g_k = first_global + k - 1;
			Mv2ave -= star_m[g_k] * madhoc * (a * a + b * b);
			Mave -= star_m[g_k] * madhoc;
		}
		for (long k = simaxlast + 1; k <= simax; k++) { /* the stars entering it, :2951-2968 */
			const long g_k = first_global + k - 1;
			double a, b;
			if (k > n_limit) { a = ghost_next_vr[k - n_limit - 1]; b = ghost_next_vt[k - n_limit - 1]; }
			else if (k < 1) { a = ghost_prev_vr[k + p - 1]; b = ghost_prev_vt[k + p - 1]; }
			else { a = vr[k - 1]; b = vt[k - 1]; }
			Mv2ave += star_m[g_k] * madhoc * (a * a + b * b);
			Mave += star_m[g_k] * madhoc;
		}
		sig_r_or_mave[si] = r_0_mave_1 == 0 ? star_r[g_si] : Mave / 2. / p; /* :2971-2974 */
		sig_sigma[si] = sqrt(Mv2ave / Mave);                                 /* :2977 */
		siminlast = simin;
		simaxlast = simax;
	}
}

/* ---- dynamics_apply, src/cmc/cmc_dynamics.c:206-500: the pair loop without the strong encounters ----------------------------------------------- */

#define ORC_PI 3.14159265358979323      /* include/cmc/cmc.h:62 */
#define ORC_FB_PI 3.141592653589793238462643 /* include/fewbody-0.24/fewbody.h:35 */
#define ORC_MSUN 1.989e+33              /* include/cmc/cmc.h:66, fewbody.h:38 */
#define ORC_XBS 2.0                     /* include/cmc/cmc.h:83 */
#define ORC_XBB 2.0                     /* include/cmc/cmc.h:85 */
static double orc_sqr(double x) { return x * x; }
static double orc_cub(double x) { return x * x * x; }

/* calc_n_local, cmc_dynamics_helper.c:768-786 */
static double orc_calc_n_local(const double *star_r, long k, long p, long N_LIMIT)
{
	long kmin = k - p, kmax = k + p;
	if (kmin < 0) { kmin = 0; kmax = 2 * p + 1; }
	else if (kmax > N_LIMIT) { kmax = N_LIMIT; kmin = N_LIMIT - 2 * p - 1; }
	return (2.0 * ((double) p)) * 3.0 / (4.0 * ORC_PI * (orc_cub(star_r[kmax]) - orc_cub(star_r[kmin])));
}

long orc_dynamics_apply_pairs(const orc_tables_t *t, long n_local, long first_global, double dt, const orc_dyn_params_t *dp, const double *draws,
			      double *vr, double *vt, double *E, double *J, double *Y, const long *binind, const int *se_k, const double *rad,
			      const double *bin_a, const long *threebb, long *pair_k, long *pair_kp, int *pair_status, double *pair_rperi,
			      double *pair_W, double *pair_rcm, orc_relax_stats_t *stats)
{
	const double relbeta[4] = {ORC_PI / 2.0, ORC_PI / 4.0, ORC_PI / 8.0, ORC_PI / 16.0}; /* :28 */
	const double madhoc = t->madhoc, N_STAR = (double) t->N_STAR;
	const long N_LIMIT = t->N_MAX, p = dp->p;
	long npairs = 0, si = 1;
	memset(stats, 0, sizeof *stats);
	while (si <= n_local - n_local % 2 - 1) { /* :216 */
		long k, kp, j;
		double v[4], vp[4], w[4], W, rcm, rperi = 0.0, S, n_local_dens, mass_k, mass_kp, P_enc;
		while (threebb && threebb[si - 1] == 1) si += 1; /* :226-228 */
		k = si;
		si += 1;
		while (threebb && threebb[si - 1] == 1) si += 1; /* :233-236 */
		kp = si;
		si += 1;
		const long g_k = first_global + k - 1, g_kp = first_global + kp - 1; /* get_global_idx of a star that is not new */
		/* calc_encounter_dyns(k, kp, ..., setY = 1), cmc_dynamics_helper.c:831-870 */
		Y[k - 1] = draws[3 * npairs];
		Y[kp - 1] = Y[k - 1];
		const double phi = Y[k - 1] * 2.0 * ORC_PI;
		v[1] = vt[k - 1]; v[2] = 0.0; v[3] = vr[k - 1];
		vp[1] = vt[kp - 1] * cos(phi); vp[2] = vt[kp - 1] * sin(phi); vp[3] = vr[kp - 1];
		for (j = 1; j <= 3; j++) w[j] = vp[j] - v[j];
		W = sqrt(orc_sqr(w[1]) + orc_sqr(w[2]) + orc_sqr(w[3]));
		rcm = (t->m[g_k] * t->r[g_k] + t->m[g_kp] * t->r[g_kp]) / (t->m[g_k] + t->m[g_kp]);
		n_local_dens = orc_calc_n_local(t->r, g_k, p, N_LIMIT); /* :249 */
		mass_k = t->m[g_k];
		mass_kp = t->m[g_kp];
		const long bk = binind[k - 1], bkp = binind[kp - 1];
		if (bk > 0 && bkp > 0) { /* :254-262 */
			rperi = ORC_XBB * (bin_a[bk] + bin_a[bkp]);
			S = dp->binbin ? ORC_PI * orc_sqr(rperi) * (1.0 + 2.0 * madhoc * (mass_k + mass_kp) / (rperi * orc_sqr(W))) : 0.0;
		} else if (bk > 0 || bkp > 0) { /* :263-280 */
			rperi = ORC_XBS * bin_a[bk > 0 ? bk : bkp];
			S = dp->binsingle ? ORC_PI * orc_sqr(rperi) * (1.0 + 2.0 * madhoc * (mass_k + mass_kp) / (rperi * orc_sqr(W))) : 0.0;
		} else if (dp->ss_collision) { /* :282-428 */
			const int kk = se_k[k - 1], kkp = se_k[kp - 1];
			const double rk = rad[k - 1], rkp = rad[kp - 1];
			double S_tmp = 0.0, collisions_multiple, S_coll;
			/* :285-303 is overwritten by :305-313 whatever CO_TDE says: WD_TC > 0 and two white dwarfs -> WD_TC, everything else COLL_FACTOR */
			if (dp->wd_tc > 0 && (kkp >= 10 && kkp <= 12) && (kk >= 10 && kk <= 12)) collisions_multiple = dp->wd_tc;
			else collisions_multiple = dp->coll_factor;
			if (dp->tidal_capture) { /* :360-374 (Lombardi et al. 2006) */
				double S_lombardi;
				if ((kk <= 1 || kk >= 10 || kk == 7) && (kkp >= 2 && kkp <= 9 && kkp != 7)) {
					rperi = 1.3 * rkp;
					S_lombardi = ORC_PI * orc_sqr(rperi) * (1.0 + 2.0 * madhoc * (mass_k + mass_kp) / (rperi * orc_sqr(W)));
				} else if ((kkp <= 1 || kkp >= 10 || kkp == 7) && (kk >= 2 && kk <= 9 && kk != 7)) {
					rperi = 1.3 * rk;
					S_lombardi = ORC_PI * orc_sqr(rperi) * (1.0 + 2.0 * madhoc * (mass_k + mass_kp) / (rperi * orc_sqr(W)));
				} else S_lombardi = 0.0;
				S_tmp = S_tmp > S_lombardi ? S_tmp : S_lombardi;
			}
			if (dp->bh_capture && kk == 14 && kkp == 14) { /* :378-389 (Quinlan & Shapiro) */
				const double clight10o7 = pow(2.9979e10 / (dp->units_l / dp->units_t), 1.428571);
				rperi = 2.6789966 * madhoc * pow(mass_k * mass_kp, 0.28571428) * pow(mass_k + mass_kp, 0.42857143) / pow(W, 0.57142857) / clight10o7;
				const double S_brem = ORC_PI * orc_sqr(rperi) * (1.0 + 2.0 * madhoc * (mass_k + mass_kp) / (rperi * orc_sqr(W)));
				S_tmp = S_tmp > S_brem ? S_tmp : S_brem;
			}
			if (dp->tc_factor > 1 && (kk <= 1 || kk == 7 || kk >= 10) && (kkp <= 1 || kkp == 7 || kkp >= 10)) { /* :393-402 */
				rperi = dp->tc_factor * (rk + rkp);
				const double S_tc_simple = ORC_PI * orc_sqr(rperi) * (1.0 + 2.0 * madhoc * (mass_k + mass_kp) / (rperi * orc_sqr(W)));
				S_tmp = S_tmp > S_tc_simple ? S_tmp : S_tc_simple;
			}
			rperi = collisions_multiple * (rk + rkp); /* :405-406 */
			S_coll = ORC_PI * orc_sqr(rperi) * (1.0 + 2.0 * madhoc * (mass_k + mass_kp) / (rperi * orc_sqr(W)));
			S = S_coll > S_tmp ? S_coll : S_tmp; /* :409-410 */
			rperi = madhoc * (mass_k + mass_kp) / orc_sqr(W) * (-1.0 + sqrt(1.0 + S / ORC_FB_PI * orc_sqr(W * W / (madhoc * mass_k + madhoc * mass_kp))));
		} else S = 0.0;
		P_enc = n_local_dens * W * S * (dt * N_STAR / log(dp->gamma * N_STAR)); /* :436 */
		pair_k[npairs] = k; pair_kp[npairs] = kp;
		pair_rperi[npairs] = rperi; pair_W[npairs] = W; pair_rcm[npairs] = rcm;
		if (draws[3 * npairs + 1] < P_enc) { /* :444-467 */
			pair_status[npairs] = (bk > 0 || bkp > 0) ? 2 : 1;
		} else if (dp->relaxation) { /* :468-540 */
			double w1[4], w2[4], w_new[4], v_new[4], vp_new[4], wp, beta, psi;
			const double Trel12 = (ORC_PI / 32.0) * orc_cub(W) / (N_STAR * n_local_dens * orc_sqr((mass_k + mass_kp) * madhoc));
			beta = (ORC_PI / 2.0) * sqrt(dt / Trel12);
			stats->Nrel++;
			for (int i = 0; i < 4; i++) {
				if (beta > relbeta[i]) {
					stats->Nrelbeta[i]++;
					stats->qaverelbeta[i] += (mass_k > mass_kp ? mass_k : mass_kp) / (mass_k < mass_kp ? mass_k : mass_kp);
					stats->maverelbeta[i] += (mass_k + mass_kp) / 2.0 * dp->units_mstar / ORC_MSUN;
					stats->raverelbeta[i] += rcm;
				}
			}
			if (beta > ORC_PI / 2.0) beta = ORC_PI / 2.0;
			wp = sqrt(orc_sqr(w[1]) + orc_sqr(w[2]));
			w1[1] = -w[2] * W / wp; w1[2] = w[1] * W / wp; w1[3] = 0.0;
			w2[1] = -w[1] * w[3] / wp; w2[2] = -w[2] * w[3] / wp; w2[3] = wp;
			psi = draws[3 * npairs + 2] * 2 * ORC_PI;
			for (j = 1; j <= 3; j++) w_new[j] = w[j] * cos(beta) + w1[j] * sin(beta) * cos(psi) + w2[j] * sin(beta) * sin(psi);
			for (j = 1; j <= 3; j++) {
				v_new[j] = v[j] - mass_kp / (mass_k + mass_kp) * (w_new[j] - w[j]);
				vp_new[j] = vp[j] + mass_k / (mass_k + mass_kp) * (w_new[j] - w[j]);
			}
			vr[k - 1] = v_new[3];
			vt[k - 1] = sqrt(orc_sqr(v_new[1]) + orc_sqr(v_new[2]));
			vr[kp - 1] = vp_new[3];
			vt[kp - 1] = sqrt(orc_sqr(vp_new[1]) + orc_sqr(vp_new[2]));
			/* set_star_EJ, cmc_dynamics_helper.c:878-882 */
			E[k - 1] = t->phi[g_k] + 0.5 * (orc_sqr(vr[k - 1]) + orc_sqr(vt[k - 1]));
			J[k - 1] = t->r[g_k] * vt[k - 1];
			E[kp - 1] = t->phi[g_kp] + 0.5 * (orc_sqr(vr[kp - 1]) + orc_sqr(vt[kp - 1]));
			J[kp - 1] = t->r[g_kp] * vt[kp - 1];
			pair_status[npairs] = 0;
		} else pair_status[npairs] = 3;
		npairs++;
	}
	return npairs;
}

/* ---- simul_relax_new, src/cmc/cmc_dynamics_helper.c:2526-2580 ------------------------------------------------------------------------------- */

double orc_simul_relax_new(const orc_tables_t *t, long n_local, long first_global, long p, double thetasemax, const double *vr, const double *vt)
{
	double dtmin = INFINITY;
	for (long si = 1 + p; si < n_local - p; si += 2 * p) { /* :2539 */
		const long simin = si - p, simax = simin + (2 * p - 1);
		double Mv2ave = 0.0, Mave = 0.0, M2ave = 0.0;
		for (long k = simin; k <= simax; k++) {
			const double tmp = t->m[first_global + k - 1] * t->madhoc;
			Mv2ave += tmp * (orc_sqr(vr[k - 1]) + orc_sqr(vt[k - 1]));
			Mave += tmp;
			M2ave += orc_sqr(tmp);
		}
		Mv2ave /= (double) (2 * p);
		Mave /= (double) (2 * p);
		M2ave /= (double) (2 * p);
		const double sigma = sqrt(Mv2ave / Mave);
		const double W = 4.0 * sigma / sqrt(3.0 * ORC_PI);
		const double n_local_dens = orc_calc_n_local(t->r, first_global + si - 1, p, t->N_MAX);
		const double dt = orc_sqr(2.0 * thetasemax / ORC_PI) * (ORC_PI / 32.0) * orc_cub(W) / (((double) t->N_STAR) * n_local_dens * (4.0 * M2ave));
		dtmin = dt < dtmin ? dt : dtmin; /* MIN(dtmin, dt), include/cmc/cmc.h */
	}
	return dtmin;
}

/* ---- tidally_strip_stars, src/cmc/cmc_evolution_thr.c:443-645: the per-star criterion of one pass of its loop ------------------------------ */

long orc_tidally_strip_select(long n_local, int tidal_treatment, double threshold, const double *r_apo, const double *E, double *rnew,
			      double *vrnew, double *vtnew, long *idx_out)
{
	long count = 0;
	for (long i = 0; i < n_local; i++) {
		/* :479 radial cut-off (r_apo > Rtidal), :555 Giersz (E > gierszalpha * phi_rtidal); both only for stars still in the cluster */
		const int strip = tidal_treatment == 0 ? (r_apo[i] > threshold && rnew[i] < 1000000) : (tidal_treatment == 1 ? (E[i] > threshold && rnew[i] < 1000000) : 0);
		if (!strip) continue;
		rnew[i] = ORC_SF_INFINITY; /* :481-483 / :557-559 */
		vrnew[i] = 0.0;
		vtnew[i] = 0.0;
		idx_out[count++] = i + 1;
	}
	return count;
}

/* ---- ComputeEnergy's second loop, src/cmc/cmc_utils.c:619-636: the rank's energy sums ----------------------------------------------------- */

void orc_compute_energy_totals(const orc_tables_t *t, long n_local, long first_global, const double *vr, const double *vt, const double *Eint,
			       const long *binind, const int *bin_inuse, const double *bin_m1, const double *bin_m2, const double *bin_a,
			       const double *bin_Eint1, const double *bin_Eint2, double out[4])
{
	const double phi0 = t->phi[0];
	double K = 0.0, P = 0.0, Ei = 0.0, Eb = 0.0;
	for (long i = 1; i <= n_local; i++) {
		const long j = first_global + i - 1;
		K += 0.5 * (vr[i - 1] * vr[i - 1] + vt[i - 1] * vt[i - 1]) * t->m[j] * t->madhoc;
		P += t->phi[j] * t->m[j] * t->madhoc;
		P += phi0 * t->cenma_m * t->madhoc / t->N_MAX;
		const long b = binind[i - 1];
		if (b == 0) Ei += Eint[i - 1];
		else if (bin_inuse[b]) {
			Eb += -(bin_m1[b] * t->madhoc) * (bin_m2[b] * t->madhoc) / (2.0 * bin_a[b]);
			Ei += bin_Eint1[b] + bin_Eint2[b];
		}
	}
	out[0] = K; out[1] = P; out[2] = Ei; out[3] = Eb; /* the caller: P *= 0.5, tot = K + P + Eint + Eb, Allreduce, + cenma.E + escaped (:638-655) */
}

/* ---- comp_multi_mass_percent, src/cmc/cmc_utils.c:769-861 (+ find_stars_mass_bin :755-764, m_binsearch include/common/gensearch.h) ------- */

static long orc_mass_bin(const double *mass_bins, long n_bins, double smass)
{
	if (smass < mass_bins[0] || smass > mass_bins[n_bins - 1]) return -1;
	long lo = 0, hi = n_bins - 1; /* largest index with mass_bins[index] < smass (ascending bins) */
	do {
		const long k = (lo + hi + 1) / 2;
		if (mass_bins[k] < smass) lo = k; else hi = k - 1;
	} while (hi != lo);
	return lo;
}

void orc_comp_multi_mass_percent(const double *star_r, const double *star_m, long N_MAX, const double *mass_bins, long n_bins, double solar_mass_dyn,
				 const double *mass_pc, long count, double *multi_mass_r /* [n_bins][count] */)
{
	if (n_bins <= 1) return;
	for (long b = 0; b < n_bins; b++) {
		/* :787-793 the bin's total mass, star count and last radius, summed in star order */
		double mtotal = 0.0, r_last = 0.0;
		long number = 0;
		for (long i = 1; i <= N_MAX; i++)
			if (orc_mass_bin(mass_bins, n_bins, star_m[i] / solar_mass_dyn) == b) { mtotal += star_m[i]; number++; r_last = star_r[i]; }
		if (number == 0) continue;
		if (number == 1) { for (long c = 0; c < count; c++) multi_mass_r[b * count + c] = r_last; continue; } /* :836-839 */
		/* :808-818 (rs, percents) with the (0, 0) point, then gsl_interp_linear of rs over percents at mass_pc (:840-844) */
		for (long c = 0; c < count; c++) {
			const double x = mass_pc[c];
			double mcount = 0.0, x_lo = 0.0, y_lo = 0.0, val = 0.0;
			int found = 0;
			long seen = 0;
			for (long i = 1; i <= N_MAX && !found; i++) {
				if (orc_mass_bin(mass_bins, n_bins, star_m[i] / solar_mass_dyn) != b) continue;
				mcount += star_m[i];
				seen++;
				const double x_hi = mcount / mtotal, y_hi = star_r[i];
				/* gsl_interp_bsearch: the interval [x_lo, x_hi) holding x, the last one closed */
				if (x < x_hi || seen == number) {
					const double dx = x_hi - x_lo;
					val = dx > 0.0 ? y_lo + (x - x_lo) / dx * (y_hi - y_lo) : 0.0;
					found = 1;
				}
				x_lo = x_hi; y_lo = y_hi;
			}
			multi_mass_r[b * count + c] = (x < 0.0 || x > 1.0) ? NAN : val;
		}
	}
}

/* clusdyn_calculate, src/cmc/cmc_utils.c:1613-1624 */
double orc_clusdyn_calculate(const double *star_r, const double *star_m, long N_STAR, double Mtotal)
{
	double m = 0.0;
	long k = 1;
	while (m < 0.5 * Mtotal) {
		m += star_m[k] / N_STAR;
		k++;
	}
	return star_r[k];
}

/* TEST INFRASTRUCTURE ONLY (oracle/): the seven GSL root-solver entry points the reference's
 * find_root_vr() reaches (reference src/cmc/cmc_nr.c:524-587), plus fb_sqr and close_buffers.
 * GSL is an un-vendored dependency of the reference (version unpinned: find_package(GSL REQUIRED),
 * reference CMakeLists.txt:62).  This is a restatement of the published Brent-Dekker method
 * (R. P. Brent, "Algorithms for Minimization without Derivatives", 1973, ch. 4) with the
 * bracket bookkeeping and the interval / delta stopping tests as documented in the GSL manual
 * ("One dimensional Root-Finding": |a-b| < epsabs + epsrel*min(|a|,|b|)).
 * PARITY UNPINNED for this path: no reference test reaches it deterministically. */
#include <math.h>
#include <stdlib.h>
#include "gsl/gsl_errno.h"
#include "gsl/gsl_roots.h"

typedef struct { double a, b, c, d, e, fa, fb, fc; } brent_state_t;

static const gsl_root_fsolver_type brent_type = { "brent", sizeof(brent_state_t) };
const gsl_root_fsolver_type *gsl_root_fsolver_brent = &brent_type;
const gsl_root_fsolver_type *gsl_root_fsolver_bisection = &brent_type;

gsl_root_fsolver *gsl_root_fsolver_alloc(const gsl_root_fsolver_type *T)
{
	gsl_root_fsolver *s = calloc(1, sizeof(*s));
	s->type = T;
	s->state = calloc(1, T->size);
	return s;
}

void gsl_root_fsolver_free(gsl_root_fsolver *s) { if (s) { free(s->state); free(s); } }

int gsl_root_fsolver_set(gsl_root_fsolver *s, gsl_function *f, double x_lower, double x_upper)
{
	brent_state_t *st = s->state;
	if (x_lower > x_upper) return GSL_EINVAL;
	s->function = f;
	s->x_lower = x_lower;
	s->x_upper = x_upper;
	s->root = 0.5 * (x_lower + x_upper);
	double fl = GSL_FN_EVAL(f, x_lower), fu = GSL_FN_EVAL(f, x_upper);
	if (!isfinite(fl) || !isfinite(fu)) return GSL_EBADFUNC;
	st->a = x_lower; st->fa = fl;
	st->b = x_upper; st->fb = fu;
	st->c = x_upper; st->fc = fu;
	st->d = x_upper - x_lower;
	st->e = x_upper - x_lower;
	if ((fl < 0.0 && fu < 0.0) || (fl > 0.0 && fu > 0.0)) return GSL_EINVAL;
	return GSL_SUCCESS;
}

int gsl_root_fsolver_iterate(gsl_root_fsolver *s)
{
	brent_state_t *st = s->state;
	double a = st->a, b = st->b, c = st->c, d = st->d, e = st->e, fa = st->fa, fb = st->fb, fc = st->fc;
	int ac_equal = 0;
	double tol, m;

	if ((fb < 0 && fc < 0) || (fb > 0 && fc > 0)) { ac_equal = 1; c = a; fc = fa; d = b - a; e = b - a; }
	if (fabs(fc) < fabs(fb)) { ac_equal = 1; a = b; b = c; c = a; fa = fb; fb = fc; fc = fa; }
	tol = 0.5 * GSL_DBL_EPSILON * fabs(b);
	m = 0.5 * (c - b);
	if (fb == 0) {
		s->root = b; s->x_lower = b; s->x_upper = b;
		goto store;
	}
	if (fabs(m) <= tol) {
		s->root = b;
		if (b < c) { s->x_lower = b; s->x_upper = c; } else { s->x_lower = c; s->x_upper = b; }
		goto store;
	}
	if (fabs(e) < tol || fabs(fa) <= fabs(fb)) {
		d = m; e = m;                              /* bisection */
	} else {
		double p, q, r, sfrac = fb / fa;           /* inverse quadratic / secant */
		if (ac_equal) { p = 2 * m * sfrac; q = 1 - sfrac; }
		else { q = fa / fc; r = fb / fc; p = sfrac * (2 * m * q * (q - r) - (b - a) * (r - 1)); q = (q - 1) * (r - 1) * (sfrac - 1); }
		if (p > 0) q = -q; else p = -p;
		if (2 * p < GSL_MIN(3 * m * q - fabs(tol * q), fabs(e * q))) { e = d; d = p / q; }
		else { d = m; e = m; }
	}
	a = b; fa = fb;
	if (fabs(d) > tol) b += d; else b += (m > 0 ? +tol : -tol);
	fb = GSL_FN_EVAL(s->function, b);
	if (!isfinite(fb)) return GSL_EBADFUNC;
	s->root = b;
	if ((fb < 0 && fc < 0) || (fb > 0 && fc > 0)) c = a;
	if (b < c) { s->x_lower = b; s->x_upper = c; } else { s->x_lower = c; s->x_upper = b; }
store:
	st->a = a; st->b = b; st->c = c; st->d = d; st->e = e; st->fa = fa; st->fb = fb; st->fc = fc;
	return GSL_SUCCESS;
}

double gsl_root_fsolver_root(const gsl_root_fsolver *s) { return s->root; }
double gsl_root_fsolver_x_lower(const gsl_root_fsolver *s) { return s->x_lower; }
double gsl_root_fsolver_x_upper(const gsl_root_fsolver *s) { return s->x_upper; }

int gsl_root_test_interval(double x_lower, double x_upper, double epsabs, double epsrel)
{
	double al = fabs(x_lower), au = fabs(x_upper), min_abs;
	if ((x_lower > 0.0 && x_upper > 0.0) || (x_lower < 0.0 && x_upper < 0.0)) min_abs = GSL_MIN(al, au);
	else min_abs = 0;
	return fabs(x_upper - x_lower) < epsabs + epsrel * min_abs ? GSL_SUCCESS : GSL_CONTINUE;
}

int gsl_root_test_delta(double x1, double x0, double epsabs, double epsrel)
{
	double tolerance = epsabs + epsrel * fabs(x1);
	return (fabs(x1 - x0) < tolerance || x1 == x0) ? GSL_SUCCESS : GSL_CONTINUE;
}

/* reference src/fewbody-0.24/fewbody_utils.c:66 (x*x); only used for MINIMUM_R bookkeeping */
double fb_sqr(double x) { return x * x; }
/* reference cmc_io.c close_buffers(): flushes log files on abnormal exit; nothing is open here */
void close_buffers(void) {}

/* ---- gsl_spline with gsl_interp_linear, as comp_multi_mass_percent uses it (reference src/cmc/cmc_utils.c:820-848) -------------------------
 * GSL is un-vendored and its version unpinned (CMakeLists.txt:62): this restates the published algorithm of interpolation/linear.c
 * (linear_eval) and interpolation/bsearch.c (gsl_interp_bsearch: the index i with x[i] <= x < x[i+1], the last interval closed) -- "parity
 * unpinned" for whatever goes through it. */
#include <stdlib.h>
#include <string.h>
#include "gsl/gsl_spline.h"
static const gsl_interp_type shim_linear = {"linear"}, shim_cspline = {"cspline"};
const gsl_interp_type *gsl_interp_linear = &shim_linear;
const gsl_interp_type *gsl_interp_cspline = &shim_cspline;
gsl_interp_accel *gsl_interp_accel_alloc(void) { return calloc(1, sizeof(gsl_interp_accel)); }
void gsl_interp_accel_free(gsl_interp_accel *a) { free(a); }
gsl_spline *gsl_spline_alloc(const gsl_interp_type *T, size_t size)
{
	gsl_spline *s = calloc(1, sizeof(*s));
	s->type = T; s->size = size;
	s->x = malloc(size * sizeof(double)); s->y = malloc(size * sizeof(double));
	return s;
}
int gsl_spline_init(gsl_spline *s, const double *x, const double *y, size_t size)
{
	memcpy(s->x, x, size * sizeof(double));
	memcpy(s->y, y, size * sizeof(double));
	return GSL_SUCCESS;
}
double gsl_spline_eval(const gsl_spline *s, double x, gsl_interp_accel *a)
{
	size_t ilo = 0, ihi = s->size - 1;
	(void) a; /* the accelerator only caches the index the search below returns */
	if (x < s->x[0] || x > s->x[s->size - 1]) return NAN; /* GSL_EDOM -> GSL_NAN */
	while (ihi > ilo + 1) {
		size_t i = (ihi + ilo) / 2;
		if (s->x[i] > x) ihi = i; else ilo = i;
	}
	{
		const double x_lo = s->x[ilo], x_hi = s->x[ilo + 1], y_lo = s->y[ilo], y_hi = s->y[ilo + 1], dx = x_hi - x_lo;
		if (dx > 0.0) return y_lo + (x - x_lo) / dx * (y_hi - y_lo);
		return 0.0; /* GSL_EINVAL */
	}
}
void gsl_spline_free(gsl_spline *s) { free(s->x); free(s->y); free(s); }

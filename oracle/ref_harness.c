/* TEST INFRASTRUCTURE ONLY (oracle/): drives the UNMODIFIED reference implementation of the
 * hot path (compiled from /root/reference by oracle/Makefile into oracle/_ref/) so that the
 * CPU restatement in oracle/cmc_oracle.c and the CUDA path can be checked against the real thing.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may use it.
 *
 * The reference keeps all state in process globals (include/cmc/cmc_vars.h); this file allocates
 * and fills them the way the reference's own start-up does:
 *   star/binary calloc           cmc_io.c:2099-2101
 *   star_r/star_m/star_phi       cmc_io.c:3431-3437
 *   sentinels star_r[0]=ZERO     cmc_fits.c:128-136
 *   madhoc = 1/N_STAR            cmc_utils.c:1183
 *   rng                          cmc_utils.c:2364-2375 (set_rng_states)
 *   q_root (Brent solver)        cmc_io.c:3398
 * and then calls the reference's own entry points
 *   new_orbits_calculate (cmc_utils.c:2194), qsorts_new (:2232), post_sort_comm (:2259),
 *   calc_potential_new (:2006).
 * Every uniform draw the orbit step consumes is recorded (taus113's rng_t113_dbl_new is
 * renamed at compile time with -D and wrapped here) so the same draws can be fed to the
 * oracle port and to the CUDA kernel. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include "cmc.h"
#include "cmc_vars.h"

#define EXPORT __attribute__((visibility("default")))

/* the real generator, renamed by -Drng_t113_dbl_new=ref_real_rng_t113_dbl_new when
 * taus113-v2.c is compiled; every other TU calls the wrapper below */
double ref_real_rng_t113_dbl_new(struct rng_t113_state *state);

static double *draw_log = NULL;
static long draw_cap = 0, draw_cnt = 0;
static int draw_replay = 0;      /* 1: serve draws from draw_log instead of the generator */
static long draw_pos = 0;

double rng_t113_dbl_new(struct rng_t113_state *state)
{
	if (draw_replay) {
		if (draw_pos >= draw_cnt) { fprintf(stderr, "ref_harness: replay ran out of draws\n"); exit(7); }
		return draw_log[draw_pos++];
	}
	double u = ref_real_rng_t113_dbl_new(state);
	if (draw_log) {
		if (draw_cnt == draw_cap) {
			draw_cap = draw_cap ? 2 * draw_cap : (1 << 20);
			draw_log = realloc(draw_log, draw_cap * sizeof(double));
		}
		draw_log[draw_cnt++] = u;
	}
	return u;
}

EXPORT long ref_sizeof_star(void) { return (long) sizeof(star_t); }
EXPORT long ref_sizeof_binary(void) { return (long) sizeof(binary_t); }
EXPORT long ref_sizeof_orbit_rs(void) { return (long) sizeof(orbit_rs_t); }

/* offsets of the star_t / binary_t fields the path touches (include/cmc/cmc.h:409-601, :231-390) */
EXPORT long ref_star_offset(const char *f)
{
#define O(n) if (!strcmp(f, #n)) return (long) offsetof(star_t, n)
	O(r); O(vr); O(vt); O(m); O(E); O(J); O(EI); O(Eint); O(rnew); O(vrnew); O(vtnew); O(rOld); O(X); O(Y);
	O(r_peri); O(r_apo); O(phi); O(interacted); O(binind); O(id); O(rad); O(Uoldrold); O(Uoldrnew); O(vtold); O(vrold);
	O(se_k); O(threebb_interacted);
#undef O
	return -1;
}
EXPORT long ref_binary_offset(const char *f)
{
#define O(n) if (!strcmp(f, #n)) return (long) offsetof(binary_t, n)
	O(id1); O(id2); O(rad1); O(rad2); O(m1); O(m2); O(Eint1); O(Eint2); O(a); O(e); O(inuse);
#undef O
	return -1;
}

EXPORT int ref_init(long n_star, long n_star_dim, long n_table_dim, long n_bin_dim, long min_chunk, int samplesize, long idum, double orbit_r_in)
{
	MPI_Comm_size(MPI_COMM_WORLD, &procs);
	MPI_Comm_rank(MPI_COMM_WORLD, &myid);
	debug = 0; quiet = 1;
	N_STAR_DIM_OPT = n_star_dim;   /* star[]: this rank's share (cmc_io.c:2092-2099) */
	N_STAR_DIM = n_table_dim;      /* star_r / star_m / star_phi: the whole cluster (cmc_io.c:3431-3437) */
	N_BIN_DIM = N_BIN_DIM_OPT = n_bin_dim;
	star = calloc(n_star_dim, sizeof(star_t));
	binary = calloc(n_bin_dim, sizeof(binary_t));
	/* the replicated tables are sized for the whole cluster, the star array for this rank's share (cmc_io.c:2092-2099, 3431-3437) */
	star_r = calloc(n_table_dim, sizeof(double));
	star_m = calloc(n_table_dim, sizeof(double));
	star_phi = calloc(n_table_dim, sizeof(double));
	if (!star || !binary || !star_r || !star_m || !star_phi) return -1;
	Start = calloc(procs, sizeof(int)); End = calloc(procs, sizeof(int));
	mpiDisp = calloc(procs, sizeof(int)); mpiLen = calloc(procs, sizeof(int));
	clus.N_STAR = clus.N_STAR_NEW = clus.N_MAX = n_star;
	clus.N_BINARY = 0;
	madhoc = 1.0 / (double) clus.N_STAR;
	MIN_CHUNK_SIZE = min_chunk;
	SAMPLESIZE = samplesize;
	SEARCH_GRID = 0;
	IDUM = idum;
	orbit_r = orbit_r_in;
	APSIDES_PRECISION = 1.0e-11; APSIDES_MAX_ITER = 100; APSIDES_CONVERGENCE = 5.e-13;   /* cmc_io.c:1752-1754 */
	q_root = gsl_root_fsolver_alloc(gsl_root_fsolver_brent);
	units.t = units.m = units.l = units.E = units.mstar = 1.0;
	cenma.m = cenma.m_new = cenma.E = cenma.E_new = 0.0;
	/* a fresh run: the accumulators are process globals and a test process initialises the library more than once */
	Eescaped = Jescaped = Eintescaped = Ebescaped = TidalMassLoss = Etidal = 0.0;
	Eescaped_old = Jescaped_old = Eintescaped_old = Ebescaped_old = TidalMassLoss_old = Etidal_old = 0.0;
	OldTidalMassLoss = DTidalMassLoss = 0.0;
	Eoops = 0.0; max_r = 0.0;
	tcount = 1; TotalTime = 0.0;
	mpi_escfile_buf = malloc(STR_BUF_LEN);
	mpi_escfile_wrbuf = malloc(1 << 26);
	mpi_escfile_wrbuf[0] = 0; mpi_escfile_len = 0;
	set_rng_states();
	findLimits(clus.N_MAX, MIN_CHUNK_SIZE);
	mpiFindIndicesCustom(clus.N_MAX, MIN_CHUNK_SIZE, myid, &mpiBegin, &mpiEnd);
	clus.N_MAX_NEW = mpiEnd - mpiBegin + 1;
	star_r[0] = ZERO; star_m[0] = 0.0;
	star[0].r = ZERO; star[0].m = 0.0;
	return 0;
}

EXPORT void ref_free(void)
{
	free(star); free(binary); free(star_r); free(star_m); free(star_phi);
	free(Start); free(End); free(mpiDisp); free(mpiLen); free(curr_st);
	free(mpi_escfile_buf); free(mpi_escfile_wrbuf); free(draw_log);
	gsl_root_fsolver_free(q_root);
	star = NULL; binary = NULL; star_r = star_m = star_phi = NULL; draw_log = NULL; draw_cap = draw_cnt = 0;
}

EXPORT void *ref_star(void) { return star; }
EXPORT void *ref_binary(void) { return binary; }
EXPORT double *ref_star_r(void) { return star_r; }
EXPORT double *ref_star_m(void) { return star_m; }
EXPORT double *ref_star_phi(void) { return star_phi; }
EXPORT int *ref_Start(void) { return Start; }
EXPORT int *ref_End(void) { return End; }
EXPORT int ref_procs(void) { return procs; }
EXPORT int ref_myid(void) { return myid; }

EXPORT void ref_set_scalars(long N_MAX, long N_MAX_NEW, long N_BINARY, double cenma_m, double cenma_m_new, double Rtidal_in, double Mtotal_in)
{
	clus.N_MAX = N_MAX; clus.N_MAX_NEW = N_MAX_NEW; clus.N_BINARY = N_BINARY; N_b_local = N_BINARY;
	cenma.m = cenma_m; cenma.m_new = cenma_m_new; Rtidal = Rtidal_in; Mtotal = Mtotal_in;
}

/* out[0..15]: N_MAX, N_MAX_NEW, Mtotal, Rtidal, cenma.m, max_r, TidalMassLoss, Eescaped, Jescaped,
 * Eintescaped, Ebescaped, Etidal, mpiBegin, mpiEnd, N_b_local, escfile bytes */
EXPORT void ref_get_scalars(double *out)
{
	out[0] = clus.N_MAX; out[1] = clus.N_MAX_NEW; out[2] = Mtotal; out[3] = Rtidal; out[4] = cenma.m;
	out[5] = max_r; out[6] = TidalMassLoss; out[7] = Eescaped; out[8] = Jescaped; out[9] = Eintescaped;
	out[10] = Ebescaped; out[11] = Etidal; out[12] = mpiBegin; out[13] = mpiEnd; out[14] = N_b_local;
	out[15] = (double) mpi_escfile_len;
}

EXPORT void ref_find_limits(long N, int blk) { findLimits(N, blk); }
EXPORT void ref_find_indices_custom(long N, int blk, int i, int *b, int *e) { mpiFindIndicesCustom(N, blk, i, b, e); }

EXPORT void ref_calc_potential_new(void) { calc_potential_new(); }
EXPORT void ref_qsorts_new(void) { qsorts_new(); }
EXPORT void ref_post_sort_comm(void) { post_sort_comm(); }

/* draw recording: mode 0 = off, 1 = record from taus113, 2 = replay the recorded log */
EXPORT void ref_draw_mode(int mode)
{
	if (mode == 1) {
		if (!draw_log) { draw_cap = 1 << 20; draw_log = malloc(draw_cap * sizeof(double)); }
		draw_cnt = 0; draw_replay = 0;
	} else if (mode == 2) { draw_replay = 1; draw_pos = 0; }
	else { draw_replay = 0; }
}
EXPORT long ref_draw_count(void) { return draw_cnt; }
EXPORT const double *ref_draw_log(void) { return draw_log; }
EXPORT void ref_set_draws(const double *d, long n)
{
	if (n > draw_cap) { draw_cap = n; draw_log = realloc(draw_log, n * sizeof(double)); }
	memcpy(draw_log, d, n * sizeof(double)); draw_cnt = n;
}

EXPORT void ref_new_orbits_calculate(void)
{
	mpi_escfile_wrbuf[0] = 0; mpi_escfile_len = 0;
	new_orbits_calculate();
}

/* single-orbit probes for unit-level parity */
EXPORT double ref_potential(double r) { return potential(r); }
EXPORT long ref_find_zero_Q(long j, long kmin, long kmax, double E, double J) { return find_zero_Q(j, kmin, kmax, E, J); }
EXPORT double ref_function_q(long j, double r, double pot, double E, double J) { return function_q(j, r, pot, E, J); }
EXPORT void ref_calc_orbit_new(long index, double E, double J, double *out)
{
	orbit_rs_t o = calc_orbit_new(index, E, J);
	out[0] = o.rp; out[1] = o.ra; out[2] = o.dQdrp; out[3] = o.dQdra; out[4] = o.kmax; out[5] = o.kmin; out[6] = o.circular_flag;
}
EXPORT double ref_find_root_vr(long index, long k, double E, double J) { return find_root_vr(index, k, E, J); }

/* taus113 known-answer access (src/libs/taus113-v2.c:166-216, 253-270) */
EXPORT void ref_rng_reset(unsigned long seed, unsigned int *z) { struct rng_t113_state s; int i; reset_rng_t113_new(seed, &s); for (i = 0; i < 4; i++) z[i] = (unsigned int) s.z[i]; }
EXPORT void ref_rng_jump80(unsigned int *z) { struct rng_t113_state s; int i; for (i = 0; i < 4; i++) s.z[i] = z[i]; s = rng_t113_jump(s, JPoly_2_80); for (i = 0; i < 4; i++) z[i] = (unsigned int) s.z[i]; }
EXPORT void ref_rng_jump20(unsigned int *z) { struct rng_t113_state s; int i; for (i = 0; i < 4; i++) s.z[i] = z[i]; s = rng_t113_jump(s, JPoly_2_20); for (i = 0; i < 4; i++) z[i] = (unsigned int) s.z[i]; }
EXPORT double ref_rng_next(unsigned int *z) { struct rng_t113_state s; int i; double u; for (i = 0; i < 4; i++) s.z[i] = z[i]; u = ref_real_rng_t113_dbl_new(&s); for (i = 0; i < 4; i++) z[i] = (unsigned int) s.z[i]; return u; }

/* multi-rank runs: fork nprocs-1 copies of the calling process (mini_mpi.c).  Every copy returns from this call with
 * its own rank; call BEFORE ref_init.  ref_join(): children exit, rank 0 waits for them. */
EXPORT int ref_spawn(int nprocs, size_t window_bytes) { return mini_mpi_spawn(nprocs, window_bytes); }
EXPORT int ref_join(void) { return mini_mpi_join(); }
EXPORT int ref_rank(void) { int r; MPI_Comm_rank(MPI_COMM_WORLD, &r); return r; }
EXPORT void ref_barrier(void) { MPI_Barrier(MPI_COMM_WORLD); }
/* the reference's own ComputeEnergy (cmc_utils.c:604-656): E, J of the local stars from the new potential */
EXPORT void ref_compute_energy(void) { ComputeEnergy(); }
EXPORT double ref_clusdyn_rh(void) { clusdyn_calculate(); return clusdyn.rh; }
EXPORT void ref_energy_conservation2(void) { energy_conservation2(); }
EXPORT void ref_compute_intermediate_energy(void) { ComputeIntermediateEnergy(); }

/* energy_conservation1 (cmc_utils.c:2153-2174) and energy_conservation3 -> set_velocities3 (:254-363) on the harness's cluster state */
EXPORT void ref_energy_conservation1(void) { energy_conservation1(); }
EXPORT void ref_energy_conservation3(void) { energy_conservation3(); }
EXPORT double ref_get_Eoops(void) { return Eoops; }
EXPORT void ref_set_Eoops(double v) { Eoops = v; }

/* calc_sigma_r (src/cmc/cmc_dynamics_helper.c:2821-2975) on the harness's cluster state: out_a / out_sigma [n_limit + 1], entry 0 unused */
EXPORT void ref_calc_sigma_r(long p, long n_limit, int r_0_mave_1, double *out_a, double *out_sigma)
{
	long n = 0;
	calc_sigma_r(p, n_limit, out_a, out_sigma, &n, r_0_mave_1);
}

/* tidally_strip_stars (src/cmc/cmc_evolution_thr.c:443-645) on the harness's cluster state; out[0..2] = Rtidal, TidalMassLoss, DTidalMassLoss */
EXPORT void ref_tidally_strip_stars(long treatment, double gamma, double *out)
{
	TIDAL_TREATMENT = treatment;
	GAMMA = gamma;
	if (!mpi_logfile_buf) { mpi_logfile_buf = malloc(STR_BUF_LEN); mpi_logfile_wrbuf = malloc(1 << 20); }
	mpi_logfile_wrbuf[0] = 0; mpi_logfile_len = 0;
	mpi_escfile_wrbuf[0] = 0; mpi_escfile_len = 0;
	tidally_strip_stars();
	out[0] = Rtidal; out[1] = TidalMassLoss; out[2] = DTidalMassLoss;
}

/* Etotal after ComputeEnergy (src/cmc/cmc_utils.c:604-656): tot, K, P, Eint, Eb */
EXPORT void ref_get_etotal(double *out) { out[0] = Etotal.tot; out[1] = Etotal.K; out[2] = Etotal.P; out[3] = Etotal.Eint; out[4] = Etotal.Eb; }

/* comp_multi_mass_percent (src/cmc/cmc_utils.c:769-861) on the harness's tables: out[n_bins][count] (entries of empty bins stay -1) */
EXPORT void ref_comp_multi_mass_percent(const double *bins, long n_bins, double solar_mass_dyn, const double *pc, long count, double *out)
{
	long i, j;
	NO_MASS_BINS = n_bins; SOLAR_MASS_DYN = solar_mass_dyn; MASS_PC_COUNT = count;
	mass_bins = (double *) malloc(n_bins * sizeof(double));
	mass_pc = (double *) malloc(count * sizeof(double));
	memcpy(mass_bins, bins, n_bins * sizeof(double));
	memcpy(mass_pc, pc, count * sizeof(double));
	multi_mass_r = (double **) malloc(n_bins * sizeof(double *));
	for (i = 0; i < n_bins; i++) { multi_mass_r[i] = (double *) malloc(count * sizeof(double)); for (j = 0; j < count; j++) multi_mass_r[i][j] = -1.0; }
	comp_multi_mass_percent();
	for (i = 0; i < n_bins; i++) { memcpy(out + i * count, multi_mass_r[i], count * sizeof(double)); free(multi_mass_r[i]); }
	free(multi_mass_r); free(mass_bins); free(mass_pc);
	multi_mass_r = NULL; mass_bins = NULL; mass_pc = NULL;
}

/* comp_mass_percent (src/cmc/cmc_utils.c:872-985) on the harness's cluster state: thresholds in, the eight result arrays out
 * (out[8][count], in the order mass_r, ave_mass_r, no_star_r, densities_r, ke_rad_r, ke_tan_r, v2_rad_r, v2_tan_r; on ranks other than 0
 * the last four hold what that rank computed, i.e. nothing meaningful -- the reference prints from rank 0 only). */
EXPORT void ref_comp_mass_percent(const double *pc, long count, int bh_include, double minimum_r, double initial_total_mass_in, double *out)
{
	long i;
	double **arr[8] = {&mass_r, &ave_mass_r, &no_star_r, &densities_r, &ke_rad_r, &ke_tan_r, &v2_rad_r, &v2_tan_r};
	MASS_PC_COUNT = count;
	MASS_PC_BH_INCLUDE = bh_include;
	MINIMUM_R = minimum_r;
	initial_total_mass = initial_total_mass_in;
	mass_pc = (double *) malloc(count * sizeof(double));
	for (i = 0; i < count; i++) mass_pc[i] = pc[i];
	for (i = 0; i < 8; i++) { long j; *arr[i] = (double *) malloc(count * sizeof(double)); for (j = 0; j < count; j++) (*arr[i])[j] = -1.0; }
	comp_mass_percent();
	for (i = 0; i < 8; i++) { memcpy(out + i * count, *arr[i], count * sizeof(double)); free(*arr[i]); *arr[i] = NULL; }
	free(mass_pc); mass_pc = NULL;
}

/* dynamics_apply (src/cmc/cmc_dynamics.c:17-542) on the harness's cluster state.  The strong-encounter routines it calls are re-bound at link
 * time (ld --wrap, oracle/Makefile) to the recorders below: an encounter pair is logged (type 0 = single-single, 1 = with a binary; k, kp,
 * then rperi, W, rcm as doubles) and left alone, so that everything else -- calc_encounter_dyns, calc_n_local, the cross sections, the
 * encounter probability, the two-body relaxation branch and the relaxation statistics -- is the reference's own code.
 * flags: RELAXATION, SS_COLLISION, BINSINGLE, BINBIN, TIDAL_CAPTURE, TC_POLYTROPE, CO_TDE, BH_CAPTURE, BH_LOSS_CONE, THREEBODYBINARIES, AVEKERNEL
 * par:   GAMMA, COLL_FACTOR, TC_FACTOR, WD_TC, units.mstar, units.l, units.t */
static long *enc_log_k = NULL; static double *enc_log_d = NULL; static long enc_log_cap = 0, enc_log_n = 0;
static void enc_record(long type, long k, long kp, double rperi, double W, double rcm)
{
	if (enc_log_n < enc_log_cap) {
		enc_log_k[3 * enc_log_n] = type; enc_log_k[3 * enc_log_n + 1] = k; enc_log_k[3 * enc_log_n + 2] = kp;
		enc_log_d[3 * enc_log_n] = rperi; enc_log_d[3 * enc_log_n + 1] = W; enc_log_d[3 * enc_log_n + 2] = rcm;
	}
	enc_log_n++;
}
void __wrap_sscollision_do(long k, long kp, double rperi, double w[4], double W, double rcm, double vcm[4], gsl_rng *rng)
{ (void) w; (void) vcm; (void) rng; enc_record(0, k, kp, rperi, W, rcm); }
void __wrap_binint_do(long k, long kp, double rperi, double w[4], double W, double rcm, double vcm[4], gsl_rng *rng)
{ (void) w; (void) vcm; (void) rng; enc_record(1, k, kp, rperi, W, rcm); }
void __wrap_break_wide_binaries(struct rng_t113_state *rng_st) { (void) rng_st; }
void __wrap_bh_rand_walk(long index, double v[4], double vcm[4], double beta, double dt)
{ (void) index; (void) v; (void) vcm; (void) beta; (void) dt; fprintf(stderr, "ref_harness: bh_rand_walk is outside the harness\n"); exit(8); }
void __wrap_make_threebodybinary(void) { fprintf(stderr, "ref_harness: make_threebodybinary is outside the harness\n"); exit(8); }
double __wrap_sigma_tc_nd(double n, double m1, double r1, double m2, double vinf)
{ (void) n; (void) m1; (void) r1; (void) m2; (void) vinf; fprintf(stderr, "ref_harness: TC_POLYTROPE is outside the harness\n"); exit(8); }
double __wrap_sigma_tc_nn(double na, double ma, double ra, double nb, double mb, double rb, double vinf)
{ (void) na; (void) ma; (void) ra; (void) nb; (void) mb; (void) rb; (void) vinf; fprintf(stderr, "ref_harness: TC_POLYTROPE is outside the harness\n"); exit(8); }

EXPORT long ref_dynamics_apply(double dt, const long *flags, const double *par, long *log_k, double *log_d, long log_cap, char *relax_text, long relax_cap)
{
	RELAXATION = flags[0]; SS_COLLISION = flags[1]; BINSINGLE = (int) flags[2]; BINBIN = (int) flags[3]; TIDAL_CAPTURE = flags[4];
	TC_POLYTROPE = flags[5]; CO_TDE = (int) flags[6]; BH_CAPTURE = flags[7]; BH_LOSS_CONE = flags[8]; THREEBODYBINARIES = (int) flags[9];
	AVEKERNEL = flags[10];
	GAMMA = par[0]; COLL_FACTOR = par[1]; TC_FACTOR = par[2]; WD_TC = par[3]; units.mstar = par[4]; units.l = par[5]; units.t = par[6];
	if (!mpi_logfile_buf) { mpi_logfile_buf = malloc(STR_BUF_LEN); mpi_logfile_wrbuf = malloc(1 << 20); }
	if (!mpi_relaxationfile_buf) { mpi_relaxationfile_buf = malloc(STR_BUF_LEN); mpi_relaxationfile_wrbuf = malloc(1 << 16); }
	mpi_logfile_wrbuf[0] = 0; mpi_logfile_len = 0;
	mpi_relaxationfile_wrbuf[0] = 0; mpi_relaxationfile_len = 0;
	{ /* sized for this run's arrays (a test process initialises the library more than once) */
		free(sigma_array.r); free(sigma_array.sigma);
		sigma_array.n = 0;
		sigma_array.r = (double *) calloc(N_STAR_DIM_OPT, sizeof(double));      /* cmc_io.c:2103-2104 */
		sigma_array.sigma = (double *) calloc(N_STAR_DIM_OPT, sizeof(double));
	}
	enc_log_k = log_k; enc_log_d = log_d; enc_log_cap = log_cap; enc_log_n = 0;
	dynamics_apply(dt, NULL);
	if (relax_text && relax_cap > 0) { strncpy(relax_text, mpi_relaxationfile_wrbuf, relax_cap - 1); relax_text[relax_cap - 1] = 0; }
	return enc_log_n;
}
EXPORT const double *ref_sigma_array(int which) { return which ? sigma_array.sigma : sigma_array.r; }

/* simul_relax_new (src/cmc/cmc_dynamics_helper.c:2526-2580) on the harness's cluster state */
EXPORT double ref_simul_relax_new(long avekernel, double thetasemax) { AVEKERNEL = avekernel; THETASEMAX = thetasemax; return simul_relax_new(); }

/* timing leg for bench.py: one pass of orbit step -> (caller refreshes r) -> sort -> post_sort_comm -> potential.
 * Brackets follow timeStartSimple/timeEndSimple (cmc_utils.c:1681-1707): barrier, wall clock. */
EXPORT void ref_time_step(double *t_orb, double *t_sort, double *t_comm, double *t_pot)
{
	double t0;
	long i;
	MPI_Barrier(MPI_COMM_WORLD); t0 = MPI_Wtime();
	ref_new_orbits_calculate();
	MPI_Barrier(MPI_COMM_WORLD); *t_orb = MPI_Wtime() - t0;
	/* what ComputeIntermediateEnergy does to create the sort key (cmc_utils.c:368-391): r <- rnew, m <- star_m */
	for (i = 1; i <= clus.N_MAX_NEW; i++) {
		star[i].rOld = star_r[get_global_idx(i)]; star[i].r = star[i].rnew; star[i].vr = star[i].vrnew; star[i].vt = star[i].vtnew;
		star[i].m = star_m[get_global_idx(i)];
	}
	MPI_Barrier(MPI_COMM_WORLD); t0 = MPI_Wtime();
	qsorts_new();
	MPI_Barrier(MPI_COMM_WORLD); *t_sort = MPI_Wtime() - t0; t0 = MPI_Wtime();
	post_sort_comm();
	MPI_Barrier(MPI_COMM_WORLD); *t_comm = MPI_Wtime() - t0; t0 = MPI_Wtime();
	calc_potential_new();
	MPI_Barrier(MPI_COMM_WORLD); *t_pot = MPI_Wtime() - t0;
}

/* write_snapshot (src/cmc/cmc_io.c:2960-3345) compiled unmodified; the HDF5 calls it makes are recorders: the rows it hands to
 * H5TBmake_table (rank 0) / H5TBappend_records (the other ranks) are what the test compares -- the N-sized part of a snapshot.  The table's
 * geometry (field names, offsets, types, record size) is recorded as well. */
static void *snap_rows = NULL;
static long snap_nrecords = 0, snap_row_size = 0, snap_nfields = 0;
static size_t snap_offsets[80];
static long snap_types[80];
static char snap_names[80][32];
long H5Fcreate(const char *name, unsigned flags, long a, long b) { (void) name; (void) flags; (void) a; (void) b; return 1; }
long H5Fopen(const char *name, unsigned flags, long a) { (void) name; (void) flags; (void) a; return 1; }
int H5Fclose(long f) { (void) f; return 0; }
int H5Zfilter_avail(int filter) { (void) filter; return 1; }
int H5Zget_filter_info(int filter, unsigned int *info) { (void) filter; *info = 3u; return 0; }
static void snap_keep(long nrecords, size_t type_size, const void *data)
{
	free(snap_rows);
	snap_nrecords = nrecords; snap_row_size = (long) type_size;
	snap_rows = malloc((size_t) (nrecords > 0 ? nrecords : 1) * type_size);
	memcpy(snap_rows, data, (size_t) nrecords * type_size);
}
int H5TBmake_table(const char *title, long loc, const char *dset, unsigned long long nfields, unsigned long long nrecords, size_t type_size,
		   const char *field_names[], const size_t *field_offset, const long *field_types, unsigned long long chunk_size, void *fill_data,
		   int compress, const void *data)
{
	unsigned long long i;
	(void) title; (void) loc; (void) dset; (void) chunk_size; (void) fill_data; (void) compress;
	snap_nfields = (long) nfields;
	for (i = 0; i < nfields && i < 80; i++) {
		snap_offsets[i] = field_offset[i]; snap_types[i] = field_types[i];
		strncpy(snap_names[i], field_names[i], 31); snap_names[i][31] = 0;
	}
	snap_keep((long) nrecords, type_size, data);
	return 0;
}
int H5TBappend_records(long loc, const char *dset, unsigned long long nrecords, size_t type_size, const size_t *field_offset, const size_t *dst_sizes,
		       const void *data)
{
	(void) loc; (void) dset; (void) field_offset; (void) dst_sizes;
	snap_keep((long) nrecords, type_size, data);
	return 0;
}
/* units_m, units_l: units.m, units.l; sigma: sigma_array.sigma[1 .. n_local] (used for binaries' eta only); returns the number of rows */
EXPORT long ref_write_snapshot(int bh_only, double units_m, double units_l, const double *sigma, long n_sigma, long avekernel)
{
	long i;
	units.m = units_m; units.l = units_l;
	AVEKERNEL = avekernel;
	if (sigma) {
		if (sigma_array.n < n_sigma || !sigma_array.sigma) {
			sigma_array.r = realloc(sigma_array.r, (size_t) (n_sigma + 2) * sizeof(double));
			sigma_array.sigma = realloc(sigma_array.sigma, (size_t) (n_sigma + 2) * sizeof(double));
			sigma_array.n = n_sigma;
		}
		for (i = 1; i <= n_sigma; i++) sigma_array.sigma[i] = sigma[i - 1];
	}
	write_snapshot("unused.h5", bh_only, "snapshot");
	return snap_nrecords;
}
EXPORT const void *ref_snapshot_rows(void) { return snap_rows; }
EXPORT long ref_snapshot_row_size(void) { return snap_row_size; }
EXPORT long ref_snapshot_nfields(void) { return snap_nfields; }
EXPORT long ref_snapshot_field(long i, char *name, long *type) { strcpy(name, snap_names[i]); *type = snap_types[i]; return (long) snap_offsets[i]; }

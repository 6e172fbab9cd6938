/* TEST INFRASTRUCTURE ONLY.  CPU restatement ("port") of the reference's per-timestep
 * data-parallel core: radius sort, potential rebuild, new-orbit step.  Explicit arguments
 * instead of the reference's process globals; every function cites the reference file:line
 * it follows (paths relative to the reference root).
 *
 * Pinning: the reference's own tests hold no golden vector for this path, so this
 * port is pinned against the reference ITSELF, compiled unmodified into oracle/_ref/libcmc_ref.so and
 * run on identical inputs (tests/test_oracle_vs_ref.py), and against fixtures that run generated
 * (tests/golden/, tests/golden/make_golden.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it. */
#ifndef CMC_ORACLE_H
#define CMC_ORACLE_H

#define ORC_SF_INFINITY 1.0e10 /* include/cmc/cmc.h:58 */
#define ORC_ZERO 1.0e-20       /* include/cmc/cmc.h:60 */
#define ORC_N_TRY 50000        /* include/cmc/cmc.h:64 */

/* per-star outcome of the orbit step (the branches of get_positions_loop, cmc_evolution_thr.c:802-933) */
enum {
	ORC_ST_OK = 0,        /* new radius sampled                               :857-933 */
	ORC_ST_MASSLESS = 1,  /* star_m < ZERO -> destroy_obj                     :816-820 */
	ORC_ST_UNBOUND = 2,   /* E >= 0 -> remove_star                            :823-829 */
	ORC_ST_CIRCULAR = 3,  /* circular_flag -> keeps r, vr, vt                 :835-839 */
	ORC_ST_TIDAL = 4,     /* rmax >= Rtidal -> remove_star                    :848-855 */
	ORC_ST_NTRY = 5       /* N_TRY exceeded (fatal in the reference)          :890-893 */
};

/* the replicated tables every function on the path reads (include/cmc/cmc_vars.h:218-226) */
typedef struct {
	const double *r, *m, *phi; /* star_r, star_m, star_phi: [0] sentinel, [1..N_MAX], [N_MAX+1] = SF_INFINITY/0 */
	long N_MAX;
	long N_STAR;               /* clus.N_STAR: the 1/N_STAR in MPI_PHI_S */
	double cenma_m;            /* cenma.m */
	double madhoc;             /* 1/N_STAR (cmc_utils.c:1183) */
} orc_tables_t;

typedef struct { /* orbit_rs_t, include/cmc/cmc.h:1621-1650 */
	double rp, ra, dQdrp, dQdra;
	long kmax, kmin;
	int circular_flag;
	int used_brent; /* not in the reference: counts find_root_vr fallbacks */
} orc_orbit_t;

#ifdef __cplusplus
extern "C" {
#endif

/* partition arithmetic: findLimits/findIndices cmc_utils.c:2299-2333 (== mpiFindIndicesCustom cmc_mpi.c:77-114) */
void orc_find_limits(long N, int blk, int procs, int *Start, int *End);

/* potential_calculate, cmc_utils.c:675-736.  star_r[N_MAX+1], star_phi[0..N_MAX+1] are written.  Returns new N_MAX. */
long orc_potential_calculate(double *star_r, const double *star_m, double *star_phi, long N_MAX_in, long N_STAR,
			     double *cenma_m, double *cenma_m_new, double madhoc, double orbit_r, double *Mtotal, double *Rtidal);

long orc_FindZero_r(const double *star_r, long kmin, long kmax, double r);          /* cmc_nr.c:62-82 */
double orc_potential(const orc_tables_t *t, double r);                               /* cmc_utils.c:112-165 */
double orc_function_q(const orc_tables_t *t, long j, long double r, long double pot, long double E, long double J); /* cmc_utils.c:1654-1662 */
long orc_find_zero_Q(const orc_tables_t *t, long j, long kmin, long kmax, long double E, long double J);           /* cmc_nr.c:316-383 */
/* calc_orbit_new, cmc_orbit.c:413-540.  g_si = global index of the star; is_new/rnew: star created this step (cmc_orbit.c:39-41) */
void orc_calc_orbit_new(const orc_tables_t *t, long g_si, int is_new, double rnew_of_new, double E, double J, orc_orbit_t *o);
double orc_find_root_vr(const orc_tables_t *t, long index, long k, double E, double J); /* cmc_nr.c:500-603 */

/* get_positions_loop, cmc_evolution_thr.c:784-937, for n_local stars whose global indices are g_idx[0..n_local).
 * E_in is star[j].E (WITHOUT the self term; the loop adds it, :806).
 * draws: the uniform stream the loop consumes, in order.  draw_off[j] (out, n_local+1 entries) = index of the first
 * draw star j consumed; draw_off[n_local] = total consumed.  If draws_per_star_off != NULL it is used as INPUT
 * (CSR offsets; star j reads draws[off[j]..off[j+1])) instead of the single sequential stream.
 * Outputs are per-star arrays; stars whose status is not ORC_ST_OK/CIRCULAR get rnew = SF_INFINITY, vrnew = vtnew = 0
 * like remove_star/destroy_obj do (cmc_evolution_thr.c:662-664, 760-765).
 * Returns max_rad. */
double orc_get_positions(const orc_tables_t *t, long n_local, const long *g_idx, const unsigned char *is_new,
			 const double *E_in, const double *J_in, const double *vr_in, const double *vt_in, const double *rnew_in,
			 double Rtidal, const double *draws, long n_draws, const long *csr_off_in, long *draw_off_out,
			 int *status, double *rnew, double *vrnew, double *vtnew, double *X, double *r_peri, double *r_apo,
			 long *kmin_out, long *kmax_out, long *n_brent);

/* Radius sort of a local star array + binary re-pack, single rank or the global result of the P-rank
 * sample sort (cmc_sort.c:375-629 + load_balance :634-791): ascending by the double at key_off, trailing
 * keys == SF_INFINITY trimmed (:343-354), binaries compacted in star order with binind = 1,2,... (:571-579, 757-766),
 * binary[nb+1..n_bin_dim-1] zeroed (:541).  star/binary point at element 1 (the reference passes star+1, binary+1).
 * perm_out[i] = original 0-based position of the record now at i.  Returns the trimmed count. */
long orc_sort_stars(unsigned char *star1, long n, long sizeof_star, long key_off, long binind_off,
		    unsigned char *binary1, long n_bin_dim, long sizeof_binary, long *n_bin_out, long *perm_out);

/* The two loops of the energy scheme that look the potential up at the NEW radii, between the orbit step and the sort:
 *   energy_conservation2, src/cmc/cmc_utils.c:2217-2226:  Uoldrnew[i] = potential(rnew[i]) + MPI_PHI_S(rnew[i], g_i)  for every local star
 *   ComputeIntermediateEnergy's first loop, :371-378:       EI[i] = sqr(vr[i]) + sqr(vt[i]) + star_phi[g_i] - potential(rnew[i])  where rnew[i] < 1e6
 * g_idx: global index of each local star.  EI entries of escaped stars are left untouched. */
void orc_energy_conservation2(const orc_tables_t *t, long n_local, const long *g_idx, const double *rnew, double *Uoldrnew);
void orc_intermediate_energy(const orc_tables_t *t, long n_local, const long *g_idx, const double *rnew, const double *vr, const double *vt, double *EI);

/* The energy scheme's velocity correction after the potential rebuild (the last of the four potential() loops):
 *   energy_conservation1, src/cmc/cmc_utils.c:2153-2174: vrold, vtold <- vr, vt; Uoldrold = star_phi[g] + MPI_PHI_S(star_r[g], g)
 *   set_velocities3, :254-352 (called through energy_conservation3, :360-363), as ONE rank sees it:
 *     _first  = the loop :268-311 over the local stars with interacted == 0: the predicted squared speed vnew2 (one potential(rOld) each);
 *               vnew2 <= 0 adds the star's kinetic-energy deficit to the running Eexcess, otherwise (vr, vt) are rescaled to vnew2,
 *               copied to (vrnew, vtnew), and while Eexcess > 0 the star gives up MIN(0.8 * its kinetic energy, Eexcess).  The carry
 *               runs serially from star to star.  vnew2_arr[i] = (vnew2 > 0).  Returns the rank's leftover Eexcess.
 *     _second = the loop :323-343 of one pass of the do-while: starts from the Eexcess received from the previous rank (0 on rank 0)
 *               and lets every star with vnew2_arr == 1 give up MIN(v2*m/2 - 0.8*m*v2_new/2, Eexcess) while Eexcess > 0.  Returns
 *               what is left.  The caller carries the Send/Recv (rank i -> i+1), repeats while the all-reduced sum is > 0 (:345-347) and
 *               finally Eoops += -Eexcess*madhoc (:351).  On a single rank the received value is 0: the first loop's leftover is dropped. */
void orc_energy_conservation1(const orc_tables_t *t, long n_local, const long *g_idx, const double *vr, const double *vt, double *vrold,
			      double *vtold, double *Uoldrold);
double orc_set_velocities3_first(const orc_tables_t *t, long n_local, const long *g_idx, const long *interacted, const double *rOld,
				 const double *vrold, const double *vtold, const double *Uoldrold, const double *Uoldrnew, double *vr, double *vt,
				 double *vrnew, double *vtnew, char *vnew2_arr);
double orc_set_velocities3_second(const orc_tables_t *t, long n_local, const long *g_idx, const long *interacted, const char *vnew2_arr,
				  double Eexcess, double *vr, double *vt, const double *vrnew, const double *vtnew);

/* calc_sigma_r, src/cmc/cmc_dynamics_helper.c:2821-2975: the local 3-D velocity dispersion around every local star,
 * sigma = sqrt(sum m v^2 / sum m) over a window of 2p stars (p before, p - 1 after), as ONE rank computes it: two running sums that every
 * star updates by subtracting the stars that left the window and adding those that entered, in that order, each a separately rounded
 * operation; the window stops sliding at both ends of the cluster (:2923-2931).  A rank other than the first starts from an empty window p
 * stars before its first one.  vr, vt: the local stars 1..n_limit at [0..n_limit-1]; ghost_prev_* [p]: the last p stars of the previous
 * rank, ghost_next_* [p]: the first p of the next (the reference exchanges them with MPI_Send/Recv, :2862-2900); first_global = Start[myid].
 * Outputs [n_limit + 1], entry 0 unused: sig_r_or_mave = star_r[g] (r_0_mave_1 == 0) or the mean mass Mave / 2 / p; sig_sigma. */
void orc_calc_sigma_r(long p, long n_limit, long first_global, long N_MAX, int is_first_rank, const double *star_r, const double *star_m,
		      double madhoc, const double *vr, const double *vt, const double *ghost_prev_vr, const double *ghost_prev_vt,
		      const double *ghost_next_vr, const double *ghost_next_vt, int r_0_mave_1, double *sig_r_or_mave, double *sig_sigma);

/* dynamics_apply, src/cmc/cmc_dynamics.c:17-542: the loop over neighbouring pairs (:216-500) for everything that does not need
 * fewbody / collisions / BSE -- the pair's kinematics (calc_encounter_dyns, cmc_dynamics_helper.c:831-870), the local density (calc_n_local,
 * :768-786), the cross section S of the pair's type (:250-432), the encounter probability (:436), the decision (:444) and the two-body
 * relaxation branch (:468-500: Trel12, beta, the rotated relative velocity, new vr / vt, set_star_EJ) with its scattering-angle statistics
 * (:473-482).  A pair whose draw falls below P_enc is an ENCOUNTER: it is reported (status, rperi, W, rcm) and left to the caller, which
 * runs binint_do / sscollision_do on it like the reference does.
 * Pairs are (1,2), (3,4), ... of the local stars, skipping stars with threebb_interacted == 1 (:226-240; threebb may be NULL).
 * draws: three uniforms per pair in pair order -- Y (the angle between the tangential velocities, calc_encounter_dyns), the encounter draw,
 * psi (consumed only by a relaxed pair; an encounter pair leaves its third draw unused).
 * Star fields are the local stars 1..n_local at [0..n_local-1]; bin_a is indexed by binind (entry 0 unused; may be NULL without binaries).
 * TC_POLYTROPE, BH_LOSS_CONE with a central mass and THREEBODYBINARIES' own loop (:60-204) are outside this restatement.
 * pair_k / pair_kp [n_local / 2]: 1-based local indices of each pair's stars; pair_status: 0 relaxed, 1 single-single encounter, 2 encounter
 * with a binary, 3 untouched (RELAXATION == 0).  Returns the number of pairs. */
typedef struct {
	long relaxation, ss_collision, binsingle, binbin, tidal_capture, co_tde, bh_capture, p /* AVEKERNEL */;
	double gamma, coll_factor, tc_factor, wd_tc, units_mstar, units_l, units_t;
} orc_dyn_params_t;
typedef struct {
	long Nrel, Nrelbeta[4];
	double qaverelbeta[4], maverelbeta[4], raverelbeta[4];
} orc_relax_stats_t;
long orc_dynamics_apply_pairs(const orc_tables_t *t, long n_local, long first_global, double dt, const orc_dyn_params_t *dp, const double *draws,
			      double *vr, double *vt, double *E, double *J, double *Y, const long *binind, const int *se_k, const double *rad,
			      const double *bin_a, const long *threebb, long *pair_k, long *pair_kp, int *pair_status, double *pair_rperi,
			      double *pair_W, double *pair_rcm, orc_relax_stats_t *stats);

/* simul_relax_new, src/cmc/cmc_dynamics_helper.c:2526-2580: the relaxation time step of GetTimeStep (src/cmc/cmc_evolution_thr.c:317-331) as
 * ONE rank computes it -- the local stars in consecutive blocks of 2p (the last, incomplete block is left out: si < n_local - p), per block
 * the freshly summed m v^2, m, m^2, the local density at the block's centre, dt = (2 THETASEMAX / PI)^2 (PI / 32) W^3 / (N_STAR n (4 <m^2>)),
 * and the minimum over the blocks (+inf when there is no block).  The caller takes the minimum over the ranks (MPI_Allreduce MIN). */
double orc_simul_relax_new(const orc_tables_t *t, long n_local, long first_global, long p, double thetasemax, const double *vr, const double *vt);

/* tidally_strip_stars, src/cmc/cmc_evolution_thr.c:443-645 (between energy_conservation2 and ComputeIntermediateEnergy, cmc.c:321-322): one
 * pass of its do-while loop over the local stars.  A star still in the cluster (rnew < 1e6) is stripped when r_apo > threshold = Rtidal
 * (TIDAL_TREATMENT 0, :479) or E > threshold = gierszalpha * phi_rtidal (TIDAL_TREATMENT 1, :555): rnew = SF_INFINITY, vrnew = vtnew = 0.
 * idx_out receives the 1-based local indices in star order; returns their number.  The caller does what the reference does per stripped
 * star in that order (the escaper sums, the log line, destroy_obj), recomputes Rtidal from the mass lost and repeats while a pass strips
 * anything (:624): tests/oraclelib.py:tidally_strip_stars. */
long orc_tidally_strip_select(long n_local, int tidal_treatment, double threshold, const double *r_apo, const double *E, double *rnew,
			      double *vrnew, double *vtnew, long *idx_out);

/* ComputeEnergy's second loop, src/cmc/cmc_utils.c:619-636: the four sequentially summed energies of one rank's stars -- out[0] kinetic
 * (buf_reduce[1]), out[1] potential BEFORE the factor 0.5 (buf_reduce[2]: two additions per star, the star's own term and the central
 * mass's share), out[2] internal (buf_reduce[3]), out[3] binary binding (buf_reduce[4]).  binind[i-1] of local star i; the bin_* arrays
 * are indexed by binind (entry 0 unused). */
void orc_compute_energy_totals(const orc_tables_t *t, long n_local, long first_global, const double *vr, const double *vt, const double *Eint,
			       const long *binind, const int *bin_inuse, const double *bin_m1, const double *bin_m2, const double *bin_a,
			       const double *bin_Eint1, const double *bin_Eint2, double out[4]);

/* comp_multi_mass_percent, src/cmc/cmc_utils.c:769-861: Lagrange radii per mass bin.  A star belongs to bin b when mass_bins[b] <
 * star_m/SOLAR_MASS_DYN <= mass_bins[b+1] (find_stars_mass_bin: m_binsearch, stars outside [bins[0], bins[n-1]] belong to none); per bin the
 * cumulative mass fraction in star order (sequential sums) against the radius, linearly interpolated at mass_pc (gsl_interp_linear of the
 * un-vendored GSL, restated: PARITY UNPINNED for the interpolation step).  multi_mass_r[n_bins][count]: entries of empty bins are left
 * untouched, a bin of one star gets that star's radius. */
void orc_comp_multi_mass_percent(const double *star_r, const double *star_m, long N_MAX, const double *mass_bins, long n_bins, double solar_mass_dyn,
				 const double *mass_pc, long count, double *multi_mass_r);

/* comp_mass_percent, src/cmc/cmc_utils.c:872-985: Lagrange radii and the cumulative kinetic sums at them, as ONE rank of P sees it.
 *   star_r, star_m: the replicated tables; vr, vt: star[k].vr / .vt of this rank's local stars k = 1..n_local at vr[k-1], vt[k-1];
 *   first_global = Start[myid] (global index of local star 1; the function is called after the sort, so there are no new stars).
 *   mprev starts at cenma.m*madhoc when MASS_PC_BH_INCLUDE (:878-892) and the leading thresholds it already exceeds get
 *   mass_r = MINIMUM_R and zeros; then for k = 1..N_MAX: mprev += star_m[k]/N_STAR, and when mprev > mass_pc[mcount]*Mtotal the entry
 *   mcount is recorded at THIS k and mcount advances by one (:930-977) -- one threshold per k.
 *   The four cumulative sums (:907-913) are rank-local sequential sums from zero; for an index owned by a rank other than 0 the value is
 *   prefix[i] + local[i] (prefix = MPI_Exscan of the ranks' totals, :925-931,945-951); rank 0 copies its local value (:962-968).
 *   Outputs [count]: mass_r, ave_mass_r, no_star_r, densities_r on every rank (replicated loop); ke_rad_r, ke_tan_r, v2_rad_r, v2_tan_r only
 *   at the entries whose index this rank owns (others are left untouched; owner[] (may be NULL) receives 1 there, 0 elsewhere).
 *   totals[4] = the rank's local sums over all its stars (what it contributes to the Exscan).  Returns mcount: entries filled. */
long orc_comp_mass_percent(const double *star_r, const double *star_m, long N_MAX, long N_STAR, const double *vr, const double *vt, long n_local,
			   long first_global, const double prefix[4], int is_rank0, double cenma_m, double madhoc, double Mtotal,
			   double initial_total_mass, int bh_include, double minimum_r, const double *mass_pc, long count, double *mass_r,
			   double *ave_mass_r, double *no_star_r, double *densities_r, double *ke_rad_r, double *ke_tan_r, double *v2_rad_r,
			   double *v2_tan_r, int *owner, double totals[4]);

/* clusdyn_calculate, src/cmc/cmc_utils.c:1613-1624: the half-mass radius clusdyn.rh = star_r[k], k = one past the first index at which the
 * sequentially summed star_m[k]/N_STAR reaches 0.5 Mtotal */
double orc_clusdyn_calculate(const double *star_r, const double *star_m, long N_STAR, double Mtotal);

/* taus113 (src/libs/taus113-v2.c:166-216 seeding, :218-251 generator, :253-270 jump) */
void orc_t113_seed(unsigned long s, unsigned int z[4]);
double orc_t113_dbl(unsigned int z[4]);
void orc_t113_jump(unsigned int z[4], const unsigned int jpoly[4]);
void orc_t113_fill(unsigned int z[4], double *out, long n);
extern const unsigned int orc_JPoly_2_20[4], orc_JPoly_2_80[4];

#ifdef __cplusplus
}
#endif
#endif

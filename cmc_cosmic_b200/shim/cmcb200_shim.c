/* Reference-side binding of libcmcb200.so: defines CMC's four hot-path entry points
 *     void new_orbits_calculate(void)   src/cmc/cmc_utils.c:2194-2211   (get_positions: src/cmc/cmc_evolution_thr.c:944-1023)
 *     void qsorts_new(void)             src/cmc/cmc_utils.c:2232-2254   (sample_sort:   src/cmc/cmc_sort.c:375-629)
 *     void post_sort_comm(void)         src/cmc/cmc_utils.c:2259-2288
 *     void calc_potential_new(void)     src/cmc/cmc_utils.c:2006-2012   (potential_calculate: :675-736)
 * on the reference's own process globals (include/cmc/cmc_vars.h) and forwards to the C ABI in include/cmcb200.h.
 * This file is compiled against the reference's headers and linked INSTEAD of those four definitions
 * (-Wl,--wrap=..., objcopy --weaken-symbol, or an #ifndef around them); everything else in CMC --
 * relaxation, stellar evolution, fewbody, I/O -- is untouched and keeps calling these names.
 *
 * Host-side side effects of the orbit step (remove_star / destroy_obj / count_esc_bhs: escaper accumulators, the esc
 * log, zeroed records; src/cmc/cmc_evolution_thr.c:655-771, src/cmc/cmc_dynamics_helper.c:131-153) are replayed here,
 * in star order, by the reference's own functions from the per-star status codes the kernel returns -- so the
 * accumulators see the same additions in the same order as in get_positions_loop.
 *
 * Single rank per GPU.  procs > 1 goes through the radius-sharded entry points (cmcb200_shard_*).
 *
 * RESIDENT MODE (-DCMCB200_SHIM_RESIDENT, linked with --wrap): the whole block of the main loop from energy_conservation1() to ComputeEnergy()
 * (src/cmc/cmc.c:302-353) is bound to the device-resident pipeline, so the star records cross the host link ONCE each way per timestep:
 *     energy_conservation1        opens the window: cmcb200_resident_begin starts the upload of star[] / binary[] / star_r / star_m / star_phi
 *     new_orbits_calculate        the orbit step, chunk by chunk behind the arriving records; then energy_conservation1's pass; host side
 *                                 effects (remove_star ...) replayed by the reference's own functions and mirrored to the device
 *     energy_conservation2, tidally_strip_stars, ComputeIntermediateEnergy, qsorts_new, post_sort_comm, calc_potential_new (tables come
 *     back here: potential() has host callers), energy_conservation3, comp_mass_percent
 *     ComputeEnergy               E, J, the energy sums; closes the window: cmcb200_resident_end downloads records, binaries and tables
 * Outside a window (start-up: src/cmc/cmc.c:133-167) the four hot-path functions take the host-buffer form and the others the reference's
 * own code (__real_*).  Between the first and the last call of a window the host must not touch star[] / binary[]: true for the block above
 * when STAR_AGING_SCHEME == 0 (toy_rejuvenation is then a no-op); with procs > 1 or STAR_AGING_SCHEME > 0 no window is opened. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "cmc.h"
#include "cmc_vars.h"
#include "cmcb200.h"

#ifdef CMCB200_SHIM_WRAP /* linked with -Wl,--wrap=new_orbits_calculate,--wrap=qsorts_new,... */
#define SHIM(f) __wrap_##f
#else
#define SHIM(f) f
#endif

static cmcb200_ctx *ctx = NULL;
static long ctx_star_dim = 0, ctx_bin_dim = 0;
static int *status_buf = NULL;
static long status_cap = 0;
/* parity mode: a recorded uniform stream in CSR form (NULL = production mode, counter-based Philox on the device) */
static const double *shim_draws = NULL;
static const long *shim_draw_off = NULL;
static cmcb200_orbit_result last_orbit;

static void shim_fail(const char *where)
{
	/* the reference's error convention: eprintf + exit_cleanly (src/cmc/cmc_utils.c:208-220) */
	eprintf("%s: %s\n", where, cmcb200_last_error());
	exit_cleanly(-1, where);
}

static void *pinned[5]; /* the arrays shim_init registered; they must be unregistered before the caller frees them */
static void shim_unpin(void)
{
	int i;
	for (i = 0; i < 5; i++) { if (pinned[i]) cmcb200_host_unregister(pinned[i]); pinned[i] = NULL; }
}

static void shim_init(void)
{
	long lay[16];
	int dev = 0;
	const char *e = getenv("CMCB200_DEVICE");
	if (ctx && ctx_star_dim == N_STAR_DIM_OPT && ctx_bin_dim == N_BIN_DIM_OPT) return;
	if (ctx) { shim_unpin(); cmcb200_destroy(ctx); ctx = NULL; } /* the arrays were re-dimensioned (a new run in the same process) */
	cmcb200_layout(lay);
	if (lay[0] != (long) sizeof(star_t) || lay[1] != (long) sizeof(binary_t) || lay[2] != (long) offsetof(star_t, r) ||
	    lay[6] != (long) offsetof(star_t, E) || lay[7] != (long) offsetof(star_t, J) || lay[8] != (long) offsetof(star_t, rnew) ||
	    lay[12] != (long) offsetof(star_t, X) || lay[13] != (long) offsetof(star_t, r_peri) || lay[15] != (long) offsetof(star_t, binind)) {
		eprintf("cmcb200 shim: libcmcb200.so was built for a different star_t/binary_t layout\n");
		exit_cleanly(-1, __FUNCTION__);
	}
	if (e) dev = atoi(e);
	else dev = myid; /* one GPU per rank of the node */
	if (procs == 1) {
		if (cmcb200_create(&ctx, dev, N_STAR_DIM_OPT, N_BIN_DIM_OPT)) shim_fail("cmcb200_create");
	} else {
		/* radius-sharded operation (include/cmcb200.h): this rank's share of star[] / binary[], the whole tables; the ranks map each other's
		 * device buffers once (CUDA IPC handles exchanged with one MPI_Allgather) and later store records straight into them over NVLink */
		const int hb = cmcb200_shard_handle_bytes();
		char *mine = (char *) malloc(hb), *all = (char *) malloc((size_t) hb * procs);
		if (cmcb200_create_sharded(&ctx, dev, N_STAR_DIM_OPT, N_STAR_DIM, N_BIN_DIM_OPT)) shim_fail("cmcb200_create_sharded");
		if (cmcb200_shard_export(ctx, myid, procs, mine)) shim_fail("cmcb200_shard_export");
		MPI_Allgather(mine, hb, MPI_BYTE, all, hb, MPI_BYTE, MPI_COMM_WORLD);
		if (cmcb200_shard_attach(ctx, all)) shim_fail("cmcb200_shard_attach");
		cmcb200_shard_fuse_transfer(ctx, 0); /* qsorts_new sorts on .r: the caller has made the r <- rnew transfer itself */
		free(mine); free(all);
	}
	ctx_star_dim = N_STAR_DIM_OPT;
	ctx_bin_dim = N_BIN_DIM_OPT;
	/* the arrays are allocated once (src/cmc/cmc_io.c:2099-2101, 3434-3436): pin them so that every copy is DMA at link rate.  A failure
	 * (locked-memory limit) is not fatal -- pageable memory only copies slower. */
	if (!getenv("CMCB200_NO_PIN")) {
		void *arr[5] = {star, binary, star_r, star_m, star_phi};
		size_t bytes[5] = {(size_t) N_STAR_DIM_OPT * sizeof(star_t), (size_t) N_BIN_DIM_OPT * sizeof(binary_t), (size_t) N_STAR_DIM * sizeof(double),
				   (size_t) N_STAR_DIM * sizeof(double), (size_t) N_STAR_DIM * sizeof(double)};
		int i;
		for (i = 0; i < 5; i++) pinned[i] = cmcb200_host_register(arr[i], bytes[i]) == 0 ? arr[i] : NULL;
	}
}

#define SHIM_EXPORT __attribute__((visibility("default")))
static void shim_reset_resident(void);
SHIM_EXPORT void cmcb200_shim_set_draws(const double *draws, const long *draw_off) { shim_draws = draws; shim_draw_off = draw_off; }
SHIM_EXPORT const cmcb200_orbit_result *cmcb200_shim_last_orbit(void) { return &last_orbit; }
SHIM_EXPORT void cmcb200_shim_shutdown(void)
{
	shim_unpin();
	if (ctx && procs > 1) { cmcb200_sync(ctx); MPI_Barrier(MPI_COMM_WORLD); cmcb200_shard_detach(ctx); MPI_Barrier(MPI_COMM_WORLD); } /* nobody unmaps a buffer a peer still uses */
	if (ctx) cmcb200_destroy(ctx);
	ctx = NULL; free(status_buf); status_buf = NULL; status_cap = 0;
	shim_draws = NULL; shim_draw_off = NULL; /* the caller's arrays: not ours to keep beyond a run */
	shim_reset_resident();
}

/* src/cmc/cmc_utils.c:2194-2211 */
#ifdef CMCB200_SHIM_RESIDENT
static void res_new_orbits_calculate(void);
static void res_qsorts_new(void);
static void res_calc_potential_new(void);
static int win_open;
#define RES_ROUTE(f) do { if (win_open) { f(); return; } } while (0)
#else
#define RES_ROUTE(f) do { } while (0)
#endif

void SHIM(new_orbits_calculate)(void)
{
	cmcb200_orbit_params p;
	double phi_rtidal, phi_zero;
	long j, n = clus.N_MAX_NEW, n_old = End[myid] - Start[myid] + 1;
	RES_ROUTE(res_new_orbits_calculate);
	shim_init();
	memset(&p, 0, sizeof p);                         /* every field the shim does not set (fuse_compute_energy: resident pipeline only) is 0 */
	OldTidalMassLoss = TidalMassLoss_old;
	phi_rtidal = potential(Rtidal);                  /* get_positions, cmc_evolution_thr.c:963-964 */
	phi_zero = potential(0.0);
	if (n > status_cap) { status_cap = n + n / 4 + 64; status_buf = realloc(status_buf, status_cap * sizeof(int)); }
	p.n_local = n;
	p.n_old_local = n_old;
	p.first_global = Start[myid];
	p.N_MAX = clus.N_MAX;
	p.N_STAR = clus.N_STAR;
	p.n_table = clus.N_MAX + 2 + (n > n_old ? n - n_old : 0);
	p.cenma_m = cenma.m;
	p.Rtidal = Rtidal;
	p.draws = shim_draws;
	p.draw_off = shim_draw_off;
	p.seed = (unsigned long long) IDUM;
	p.tcount = (unsigned long long) tcount;
	if (cmcb200_new_orbits_calculate(ctx, star + 1, star_r, star_m, star_phi, &p, status_buf, NULL, NULL, &last_orbit)) shim_fail("new_orbits_calculate");
	for (j = 1; j <= n; j++) {                       /* the branches of get_positions_loop that touch host state, in star order */
		switch (status_buf[j - 1]) {
		case CMCB200_ORBIT_MASSLESS: destroy_obj(j); break;                                           /* :816-820 */
		case CMCB200_ORBIT_UNBOUND: count_esc_bhs(j); remove_star(j, phi_rtidal, phi_zero); break;    /* :823-829 */
		case CMCB200_ORBIT_TIDAL: remove_star(j, phi_rtidal, phi_zero); break;                        /* :848-855 */
		case CMCB200_ORBIT_NTRY:                                                                      /* :890-893 */
			eprintf("N_TRY exceeded\n");
			exit_cleanly(-1, __FUNCTION__);
		default: break;
		}
	}
	max_r = last_orbit.max_r;
	MPI_Allreduce(MPI_IN_PLACE, &max_r, 1, MPI_DOUBLE, MPI_MAX, MPI_COMM_WORLD);
	MPI_Allreduce(MPI_IN_PLACE, &TidalMassLoss, 1, MPI_DOUBLE, MPI_SUM, MPI_COMM_WORLD);
	TidalMassLoss += TidalMassLoss_old;
}

/* procs > 1: sample_sort + load_balance + the Allgatherv pair of post_sort_comm (src/cmc/cmc_sort.c:375-791, src/cmc/cmc_utils.c:2269-2280) as
 * the radius-sharded sequence of include/cmcb200.h.  MPI supplies two small collectives and the barriers between the phases; the
 * records travel GPU to GPU.  Every rank ends with its findLimits share of the globally sorted stars, the binaries that follow them
 * (binind = 1, 2, ..) and the full star_r / star_m tables -- what the reference's code leaves, bit for bit. */
static void shim_sharded_sort(long n, long nb)
{
	long keep = 0, total = 0, nl, nb_out = 0, *keep_all = (long *) malloc(procs * sizeof(long));
	int q;
	if (cmcb200_upload(ctx, star + 1, n, binary + 1, nb, star_r, star_m, NULL, 1)) shim_fail("upload");          /* [0]: the table sentinel */
	if (cmcb200_shard_sort_local(ctx, n, &keep)) shim_fail("shard_sort_local");
	MPI_Allgather(&keep, 1, MPI_LONG, keep_all, 1, MPI_LONG, MPI_COMM_WORLD);                                    /* also: every rank's keys have landed */
	for (q = 0; q < procs; q++) total += keep_all[q];
	findLimits(total, MIN_CHUNK_SIZE);               /* the partition the reference's load_balance realises (cmc_sort.c:421, 634-791) */
	if (cmcb200_shard_exchange(ctx, keep_all, Start, End, 0, 0, 0) || cmcb200_sync(ctx)) shim_fail("shard_exchange");
	MPI_Barrier(MPI_COMM_WORLD);
	nl = End[myid] - Start[myid] + 1;
	if (cmcb200_shard_publish(ctx, nl, Start[myid], N_BIN_DIM_OPT - 1, &nb_out)) shim_fail("shard_publish");
	MPI_Barrier(MPI_COMM_WORLD);
	cmcb200_shard_finish(ctx);
	if (cmcb200_download(ctx, star + 1, nl, binary + 1, nb_out, star_r, star_m, NULL, total + 1)) shim_fail("download");
	if (N_BIN_DIM_OPT - 1 > nb_out) memset(binary + 1 + nb_out, 0, (size_t) (N_BIN_DIM_OPT - 1 - nb_out) * sizeof(binary_t));    /* cmc_sort.c:541 */
	clus.N_MAX = total;                              /* sample_sort's return value (cmc_sort.c:628) */
	clus.N_MAX_NEW = nl;
	N_b_local = nb_out;                              /* cmc_sort.c:741 */
	free(keep_all);
}

/* src/cmc/cmc_utils.c:2232-2254; the r/m extraction of post_sort_comm (:2269-2280) is fused into the same device pass */
void SHIM(qsorts_new)(void)
{
	long n = clus.N_MAX_NEW, nb = N_b_local, i;
	RES_ROUTE(res_qsorts_new);
	shim_init();
	/* binaries referenced by the stars being sorted occupy binary[1..max binind]; N_b_local is only refreshed by load_balance */
	for (i = 1; i <= n; i++) if (star[i].binind > nb) nb = star[i].binind;
	if (procs > 1) { shim_sharded_sort(n, nb); return; }
	if (cmcb200_qsorts_new(ctx, star + 1, &n, binary + 1, N_BIN_DIM_OPT, &nb, star_r, star_m)) shim_fail("qsorts_new");
	clus.N_MAX = n;                                  /* sample_sort's return value: global count after trimming (cmc_sort.c:628) */
	clus.N_MAX_NEW = n;
	N_b_local = nb;                                  /* cmc_sort.c:741 */
	findLimits(clus.N_MAX, MIN_CHUNK_SIZE);          /* find_expected_count (cmc_sort.c:421, :260-273) leaves Start/End for the new N_MAX */
}

/* src/cmc/cmc_utils.c:2259-2288: star_r / star_m were already written by qsorts_new; on one rank the reductions are identities */
void SHIM(post_sort_comm)(void)
{
	MPI_Allreduce(MPI_IN_PLACE, &cenma.m_new, 1, MPI_DOUBLE, MPI_SUM, MPI_COMM_WORLD);
	MPI_Allreduce(MPI_IN_PLACE, &cenma.E_new, 1, MPI_DOUBLE, MPI_SUM, MPI_COMM_WORLD);
}

/* src/cmc/cmc_utils.c:2006-2012 -> potential_calculate :675-736 */
void SHIM(calc_potential_new)(void)
{
	long nmax = clus.N_MAX;
	RES_ROUTE(res_calc_potential_new);
	shim_init();
	if (cmcb200_calc_potential_new(ctx, star_r, star_m, star_phi, &nmax, clus.N_STAR, &cenma.m, &cenma.m_new, orbit_r, &Mtotal, &Rtidal))
		shim_fail("calc_potential_new");
	clus.N_MAX = nmax;
	mpiFindIndicesCustom(clus.N_MAX, MIN_CHUNK_SIZE, myid, &mpiBegin, &mpiEnd);
}


#ifdef CMCB200_SHIM_RESIDENT
#ifndef CMCB200_SHIM_WRAP
#error "the resident mode needs the --wrap build: outside a window it calls the reference's own functions as __real_*"
#endif
void __real_energy_conservation1(void);
void __real_energy_conservation2(void);
void __real_energy_conservation3(void);
void __real_tidally_strip_stars(void);
void __real_ComputeIntermediateEnergy(void);
void __real_ComputeEnergy(void);
void __real_comp_mass_percent(void);
void __real_dynamics_apply(double dt, gsl_rng *rng);

static int win_open = 0;      /* (declared above) the records live on the device; the host copy of star[] / binary[] is stale */
static int ec1_pending = 0;   /* energy_conservation1's pass has been deferred behind the orbit step (same result: see new_orbits_calculate) */
static unsigned long long dirty_words = ~0ull; /* what the host may have written since the last window (cmcb200_shim_set_dirty_words) */
static long win_n_bin = 0;    /* binary slots uploaded */

/* the caller's promise for cmcb200_resident_begin: bit w = word w of star_t may have changed since the last ComputeEnergy() */
SHIM_EXPORT void cmcb200_shim_set_dirty_words(unsigned long long mask) { dirty_words = mask; }
SHIM_EXPORT int cmcb200_shim_window_open(void) { return win_open; }

static void res_params(cmcb200_orbit_params *p)
{
	long n = clus.N_MAX_NEW, n_old = End[myid] - Start[myid] + 1;
	memset(p, 0, sizeof *p);
	p->n_local = n;
	p->n_old_local = n_old;
	p->first_global = Start[myid];
	p->N_MAX = clus.N_MAX;
	p->N_STAR = clus.N_STAR;
	p->n_table = clus.N_MAX + 2 + (n > n_old ? n - n_old : 0);
	p->cenma_m = cenma.m;
	p->Rtidal = Rtidal;
	p->draws = shim_draws;
	p->draw_off = shim_draw_off;
	p->seed = (unsigned long long) IDUM;
	p->tcount = (unsigned long long) tcount;
}

static int res_allowed(void) { return procs == 1 && STAR_AGING_SCHEME == 0 && !getenv("CMCB200_NO_RESIDENT"); }

static void res_flush_ec1(void)
{
	cmcb200_orbit_params p;
	if (!ec1_pending) return;
	res_params(&p);
	if (cmcb200_dev_energy_conservation1(ctx, &p)) shim_fail("energy_conservation1");
	ec1_pending = 0;
}

/* one record / table entry / binary back to the device after the reference's host code rewrote it (remove_star, destroy_obj) */
static void res_push_star(long j, long k)
{
	int g = get_global_idx(j);
	if (cmcb200_dev_put(ctx, star + j, j, k > 0 ? binary + k : NULL, k, star_m + g, g)) shim_fail("dev_put");
}

/* src/cmc/cmc_utils.c:2153-2174.  Opens the window (unless dynamics_apply did).  The pass itself reads vr, vt and the tables and writes vrold, vtold, Uoldrold; the orbit
 * step that follows reads and writes none of these, so the pass is queued BEHIND it (new_orbits_calculate below) -- the orbit step can then
 * start on the first uploaded chunk instead of waiting for the last one. */
static void res_open_window(void)
{
	long n_new;
	shim_init();
	n_new = clus.N_MAX_NEW - (End[myid] - Start[myid] + 1);
	win_n_bin = N_BIN_DIM_OPT - 1; /* a star may point at any slot (binaries are created in free slots between two sorts) */
	if (cmcb200_resident_begin(ctx, star + 1, clus.N_MAX_NEW, binary + 1, win_n_bin, star_r, star_m, star_phi, clus.N_MAX + 2 + (n_new > 0 ? n_new : 0), dirty_words))
		shim_fail("resident_begin");
	win_open = 1;
}

void SHIM(energy_conservation1)(void)
{
	if (win_open) { ec1_pending = 1; return; } /* dynamics_apply opened the window already */
	if (!res_allowed()) { __real_energy_conservation1(); return; }
	res_open_window();
	ec1_pending = 1;
}

/* src/cmc/cmc_dynamics.c:17-542.  With nothing between it and energy_conservation1 that reads the records on the host
 * (STELLAR_EVOLUTION == 0: src/cmc/cmc.c:286-302) the window opens HERE: calc_sigma_r and the loop over neighbouring pairs run on the
 * resident records behind the upload, and only the pairs whose encounter draw fell below P_enc come back for the reference's own
 * sscollision_do / binint_do.  The reference's host code that follows the loop (break_wide_binaries) walks every binary and dumps excess
 * energy on the stars behind it, so the binding is taken for clusters WITHOUT binaries only; everything else runs __real_dynamics_apply. */
static const double *shim_dyn_draws = NULL;
SHIM_EXPORT void cmcb200_shim_set_dynamics_draws(const double *draws) { shim_dyn_draws = draws; }
static cmcb200_dynamics_result last_dynamics;
SHIM_EXPORT const cmcb200_dynamics_result *cmcb200_shim_last_dynamics(void) { return &last_dynamics; }
/* flushes what is pending and closes the window (tests; a caller that must read star[] in the middle of a timestep) */
SHIM_EXPORT void cmcb200_shim_close_window(void)
{
	if (!win_open) return;
	res_flush_ec1();
	if (cmcb200_resident_end(ctx, star + 1, clus.N_MAX_NEW, binary + 1, win_n_bin, N_BIN_DIM_OPT, NULL, NULL, NULL, 0)) shim_fail("resident_end");
	win_open = 0;
}

static int res_dynamics_allowed(void)
{
	return res_allowed() && STELLAR_EVOLUTION == 0 && !THREEBODYBINARIES && !TC_POLYTROPE && !(cenma.m > 0.0 && BH_LOSS_CONE) && N_b_local == 0 &&
	       clus.N_BINARY == 0 && clus.N_MAX_NEW == End[myid] - Start[myid] + 1 && !getenv("CMCB200_NO_DEVICE_DYNAMICS");
}

void SHIM(dynamics_apply)(double dt, gsl_rng *rng)
{
	cmcb200_sigma_params sp;
	cmcb200_dynamics_params dp;
	long n = mpiEnd - mpiBegin + 1, ne, i;
	int j;
	double relbeta[4] = {PI / 2.0, PI / 4.0, PI / 8.0, PI / 16.0};
	if (!res_dynamics_allowed()) { __real_dynamics_apply(dt, rng); return; }
	res_open_window();
	/* :37 -- sigma_array is read by host code (bh_rand_walk, the binary routines): it comes back */
	memset(&sp, 0, sizeof sp);
	sp.p = AVEKERNEL; sp.n_local = n; sp.first_global = Start[myid]; sp.N_MAX = clus.N_MAX; sp.N_STAR = clus.N_STAR; sp.r_0_mave_1 = 0;
	if (cmcb200_dev_calc_sigma_r(ctx, &sp, sigma_array.r, sigma_array.sigma, NULL)) shim_fail("calc_sigma_r");
	sigma_array.n = n;
	if (tcount == 1) { /* :40-46 */
		pararootfprintf(relaxationfile, "# time");
		for (j = 0; j < 4; j++) { pararootfprintf(relaxationfile, " thetase>%g:f,q,<M>,<r>", relbeta[j]); }
		pararootfprintf(relaxationfile, "\n");
	}
	DE_bb = 0.0; /* :55-56 */
	DE_bs = 0.0;
	rootgprintf("%s(): performing interactions:\n", "dynamics_apply");
	pararootfprintf(logfile, "%s(): performing interactions:", "dynamics_apply");
	memset(&dp, 0, sizeof dp);
	dp.n_local = n; dp.first_global = Start[myid]; dp.N_MAX = clus.N_MAX; dp.N_STAR = clus.N_STAR; dp.dt = dt;
	dp.relaxation = RELAXATION; dp.ss_collision = SS_COLLISION; dp.binsingle = BINSINGLE; dp.binbin = BINBIN; dp.tidal_capture = TIDAL_CAPTURE;
	dp.tc_polytrope = TC_POLYTROPE; dp.bh_capture = BH_CAPTURE; dp.bh_loss_cone = BH_LOSS_CONE; dp.avekernel = AVEKERNEL; dp.honor_threebb = 0;
	dp.cenma_m = cenma.m; dp.gamma = GAMMA; dp.coll_factor = COLL_FACTOR; dp.tc_factor = TC_FACTOR; dp.wd_tc = WD_TC;
	dp.units_mstar = units.mstar; dp.units_l = units.l; dp.units_t = units.t;
	dp.draws = shim_dyn_draws; dp.seed = (unsigned long long) IDUM; dp.tcount = (unsigned long long) tcount;
	if (cmcb200_dev_dynamics_apply(ctx, &dp, &last_dynamics)) shim_fail("dynamics_apply");
	ne = last_dynamics.n_encounters;
	if (ne > 0) { /* :444-467: the strong encounters are the reference's own code, pair by pair in the loop's order */
		long *ek = (long *) malloc(ne * sizeof(long)), *ekp = (long *) malloc(ne * sizeof(long)), cnt = 0;
		int *et = (int *) malloc(ne * sizeof(int));
		double *ed = (double *) malloc(3 * ne * sizeof(double));
		if (cmcb200_dev_dynamics_encounters(ctx, ne, ek, ekp, et, ed, ed + ne, ed + 2 * ne, &cnt)) shim_fail("dynamics_encounters");
		for (i = 0; i < cnt; i++) {
			double v[4], vp[4], w[4], W, rcm, vcm[4];
			long k = ek[i], kp = ekp[i], n_before = clus.N_MAX_NEW, q;
			if (cmcb200_dev_get(ctx, star + k, k, NULL, 0) || cmcb200_dev_get(ctx, star + kp, kp, NULL, 0)) shim_fail("dev_get");
			calc_encounter_dyns(k, kp, v, vp, w, &W, &rcm, vcm, rng, 0); /* star[k].Y holds the pair's angle */
			print_interaction_status("SS");
			sscollision_do(k, kp, ed[i], w, W, rcm, vcm, rng);
			res_push_star(k, star[k].binind);
			res_push_star(kp, star[kp].binind);
			for (q = n_before + 1; q <= clus.N_MAX_NEW; q++) res_push_star(q, star[q].binind); /* collision products */
		}
		free(ek); free(ekp); free(et); free(ed);
	}
	{ /* :503-531 */
		long Nrel = last_dynamics.Nrel;
		double tmpTimeStart = timeStartSimple();
		double bd[3][4], bdr[3][4];
		long bl[5], blr[5];
		for (j = 0; j < 4; j++) { bd[0][j] = last_dynamics.qaverelbeta[j]; bd[1][j] = last_dynamics.maverelbeta[j]; bd[2][j] = last_dynamics.raverelbeta[j]; bl[j] = last_dynamics.Nrelbeta[j]; }
		bl[4] = Nrel;
		MPI_Reduce(bd, bdr, 12, MPI_DOUBLE, MPI_SUM, 0, MPI_COMM_WORLD);
		MPI_Reduce(bl, blr, 5, MPI_LONG, MPI_SUM, 0, MPI_COMM_WORLD);
		Nrel = blr[4];
		timeEndSimple(tmpTimeStart, &t_comm);
		pararootfprintf(relaxationfile, "%g", TotalTime);
		for (j = 0; j < 4; j++) {
			pararootfprintf(relaxationfile, " %g %g %g %g", ((double) blr[j]) / ((double) Nrel), bdr[0][j] / ((double) blr[j]), bdr[1][j] / ((double) blr[j]),
					bdr[2][j] / ((double) blr[j]));
		}
		pararootfprintf(relaxationfile, "\n");
	}
	rootgprintf("\n");
	pararootfprintf(logfile, "\n");
	/* break_wide_binaries(curr_st) (:541): without binaries it only adds up star[].E_excess, which fewbody alone sets -- nothing to do */
}

static void res_new_orbits_calculate(void)
{
	cmcb200_orbit_params p;
	double phi_rtidal, phi_zero;
	long j, n = clus.N_MAX_NEW;
	unsigned long long flagged;
	OldTidalMassLoss = TidalMassLoss_old;
	phi_rtidal = potential(Rtidal);                  /* get_positions, cmc_evolution_thr.c:963-964: the host tables are the uploaded ones */
	phi_zero = potential(0.0);
	res_params(&p);
	if (cmcb200_dev_new_orbits(ctx, &p, 1, &last_orbit)) shim_fail("new_orbits_calculate");
	res_flush_ec1();                                 /* before the removed stars' records are rewritten below: like the reference, where it ran first */
	flagged = last_orbit.n_status[CMCB200_ORBIT_MASSLESS] + last_orbit.n_status[CMCB200_ORBIT_UNBOUND] + last_orbit.n_status[CMCB200_ORBIT_TIDAL] +
		  last_orbit.n_status[CMCB200_ORBIT_NTRY];
	if (flagged) {                                   /* rare: a few stars per step */
		if (n > status_cap) { status_cap = n + n / 4 + 64; status_buf = realloc(status_buf, status_cap * sizeof(int)); }
		if (cmcb200_dev_read_status(ctx, status_buf, NULL, NULL, n)) shim_fail("dev_read_status");
		for (j = 1; j <= n; j++) {
			int st = status_buf[j - 1];
			if (st == CMCB200_ORBIT_OK || st == CMCB200_ORBIT_CIRCULAR) continue;
			/* the reference's functions read the star's record as the loop body left it (r_peri / r_apo of a tidally stripped star are logged) */
			long k;
			if (cmcb200_dev_get(ctx, star + j, j, NULL, 0)) shim_fail("dev_get");
			k = star[j].binind;                      /* destroy_obj zeroes the record: remember which binary it destroys with it */
			switch (st) {
			case CMCB200_ORBIT_MASSLESS: destroy_obj(j); break;                                           /* :816-820 */
			case CMCB200_ORBIT_UNBOUND: count_esc_bhs(j); remove_star(j, phi_rtidal, phi_zero); break;    /* :823-829 */
			case CMCB200_ORBIT_TIDAL: remove_star(j, phi_rtidal, phi_zero); break;                        /* :848-855 */
			default:                                                                                      /* :890-893 */
				eprintf("N_TRY exceeded\n");
				exit_cleanly(-1, __FUNCTION__);
			}
			res_push_star(j, k);
		}
	}
	max_r = last_orbit.max_r;
	MPI_Allreduce(MPI_IN_PLACE, &max_r, 1, MPI_DOUBLE, MPI_MAX, MPI_COMM_WORLD);
	MPI_Allreduce(MPI_IN_PLACE, &TidalMassLoss, 1, MPI_DOUBLE, MPI_SUM, MPI_COMM_WORLD);
	TidalMassLoss += TidalMassLoss_old;
}

/* src/cmc/cmc_utils.c:2217-2226 */
void SHIM(energy_conservation2)(void)
{
	cmcb200_orbit_params p;
	if (!win_open) { __real_energy_conservation2(); return; }
	res_flush_ec1();
	res_params(&p);
	if (cmcb200_dev_energy_conservation2(ctx, &p)) shim_fail("energy_conservation2");
}

/* src/cmc/cmc_evolution_thr.c:443-645: the criterion runs over the resident records (one device pass per iteration of the reference's
 * do-while); the loop around it, and for every stripped star the sums, the esc log line and destroy_obj, are host code in the reference's
 * order.  The body below restates :484-547 for one star (a maintainer would factor it out of the loop instead). */
static void res_strip_one(long i, double phi_rtidal, double phi_zero)
{
	int g_i = get_global_idx(i);
	double m = star_m[g_i], r = star_r[g_i];
	long k;
	star[i].rnew = SF_INFINITY;
	star[i].vrnew = 0.0;
	star[i].vtnew = 0.0;
	Eescaped += star[i].E * m / clus.N_STAR;
	Jescaped += star[i].J * m / clus.N_STAR;
	if (star[i].binind == 0) {
		Eintescaped += star[i].Eint;
	} else {
		Ebescaped += -(binary[star[i].binind].m1 / clus.N_STAR) * (binary[star[i].binind].m2 / clus.N_STAR) / (2.0 * binary[star[i].binind].a);
		Eintescaped += binary[star[i].binind].Eint1 + binary[star[i].binind].Eint2;
	}
	DTidalMassLoss += m / clus.N_STAR;
	Etidal += star[i].E * m / clus.N_STAR;
	parafprintf(escfile, "%ld %.8g %.8g %.8g %.8g %.8g %.8g %.8g %.8g %.8g %.8g %.8g %.8g %ld ", tcount, TotalTime, m * (units.m / clus.N_STAR) / MSUN, r,
		    star[i].vr, star[i].vt, star[i].r_peri, star[i].r_apo, Rtidal, phi_rtidal, phi_zero, star[i].E, star[i].J, star[i].id);
	k = star[i].binind;
	if (k) parafprintf(escfile, "1 %.8g %.8g %ld %ld %.8g %.8g ", binary[k].m1 * (units.m / clus.N_STAR) / MSUN, binary[k].m2 * (units.m / clus.N_STAR) / MSUN,
			   binary[k].id1, binary[k].id2, binary[k].a * units.l / AU, binary[k].e);
	else parafprintf(escfile, "0 0 0 0 0 0 0 ");
	if (k == 0)
		parafprintf(escfile, "%d na na na na na na na na na na na na na na na na na na na na na na na na na na na na na na na na na na na %g na na %g %g %g",
			    star[i].se_k, star[i].se_bhspin, star[i].se_ospin, star[i].se_scm_B, star[i].se_scm_formation);
	else
		parafprintf(escfile, "na %d %d %g %g %g %g %g %g %g %g %g %g %g %g %g %g %g %g %g %g %g %g %g %g %g %g %g %g %g %g %g %g %g %g %g na %g %g na na na",
			    binary[k].bse_kw[0], binary[k].bse_kw[1], binary[k].bse_radius[0], binary[k].bse_radius[1], binary[k].bse_tb, binary[k].bse_lum[0],
			    binary[k].bse_lum[1], binary[k].bse_massc[0], binary[k].bse_massc[1], binary[k].bse_radc[0], binary[k].bse_radc[1], binary[k].bse_menv[0],
			    binary[k].bse_menv[1], binary[k].bse_renv[0], binary[k].bse_renv[1], binary[k].bse_tms[0], binary[k].bse_tms[1], binary[k].bse_bcm_dmdt[0],
			    binary[k].bse_bcm_dmdt[1], binary[k].bse_bcm_radrol[0], binary[k].bse_bcm_radrol[1], binary[k].bse_ospin[0], binary[k].bse_ospin[1],
			    binary[k].bse_bcm_B[0], binary[k].bse_bcm_B[1], binary[k].bse_bcm_formation[0], binary[k].bse_bcm_formation[1], binary[k].bse_bacc[0],
			    binary[k].bse_bacc[1], binary[k].bse_tacc[0], binary[k].bse_tacc[1], binary[k].bse_mass0[0], binary[k].bse_mass0[1], binary[k].bse_epoch[0],
			    binary[k].bse_epoch[1], binary[k].bse_bhspin[0], binary[k].bse_bhspin[1]);
	parafprintf(escfile, "\n");
	destroy_obj(i);
}

void SHIM(tidally_strip_stars)(void)
{
	static long *idx = NULL, cap = 0;
	cmcb200_orbit_params p;
	double phi_rtidal, phi_zero, thr;
	long j = 0, n, q;
	int rc;
	if (!win_open) { __real_tidally_strip_stars(); return; }
	res_flush_ec1();
	if (!idx) { cap = 4096; idx = malloc(cap * sizeof(long)); }
	Etidal = 0.0;
	DTidalMassLoss = TidalMassLoss - OldTidalMassLoss;
	rootgprintf("tidally_strip_stars(): iteration %ld: OldTidalMassLoss=%.6g DTidalMassLoss=%.6g\n", j, OldTidalMassLoss, DTidalMassLoss);
	pararootfprintf(logfile, "tidally_strip_stars(): iteration %ld: OldTidalMassLoss=%.6g DTidalMassLoss=%.6g\n", j, OldTidalMassLoss, DTidalMassLoss);
	do {
		Rtidal = orbit_r * pow(Mtotal - (TidalMassLoss - OldTidalMassLoss), 1.0 / 3.0);
		phi_rtidal = potential(Rtidal);
		phi_zero = potential(0.0);
		DTidalMassLoss = 0.0;
		thr = TIDAL_TREATMENT == 0 ? Rtidal : (1.5 - 3.0 * pow(log(GAMMA * ((double) clus.N_STAR)) / ((double) clus.N_STAR), 0.25)) * phi_rtidal;
		res_params(&p);
		for (;;) { /* the criterion over all resident records; nothing is marked when the list does not fit */
			rc = cmcb200_dev_tidally_strip_select(ctx, &p, (int) TIDAL_TREATMENT, thr, idx, cap, &n, NULL, NULL, NULL);
			if (rc == -3 && n > cap) { cap = n + 1024; idx = realloc(idx, cap * sizeof(long)); continue; }
			if (rc) shim_fail("tidally_strip_stars");
			break;
		}
		for (q = 0; q < n; q++) { /* star order: the sums see the reference's additions in the reference's order */
			long i = idx[q], k;
			if (cmcb200_dev_get(ctx, star + i, i, NULL, 0)) shim_fail("dev_get");
			k = star[i].binind;
			res_strip_one(i, phi_rtidal, phi_zero);
			res_push_star(i, k);
		}
		j++;
		MPI_Allreduce(MPI_IN_PLACE, &DTidalMassLoss, 1, MPI_DOUBLE, MPI_SUM, MPI_COMM_WORLD);
		TidalMassLoss += DTidalMassLoss;
		rootgprintf("tidally_strip_stars(): iteration %ld: TidalMassLoss=%.6g DTidalMassLoss=%.6g\n", j, TidalMassLoss, DTidalMassLoss);
		pararootfprintf(logfile, "tidally_strip_stars(): iteration %ld: TidalMassLoss=%.6g DTidalMassLoss=%.6g\n", j, TidalMassLoss, DTidalMassLoss);
	} while (DTidalMassLoss > 0);
	{
		double buf_reduce[5] = {Eescaped, Jescaped, Eintescaped, Ebescaped, Etidal};
		MPI_Allreduce(MPI_IN_PLACE, buf_reduce, 5, MPI_DOUBLE, MPI_SUM, MPI_COMM_WORLD);
		Eescaped = buf_reduce[0]; Jescaped = buf_reduce[1]; Eintescaped = buf_reduce[2]; Ebescaped = buf_reduce[3]; Etidal = buf_reduce[4];
	}
}

/* src/cmc/cmc_utils.c:368-391 */
void SHIM(ComputeIntermediateEnergy)(void)
{
	cmcb200_orbit_params p;
	if (!win_open) { __real_ComputeIntermediateEnergy(); return; }
	res_flush_ec1();
	res_params(&p);
	if (cmcb200_dev_compute_intermediate_energy(ctx, &p)) shim_fail("ComputeIntermediateEnergy");
}

static void res_qsorts_new(void)
{
	long n = clus.N_MAX_NEW, nb = 0;
	res_flush_ec1();
	if (cmcb200_dev_qsorts_new(ctx, &n, win_n_bin, &nb)) shim_fail("qsorts_new");
	clus.N_MAX = n;
	clus.N_MAX_NEW = n;
	N_b_local = nb;
	win_n_bin = nb;
	findLimits(clus.N_MAX, MIN_CHUNK_SIZE);
}

static void res_calc_potential_new(void)
{
	long nmax = clus.N_MAX;
	if (cmcb200_dev_calc_potential_new(ctx, &nmax, clus.N_STAR, &cenma.m, &cenma.m_new, orbit_r, &Mtotal, &Rtidal)) shim_fail("calc_potential_new");
	clus.N_MAX = nmax;
	/* potential() / calc_orbit_new() / FindZero_r() have host callers: the tables come back now, the records at the end of the window */
	if (cmcb200_download(ctx, NULL, 0, NULL, 0, star_r, star_m, star_phi, clus.N_MAX + 2)) shim_fail("download(tables)");
	mpiFindIndicesCustom(clus.N_MAX, MIN_CHUNK_SIZE, myid, &mpiBegin, &mpiEnd);
}

/* src/cmc/cmc_utils.c:254-363 */
void SHIM(energy_conservation3)(void)
{
	cmcb200_orbit_params p;
	double Eexcess = 0.0, Eexcess_prev = 0.0, Eexcess_check = 0.0;
	MPI_Status stat;
	if (!win_open) { __real_energy_conservation3(); return; }
	res_params(&p);
	if (cmcb200_dev_set_velocities3(ctx, &p, &Eexcess)) shim_fail("set_velocities3");
	do { /* :315-347: what a rank could not absorb travels to the next one */
		if (myid != procs - 1) MPI_Send(&Eexcess, 1, MPI_DOUBLE, myid + 1, 0, MPI_COMM_WORLD);
		if (myid != 0) MPI_Recv(&Eexcess_prev, 1, MPI_DOUBLE, myid - 1, 0, MPI_COMM_WORLD, &stat);
		Eexcess = Eexcess_prev;
		if (Eexcess > 0 && cmcb200_dev_set_velocities3_pass(ctx, &p, Eexcess, &Eexcess)) shim_fail("set_velocities3 (second loop)");
		MPI_Allreduce(&Eexcess, &Eexcess_check, 1, MPI_DOUBLE, MPI_SUM, MPI_COMM_WORLD);
	} while (Eexcess_check > 0);
	Eoops += -Eexcess * madhoc;
}

/* src/cmc/cmc_utils.c:872-985 */
void SHIM(comp_mass_percent)(void)
{
	cmcb200_mass_pc_params p;
	cmcb200_mass_pc_result r;
	if (!win_open) { __real_comp_mass_percent(); return; }
	memset(&p, 0, sizeof p); memset(&r, 0, sizeof r);
	p.count = MASS_PC_COUNT; p.mass_pc = mass_pc; p.bh_include = MASS_PC_BH_INCLUDE; p.minimum_r = MINIMUM_R;
	p.initial_total_mass = initial_total_mass; p.Mtotal = Mtotal; p.cenma_m = cenma.m;
	p.N_MAX = clus.N_MAX; p.N_STAR = clus.N_STAR; p.n_local = clus.N_MAX_NEW; p.first_global = Start[myid];
	r.mass_r = mass_r; r.ave_mass_r = ave_mass_r; r.no_star_r = no_star_r; r.densities_r = densities_r;
	r.ke_rad_r = ke_rad_r; r.ke_tan_r = ke_tan_r; r.v2_rad_r = v2_rad_r; r.v2_tan_r = v2_tan_r;
	if (cmcb200_dev_comp_mass_percent(ctx, &p, &r)) shim_fail("comp_mass_percent");
}

/* src/cmc/cmc_utils.c:604-656.  Closes the window. */
void SHIM(ComputeEnergy)(void)
{
	double buf_reduce[5], s[4];
	long n = mpiEnd - mpiBegin + 1;
	if (!win_open) { __real_ComputeEnergy(); return; }
	if (cmcb200_dev_compute_energy(ctx, n, Start[myid])) shim_fail("ComputeEnergy");
	if (cmcb200_dev_compute_energy_totals(ctx, n, Start[myid], clus.N_MAX, clus.N_STAR, cenma.m, s)) shim_fail("ComputeEnergy (sums)");
	buf_reduce[1] = s[0]; buf_reduce[2] = s[1]; buf_reduce[3] = s[2]; buf_reduce[4] = s[3];
	buf_reduce[2] *= 0.5;
	buf_reduce[0] = buf_reduce[1] + buf_reduce[2] + buf_reduce[3] + buf_reduce[4];
	MPI_Allreduce(MPI_IN_PLACE, buf_reduce, 5, MPI_DOUBLE, MPI_SUM, MPI_COMM_WORLD);
	Etotal.tot = buf_reduce[0]; Etotal.K = buf_reduce[1]; Etotal.P = buf_reduce[2]; Etotal.Eint = buf_reduce[3]; Etotal.Eb = buf_reduce[4];
	cenma.E += cenma.E_new;
	cenma.E_new = 0.0;
	Etotal.tot += cenma.E + Eescaped + Ebescaped + Eintescaped;
	/* everything the host reads next (reset_interaction_flags, update_vars, the output routines, dynamics_apply of the next step) */
	if (cmcb200_resident_end(ctx, star + 1, clus.N_MAX_NEW, binary + 1, win_n_bin, N_BIN_DIM_OPT, NULL, NULL, NULL, 0)) shim_fail("resident_end");
	win_open = 0;
}
static void shim_reset_resident(void) { shim_dyn_draws = NULL; win_open = 0; ec1_pending = 0; dirty_words = ~0ull; }
#else
static void shim_reset_resident(void) {}
#endif /* CMCB200_SHIM_RESIDENT */

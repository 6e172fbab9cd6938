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
 * Single rank per GPU.  procs > 1 goes through the radius-sharded entry points (cmcb200_shard_*). */
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
SHIM_EXPORT void cmcb200_shim_set_draws(const double *draws, const long *draw_off) { shim_draws = draws; shim_draw_off = draw_off; }
SHIM_EXPORT const cmcb200_orbit_result *cmcb200_shim_last_orbit(void) { return &last_orbit; }
SHIM_EXPORT void cmcb200_shim_shutdown(void)
{
	shim_unpin();
	if (ctx && procs > 1) { cmcb200_sync(ctx); MPI_Barrier(MPI_COMM_WORLD); cmcb200_shard_detach(ctx); MPI_Barrier(MPI_COMM_WORLD); } /* nobody unmaps a buffer a peer still uses */
	if (ctx) cmcb200_destroy(ctx);
	ctx = NULL; free(status_buf); status_buf = NULL; status_cap = 0;
}

/* src/cmc/cmc_utils.c:2194-2211 */
void SHIM(new_orbits_calculate)(void)
{
	cmcb200_orbit_params p;
	double phi_rtidal, phi_zero;
	long j, n = clus.N_MAX_NEW, n_old = End[myid] - Start[myid] + 1;
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
	shim_init();
	if (cmcb200_calc_potential_new(ctx, star_r, star_m, star_phi, &nmax, clus.N_STAR, &cenma.m, &cenma.m_new, orbit_r, &Mtotal, &Rtidal))
		shim_fail("calc_potential_new");
	clus.N_MAX = nmax;
	mpiFindIndicesCustom(clus.N_MAX, MIN_CHUNK_SIZE, myid, &mpiBegin, &mpiEnd);
}

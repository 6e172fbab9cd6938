/* cmcb200 -- B200-native drop-in for the per-timestep data-parallel core of ClusterMonteCarlo/CMC-COSMIC:
 * radius sort, potential rebuild, per-star new-orbit step.
 *
 * C ABI: plain pointers and sizes, no C++/torch types.  Paths cite the reference tree (ClusterMonteCarlo/CMC-COSMIC).
 *
 * The reference has no plugin API: the boundary is four `void f(void)` functions that main() calls on process
 * globals (src/cmc/cmc.c:311,335,339,344):
 *     void new_orbits_calculate(void)   include/cmc/cmc.h:1801   -> cmcb200_new_orbits_calculate
 *     void qsorts_new(void)             include/cmc/cmc.h:1922   -> cmcb200_qsorts_new
 *     void post_sort_comm(void)         include/cmc/cmc.h:1805   -> (fused into cmcb200_qsorts_new: star_r / star_m out)
 *     void calc_potential_new(void)     include/cmc/cmc.h:1791   -> cmcb200_calc_potential_new
 * Each entry point below takes what the corresponding function reads from include/cmc/cmc_vars.h as explicit
 * arguments and returns what it writes.  INTEGRATION.md shows the shim that defines the four functions on the
 * reference's globals and forwards here.
 *
 * Conventions
 *  - All functions return 0 on success, non-zero on failure; cmcb200_last_error() describes the failure.  The
 *    reference's convention (eprintf + exit_cleanly -> MPI_Abort, src/cmc/cmc_utils.c:208-220) is applied by the shim.
 *  - "host" pointers are ordinary host memory owned by the caller (pinned memory makes the copies faster but is not
 *    required).  The *_dev entry points work on the context's device-resident mirrors and move no bulk data.
 *  - star_t / binary_t keep the reference's layout byte for byte (include/cmc/cmc.h:409-601, :231-390;
 *    sizeof = 424 / 464); cmcb200_layout() reports the constants the kernels were built with.
 *  - Index conventions are the reference's: star[0] / star_r[0] is the sentinel, live stars are 1..N.
 *  - One context drives one GPU from one host thread (the reference runs one thread per MPI rank).
 */
#ifndef CMCB200_H
#define CMCB200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cmcb200_ctx cmcb200_ctx;

#define CMCB200_SF_INFINITY 1.0e10 /* include/cmc/cmc.h:58 */
#define CMCB200_ZERO 1.0e-20       /* include/cmc/cmc.h:60 */

/* per-star outcome of the orbit step: the branches of get_positions_loop (src/cmc/cmc_evolution_thr.c:802-933) */
enum {
	CMCB200_ORBIT_OK = 0,       /* new radius sampled */
	CMCB200_ORBIT_MASSLESS = 1, /* star_m < ZERO: the shim must call destroy_obj(j)                              :816-820 */
	CMCB200_ORBIT_UNBOUND = 2,  /* E >= 0: the shim must call count_esc_bhs(j); remove_star(j, ...)              :823-829 */
	CMCB200_ORBIT_CIRCULAR = 3, /* circular_flag: rnew/vrnew/vtnew = r/vr/vt                                     :835-839 */
	CMCB200_ORBIT_TIDAL = 4,    /* rmax >= Rtidal: r_peri/r_apo set; the shim must call remove_star(j, ...)      :848-855 */
	CMCB200_ORBIT_NTRY = 5      /* N_TRY exceeded: fatal in the reference (exit_cleanly)                         :890-893 */
};

/* ---- life cycle -------------------------------------------------------------------------------------------- */

/* device: CUDA ordinal.  n_star_dim / n_bin_dim: capacity in records of star[] / binary[] (N_STAR_DIM_OPT, N_BIN_DIM_OPT,
 * src/cmc/cmc_io.c:2092-2101) -- also the capacity of star_r / star_m / star_phi (src/cmc/cmc_io.c:3431-3437). */
int cmcb200_create(cmcb200_ctx **ctx, int device, long n_star_dim, long n_bin_dim);
/* the same with a separate capacity for the replicated tables: with the star array sharded over several ranks n_star_dim is this rank's
 * share of star[] (N_STAR_DIM_OPT) while star_r / star_m / star_phi hold the whole cluster (N_STAR_DIM, src/cmc/cmc_io.c:3431-3437) */
int cmcb200_create_sharded(cmcb200_ctx **ctx, int device, long n_star_dim, long n_table_dim, long n_bin_dim);
void cmcb200_destroy(cmcb200_ctx *ctx);
const char *cmcb200_last_error(void);
/* run the context's kernels on an existing CUDA stream (cudaStream_t passed as void*); default: a private stream */
int cmcb200_set_stream(cmcb200_ctx *ctx, void *cuda_stream);
/* sizeof(star_t), sizeof(binary_t) and the byte offsets of r, vr, vt, m, E, J, rnew, vrnew, vtnew, rOld, X, r_peri, r_apo,
 * binind the library was built for (out[0..15]); the shim static-asserts them against the reference headers */
void cmcb200_layout(long out[16]);

/* ---- partition arithmetic (integer, bit-exact) --------------------------------------------------------------- */

/* findLimits (src/cmc/cmc_utils.c:2319-2333) == mpiFindIndicesCustom per rank (src/cmc/cmc_mpi.c:100-114):
 * Start[i], End[i] (1-based, inclusive) of rank i when N stars are dealt in blocks of blk to procs ranks. */
void cmcb200_find_limits(long N, int blk, int procs, int *Start, int *End);

/* ---- calc_potential_new -> potential_calculate (src/cmc/cmc_utils.c:2006-2012, 675-736) ----------------------- */

/* Host-buffer form.  star_r, star_m: the replicated tables ([0] sentinel .. [*N_MAX + 1]); star_phi is written at
 * [0 .. new N_MAX + 1]; star_r[new N_MAX + 1] = SF_INFINITY like the reference (:713).
 * *N_MAX: clus.N_MAX in/out.  N_STAR: clus.N_STAR.  cenma_m / cenma_m_new: cenma.m, cenma.m_new in/out (:699-700).
 * orbit_r -> *Rtidal = orbit_r * Mtotal^(1/3) (:707).  Results carry the reference's sequential FP64 roundings bit for bit. */
int cmcb200_calc_potential_new(cmcb200_ctx *ctx, double *star_r, const double *star_m, double *star_phi, long *N_MAX, long N_STAR,
			       double *cenma_m, double *cenma_m_new, double orbit_r, double *Mtotal, double *Rtidal);

/* ---- qsorts_new + post_sort_comm (src/cmc/cmc_utils.c:2232-2288; src/cmc/cmc_sort.c:375-791), one rank ------------ */

/* Host-buffer form.  star1 = &star[1], binary1 = &binary[1] (what the reference passes to sample_sort).
 * *n_local: clus.N_MAX_NEW in; out = stars kept (trailing r == SF_INFINITY trimmed, cmc_sort.c:343-354) = new clus.N_MAX
 * and clus.N_MAX_NEW on one rank.  Records [0, n_in) are permuted into ascending r (ties keep input order); binaries are
 * re-packed in star order, binind renumbered 1,2,.. and binary1[*n_bin .. n_bin_dim-2] zeroed (cmc_sort.c:541,571-579).
 * star_r / star_m (may be NULL): receive star[i].r / star[i].m at [1 .. *n_local] (post_sort_comm). */
int cmcb200_qsorts_new(cmcb200_ctx *ctx, void *star1, long *n_local, void *binary1, long n_bin_dim, long *n_bin,
		       double *star_r, double *star_m);

/* ---- new_orbits_calculate -> get_positions (src/cmc/cmc_utils.c:2194-2211; src/cmc/cmc_evolution_thr.c:784-1023) --- */

typedef struct {
	long n_local;       /* clus.N_MAX_NEW */
	long n_old_local;   /* End[myid] - Start[myid] + 1; local stars beyond it were created this step (get_global_idx) */
	long first_global;  /* Start[myid] */
	long N_MAX;         /* clus.N_MAX */
	long N_STAR;        /* clus.N_STAR */
	long n_table;       /* entries of star_r/star_m/star_phi to make visible (>= N_MAX + 2 + number of new stars) */
	double cenma_m;     /* cenma.m */
	double Rtidal;
	/* uniform draws.  draws != NULL: a recorded stream in CSR form -- star j (0-based local) consumes
	 * draws[draw_off[j] .. draw_off[j+1]) in order (X, then the g0 draw, per try; one more for the sign of vr):
	 * "the same uniform random draws fed in as an input array".  draws == NULL: a counter-based Philox4x32-10 stream per
	 * star keyed by (seed, tcount, global index); statistically equivalent to the reference's per-rank taus113 stream. */
	const double *draws;
	const long *draw_off;
	unsigned long long seed, tcount;
	/* resident pipeline only: != 0 computes E = star_phi[g] + (vr^2 + vt^2)/2, J = star_r[g] vt of every star on the fly and stores them
	 * -- the first loop of ComputeEnergy (src/cmc/cmc_utils.c:612-616) fused into the orbit step instead of a pass of its own */
	int fuse_compute_energy;
} cmcb200_orbit_params;

typedef struct {
	double max_r;                 /* max over sampled stars of the new radius (get_positions' return value) */
	unsigned long long n_draws;   /* uniforms consumed */
	unsigned long long n_exact;   /* Q evaluations that needed the exact 80-bit replay */
	unsigned long long n_brent;   /* find_root_vr fallbacks */
	unsigned long long n_status[6];
	unsigned long long n_deferred; /* stars whose own Q was <= 0: their O(N) linear scan (src/cmc/cmc_orbit.c:46-57) ran as a grid-wide find-first */
} cmcb200_orbit_result;

/* Host-buffer form.  star1 = &star[1]: reads .E .J .vr .vt (.rnew for new stars); writes .rnew .vrnew .vtnew .X .r_peri
 * .r_apo exactly where the reference's loop body does.  Only bytes 0..127 of each record are copied in and bytes 64..127 copied
 * back (strided DMA, pipelined over chunks of the star range with the kernels).  status (may be NULL) receives one CMCB200_ORBIT_* per star; the
 * shim replays the host-side side effects (remove_star / destroy_obj / count_esc_bhs, esc log) for the flagged stars in
 * star order.  kmin / kmax (may be NULL): the bracket indices of calc_orbit_new, for parity tests. */
int cmcb200_new_orbits_calculate(cmcb200_ctx *ctx, void *star1, const double *star_r, const double *star_m, const double *star_phi,
				 const cmcb200_orbit_params *p, int *status, long *kmin, long *kmax, cmcb200_orbit_result *res);

/* ---- host memory ---------------------------------------------------------------------------------------------------- */

/* Pin a host array the host-buffer entry points are called with (star[], binary[], star_r/m/phi: the reference callocs / mallocs them
 * once, src/cmc/cmc_io.c:2099-2101, 3434-3436, so the shim registers them once at start-up).  Optional: pageable memory works, but only
 * pinned memory is copied by DMA at PCIe rate and asynchronously, which is what lets cmcb200_new_orbits_calculate overlap its copy-in,
 * kernels and copy-out. */
int cmcb200_host_register(void *ptr, size_t bytes);
int cmcb200_host_unregister(void *ptr);
/* bytes the host-buffer entry points have moved host->device / device->host since creation (or the last reset) */
int cmcb200_transfer_bytes(cmcb200_ctx *ctx, long *h2d, long *d2h, int reset);

/* ---- device-resident pipeline ------------------------------------------------------------------------------------ */
/* The star array lives in HBM across timesteps; the host touches O(encounters) records through the get/put calls. */

int cmcb200_upload(cmcb200_ctx *ctx, const void *star1, long n_local, const void *binary1, long n_bin,
		   const double *star_r, const double *star_m, const double *star_phi, long n_table);
int cmcb200_download(cmcb200_ctx *ctx, void *star1, long n_local, void *binary1, long n_bin,
		     double *star_r, double *star_m, double *star_phi, long n_table);

/* Resident timestep on HOST arrays (the residency protocol of the drop-in shim): the records cross the host link once each
 * way per timestep instead of once per entry point.
 *   cmcb200_resident_begin   starts the upload of the caller's arrays and returns at once: tables first, then the records in row chunks, then
 *                            the binaries, on the context's copy stream.  The cmcb200_dev_* calls that follow are ordered behind it ON THE
 *                            DEVICE; cmcb200_dev_new_orbits runs chunk by chunk behind the arriving records, so the orbit step hides under
 *                            the upload.  The caller must not write the arrays until a call that returns results computed from all of them
 *                            has returned (any cmcb200_dev_* call that synchronises: the orbit step, the sort, the potential).
 *                            dirty_words: bit w set = the 8-byte word w of the records (0 <= w < 53) may differ from the copy the last
 *                            cmcb200_resident_end left on the device; ~0ull = everything (always used when there is no such copy or n_local
 *                            changed).  Clean words are not sent -- the caller's promise, e.g. a run without stellar evolution never
 *                            writes the se_* words between two timesteps.  Runs of dirty words travel as strided (2-D) DMA copies.
 *   cmcb200_resident_end     downloads records, binaries (the tail of binary1 up to n_bin_dim - 1 is zeroed like cmc_sort.c:541 does) and
 *                            tables, and waits: the host arrays are coherent with the device again.
 * Any pointer may be NULL (not moved). */
int cmcb200_resident_begin(cmcb200_ctx *ctx, const void *star1, long n_local, const void *binary1, long n_bin,
			   const double *star_r, const double *star_m, const double *star_phi, long n_table, unsigned long long dirty_words);
/* The same for a radius-sharded context (cmcb200_shard_attach done).  The host tables are replicated on every rank, so each rank sends only its
 * own slice [slice_first, slice_first + slice_n) of star_r / star_m / star_phi (and the sentinel [0]) over its host link and the ranks hand the
 * slices to each other over NVLink (peer stores + a device-side barrier).  The caller gives rank q the slice Start[q] .. End[q] and extends
 * the LAST rank's slice to n_table - 1, so that the slices cover [1, n_table). */
int cmcb200_resident_begin_sharded(cmcb200_ctx *ctx, const void *star1, long n_local, const void *binary1, long n_bin, const double *star_r,
				   const double *star_m, const double *star_phi, long n_table, unsigned long long dirty_words, long slice_first, long slice_n);
int cmcb200_resident_end(cmcb200_ctx *ctx, void *star1, long n_local, void *binary1, long n_bin, long n_bin_dim,
			 double *star_r, double *star_m, double *star_phi, long n_table);

/* Single records between two bulk transfers -- what the reference's host code needs when it handles one star of a resident window
 * (remove_star / destroy_obj for an escaper: reads the record, zeroes it, sets star_m[g] = DBL_MIN, destroys its binary).
 *   cmcb200_dev_get   star_rec <- resident record j (1-based local index); binary_rec (may be NULL) <- resident binary slot b
 *   cmcb200_dev_put   the other way; star_m_g (may be NULL): the new value of the resident table entry star_m[g]
 * Both wait for the copy. */
int cmcb200_dev_get(cmcb200_ctx *ctx, void *star_rec, long j, void *binary_rec, long b);
int cmcb200_dev_put(cmcb200_ctx *ctx, const void *star_rec, long j, const void *binary_rec, long b, const double *star_m_g, long g);

/* orbit step on the resident arrays (same semantics as cmcb200_new_orbits_calculate; star_r/m/phi = resident tables) */
int cmcb200_dev_new_orbits(cmcb200_ctx *ctx, const cmcb200_orbit_params *p, int want_status, cmcb200_orbit_result *res);
/* the orbit step queued without waiting for it; cmcb200_dev_orbit_result waits for the stream and returns its statistics */
int cmcb200_dev_new_orbits_async(cmcb200_ctx *ctx, const cmcb200_orbit_params *p, int want_status);
int cmcb200_dev_orbit_result(cmcb200_ctx *ctx, const cmcb200_orbit_params *p, cmcb200_orbit_result *res);
/* the transfer half of ComputeIntermediateEnergy (src/cmc/cmc_utils.c:380-390): rOld <- star_r, r <- rnew, m <- star_m,
 * vr <- vrnew, vt <- vtnew for local stars 1..n_local -- this is what creates the sort key */
int cmcb200_dev_transfer_new_positions(cmcb200_ctx *ctx, long n_local, long n_old_local, long first_global, long N_MAX);
/* The two loops of the energy scheme between the orbit step and the sort that look the potential up at the new radii (one
 * potential() per star, src/cmc/cmc_utils.c:112-165, through the same coarse-to-fine search as the orbit step's rejection rounds).
 * p: the partition and table fields of the orbit step (n_local, n_old_local, first_global, N_MAX, N_STAR, n_table, cenma_m).
 *   energy_conservation2 (src/cmc/cmc_utils.c:2217-2226):  star[i].Uoldrnew = potential(star[i].rnew) + MPI_PHI_S(star[i].rnew, g_i)
 *   ComputeIntermediateEnergy (:368-391):                   star[j].EI = sqr(vr) + sqr(vt) + star_phi[g_j] - potential(star[j].rnew) where
 *                                                           rnew < 1e6, then the transfer (cmcb200_dev_transfer_new_positions)
 * A lookup whose bracket check fails (fatal in the reference, :144-152) makes the call return non-zero. */
int cmcb200_dev_energy_conservation2(cmcb200_ctx *ctx, const cmcb200_orbit_params *p);
int cmcb200_dev_compute_intermediate_energy(cmcb200_ctx *ctx, const cmcb200_orbit_params *p);
/* Both of the above from ONE lookup per star (they look up the same potential(rnew)): Uoldrnew and EI, WITHOUT the transfer -- follow with
 * cmcb200_dev_transfer_and_sort, which applies it while the sort moves the records.  Equivalent to the reference's sequence
 * energy_conservation2(); ComputeIntermediateEnergy(); qsorts_new(); (src/cmc/cmc.c:317-335) when nothing in between touches the stars,
 * i.e. TIDALLY_STRIP_STARS == 0 (:321-322). */
int cmcb200_dev_energy_new_radii(cmcb200_ctx *ctx, const cmcb200_orbit_params *p);
/* The rest of the energy scheme (Stodolkiewicz), on the resident arrays; p as above, the records are the sorted ones after the potential rebuild.
 *   energy_conservation1 (src/cmc/cmc_utils.c:2153-2174):  star[i].vrold/.vtold <- .vr/.vt;  .Uoldrold = star_phi[g_i] + MPI_PHI_S(star_r[g_i], g_i)
 *   set_velocities3 (:254-352, called through energy_conservation3 :360-363), first loop (:268-311): for every star with interacted == 0 the
 *     predicted squared speed vnew2 (one potential(rOld) lookup); vnew2 > 0: (vr, vt) rescaled to it and copied to (vrnew, vtnew); vnew2 <= 0:
 *     the star's deficit joins Eexcess, which the following stars absorb (at most 0.8 of their kinetic energy each) -- a carry the reference
 *     runs serially from star to star, reproduced here in star order (csrc/velocities_core.h).  *Eexcess_left = what the rank is left with.
 *   On ONE rank that is the whole function: the reference resets Eexcess to the value received from the previous rank, 0 (:322), before the
 *   second loop, so the leftover is dropped and Eoops does not change.
 *   On P ranks the caller carries the reference's message pattern (:315-347): repeat { send Eexcess to rank+1; Eexcess <- received (0 on rank 0);
 *   cmcb200_dev_set_velocities3_pass(Eexcess) = one pass of the second loop (:323-343) over this rank's stars, in star order; all-reduce the
 *   leftovers } while their sum > 0, then Eoops += -Eexcess * madhoc (:351).
 * A potential() lookup whose bracket check fails (fatal in the reference, :144-152) makes the call return non-zero. */
int cmcb200_dev_energy_conservation1(cmcb200_ctx *ctx, const cmcb200_orbit_params *p);
int cmcb200_dev_set_velocities3(cmcb200_ctx *ctx, const cmcb200_orbit_params *p, double *Eexcess_left);
int cmcb200_dev_set_velocities3_pass(cmcb200_ctx *ctx, const cmcb200_orbit_params *p, double Eexcess_in, double *Eexcess_left);
/* potential() (src/cmc/cmc_utils.c:112-165) at n radii on the resident tables: out[i] = potential(r[i]), bit for bit (r, out: host arrays) */
int cmcb200_dev_potential(cmcb200_ctx *ctx, const cmcb200_orbit_params *p, const double *r, long n, double *out);
/* One pass of the loop of tidally_strip_stars (src/cmc/cmc_evolution_thr.c:443-645; between energy_conservation2 and
 * ComputeIntermediateEnergy when TIDALLY_STRIP_STARS > 0, src/cmc/cmc.c:321-322) over the resident records: a star with rnew < 1e6 is
 * stripped when r_apo > threshold = Rtidal (tidal_treatment 0, :479) or E > threshold = gierszalpha * phi_rtidal (tidal_treatment 1, :555):
 * rnew = SF_INFINITY, vrnew = vtnew = 0.  idx_out[cap] receives the 1-based local indices in star order, *count their number; star_out
 * (may be NULL) the stripped stars' 424-byte records and binary_out (may be NULL) the 464-byte record of each one that is a binary, both
 * at position k of the list -- what the caller needs for the reference's per-star bookkeeping (:484-496: Eescaped, Jescaped, Eintescaped,
 * Ebescaped, DTidalMassLoss, Etidal, the esc log line), which it does in that order, recomputes Rtidal from the mass lost and calls again
 * while a pass strips anything (:624).  m_out[cap] (may be NULL): star_m[g] of each listed star (:471) from the resident table -- the mass
 * the sums use; the caller's host copy of star_m may predate the last sort.  The pass is two-phase: the criterion is evaluated first and the
 * records are only marked once the list is known to fit.  More than cap hits in one pass: return -3 with *count = the number of hits and NO
 * record changed, so the caller can come back with a longer list. */
int cmcb200_dev_tidally_strip_select(cmcb200_ctx *ctx, const cmcb200_orbit_params *p, int tidal_treatment, double threshold, long *idx_out,
				     long cap, long *count, void *star_out, void *binary_out, double *m_out);
/* qsorts_new + post_sort_comm on the resident arrays; *n_local in/out, *n_bin out */
int cmcb200_dev_qsorts_new(cmcb200_ctx *ctx, long *n_local, long n_bin_clear, long *n_bin);
/* cmcb200_dev_transfer_new_positions + cmcb200_dev_qsorts_new in one pass over the records: the sort key is rnew and the transfer is
 * applied to each 424-byte record while the gather moves it (n_old_local / first_global / N_MAX: those of the orbit step) */
int cmcb200_dev_transfer_and_sort(cmcb200_ctx *ctx, long *n_local, long n_old_local, long first_global, long N_MAX, long n_bin_clear, long *n_bin);
/* The three calls above in one: cmcb200_dev_new_orbits -> cmcb200_dev_transfer_and_sort -> cmcb200_dev_calc_potential_new (src/cmc/cmc.c:311-344
 * without the energy scheme), everything queued on the context's stream and ONE host wait at the end instead of three -- what a resident
 * timestep costs at N <= 1e5, where the waits and the launch gaps outweigh the kernels.  Because the orbit step and the sort are queued by the
 * same call, the orbit kernels also leave the sort's keys (rnew as an order-preserving integer) in the sort's key buffer, and the sort
 * skips its pass over the records that would collect them.  Bit-identical with the three calls. */
typedef struct {
	long n_local, n_bin;  /* after the sort (stars at r = SF_INFINITY trimmed, binaries re-packed) */
	long N_MAX;           /* and what calc_potential_new leaves */
	double cenma_m, cenma_m_new, Mtotal, Rtidal;
} cmcb200_step_result;
int cmcb200_dev_step(cmcb200_ctx *ctx, const cmcb200_orbit_params *p, long n_bin_clear, double cenma_m_new, double orbit_r,
		     cmcb200_orbit_result *orbit, cmcb200_step_result *out);
/* calc_potential_new on the resident tables */
int cmcb200_dev_calc_potential_new(cmcb200_ctx *ctx, long *N_MAX, long N_STAR, double *cenma_m, double *cenma_m_new, double orbit_r,
				   double *Mtotal, double *Rtidal);
/* E, J of every star from the new potential: the first loop of ComputeEnergy (src/cmc/cmc_utils.c:612-616) */
int cmcb200_dev_compute_energy(cmcb200_ctx *ctx, long n_local, long first_global);
/* The second loop of ComputeEnergy (:619-636): this rank's sequentially summed energies, bit for bit -- out[0] kinetic (buf_reduce[1]),
 * out[1] potential before the factor 0.5 (buf_reduce[2]), out[2] internal (buf_reduce[3]), out[3] binary binding (buf_reduce[4]).  The
 * caller finishes like the reference (:638-655): out[1] *= 0.5, tot = the four added in that order, MPI_Allreduce, + cenma.E + the
 * escaped energies. */
int cmcb200_dev_compute_energy_totals(cmcb200_ctx *ctx, long n_local, long first_global, long N_MAX, long N_STAR, double cenma_m, double out[4]);
/* per-star outputs of the last dev orbit step (status needs want_status) */
int cmcb200_dev_read_status(cmcb200_ctx *ctx, int *status, long *kmin, long *kmax, long n_local);
/* wait for everything queued on the context's stream */
int cmcb200_sync(cmcb200_ctx *ctx);

/* ---- comp_mass_percent (src/cmc/cmc_utils.c:872-985): Lagrange radii on the resident arrays -------------------------------------- */

typedef struct {
	long count;                /* MASS_PC_COUNT (<= 256) */
	const double *mass_pc;     /* [count] mass fractions, as in the reference's mass_pc[] */
	int bh_include;            /* MASS_PC_BH_INCLUDE */
	double minimum_r;          /* MINIMUM_R */
	double initial_total_mass, Mtotal, cenma_m;
	long N_MAX, N_STAR;        /* clus.N_MAX, clus.N_STAR */
	long n_local, first_global; /* clus.N_MAX_NEW, Start[myid]: the local stars whose vr / vt enter the cumulative sums */
} cmcb200_mass_pc_params;

typedef struct {
	/* caller's arrays [count], the reference's globals of the same names (include/cmc/cmc_vars.h:195-196); entries the loop never reaches
	 * are left untouched, like the reference leaves them */
	double *mass_r, *ave_mass_r, *no_star_r, *densities_r, *ke_rad_r, *ke_tan_r, *v2_rad_r, *v2_tan_r;
	int *owner;                /* (may be NULL) 1 where this rank owns the recorded index, i.e. where the four sums were written */
	double totals[4];          /* the rank's sums over all its local stars: its contribution to the MPI_Exscan (:915-931) */
	long filled;               /* mcount at exit */
} cmcb200_mass_pc_result;

/* The enclosed-mass recurrence runs on the replicated tables, the four cumulative kinetic sums on this rank's resident records, all
 * with the reference's sequential roundings.  The sums are RANK-LOCAL: for an index owned by a rank other than 0 the reference
 * reports prefix + local (prefix = MPI_Exscan of the totals, :945-951) -- one addition the caller performs.  On one rank the
 * results are the reference's as they are. */
int cmcb200_dev_comp_mass_percent(cmcb200_ctx *ctx, const cmcb200_mass_pc_params *p, cmcb200_mass_pc_result *res);
/* clusdyn_calculate (src/cmc/cmc_utils.c:1613-1624) on the resident tables: *rh = clusdyn.rh, the half-mass radius, from the same
 * sequentially rounded enclosed-mass sums */
int cmcb200_dev_clusdyn_calculate(cmcb200_ctx *ctx, long N_MAX, long N_STAR, double Mtotal, double *rh);

/* comp_multi_mass_percent (src/cmc/cmc_utils.c:769-861) on the resident tables: Lagrange radii per mass bin.  mass_bins[n_bins] ascending (the
 * reference's mass_bins[], in solar masses; a star belongs to the bin find_stars_mass_bin(star_m / SOLAR_MASS_DYN) gives, :756-764, -1 = none),
 * mass_pc[count]; multi_mass_r[n_bins][count] row-major receives the radii -- the row of a bin without stars is left untouched, a bin with one
 * star gets that star's radius in every entry, like the reference (:845-850).  The per-bin mass sums carry the reference's sequential roundings;
 * the interpolation is gsl_interp_linear restated (GSL is un-vendored: that one step is parity-unpinned, on both sides of the comparison). */
int cmcb200_dev_comp_multi_mass_percent(cmcb200_ctx *ctx, long N_MAX, const double *mass_bins, long n_bins, double solar_mass_dyn,
					const double *mass_pc, long count, double *multi_mass_r);

/* ---- HDF5 snapshot table (src/cmc/cmc_io.c:2960-3345 write_snapshot; include/cmc/cmc.h:136-200 `Snapshot`) ------------------- */
/* The rows write_snapshot() builds in its loop over the stars (:3176-3327) -- 62 fields, 488 bytes each, the struct the reference hands to
 * H5TBmake_table / H5TBappend_records (:3333-3343) -- built on the device from the resident star[], binary[], star_r / star_m / star_phi and
 * written to rows_out, in star order, bit for bit what the reference's loop produces.  The HDF5 calls stay in the reference (the library is
 * not part of this path): write_snapshot() calls this instead of running its loop.
 *   sigma      sigma_array.sigma ([n_local + 1], entry 0 unused; read for binaries' eta only), or NULL: the result cmcb200_dev_calc_sigma_r
 *              left on the device
 *   nan_value  the caller's `0.0 / 0.0` (:3244-3276 fill 33 fields of a single star with it; its bit pattern is the compiler's)
 *   bh_only    != 0: only stars with se_k == 14 or a binary component of type 14 (:3170-3171, :3188)
 * Returns 0, or -3 when the caller's buffer is too small (*n_rows = rows needed, nothing written). */
typedef struct {
	long n_local;        /* clus.N_MAX_NEW */
	long n_old_local, first_global, N_MAX; /* get_global_idx (src/cmc/cmc_utils.c:2384-2407), as in cmcb200_orbit_params */
	long N_STAR;         /* clus.N_STAR */
	double units_m, units_l, madhoc; /* units.m, units.l, madhoc */
	long avekernel;      /* AVEKERNEL (calc_average_mass_sqr, src/cmc/cmc_dynamics_helper.c:2997-3022) */
	double nan_value;
	int bh_only;
} cmcb200_snapshot_params;
int cmcb200_dev_snapshot_rows(cmcb200_ctx *ctx, const cmcb200_snapshot_params *p, const double *sigma, void *rows_out, long cap_rows, long *n_rows);

/* ---- restart file (src/cmc/cmc_io.c:3829-3884) written while the records are resident -------------------------------------- */
/* The file save_restart_file() writes, byte for byte: curr_st (struct rng_t113_state), restart_struct_t, clus_struct_t -- opaque blocks the
 * caller fills exactly as the reference does (save_global_vars) -- then star[0 .. N_STAR_DIM_OPT) and binary[0 .. N_BIN_DIM_OPT) whole, then
 * snapshot_window_counters (:3862-3867).  star[1 .. n_resident] and binary[1 .. n_bin_resident] are taken from the DEVICE (streamed through
 * pinned staging buffers, copy and fwrite overlapped), everything else from the caller's arrays: a checkpoint inside a resident window
 * needs no refresh of the host's 424-byte records.  load_restart_file() (:3886-3935) reads the file unchanged. */
int cmcb200_restart_write(cmcb200_ctx *ctx, const char *path, const void *rng_state, long rng_bytes, const void *restart_struct, long restart_bytes,
			  const void *clus, long clus_bytes, const void *star0, long n_star_dim_opt, long n_resident, const void *binary0,
			  long n_bin_dim_opt, long n_bin_resident, const int *snapshot_window_counters, long snapshot_window_count);

/* ---- dynamics_apply (src/cmc/cmc_dynamics.c:17-542): the loop over neighbouring pairs on the resident arrays ---------------- */
/* Everything of the pair loop (:206-500) that is not a strong encounter: per pair (k, kp) = local stars (1,2), (3,4), ... the kinematics of
 * calc_encounter_dyns (src/cmc/cmc_dynamics_helper.c:831-870; star[k].Y = star[kp].Y = the pair's first draw), calc_n_local (:768-786), the
 * cross section of the pair's type (cmc_dynamics.c:250-432: binary-binary, binary-single, single-single with the sticky-sphere, Lombardi,
 * Quinlan-Shapiro and TC_FACTOR terms), P_enc (:436) and the decision (:444).  A pair whose second draw is < P_enc is an ENCOUNTER: it is
 * listed and left untouched for the caller, which runs the reference's binint_do / sscollision_do on it (cmcb200_dev_get / _put).  Every
 * other pair is relaxed when RELAXATION != 0 (:468-540): Trel12, beta, the relative velocity rotated by (beta, psi = third draw), new
 * star[].vr, .vt and set_star_EJ's .E, .J for both stars.  The scattering-angle statistics of :473-482 come back in the result (sums in a
 * fixed, run-to-run reproducible order that is not the reference's serial order: they agree to rounding and are only printed with %g).
 * sin / cos / log / pow are CUDA's double-precision routines (<= 2 ulp from glibc's), so vr, vt, E, J agree with the reference to ~1e-15
 * relative, not bit for bit; the parity tests state 1e-12.
 * Not handled here (the call fails): TC_POLYTROPE; BH_LOSS_CONE with a central mass (bh_rand_walk).  The three-body loop (:60-204) stays
 * with the caller; with honor_threebb != 0 the stars it flagged (star[].threebb_interacted == 1) are skipped when pairs are formed (:226-240). */
typedef struct {
	long n_local;        /* mpiEnd - mpiBegin + 1 */
	long first_global;   /* Start[myid] */
	long N_MAX, N_STAR;  /* clus.N_MAX (N_LIMIT of calc_n_local), clus.N_STAR */
	double dt;
	long relaxation, ss_collision, binsingle, binbin, tidal_capture, tc_polytrope, bh_capture, bh_loss_cone; /* the reference's switches */
	long avekernel;      /* AVEKERNEL */
	long honor_threebb;
	double cenma_m;
	double gamma, coll_factor, tc_factor, wd_tc, units_mstar, units_l, units_t;
	/* draws != NULL: three uniforms per pair in pair order (host array: Y, the encounter draw, psi) -- "the same uniform random draws fed in
	 * as an input array"; NULL: Philox4x32-10 keyed by (seed, tcount, global index of the pair's first star) */
	const double *draws;
	unsigned long long seed, tcount;
} cmcb200_dynamics_params;
typedef struct {
	long n_pairs;       /* pairs formed (without skipped stars: n_local / 2) */
	long n_encounters;  /* pairs listed for the caller */
	long Nrel, Nrelbeta[4];
	double qaverelbeta[4], maverelbeta[4], raverelbeta[4];
} cmcb200_dynamics_result;
int cmcb200_dev_dynamics_apply(cmcb200_ctx *ctx, const cmcb200_dynamics_params *p, cmcb200_dynamics_result *res);
/* the encounter pairs of the last cmcb200_dev_dynamics_apply in pair order: 1-based local indices of both stars, the type (1 single-single,
 * 2 with a binary) and what the reference hands to sscollision_do / binint_do (rperi, W, rcm).  More than cap pairs: -3, *count set. */
int cmcb200_dev_dynamics_encounters(cmcb200_ctx *ctx, long cap, long *k, long *kp, int *type, double *rperi, double *W, double *rcm, long *count);

/* simul_relax_new (src/cmc/cmc_dynamics_helper.c:2526-2580), the relaxation time step GetTimeStep starts from (src/cmc/cmc_evolution_thr.c:317-331),
 * on the resident arrays: the minimum over this rank's blocks of 2 AVEKERNEL stars of (2 THETASEMAX / PI)^2 (PI / 32) W^3 / (N_STAR n (4 <m^2>)),
 * bit for bit (fresh block sums in star order, IEEE sqrt / divisions); +inf when the rank holds no complete block.  The caller takes the
 * minimum over the ranks (MPI_Allreduce MIN, :2577) and applies calc_timestep's rules (src/cmc/cmc_utils.c:2132-2148). */
int cmcb200_dev_simul_relax(cmcb200_ctx *ctx, long n_local, long first_global, long N_MAX, long N_STAR, long avekernel, double thetasemax, double *DTrel);

/* ---- calc_sigma_r (src/cmc/cmc_dynamics_helper.c:2821-2975) on the resident arrays -------------------------------------- */
/* sigma = sqrt(sum m v^2 / sum m) over the window of 2p stars around every local star, from the reference's two running sums: each star
 * subtracts the star that left the window and adds the one that entered, every operation rounded on its own, so the values carry the
 * rounding history of the whole rank -- reproduced BIT FOR BIT (csrc/sigma_scan.h).  sig_r_or_mave / sig_sigma (host, [n_local + 1], entry
 * 0 untouched; either may be NULL): sigma_array.r / .sigma of dynamics_apply (cmc_dynamics.c:37; r_0_mave_1 == 0) or the mean mass and sigma
 * of the three-body loop (:94; r_0_mave_1 == 1).  The results also stay on the device (cmcb200_dev_sigma_arrays).
 * On P ranks every rank starts from an empty window p stars before its first star, like the reference: ghost_prev_* = vr / vt of the last p
 * stars of the previous rank, ghost_next_* of the first p of the next (host arrays of p doubles; NULL on the first / last rank) -- the
 * MPI_Send / MPI_Recv pair of :2862-2900 is the caller's.
 * stats (may be NULL): [0] segments replayed because a band straddled a binade boundary, [1] segments replayed because they climb too many
 * binades, [2] tiles walked segment by segment, [3] band broken (whole chain replayed by one thread), [4] rerun sequentially, [5] tiles. */
typedef struct {
	long p;             /* AVEKERNEL / BH_AVEKERNEL */
	long n_local;       /* N_LIMIT: mpiEnd - mpiBegin + 1 */
	long first_global;  /* Start[myid] */
	long N_MAX, N_STAR;
	int r_0_mave_1;
	int force_sequential; /* != 0: one thread replays the loop (diagnostic) */
	const double *ghost_prev_vr, *ghost_prev_vt, *ghost_next_vr, *ghost_next_vt;
} cmcb200_sigma_params;
int cmcb200_dev_calc_sigma_r(cmcb200_ctx *ctx, const cmcb200_sigma_params *p, double *sig_r_or_mave, double *sig_sigma, long stats[6]);
int cmcb200_dev_sigma_arrays(cmcb200_ctx *ctx, const double **sig_r_or_mave, const double **sig_sigma);

/* ---- radius-sharded operation: one rank per GPU of an NVLink node ---------------------------------------------------------------- */
/* Replaces sample_sort + load_balance + post_sort_comm (src/cmc/cmc_sort.c:375-791, src/cmc/cmc_utils.c:2259-2288) for procs > 1.
 * The star array is sharded by radius rank exactly like the reference's MPI partition (findLimits), every rank keeps the full
 * star_r / star_m / star_phi tables.  Ranks exchange data by storing straight into each other's device memory (CUDA IPC + NVLink);
 * the HOST layer only moves a few bytes of metadata and provides the barriers between the phases (MPI in CMC, torch.distributed in
 * the Python mirror).  Per timestep, on every rank:
 *     cmcb200_dev_new_orbits(... first_global = Start[rank] ...)
 *     cmcb200_shard_sort_local(ctx, n_local, &n_keep)                   host: allgather n_keep -> findLimits -> Start/End;  BARRIER
 *     cmcb200_shard_exchange(ctx, n_keep_all, Start, End, ...)           BARRIER
 *     cmcb200_shard_publish(ctx, End[rank]-Start[rank]+1, Start[rank], ...)   BARRIER
 *     cmcb200_shard_finish(ctx)
 *     cmcb200_dev_calc_potential_new(...)                                 (replicated on every rank, like potential_calculate)
 * Set-up, once: cmcb200_create_sharded; cmcb200_shard_export -> host allgathers the handle blobs -> cmcb200_shard_attach. */
int cmcb200_shard_handle_bytes(void);
int cmcb200_shard_export(cmcb200_ctx *ctx, int rank, int nranks, void *handles);
int cmcb200_shard_attach(cmcb200_ctx *ctx, const void *all_handles);
int cmcb200_shard_detach(cmcb200_ctx *ctx);
/* fused != 0 (default): the sort key is rnew and the r <- rnew transfer rides on the exchange; 0: key r, records unchanged (what the reference's
 * qsorts_new does after ComputeIntermediateEnergy; n_old_local / first_global / N_MAX of cmcb200_shard_exchange are then ignored) */
int cmcb200_shard_fuse_transfer(cmcb200_ctx *ctx, int fused);
int cmcb200_shard_sort_local(cmcb200_ctx *ctx, long n_local, long *n_keep);
int cmcb200_shard_exchange(cmcb200_ctx *ctx, const long *n_keep_all, const int *Start, const int *End, long n_old_local, long first_global, long N_MAX);
int cmcb200_shard_publish(cmcb200_ctx *ctx, long n_new_local, long start_global, long n_bin_clear, long *n_bin);
int cmcb200_shard_finish(cmcb200_ctx *ctx);
/* The same four calls as ONE asynchronous sequence: the ranks meet at three barriers made of flags in each other's device memory
 * (st.release.sys / ld.acquire.sys over NVLink), the counts travel with the first barrier, every device plans the new partition itself
 * (findLimits with blk = min_chunk) and the kernels behind take their extents from that plan.  The host only enqueues;
 * cmcb200_shard_step_result waits for the stream and returns what the reference's globals need: n_keep_all[procs], Start / End [procs] of the new
 * partition (clus.N_MAX = End[procs-1]), the local binary count.  A barrier that a peer never reaches times out (4 s) and the result call fails.
 * The new star_r / star_m tables do not wait for the records: every rank sends (global rank, star_m) of its sorted run to all ranks, each scatters
 * them (with the radii it already holds as keys) into its own tables on a second stream, and -- with pot_N_STAR > 0 -- the potential rebuild
 * (cmcb200_dev_calc_potential_new's kernels, with cenma.m / cenma.m_new as given) is queued right behind, so that tables and potential
 * overlap the record exchange.  pot_out[4] (may be NULL) then receives N_MAX, cenma.m, cenma.m_new, Mtotal. */
int cmcb200_shard_step_async(cmcb200_ctx *ctx, long n_local, long n_old_local, long first_global, long N_MAX, int min_chunk, long n_bin_clear,
			     long pot_N_STAR, double pot_cenma_m, double pot_cenma_m_new);
int cmcb200_shard_step_result(cmcb200_ctx *ctx, long *n_keep_all, int *Start, int *End, long *n_bin, double pot_out[4]);

/* ---- instrumentation --------------------------------------------------------------------------------------------- */

/* With profiling on, every stage is bracketed by CUDA events on the launching stream.  cmcb200_profile_read returns the
 * accumulated milliseconds and launch counts since the last reset for: 0 keys+histogram, 1 radix passes, 2 record
 * gather, 3 binaries, 4 potential, 5 orbit stage A (orbits of all stars), 6 transfer, 7 energy, 8 h2d, 9 d2h, 10 orbit linear scans,
 * 11 orbit stage B (rejection rounds), 12 orbit lookup tables, 13 sharded: key push, 14 sharded: global ranks, 15 sharded: table publish,
 * 16 Lagrange radii, 17 dynamics_apply's pair loop, 18 calc_sigma_r. */
#define CMCB200_N_STAGES 19
int cmcb200_profile(cmcb200_ctx *ctx, int on);
int cmcb200_profile_read(cmcb200_ctx *ctx, double ms[CMCB200_N_STAGES], long launches[CMCB200_N_STAGES], int reset);
/* kernels launched by this context since it was created */
long cmcb200_launch_count(cmcb200_ctx *ctx);
/* exact-sum diagnostics of the last potential rebuild: events resolved sequentially, full sequential fallbacks */
int cmcb200_potential_stats(cmcb200_ctx *ctx, long *n_events, long *n_fallbacks);

#ifdef __cplusplus
}
#endif
#endif

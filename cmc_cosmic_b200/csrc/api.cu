// C ABI of libcmcb200.so (include/cmcb200.h): context, device-resident mirrors, host<->device staging.
#include <cmath>
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <utility>
#include <vector>
#include "ctx.cuh"

constexpr int ORBIT_MAX_CHUNKS = 16; // pipeline chunks of the host-buffer orbit step (their statistics live in the 4 KB pinned scratch)

namespace cmcb {

static thread_local char g_err[512] = "";
void set_error(const char *fmt, ...)
{
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(g_err, sizeof(g_err), fmt, ap);
	va_end(ap);
}
const char *get_error() { return g_err; }

} // namespace cmcb

extern "C" const char *cmcb200_last_error(void) { return get_error(); }

extern "C" void cmcb200_layout(long out[16])
{
	const long v[16] = {CMCB_SIZEOF_STAR, CMCB_SIZEOF_BINARY, CMCB_OFF_R, CMCB_OFF_VR, CMCB_OFF_VT, CMCB_OFF_M, CMCB_OFF_E, CMCB_OFF_J,
			    CMCB_OFF_RNEW, CMCB_OFF_VRNEW, CMCB_OFF_VTNEW, CMCB_OFF_ROLD, CMCB_OFF_X, CMCB_OFF_R_PERI, CMCB_OFF_R_APO, CMCB_OFF_BININD};
	memcpy(out, v, sizeof(v));
}

// findIndices / findLimits, reference src/cmc/cmc_utils.c:2299-2333
extern "C" void cmcb200_find_limits(long N, int blk, int procs, int *Start, int *End)
{
	const long blocks = N / blk, remain = N % blk;
	const long chunk = (blocks / procs) * blk, extra = blocks % procs;
	for (int i = 0; i < procs; i++) {
		long b, e;
		if (i < extra) { b = i * chunk + (long) i * blk + 1; e = b + chunk + blk - 1; }
		else { b = i * chunk + extra * blk + 1; e = b + chunk - 1; }
		Start[i] = (int) b;
		End[i] = (int) e;
	}
	End[procs - 1] += (int) remain;
}

extern "C" int cmcb200_create(cmcb200_ctx **out, int device, long n_star_dim, long n_bin_dim)
{
	return cmcb200_create_sharded(out, device, n_star_dim, n_star_dim, n_bin_dim);
}

extern "C" int cmcb200_create_sharded(cmcb200_ctx **out, int device, long n_star_dim, long n_table_dim, long n_bin_dim)
{
	*out = nullptr;
	if (n_star_dim < 4 || n_bin_dim < 2) { set_error("cmcb200_create: n_star_dim >= 4 and n_bin_dim >= 2 required"); return -1; }
	if (n_table_dim < n_star_dim) n_table_dim = n_star_dim;
	if (n_table_dim >= (1L << 30)) { set_error("cmcb200_create: dimensions must be < 2^30 (30-bit look-back words, 32-bit table indices)"); return -1; }
	int ndev = 0;
	CMCB_CUDA_OK(cudaGetDeviceCount(&ndev));
	if (device < 0 || device >= ndev) { set_error("cmcb200_create: no CUDA device %d (found %d)", device, ndev); return -1; }
	CMCB_CUDA_OK(cudaSetDevice(device));
	cudaDeviceProp prop;
	CMCB_CUDA_OK(cudaGetDeviceProperties(&prop, device));
	if (prop.major < 10) { set_error("cmcb200_create: built for sm_100a, device %d is sm_%d%d", device, prop.major, prop.minor); return -1; }
	cmcb200_ctx *c = new cmcb200_ctx();
	c->device = device;
	c->n_star_dim = n_star_dim;
	c->n_table_dim = n_table_dim;
	c->rank = 0; c->nranks = 1; c->r_parity = 0;
	c->n_bin_dim = n_bin_dim;
	c->orbit_ladders = 9; // bit 6 (off; needs a build with CMCB_SLAD_F32_TOP): FP32 copy of the sample ladder's coarsest level (its probes screened in FP32 arithmetic); bit 5: the rejection rounds as lookup + try kernels (off: measured slower), bit 4: r-ladder lines read whole (off); bit 3: sample ladder for the index bisections of stage A (find_zero_Q_sampled); bit 0: r ladder (potential() lookups of the rejection rounds), bit 1: Q ladder (index bisections of stage A), bit 2: FP32 screening copy of the (r, Phi) table for the index bisections
	c->bytes_h2d = c->bytes_d2h = 0;
	c->orbit_ctas = 4;
	c->orbit_max_chunks = 8;
	if (const char *e = getenv("CMCB200_ORBIT_CHUNKS")) c->orbit_max_chunks = atoi(e) < 1 ? 1 : (atoi(e) > ORBIT_MAX_CHUNKS ? ORBIT_MAX_CHUNKS : atoi(e));
	if (const char *e = getenv("CMCB200_ORBIT_CTAS")) c->orbit_ctas = atoi(e);
	c->step_emit_keys = true; c->orbit_emit_keys = false;
	if (const char *e = getenv("CMCB200_STEP_KEYS")) c->step_emit_keys = atoi(e) != 0;
	if (const char *e = getenv("CMCB200_ORBIT_LADDERS")) c->orbit_ladders = atoi(e); // 0: A/B measurement of the plain bisections
	CMCB_CUDA_OK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
	c->own_stream = true;
	c->l2_persist_pct = 0;
	c->l2_window_max = (size_t) prop.accessPolicyMaxWindowSize;
	if (const char *e = getenv("CMCB200_L2_PERSIST")) c->l2_persist_pct = atoi(e);
	if (c->l2_persist_pct > 0 && prop.persistingL2CacheMaxSize > 0) CMCB_CUDA_OK(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t) prop.persistingL2CacheMaxSize));
	else c->l2_persist_pct = 0;
	const size_t sb = (size_t) n_star_dim * CMCB_SIZEOF_STAR + 64, bb = (size_t) n_bin_dim * CMCB_SIZEOF_BINARY + 64;
	for (int i = 0; i < 2; i++) {
		CMCB_CUDA_OK(cudaMalloc((void **) &c->d_star[i], sb));
		CMCB_CUDA_OK(cudaMemsetAsync(c->d_star[i], 0, sb, c->stream));
		CMCB_CUDA_OK(cudaMalloc((void **) &c->d_bin[i], bb));
		CMCB_CUDA_OK(cudaMemsetAsync(c->d_bin[i], 0, bb, c->stream));
	}
	const size_t tb = (size_t) n_table_dim * sizeof(double) + 256; // readable a line past the end (16-ary index windows)
	CMCB_CUDA_OK(cudaMalloc((void **) &c->d_r, tb));
	CMCB_CUDA_OK(cudaMalloc((void **) &c->d_m, tb));
	CMCB_CUDA_OK(cudaMalloc((void **) &c->d_phi, tb));
	CMCB_CUDA_OK(cudaMalloc((void **) &c->d_r_alt, tb));
	CMCB_CUDA_OK(cudaMalloc((void **) &c->d_m_alt, tb));
	CMCB_CUDA_OK(cudaMemsetAsync(c->d_r_alt, 0, tb, c->stream));
	CMCB_CUDA_OK(cudaMemsetAsync(c->d_m_alt, 0, tb, c->stream));
	CMCB_CUDA_OK(cudaMalloc((void **) &c->d_mprev, tb));
	CMCB_CUDA_OK(cudaMalloc((void **) &c->d_rp, (size_t) (n_table_dim + 16) * sizeof(RPhi))); // the r ladder's last step reads whole 8-entry lines
	CMCB_CUDA_OK(cudaMemsetAsync(c->d_rp, 0, (size_t) (n_table_dim + 16) * sizeof(RPhi), c->stream));
	CMCB_CUDA_OK(cudaMalloc((void **) &c->d_rpf, (size_t) (n_table_dim + 16) * sizeof(RPhiF)));
	CMCB_CUDA_OK(cudaMemsetAsync(c->d_rpf, 0, (size_t) (n_table_dim + 16) * sizeof(RPhiF), c->stream));
	c->lad.q_cap = ladder_q_capacity(n_table_dim);
	c->lad.r_cap = ladder_r_capacity(n_table_dim);
	CMCB_CUDA_OK(cudaMalloc((void **) &c->lad.q, (size_t) c->lad.q_cap * sizeof(LBlock)));
	CMCB_CUDA_OK(cudaMalloc((void **) &c->lad.r, (size_t) c->lad.r_cap * sizeof(double)));
	CMCB_CUDA_OK(cudaMalloc((void **) &c->lad.unsorted, sizeof(int)));
	c->lad.q_max_shift = 64; c->lad.q_max_width_log2 = 31; c->lad.q_factor_log2 = QLAD_FACTOR_LOG2;
	if (const char *e = getenv("CMCB200_QLAD_MAX_SHIFT")) c->lad.q_max_shift = atoi(e);
	if (const char *e = getenv("CMCB200_QLAD_MAX_WIDTH_LOG2")) c->lad.q_max_width_log2 = atoi(e);
	if (const char *e = getenv("CMCB200_QLAD_FACTOR_LOG2")) c->lad.q_factor_log2 = atoi(e);
	CMCB_CUDA_OK(cudaMalloc((void **) &c->d_defer_star, (size_t) ORBIT_DEFER_CAP * sizeof(long)));
	CMCB_CUDA_OK(cudaMalloc((void **) &c->d_defer_found, (size_t) ORBIT_DEFER_CAP * sizeof(long)));
	CMCB_CUDA_OK(cudaMalloc((void **) &c->d_oflag, (size_t) n_star_dim));
	CMCB_CUDA_OK(cudaMalloc((void **) &c->d_defer_first, (size_t) ORBIT_DEFER_CAP * sizeof(unsigned)));
	CMCB_CUDA_OK(cudaMalloc((void **) &c->d_job, (size_t) 6 * n_star_dim * sizeof(double)));
	CMCB_CUDA_OK(cudaMalloc((void **) &c->d_pend[0], (size_t) n_star_dim * sizeof(int)));
	CMCB_CUDA_OK(cudaMalloc((void **) &c->d_pend[1], (size_t) n_star_dim * sizeof(int)));
	CMCB_CUDA_OK(cudaMemsetAsync(c->d_r, 0, tb, c->stream));
	CMCB_CUDA_OK(cudaMemsetAsync(c->d_m, 0, tb, c->stream));
	CMCB_CUDA_OK(cudaMemsetAsync(c->d_phi, 0, tb, c->stream));
	CMCB_CUDA_OK(cudaMalloc((void **) &c->sbuf.keys_a, (size_t) n_star_dim * 8));
	CMCB_CUDA_OK(cudaMalloc((void **) &c->sbuf.keys_b, (size_t) n_star_dim * 8));
	CMCB_CUDA_OK(cudaMalloc((void **) &c->sbuf.vals_a, (size_t) n_star_dim * 4));
	CMCB_CUDA_OK(cudaMalloc((void **) &c->sbuf.vals_b, (size_t) n_star_dim * 4));
	CMCB_CUDA_OK(cudaMalloc((void **) &c->d_sort, sizeof(SortState)));
	const long ntiles = (n_star_dim + SORT_TILE - 1) / SORT_TILE;
	c->tile_hist_words = ntiles * 256 * 8;
	CMCB_CUDA_OK(cudaMalloc((void **) &c->d_tile_hist, (size_t) c->tile_hist_words * 4));
	CMCB_CUDA_OK(cudaMalloc((void **) &c->d_binflag, (size_t) n_star_dim));
	CMCB_CUDA_OK(cudaMalloc((void **) &c->d_bin_block_cnt, (size_t) (n_star_dim / 2048 + 2) * 4));
	CMCB_CUDA_OK(cudaMalloc((void **) &c->d_pot, sizeof(PotScalars)));
	const size_t wb = pot_work_bytes(n_table_dim);
	CMCB_CUDA_OK(cudaMalloc(&c->d_seqwork, wb));
	CMCB_CUDA_OK(cudaMemsetAsync(c->d_seqwork, 0, wb, c->stream));
	pot_work_bind(c->seqwork, c->d_seqwork, n_table_dim);
	CMCB_CUDA_OK(cudaMalloc((void **) &c->d_ostats, sizeof(OrbitStats)));
	CMCB_CUDA_OK(cudaMalloc((void **) &c->d_status, (size_t) n_star_dim * sizeof(int)));
	CMCB_CUDA_OK(cudaMalloc((void **) &c->d_kmin, (size_t) n_star_dim * sizeof(long)));
	CMCB_CUDA_OK(cudaMalloc((void **) &c->d_kmax, (size_t) n_star_dim * sizeof(long)));
	CMCB_CUDA_OK(cudaMallocHost(&c->h_pinned, 8192)); // [0, 4096): scratch of whichever call reads scalars back; [4096, 8192): the orbit step's per-chunk statistics
	if (sort_setup()) { set_error("cmcb200_create: cannot raise the dynamic shared memory limit of the sort kernel"); return -1; }
	CMCB_CUDA_OK(cudaStreamSynchronize(c->stream));
	*out = c;
	return 0;
}

extern "C" void cmcb200_destroy(cmcb200_ctx *c)
{
	if (!c) return;
	cudaSetDevice(c->device);
	cudaStreamSynchronize(c->stream);
	for (int i = 0; i < 2; i++) { cudaFree(c->d_star[i]); cudaFree(c->d_bin[i]); }
	cudaFree(c->d_r); cudaFree(c->d_m); cudaFree(c->d_phi); cudaFree(c->d_mprev); cudaFree(c->d_r_alt); cudaFree(c->d_m_alt);
	cudaFree(c->d_rp); cudaFree(c->d_rpf); cudaFree(c->d_defer_star); cudaFree(c->d_defer_found);
	cudaFree(c->d_defer_first); cudaFree(c->lad.q); cudaFree(c->lad.r); cudaFree(c->lad.unsorted);
	if (c->d_keys_all) { cudaFree(c->d_keys_all); cudaFree(c->d_src_all); cudaFree(c->d_dest_pos); cudaFree(c->d_dest_rank); cudaFree(c->d_shard_meta); cudaFree(c->d_shard_sync); cudaFree(c->d_rank_all); cudaFree(c->d_m_all);
		if (c->s_aux) { cudaStreamDestroy(c->s_aux); cudaEventDestroy(c->ev_fork); cudaEventDestroy(c->ev_join); } }
	cudaFree(c->d_oflag); cudaFree(c->d_job); cudaFree(c->d_pend[0]); cudaFree(c->d_pend[1]);
	cudaFree(c->sbuf.keys_a); cudaFree(c->sbuf.keys_b); cudaFree(c->sbuf.vals_a); cudaFree(c->sbuf.vals_b);
	cudaFree(c->d_sort); cudaFree(c->d_tile_hist); cudaFree(c->d_binflag); cudaFree(c->d_bin_block_cnt);
	cudaFree(c->d_pot); cudaFree(c->d_seqwork); cudaFree(c->d_ostats); cudaFree(c->d_status); cudaFree(c->d_kmin); cudaFree(c->d_kmax);
	if (c->d_lagrange) cudaFree(c->d_lagrange);
	if (c->d_multimass) cudaFree(c->d_multimass);
	if (c->d_sv3) cudaFree(c->d_sv3);
	dynamics_free(c);
	sigma_free(c);
	if (c->d_seqwork2) cudaFree(c->d_seqwork2);
	if (c->d_draws) cudaFree(c->d_draws);
	if (c->d_draw_off) cudaFree(c->d_draw_off);
	cudaFreeHost(c->h_pinned);
	for (auto &e : c->events) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
	for (auto &e : c->pool) cudaEventDestroy(e);
	if (c->s_in) {
		cudaStreamDestroy(c->s_in); cudaStreamDestroy(c->s_out);
		for (int k = 0; k < 16; k++) { cudaEventDestroy(c->ev_in[k]); cudaEventDestroy(c->ev_done[k]); }
		cudaEventDestroy(c->ev_tab); cudaEventDestroy(c->ev_bin);
	}
	if (c->own_stream) cudaStreamDestroy(c->stream);
	delete c;
}

extern "C" int cmcb200_set_stream(cmcb200_ctx *c, void *s)
{
	CMCB_CUDA_OK(cudaStreamSynchronize(c->stream));
	if (c->own_stream) cudaStreamDestroy(c->stream);
	c->stream = (cudaStream_t) s;
	c->own_stream = false;
	return 0;
}

extern "C" int cmcb200_sync(cmcb200_ctx *c)
{
	CMCB_CUDA_OK(cudaStreamSynchronize(c->stream));
	return 0;
}

extern "C" int cmcb200_profile(cmcb200_ctx *c, int on)
{
	c->profiling = on != 0;
	return 0;
}

extern "C" int cmcb200_profile_read(cmcb200_ctx *c, double ms[CMCB200_N_STAGES], long launches[CMCB200_N_STAGES], int reset)
{
	CMCB_CUDA_OK(cudaStreamSynchronize(c->stream));
	for (auto &e : c->events) {
		float t = 0.f;
		if (cudaEventElapsedTime(&t, e.a, e.b) == cudaSuccess) { c->acc_ms[e.stage] += t; c->acc_n[e.stage]++; }
		c->pool.push_back(e.a);
		c->pool.push_back(e.b);
	}
	c->events.clear();
	for (int i = 0; i < CMCB200_N_STAGES; i++) { ms[i] = c->acc_ms[i]; launches[i] = c->acc_n[i]; }
	if (reset) for (int i = 0; i < CMCB200_N_STAGES; i++) { c->acc_ms[i] = 0; c->acc_n[i] = 0; }
	return 0;
}

extern "C" long cmcb200_launch_count(cmcb200_ctx *c) { return c->launches; }

extern "C" int cmcb200_potential_stats(cmcb200_ctx *c, long *n_events, long *n_fallbacks)
{
	long h[2];
	CMCB_CUDA_OK(cudaMemcpyAsync(h, c->seqwork.stats, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
	CMCB_CUDA_OK(cudaStreamSynchronize(c->stream));
	*n_events = h[0];
	*n_fallbacks = h[1];
	return 0;
}

// ---- potential ------------------------------------------------------------------------------------------------

// the rebuild queued on `stream` (the context's own, or the second stream of the asynchronous sharded step).  n_max_upper: N_MAX on entry -- the
// new N_MAX is the number of leading stars with r < SF_INFINITY (cmc_utils.c:685-696), found on the device
int potential_enqueue(cmcb200_ctx *c, long n_max_upper, long N_STAR, double cenma_m, double cenma_m_new, cudaStream_t stream)
{
	cudaStream_t keep = c->stream;
	c->stream = stream; // the stage timer records on the stream the kernels run on
	{
		StageTimer t(c, ST_POTENTIAL);
		cudaMemsetAsync(c->seqwork.stats, 0, 2 * sizeof(long), stream);
		PotArgs a;
		a.r = c->d_r; a.m = c->d_m; a.phi = c->d_phi; a.mprev = c->d_mprev; a.sc = c->d_pot; a.work = c->seqwork;
		a.n_max_in = n_max_upper; a.n_star = (double) N_STAR; a.madhoc = 1.0 / ((double) N_STAR);
		a.cenma_m = cenma_m; a.cenma_m_new = cenma_m_new;
		c->launches += potential_calculate_device(a, stream);
	}
	c->stream = keep;
	CMCB_CUDA_OK(cudaGetLastError());
	return 0;
}

extern "C" int cmcb200_dev_calc_potential_new(cmcb200_ctx *c, long *N_MAX, long N_STAR, double *cenma_m, double *cenma_m_new, double orbit_r,
						      double *Mtotal, double *Rtidal)
{
	if (*N_MAX + 2 > c->n_table_dim) { set_error("calc_potential_new: N_MAX + 2 = %ld exceeds the table capacity %ld", *N_MAX + 2, c->n_table_dim); return -1; }
	CMCB_CUDA_OK(cudaSetDevice(c->device));
	if (join_uploads(c)) return -1;
	if (potential_enqueue(c, *N_MAX, N_STAR, *cenma_m, *cenma_m_new, c->stream)) return -1;
	PotScalars *h = (PotScalars *) c->h_pinned;
	CMCB_CUDA_OK(cudaMemcpyAsync(h, c->d_pot, sizeof(PotScalars), cudaMemcpyDeviceToHost, c->stream));
	CMCB_CUDA_OK(cudaStreamSynchronize(c->stream));
	*N_MAX = h->n_max;
	*cenma_m = h->cenma_m;
	*cenma_m_new = h->cenma_m_new;
	*Mtotal = h->Mtotal;
	// libm pow on the host: the same correctly rounded glibc routine the reference calls (cmc_utils.c:707); CUDA's pow is not
	*Rtidal = orbit_r * pow(h->Mtotal, 1.0 / 3.0);
	return 0;
}

extern "C" int cmcb200_calc_potential_new(cmcb200_ctx *c, double *star_r, const double *star_m, double *star_phi, long *N_MAX, long N_STAR,
					  double *cenma_m, double *cenma_m_new, double orbit_r, double *Mtotal, double *Rtidal)
{
	const long n_in = *N_MAX + 2;
	if (n_in > c->n_table_dim) { set_error("calc_potential_new: N_MAX + 2 = %ld exceeds the table capacity %ld", n_in, c->n_table_dim); return -1; }
	CMCB_CUDA_OK(cudaSetDevice(c->device));
	{
		StageTimer t(c, ST_H2D);
		CMCB_CUDA_OK(cudaMemcpyAsync(c->d_r, star_r, n_in * sizeof(double), cudaMemcpyHostToDevice, c->stream));
		CMCB_CUDA_OK(cudaMemcpyAsync(c->d_m, star_m, n_in * sizeof(double), cudaMemcpyHostToDevice, c->stream));
		c->bytes_h2d += 2 * n_in * (long) sizeof(double);
	}
	if (cmcb200_dev_calc_potential_new(c, N_MAX, N_STAR, cenma_m, cenma_m_new, orbit_r, Mtotal, Rtidal)) return -1;
	{
		StageTimer t(c, ST_D2H);
		CMCB_CUDA_OK(cudaMemcpyAsync(star_phi, c->d_phi, (*N_MAX + 2) * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
		c->bytes_d2h += (*N_MAX + 2) * (long) sizeof(double);
	}
	CMCB_CUDA_OK(cudaStreamSynchronize(c->stream));
	star_r[*N_MAX + 1] = CMCB200_SF_INFINITY; // cmc_utils.c:713
	return 0;
}

// ---- sort -------------------------------------------------------------------------------------------------------

// fuse: the ComputeIntermediateEnergy transfer (r <- rnew ...) happens while the records move; n_old_local / first_global / N_MAX
// are those of the orbit step that produced rnew
// sync == false: everything is only queued; the five counters land in the pinned block at `counts_at` and the caller waits and reads them itself
// keys_ready: the orbit step queued in front has written the sort keys of all n records (OrbitArgs::sort_keys)
static int sort_common(cmcb200_ctx *c, long *n_local, long n_bin_clear, long *n_bin, bool fuse, long n_old_local, long first_global, long N_MAX,
		       bool sync = true, long counts_at = 0, bool keys_ready = false)
{
	const long n = *n_local;
	if (n + 2 > c->n_star_dim) { set_error("qsorts_new: n_local + 2 = %ld exceeds n_star_dim = %ld", n + 2, c->n_star_dim); return -1; }
	if (n_bin_clear > c->n_bin_dim - 1) n_bin_clear = c->n_bin_dim - 1;
	CMCB_CUDA_OK(cudaSetDevice(c->device));
	if (join_uploads(c)) return -1;
	SortArgs a;
	a.star_in = star1(c);
	a.star_out = c->d_star[c->cur_star ^ 1] + CMCB_SIZEOF_STAR;
	a.n = n;
	a.key_off = fuse ? CMCB_OFF_RNEW : CMCB_OFF_R;
	a.buf = c->sbuf;
	a.st = c->d_sort;
	a.tile_hist = c->d_tile_hist;
	a.keep_keys = false;
	a.keys_ready = keys_ready;
	a.fuse_transfer = fuse;
	a.old_r = c->d_r; a.old_m = c->d_m;
	a.n_old_local = n_old_local; a.first_global = first_global; a.N_MAX = N_MAX;
	// the fused form reads the old tables while the new ones are written: ping-pong
	a.star_r = fuse ? c->d_r_alt : c->d_r;
	a.star_m = fuse ? c->d_m_alt : c->d_m;
	a.binflag = c->d_binflag;
	a.bin_block_cnt = c->d_bin_block_cnt;
	a.bin_in = bin1(c);
	a.bin_out = c->d_bin[c->cur_bin ^ 1] + CMCB_SIZEOF_BINARY;
	a.n_bin_clear = n_bin_clear;
	{ StageTimer t(c, ST_KEYS); c->launches += sort_stage_keys(a, c->stream); }
	{ StageTimer t(c, ST_PASSES); c->launches += sort_stage_passes(a, c->stream); }
	{
		StageTimer t(c, ST_GATHER);
		// the sentinel record star[0] travels with the array
		CMCB_CUDA_OK(cudaMemcpyAsync(c->d_star[c->cur_star ^ 1], c->d_star[c->cur_star], CMCB_SIZEOF_STAR, cudaMemcpyDeviceToDevice, c->stream));
		if (fuse) { // so does the table sentinel [0]
			CMCB_CUDA_OK(cudaMemcpyAsync(c->d_r_alt, c->d_r, sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
			CMCB_CUDA_OK(cudaMemcpyAsync(c->d_m_alt, c->d_m, sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
		}
		c->launches += sort_stage_gather(a, c->stream);
	}
	{ StageTimer t(c, ST_BINARIES); c->launches += sort_stage_binaries(a, c->stream); }
	CMCB_CUDA_OK(cudaGetLastError());
	c->cur_star ^= 1;
	c->cur_bin ^= 1;
	if (fuse) { std::swap(c->d_r, c->d_r_alt); std::swap(c->d_m, c->d_m_alt); c->r_parity ^= 1; }
	long *h = (long *) ((char *) c->h_pinned + counts_at); // n_eq_inf, n_gt_inf, n_total, n_new, n_bin
	CMCB_CUDA_OK(cudaMemcpyAsync(h, &c->d_sort->n_eq_inf, 5 * sizeof(long), cudaMemcpyDeviceToHost, c->stream));
	if (!sync) return 0;
	CMCB_CUDA_OK(cudaStreamSynchronize(c->stream));
	*n_local = h[3];
	*n_bin = h[4];
	return 0;
}

// ---- the whole resident step with ONE wait: orbit step -> r <- rnew + sort -> potential rebuild ---------------------------------------
static int orbit_common(cmcb200_ctx *c, const cmcb200_orbit_params *p, bool want_status, cmcb200_orbit_result *res, bool sync);
static void orbit_end(cmcb200_ctx *c, const cmcb200_orbit_params *p, int n_chunks, cmcb200_orbit_result *res);
__global__ void step_table_sentinel(const cmcb::SortState *st, double *star_r, double *star_m)
{
	// the entry behind the last star that stayed: potential_calculate finds N_MAX as the first radius that is not < SF_INFINITY (cmc_utils.c:685)
	star_r[1 + st->n_new] = 1.0e10;
	star_m[1 + st->n_new] = 0.0;
}
extern "C" int cmcb200_dev_step(cmcb200_ctx *c, const cmcb200_orbit_params *p, long n_bin_clear, double cenma_m_new, double orbit_r,
				cmcb200_orbit_result *orbit, cmcb200_step_result *out)
{
	CMCB_CUDA_OK(cudaSetDevice(c->device));
	if (p->n_local + 2 > c->n_table_dim) { set_error("dev_step: n_local + 2 = %ld exceeds the table capacity %ld", p->n_local + 2, c->n_table_dim); return -1; }
	// the orbit kernels leave key_from_double(rnew) of every star in the sort's key buffer: the sort skips its strided pass over the records
	// (0.2 ms of 10.3 at N = 1e7; CMCB200_STEP_KEYS=0 switches it off for A/B measurements)
	c->orbit_emit_keys = c->step_emit_keys;
	const int orc = orbit_common(c, p, false, nullptr, false);
	c->orbit_emit_keys = false;
	if (orc) return -1;
	long n = p->n_local, nb = 0;
	if (sort_common(c, &n, n_bin_clear, &nb, true, p->n_old_local, p->first_global, p->N_MAX, false, 2048, c->step_emit_keys)) return -1;
	step_table_sentinel<<<1, 1, 0, c->stream>>>(c->d_sort, c->d_r, c->d_m);
	c->launches += 1;
	if (potential_enqueue(c, p->n_local, p->N_STAR, p->cenma_m, cenma_m_new, c->stream)) return -1;
	PotScalars *h = (PotScalars *) c->h_pinned;
	CMCB_CUDA_OK(cudaMemcpyAsync(h, c->d_pot, sizeof(PotScalars), cudaMemcpyDeviceToHost, c->stream));
	CMCB_CUDA_OK(cudaStreamSynchronize(c->stream)); // the one wait of the step
	orbit_end(c, p, c->last_orbit_chunks > 0 ? c->last_orbit_chunks : 1, orbit);
	const long *cnt = (const long *) ((const char *) c->h_pinned + 2048);
	out->n_local = cnt[3];
	out->n_bin = cnt[4];
	out->N_MAX = h->n_max;
	out->cenma_m = h->cenma_m;
	out->cenma_m_new = h->cenma_m_new;
	out->Mtotal = h->Mtotal;
	out->Rtidal = orbit_r * pow(h->Mtotal, 1.0 / 3.0); // glibc's pow on the host, like the reference (cmc_utils.c:707)
	return 0;
}

extern "C" int cmcb200_dev_qsorts_new(cmcb200_ctx *c, long *n_local, long n_bin_clear, long *n_bin)
{
	return sort_common(c, n_local, n_bin_clear, n_bin, false, 0, 0, 0);
}

extern "C" int cmcb200_dev_transfer_and_sort(cmcb200_ctx *c, long *n_local, long n_old_local, long first_global, long N_MAX, long n_bin_clear, long *n_bin)
{
	return sort_common(c, n_local, n_bin_clear, n_bin, true, n_old_local, first_global, N_MAX);
}

extern "C" int cmcb200_qsorts_new(cmcb200_ctx *c, void *star1_h, long *n_local, void *binary1_h, long n_bin_dim, long *n_bin,
				  double *star_r, double *star_m)
{
	const long n = *n_local, nb_in = binary1_h ? *n_bin : 0;
	if (n + 2 > c->n_star_dim) { set_error("qsorts_new: n_local + 2 = %ld exceeds n_star_dim = %ld", n + 2, c->n_star_dim); return -1; }
	if (nb_in + 1 > c->n_bin_dim || n_bin_dim > c->n_bin_dim) { set_error("qsorts_new: binary capacity %ld exceeded", c->n_bin_dim); return -1; }
	CMCB_CUDA_OK(cudaSetDevice(c->device));
	{
		StageTimer t(c, ST_H2D);
		CMCB_CUDA_OK(cudaMemcpyAsync(star1(c), star1_h, (size_t) n * CMCB_SIZEOF_STAR, cudaMemcpyHostToDevice, c->stream));
		if (nb_in > 0) CMCB_CUDA_OK(cudaMemcpyAsync(bin1(c), binary1_h, (size_t) nb_in * CMCB_SIZEOF_BINARY, cudaMemcpyHostToDevice, c->stream));
		c->bytes_h2d += n * CMCB_SIZEOF_STAR + nb_in * CMCB_SIZEOF_BINARY;
	}
	long nb_out = 0;
	// nothing to clear on the device: the host tail is zeroed below
	if (cmcb200_dev_qsorts_new(c, n_local, 0, &nb_out)) return -1;
	{
		StageTimer t(c, ST_D2H);
		CMCB_CUDA_OK(cudaMemcpyAsync(star1_h, star1(c), (size_t) n * CMCB_SIZEOF_STAR, cudaMemcpyDeviceToHost, c->stream));
		if (binary1_h && nb_out > 0) CMCB_CUDA_OK(cudaMemcpyAsync(binary1_h, bin1(c), (size_t) nb_out * CMCB_SIZEOF_BINARY, cudaMemcpyDeviceToHost, c->stream));
		if (star_r) CMCB_CUDA_OK(cudaMemcpyAsync(star_r + 1, c->d_r + 1, (size_t) *n_local * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
		if (star_m) CMCB_CUDA_OK(cudaMemcpyAsync(star_m + 1, c->d_m + 1, (size_t) *n_local * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
		c->bytes_d2h += n * CMCB_SIZEOF_STAR + nb_out * CMCB_SIZEOF_BINARY + ((star_r ? 1 : 0) + (star_m ? 1 : 0)) * *n_local * (long) sizeof(double);
	}
	CMCB_CUDA_OK(cudaStreamSynchronize(c->stream));
	// memset(b_buf, 0, (N_BIN_DIM_OPT-1)*sizeof(binary_t)) (cmc_sort.c:541): everything past the packed binaries is zero
	if (binary1_h && n_bin_dim - 1 > nb_out)
		memset((char *) binary1_h + (size_t) nb_out * CMCB_SIZEOF_BINARY, 0, (size_t) (n_bin_dim - 1 - nb_out) * CMCB_SIZEOF_BINARY);
	if (n_bin) *n_bin = nb_out;
	return 0;
}

// ---- orbit step ---------------------------------------------------------------------------------------------------

// The orbit step in three parts so that the host-buffer form can pipeline it over chunks of the star range:
//   orbit_begin   arguments, recorded draws, lookup tables (packed table + ladders) -- once per call
//   orbit_range   the three stages on local stars [off, off + len): every star is independent of the others (its draws are keyed by its
//                 global index or come from its own CSR slice), so the result does not depend on the chunking
//   orbit_end     combines the per-chunk statistics
// Experiment knob CMCB200_L2_PERSIST: while the orbit kernels run, a random `pct` percent of the (r, Phi) table's lines are kept in L2's
// persisting carve-out (stream access-policy window), so that the 424-byte record stream cannot evict them.  bytes == 0 switches it off.
static void orbit_l2_window(cmcb200_ctx *c, size_t bytes)
{
	if (c->l2_persist_pct <= 0) return;
	cudaStreamAttrValue v;
	memset(&v, 0, sizeof(v));
	v.accessPolicyWindow.base_ptr = (void *) c->d_rp;
	v.accessPolicyWindow.num_bytes = bytes < c->l2_window_max ? bytes : c->l2_window_max;
	v.accessPolicyWindow.hitRatio = bytes ? (float) c->l2_persist_pct / 100.0f : 0.0f;
	v.accessPolicyWindow.hitProp = bytes ? cudaAccessPropertyPersisting : cudaAccessPropertyNormal;
	v.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
	cudaStreamSetAttribute(c->stream, cudaStreamAttributeAccessPolicyWindow, &v);
}

// copy-in / copy-out streams and the per-chunk events of the pipelines (host-buffer orbit step, resident timestep): created on first use
static int copy_streams(cmcb200_ctx *c)
{
	if (c->s_in) return 0;
	CMCB_CUDA_OK(cudaStreamCreateWithFlags(&c->s_in, cudaStreamNonBlocking));
	CMCB_CUDA_OK(cudaStreamCreateWithFlags(&c->s_out, cudaStreamNonBlocking));
	for (int k = 0; k < ORBIT_MAX_CHUNKS; k++) {
		CMCB_CUDA_OK(cudaEventCreateWithFlags(&c->ev_in[k], cudaEventDisableTiming));
		CMCB_CUDA_OK(cudaEventCreateWithFlags(&c->ev_done[k], cudaEventDisableTiming));
	}
	CMCB_CUDA_OK(cudaEventCreateWithFlags(&c->ev_tab, cudaEventDisableTiming));
	CMCB_CUDA_OK(cudaEventCreateWithFlags(&c->ev_bin, cudaEventDisableTiming));
	return 0;
}

int orbit_begin(cmcb200_ctx *c, const cmcb200_orbit_params *p, bool want_status, bool need_status, OrbitArgs &a)
{
	if (join_uploads(c)) return -1;
	if (p->n_local + 2 > c->n_star_dim || p->n_table > c->n_table_dim) { set_error("new_orbits_calculate: size exceeds the capacity (%ld records, %ld table entries)", c->n_star_dim, c->n_table_dim); return -1; }
	a.star1 = star1(c);
	a.n_local = p->n_local;
	a.n_old_local = p->n_old_local;
	a.first_global = p->first_global;
	a.t.rp = c->d_rp; a.t.m = c->d_m;
	a.defer_star = c->d_defer_star; a.defer_found = c->d_defer_found;
	a.fuse_energy = p->fuse_compute_energy != 0;
	a.prepare_ctas = c->orbit_ctas;
	a.defer_first = c->d_defer_first;
	a.flag = c->d_oflag; a.pend_list[0] = c->d_pend[0]; a.pend_list[1] = c->d_pend[1];
	a.sort_keys = c->orbit_emit_keys ? (unsigned long long *) c->sbuf.keys_a : nullptr;
	a.job = c->d_job; a.job_stride = c->n_star_dim;
	a.split_lookup = (c->orbit_ladders & 32) != 0 && (c->orbit_ladders & 1);
	a.t.N_MAX = p->N_MAX;
	a.t.invN = 1.0 / p->N_STAR;
	a.t.madhoc = 1.0 / ((double) p->N_STAR);
	a.t.cenma_m = p->cenma_m;
	a.Rtidal = p->Rtidal;
	a.draws = nullptr; a.draw_off = nullptr;
	a.seed = p->seed; a.tcount = p->tcount;
	if (p->draws) {
		const long nd = p->draw_off[p->n_local];
		if (ensure((void **) &c->d_draws, &c->draws_cap, nd + 1, sizeof(double))) return -1;
		if (ensure((void **) &c->d_draw_off, &c->draw_off_cap, p->n_local + 1, sizeof(long))) return -1;
		StageTimer t(c, ST_H2D);
		CMCB_CUDA_OK(cudaMemcpyAsync(c->d_draws, p->draws, (size_t) nd * sizeof(double), cudaMemcpyHostToDevice, c->stream));
		CMCB_CUDA_OK(cudaMemcpyAsync(c->d_draw_off, p->draw_off, (size_t) (p->n_local + 1) * sizeof(long), cudaMemcpyHostToDevice, c->stream));
		a.draws = c->d_draws;
		a.draw_off = c->d_draw_off;
	}
	a.status = (want_status || need_status) ? c->d_status : nullptr;
	a.kmin = want_status ? c->d_kmin : nullptr;
	a.kmax = want_status ? c->d_kmax : nullptr;
	a.n_draws = nullptr;
	a.stats = c->d_ostats;
	{
		StageTimer t(c, ST_ORBIT_TABLES);
		LadderBuffers lb = c->lad;
		if (!(c->orbit_ladders & 1)) lb.r = nullptr;
		lb.r_whole_lines = (c->orbit_ladders & 16) != 0;
		lb.sampled = (c->orbit_ladders & 8) && !(c->orbit_ladders & 2); // the sample ladder lives in the Q ladder's buffer
		lb.f32_top = (c->orbit_ladders & 64) != 0;
		if (!(c->orbit_ladders & 2) && !lb.sampled) lb.q = nullptr;
		c->launches += orbit_build_tables_device(c->d_r, c->d_phi, c->d_rp, (c->orbit_ladders & 4) ? c->d_rpf : nullptr, p->n_table, p->N_MAX + 1, lb, a, c->stream);
	}
	orbit_l2_window(c, (size_t) p->n_table * sizeof(RPhi));
	return 0;
}


static int orbit_range(cmcb200_ctx *c, const OrbitArgs &whole, long off, long len, int chunk)
{
	OrbitArgs a = whole;
	a.star1 += off * CMCB_SIZEOF_STAR;
	a.n_local = len;
	a.n_old_local = whole.n_old_local - off; // may be <= 0: every star of the chunk was created this step (get_global_idx, cmc_utils.c:2384-2407)
	a.first_global = whole.first_global + off;
	a.flag += off;
	if (a.sort_keys) a.sort_keys += off;
	a.job += off;
	if (a.draw_off) a.draw_off += off;
	if (a.status) a.status += off;
	if (a.kmin) { a.kmin += off; a.kmax += off; }
	{ StageTimer t(c, ST_ORBIT); c->launches += orbit_stage_prepare(a, c->stream); }
	{ StageTimer t(c, ST_ORBIT_SCAN); c->launches += orbit_stage_scan(a, c->stream); }
	{ StageTimer t(c, ST_ORBIT_TRIES); c->launches += orbit_stage_tries(a, c->stream); }
	CMCB_CUDA_OK(cudaGetLastError());
	CMCB_CUDA_OK(cudaMemcpyAsync((OrbitStats *) ((char *) c->h_pinned + 4096) + chunk, c->d_ostats, sizeof(OrbitStats), cudaMemcpyDeviceToHost, c->stream));
	return 0;
}

// after the stream has been synchronised
static void orbit_end(cmcb200_ctx *c, const cmcb200_orbit_params *p, int n_chunks, cmcb200_orbit_result *res)
{
	orbit_l2_window(c, 0);
	if (!res) return;
	const OrbitStats *h = (const OrbitStats *) ((const char *) c->h_pinned + 4096);
	unsigned long long max_bits = 0, other = 0;
	memset(res, 0, sizeof(*res));
	for (int k = 0; k < n_chunks; k++) {
		if (h[k].max_rad_bits > max_bits) max_bits = h[k].max_rad_bits; // positive doubles order like their bit patterns
		res->n_draws += h[k].n_draws; res->n_exact += h[k].n_exact; res->n_brent += h[k].n_brent;
		for (int i = 1; i < 6; i++) { res->n_status[i] += h[k].n_status[i]; other += h[k].n_status[i]; }
		res->n_deferred += h[k].n_defer;
	}
	memcpy(&res->max_r, &max_bits, 8);
	res->n_status[0] = (unsigned long long) p->n_local - other;
}

static int orbit_common(cmcb200_ctx *c, const cmcb200_orbit_params *p, bool want_status, cmcb200_orbit_result *res, bool wait)
{
	OrbitArgs a;
	int n_chunks = 1;
	if (c->up_pending && p->n_local == c->up_n && c->up_chunks > 1) {
		// cmcb200_resident_begin's upload is still in flight: the step runs chunk by chunk behind it (every star is independent of the
		// others, so the result does not depend on the chunking); only the tables must be complete before the first kernel
		CMCB_CUDA_OK(cudaStreamWaitEvent(c->stream, c->ev_tab, 0));
		c->up_defer_join = true;
		const int rc = orbit_begin(c, p, want_status, false, a);
		c->up_defer_join = false;
		if (rc) return -1;
		n_chunks = c->up_chunks;
		for (int k = 0; k < n_chunks; k++) {
			const long off = k * c->up_per, len = p->n_local - off < c->up_per ? p->n_local - off : c->up_per;
			CMCB_CUDA_OK(cudaStreamWaitEvent(c->stream, c->ev_in[k], 0));
			if (orbit_range(c, a, off, len > 0 ? len : 0, k)) return -1;
		}
		CMCB_CUDA_OK(cudaStreamWaitEvent(c->stream, c->ev_bin, 0));
		c->up_pending = false;
	} else {
		if (orbit_begin(c, p, want_status, false, a)) return -1;
		if (orbit_range(c, a, 0, p->n_local, 0)) return -1;
	}
	c->last_orbit_chunks = n_chunks;
	if (!wait) return 0;
	CMCB_CUDA_OK(cudaStreamSynchronize(c->stream));
	orbit_end(c, p, n_chunks, res);
	return 0;
}

extern "C" int cmcb200_dev_new_orbits(cmcb200_ctx *c, const cmcb200_orbit_params *p, int want_status, cmcb200_orbit_result *res)
{
	CMCB_CUDA_OK(cudaSetDevice(c->device));
	return orbit_common(c, p, want_status != 0, res, true);
}

// the same without waiting: the kernels are queued, the statistics are fetched by cmcb200_dev_orbit_result once something else has
// synchronised the stream (or by that call itself)
extern "C" int cmcb200_dev_new_orbits_async(cmcb200_ctx *c, const cmcb200_orbit_params *p, int want_status)
{
	CMCB_CUDA_OK(cudaSetDevice(c->device));
	return orbit_common(c, p, want_status != 0, nullptr, false);
}

extern "C" int cmcb200_dev_orbit_result(cmcb200_ctx *c, const cmcb200_orbit_params *p, cmcb200_orbit_result *res)
{
	CMCB_CUDA_OK(cudaSetDevice(c->device));
	CMCB_CUDA_OK(cudaStreamSynchronize(c->stream));
	orbit_end(c, p, c->last_orbit_chunks > 0 ? c->last_orbit_chunks : 1, res);
	return 0;
}

// energy_conservation2 / ComputeIntermediateEnergy on the resident arrays (orbit.cu: energy_rnew_kernel)
static int energy_rnew(cmcb200_ctx *c, const cmcb200_orbit_params *p, bool uold, bool ei, const char *what)
{
	CMCB_CUDA_OK(cudaSetDevice(c->device));
	OrbitArgs a;
	if (orbit_begin(c, p, false, false, a)) return -1;
	{ StageTimer t(c, ST_ENERGY); c->launches += orbit_energy_rnew(a, uold, ei, c->stream); }
	CMCB_CUDA_OK(cudaGetLastError());
	OrbitStats *h = (OrbitStats *) c->h_pinned;
	CMCB_CUDA_OK(cudaMemcpyAsync(h, c->d_ostats, sizeof(OrbitStats), cudaMemcpyDeviceToHost, c->stream));
	CMCB_CUDA_OK(cudaStreamSynchronize(c->stream));
	if (h->n_status[ORBIT_NTRY]) { // eprintf + exit_cleanly(-2) in the reference (cmc_utils.c:144-152)
		set_error("%s: binary search (FindZero_r) failed for %llu stars", what, h->n_status[ORBIT_NTRY]);
		return -2;
	}
	return 0;
}

extern "C" int cmcb200_dev_energy_conservation2(cmcb200_ctx *c, const cmcb200_orbit_params *p)
{
	return energy_rnew(c, p, true, false, "energy_conservation2");
}

extern "C" int cmcb200_dev_compute_intermediate_energy(cmcb200_ctx *c, const cmcb200_orbit_params *p)
{
	if (energy_rnew(c, p, false, true, "ComputeIntermediateEnergy")) return -1;
	return cmcb200_dev_transfer_new_positions(c, p->n_local, p->n_old_local, p->first_global, p->N_MAX);
}

extern "C" int cmcb200_dev_energy_new_radii(cmcb200_ctx *c, const cmcb200_orbit_params *p)
{
	return energy_rnew(c, p, true, true, "energy_conservation2 + ComputeIntermediateEnergy");
}

extern "C" int cmcb200_dev_read_status(cmcb200_ctx *c, int *status, long *kmin, long *kmax, long n_local)
{
	if (join_uploads(c)) return -1;
	StageTimer t(c, ST_D2H);
	if (status) CMCB_CUDA_OK(cudaMemcpyAsync(status, c->d_status, (size_t) n_local * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
	if (kmin) CMCB_CUDA_OK(cudaMemcpyAsync(kmin, c->d_kmin, (size_t) n_local * sizeof(long), cudaMemcpyDeviceToHost, c->stream));
	if (kmax) CMCB_CUDA_OK(cudaMemcpyAsync(kmax, c->d_kmax, (size_t) n_local * sizeof(long), cudaMemcpyDeviceToHost, c->stream));
	CMCB_CUDA_OK(cudaStreamSynchronize(c->stream));
	return 0;
}

extern "C" int cmcb200_host_register(void *ptr, size_t bytes)
{
	cudaError_t e = cudaHostRegister(ptr, bytes, cudaHostRegisterPortable | cudaHostRegisterMapped);
	if (e == cudaErrorHostMemoryAlreadyRegistered) { cudaGetLastError(); return 0; }
	CMCB_CUDA_OK(e);
	return 0;
}

extern "C" int cmcb200_host_unregister(void *ptr)
{
	cudaError_t e = cudaHostUnregister(ptr);
	if (e == cudaErrorHostMemoryNotRegistered) { cudaGetLastError(); return 0; }
	CMCB_CUDA_OK(e);
	return 0;
}

extern "C" int cmcb200_transfer_bytes(cmcb200_ctx *c, long *h2d, long *d2h, int reset)
{
	if (h2d) *h2d = c->bytes_h2d;
	if (d2h) *d2h = c->bytes_d2h;
	if (reset) c->bytes_h2d = c->bytes_d2h = 0;
	return 0;
}

// Host-buffer form.  Of each 424-byte record the step reads bytes 0..47 (r, vr, vt, m, E, J; rnew at 64 for stars created this step) and
// writes inside bytes 64..127 (rnew, vrnew, vtnew, X, r_peri, r_apo -- which of them depends on the star's outcome; rOld and Y in between
// are not touched).  So only the first 128 bytes of every record travel to the device and bytes 64..127 travel back, as strided (2-D) DMA
// copies: the device copy of bytes 64..127 starts as the host's, the kernels store exactly where the reference's loop body stores, and the
// unconditional copy back therefore changes nothing else.  The star range is cut into chunks that flow through three streams
// (copy in -> kernels -> copy out), so the two copy directions and the kernels overlap.
extern "C" int cmcb200_new_orbits_calculate(cmcb200_ctx *c, void *star1_h, const double *star_r, const double *star_m, const double *star_phi,
					    const cmcb200_orbit_params *p, int *status, long *kmin, long *kmax, cmcb200_orbit_result *res)
{
	CMCB_CUDA_OK(cudaSetDevice(c->device));
	// E and J are the caller's here (the host records hold them; remove_star reads them afterwards): the fused ComputeEnergy belongs to the resident pipeline
	if (p->fuse_compute_energy) { set_error("new_orbits_calculate: fuse_compute_energy is for the resident pipeline (cmcb200_dev_new_orbits) only"); return -1; }
	if (copy_streams(c) || join_uploads(c)) return -1;
	const bool want = status || kmin || kmax;
	const long n = p->n_local;
	int n_chunks = (int) ((n + (1L << 19) - 1) >> 19); // >= 512 K stars per chunk: the kernels of a chunk still fill the GPU
	if (n_chunks > c->orbit_max_chunks) n_chunks = c->orbit_max_chunks;
	if (n_chunks < 1) n_chunks = 1;
	const long per = (n + n_chunks - 1) / n_chunks;
	const size_t pitch = CMCB_SIZEOF_STAR;
	char *dev1 = star1(c);
	// copies in: queued up front on their own stream, behind whatever the context's stream still has in flight (cmcb200_upload is asynchronous)
	CMCB_CUDA_OK(cudaEventRecord(c->ev_done[ORBIT_MAX_CHUNKS - 1], c->stream));
	CMCB_CUDA_OK(cudaStreamWaitEvent(c->s_in, c->ev_done[ORBIT_MAX_CHUNKS - 1], 0));
	for (int k = 0; k < n_chunks; k++) {
		const long off = k * per, len = n - off < per ? n - off : per;
		if (len > 0) CMCB_CUDA_OK(cudaMemcpy2DAsync(dev1 + off * pitch, pitch, (const char *) star1_h + off * pitch, pitch, 128, (size_t) len, cudaMemcpyHostToDevice, c->s_in));
		CMCB_CUDA_OK(cudaEventRecord(c->ev_in[k], c->s_in));
	}
	c->bytes_h2d += n * 128;
	{
		StageTimer t(c, ST_H2D);
		CMCB_CUDA_OK(cudaMemcpyAsync(c->d_r, star_r, (size_t) p->n_table * sizeof(double), cudaMemcpyHostToDevice, c->stream));
		CMCB_CUDA_OK(cudaMemcpyAsync(c->d_m, star_m, (size_t) p->n_table * sizeof(double), cudaMemcpyHostToDevice, c->stream));
		CMCB_CUDA_OK(cudaMemcpyAsync(c->d_phi, star_phi, (size_t) p->n_table * sizeof(double), cudaMemcpyHostToDevice, c->stream));
		c->bytes_h2d += 3 * p->n_table * (long) sizeof(double);
	}
	OrbitArgs a;
	if (orbit_begin(c, p, want, false, a)) return -1;
	for (int k = 0; k < n_chunks; k++) {
		const long off = k * per, len = n - off < per ? n - off : per;
		CMCB_CUDA_OK(cudaStreamWaitEvent(c->stream, c->ev_in[k], 0));
		if (orbit_range(c, a, off, len > 0 ? len : 0, k)) return -1;
		CMCB_CUDA_OK(cudaEventRecord(c->ev_done[k], c->stream));
		CMCB_CUDA_OK(cudaStreamWaitEvent(c->s_out, c->ev_done[k], 0));
		if (len > 0) CMCB_CUDA_OK(cudaMemcpy2DAsync((char *) star1_h + off * pitch + 64, pitch, dev1 + off * pitch + 64, pitch, 64, (size_t) len, cudaMemcpyDeviceToHost, c->s_out));
	}
	c->bytes_d2h += n * 64;
	CMCB_CUDA_OK(cudaStreamSynchronize(c->stream));
	CMCB_CUDA_OK(cudaStreamSynchronize(c->s_out));
	orbit_end(c, p, n_chunks, res);
	if (want) return cmcb200_dev_read_status(c, status, kmin, kmax, p->n_local);
	return 0;
}

// ---- resident-array plumbing ----------------------------------------------------------------------------------------

extern "C" int cmcb200_upload(cmcb200_ctx *c, const void *star1_h, long n_local, const void *binary1_h, long n_bin,
			      const double *star_r, const double *star_m, const double *star_phi, long n_table)
{
	if (n_local + 2 > c->n_star_dim || n_table > c->n_table_dim || n_bin + 1 > c->n_bin_dim) { set_error("upload: sizes exceed the context capacity"); return -1; }
	CMCB_CUDA_OK(cudaSetDevice(c->device));
	if (join_uploads(c)) return -1;
	c->resident_valid = false;
	StageTimer t(c, ST_H2D);
	c->bytes_h2d += (star1_h ? n_local * CMCB_SIZEOF_STAR : 0) + (binary1_h && n_bin > 0 ? n_bin * CMCB_SIZEOF_BINARY : 0)
			+ ((star_r ? 1 : 0) + (star_m ? 1 : 0) + (star_phi ? 1 : 0)) * n_table * (long) sizeof(double);
	if (star1_h) CMCB_CUDA_OK(cudaMemcpyAsync(star1(c), star1_h, (size_t) n_local * CMCB_SIZEOF_STAR, cudaMemcpyHostToDevice, c->stream));
	if (binary1_h && n_bin > 0) CMCB_CUDA_OK(cudaMemcpyAsync(bin1(c), binary1_h, (size_t) n_bin * CMCB_SIZEOF_BINARY, cudaMemcpyHostToDevice, c->stream));
	if (star_r) CMCB_CUDA_OK(cudaMemcpyAsync(c->d_r, star_r, (size_t) n_table * sizeof(double), cudaMemcpyHostToDevice, c->stream));
	if (star_m) CMCB_CUDA_OK(cudaMemcpyAsync(c->d_m, star_m, (size_t) n_table * sizeof(double), cudaMemcpyHostToDevice, c->stream));
	if (star_phi) CMCB_CUDA_OK(cudaMemcpyAsync(c->d_phi, star_phi, (size_t) n_table * sizeof(double), cudaMemcpyHostToDevice, c->stream));
	return 0;
}

extern "C" int cmcb200_download(cmcb200_ctx *c, void *star1_h, long n_local, void *binary1_h, long n_bin,
				double *star_r, double *star_m, double *star_phi, long n_table)
{
	CMCB_CUDA_OK(cudaSetDevice(c->device));
	if (join_uploads(c)) return -1;
	c->bytes_d2h += (star1_h ? n_local * CMCB_SIZEOF_STAR : 0) + (binary1_h && n_bin > 0 ? n_bin * CMCB_SIZEOF_BINARY : 0)
			+ ((star_r ? 1 : 0) + (star_m ? 1 : 0) + (star_phi ? 1 : 0)) * n_table * (long) sizeof(double);
	{
		StageTimer t(c, ST_D2H);
		if (star1_h) CMCB_CUDA_OK(cudaMemcpyAsync(star1_h, star1(c), (size_t) n_local * CMCB_SIZEOF_STAR, cudaMemcpyDeviceToHost, c->stream));
		if (binary1_h && n_bin > 0) CMCB_CUDA_OK(cudaMemcpyAsync(binary1_h, bin1(c), (size_t) n_bin * CMCB_SIZEOF_BINARY, cudaMemcpyDeviceToHost, c->stream));
		if (star_r) CMCB_CUDA_OK(cudaMemcpyAsync(star_r, c->d_r, (size_t) n_table * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
		if (star_m) CMCB_CUDA_OK(cudaMemcpyAsync(star_m, c->d_m, (size_t) n_table * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
		if (star_phi) CMCB_CUDA_OK(cudaMemcpyAsync(star_phi, c->d_phi, (size_t) n_table * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
	}
	CMCB_CUDA_OK(cudaStreamSynchronize(c->stream));
	return 0;
}

// ---- restart file (reference src/cmc/cmc_io.c:3829-3884) written from the resident arrays ---------------------------------------------
// save_restart_file() fwrites curr_st, restart_struct, clus, the WHOLE star[] and binary[] arrays (N_STAR_DIM_OPT / N_BIN_DIM_OPT records,
// :3862-3867) and the snapshot-window counters.  While the records live on the device the host's star[1 .. n] / binary[1 .. nb] are stale:
// those byte ranges are streamed from HBM through two pinned staging buffers (the copy of chunk i + 1 runs while chunk i is written), the
// rest -- the three opaque header blocks, star[0], the unused tail of both arrays, the counters -- comes from the caller's memory.
static int restart_stream(cmcb200_ctx *c, FILE *f, const char *dev, size_t bytes, char *stage[2], size_t chunk, cudaEvent_t ev[2])
{
	size_t off = 0;
	int cur = 0;
	size_t len[2] = {0, 0};
	if (bytes == 0) return 0;
	len[0] = bytes < chunk ? bytes : chunk;
	CMCB_CUDA_OK(cudaMemcpyAsync(stage[0], dev, len[0], cudaMemcpyDeviceToHost, c->stream));
	CMCB_CUDA_OK(cudaEventRecord(ev[0], c->stream));
	while (off < bytes) {
		const size_t next_off = off + len[cur];
		if (next_off < bytes) {
			len[cur ^ 1] = bytes - next_off < chunk ? bytes - next_off : chunk;
			CMCB_CUDA_OK(cudaMemcpyAsync(stage[cur ^ 1], dev + next_off, len[cur ^ 1], cudaMemcpyDeviceToHost, c->stream));
			CMCB_CUDA_OK(cudaEventRecord(ev[cur ^ 1], c->stream));
		}
		CMCB_CUDA_OK(cudaEventSynchronize(ev[cur]));
		if (fwrite(stage[cur], 1, len[cur], f) != len[cur]) { set_error("restart_write: short write"); return -1; }
		off = next_off;
		cur ^= 1;
	}
	c->bytes_d2h += (long) bytes;
	return 0;
}

extern "C" int cmcb200_restart_write(cmcb200_ctx *c, const char *path, const void *rng_state, long rng_bytes, const void *restart_struct, long restart_bytes,
				     const void *clus, long clus_bytes, const void *star0, long n_star_dim_opt, long n_resident, const void *binary0,
				     long n_bin_dim_opt, long n_bin_resident, const int *snapshot_window_counters, long snapshot_window_count)
{
	CMCB_CUDA_OK(cudaSetDevice(c->device));
	if (join_uploads(c)) return -1;
	if (n_resident + 1 > n_star_dim_opt || n_resident + 1 > c->n_star_dim || n_bin_resident + 1 > n_bin_dim_opt || n_bin_resident + 1 > c->n_bin_dim) {
		set_error("restart_write: %ld stars / %ld binaries resident do not fit the arrays (%ld / %ld)", n_resident, n_bin_resident, n_star_dim_opt, n_bin_dim_opt);
		return -1;
	}
	FILE *f = fopen(path, "wb");
	if (!f) { set_error("restart_write: cannot open %s", path); return -1; }
	const size_t chunk = (size_t) 32 << 20;
	char *stage[2] = {nullptr, nullptr};
	cudaEvent_t ev[2];
	int rc = 0;
	if (cudaMallocHost(&stage[0], chunk) != cudaSuccess || cudaMallocHost(&stage[1], chunk) != cudaSuccess) { set_error("restart_write: no pinned staging memory"); rc = -1; }
	cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming);
	cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming);
	auto put = [&](const void *p, size_t n) { if (!rc && n && fwrite(p, 1, n, f) != n) { set_error("restart_write: short write"); rc = -1; } };
	put(rng_state, (size_t) rng_bytes);           // :3862
	put(restart_struct, (size_t) restart_bytes);  // :3863
	put(clus, (size_t) clus_bytes);               // :3864
	put(star0, CMCB_SIZEOF_STAR);                 // :3865: star[0], then the resident records, then the unused tail
	if (!rc) rc = restart_stream(c, f, star1(c), (size_t) n_resident * CMCB_SIZEOF_STAR, stage, chunk, ev);
	put((const char *) star0 + (size_t) (n_resident + 1) * CMCB_SIZEOF_STAR, (size_t) (n_star_dim_opt - n_resident - 1) * CMCB_SIZEOF_STAR);
	put(binary0, CMCB_SIZEOF_BINARY);             // :3866
	if (!rc) rc = restart_stream(c, f, bin1(c), (size_t) n_bin_resident * CMCB_SIZEOF_BINARY, stage, chunk, ev);
	put((const char *) binary0 + (size_t) (n_bin_resident + 1) * CMCB_SIZEOF_BINARY, (size_t) (n_bin_dim_opt - n_bin_resident - 1) * CMCB_SIZEOF_BINARY);
	put(snapshot_window_counters, (size_t) snapshot_window_count * sizeof(int)); // :3867
	if (fclose(f) != 0 && !rc) { set_error("restart_write: close failed"); rc = -1; }
	cudaEventDestroy(ev[0]); cudaEventDestroy(ev[1]);
	if (stage[0]) cudaFreeHost(stage[0]);
	if (stage[1]) cudaFreeHost(stage[1]);
	return rc;
}

// ---- resident timestep on host arrays: the records cross the host link once each way ---------------------------------------------

// Runs of consecutive dirty 8-byte words of a record (bit w of the mask = word w); returns the number of runs
static int dirty_runs(unsigned long long mask, int lo[CMCB_STAR_WORDS], int len[CMCB_STAR_WORDS])
{
	int n = 0;
	for (int w = 0; w < CMCB_STAR_WORDS;) {
		if (!((mask >> w) & 1ull)) { w++; continue; }
		int e = w;
		while (e < CMCB_STAR_WORDS && ((mask >> e) & 1ull)) e++;
		lo[n] = w; len[n] = e - w; n++;
		w = e;
	}
	return n;
}

static int resident_begin_impl(cmcb200_ctx *c, const void *star1_h, long n_local, const void *binary1_h, long n_bin, const double *star_r, const double *star_m,
			       const double *star_phi, long n_table, unsigned long long dirty_words, long slice_first, long slice_n);

extern "C" int cmcb200_resident_begin(cmcb200_ctx *c, const void *star1_h, long n_local, const void *binary1_h, long n_bin,
					      const double *star_r, const double *star_m, const double *star_phi, long n_table, unsigned long long dirty_words)
{
	return resident_begin_impl(c, star1_h, n_local, binary1_h, n_bin, star_r, star_m, star_phi, n_table, dirty_words, 0, -1);
}

// Radius-sharded form: every rank holds the same replicated host tables, so each one sends only ITS slice [slice_first, slice_first + slice_n) of
// star_r / star_m / star_phi (plus the sentinel [0] and everything behind the last rank's slice: N_MAX + 1 and the stars created this step)
// over its host link and the ranks hand the slices to each other over NVLink (peer stores + a device-side barrier) -- (P - 1) / P of the
// table upload leaves the host links.
extern "C" int cmcb200_resident_begin_sharded(cmcb200_ctx *c, const void *star1_h, long n_local, const void *binary1_h, long n_bin, const double *star_r,
						      const double *star_m, const double *star_phi, long n_table, unsigned long long dirty_words, long slice_first, long slice_n)
{
	if (c->nranks <= 1 || !star_r || !star_m || !star_phi) slice_n = -1;
	return resident_begin_impl(c, star1_h, n_local, binary1_h, n_bin, star_r, star_m, star_phi, n_table, dirty_words, slice_first, slice_n);
}

static int resident_begin_impl(cmcb200_ctx *c, const void *star1_h, long n_local, const void *binary1_h, long n_bin, const double *star_r, const double *star_m,
			       const double *star_phi, long n_table, unsigned long long dirty_words, long slice_first, long slice_n)
{
	if (n_local + 2 > c->n_star_dim || n_table > c->n_table_dim || n_bin + 1 > c->n_bin_dim) { set_error("resident_begin: sizes exceed the context capacity"); return -1; }
	CMCB_CUDA_OK(cudaSetDevice(c->device));
	if (copy_streams(c) || join_uploads(c)) return -1;
	const unsigned long long all = (1ull << CMCB_STAR_WORDS) - 1ull;
	dirty_words &= all;
	// a partial upload patches the copy the last cmcb200_resident_end left behind: same records, same order
	if (!(c->resident_valid && c->resident_n == n_local)) dirty_words = all;
	c->resident_valid = false;
	// the copies start behind whatever the context's stream still has in flight
	CMCB_CUDA_OK(cudaEventRecord(c->ev_done[ORBIT_MAX_CHUNKS - 1], c->stream));
	CMCB_CUDA_OK(cudaStreamWaitEvent(c->s_in, c->ev_done[ORBIT_MAX_CHUNKS - 1], 0));
	if (slice_n >= 0) {
		// this rank's slice and the sentinel [0]; the caller extends the LAST rank's slice to n_table - 1 (N_MAX + 1 and the stars created this
		// step lie behind the last shard), so the slices of all ranks cover [1, n_table)
		if (slice_first < 1 || slice_first + slice_n > n_table) { set_error("resident_begin_sharded: slice [%ld, %ld) outside the tables", slice_first, slice_first + slice_n); return -1; }
		const double *src[3] = {star_r, star_m, star_phi};
		double *dst[3] = {c->d_r, c->d_m, c->d_phi};
		for (int t = 0; t < 3; t++) {
			CMCB_CUDA_OK(cudaMemcpyAsync(dst[t], src[t], sizeof(double), cudaMemcpyHostToDevice, c->s_in));
			if (slice_n > 0) CMCB_CUDA_OK(cudaMemcpyAsync(dst[t] + slice_first, src[t] + slice_first, (size_t) slice_n * sizeof(double), cudaMemcpyHostToDevice, c->s_in));
		}
		c->bytes_h2d += 3 * (slice_n + 1) * (long) sizeof(double);
	} else {
		if (star_r) CMCB_CUDA_OK(cudaMemcpyAsync(c->d_r, star_r, (size_t) n_table * sizeof(double), cudaMemcpyHostToDevice, c->s_in));
		if (star_m) CMCB_CUDA_OK(cudaMemcpyAsync(c->d_m, star_m, (size_t) n_table * sizeof(double), cudaMemcpyHostToDevice, c->s_in));
		if (star_phi) CMCB_CUDA_OK(cudaMemcpyAsync(c->d_phi, star_phi, (size_t) n_table * sizeof(double), cudaMemcpyHostToDevice, c->s_in));
		c->bytes_h2d += ((star_r ? 1 : 0) + (star_m ? 1 : 0) + (star_phi ? 1 : 0)) * n_table * (long) sizeof(double);
	}
	CMCB_CUDA_OK(cudaEventRecord(c->ev_tab, c->s_in));
	if (slice_n >= 0) { // the slices travel on to the peers behind the upload, on the context's stream; the kernels queued later find complete tables
		CMCB_CUDA_OK(cudaStreamWaitEvent(c->stream, c->ev_tab, 0));
		if (shard_share_table_slices(c, slice_first, slice_n)) return -1;
	}
	int n_chunks = (int) ((n_local + (1L << 19) - 1) >> 19); // >= 512 K stars per chunk: the kernels of a chunk still fill the GPU
	if (n_chunks > c->orbit_max_chunks) n_chunks = c->orbit_max_chunks;
	if (n_chunks < 1) n_chunks = 1;
	const long per = (n_local + n_chunks - 1) / n_chunks;
	int lo[CMCB_STAR_WORDS], len[CMCB_STAR_WORDS];
	const int n_runs = dirty_words == all ? 0 : dirty_runs(dirty_words, lo, len);
	char *dev1 = star1(c);
	const size_t pitch = CMCB_SIZEOF_STAR;
	for (int k = 0; k < n_chunks; k++) {
		const long off = k * per, n = n_local - off < per ? n_local - off : per;
		if (n > 0 && star1_h) {
			if (dirty_words == all) {
				CMCB_CUDA_OK(cudaMemcpyAsync(dev1 + off * pitch, (const char *) star1_h + off * pitch, (size_t) n * pitch, cudaMemcpyHostToDevice, c->s_in));
				c->bytes_h2d += n * (long) pitch;
			} else {
				for (int r = 0; r < n_runs; r++) { // strided (2-D) DMA of each run of dirty words
					CMCB_CUDA_OK(cudaMemcpy2DAsync(dev1 + off * pitch + lo[r] * 8, pitch, (const char *) star1_h + off * pitch + lo[r] * 8, pitch, (size_t) len[r] * 8,
								       (size_t) n, cudaMemcpyHostToDevice, c->s_in));
					c->bytes_h2d += n * (long) len[r] * 8;
				}
			}
		}
		CMCB_CUDA_OK(cudaEventRecord(c->ev_in[k], c->s_in));
	}
	if (binary1_h && n_bin > 0) {
		CMCB_CUDA_OK(cudaMemcpyAsync(bin1(c), binary1_h, (size_t) n_bin * CMCB_SIZEOF_BINARY, cudaMemcpyHostToDevice, c->s_in));
		c->bytes_h2d += n_bin * CMCB_SIZEOF_BINARY;
	}
	CMCB_CUDA_OK(cudaEventRecord(c->ev_bin, c->s_in));
	c->up_pending = true;
	c->up_chunks = n_chunks;
	c->up_per = per;
	c->up_n = n_local;
	return 0;
}

extern "C" int cmcb200_resident_end(cmcb200_ctx *c, void *star1_h, long n_local, void *binary1_h, long n_bin, long n_bin_dim,
					    double *star_r, double *star_m, double *star_phi, long n_table)
{
	if (cmcb200_download(c, star1_h, n_local, binary1_h, n_bin, star_r, star_m, star_phi, n_table)) return -1;
	// memset(b_buf, 0, (N_BIN_DIM_OPT-1)*sizeof(binary_t)) (cmc_sort.c:541): everything past the packed binaries is zero
	if (binary1_h && n_bin_dim - 1 > n_bin) memset((char *) binary1_h + (size_t) n_bin * CMCB_SIZEOF_BINARY, 0, (size_t) (n_bin_dim - 1 - n_bin) * CMCB_SIZEOF_BINARY);
	c->resident_valid = star1_h != nullptr;
	c->resident_n = n_local;
	return 0;
}

extern "C" int cmcb200_dev_get(cmcb200_ctx *c, void *star_rec, long j, void *binary_rec, long b)
{
	if (j < 1 || j + 1 > c->n_star_dim || (binary_rec && (b < 1 || b >= c->n_bin_dim))) { set_error("dev_get: index out of range"); return -1; }
	CMCB_CUDA_OK(cudaSetDevice(c->device));
	if (join_uploads(c)) return -1;
	if (star_rec) CMCB_CUDA_OK(cudaMemcpyAsync(star_rec, star1(c) + (j - 1) * CMCB_SIZEOF_STAR, CMCB_SIZEOF_STAR, cudaMemcpyDeviceToHost, c->stream));
	if (binary_rec) CMCB_CUDA_OK(cudaMemcpyAsync(binary_rec, bin1(c) + (b - 1) * CMCB_SIZEOF_BINARY, CMCB_SIZEOF_BINARY, cudaMemcpyDeviceToHost, c->stream));
	CMCB_CUDA_OK(cudaStreamSynchronize(c->stream));
	return 0;
}

extern "C" int cmcb200_dev_put(cmcb200_ctx *c, const void *star_rec, long j, const void *binary_rec, long b, const double *star_m_g, long g)
{
	if (j < 1 || j + 1 > c->n_star_dim || (binary_rec && (b < 1 || b >= c->n_bin_dim)) || (star_m_g && (g < 0 || g >= c->n_table_dim))) { set_error("dev_put: index out of range"); return -1; }
	CMCB_CUDA_OK(cudaSetDevice(c->device));
	if (join_uploads(c)) return -1;
	if (star_rec) CMCB_CUDA_OK(cudaMemcpyAsync(star1(c) + (j - 1) * CMCB_SIZEOF_STAR, star_rec, CMCB_SIZEOF_STAR, cudaMemcpyHostToDevice, c->stream));
	if (binary_rec) CMCB_CUDA_OK(cudaMemcpyAsync(bin1(c) + (b - 1) * CMCB_SIZEOF_BINARY, binary_rec, CMCB_SIZEOF_BINARY, cudaMemcpyHostToDevice, c->stream));
	if (star_m_g) CMCB_CUDA_OK(cudaMemcpyAsync(c->d_m + g, star_m_g, sizeof(double), cudaMemcpyHostToDevice, c->stream));
	CMCB_CUDA_OK(cudaStreamSynchronize(c->stream));
	return 0;
}

// ---- the two per-star elementwise steps that chain the pipeline on the device -----------------------------------------

namespace cmcb {

// ComputeIntermediateEnergy's transfer loop (reference src/cmc/cmc_utils.c:380-390)
__global__ void transfer_new_positions_kernel(char *star1, long n_local, long n_old_local, long first_global, long N_MAX,
					      const double *__restrict__ star_r, const double *__restrict__ star_m)
{
	long j0 = blockIdx.x * (long) blockDim.x + threadIdx.x;
	if (j0 >= n_local) return;
	long j = j0 + 1;
	long g = j > n_old_local ? 1 + N_MAX + (j - n_old_local) : first_global + j - 1;
	char *rec = star1 + j0 * CMCB_SIZEOF_STAR;
	*(double *) (rec + CMCB_OFF_ROLD) = star_r[g];
	*(double *) (rec + CMCB_OFF_R) = *(const double *) (rec + CMCB_OFF_RNEW);
	*(double *) (rec + CMCB_OFF_M) = star_m[g];
	*(double *) (rec + CMCB_OFF_VR) = *(const double *) (rec + CMCB_OFF_VRNEW);
	*(double *) (rec + CMCB_OFF_VT) = *(const double *) (rec + CMCB_OFF_VTNEW);
}

// first loop of ComputeEnergy (reference src/cmc/cmc_utils.c:612-616)
__global__ void compute_energy_kernel(char *star1, long n_local, long first_global, const double *__restrict__ star_r, const double *__restrict__ star_phi)
{
	long j0 = blockIdx.x * (long) blockDim.x + threadIdx.x;
	if (j0 >= n_local) return;
	long g = first_global + j0;
	char *rec = star1 + j0 * CMCB_SIZEOF_STAR;
	double vr = *(const double *) (rec + CMCB_OFF_VR), vt = *(const double *) (rec + CMCB_OFF_VT);
	*(double *) (rec + CMCB_OFF_E) = dadd(star_phi[g], dmul(0.5, dadd(dmul(vr, vr), dmul(vt, vt))));
	*(double *) (rec + CMCB_OFF_J) = dmul(star_r[g], vt);
}

} // namespace cmcb

extern "C" int cmcb200_dev_transfer_new_positions(cmcb200_ctx *c, long n_local, long n_old_local, long first_global, long N_MAX)
{
	CMCB_CUDA_OK(cudaSetDevice(c->device));
	if (join_uploads(c)) return -1;
	StageTimer t(c, ST_TRANSFER);
	if (n_local > 0)
	{
		transfer_new_positions_kernel<<<(unsigned) ((n_local + 255) / 256), 256, 0, c->stream>>>(star1(c), n_local, n_old_local, first_global, N_MAX, c->d_r, c->d_m);
		c->launches++;
	}
	CMCB_CUDA_OK(cudaGetLastError());
	return 0;
}

extern "C" int cmcb200_dev_compute_energy(cmcb200_ctx *c, long n_local, long first_global)
{
	CMCB_CUDA_OK(cudaSetDevice(c->device));
	if (join_uploads(c)) return -1;
	StageTimer t(c, ST_ENERGY);
	if (n_local > 0)
	{
		compute_energy_kernel<<<(unsigned) ((n_local + 255) / 256), 256, 0, c->stream>>>(star1(c), n_local, first_global, c->d_r, c->d_phi);
		c->launches++;
	}
	CMCB_CUDA_OK(cudaGetLastError());
	return 0;
}

// The context behind the C ABI (include/cmcb200.h): device-resident mirrors, scratch, stage profiling.  Shared by api.cu and shard.cu.
#pragma once
#include <vector>
#include "../../include/cmcb200.h"
#include "common.cuh"
#include "orbit.cuh"
#include "potential.cuh"
#include "sort.cuh"
#include "star_layout.h"

using namespace cmcb;

constexpr int SHARD_MAX_RANKS = 16;
enum { SH_STAR0 = 0, SH_STAR1, SH_BIN0, SH_BIN1, SH_R, SH_M, SH_R_ALT, SH_M_ALT, SH_KEYS_ALL, SH_SRC_ALL, SH_SYNC, SH_PHI, SH_RANK_ALL, SH_M_ALL, SH_N_BUFFERS }; // peer-visible buffers

enum { ST_KEYS = 0, ST_PASSES, ST_GATHER, ST_BINARIES, ST_POTENTIAL, ST_ORBIT, ST_TRANSFER, ST_ENERGY, ST_H2D, ST_D2H, ST_ORBIT_SCAN, ST_ORBIT_TRIES, ST_ORBIT_TABLES, ST_SHARD_KEYS, ST_SHARD_RANK, ST_SHARD_TABLES, ST_LAGRANGE, ST_DYNAMICS, ST_SIGMA };

struct StageEvent {
	int stage;
	cudaEvent_t a, b;
};

struct cmcb200_ctx {
	int device;
	long n_star_dim, n_bin_dim;
	long n_table_dim; // capacity of star_r / star_m / star_phi (the whole cluster; n_star_dim is this rank's share of the records)
	cudaStream_t stream;
	bool own_stream;
	// resident arrays.  d_star[cur] holds the live star array: record 0 = sentinel (star[0]), star1 = base + 424.
	char *d_star[2];
	char *d_bin[2]; // record 0 = binary[0] (unused by the reference), binary1 = base + 464
	int cur_star, cur_bin;
	double *d_r, *d_m, *d_phi, *d_mprev;
	double *d_r_alt, *d_m_alt; // second copy of star_r / star_m: the fused transfer+sort reads the old tables while it writes the new ones
	RPhi *d_rp;       // packed (r, phi) lookup table of the orbit step
	RPhiF *d_rpf;     // its FP32 screening copy (orbit_core.h: sign_Q_screened)
	long *d_defer_star, *d_defer_found;
	unsigned *d_defer_first;
	LadderBuffers lad; // coarse levels of the lookup tables (orbit_core.h "ladders")
	unsigned char *d_oflag;
	double *d_job;     // [6][n_star_dim] jobs of the pending stars (orbit.cuh: OrbitArgs::job)
	int *d_pend[2];
	int orbit_max_chunks; // host-buffer orbit step: at most this many pipeline chunks (CMCB200_ORBIT_CHUNKS, default 8)
	int orbit_ctas;    // tuning knob CMCB200_ORBIT_CTAS: resident CTAs per SM orbit_prepare_kernel is compiled for (4 or 3)
	int l2_persist_pct;    // CMCB200_L2_PERSIST: > 0 = fraction (percent) of the (r, Phi) table pinned in L2's persisting carve-out during the orbit step
	size_t l2_window_max;  // cudaDeviceProp.accessPolicyMaxWindowSize
	int orbit_ladders; // CMCB200_ORBIT_LADDERS: bit 3 = sample ladder for find_zero_Q (default on), bit 0 = r ladder for potential() (default on), bit 2 = FP32 screening table for find_zero_Q (default off: measured no faster), bit 1 = Q ladder for find_zero_Q (default off:
	                   // measured slower than the plain bisection on B200); 0 = every probe reads the fine table
	// sort scratch
	SortBuffers sbuf;
	SortState *d_sort;
	unsigned *d_tile_hist;
	long tile_hist_words;
	unsigned char *d_binflag;
	unsigned *d_bin_block_cnt;
	// potential scratch
	PotScalars *d_pot;
	void *d_seqwork;
	SeqWork seqwork;
	// orbit scratch
	OrbitStats *d_ostats;
	void *d_lagrange; // device result block of comp_mass_percent (lagrange.cu), allocated on first use
	void *d_multimass; // argument / result block of comp_multi_mass_percent (lagrange.cu), allocated on first use
	void *d_seqwork2; // exact-scan scratch for chains of 2 n_star_dim elements (ComputeEnergy's potential sum), allocated on first use
	void *d_dynamics; // scratch of dynamics_apply's pair loop (dynamics.cu), a host struct holding device buffers, allocated on first use
	unsigned long long *d_dtmin; // result word of simul_relax_new (dynamics.cu), allocated on first use
	void *d_sigma;    // scratch of calc_sigma_r (sigma.cu), likewise
	void *d_sv3;      // scratch of set_velocities3 (velocities.cu), allocated on first use for n_star_dim stars
	int *d_status;
	long *d_kmin, *d_kmax;
	double *d_draws;
	long *d_draw_off;
	long draws_cap, draw_off_cap;
	// pinned scalars for read-back
	void *h_pinned;
	// profiling
	bool profiling;
	std::vector<StageEvent> events;
	std::vector<cudaEvent_t> pool;
	double acc_ms[CMCB200_N_STAGES];
	long acc_n[CMCB200_N_STAGES];
	long launches; // kernels launched since creation
	bool step_emit_keys;   // cmcb200_dev_step: the orbit step writes the sort keys (default on; CMCB200_STEP_KEYS=0 for A/B)
	bool orbit_emit_keys;  // set around the orbit step of cmcb200_dev_step only: every other caller leaves OrbitArgs::sort_keys null
	int last_orbit_chunks; // chunks of the last orbit step (their statistics wait in the pinned block for cmcb200_dev_orbit_result)
	long bytes_h2d, bytes_d2h; // bytes the host-buffer entry points moved over PCIe (cmcb200_transfer_bytes)
	// host-buffer orbit step: copy-in and copy-out streams and per-chunk events of the pipeline (created on first use)
	cudaStream_t s_in, s_out;
	cudaEvent_t ev_in[16], ev_done[16];
	// resident timestep on host arrays (cmcb200_resident_begin / _end, api.cu): the upload runs in row chunks on s_in.  up_pending: the
	// context's stream has not yet been ordered behind all of it (join_uploads); up_defer_join: the chunk-wise orbit step orders itself
	bool up_pending, up_defer_join;
	int up_chunks;
	long up_per, up_n;
	cudaEvent_t ev_tab, ev_bin;
	bool resident_valid; // the device records mirror the host's records [1 .. resident_n] as of the last cmcb200_resident_end
	long resident_n;
	// radius-sharded operation (shard.cu): this context is rank `rank` of `nranks` GPUs of one node
	int rank, nranks;
	void *own_buf[SH_N_BUFFERS];                    // this rank's peer-visible buffers, in the order of the SH_* enum
	void *peer_buf[SHARD_MAX_RANKS][SH_N_BUFFERS];  // the same buffers of every rank, mapped with cudaIpcOpenMemHandle ([rank] = own pointers)
	uint64_t *d_keys_all;                           // [nranks][keys_stride] sorted keys of every rank (pushed by their owners)
	unsigned char *d_src_all;                       // [n_star_dim] rank each record of the new shard came from (written by the pushing ranks)
	long keys_stride;
	unsigned *d_dest_pos;                           // [n] final local position of each local sorted element on its destination rank
	unsigned char *d_dest_rank;                     // [n] destination rank
	long *d_shard_meta;                             // device copy of (n_keep[q], start0[q]) per rank
	unsigned *d_rank_all;                           // [nranks][keys_stride] global rank of every element of every rank's sorted run (pushed by their owners)
	double *d_m_all;                                // [nranks][keys_stride] star_m of the same elements
	cudaStream_t s_aux;                             // second stream of the asynchronous sharded step: tables + potential while the records travel
	cudaEvent_t ev_fork, ev_join;
	bool pot_queued;                                // the last asynchronous step queued the potential rebuild behind its tables
	void *d_shard_sync;                             // ShardSync (shard.cu): flags and posted counts the peers store into -- the device-side barriers
	unsigned long long shard_epoch;                 // barrier generation, advanced identically on every rank
	unsigned long long table_epoch;                 // generation of the table-sharing barrier (phase 3) of cmcb200_resident_begin_sharded
	int shard_unfused;                              // 1: the sharded sort keys on r and moves the records unchanged (cmcb200_shard_fuse_transfer)
	int r_parity;                                   // which of (d_r, d_r_alt) is the buffer SH_R (flips with every fused sort)
};

// shard.cu: this rank's slice [first, first + n) of star_r / star_m / star_phi stored into every peer's tables, then a device-side barrier --
// queued on the context's stream (cmcb200_resident_begin_sharded)
int shard_share_table_slices(cmcb200_ctx *c, long first, long n);
// api.cu: the potential rebuild queued on `stream` (no wait, no read-back); n_max_upper bounds the new N_MAX
int potential_enqueue(cmcb200_ctx *c, long n_max_upper, long N_STAR, double cenma_m, double cenma_m_new, cudaStream_t stream);

// dynamics.cu / sigma.cu: release the scratch behind d_dynamics / d_sigma
void dynamics_free(cmcb200_ctx *c);
void sigma_free(cmcb200_ctx *c);

// fills the orbit-step argument block from the context and the caller's parameters and (re)builds the lookup tables (api.cu)
int orbit_begin(cmcb200_ctx *c, const cmcb200_orbit_params *p, bool want_status, bool need_status, OrbitArgs &a);

// every entry point that works on the resident arrays orders the context's stream behind an upload cmcb200_resident_begin left in flight
static inline int join_uploads(cmcb200_ctx *c)
{
	if (!c->up_pending || c->up_defer_join) return 0;
	CMCB_CUDA_OK(cudaStreamWaitEvent(c->stream, c->ev_tab, 0));
	for (int k = 0; k < c->up_chunks; k++) CMCB_CUDA_OK(cudaStreamWaitEvent(c->stream, c->ev_in[k], 0));
	CMCB_CUDA_OK(cudaStreamWaitEvent(c->stream, c->ev_bin, 0));
	c->up_pending = false;
	return 0;
}

static inline char *star1(cmcb200_ctx *c) { return c->d_star[c->cur_star] + CMCB_SIZEOF_STAR; }
static inline char *bin1(cmcb200_ctx *c) { return c->d_bin[c->cur_bin] + CMCB_SIZEOF_BINARY; }

namespace {

struct StageTimer { // brackets a stage with events when profiling is on
	cmcb200_ctx *c;
	int idx;
	StageTimer(cmcb200_ctx *ctx, int stage) : c(ctx), idx(-1)
	{
		if (!c->profiling) return;
		StageEvent e;
		e.stage = stage;
		for (cudaEvent_t *p : {&e.a, &e.b}) {
			if (!c->pool.empty()) { *p = c->pool.back(); c->pool.pop_back(); }
			else cudaEventCreate(p);
		}
		cudaEventRecord(e.a, c->stream);
		c->events.push_back(e);
		idx = (int) c->events.size() - 1;
	}
	~StageTimer()
	{
		if (idx >= 0) cudaEventRecord(c->events[idx].b, c->stream);
	}
};

inline int ensure(void **p, long *cap, long need, size_t elem)
{
	if (need <= *cap) return 0;
	if (*p) cudaFree(*p);
	*p = nullptr;
	long ncap = need + need / 4 + 1024;
	CMCB_CUDA_OK(cudaMalloc(p, (size_t) ncap * elem));
	*cap = ncap;
	return 0;
}

} // namespace


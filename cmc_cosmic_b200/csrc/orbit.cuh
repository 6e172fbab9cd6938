#pragma once
#include "common.cuh"
#include "orbit_core.h"
#include "star_layout.h"
#include "sort.cuh"

namespace cmcb {

constexpr int ORBIT_THREADS = 256;          // persistent CTAs: each stages the coarsest ladder level in shared memory once
constexpr int ORBIT_CTAS_PER_SM = 4;        // 1024 threads and <= 160 KB of staged tables per SM
constexpr int ORBIT_SCAN_THREADS = 256;
constexpr int ORBIT_SCAN_CHUNK = 4096;      // table entries per work item of the cooperative linear scan
constexpr int ORBIT_DEFER_CAP = 65536;      // queued scans per launch (beyond it a star scans in its own thread)
constexpr long ORBIT_NOT_FOUND = 0x7fffffffffffffffL;
constexpr int ORBIT_TRY_THREADS = 256;
constexpr int ORBIT_MAX_ROUNDS = 8;         // launches of the rejection-sampling stage

struct OrbitStats { // device-resident, zeroed per launch
	unsigned long long max_rad_bits; // max over ORBIT_OK stars of rnew (positive doubles order like their bit patterns)
	unsigned long long n_draws, n_exact, n_brent;
	unsigned long long n_status[6];
	unsigned n_defer;     // stars queued for the cooperative linear scan
	unsigned round_count[ORBIT_MAX_ROUNDS + 2]; // stars pending at the start of each rejection round
};

struct OrbitArgs {
	char *star1;        // &star[1] on the device (AoS, reference layout)
	long n_local;       // clus.N_MAX_NEW: local stars, including stars created this step
	long n_old_local;   // End[myid]-Start[myid]+1: stars beyond are new (get_global_idx, cmc_utils.c:2384-2407)
	long first_global;  // Start[myid]
	OrbitTables t;      // device tables
	double Rtidal;
	// draws: either a recorded stream in CSR form (parity mode) ...
	const double *draws;
	const long *draw_off; // [n_local + 1]
	// ... or a counter-based generator keyed by (seed, tcount, global star index) (production mode, draws == nullptr)
	unsigned long long seed;
	unsigned long long tcount;
	// optional per-star diagnostics (may be null)
	int *status;
	long *kmin, *kmax;
	int *n_draws;
	OrbitStats *stats;
	unsigned long long *sort_keys; // [n_local] or null: key_from_double(star[j].rnew) as the step leaves it -- the radix sort that follows (cmcb200_dev_step)
	                               // then needs no pass over the 424-byte records to collect its keys
	unsigned char *flag; // [n_local] 1: orbit known, new radius still to be drawn
	// the pending stars' jobs as five compact arrays (component c of local star j0 at job[c * job_stride + j0]: envelope F, r_min, r_max, E, J) --
	// stage A writes them coalesced and the rejection rounds read them coalesced instead of five scattered fields of the 424-byte records
	double *job;
	long job_stride;
	bool split_lookup; // the rejection rounds run as lookup kernel + try kernel (CMCB200_ORBIT_LADDERS bit 5, off: measured slower); a sixth job component holds potential(r)
	int *pend_list[2]; // [n_local] each: compacted lists of pending stars, ping-pong between rounds
	bool fuse_energy;  // E, J of every star are computed from the table and (vr, vt) on the fly and stored (ComputeEnergy's first loop)
	const int *lad_unsorted; // device flag set by the table builder when star_r is not sorted (or Phi holds a NaN): ladders off
	int prepare_ctas;  // register-cap variant of orbit_prepare_kernel: 4 (default) or 3 resident CTAs per SM (CMCB200_ORBIT_CTAS)
	unsigned q_top_bytes, r_top_bytes; // size of the coarsest level of each ladder (what a CTA stages), multiples of 16
	unsigned s_topf_bytes;             // bytes of the sample ladder's coarsest level rounded to FP32 (t.slad.topf; 0: no FP32 screening)
	unsigned s_top_bytes;              // the same for the sample ladder (it shares the Q ladder's buffer: one of the two is in use)
	long *defer_star;  // [ORBIT_DEFER_CAP] local 0-based index of each queued star
	unsigned *defer_first; // [ORBIT_DEFER_CAP] lowest 256-entry sub-chunk of the table the bound could not prove free of Q > 0 (0xffffffff: none)
	long *defer_found; // [ORBIT_DEFER_CAP] first k with Q > 0 (ORBIT_NOT_FOUND: none)
};

// device storage of the ladders (sized for the table capacity at context creation)
struct LadderBuffers {
	LBlock *q;     // all levels of the Q ladder back to back, finest first
	double *r;     // all levels of the r ladder back to back, finest first
	long q_cap, r_cap; // capacity in blocks / doubles
	int *unsorted; // device flag
	bool r_whole_lines; // the r ladder's lines are read whole (tuning knob: CMCB200_ORBIT_LADDERS bit 4)
	bool sampled;  // build the sample ladder (in q's memory) instead of the Q ladder
	bool f32_top;  // with it, an FP32 copy of its coarsest level for FP32 screening of that level's probes (CMCB200_ORBIT_LADDERS bit 6, default on)
	int q_max_shift, q_max_width_log2, q_factor_log2; // tuning knobs of the Q ladder (CMCB200_QLAD_MAX_SHIFT / _MAX_WIDTH_LOG2 / _FACTOR_LOG2)
};
long ladder_q_capacity(long n_table_dim);
long ladder_r_capacity(long n_table_dim);
// packs (r, Phi) of entries [0, n) and builds both ladders over [0, M]; fills t.rp-independent ladder descriptors of `a`
int orbit_build_tables_device(const double *r, const double *phi, RPhi *rp, RPhiF *rpf, long n, long M, const LadderBuffers &lb, OrbitArgs &a, cudaStream_t s);
int orbit_stage_prepare(const OrbitArgs &a, cudaStream_t s);
int orbit_stage_scan(const OrbitArgs &a, cudaStream_t s);
int orbit_stage_tries(const OrbitArgs &a, cudaStream_t s);
int orbit_energy_rnew(const OrbitArgs &a, bool uold, bool ei, cudaStream_t s);
int orbit_reset_stats(const OrbitArgs &a, cudaStream_t s); // zeroes a.stats (one launch)

// ---- device helpers shared by orbit.cu and velocities.cu ----------------------------------------------------------------
__device__ __forceinline__ double ld_field(const char *rec, int off) { return *(const double *) (rec + off); }
__device__ __forceinline__ void st_field(char *rec, int off, double v) { *(double *) (rec + off) = v; }

// Stages `bytes` of a ladder's coarsest level into this CTA's shared memory with one TMA bulk copy and returns the shared address
// (nullptr when the ladder is off).  Called by all threads of the CTA, once.
__device__ __forceinline__ const void *stage_top_level(const void *level, unsigned bytes, bool on, unsigned char *smem, unsigned long long *bar)
{
	if (!on) return nullptr; // CTA-uniform
	if (threadIdx.x == 0) mbar_init(bar, 1);
	__syncthreads();
	if (threadIdx.x == 0) bulk_copy_g2s(smem, level, bytes, bar);
	mbar_wait(bar, 0);
	return smem;
}


} // namespace cmcb

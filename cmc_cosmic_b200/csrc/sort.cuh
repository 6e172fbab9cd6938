#pragma once
#include "common.cuh"
#include "star_layout.h"

namespace cmcb {

// 512 threads x 8 items: the same 4096-pair tile as 256 x 16, but the per-thread arrays fit in 64 registers, so two CTAs = 32 warps are
// resident per SM instead of 16 (the pass is bound by the latency of its look-back and shared-memory phases, not by HBM)
constexpr int SORT_THREADS = 256;
constexpr int SORT_WARPS = SORT_THREADS / 32;
constexpr int SORT_ITEMS = 16;
constexpr int SORT_TILE = SORT_THREADS * SORT_ITEMS; // 4096 (key, index) pairs per CTA iteration
constexpr int SORT_CTAS_PER_SM = 2;

#ifdef __CUDACC__
// total order of IEEE doubles as unsigned integers: the radix sort's key (also written by the orbit step, OrbitArgs::sort_keys)
__device__ __forceinline__ uint64_t key_from_double(double r)
{
	uint64_t b = (uint64_t) __double_as_longlong(r);
	return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
#endif

struct SortState { // device-resident bookkeeping of one sort
	unsigned hist[8][256];    // digit histograms of the 8 LSD passes
	unsigned base[8][256];    // exclusive scans of them
	int skip[8];              // pass is the identity (all keys share the digit)
	int src_buf[8];           // buffer (0 = A, 1 = B) the pass reads
	int final_buf;            // buffer holding the sorted permutation
	unsigned tile_counter[8]; // dynamic tile ids
	long n_eq_inf, n_gt_inf;  // keys == / > SF_INFINITY
	long n_total;             // records sorted
	long n_new;               // records kept after trimming stripped stars
	long n_bin;               // binaries after compaction
};

struct SortBuffers {
	uint64_t *keys_a, *keys_b;
	unsigned *vals_a, *vals_b;
};

struct SortArgs {
	const char *star_in; // &star[1] (device), n records
	char *star_out;      // &star_sorted[1]
	long n;
	long key_off;        // byte offset of the sort key inside star_t (r)
	SortBuffers buf;
	SortState *st;
	unsigned *tile_hist; // [8][ntiles][256] look-back words
	double *star_r, *star_m; // SoA tables (index 0 = sentinel), written at [1..n_new]
	unsigned char *binflag;  // [n]
	unsigned *bin_block_cnt; // [ceil(n/2048)]
	const char *bin_in;  // &binary[1] before the sort (may be null)
	char *bin_out;       // &binary[1] after
	long n_bin_clear;    // zero bin_out records [n_bin, n_bin_clear)
	// fused ComputeIntermediateEnergy transfer (reference src/cmc/cmc_utils.c:380-390): the key is rnew, and while a record moves
	// r <- rnew, vr <- vrnew, vt <- vtnew, m <- star_m[g], rOld <- star_r[g] (g = the record's global index BEFORE the sort)
	bool keys_ready;     // buf.keys_a already holds the keys of the n records (the orbit step wrote them, cmcb200_dev_step): no extraction pass
	bool keep_keys;      // materialise the sorted keys too (the sharded sort ranks them against the other ranks' keys)
	bool fuse_transfer;
	const double *old_r, *old_m; // the tables the transfer reads (must not alias star_r / star_m, which the gather writes)
	long n_old_local, first_global, N_MAX;
};

int sort_setup();
size_t sort_smem_bytes();
int sort_stage_keys(const SortArgs &a, cudaStream_t s);
int sort_stage_passes(const SortArgs &a, cudaStream_t s);
int sort_stage_gather(const SortArgs &a, cudaStream_t s);
int sort_stage_binaries(const SortArgs &a, cudaStream_t s);
int sort_binary_count_scan(const unsigned char *binflag, SortState *st, unsigned *bin_block_cnt, long n, cudaStream_t s);
int sort_binary_count_scan_planned(const unsigned char *binflag, SortState *st, unsigned *bin_block_cnt, long n_upper, cudaStream_t s);
int sort_binary_clear_device(char *bin_out, const SortState *st, long n_clear, cudaStream_t s);

} // namespace cmcb

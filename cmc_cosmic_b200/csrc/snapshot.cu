// The rows of the reference's HDF5 snapshot table built from the RESIDENT arrays: write_snapshot()'s loop over the stars
// (reference src/cmc/cmc_io.c:3176-3327) as a kernel -- one thread per row, snapshot_core.h holds the row itself -- streamed to the caller's
// buffer in chunks.  The HDF5 calls that store the rows (H5TBmake_table on rank 0, H5TBappend_records on the others, :3333-3343) stay in
// the reference: it hands them this buffer instead of the one its own loop filled.
#include <algorithm>
#include <cstring>
#include "../../include/cmcb200.h"
#include "common.cuh"
#include "ctx.cuh"
#include "snapshot_core.h"

using namespace cmcb;

namespace {

constexpr long SNAP_CHUNK_ROWS = 1L << 18; // 128 MB of rows per device buffer

// rows [0, cnt) of this launch: local star list[k] (bh_only) or first + k
__global__ void __launch_bounds__(256) snapshot_rows_kernel(SnapArgs a, const long *__restrict__ list, long first, long cnt, uint64_t *__restrict__ rows)
{
	for (long k = blockIdx.x * 256L + threadIdx.x; k < cnt; k += gridDim.x * 256L)
		snap_row(a, list ? list[k] : first + k, rows + k * SNAP_ROW_WORDS);
}

// bh_only: the stars write_snapshot keeps (:3170-3171), in any order -- the caller sorts the list
__global__ void __launch_bounds__(256) snapshot_select_bh_kernel(SnapArgs a, long *__restrict__ list, unsigned long long *count)
{
	for (long i = 1 + blockIdx.x * 256L + threadIdx.x; i <= a.n_local; i += gridDim.x * 256L)
		if (snap_is_bh(a, i)) list[atomicAdd(count, 1ull)] = i;
}

unsigned blocks_for(long n) { long b = (n + 255) / 256; return (unsigned) (b < 1 ? 1 : (b > NUM_SMS_B200 * 8L ? NUM_SMS_B200 * 8L : b)); }

} // namespace

extern "C" int cmcb200_dev_snapshot_rows(cmcb200_ctx *c, const cmcb200_snapshot_params *p, const double *sigma, void *rows_out, long cap_rows, long *n_rows)
{
	CMCB_CUDA_OK(cudaSetDevice(c->device));
	if (join_uploads(c)) return -1;
	if (p->n_local + 2 > c->n_star_dim) { set_error("snapshot_rows: n_local + 2 = %ld exceeds n_star_dim = %ld", p->n_local + 2, c->n_star_dim); return -1; }
	if (p->avekernel < 1) { set_error("snapshot_rows: AVEKERNEL = %ld", p->avekernel); return -1; }
	{ // the highest table entry a row may read: stars created this step, and on P > 1 ranks the unclamped mass window (snapshot_core.h)
		const long over = (p->n_local > p->n_old_local ? p->n_local - p->n_old_local : 0) + (p->n_old_local < p->N_MAX ? 2 * p->avekernel : 0);
		if (1 + p->N_MAX + over >= c->n_table_dim) { set_error("snapshot_rows: table entry %ld beyond the capacity %ld", 1 + p->N_MAX + over, c->n_table_dim); return -1; }
	}
	double *d_sigma_tmp = nullptr;
	long *d_list = nullptr;
	uint64_t *d_rows = nullptr;
	unsigned long long *d_count = nullptr;
	long *h_list = nullptr;
	int rc = -1;
	SnapArgs a;
	a.star1 = star1(c); a.bin0 = c->d_bin[c->cur_bin];
	a.star_r = c->d_r; a.star_m = c->d_m; a.star_phi = c->d_phi;
	a.sigma = nullptr;
	a.n_local = p->n_local; a.n_old_local = p->n_old_local; a.first_global = p->first_global; a.N_MAX = p->N_MAX; a.N_STAR = p->N_STAR;
	a.units_m = p->units_m; a.units_l = p->units_l; a.madhoc = p->madhoc; a.avekernel = p->avekernel;
	memcpy(&a.nan_bits, &p->nan_value, 8);
	long n = p->n_local;
	do {
		if (sigma) { // sigma_array.sigma of the caller, [n_local + 1]
			if (cudaMalloc((void **) &d_sigma_tmp, (size_t) (p->n_local + 1) * sizeof(double)) != cudaSuccess) { set_error("snapshot_rows: out of device memory (sigma)"); break; }
			if (cudaMemcpyAsync(d_sigma_tmp, sigma, (size_t) (p->n_local + 1) * sizeof(double), cudaMemcpyHostToDevice, c->stream) != cudaSuccess) { set_error("snapshot_rows: sigma upload failed"); break; }
			a.sigma = d_sigma_tmp;
		} else if (c->d_sigma) { // what cmcb200_dev_calc_sigma_r left on the device
			const double *d_a, *d_s;
			if (cmcb200_dev_sigma_arrays(c, &d_a, &d_s)) break;
			a.sigma = d_s;
		}
		if (p->bh_only && p->n_local > 0) {
			if (cudaMalloc((void **) &d_list, (size_t) p->n_local * sizeof(long)) != cudaSuccess || cudaMalloc((void **) &d_count, 8) != cudaSuccess) { set_error("snapshot_rows: out of device memory (list)"); break; }
			cudaMemsetAsync(d_count, 0, 8, c->stream);
			snapshot_select_bh_kernel<<<blocks_for(p->n_local), 256, 0, c->stream>>>(a, d_list, d_count);
			c->launches++;
			unsigned long long cnt = 0;
			if (cudaMemcpyAsync(&cnt, d_count, 8, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess || cudaStreamSynchronize(c->stream) != cudaSuccess) { set_error("snapshot_rows: %s", cudaGetErrorString(cudaGetLastError())); break; }
			n = (long) cnt;
			if (n > 0) { // star order, like the reference's loop
				h_list = (long *) malloc((size_t) n * sizeof(long));
				if (!h_list) { set_error("snapshot_rows: out of host memory"); break; }
				cudaMemcpy(h_list, d_list, (size_t) n * sizeof(long), cudaMemcpyDeviceToHost);
				std::sort(h_list, h_list + n);
				cudaMemcpy(d_list, h_list, (size_t) n * sizeof(long), cudaMemcpyHostToDevice);
			}
		}
		*n_rows = n;
		if (n > cap_rows) { set_error("snapshot_rows: %ld rows, the caller's buffer holds %ld", n, cap_rows); rc = -3; break; }
		if (n > 0) {
			const long chunk = n < SNAP_CHUNK_ROWS ? n : SNAP_CHUNK_ROWS;
			if (cudaMalloc((void **) &d_rows, (size_t) chunk * SNAP_ROW_BYTES) != cudaSuccess) { set_error("snapshot_rows: out of device memory (rows)"); break; }
			bool ok = true;
			for (long off = 0; off < n && ok; off += chunk) {
				const long cnt = n - off < chunk ? n - off : chunk;
				snapshot_rows_kernel<<<blocks_for(cnt), 256, 0, c->stream>>>(a, p->bh_only ? d_list + off : nullptr, 1 + off, cnt, d_rows);
				c->launches++;
				ok = cudaMemcpyAsync((char *) rows_out + off * SNAP_ROW_BYTES, d_rows, (size_t) cnt * SNAP_ROW_BYTES, cudaMemcpyDeviceToHost, c->stream) == cudaSuccess
				     && cudaStreamSynchronize(c->stream) == cudaSuccess; // one device buffer: the next chunk overwrites it
				c->bytes_d2h += cnt * SNAP_ROW_BYTES;
			}
			if (!ok) { set_error("snapshot_rows: %s", cudaGetErrorString(cudaGetLastError())); break; }
		}
		rc = 0;
	} while (0);
	cudaStreamSynchronize(c->stream);
	cudaFree(d_sigma_tmp); cudaFree(d_list); cudaFree(d_rows); cudaFree(d_count); free(h_list);
	return rc;
}

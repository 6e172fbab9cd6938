// Probe for the column-group pipeline of the resident timestep: 424-byte records in pinned host memory, a byte range [off, off + w) of every
// record copied host->device / device->host (a) by a kernel through the mapped alias, (b) by the DMA engine (cudaMemcpy2DAsync), alone and
// with the two directions running at once; then the whole schedule (hot group up, kernels, cold groups up while sorted groups go down).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/pcie_probe2 tools/pcie_probe2.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <cuda_runtime.h>
#define OK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)
constexpr long REC = 424;

// words [w0, w0 + nw) of records [0, n): dst and src both laid out as 424-byte records
__global__ void copy_cols(const unsigned long long *__restrict__ src, unsigned long long *__restrict__ dst, long n, int w0, int nw)
{
	const long total = n * nw, stride = (long) gridDim.x * blockDim.x;
	long t = blockIdx.x * (long) blockDim.x + threadIdx.x;
	for (; t + 3 * stride < total; t += 4 * stride) {
		unsigned long long v[4]; long a[4];
#pragma unroll
		for (int u = 0; u < 4; u++) { long tt = t + u * stride; long i = tt / nw; int j = (int) (tt - i * nw); a[u] = i * (REC / 8) + w0 + j; v[u] = src[a[u]]; }
#pragma unroll
		for (int u = 0; u < 4; u++) dst[a[u]] = v[u];
	}
	for (; t < total; t += stride) { long i = t / nw; int j = (int) (t - i * nw); long a = i * (REC / 8) + w0 + j; dst[a] = src[a]; }
}

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main(int argc, char **argv)
{
	long N = argc > 1 ? atol(argv[1]) : 10000000;
	size_t bytes = (size_t) N * REC + 64;
	char *h, *h2, *d, *d2;
	OK(cudaSetDeviceFlags(cudaDeviceMapHost));
	OK(cudaMallocHost(&h, bytes)); OK(cudaMallocHost(&h2, bytes));
	memset(h, 1, bytes); memset(h2, 2, bytes);
	OK(cudaMalloc(&d, bytes)); OK(cudaMalloc(&d2, bytes));
	OK(cudaMemset(d, 0, bytes)); OK(cudaMemset(d2, 0, bytes));
	cudaStream_t s1, s2, s3; OK(cudaStreamCreate(&s1)); OK(cudaStreamCreate(&s2)); OK(cudaStreamCreate(&s3));
	// records start 8 bytes into the allocation: only 8-byte alignment, like &star[1]
	unsigned long long *hw = (unsigned long long *) (h + 8), *hw2 = (unsigned long long *) (h2 + 8), *dw = (unsigned long long *) (d + 8), *dw2 = (unsigned long long *) (d2 + 8);
	auto timeit = [&](auto f, int reps = 3) { f(); OK(cudaDeviceSynchronize()); double t0 = now(); for (int i = 0; i < reps; i++) f(); OK(cudaDeviceSynchronize()); return (now() - t0) / reps; };
	const int grids[] = {148 * 2, 148 * 8};
	const int widths[][2] = {{0, 128}, {0, 160}, {128, 148}, {128, 296}, {160, 88}, {160, 264}, {0, 424}};
	for (auto &wd : widths) {
		int w0 = wd[0] / 8, nw = wd[1] / 8;
		double gb = (double) N * wd[1] / 1e9;
		for (int g : grids) {
			double tu = timeit([&] { copy_cols<<<g, 512, 0, s1>>>(hw, dw, N, w0, nw); });
			double td = timeit([&] { copy_cols<<<g, 512, 0, s2>>>(dw2, hw2, N, w0, nw); });
			double tb = timeit([&] { copy_cols<<<g, 512, 0, s1>>>(hw, dw, N, w0, nw); copy_cols<<<g, 512, 0, s2>>>(dw2, hw2, N, w0, nw); });
			printf("kernel  off %3d w %3d grid %4d: up %5.1f GB/s (%5.1f ms)  down %5.1f GB/s (%5.1f ms)  both at once %5.1f ms\n", wd[0], wd[1], g, gb / tu, tu * 1e3, gb / td, td * 1e3, tb * 1e3);
		}
		double tu = timeit([&] { OK(cudaMemcpy2DAsync(d + 8 + wd[0], REC, h + 8 + wd[0], REC, wd[1], N, cudaMemcpyHostToDevice, s1)); });
		double td = timeit([&] { OK(cudaMemcpy2DAsync(h2 + 8 + wd[0], REC, d2 + 8 + wd[0], REC, wd[1], N, cudaMemcpyDeviceToHost, s2)); });
		double tp = timeit([&] { OK(cudaMemcpy2DAsync(h2 + 8 + wd[0], REC, d2, wd[1], wd[1], N, cudaMemcpyDeviceToHost, s2)); });
		double tb = timeit([&] { OK(cudaMemcpy2DAsync(d + 8 + wd[0], REC, h + 8 + wd[0], REC, wd[1], N, cudaMemcpyHostToDevice, s1));
					 OK(cudaMemcpy2DAsync(h2 + 8 + wd[0], REC, d2 + 8 + wd[0], REC, wd[1], N, cudaMemcpyDeviceToHost, s2)); });
		printf("2-D DMA off %3d w %3d          : up %5.1f GB/s (%5.1f ms)  down %5.1f GB/s (%5.1f ms)  down from packed %5.1f ms  both at once %5.1f ms\n", wd[0], wd[1], gb / tu, tu * 1e3, gb / td, td * 1e3, tp * 1e3, tb * 1e3);
	}
	// the schedule: groups G0 = [0,160) hot, then cold groups; group g goes down (from d2, standing for the sorted array) once it is up and the
	// permutation is known (a 12 ms stand-in kernel wait after the hot group)
	for (int ncold = 1; ncold <= 4; ncold++) for (int mode = 0; mode < 2; mode++) {
		int off[6], wid[6], ng = 1 + ncold;
		off[0] = 0; wid[0] = 160;
		int rest = 264 / 8, o = 160;
		for (int g = 1; g < ng; g++) { int w = (rest / ncold + (g <= rest % ncold ? 1 : 0)) * 8; off[g] = o; wid[g] = w; o += w; }
		cudaEvent_t up[6]; for (int g = 0; g < ng; g++) OK(cudaEventCreateWithFlags(&up[g], cudaEventDisableTiming));
		auto run = [&] {
			for (int g = 0; g < ng; g++) {
				if (mode == 0) copy_cols<<<148 * 8, 512, 0, s1>>>(hw, dw, N, off[g] / 8, wid[g] / 8);
				else OK(cudaMemcpy2DAsync(d + 8 + off[g], REC, h + 8 + off[g], REC, wid[g], N, cudaMemcpyHostToDevice, s1));
				OK(cudaEventRecord(up[g], s1));
			}
			// downloads: cold groups in order as they arrive, the hot group last (it is still being modified by the later passes)
			for (int k = 1; k <= ng; k++) {
				int g = k % ng;
				OK(cudaStreamWaitEvent(s2, up[g == 0 ? ng - 1 : g], 0));
				if (mode == 0) copy_cols<<<148 * 8, 512, 0, s2>>>(dw2, hw2, N, off[g] / 8, wid[g] / 8);
				else OK(cudaMemcpy2DAsync(h2 + 8 + off[g], REC, d2 + 8 + off[g], REC, wid[g], N, cudaMemcpyDeviceToHost, s2));
			}
		};
		double t = timeit(run, 2);
		printf("schedule %s, hot 160 B + %d cold groups, hot group down last: %6.1f ms\n", mode ? "2-D DMA" : "kernels", ncold, t * 1e3);
	}
	double t = timeit([&] { OK(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, s1)); OK(cudaMemcpyAsync(h2, d2, bytes, cudaMemcpyDeviceToHost, s1)); }, 2);
	printf("whole records up then down (one stream): %6.1f ms\n", t * 1e3);
	return 0;
}

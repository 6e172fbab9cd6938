/* TEST INFRASTRUCTURE ONLY (oracle/): single-node mini-MPI for running the reference's
 * hot-path sources as P processes on the host cores (CPU baseline + multi-rank parity).
 * Implements exactly the calls that path reaches (call sites: reference
 * src/cmc/cmc_sort.c:401-607,694-750; src/cmc/cmc_utils.c:217,1686-1704,2206-2207,2236-2253,2279-2283).
 * Ranks are forked processes (the reference keeps all state in process globals);
 * data moves through per-rank windows in one MAP_SHARED region.
 * With no mini_mpi_spawn() call there is one rank and every collective is a local copy. */
#define _GNU_SOURCE
#include "mpi.h"
#include <sched.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/wait.h>
#include <time.h>
#include <unistd.h>

#define MM_MAX_RANKS 256
#define MM_MAIL_BYTES 1024

typedef struct {
	volatile int bar_count;
	volatile int bar_sense;
	volatile int abort_flag;
	int nprocs;
	size_t window_bytes;
	/* per-rank published metadata for v-collectives */
	int counts[MM_MAX_RANKS][MM_MAX_RANKS];
	int displs[MM_MAX_RANKS][MM_MAX_RANKS];
	size_t nbytes[MM_MAX_RANKS];
	/* point-to-point: one buffered mailbox per (destination, source) pair -- comp_mass_percent sends 4 doubles to rank 0, set_velocities3
	 * one, calc_sigma_r 2 * AVEKERNEL doubles of ghost velocities to both neighbours BEFORE it receives (so a send must not wait for its
	 * receive).  The region is an anonymous mapping: pages nobody writes cost nothing. */
	volatile int mail_full[MM_MAX_RANKS][MM_MAX_RANKS];
	char mail[MM_MAX_RANKS][MM_MAX_RANKS][MM_MAIL_BYTES];
} mm_ctrl_t;

static mm_ctrl_t *mm = NULL;
static char *mm_windows = NULL;
static int mm_rank = 0, mm_size = 1, mm_local_sense = 0;
static pid_t mm_children[MM_MAX_RANKS];
static char mm_self_mail[MM_MAIL_BYTES];

static char *win(int r) { return mm_windows + (size_t) r * mm->window_bytes; }

static void mm_barrier(void)
{
	if (mm_size == 1) return;
	mm_local_sense = !mm_local_sense;
	if (__sync_add_and_fetch(&mm->bar_count, 1) == mm_size) {
		mm->bar_count = 0;
		__sync_synchronize();
		mm->bar_sense = mm_local_sense;
	} else {
		while (mm->bar_sense != mm_local_sense) {
			if (mm->abort_flag) _exit(3);
			sched_yield();
		}
	}
	__sync_synchronize();
}

static void mm_need(size_t bytes, const char *what)
{
	if (mm_size > 1 && bytes > mm->window_bytes) {
		fprintf(stderr, "mini_mpi: %s needs %zu B but the rank window is %zu B\n", what, bytes, mm->window_bytes);
		mm->abort_flag = 1;
		_exit(4);
	}
}

int mini_mpi_spawn(int nprocs, size_t window_bytes)
{
	if (nprocs < 1 || nprocs > MM_MAX_RANKS) return -1;
	if (nprocs == 1) return 0;
	size_t ctrl = (sizeof(mm_ctrl_t) + 4095) & ~(size_t) 4095;
	window_bytes = (window_bytes + 4095) & ~(size_t) 4095;
	char *base = mmap(NULL, ctrl + window_bytes * nprocs, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
	if (base == MAP_FAILED) { perror("mini_mpi mmap"); return -1; }
	mm = (mm_ctrl_t *) base; /* zero-filled by the kernel */
	mm->nprocs = nprocs;
	mm->window_bytes = window_bytes;
	mm_windows = base + ctrl;
	mm_size = nprocs;
	fflush(NULL);
	for (int r = 1; r < nprocs; r++) {
		pid_t p = fork();
		if (p < 0) { perror("mini_mpi fork"); return -1; }
		if (p == 0) { mm_rank = r; return 0; }
		mm_children[r] = p;
	}
	mm_rank = 0;
	return 0;
}

int mini_mpi_join(void)
{
	if (mm_size == 1) return 0;
	fflush(NULL);
	if (mm_rank != 0) _exit(0);
	int worst = 0;
	for (int r = 1; r < mm_size; r++) {
		int st = 0;
		waitpid(mm_children[r], &st, 0);
		if (!WIFEXITED(st) || WEXITSTATUS(st) != 0) worst = 1;
	}
	mm_size = 1;
	return worst;
}

int MPI_Init(int *argc, char ***argv) { (void) argc; (void) argv; return 0; }
int MPI_Finalize(void) { return 0; }
int MPI_Abort(MPI_Comm c, int code) { (void) c; if (mm) mm->abort_flag = 1; fflush(NULL); _exit(code ? code : 1); }
int MPI_Comm_size(MPI_Comm c, int *n) { (void) c; *n = mm_size; return 0; }
int MPI_Comm_rank(MPI_Comm c, int *r) { (void) c; *r = mm_rank; return 0; }
int MPI_Barrier(MPI_Comm c) { (void) c; mm_barrier(); return 0; }
double MPI_Wtime(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }

int MPI_Type_contiguous(int n, MPI_Datatype old, MPI_Datatype *newt)
{ *newt = (MINI_MPI_KIND_DERIVED << 24) | (n * MINI_MPI_SIZE(old)); return 0; }
int MPI_Type_commit(MPI_Datatype *t) { (void) t; return 0; }
int MPI_Type_free(MPI_Datatype *t) { *t = MPI_DATATYPE_NULL; return 0; }
int MPI_Type_size(MPI_Datatype t, int *sz) { *sz = MINI_MPI_SIZE(t); return 0; }

static void reduce_into(void *acc, const void *in, int n, MPI_Datatype t, MPI_Op op)
{
	int kind = MINI_MPI_KIND(t);
#define RED(T) do { T *a = (T *) acc; const T *b = (const T *) in; for (int i = 0; i < n; i++) { \
	if (op == MPI_SUM) a[i] += b[i]; else if (op == MPI_MAX) { if (b[i] > a[i]) a[i] = b[i]; } else { if (b[i] < a[i]) a[i] = b[i]; } } } while (0)
	if (kind == MINI_MPI_KIND(MPI_INT)) RED(int);
	else if (kind == MINI_MPI_KIND(MPI_LONG) || kind == MINI_MPI_KIND(MPI_LONG_LONG)) RED(long);
	else if (kind == MINI_MPI_KIND(MPI_DOUBLE)) RED(double);
	else if (kind == MINI_MPI_KIND(MPI_FLOAT)) RED(float);
	else if (kind == MINI_MPI_KIND(MPI_UNSIGNED)) RED(unsigned);
	else { fprintf(stderr, "mini_mpi: reduce on unsupported datatype %d\n", t); MPI_Abort(0, 5); }
#undef RED
}

/* rank order 0..P-1, like a linear reduce; every rank evaluates the same order, so all results are identical */
int MPI_Allreduce(const void *s, void *r, int n, MPI_Datatype t, MPI_Op op, MPI_Comm c)
{
	(void) c;
	size_t bytes = (size_t) n * MINI_MPI_SIZE(t);
	if (mm_size == 1) { if (s != MPI_IN_PLACE) memcpy(r, s, bytes); return 0; }
	mm_need(bytes, "Allreduce");
	memcpy(win(mm_rank), s == MPI_IN_PLACE ? r : s, bytes);
	mm_barrier();
	memcpy(r, win(0), bytes);
	for (int q = 1; q < mm_size; q++) reduce_into(r, win(q), n, t, op);
	mm_barrier();
	return 0;
}

/* the result lands on every rank (a superset of what MPI_Reduce promises the root) */
int MPI_Reduce(const void *s, void *r, int n, MPI_Datatype t, MPI_Op op, int root, MPI_Comm c)
{
	(void) root;
	return MPI_Allreduce(s, r, n, t, op, c);
}

int MPI_Bcast(void *b, int n, MPI_Datatype t, int root, MPI_Comm c)
{
	(void) c;
	size_t bytes = (size_t) n * MINI_MPI_SIZE(t);
	if (mm_size == 1) return 0;
	mm_need(bytes, "Bcast");
	if (mm_rank == root) memcpy(win(root), b, bytes);
	mm_barrier();
	if (mm_rank != root) memcpy(b, win(root), bytes);
	mm_barrier();
	return 0;
}

int MPI_Allgatherv(const void *s, int sn, MPI_Datatype st, void *r, const int *rn, const int *disp, MPI_Datatype rt, MPI_Comm c)
{
	(void) c;
	size_t esz = MINI_MPI_SIZE(rt);
	size_t sbytes = (size_t) sn * MINI_MPI_SIZE(st);
	if (mm_size == 1) { memmove((char *) r + (size_t) disp[0] * esz, s, sbytes); return 0; }
	mm_need(sbytes, "Allgather(v)");
	memcpy(win(mm_rank), s, sbytes);
	mm_barrier();
	for (int q = 0; q < mm_size; q++)
		memcpy((char *) r + (size_t) disp[q] * esz, win(q), (size_t) rn[q] * esz);
	mm_barrier();
	return 0;
}

int MPI_Allgather(const void *s, int sn, MPI_Datatype st, void *r, int rn, MPI_Datatype rt, MPI_Comm c)
{
	int cnt[MM_MAX_RANKS], dsp[MM_MAX_RANKS];
	for (int q = 0; q < mm_size; q++) { cnt[q] = rn; dsp[q] = q * rn; }
	return MPI_Allgatherv(s, sn, st, r, cnt, dsp, rt, c);
}

int MPI_Gather(const void *s, int sn, MPI_Datatype st, void *r, int rn, MPI_Datatype rt, int root, MPI_Comm c)
{
	(void) c;
	size_t sbytes = (size_t) sn * MINI_MPI_SIZE(st), rbytes = (size_t) rn * MINI_MPI_SIZE(rt);
	if (mm_size == 1) { memmove(r, s, sbytes); return 0; }
	mm_need(sbytes, "Gather");
	memcpy(win(mm_rank), s, sbytes);
	mm_barrier();
	if (mm_rank == root)
		for (int q = 0; q < mm_size; q++) memcpy((char *) r + q * rbytes, win(q), rbytes);
	mm_barrier();
	return 0;
}

int MPI_Alltoallv(const void *s, const int *sn, const int *sd, MPI_Datatype st, void *r, const int *rn, const int *rd, MPI_Datatype rt, MPI_Comm c)
{
	(void) c;
	size_t ssz = MINI_MPI_SIZE(st), rsz = MINI_MPI_SIZE(rt);
	if (mm_size == 1) { memmove((char *) r + (size_t) rd[0] * rsz, (const char *) s + (size_t) sd[0] * ssz, (size_t) sn[0] * ssz); return 0; }
	/* stage the used extent of the send buffer, publish counts/displs, let receivers pull */
	size_t lo = (size_t) -1, hi = 0;
	for (int q = 0; q < mm_size; q++) {
		mm->counts[mm_rank][q] = sn[q];
		if (sn[q] == 0) continue;
		size_t a = (size_t) sd[q] * ssz, b = a + (size_t) sn[q] * ssz;
		if (a < lo) lo = a;
		if (b > hi) hi = b;
	}
	if (hi < lo) lo = hi = 0;
	mm_need(hi - lo, "Alltoallv");
	for (int q = 0; q < mm_size; q++) mm->displs[mm_rank][q] = sn[q] ? (int) (((size_t) sd[q] * ssz - lo) / ssz) : 0;
	memcpy(win(mm_rank), (const char *) s + lo, hi - lo);
	mm_barrier();
	for (int q = 0; q < mm_size; q++) {
		int cnt = mm->counts[q][mm_rank];
		if (cnt != rn[q]) { fprintf(stderr, "mini_mpi: Alltoallv count mismatch %d<-%d: %d vs %d\n", mm_rank, q, rn[q], cnt); MPI_Abort(0, 6); }
		memcpy((char *) r + (size_t) rd[q] * rsz, win(q) + (size_t) mm->displs[q][mm_rank] * ssz, (size_t) cnt * ssz);
	}
	mm_barrier();
	return 0;
}

int MPI_Alltoall(const void *s, int sn, MPI_Datatype st, void *r, int rn, MPI_Datatype rt, MPI_Comm c)
{
	int scnt[MM_MAX_RANKS], sdsp[MM_MAX_RANKS], rcnt[MM_MAX_RANKS], rdsp[MM_MAX_RANKS];
	for (int q = 0; q < mm_size; q++) { scnt[q] = sn; sdsp[q] = q * sn; rcnt[q] = rn; rdsp[q] = q * rn; }
	return MPI_Alltoallv(s, scnt, sdsp, st, r, rcnt, rdsp, rt, c);
}

/* MPI_Exscan (reference src/cmc/cmc_utils.c:925): rank i receives op(send_0, .., send_{i-1}), combined in rank order; rank 0's
 * receive buffer is left untouched (undefined in MPI).  Only MPI_SUM over doubles is reached. */
int MPI_Exscan(const void *s, void *r, int n, MPI_Datatype t, MPI_Op op, MPI_Comm c)
{
	(void) c; (void) op;
	if (mm_size == 1) return 0;
	size_t bytes = (size_t) n * MINI_MPI_SIZE(t);
	mm_need(bytes, "MPI_Exscan");
	memcpy(win(mm_rank), s, bytes);
	mm_barrier();
	if (mm_rank > 0) {
		double *out = (double *) r;
		for (int i = 0; i < n; i++) out[i] = ((const double *) win(0))[i];
		for (int q = 1; q < mm_rank; q++)
			for (int i = 0; i < n; i++) out[i] += ((const double *) win(q))[i];
	}
	mm_barrier();
	return 0;
}

int MPI_Send(const void *b, int n, MPI_Datatype t, int dst, int tag, MPI_Comm c)
{
	(void) tag; (void) c;
	size_t bytes = (size_t) n * MINI_MPI_SIZE(t);
	if (bytes > MM_MAIL_BYTES) { fprintf(stderr, "mini_mpi: MPI_Send of %zu bytes not supported\n", bytes); _exit(5); }
	if (mm_size == 1) { memcpy(mm_self_mail, b, bytes); return 0; } /* a rank sending to itself (calc_sigma_r's ring on one rank) */
	while (mm->mail_full[dst][mm_rank]) { if (mm->abort_flag) _exit(3); sched_yield(); }
	memcpy(mm->mail[dst][mm_rank], b, bytes);
	__sync_synchronize();
	mm->mail_full[dst][mm_rank] = 1;
	return 0;
}

int MPI_Recv(void *b, int n, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Status *st)
{
	(void) tag; (void) c; (void) st;
	size_t bytes = (size_t) n * MINI_MPI_SIZE(t);
	if (bytes > MM_MAIL_BYTES) { fprintf(stderr, "mini_mpi: MPI_Recv of %zu bytes not supported\n", bytes); _exit(5); }
	if (mm_size == 1) { memcpy(b, mm_self_mail, bytes); return 0; }
	while (!mm->mail_full[mm_rank][src]) { if (mm->abort_flag) _exit(3); sched_yield(); }
	__sync_synchronize();
	memcpy(b, mm->mail[mm_rank][src], bytes);
	__sync_synchronize();
	mm->mail_full[mm_rank][src] = 0;
	return 0;
}

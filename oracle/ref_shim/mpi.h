/* TEST INFRASTRUCTURE ONLY (oracle/): minimal <mpi.h> stand-in so the reference's own
 * hot-path sources compile unmodified where they lie under /root/reference.
 * Semantics are implemented in mini_mpi.c (fork + POSIX shared memory; single node). */
#ifndef CMCB200_SHIM_MPI_H
#define CMCB200_SHIM_MPI_H
#include <stddef.h>
typedef int MPI_Comm;
typedef int MPI_Datatype;   /* value = element size in bytes */
typedef int MPI_Op;
typedef int MPI_Group;
typedef int MPI_Info;
typedef long long MPI_Offset;
typedef struct { int fd; } *MPI_File;
typedef struct { int MPI_SOURCE, MPI_TAG, MPI_ERROR; } MPI_Status;
typedef int MPI_Request;
#define MPI_COMM_WORLD 0
#define MPI_COMM_NULL (-1)
#define MPI_SUCCESS 0
#define MPI_IN_PLACE ((void *)1)
#define MPI_STATUS_IGNORE ((MPI_Status *)0)
#define MPI_STATUSES_IGNORE ((MPI_Status *)0)
#define MPI_INFO_NULL 0
#define MPI_DATATYPE_NULL 0
/* a datatype is (kind << 24) | size_in_bytes; kind selects the reduce arithmetic */
#define MINI_MPI_KIND(t) ((t) >> 24)
#define MINI_MPI_SIZE(t) ((t) & 0xffffff)
#define MPI_BYTE      ((1 << 24) | 1)
#define MPI_CHAR      ((2 << 24) | 1)
#define MPI_INT       ((3 << 24) | 4)
#define MPI_LONG      ((4 << 24) | 8)
#define MPI_LONG_LONG ((5 << 24) | 8)
#define MPI_DOUBLE    ((6 << 24) | 8)
#define MPI_FLOAT     ((7 << 24) | 4)
#define MPI_UNSIGNED  ((8 << 24) | 4)
#define MINI_MPI_KIND_DERIVED 9
#define MPI_SUM 1
#define MPI_MAX 2
#define MPI_MIN 3
#define MPI_MODE_CREATE 1
#define MPI_MODE_WRONLY 2
#define MPI_MODE_RDWR 4
#define MPI_MODE_APPEND 8
#define MPI_SEEK_SET 0
int MPI_Init(int *argc, char ***argv);
int MPI_Finalize(void);
int MPI_Abort(MPI_Comm c, int code);
int MPI_Comm_size(MPI_Comm c, int *n);
int MPI_Comm_rank(MPI_Comm c, int *r);
int MPI_Barrier(MPI_Comm c);
double MPI_Wtime(void);
int MPI_Allreduce(const void *s, void *r, int n, MPI_Datatype t, MPI_Op op, MPI_Comm c);
int MPI_Reduce(const void *s, void *r, int n, MPI_Datatype t, MPI_Op op, int root, MPI_Comm c);
int MPI_Reduce(const void *s, void *r, int n, MPI_Datatype t, MPI_Op op, int root, MPI_Comm c);
int MPI_Exscan(const void *s, void *r, int n, MPI_Datatype t, MPI_Op op, MPI_Comm c);
int MPI_Bcast(void *b, int n, MPI_Datatype t, int root, MPI_Comm c);
int MPI_Gather(const void *s, int sn, MPI_Datatype st, void *r, int rn, MPI_Datatype rt, int root, MPI_Comm c);
int MPI_Gatherv(const void *s, int sn, MPI_Datatype st, void *r, const int *rn, const int *disp, MPI_Datatype rt, int root, MPI_Comm c);
int MPI_Scatterv(const void *s, const int *sn, const int *disp, MPI_Datatype st, void *r, int rn, MPI_Datatype rt, int root, MPI_Comm c);
int MPI_Allgather(const void *s, int sn, MPI_Datatype st, void *r, int rn, MPI_Datatype rt, MPI_Comm c);
int MPI_Allgatherv(const void *s, int sn, MPI_Datatype st, void *r, const int *rn, const int *disp, MPI_Datatype rt, MPI_Comm c);
int MPI_Alltoall(const void *s, int sn, MPI_Datatype st, void *r, int rn, MPI_Datatype rt, MPI_Comm c);
int MPI_Alltoallv(const void *s, const int *sn, const int *sd, MPI_Datatype st, void *r, const int *rn, const int *rd, MPI_Datatype rt, MPI_Comm c);
int MPI_Send(const void *b, int n, MPI_Datatype t, int dst, int tag, MPI_Comm c);
int MPI_Recv(void *b, int n, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Status *st);
int MPI_Isend(const void *b, int n, MPI_Datatype t, int dst, int tag, MPI_Comm c, MPI_Request *rq);
int MPI_Irecv(void *b, int n, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Request *rq);
int MPI_Wait(MPI_Request *rq, MPI_Status *st);
int MPI_Waitall(int n, MPI_Request *rq, MPI_Status *st);
int MPI_Sendrecv(const void *sb, int sn, MPI_Datatype st, int dst, int stag, void *rb, int rn, MPI_Datatype rt, int src, int rtag, MPI_Comm c, MPI_Status *status);
int MPI_Type_contiguous(int n, MPI_Datatype old, MPI_Datatype *newt);
int MPI_Type_commit(MPI_Datatype *t);
int MPI_Type_free(MPI_Datatype *t);
int MPI_Type_size(MPI_Datatype t, int *sz);
int MPI_Comm_group(MPI_Comm c, MPI_Group *g);
int MPI_Group_incl(MPI_Group g, int n, const int *ranks, MPI_Group *ng);
int MPI_Comm_create(MPI_Comm c, MPI_Group g, MPI_Comm *nc);
int MPI_File_open(MPI_Comm c, const char *fn, int mode, MPI_Info info, MPI_File *fh);
int MPI_File_close(MPI_File *fh);
int MPI_File_set_size(MPI_File fh, MPI_Offset sz);
int MPI_File_write(MPI_File fh, const void *b, int n, MPI_Datatype t, MPI_Status *st);
int MPI_File_write_at_all(MPI_File fh, MPI_Offset off, const void *b, int n, MPI_Datatype t, MPI_Status *st);
int MPI_File_write_at(MPI_File fh, MPI_Offset off, const void *b, int n, MPI_Datatype t, MPI_Status *st);
int MPI_File_seek(MPI_File fh, MPI_Offset off, int whence);
int MPI_File_get_position(MPI_File fh, MPI_Offset *off);
/* mini-MPI control (not part of MPI): fork nprocs-1 children that continue from the call site.
 * window_bytes = per-rank shared staging window; must hold the largest single send buffer. */
int mini_mpi_spawn(int nprocs, size_t window_bytes);
/* children exit(0) here; rank 0 waits for them and returns their worst exit status */
int mini_mpi_join(void);
#endif

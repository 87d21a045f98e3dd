/* oracle/shmpi/mpi.h -- TEST INFRASTRUCTURE ONLY.
 *
 * A single-node, shared-memory implementation of the subset of MPI-1/2 that IMD's MPI build uses
 * (SURVEY.md section 8f rank 2), so that the UNMODIFIED reference `imd_mpi_nve_eam_nbl` can be compiled
 * here (no MPI in this image) and timed on the GPU box's host cores as "IMD's own MPI CPU build".
 *
 * Process model: MPI_Init() forks SHMPI_NP-1 children (environment variable, default 1); ranks talk through
 * one anonymous MAP_SHARED region: a sense-reversing barrier, per-rank slots for the collectives and one
 * single-producer/single-consumer byte ring per ordered rank pair for point-to-point traffic.
 * Calls covered are exactly those found in /root/reference/src for the 3-D MPI build:
 *   Init Finalize Abort Comm_size Comm_rank Barrier Wtime Bcast Reduce Allreduce Send Recv Sendrecv Isend Irecv
 *   Wait Waitall Waitany Get_count Cart_create Cart_coords Cart_rank Alloc_mem Free_mem
 *   (src/imd_mpi_util.c:43-82, 263-339; src/imd_geom_mpi_3d.c:43-53; src/imd_comm_force_3d.c:268-395;
 *    src/imd_fix_cells_3d.c:209-437; src/imd_io.c:136-147; src/imd_io_3d.c:678-680).
 * Nothing under imd_b200/ includes or links this.
 */
#ifndef SHMPI_MPI_H
#define SHMPI_MPI_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MPI_VERSION    2
#define MPI_SUBVERSION 0

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Info;
typedef long MPI_Aint;
typedef struct shmpi_request *MPI_Request;

typedef struct {
  int MPI_SOURCE;
  int MPI_TAG;
  int MPI_ERROR;
  long shmpi_bytes;
} MPI_Status;

#define MPI_SUCCESS      0
#define MPI_COMM_WORLD   ((MPI_Comm)1)
#define MPI_COMM_NULL    ((MPI_Comm)0)
#define MPI_INFO_NULL    ((MPI_Info)0)
#define MPI_REQUEST_NULL ((MPI_Request)0)
#define MPI_STATUS_IGNORE   ((MPI_Status *)0)
#define MPI_STATUSES_IGNORE ((MPI_Status *)0)
#define MPI_ANY_SOURCE   (-1)
#define MPI_ANY_TAG      (-1)
#define MPI_PROC_NULL    (-2)
#define MPI_UNDEFINED    (-32766)

/* datatypes: the value is the element size in the low byte, a kind code above it */
#define SHMPI_DT(kind, size) (((kind) << 8) | (size))
#define MPI_CHAR           SHMPI_DT(1, 1)
#define MPI_BYTE           SHMPI_DT(2, 1)
#define MPI_SHORT          SHMPI_DT(3, 2)
#define MPI_INT            SHMPI_DT(4, 4)
#define MPI_LONG           SHMPI_DT(5, 8)
#define MPI_FLOAT          SHMPI_DT(6, 4)
#define MPI_DOUBLE         SHMPI_DT(7, 8)
#define MPI_UNSIGNED       SHMPI_DT(8, 4)
#define MPI_UNSIGNED_CHAR  SHMPI_DT(9, 1)
#define MPI_UNSIGNED_LONG  SHMPI_DT(10, 8)
#define MPI_LONG_LONG      SHMPI_DT(11, 8)

#define MPI_SUM  1
#define MPI_MAX  2
#define MPI_MIN  3
#define MPI_PROD 4

int MPI_Init(int *argc, char ***argv);
int MPI_Finalize(void);
int MPI_Abort(MPI_Comm comm, int code);
int MPI_Comm_size(MPI_Comm comm, int *size);
int MPI_Comm_rank(MPI_Comm comm, int *rank);
int MPI_Barrier(MPI_Comm comm);
double MPI_Wtime(void);

int MPI_Bcast(void *buf, int count, MPI_Datatype dt, int root, MPI_Comm comm);
int MPI_Reduce(const void *sbuf, void *rbuf, int count, MPI_Datatype dt, MPI_Op op, int root, MPI_Comm comm);
int MPI_Allreduce(const void *sbuf, void *rbuf, int count, MPI_Datatype dt, MPI_Op op, MPI_Comm comm);

int MPI_Send(const void *buf, int count, MPI_Datatype dt, int dest, int tag, MPI_Comm comm);
int MPI_Recv(void *buf, int count, MPI_Datatype dt, int src, int tag, MPI_Comm comm, MPI_Status *st);
int MPI_Sendrecv(const void *sbuf, int scount, MPI_Datatype sdt, int dest, int stag,
                 void *rbuf, int rcount, MPI_Datatype rdt, int src, int rtag, MPI_Comm comm, MPI_Status *st);
int MPI_Isend(const void *buf, int count, MPI_Datatype dt, int dest, int tag, MPI_Comm comm, MPI_Request *req);
int MPI_Irecv(void *buf, int count, MPI_Datatype dt, int src, int tag, MPI_Comm comm, MPI_Request *req);
int MPI_Wait(MPI_Request *req, MPI_Status *st);
int MPI_Waitall(int n, MPI_Request *reqs, MPI_Status *sts);
int MPI_Waitany(int n, MPI_Request *reqs, int *index, MPI_Status *st);
int MPI_Get_count(const MPI_Status *st, MPI_Datatype dt, int *count);

int MPI_Cart_create(MPI_Comm comm, int ndims, const int *dims, const int *periods, int reorder, MPI_Comm *cart);
int MPI_Cart_coords(MPI_Comm comm, int rank, int maxdims, int *coords);
int MPI_Cart_rank(MPI_Comm comm, const int *coords, int *rank);

int MPI_Alloc_mem(MPI_Aint size, MPI_Info info, void *baseptr);
int MPI_Free_mem(void *base);

#ifdef __cplusplus
}
#endif
#endif

/* TEST INFRASTRUCTURE ONLY (oracle).  In-process MPI stand-in used to compile the
 * reference's own FSILS / solver sources (under /root/reference, never copied) into
 * oracle/_ref/.  One "rank" = one host thread of the same process; the functions are
 * implemented in mpi_stub.cpp.  With a single thread it degenerates to a serial MPI.
 * Only the subset of MPI the reference hot path uses is provided
 * (Code/Source/liner_solver/{lhs,in_commu,dot,norm,bcast,bc,precond,ns_solver}.cpp,
 *  Code/Source/solver/{CmMod,all_fun}.cpp).
 */
#ifndef ORACLE_MPI_STUB_H
#define ORACLE_MPI_STUB_H

#ifdef __cplusplus
extern "C" {
#endif

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef struct { int idx; } MPI_Request;
typedef struct { int MPI_SOURCE, MPI_TAG, MPI_ERROR; } MPI_Status;

#define MPI_COMM_WORLD 0
#define MPI_SUCCESS 0
#define MPI_STATUS_SIZE 3
#define MPI_STATUS_IGNORE ((MPI_Status*)0)

#define MPI_INTEGER 1
#define MPI_INT 1
#define MPI_DOUBLE_PRECISION 2
#define MPI_DOUBLE 2
#define MPI_LOGICAL 3
#define MPI_CXX_BOOL 4
#define MPI_CHARACTER 5
#define MPI_CHAR 5

#define MPI_SUM 1
#define MPI_MAX 2
#define MPI_MIN 3

int MPI_Init(int*, char***);
int MPI_Finalize(void);
int MPI_Comm_rank(MPI_Comm, int*);
int MPI_Comm_size(MPI_Comm, int*);
int MPI_Barrier(MPI_Comm);
int MPI_Allreduce(const void* s, void* r, int n, MPI_Datatype t, MPI_Op op, MPI_Comm c);
int MPI_Allgather(const void* s, int sn, MPI_Datatype st, void* r, int rn, MPI_Datatype rt, MPI_Comm c);
int MPI_Allgatherv(const void* s, int sn, MPI_Datatype st, void* r, const int* rc, const int* disp, MPI_Datatype rt, MPI_Comm c);
int MPI_Gatherv(const void* s, int sn, MPI_Datatype st, void* r, const int* rc, const int* disp, MPI_Datatype rt, int root, MPI_Comm c);
int MPI_Bcast(void* b, int n, MPI_Datatype t, int root, MPI_Comm c);
int MPI_Send(const void* b, int n, MPI_Datatype t, int dst, int tag, MPI_Comm c);
int MPI_Recv(void* b, int n, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Status* st);
int MPI_Isend(const void* b, int n, MPI_Datatype t, int dst, int tag, MPI_Comm c, MPI_Request* rq);
int MPI_Irecv(void* b, int n, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Request* rq);
int MPI_Wait(MPI_Request* rq, MPI_Status* st);
int MPI_Reduce(const void* s, void* r, int n, MPI_Datatype t, MPI_Op op, int root, MPI_Comm c);
int MPI_Gather(const void* s, int sn, MPI_Datatype st, void* r, int rn, MPI_Datatype rt, int root, MPI_Comm c);
int MPI_Scatterv(const void* s, const int* sc, const int* disp, MPI_Datatype st, void* r, int rn, MPI_Datatype rt, int root, MPI_Comm c);

/* MPI-IO: the reference only uses it for its partition cache file (distribute.cpp:1649-1653, 1724-1728: partitioning.bin).  The
 * stand-in never finds such a file and never writes one, so every run partitions afresh. */
typedef struct { int unused; } MPI_File;
typedef long long MPI_Offset;
typedef int MPI_Info;
#define MPI_INFO_NULL 0
#define MPI_MODE_RDONLY 2
#define MPI_MODE_WRONLY 8
#define MPI_MODE_CREATE 1
int MPI_File_open(MPI_Comm c, const char* name, int mode, MPI_Info info, MPI_File* f);
int MPI_File_set_view(MPI_File f, MPI_Offset disp, MPI_Datatype et, MPI_Datatype ft, const char* rep, MPI_Info info);
int MPI_File_read(MPI_File f, void* buf, int n, MPI_Datatype t, MPI_Status* st);
int MPI_File_write(MPI_File f, const void* buf, int n, MPI_Datatype t, MPI_Status* st);
int MPI_File_close(MPI_File* f);

/* stub control (called by the harness, not by the reference) */
void mpistub_set_world(int size);      /* before spawning rank threads */
void mpistub_bind_rank(int rank);      /* first call in every rank thread */

#ifdef __cplusplus
}
#endif
#endif

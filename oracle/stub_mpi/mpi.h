/* Single-process stand-in for <mpi.h>.
 *
 * TEST INFRASTRUCTURE ONLY (oracle build). The SPADE reference includes "mpi.h"
 * unconditionally (reference: src/parallel/compute_pool.h:10) but only needs it for
 * rank bookkeeping when every "rank" is a std::thread of one process. This header gives the
 * ~40 symbols the reference names single-node semantics: one node, rank 0, collectives copy
 * their input to their output.
 */
#ifndef SPB_ORACLE_STUB_MPI_H
#define SPB_ORACLE_STUB_MPI_H
#include <stddef.h>
#include <string.h>
#include <stdio.h>

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Request;
typedef int MPI_Info;
typedef long long MPI_Offset;
typedef struct { int MPI_SOURCE, MPI_TAG, MPI_ERROR; } MPI_Status;
typedef FILE* MPI_File;

#define MPI_COMM_WORLD 0
#define MPI_SUCCESS 0
#define MPI_CHAR 1
#define MPI_INT 4
#define MPI_FLOAT 5
#define MPI_DOUBLE 8
#define MPI_UINT64_T 9
#define MPI_SUM 1
#define MPI_MIN 2
#define MPI_MAX 3
#define MPI_THREAD_FUNNELED 1
#define MPI_MODE_RDWR 2
#define MPI_MODE_CREATE 1
#define MPI_INFO_NULL 0
#define MPI_STATUS_IGNORE ((MPI_Status*)0)
#define MPI_STATUSES_IGNORE ((MPI_Status*)0)
#define MPI_MAX_ERROR_STRING 64

static inline size_t spb_stub_mpi_size(MPI_Datatype t)
{
    switch (t) { case MPI_CHAR: return 1; case MPI_INT: return 4; case MPI_FLOAT: return 4; default: return 8; }
}
static inline int MPI_Init(int*, char***) { return MPI_SUCCESS; }
static inline int MPI_Init_thread(int*, char***, int req, int* prov) { if (prov) *prov = req; return MPI_SUCCESS; }
static inline int MPI_Finalize(void) { return MPI_SUCCESS; }
static inline int MPI_Comm_rank(MPI_Comm, int* r) { *r = 0; return MPI_SUCCESS; }
static inline int MPI_Comm_size(MPI_Comm, int* s) { *s = 1; return MPI_SUCCESS; }
static inline int MPI_Barrier(MPI_Comm) { return MPI_SUCCESS; }
static inline int MPI_Allreduce(const void* in, void* out, int n, MPI_Datatype t, MPI_Op, MPI_Comm)
{ if (in != out) memcpy(out, in, n*spb_stub_mpi_size(t)); return MPI_SUCCESS; }
static inline int MPI_Allgather(const void* in, int n, MPI_Datatype t, void* out, int, MPI_Datatype, MPI_Comm)
{ if (in != out) memcpy(out, in, n*spb_stub_mpi_size(t)); return MPI_SUCCESS; }
static inline int MPI_Gather(const void* in, int n, MPI_Datatype t, void* out, int, MPI_Datatype, int, MPI_Comm)
{ if (in != out) memcpy(out, in, n*spb_stub_mpi_size(t)); return MPI_SUCCESS; }
static inline int MPI_Gatherv(const void* in, int n, MPI_Datatype t, void* out, const int*, const int* displs, MPI_Datatype, int, MPI_Comm)
{ memcpy((char*)out + (displs ? displs[0] : 0)*spb_stub_mpi_size(t), in, n*spb_stub_mpi_size(t)); return MPI_SUCCESS; }
static inline int MPI_Irecv(void*, int, MPI_Datatype, int, int, MPI_Comm, MPI_Request* r) { if (r) *r = 0; return MPI_SUCCESS; }
static inline int MPI_Ssend(const void*, int, MPI_Datatype, int, int, MPI_Comm) { return MPI_SUCCESS; }
static inline int MPI_Waitall(int, MPI_Request*, MPI_Status*) { return MPI_SUCCESS; }
static inline int MPI_Error_string(int, char* s, int* len) { strcpy(s, "stub-mpi"); if (len) *len = 8; return MPI_SUCCESS; }
static inline int MPI_File_open(MPI_Comm, const char* fn, int, MPI_Info, MPI_File* fh)
{ *fh = fopen(fn, "r+b"); if (!*fh) *fh = fopen(fn, "w+b"); return *fh ? MPI_SUCCESS : 1; }
static inline int MPI_File_close(MPI_File* fh) { if (*fh) fclose(*fh); *fh = 0; return MPI_SUCCESS; }
static inline int MPI_File_write_at(MPI_File fh, MPI_Offset off, const void* buf, int n, MPI_Datatype t, MPI_Status*)
{ fseek(fh, (long)off, SEEK_SET); fwrite(buf, spb_stub_mpi_size(t), n, fh); return MPI_SUCCESS; }
static inline int MPI_File_read_at(MPI_File fh, MPI_Offset off, void* buf, int n, MPI_Datatype t, MPI_Status*)
{ fseek(fh, (long)off, SEEK_SET); size_t got = fread(buf, spb_stub_mpi_size(t), n, fh); (void)got; return MPI_SUCCESS; }
#endif

/*
 * TEST INFRASTRUCTURE ONLY -- serial, single-rank MPI stand-in.
 *
 * Lets the reference's hot-path sources (src/pb/*.cc, src/linear_algebra/
 * mputils.cc, src/Preconditioning.cc, ...) compile UNMODIFIED in a container
 * that has no MPI installation.  Every communicator has
 * exactly one rank: collectives copy sendbuf -> recvbuf, the Cartesian
 * topology is 1x1x1 periodic (every neighbour is rank 0 itself), and
 * point-to-point calls are never reached by the reference when a direction
 * has a single task (it takes its local periodic-wrap branch instead,
 * src/pb/GridFuncVector.cc:586-603, src/pb/GridFunc.cc:1953-1967).
 *
 * This header is ours (not copied from any MPI distribution); it implements
 * only the ~45 entry points those sources name.
 */
#ifndef MGMOL_B200_ORACLE_MPI_STUB_H
#define MGMOL_B200_ORACLE_MPI_STUB_H

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Request;
typedef int MPI_Group;
typedef int MPI_Info;
typedef long MPI_Aint;
typedef struct MPI_Status
{
    int MPI_SOURCE;
    int MPI_TAG;
    int MPI_ERROR;
} MPI_Status;

#define MPI_SUCCESS 0
#define MPI_COMM_NULL ((MPI_Comm)0)
#define MPI_COMM_WORLD ((MPI_Comm)1)
#define MPI_COMM_SELF ((MPI_Comm)2)
#define MPI_REQUEST_NULL ((MPI_Request)0)
#define MPI_STATUS_IGNORE ((MPI_Status*)0)
#define MPI_STATUSES_IGNORE ((MPI_Status*)0)
#define MPI_MAX_PROCESSOR_NAME 256
#define MPI_MAX_ERROR_STRING 256
#define MPI_ANY_SOURCE (-1)
#define MPI_ANY_TAG (-1)
#define MPI_PROC_NULL (-2)
#define MPI_UNDEFINED (-32766)
#define MPI_IN_PLACE ((void*)1)
#define MPI_INFO_NULL ((MPI_Info)0)

/* datatypes: value = size in bytes, tagged in the high bits to stay distinct */
#define MPI_CHAR ((MPI_Datatype)0x0101)
#define MPI_BYTE ((MPI_Datatype)0x0201)
#define MPI_SHORT ((MPI_Datatype)0x0302)
#define MPI_UNSIGNED_SHORT ((MPI_Datatype)0x0402)
#define MPI_INT ((MPI_Datatype)0x0504)
#define MPI_UNSIGNED ((MPI_Datatype)0x0604)
#define MPI_LONG ((MPI_Datatype)0x0708)
#define MPI_UNSIGNED_LONG ((MPI_Datatype)0x0808)
#define MPI_LONG_LONG ((MPI_Datatype)0x0908)
#define MPI_FLOAT ((MPI_Datatype)0x0a04)
#define MPI_DOUBLE ((MPI_Datatype)0x0b08)
#define MPI_C_BOOL ((MPI_Datatype)0x0c01)
#define MPI_UNSIGNED_CHAR ((MPI_Datatype)0x0d01)
#define MPI_LONG_DOUBLE ((MPI_Datatype)0x0e10)
#define MPI_DOUBLE_INT ((MPI_Datatype)0x0f10)
#define MPI_FLOAT_INT ((MPI_Datatype)0x1008)

#define MPI_SUM 1
#define MPI_MAX 2
#define MPI_MIN 3
#define MPI_PROD 4
#define MPI_LOR 5
#define MPI_LAND 6
#define MPI_MAXLOC 7
#define MPI_MINLOC 8

static inline size_t mpi_stub_sizeof(MPI_Datatype t) { return (size_t)(t & 0xff); }

static inline void mpi_stub_copy(
    const void* s, void* r, int count, MPI_Datatype t)
{
    if (s != MPI_IN_PLACE && s != r && count > 0)
        memcpy(r, s, (size_t)count * mpi_stub_sizeof(t));
}

static inline int MPI_Init(int* argc, char*** argv)
{
    (void)argc;
    (void)argv;
    return MPI_SUCCESS;
}
static inline int MPI_Initialized(int* flag)
{
    *flag = 1;
    return MPI_SUCCESS;
}
static inline int MPI_Finalize(void) { return MPI_SUCCESS; }
static inline int MPI_Abort(MPI_Comm c, int code)
{
    (void)c;
    fprintf(stderr, "[mpi stub] MPI_Abort(%d)\n", code);
    abort();
    return MPI_SUCCESS;
}
static inline double MPI_Wtime(void)
{
    struct timeval tv;
    gettimeofday(&tv, 0);
    return (double)tv.tv_sec + 1.e-6 * (double)tv.tv_usec;
}
static inline int MPI_Comm_rank(MPI_Comm c, int* rank)
{
    (void)c;
    *rank = 0;
    return MPI_SUCCESS;
}
static inline int MPI_Comm_size(MPI_Comm c, int* size)
{
    (void)c;
    *size = 1;
    return MPI_SUCCESS;
}
static inline int MPI_Comm_free(MPI_Comm* c)
{
    *c = MPI_COMM_NULL;
    return MPI_SUCCESS;
}
static inline int MPI_Comm_dup(MPI_Comm c, MPI_Comm* n)
{
    *n = c + 16;
    return MPI_SUCCESS;
}
static inline int MPI_Comm_split(MPI_Comm c, int color, int key, MPI_Comm* n)
{
    (void)color;
    (void)key;
    *n = c + 16;
    return MPI_SUCCESS;
}
static inline int MPI_Comm_compare(MPI_Comm a, MPI_Comm b, int* result)
{
    *result = (a == b) ? 0 : 3;
    return MPI_SUCCESS;
}
static inline int MPI_Barrier(MPI_Comm c)
{
    (void)c;
    return MPI_SUCCESS;
}
static inline int MPI_Bcast(
    void* buf, int count, MPI_Datatype t, int root, MPI_Comm c)
{
    (void)buf;
    (void)count;
    (void)t;
    (void)root;
    (void)c;
    return MPI_SUCCESS;
}
static inline int MPI_Allreduce(const void* s, void* r, int count,
    MPI_Datatype t, MPI_Op op, MPI_Comm c)
{
    (void)op;
    (void)c;
    mpi_stub_copy(s, r, count, t);
    return MPI_SUCCESS;
}
static inline int MPI_Reduce(const void* s, void* r, int count, MPI_Datatype t,
    MPI_Op op, int root, MPI_Comm c)
{
    (void)op;
    (void)root;
    (void)c;
    mpi_stub_copy(s, r, count, t);
    return MPI_SUCCESS;
}
static inline int MPI_Allgather(const void* s, int scount, MPI_Datatype st,
    void* r, int rcount, MPI_Datatype rt, MPI_Comm c)
{
    (void)rcount;
    (void)rt;
    (void)c;
    mpi_stub_copy(s, r, scount, st);
    return MPI_SUCCESS;
}
static inline int MPI_Allgatherv(const void* s, int scount, MPI_Datatype st,
    void* r, const int* rcounts, const int* displs, MPI_Datatype rt, MPI_Comm c)
{
    (void)rcounts;
    (void)c;
    if (s != MPI_IN_PLACE)
        memcpy((char*)r + (size_t)displs[0] * mpi_stub_sizeof(rt), s,
            (size_t)scount * mpi_stub_sizeof(st));
    return MPI_SUCCESS;
}
static inline int MPI_Gather(const void* s, int scount, MPI_Datatype st,
    void* r, int rcount, MPI_Datatype rt, int root, MPI_Comm c)
{
    (void)rcount;
    (void)rt;
    (void)root;
    (void)c;
    mpi_stub_copy(s, r, scount, st);
    return MPI_SUCCESS;
}
static inline int MPI_Gatherv(const void* s, int scount, MPI_Datatype st,
    void* r, const int* rcounts, const int* displs, MPI_Datatype rt, int root,
    MPI_Comm c)
{
    (void)rcounts;
    (void)root;
    (void)c;
    if (s != MPI_IN_PLACE)
        memcpy((char*)r + (size_t)displs[0] * mpi_stub_sizeof(rt), s,
            (size_t)scount * mpi_stub_sizeof(st));
    return MPI_SUCCESS;
}
static inline int MPI_Scatter(const void* s, int scount, MPI_Datatype st,
    void* r, int rcount, MPI_Datatype rt, int root, MPI_Comm c)
{
    (void)scount;
    (void)st;
    (void)root;
    (void)c;
    mpi_stub_copy(s, r, rcount, rt);
    return MPI_SUCCESS;
}
static inline int MPI_Alltoall(const void* s, int scount, MPI_Datatype st,
    void* r, int rcount, MPI_Datatype rt, MPI_Comm c)
{
    (void)rcount;
    (void)rt;
    (void)c;
    mpi_stub_copy(s, r, scount, st);
    return MPI_SUCCESS;
}

/* Point-to-point: with one rank per direction the reference never posts
 * these (it wraps locally).  A send to self without a matching receive would
 * silently lose data, so fail loudly if one is ever reached. */
static inline int mpi_stub_p2p_unreachable(const char* what)
{
    fprintf(stderr,
        "[mpi stub] %s reached: the serial oracle supports one rank only\n",
        what);
    abort();
    return 1;
}
static inline int MPI_Send(const void* b, int n, MPI_Datatype t, int dest,
    int tag, MPI_Comm c)
{
    (void)b;
    (void)n;
    (void)t;
    (void)dest;
    (void)tag;
    (void)c;
    return mpi_stub_p2p_unreachable("MPI_Send");
}
static inline int MPI_Recv(void* b, int n, MPI_Datatype t, int src, int tag,
    MPI_Comm c, MPI_Status* st)
{
    (void)b;
    (void)n;
    (void)t;
    (void)src;
    (void)tag;
    (void)c;
    (void)st;
    return mpi_stub_p2p_unreachable("MPI_Recv");
}
static inline int MPI_Isend(const void* b, int n, MPI_Datatype t, int dest,
    int tag, MPI_Comm c, MPI_Request* rq)
{
    (void)b;
    (void)n;
    (void)t;
    (void)dest;
    (void)tag;
    (void)c;
    (void)rq;
    return mpi_stub_p2p_unreachable("MPI_Isend");
}
static inline int MPI_Irecv(void* b, int n, MPI_Datatype t, int src, int tag,
    MPI_Comm c, MPI_Request* rq)
{
    (void)b;
    (void)n;
    (void)t;
    (void)src;
    (void)tag;
    (void)c;
    (void)rq;
    return mpi_stub_p2p_unreachable("MPI_Irecv");
}
static inline int MPI_Wait(MPI_Request* rq, MPI_Status* st)
{
    (void)st;
    if (rq) *rq = MPI_REQUEST_NULL;
    return MPI_SUCCESS;
}
static inline int MPI_Waitall(int n, MPI_Request* rq, MPI_Status* st)
{
    (void)st;
    for (int i = 0; i < n; i++)
        rq[i] = MPI_REQUEST_NULL;
    return MPI_SUCCESS;
}
static inline int MPI_Get_count(const MPI_Status* st, MPI_Datatype t, int* n)
{
    (void)st;
    (void)t;
    *n = 0;
    return MPI_SUCCESS;
}

/* 1x1x1 periodic Cartesian topology */
static inline int MPI_Cart_create(MPI_Comm c, int ndims, const int* dims,
    const int* periods, int reorder, MPI_Comm* cart)
{
    (void)periods;
    (void)reorder;
    for (int i = 0; i < ndims; i++)
        if (dims[i] != 1)
        {
            fprintf(stderr, "[mpi stub] MPI_Cart_create with dims != 1\n");
            abort();
        }
    *cart = c + 32;
    return MPI_SUCCESS;
}
static inline int MPI_Cart_coords(
    MPI_Comm c, int rank, int maxdims, int* coords)
{
    (void)c;
    (void)rank;
    for (int i = 0; i < maxdims; i++)
        coords[i] = 0;
    return MPI_SUCCESS;
}
static inline int MPI_Cart_rank(MPI_Comm c, const int* coords, int* rank)
{
    (void)c;
    (void)coords;
    *rank = 0;
    return MPI_SUCCESS;
}
static inline int MPI_Cart_shift(
    MPI_Comm c, int dir, int disp, int* src, int* dest)
{
    (void)c;
    (void)dir;
    (void)disp;
    *src  = 0;
    *dest = 0;
    return MPI_SUCCESS;
}
static inline int MPI_Dims_create(int nnodes, int ndims, int* dims)
{
    (void)nnodes;
    for (int i = 0; i < ndims; i++)
        dims[i] = 1;
    return MPI_SUCCESS;
}
static inline int MPI_Get_processor_name(char* name, int* len)
{
    strcpy(name, "serial-oracle");
    *len = (int)strlen(name);
    return MPI_SUCCESS;
}
static inline int PMPI_Get_processor_name(char* name, int* len)
{
    return MPI_Get_processor_name(name, len);
}
static inline int MPI_Error_string(int code, char* s, int* len)
{
    (void)code;
    strcpy(s, "mpi stub error");
    *len = (int)strlen(s);
    return MPI_SUCCESS;
}

#ifdef __cplusplus
}
#endif

#endif

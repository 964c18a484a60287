/* Single-rank MPI shim for building the reference CPU path as the parity oracle.
 *
 * TEST INFRASTRUCTURE ONLY.  This image has no MPI; the reference (LLNL/ddcMD)
 * needs <mpi.h> to compile.  On one rank every collective degenerates to a copy
 * and point-to-point is never reached (ddcUpdateTables returns early when
 * getSize(0)==1, src/ddcUpdateAll.c:84-89; ddcUpdateForce is guarded by
 * getSize(0)>1, src/ddcenergy.c:216).  A basic datatype handle is its size in
 * bytes; derived datatypes get handles >= 1000 with their size in a table.
 */
#ifndef ORACLE_MPI_SHIM_H
#define ORACLE_MPI_SHIM_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Request;
typedef int MPI_Group;
typedef long MPI_Aint;
typedef struct { int MPI_SOURCE, MPI_TAG, MPI_ERROR, count; } MPI_Status;

#define MPI_COMM_WORLD 1
#define MPI_COMM_NULL 0
#define MPI_SUCCESS 0
#define MPI_ANY_SOURCE (-1)
#define MPI_ANY_TAG (-1)
#define MPI_UNDEFINED (-32766)
#define MPI_IDENT 0
#define MPI_TAG_UB 1
#define MPI_IN_PLACE ((void *)1)
#define MPI_STATUS_IGNORE ((MPI_Status *)0)
#define MPI_STATUSES_IGNORE ((MPI_Status *)0)

#define MPI_BYTE 1
#define MPI_CHAR 1
#define MPI_UNSIGNED_SHORT 2
#define MPI_INT 4
#define MPI_UNSIGNED 4
#define MPI_FLOAT 4
#define MPI_DOUBLE 8
#define MPI_LONG_LONG 8
#define MPI_LONG_LONG_INT 8
#define MPI_UNSIGNED_LONG_LONG 8
#define MPI_DOUBLE_INT 16

#define MPI_SUM 1
#define MPI_MAX 2
#define MPI_MIN 3
#define MPI_BAND 4
#define MPI_BOR 5
#define MPI_MAXLOC 6
#define MPI_MINLOC 7

int MPI_Init(int *, char ***);
int MPI_Finalize(void);
int MPI_Abort(MPI_Comm, int);
double MPI_Wtime(void);
int MPI_Comm_rank(MPI_Comm, int *);
int MPI_Comm_size(MPI_Comm, int *);
int MPI_Barrier(MPI_Comm);
int MPI_Comm_split(MPI_Comm, int, int, MPI_Comm *);
int MPI_Comm_compare(MPI_Comm, MPI_Comm, int *);
int MPI_Comm_create(MPI_Comm, MPI_Group, MPI_Comm *);
int MPI_Comm_group(MPI_Comm, MPI_Group *);
int MPI_Group_incl(MPI_Group, int, const int *, MPI_Group *);
int MPI_Group_free(MPI_Group *);
int MPI_Comm_get_attr(MPI_Comm, int, void *, int *);
int MPI_Attr_get(MPI_Comm, int, void *, int *);
int MPI_Allreduce(const void *, void *, int, MPI_Datatype, MPI_Op, MPI_Comm);
int MPI_Reduce(const void *, void *, int, MPI_Datatype, MPI_Op, int, MPI_Comm);
int MPI_Reduce_scatter(const void *, void *, const int *, MPI_Datatype, MPI_Op, MPI_Comm);
int MPI_Scan(const void *, void *, int, MPI_Datatype, MPI_Op, MPI_Comm);
int MPI_Exscan(const void *, void *, int, MPI_Datatype, MPI_Op, MPI_Comm);
int MPI_Bcast(void *, int, MPI_Datatype, int, MPI_Comm);
int MPI_Allgather(const void *, int, MPI_Datatype, void *, int, MPI_Datatype, MPI_Comm);
int MPI_Allgatherv(const void *, int, MPI_Datatype, void *, const int *, const int *, MPI_Datatype, MPI_Comm);
int MPI_Gather(const void *, int, MPI_Datatype, void *, int, MPI_Datatype, int, MPI_Comm);
int MPI_Gatherv(const void *, int, MPI_Datatype, void *, const int *, const int *, MPI_Datatype, int, MPI_Comm);
int MPI_Alltoall(const void *, int, MPI_Datatype, void *, int, MPI_Datatype, MPI_Comm);
int MPI_Alltoallv(const void *, const int *, const int *, MPI_Datatype, void *, const int *, const int *, MPI_Datatype, MPI_Comm);
int MPI_Send(const void *, int, MPI_Datatype, int, int, MPI_Comm);
int MPI_Recv(void *, int, MPI_Datatype, int, int, MPI_Comm, MPI_Status *);
int MPI_Isend(const void *, int, MPI_Datatype, int, int, MPI_Comm, MPI_Request *);
int MPI_Irecv(void *, int, MPI_Datatype, int, int, MPI_Comm, MPI_Request *);
int MPI_Wait(MPI_Request *, MPI_Status *);
int MPI_Waitall(int, MPI_Request *, MPI_Status *);
int MPI_Waitsome(int, MPI_Request *, int *, int *, MPI_Status *);
int MPI_Iprobe(int, int, MPI_Comm, int *, MPI_Status *);
int MPI_Get_count(const MPI_Status *, MPI_Datatype, int *);
int MPI_Get_address(const void *, MPI_Aint *);
int MPI_Address(void *, MPI_Aint *);
int MPI_Type_commit(MPI_Datatype *);
int MPI_Type_free(MPI_Datatype *);
int MPI_Type_size(MPI_Datatype, int *);
int MPI_Type_contiguous(int, MPI_Datatype, MPI_Datatype *);
int MPI_Type_create_hvector(int, int, MPI_Aint, MPI_Datatype, MPI_Datatype *);
int MPI_Type_create_struct(int, const int *, const MPI_Aint *, const MPI_Datatype *, MPI_Datatype *);
int MPI_Type_struct(int, int *, MPI_Aint *, MPI_Datatype *, MPI_Datatype *);

#ifdef __cplusplus
}
#endif
#endif

/* Single-rank MPI shim (see mpi.h).  TEST INFRASTRUCTURE ONLY. */
#include "mpi.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>

#define MAX_DERIVED 8192
static int derived_size[MAX_DERIVED];
static int n_derived = 0;

static int type_bytes(MPI_Datatype t) { return t < 1000 ? t : derived_size[(t - 1000) % MAX_DERIVED]; }
static int new_type(int bytes)
{
    int id = n_derived % MAX_DERIVED;
    derived_size[id] = bytes;
    n_derived++;
    return 1000 + id;
}
static void copy_if_distinct(const void *s, void *r, int n, MPI_Datatype t)
{
    if (s != MPI_IN_PLACE && s != r) memcpy(r, s, (size_t)n * type_bytes(t));
}
static void p2p_unreachable(const char *what)
{
    fprintf(stderr, "mpi_shim: %s called on a single rank\n", what);
    abort();
}

int MPI_Init(int *a, char ***b) { (void)a; (void)b; return 0; }
int MPI_Finalize(void) { return 0; }
int MPI_Abort(MPI_Comm c, int e) { (void)c; exit(e ? e : 1); }
double MPI_Wtime(void)
{
    struct timeval t;
    gettimeofday(&t, 0);
    return t.tv_sec + 1e-6 * t.tv_usec;
}
int MPI_Comm_rank(MPI_Comm c, int *r) { (void)c; *r = 0; return 0; }
int MPI_Comm_size(MPI_Comm c, int *s) { (void)c; *s = 1; return 0; }
int MPI_Barrier(MPI_Comm c) { (void)c; return 0; }
int MPI_Comm_split(MPI_Comm c, int color, int key, MPI_Comm *n)
{
    (void)key;
    *n = (color == MPI_UNDEFINED) ? MPI_COMM_NULL : c + 1;
    return 0;
}
int MPI_Comm_compare(MPI_Comm a, MPI_Comm b, int *r) { (void)a; (void)b; *r = MPI_IDENT; return 0; }
int MPI_Comm_create(MPI_Comm c, MPI_Group g, MPI_Comm *n) { (void)g; *n = c + 1; return 0; }
int MPI_Comm_group(MPI_Comm c, MPI_Group *g) { (void)c; *g = 1; return 0; }
int MPI_Group_incl(MPI_Group g, int n, const int *r, MPI_Group *o) { (void)g; (void)n; (void)r; *o = 1; return 0; }
int MPI_Group_free(MPI_Group *g) { (void)g; return 0; }
static int tag_ub = 32767;
int MPI_Comm_get_attr(MPI_Comm c, int k, void *v, int *f) { (void)c; (void)k; *(int **)v = &tag_ub; *f = 1; return 0; }
int MPI_Attr_get(MPI_Comm c, int k, void *v, int *f) { return MPI_Comm_get_attr(c, k, v, f); }

int MPI_Allreduce(const void *s, void *r, int n, MPI_Datatype t, MPI_Op o, MPI_Comm c)
{ (void)o; (void)c; copy_if_distinct(s, r, n, t); return 0; }
int MPI_Reduce(const void *s, void *r, int n, MPI_Datatype t, MPI_Op o, int root, MPI_Comm c)
{ (void)o; (void)root; (void)c; copy_if_distinct(s, r, n, t); return 0; }
int MPI_Reduce_scatter(const void *s, void *r, const int *cn, MPI_Datatype t, MPI_Op o, MPI_Comm c)
{ (void)o; (void)c; copy_if_distinct(s, r, cn[0], t); return 0; }
int MPI_Scan(const void *s, void *r, int n, MPI_Datatype t, MPI_Op o, MPI_Comm c)
{ (void)o; (void)c; copy_if_distinct(s, r, n, t); return 0; }
int MPI_Exscan(const void *s, void *r, int n, MPI_Datatype t, MPI_Op o, MPI_Comm c)
{ (void)s; (void)r; (void)n; (void)t; (void)o; (void)c; return 0; }
int MPI_Bcast(void *b, int n, MPI_Datatype t, int root, MPI_Comm c)
{ (void)b; (void)n; (void)t; (void)root; (void)c; return 0; }
int MPI_Allgather(const void *s, int n, MPI_Datatype t, void *r, int rn, MPI_Datatype rt, MPI_Comm c)
{ (void)rn; (void)rt; (void)c; copy_if_distinct(s, r, n, t); return 0; }
int MPI_Allgatherv(const void *s, int n, MPI_Datatype t, void *r, const int *rc, const int *d, MPI_Datatype rt, MPI_Comm c)
{
    (void)rc; (void)c;
    if (s != MPI_IN_PLACE) memcpy((char *)r + (size_t)d[0] * type_bytes(rt), s, (size_t)n * type_bytes(t));
    return 0;
}
int MPI_Gather(const void *s, int n, MPI_Datatype t, void *r, int rn, MPI_Datatype rt, int root, MPI_Comm c)
{ (void)rn; (void)rt; (void)root; (void)c; copy_if_distinct(s, r, n, t); return 0; }
int MPI_Gatherv(const void *s, int n, MPI_Datatype t, void *r, const int *rc, const int *d, MPI_Datatype rt, int root, MPI_Comm c)
{
    (void)rc; (void)root; (void)c;
    if (s != MPI_IN_PLACE) memcpy((char *)r + (size_t)d[0] * type_bytes(rt), s, (size_t)n * type_bytes(t));
    return 0;
}
int MPI_Alltoall(const void *s, int n, MPI_Datatype t, void *r, int rn, MPI_Datatype rt, MPI_Comm c)
{ (void)rn; (void)rt; (void)c; copy_if_distinct(s, r, n, t); return 0; }
int MPI_Alltoallv(const void *s, const int *sc, const int *sd, MPI_Datatype t, void *r, const int *rc, const int *rd, MPI_Datatype rt, MPI_Comm c)
{
    (void)rc; (void)c;
    memcpy((char *)r + (size_t)rd[0] * type_bytes(rt), (const char *)s + (size_t)sd[0] * type_bytes(t), (size_t)sc[0] * type_bytes(t));
    return 0;
}

int MPI_Send(const void *b, int n, MPI_Datatype t, int d, int tag, MPI_Comm c)
{ (void)b; (void)n; (void)t; (void)d; (void)tag; (void)c; p2p_unreachable("MPI_Send"); return 0; }
int MPI_Recv(void *b, int n, MPI_Datatype t, int s, int tag, MPI_Comm c, MPI_Status *st)
{ (void)b; (void)n; (void)t; (void)s; (void)tag; (void)c; (void)st; p2p_unreachable("MPI_Recv"); return 0; }
int MPI_Isend(const void *b, int n, MPI_Datatype t, int d, int tag, MPI_Comm c, MPI_Request *r)
{ (void)b; (void)n; (void)t; (void)d; (void)tag; (void)c; (void)r; p2p_unreachable("MPI_Isend"); return 0; }
int MPI_Irecv(void *b, int n, MPI_Datatype t, int s, int tag, MPI_Comm c, MPI_Request *r)
{ (void)b; (void)n; (void)t; (void)s; (void)tag; (void)c; (void)r; p2p_unreachable("MPI_Irecv"); return 0; }
int MPI_Wait(MPI_Request *r, MPI_Status *s) { (void)r; (void)s; return 0; }
int MPI_Waitall(int n, MPI_Request *r, MPI_Status *s) { (void)n; (void)r; (void)s; return 0; }
int MPI_Waitsome(int n, MPI_Request *r, int *o, int *i, MPI_Status *s)
{ (void)n; (void)r; (void)i; (void)s; *o = MPI_UNDEFINED; return 0; }
int MPI_Iprobe(int s, int t, MPI_Comm c, int *f, MPI_Status *st) { (void)s; (void)t; (void)c; (void)st; *f = 0; return 0; }
int MPI_Get_count(const MPI_Status *s, MPI_Datatype t, int *n) { (void)s; (void)t; *n = 0; return 0; }
int MPI_Get_address(const void *p, MPI_Aint *a) { *a = (MPI_Aint)p; return 0; }
int MPI_Address(void *p, MPI_Aint *a) { *a = (MPI_Aint)p; return 0; }
int MPI_Type_commit(MPI_Datatype *t) { (void)t; return 0; }
int MPI_Type_free(MPI_Datatype *t) { (void)t; return 0; }
int MPI_Type_size(MPI_Datatype t, int *s) { *s = type_bytes(t); return 0; }
int MPI_Type_contiguous(int n, MPI_Datatype t, MPI_Datatype *o) { *o = new_type(n * type_bytes(t)); return 0; }
int MPI_Type_create_hvector(int n, int bl, MPI_Aint st, MPI_Datatype t, MPI_Datatype *o)
{ (void)st; *o = new_type(n * bl * type_bytes(t)); return 0; }
int MPI_Type_create_struct(int n, const int *bl, const MPI_Aint *d, const MPI_Datatype *ty, MPI_Datatype *o)
{
    (void)d;
    int s = 0;
    for (int i = 0; i < n; i++) s += bl[i] * type_bytes(ty[i]);
    *o = new_type(s);
    return 0;
}
int MPI_Type_struct(int n, int *bl, MPI_Aint *d, MPI_Datatype *ty, MPI_Datatype *o)
{ return MPI_Type_create_struct(n, bl, d, ty, o); }

// nccl.h (CPU EMULATION SHIM) - TEST INFRASTRUCTURE ONLY.
// The handful of NCCL calls the product makes, emulated between PROCESSES of one host through a POSIX shared-memory
// segment: all-reduce (sum of doubles, summed in rank order) and grouped send/recv.  Every call is collective over
// the communicator and synchronous, which is how the product uses them (one group per exchange, every rank in it).
#pragma once
#include <fcntl.h>
#include <sched.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>
#include <vector>

typedef struct emuNcclComm
{
    int rank, nranks;
    char *base;
    size_t slot;       // bytes per rank slot
    char name[64];
    unsigned sense;
} *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclSuccess = 0, ncclSystemError = 2, ncclInvalidUsage = 5 };
enum ncclDataType_t { ncclInt8 = 0, ncclInt32 = 2, ncclDouble = 8 };
static inline size_t ncclTypeSize(ncclDataType_t t) { return t == ncclDouble ? 8 : (t == ncclInt32 ? 4 : 1); }
enum ncclRedOp_t { ncclSum = 0, ncclMax = 2 };

namespace emu
{
struct NcclHdr { volatile unsigned count; volatile unsigned sense; char pad[56]; };
struct NcclOp { int send; const void *src; void *dst; size_t bytes; int peer; };
inline std::vector<NcclOp> &ncclQueue() { static std::vector<NcclOp> q; return q; }
inline int &ncclGroupDepth() { static int d = 0; return d; }
inline ncclComm_t &ncclGroupComm() { static ncclComm_t c = nullptr; return c; }
inline void ncclBarrier(ncclComm_t c)
{
    NcclHdr *h = (NcclHdr *)c->base;
    c->sense ^= 1u;
    if (__atomic_add_fetch(&h->count, 1u, __ATOMIC_ACQ_REL) == (unsigned)c->nranks)
    {
        __atomic_store_n(&h->count, 0u, __ATOMIC_RELAXED);
        __atomic_store_n(&h->sense, c->sense, __ATOMIC_RELEASE);
    }
    else
        while (__atomic_load_n(&h->sense, __ATOMIC_ACQUIRE) != c->sense) sched_yield();
}
inline char *ncclSlot(ncclComm_t c, int r) { return c->base + 4096 + (size_t)r * c->slot; }
}   // namespace emu

static inline const char *ncclGetErrorString(ncclResult_t) { return "emulated NCCL error"; }
static inline ncclResult_t ncclGetUniqueId(ncclUniqueId *u)
{
    memset(u, 0, sizeof *u);
    timespec ts;
    clock_gettime(CLOCK_REALTIME, &ts);
    snprintf(u->internal, sizeof u->internal, "/ddcb200_emu_%d_%ld", (int)getpid(), (long)ts.tv_nsec);
    return ncclSuccess;
}
static inline ncclResult_t ncclCommInitRank(ncclComm_t *out, int nranks, ncclUniqueId id, int rank)
{
    ncclComm_t c = (ncclComm_t)calloc(1, sizeof(*c));
    c->rank = rank;
    c->nranks = nranks;
    c->slot = (size_t)256 << 20;   // sparse: only touched pages exist
    strncpy(c->name, id.internal, sizeof c->name - 1);
    const size_t total = 4096 + c->slot * (size_t)nranks;
    int fd = shm_open(c->name, O_CREAT | O_RDWR, 0600);
    if (fd < 0 || ftruncate(fd, (off_t)total) != 0) return ncclSystemError;
    c->base = (char *)mmap(nullptr, total, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (c->base == (char *)MAP_FAILED) return ncclSystemError;
    c->sense = 0;
    *out = c;
    emu::ncclGroupComm() = c;      // one communicator per process: ncclGroupEnd() has no comm argument
    emu::ncclBarrier(c);
    return ncclSuccess;
}
static inline ncclResult_t ncclCommDestroy(ncclComm_t c)
{
    if (!c) return ncclSuccess;
    emu::ncclBarrier(c);
    munmap(c->base, 4096 + c->slot * (size_t)c->nranks);
    if (c->rank == 0) shm_unlink(c->name);
    free(c);
    return ncclSuccess;
}
static inline ncclResult_t ncclAllReduce(const void *src, void *dst, size_t count, ncclDataType_t, ncclRedOp_t op, ncclComm_t c, cudaStream_t)
{
    if (count * 8 > c->slot) return ncclInvalidUsage;
    memcpy(emu::ncclSlot(c, c->rank), src, count * 8);
    emu::ncclBarrier(c);
    double *d = (double *)dst;
    for (size_t k = 0; k < count; k++)
    {
        double s = op == ncclMax ? -1e308 : 0.0;
        for (int r = 0; r < c->nranks; r++)
        {
            const double v = ((const double *)emu::ncclSlot(c, r))[k];
            s = op == ncclMax ? (v > s ? v : s) : s + v;
        }
        d[k] = s;
    }
    emu::ncclBarrier(c);
    return ncclSuccess;
}
// every rank's `count` elements, in rank order, to every rank
static inline ncclResult_t ncclAllGather(const void *src, void *dst, size_t count, ncclDataType_t t, ncclComm_t c, cudaStream_t)
{
    const size_t bytes = count * ncclTypeSize(t);
    if (bytes > c->slot) return ncclInvalidUsage;
    memcpy(emu::ncclSlot(c, c->rank), src, bytes);
    emu::ncclBarrier(c);
    for (int r = 0; r < c->nranks; r++) memcpy((char *)dst + (size_t)r * bytes, emu::ncclSlot(c, r), bytes);
    emu::ncclBarrier(c);
    return ncclSuccess;
}
static inline ncclResult_t ncclGroupStart() { emu::ncclGroupDepth()++; return ncclSuccess; }
static inline ncclResult_t ncclSend(const void *src, size_t count, ncclDataType_t t, int peer, ncclComm_t c, cudaStream_t)
{
    if (!emu::ncclGroupDepth()) return ncclInvalidUsage;
    emu::ncclGroupComm() = c;
    emu::ncclQueue().push_back({1, src, nullptr, count * ncclTypeSize(t), peer});
    return ncclSuccess;
}
static inline ncclResult_t ncclRecv(void *dst, size_t count, ncclDataType_t t, int peer, ncclComm_t c, cudaStream_t)
{
    if (!emu::ncclGroupDepth()) return ncclInvalidUsage;
    emu::ncclGroupComm() = c;
    emu::ncclQueue().push_back({0, nullptr, dst, count * ncclTypeSize(t), peer});
    return ncclSuccess;
}
// The product issues exactly one group per halo exchange and every rank takes part (possibly with no messages), so
// the group end is a rendezvous: senders write region [peer] of their own slot, receivers read region [me] of the peer's.
static inline ncclResult_t ncclGroupEnd()
{
    ncclComm_t c = emu::ncclGroupComm();
    if (!c || --emu::ncclGroupDepth() != 0) return c ? ncclSuccess : ncclInvalidUsage;
    const size_t region = c->slot / (size_t)c->nranks;
    for (auto &op : emu::ncclQueue())
        if (op.send)
        {
            if (op.bytes > region) return ncclInvalidUsage;
            memcpy(emu::ncclSlot(c, c->rank) + (size_t)op.peer * region, op.src, op.bytes);
        }
    emu::ncclBarrier(c);
    for (auto &op : emu::ncclQueue())
        if (!op.send) memcpy(op.dst, emu::ncclSlot(c, op.peer) + (size_t)c->rank * region, op.bytes);
    emu::ncclBarrier(c);
    emu::ncclQueue().clear();
    return ncclSuccess;
}

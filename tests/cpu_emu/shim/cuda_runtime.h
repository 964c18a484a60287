// cuda_runtime.h (CPU EMULATION SHIM) - TEST INFRASTRUCTURE ONLY.
//
// Lets the product's CUDA sources (ddcmd_b200/csrc/api.cu + *.cuh) be compiled with g++ and executed on the
// host so the kernel LOGIC (indexing, list formats, reductions, parity arithmetic) can be exercised by the
// `-m "not gpu"` tests in a container that has no GPU.  It is never built into, loaded by, or reachable from the
// product library: the product has no CPU path.  Only tests/cpu_emu/build_emu.py puts
// this directory on an include path.
//
// Execution model: blocks run one after another; the threads of a block are fibers (hand-rolled x86-64 context
// switch) scheduled round-robin.  __syncthreads() and the *_sync warp primitives yield to the scheduler, which
// releases a barrier when every live thread of the block (resp. every live lane of the warp) has arrived.
// Memory is the host heap; streams and copies are synchronous; atomics are plain read-modify-writes.
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <algorithm>
#include <functional>
#include <vector>

#define DDCB200_EMU 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__ static

struct uint3 { unsigned x, y, z; };
struct dim3
{
    unsigned x, y, z;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct alignas(16) double2 { double x, y; };
struct alignas(32) double4 { double x, y, z, w; };
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(8) float2 { float x, y; };
struct alignas(16) int4 { int x, y, z, w; };
struct alignas(8) int2 { int x, y; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
struct alignas(8) uint2 { unsigned x, y; };
static inline double2 make_double2(double x, double y) { return double2{x, y}; }
static inline double4 make_double4(double x, double y, double z, double w) { return double4{x, y, z, w}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }
static inline int2 make_int2(int x, int y) { return int2{x, y}; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }

// ---- runtime API subset ------------------------------------------------------------------
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1 };
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
typedef struct emuStream *cudaStream_t;
typedef struct emuEvent { double t; } *cudaEvent_t;
enum { cudaStreamNonBlocking = 1 };
struct cudaDeviceProp
{
    char name[64];
    int major, minor, multiProcessorCount;
    size_t sharedMemPerBlockOptin;
};
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };

namespace emu
{
inline cudaError_t &lastError() { static cudaError_t e = cudaSuccess; return e; }
inline double now()
{
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}
}   // namespace emu

static inline const char *cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "emulated CUDA error"; }
static inline cudaError_t cudaGetLastError() { cudaError_t e = emu::lastError(); emu::lastError() = cudaSuccess; return e; }
static inline cudaError_t cudaGetDeviceCount(int *n) { *n = 16; return cudaSuccess; }   // every "device" is the host
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int)
{
    memset(p, 0, sizeof *p);
    strcpy(p->name, "cpu-emulated sm_100");
    p->major = 10;
    p->multiProcessorCount = 148;
    p->sharedMemPerBlockOptin = 227 * 1024;
    return cudaSuccess;
}
static inline cudaError_t cudaMalloc(void **p, size_t n)
{
    n = (n + 255) & ~(size_t)255;
    *p = aligned_alloc(256, n ? n : 256);
    if (!*p) return cudaErrorMemoryAllocation;
    memset(*p, 0xCD, n ? n : 256);   // poison: reads of never-written device memory show up as huge garbage
    return cudaSuccess;
}
static inline cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMallocHost(void **p, size_t n) { *p = aligned_alloc(256, ((n + 255) & ~(size_t)255) ? ((n + 255) & ~(size_t)255) : 256); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
static inline cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpy2DAsync(void *d, size_t dpitch, const void *s, size_t spitch, size_t width, size_t height, cudaMemcpyKind,
                                            cudaStream_t = nullptr)
{
    for (size_t r = 0; r < height; r++) memmove((char *)d + r * dpitch, (const char *)s + r * spitch, width);
    return cudaSuccess;
}
static inline cudaError_t cudaMemset(void *d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t = nullptr) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = (cudaStream_t)malloc(8); return cudaSuccess; }
static inline cudaError_t cudaDeviceGetStreamPriorityRange(int *least, int *greatest) { *least = 0; *greatest = -5; return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithPriority(cudaStream_t *s, unsigned, int) { *s = (cudaStream_t)malloc(8); return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t s) { free(s); return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
enum { cudaEventDisableTiming = 2 };
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { *e = (cudaEvent_t)malloc(sizeof(emuEvent)); (*e)->t = 0; return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }   // streams run in program order here
static inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = (cudaEvent_t)malloc(sizeof(emuEvent)); (*e)->t = 0; return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { free(e); return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = nullptr) { e->t = emu::now(); return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b) { *ms = (float)(b->t - a->t); return cudaSuccess; }
template <class F>
static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }

// ---- fibers and the block scheduler ---------------------------------------------------------
extern "C" void emu_switch(void **save_sp, void *load_sp);
#ifdef DDCB200_EMU_IMPL
asm(".text\n.globl emu_switch\n.type emu_switch,@function\nemu_switch:\n"
    "pushq %rbp\npushq %rbx\npushq %r12\npushq %r13\npushq %r14\npushq %r15\n"
    "movq %rsp, (%rdi)\nmovq %rsi, %rsp\n"
    "popq %r15\npopq %r14\npopq %r13\npopq %r12\npopq %rbx\npopq %rbp\nret\n");
#endif

namespace emu
{
enum { READY = 0, AT_BARRIER, AT_WARP, DONE };
struct Fiber
{
    void *sp;
    int state;
};
struct Block
{
    std::vector<Fiber> fib;
    std::vector<char> stacks;
    size_t stackSize = 256 * 1024;
    void *schedSp = nullptr;
    int cur = 0;
    int nthreads = 0;
    std::function<void()> body;
    // warp exchange: two parity buffers per warp
    uint64_t xchg[32][2][32];
    uint32_t posted[32][2];
    int wop[32];
    std::vector<char> dynSmem;
    long long launches = 0, yields = 0;
};
inline Block &blk() { static Block b; return b; }
}   // namespace emu

extern uint3 threadIdx, blockIdx;
extern dim3 blockDim, gridDim;
#ifdef DDCB200_EMU_IMPL
uint3 threadIdx, blockIdx;
dim3 blockDim, gridDim;
#endif

namespace emu
{
inline void setThread(int t)
{
    threadIdx.x = t % blockDim.x;
    threadIdx.y = (t / blockDim.x) % blockDim.y;
    threadIdx.z = t / (blockDim.x * blockDim.y);
}
inline void yield(int state)
{
    Block &b = blk();
    Fiber &f = b.fib[b.cur];
    f.state = state;
    b.yields++;
    emu_switch(&f.sp, b.schedSp);
}
static void fiberEntry()
{
    Block &b = blk();
    b.body();
    b.fib[b.cur].state = DONE;
    emu_switch(&b.fib[b.cur].sp, b.schedSp);
    abort();   // a finished fiber is never resumed
}
inline void runBlock()
{
    Block &b = blk();
    const int n = b.nthreads;
    b.fib.resize(n);
    if (b.stacks.size() < (size_t)n * b.stackSize) b.stacks.resize((size_t)n * b.stackSize);
    for (int t = 0; t < n; t++)
    {
        char *top = b.stacks.data() + (size_t)(t + 1) * b.stackSize;
        top = (char *)((uintptr_t)top & ~(uintptr_t)15);
        void **sp = (void **)(top - 64);
        for (int k = 0; k < 6; k++) sp[k] = nullptr;
        sp[6] = (void *)&fiberEntry;
        sp[7] = nullptr;
        b.fib[t].sp = sp;
        b.fib[t].state = READY;
    }
    const int nwarps = (n + 31) / 32;
    for (int w = 0; w < nwarps; w++) { b.posted[w][0] = b.posted[w][1] = 0; b.wop[w] = 0; }
    int done = 0;
    while (done < n)
    {
        bool ran = false;
        for (int t = 0; t < n; t++)
            if (b.fib[t].state == READY)
            {
                b.cur = t;
                setThread(t);
                emu_switch(&b.schedSp, b.fib[t].sp);
                ran = true;
                if (b.fib[t].state == DONE) done++;
            }
        // release warp-level rendezvous
        for (int w = 0; w < nwarps; w++)
        {
            int live = 0, atw = 0;
            for (int l = 0; l < 32 && w * 32 + l < n; l++)
            {
                const int s = b.fib[w * 32 + l].state;
                if (s != DONE) live++;
                if (s == AT_WARP) atw++;
            }
            if (live > 0 && atw == live)
            {
                for (int l = 0; l < 32 && w * 32 + l < n; l++)
                    if (b.fib[w * 32 + l].state == AT_WARP) b.fib[w * 32 + l].state = READY;
                b.posted[w][(b.wop[w] + 1) & 1] = 0;
                b.wop[w]++;
                ran = true;
            }
        }
        // release the block barrier
        int live = 0, atb = 0;
        for (int t = 0; t < n; t++)
        {
            if (b.fib[t].state != DONE) live++;
            if (b.fib[t].state == AT_BARRIER) atb++;
        }
        if (live > 0 && atb == live)
        {
            for (int t = 0; t < n; t++)
                if (b.fib[t].state == AT_BARRIER) b.fib[t].state = READY;
            ran = true;
        }
        if (!ran)
        {
            fprintf(stderr, "cpu_emu: deadlock in block (%u,%u,%u): divergent barrier or warp primitive\n", blockIdx.x, blockIdx.y, blockIdx.z);
            abort();
        }
    }
}

template <class... P>
struct Launch
{
    void (*k)(P...);
    dim3 g, b;
    size_t smem;
    template <class... A>
    void operator()(A &&...args)
    {
        Block &B = blk();
        if ((size_t)b.x * b.y * b.z > 1024 || b.x * b.y * b.z == 0 || g.x == 0 || smem > 227 * 1024)
        {
            lastError() = cudaErrorInvalidValue;   // what cudaGetLastError would report for a bad configuration
            return;
        }
        B.launches++;
        gridDim = g;
        blockDim = b;
        B.nthreads = (int)(b.x * b.y * b.z);
        B.dynSmem.assign(smem + 64, 0);
        B.body = [&]() { k(static_cast<P>(args)...); };
        for (unsigned z = 0; z < g.z; z++)
            for (unsigned y = 0; y < g.y; y++)
                for (unsigned x = 0; x < g.x; x++)
                {
                    blockIdx.x = x; blockIdx.y = y; blockIdx.z = z;
                    runBlock();
                }
    }
};
template <class... P>
inline Launch<P...> launch(void (*k)(P...), dim3 g, dim3 b, size_t smem = 0, cudaStream_t = nullptr) { return Launch<P...>{k, g, b, smem}; }

inline void *dynamicSmem() { return (void *)(((uintptr_t)blk().dynSmem.data() + 31) & ~(uintptr_t)31); }

// warp rendezvous: post a 64-bit value, wait for the live lanes, return the buffer of posted values + mask
inline const uint64_t *warpPost(uint64_t v, uint32_t &mask)
{
    Block &b = blk();
    const int w = b.cur >> 5, l = b.cur & 31, par = b.wop[w] & 1;
    b.xchg[w][par][l] = v;
    b.posted[w][par] |= 1u << l;
    yield(AT_WARP);
    mask = b.posted[w][par];
    return b.xchg[w][par];
}
template <class T>
inline uint64_t toBits(T v) { uint64_t u = 0; static_assert(sizeof(T) <= 8, "shuffle width"); memcpy(&u, &v, sizeof(T)); return u; }
template <class T>
inline T fromBits(uint64_t u) { T v; memcpy(&v, &u, sizeof(T)); return v; }
}   // namespace emu

#define LAUNCH(k, ...) emu::launch(k, __VA_ARGS__)
#define EXTERN_SHARED(type, name) type *name = (type *)emu::dynamicSmem()

static inline void __syncthreads() { emu::yield(emu::AT_BARRIER); }
static inline void __syncwarp(unsigned = 0xffffffffu) { uint32_t m; emu::warpPost(0, m); }
template <class T>
static inline T __shfl_xor_sync(unsigned, T v, int o)
{
    uint32_t m;
    const uint64_t *x = emu::warpPost(emu::toBits(v), m);
    const int src = (emu::blk().cur & 31) ^ o;
    return (m >> src) & 1u ? emu::fromBits<T>(x[src]) : v;
}
template <class T>
static inline T __shfl_sync(unsigned, T v, int src)
{
    uint32_t m;
    const uint64_t *x = emu::warpPost(emu::toBits(v), m);
    src &= 31;
    return (m >> src) & 1u ? emu::fromBits<T>(x[src]) : v;
}
template <class T>
static inline T __shfl_down_sync(unsigned, T v, int d)
{
    uint32_t m;
    const uint64_t *x = emu::warpPost(emu::toBits(v), m);
    const int src = (emu::blk().cur & 31) + d;
    return (src < 32 && ((m >> src) & 1u)) ? emu::fromBits<T>(x[src]) : v;
}
template <class T>
static inline T __shfl_up_sync(unsigned, T v, int d)
{
    uint32_t m;
    const uint64_t *x = emu::warpPost(emu::toBits(v), m);
    const int src = (emu::blk().cur & 31) - d;
    return (src >= 0 && ((m >> src) & 1u)) ? emu::fromBits<T>(x[src]) : v;
}
static inline unsigned __ballot_sync(unsigned, int pred)
{
    uint32_t m;
    const uint64_t *x = emu::warpPost(pred ? 1u : 0u, m);
    unsigned r = 0;
    for (int l = 0; l < 32; l++)
        if (((m >> l) & 1u) && x[l]) r |= 1u << l;
    return r;
}
static inline int __any_sync(unsigned mk, int pred) { return __ballot_sync(mk, pred) != 0u; }
static inline int __all_sync(unsigned mk, int pred)
{
    uint32_t m;
    const uint64_t *x = emu::warpPost(pred ? 1u : 0u, m);
    for (int l = 0; l < 32; l++)
        if (((m >> l) & 1u) && !x[l]) return 0;
    (void)mk;
    return 1;
}
static inline int __reduce_max_sync(unsigned, int v)
{
    uint32_t m;
    const uint64_t *x = emu::warpPost(emu::toBits(v), m);
    int r = v;
    for (int l = 0; l < 32; l++)
        if ((m >> l) & 1u) r = std::max(r, emu::fromBits<int>(x[l]));
    return r;
}
static inline int __reduce_min_sync(unsigned, int v)
{
    uint32_t m;
    const uint64_t *x = emu::warpPost(emu::toBits(v), m);
    int r = v;
    for (int l = 0; l < 32; l++)
        if ((m >> l) & 1u) r = std::min(r, emu::fromBits<int>(x[l]));
    return r;
}
static inline unsigned __activemask()
{
    emu::Block &b = emu::blk();
    unsigned r = 0;
    const int w = b.cur >> 5;
    for (int l = 0; l < 32 && w * 32 + l < b.nthreads; l++)
        if (b.fib[w * 32 + l].state != emu::DONE) r |= 1u << l;
    return r;
}

// ---- intrinsics --------------------------------------------------------------------------
static inline double __dadd_rn(double a, double b) { return a + b; }   // built with -ffp-contract=off: no fusion
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline double __fma_rn(double a, double b, double c) { return fma(a, b, c); }
static inline double __dsqrt_rn(double a) { return sqrt(a); }
static inline double __ull2double_rn(unsigned long long v) { return (double)v; }
static inline long long __double_as_longlong(double v) { long long r; memcpy(&r, &v, 8); return r; }
static inline float __double2float_ru(double v) { float f = (float)v; return ((double)f < v) ? nextafterf(f, INFINITY) : f; }
static inline int __double2hiint(double v) { long long r; memcpy(&r, &v, 8); return (int)(r >> 32); }
static inline double __longlong_as_double(long long v) { double r; memcpy(&r, &v, 8); return r; }
static inline int __float_as_int(float v) { int r; memcpy(&r, &v, 4); return r; }
static inline float __int_as_float(int v) { float r; memcpy(&r, &v, 4); return r; }
static inline unsigned __float_as_uint(float v) { unsigned r; memcpy(&r, &v, 4); return r; }
static inline float __uint_as_float(unsigned v) { float r; memcpy(&r, &v, 4); return r; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
static inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
static inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
static inline double rsqrt(double x) { return 1.0 / sqrt(x); }
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
using std::max;
using std::min;
static inline unsigned min(unsigned a, int b) { return a < (unsigned)b ? a : (unsigned)b; }
static inline unsigned min(int a, unsigned b) { return (unsigned)a < b ? (unsigned)a : b; }

template <class T>
static inline T atomicAdd(T *p, T v) { T o = *p; *p = o + v; return o; }
static inline int atomicAdd(int *p, unsigned v) { int o = *p; *p = o + (int)v; return o; }
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned v) { unsigned long long o = *p; *p = o + v; return o; }
template <class T>
static inline T atomicMax(T *p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <class T>
static inline T atomicMin(T *p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <class T>
static inline T atomicOr(T *p, T v) { T o = *p; *p = o | v; return o; }
template <class T>
static inline T atomicAnd(T *p, T v) { T o = *p; *p = o & v; return o; }
template <class T>
static inline T atomicXor(T *p, T v) { T o = *p; *p = o ^ v; return o; }
template <class T>
static inline T atomicExch(T *p, T v) { T o = *p; *p = v; return o; }
template <class T>
static inline T atomicCAS(T *p, T c, T v) { T o = *p; if (o == c) *p = v; return o; }

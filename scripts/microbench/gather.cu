// gather.cu - what a warp-wide gather of 32-byte (and 16-byte) records costs in the SM's load-store data pipe, by access pattern.
// Evidence for the layout of the pair rows: k_pair2 is bound by l1tex data-pipe wavefronts, one per gathered
// partner.  Every pattern reads a table that stays in L1/L2 so that the pipe, not DRAM, is measured.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gather gather.cu && ./gather
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ double4 ld32(const double4 *p)
{
    double4 r;
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p));
    return r;
}

template <int BYTES>
__global__ void __launch_bounds__(128) k_gather(const double4 *__restrict__ tab, const uint32_t *__restrict__ idx, int steps, int nThreads, double *out)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    double acc = 0.0;
    uint32_t e[4];
#pragma unroll
    for (int u = 0; u < 4; u++) e[u] = idx[(size_t)u * nThreads + t];
    for (int k = 0; k < steps; k += 4)
    {
        double4 p[4];
#pragma unroll
        for (int u = 0; u < 4; u++)
        {
            if (BYTES == 32) p[u] = ld32(tab + e[u]);
            else
            {
                const double2 q = __ldg((const double2 *)tab + e[u]);
                p[u] = make_double4(q.x, q.y, 0.0, 0.0);
            }
        }
#pragma unroll
        for (int u = 0; u < 4; u++) e[u] = (k + 4 + u < steps) ? idx[(size_t)(k + 4 + u) * nThreads + t] : 0u;
#pragma unroll
        for (int u = 0; u < 4; u++) acc += p[u].x + p[u].y + p[u].z + p[u].w;
    }
    if (acc == 12345.678) out[t] = acc;
}

static uint32_t rnd(uint64_t &s) { s = s * 6364136223846793005ull + 1442695040888963407ull; return (uint32_t)(s >> 33); }

int main()
{
    const int nSM = 148, ctasPerSM = 8, threads = 128;
    const int nThreads = nSM * ctasPerSM * threads, steps = 256;
    struct Tab { const char *name; int records; } tabs[] = {{"L1-resident (2048 rec = 64 KB)", 2048}, {"L2-resident (1M rec = 32 MB)", 1 << 20}};
    struct Pat { const char *name; int bytes; int group; int aligned; int stride16; } pats[] = {
        {"32B random per lane", 32, 1, 0, 0},
        {"32B pairs of lanes consecutive, aligned", 32, 2, 1, 0},
        {"32B quads of lanes consecutive, aligned (one line)", 32, 4, 1, 0},
        {"32B quads of lanes consecutive, unaligned", 32, 4, 0, 0},
        {"32B octets of lanes consecutive, aligned (two lines)", 32, 8, 1, 0},
        {"32B lanes l and l+16 same line (quads split across half-warps)", 32, 4, 1, 1},
        {"32B all 32 lanes consecutive (coalesced)", 32, 32, 1, 0},
        {"32B all lanes same record (broadcast)", 32, -1, 1, 0},
        {"16B random per lane", 16, 1, 0, 0},
        {"16B octets of lanes consecutive, aligned (one line)", 16, 8, 1, 0},
        {"16B quads of lanes consecutive, aligned", 16, 4, 1, 0},
    };
    double *out; cudaMalloc(&out, sizeof(double) * nThreads);
    uint32_t *dIdx; cudaMalloc(&dIdx, sizeof(uint32_t) * (size_t)nThreads * steps);
    std::vector<uint32_t> h((size_t)nThreads * steps);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("threads %d steps %d, SM clock attr %d kHz\n", nThreads, steps, clk);
    for (auto &tb : tabs)
    {
        double4 *tab; cudaMalloc(&tab, sizeof(double4) * tb.records); cudaMemset(tab, 0, sizeof(double4) * tb.records);
        printf("---- table %s\n", tb.name);
        for (auto &p : pats)
        {
            const int recs = p.bytes == 32 ? tb.records : 2 * tb.records;      // 16-byte records: twice as many in the same table
            uint64_t s = 12345;
            for (int k = 0; k < steps; k++)
                for (int w = 0; w < nThreads / 32; w++)
                {
                    uint32_t base[32];
                    for (int g = 0; g < 32; g++) base[g] = rnd(s) % recs;
                    for (int l = 0; l < 32; l++)
                    {
                        uint32_t v;
                        if (p.group < 0) v = base[0];
                        else if (p.stride16) { const int g = l % 16 / 2, m = (l / 16) * 2 + (l & 1); v = (base[g] & ~3u) + m; }
                        else
                        {
                            const int g = l / p.group, m = l % p.group;
                            uint32_t b = base[g];
                            if (p.aligned) b -= b % (p.group > 32 ? 32 : p.group);
                            v = (b + m) % recs;
                        }
                        h[(size_t)k * nThreads + w * 32 + l] = v;
                    }
                }
            cudaMemcpy(dIdx, h.data(), sizeof(uint32_t) * h.size(), cudaMemcpyHostToDevice);
            float best = 1e30f;
            for (int rep = 0; rep < 5; rep++)
            {
                cudaEventRecord(e0);
                if (p.bytes == 32) k_gather<32><<<nThreads / threads, threads>>>(tab, dIdx, steps, nThreads, out);
                else k_gather<16><<<nThreads / threads, threads>>>(tab, dIdx, steps, nThreads, out);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                if (rep > 0 && ms < best) best = ms;
            }
            const double warpInstr = (double)nThreads / 32 * steps;
            const double cyc = best * 1e-3 * 1.965e9 * nSM / warpInstr;      // SM cycles per warp-wide gather at 1965 MHz
            printf("%-66s %8.3f ms  %6.2f cycles per warp gather  (%.2f lanes/clk/SM)\n", p.name, best, cyc, 32.0 / cyc);
        }
        cudaFree(tab);
    }
    cudaError_t err = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(err));
    return err != cudaSuccess;
}

// ddc.cuh - ddc-style spatial decomposition across the GPUs of one NVSwitch box.
//
// Replaces, for the Martini step, ddcAssignment (src/ddcAssignment.c:64-107, nearest domain
// centre of a lattice of centres = a brick), ddcRuleMolecule / bioMartiniRule (molecules stay
// whole and follow their ownership bead, src/ddcRuleMolecule.c:43), ddcSendRecvTables
// (src/ddcSendRecv.c:41-277: who needs which of my beads as ghosts), ddcUpdate (per-step halo
// of ghost positions, src/ddcUpdate.c:40-88) and the 24-double reduction of eval_energyInfo
// (src/energyInfo.c:9-63).
//
// B200-first choices (DESIGN.md "Multi-GPU"):
//   * Every rank holds the static bead tables.  At a re-domain step (every DDC updateRate steps,
//     together with the list rebuild) the dynamic state is replicated once with one in-place
//     all-reduce over NVSwitch (each element has exactly one non-zero contributor, so the sum is
//     exact); ownership, ghost sets and BOTH ends of every send/recv list are then derived
//     independently - and identically - on every rank from the same replicated data.  There
//     is no handshake, no count exchange and no particle-migration message.
//   * Pair rows are full (both directions), so a cross-boundary pair is evaluated by both owners
//     and ddcUpdateForce's force back-communication has no message at all; per step the only
//     exchange is ghost positions (24 B per ghost).
//
// The geometric predicates are __host__ __device__ so the CPU-side planner used by the
// world_size-2 tests (ddcb200_ddcPlan) runs the very same code as the kernels.
#pragma once
#include "engine.cuh"

#define DDC_MAXRANKS 16
#define DDC_GHOST_BIT 0x8000000000000000ull   // bit 63 of pos4.w

struct DdcGeom
{
    int nranks, me;
    int lat[3];
    double L[3], hL[3];
    double rlist2;   // (rmax + deltaR)^2 with a relative safety margin
};

struct DdcBoxes   // bounding boxes of the ranks' local beads, relative to their brick centres
{
    double lo[DDC_MAXRANKS][3], hi[DDC_MAXRANKS][3];
};

#define HD __host__ __device__ __forceinline__

HD int ddcBrickOf(double x, double y, double z, const DdcGeom &g)
{
    const double p[3] = {x, y, z};
    int idx[3];
    for (int a = 0; a < 3; a++)
    {
        int i = (int)floor((p[a] + g.hL[a]) / g.L[a] * g.lat[a]);
        // positions are kept inside the box by backInBox_fast; a bead exactly on the upper face folds to brick 0
        i %= g.lat[a];
        if (i < 0) i += g.lat[a];
        idx[a] = i;
    }
    return idx[0] + g.lat[0] * (idx[1] + g.lat[1] * idx[2]);
}

HD void ddcBrickCentre(int r, const DdcGeom &g, double c[3])
{
    const int ix = r % g.lat[0], iy = (r / g.lat[0]) % g.lat[1], iz = r / (g.lat[0] * g.lat[1]);
    c[0] = -g.hL[0] + (ix + 0.5) * g.L[0] / g.lat[0];
    c[1] = -g.hL[1] + (iy + 0.5) * g.L[1] / g.lat[1];
    c[2] = -g.hL[2] + (iz + 0.5) * g.L[2] / g.lat[2];
}

HD double ddcMinImg(double d, double L, double hL)
{
    if (d > hL) d -= L;
    if (d < -hL) d += L;
    return d;
}

HD double ddcExcess(double d, double lo, double hi)
{
    const double a = lo - d, b = d - hi;
    const double m = a > b ? a : b;
    return m > 0.0 ? m : 0.0;
}

// Is bead p within the list range of ANY point of rank r's bounding box (any periodic image)?
// Never misses a bead that is within range of one of r's local beads (superset), see DESIGN.md.
HD bool ddcNear(double x, double y, double z, int r, const DdcGeom &g, const DdcBoxes &bx)
{
    double c[3];
    ddcBrickCentre(r, g, c);
    const double p[3] = {x, y, z};
    double e2 = 0.0;
    for (int a = 0; a < 3; a++)
    {
        const double d = ddcMinImg(p[a] - c[a], g.L[a], g.hL[a]);
        double e = ddcExcess(d, bx.lo[r][a], bx.hi[r][a]);
        const double e1 = ddcExcess(d - g.L[a], bx.lo[r][a], bx.hi[r][a]);
        const double e2a = ddcExcess(d + g.L[a], bx.lo[r][a], bx.hi[r][a]);
        if (e1 < e) e = e1;
        if (e2a < e) e = e2a;
        e2 += e * e;
    }
    return e2 < g.rlist2;
}

#if defined(__CUDACC__) || defined(DDCB200_EMU)
// monotone double <-> uint64 encoding for atomicMin/atomicMax
__device__ __forceinline__ unsigned long long encOrd(double v)
{
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double decOrd(unsigned long long e)
{
    const unsigned long long b = (e & 0x8000000000000000ull) ? (e & 0x7fffffffffffffffull) : ~e;
    return __longlong_as_double((long long)b);
}

__device__ __forceinline__ bool isGhostW(double w) { return (((unsigned long long)__double_as_longlong(w)) & DDC_GHOST_BIT) != 0ull; }

// 1. every local bead writes its state at its bead index of the (zeroed) replicated arrays
__global__ void k_ddc_scatter(int nIon, int64_t nGlobal, const double4 *__restrict__ pos, const double *__restrict__ vx,
                              const double *__restrict__ vy, const double *__restrict__ vz, double *__restrict__ gs)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nIon) return;
    const double4 p = pos[s];
    const unsigned long long w = (unsigned long long)__double_as_longlong(p.w);
    if (w & DDC_GHOST_BIT) return;
    const size_t b = (size_t)((w >> 32) & 0x7fffffffull);
    gs[b] = p.x;
    gs[(size_t)nGlobal + b] = p.y;
    gs[2 * (size_t)nGlobal + b] = p.z;
    gs[3 * (size_t)nGlobal + b] = vx[s];
    gs[4 * (size_t)nGlobal + b] = vy[s];
    gs[5 * (size_t)nGlobal + b] = vz[s];
}

// 2. owner of every bead = brick of its molecule's ownership bead; bounding box of every rank's beads
//    (encoded min/max; boxEnc[r*6 + a] = min, boxEnc[r*6 + 3 + a] = max)
__global__ void __launch_bounds__(256)
k_ddc_owner(int64_t nGlobal, const double *__restrict__ gs, const int *__restrict__ ownerBead, DdcGeom g, int *__restrict__ owner,
            unsigned long long *__restrict__ boxEnc)
{
    __shared__ unsigned long long sEnc[DDC_MAXRANKS * 6];
    for (int k = threadIdx.x; k < g.nranks * 6; k += blockDim.x) sEnc[k] = ((k % 6) < 3) ? ~0ull : 0ull;
    __syncthreads();
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b < nGlobal)
    {
        const int ob = ownerBead[b];
        const int r = ddcBrickOf(gs[ob], gs[(size_t)nGlobal + ob], gs[2 * (size_t)nGlobal + ob], g);
        owner[b] = r;
        double c[3];
        ddcBrickCentre(r, g, c);
        for (int a = 0; a < 3; a++)
        {
            const double d = ddcMinImg(gs[(size_t)a * nGlobal + b] - c[a], g.L[a], g.hL[a]);
            const unsigned long long e = encOrd(d);
            atomicMin(&sEnc[r * 6 + a], e);
            atomicMax(&sEnc[r * 6 + 3 + a], e);
        }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < g.nranks * 6; k += blockDim.x)
    {
        if ((k % 6) < 3) { if (sEnc[k] != ~0ull) atomicMin(&boxEnc[k], sEnc[k]); }
        else { if (sEnc[k] != 0ull) atomicMax(&boxEnc[k], sEnc[k]); }
    }
}

__global__ void k_ddc_boxes(int nranks, const unsigned long long *__restrict__ boxEnc, DdcBoxes *__restrict__ bx)
{
    const int k = threadIdx.x;
    if (k >= nranks * 6) return;
    const int r = k / 6, a = k % 6;
    const unsigned long long e = boxEnc[k];
    if (a < 3) bx->lo[r][a] = (e == ~0ull) ? 1e300 : decOrd(e);      // empty rank: nothing is near it
    else bx->hi[r][a - 3] = (e == 0ull) ? -1e300 : decOrd(e);
}

// 3. classify every bead for this rank: bit p (p != me) = mine and needed by rank p as a ghost;
//    bit 16+p = owned by p and needed here as a ghost; bit 31 = mine.  Counts locals / ghosts.
__global__ void __launch_bounds__(256)
k_ddc_mask(int64_t nGlobal, const double *__restrict__ gs, const int *__restrict__ owner, DdcGeom g, const DdcBoxes *__restrict__ bxp,
           uint32_t *__restrict__ mask, int *__restrict__ counters)
{
    __shared__ DdcBoxes bx;
    for (int k = threadIdx.x; k < (int)(sizeof(DdcBoxes) / sizeof(double)); k += blockDim.x) ((double *)&bx)[k] = ((const double *)bxp)[k];
    __syncthreads();
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t m = 0u;
    if (b < nGlobal)
    {
        const int o = owner[b];
        const double x = gs[b], y = gs[(size_t)nGlobal + b], z = gs[2 * (size_t)nGlobal + b];
        if (o == g.me)
        {
            m = 0x80000000u;
            for (int p = 0; p < g.nranks; p++)
                if (p != g.me && ddcNear(x, y, z, p, g, bx)) m |= 1u << p;
        }
        else if (ddcNear(x, y, z, g.me, g, bx)) m = 1u << (16 + o);
        mask[b] = m;
    }
    const unsigned nl = __popc(__ballot_sync(0xffffffffu, (m & 0x80000000u) != 0u));
    const unsigned ng = __popc(__ballot_sync(0xffffffffu, (m & 0xffff0000u) != 0u && !(m & 0x80000000u)));
    if ((threadIdx.x & 31) == 0)
    {
        if (nl) atomicAdd(&counters[0], (int)nl);
        if (ng) atomicAdd(&counters[1], (int)ng);
    }
}

// 4. rebuild the slot arrays from the replicated state: locals and ghosts in arbitrary order (the
//    cell sort that follows orders slots by (cell, sub-cell, bead) whatever the order here)
__global__ void k_ddc_select(int64_t nGlobal, const double *__restrict__ gs, const uint32_t *__restrict__ mask,
                             const uint64_t *__restrict__ wOfBead, int *__restrict__ counter, double4 *__restrict__ pos,
                             double *__restrict__ vx, double *__restrict__ vy, double *__restrict__ vz, int *__restrict__ beadOfSlot,
                             int *__restrict__ slotOfBead)
{
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nGlobal) return;
    const uint32_t m = mask[b];
    if (m == 0u) return;
    const bool local = (m & 0x80000000u) != 0u;
    const int s = atomicAdd(counter, 1);
    unsigned long long w = wOfBead[b];
    if (!local) w |= DDC_GHOST_BIT;
    pos[s] = make_double4(gs[b], gs[(size_t)nGlobal + b], gs[2 * (size_t)nGlobal + b], __longlong_as_double((long long)w));
    vx[s] = local ? gs[3 * (size_t)nGlobal + b] : 0.0;
    vy[s] = local ? gs[4 * (size_t)nGlobal + b] : 0.0;
    vz[s] = local ? gs[5 * (size_t)nGlobal + b] : 0.0;
    beadOfSlot[s] = (int)b;
    slotOfBead[b] = s;
}

// 5. ordered multi-column compaction of the mask bits into bead lists (ascending bead index), so
//    sender and receiver build the same list in the same order without talking to each other.
//    unit = one warp = 32 consecutive beads; cnt[col][unit].
__global__ void __launch_bounds__(256)
k_ddc_colcount(int64_t nGlobal, const uint32_t *__restrict__ mask, uint32_t colBits, int nUnits, int *__restrict__ cnt)
{
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t m = (b < nGlobal) ? mask[b] : 0u;
    const int unit = (int)(b >> 5);
    if (unit >= nUnits) return;
    const int lane = threadIdx.x & 31;
    uint32_t bits = colBits;
    int col = 0;
    while (bits)
    {
        const int bit = __ffs(bits) - 1;
        bits &= bits - 1;
        const unsigned v = __ballot_sync(0xffffffffu, (m >> bit) & 1u);
        if (lane == 0) cnt[(size_t)col * nUnits + unit] = __popc(v);
        col++;
    }
}

// one CTA per column: exclusive scan over the units (in place), total to colTotal[col]
__global__ void __launch_bounds__(1024)
k_ddc_colscan(int nUnits, int *__restrict__ cnt, int *__restrict__ colTotal)
{
    __shared__ int sums[1024];
    int *c = cnt + (size_t)blockIdx.x * nUnits;
    const int per = (nUnits + blockDim.x - 1) / blockDim.x;
    const int lo = min(nUnits, (int)threadIdx.x * per), hi = min(nUnits, lo + per);
    int s = 0;
    for (int i = lo; i < hi; i++) s += c[i];
    sums[threadIdx.x] = s;
    __syncthreads();
    for (int o = 1; o < blockDim.x; o <<= 1)
    {
        int v = (threadIdx.x >= o) ? sums[threadIdx.x - o] : 0;
        __syncthreads();
        sums[threadIdx.x] += v;
        __syncthreads();
    }
    int run = sums[threadIdx.x] - s;
    for (int i = lo; i < hi; i++)
    {
        const int v = c[i];
        c[i] = run;
        run += v;
    }
    if (threadIdx.x == blockDim.x - 1) colTotal[blockIdx.x] = sums[threadIdx.x];
}

__global__ void __launch_bounds__(256)
k_ddc_colscatter(int64_t nGlobal, const uint32_t *__restrict__ mask, uint32_t colBits, int nUnits, const int *__restrict__ off,
                 const int *__restrict__ colStart, int *__restrict__ list)
{
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t m = (b < nGlobal) ? mask[b] : 0u;
    const int unit = (int)(b >> 5);
    if (unit >= nUnits) return;
    const unsigned lt = (1u << (threadIdx.x & 31)) - 1u;
    uint32_t bits = colBits;
    int col = 0;
    while (bits)
    {
        const int bit = __ffs(bits) - 1;
        bits &= bits - 1;
        const bool on = (m >> bit) & 1u;
        const unsigned v = __ballot_sync(0xffffffffu, on);
        if (on) list[colStart[col] + off[(size_t)col * nUnits + unit] + __popc(v & lt)] = (int)b;
        col++;
    }
}

__global__ void k_ddc_toslots(int n, const int *__restrict__ beads, const int *__restrict__ slotOfBead, int *__restrict__ slots)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) slots[i] = slotOfBead[beads[i]];
}

// ---- per step: ghost positions ------------------------------------------------------------
__global__ void k_halo_pack(int n, const int *__restrict__ slot, const double4 *__restrict__ pos, double *__restrict__ buf)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double4 p = pos[slot[i]];
    buf[3 * (size_t)i] = p.x;
    buf[3 * (size_t)i + 1] = p.y;
    buf[3 * (size_t)i + 2] = p.z;
}

__global__ void k_halo_unpack(int n, const int *__restrict__ slot, const double *__restrict__ buf, double4 *__restrict__ pos,
                              const double *__restrict__ bx, const double *__restrict__ by, const double *__restrict__ bz, PairConst pc,
                              unsigned long long *__restrict__ dmax2)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double disp2 = 0.0;
    if (i < n)
    {
        const int s = slot[i];
        const double x = buf[3 * (size_t)i], y = buf[3 * (size_t)i + 1], z = buf[3 * (size_t)i + 2];
        pos[s].x = x;
        pos[s].y = y;
        pos[s].z = z;
        // ghosts move too: their displacement since the build enters the same bound as the locals'
        double dx = ddcMinImg(x - bx[s], pc.hxx, pc.hhx), dy = ddcMinImg(y - by[s], pc.hyy, pc.hhy), dz = ddcMinImg(z - bz[s], pc.hzz, pc.hhz);
        disp2 = dx * dx + dy * dy + dz * dz;
    }
    for (int o = 16; o > 0; o >>= 1) disp2 = fmax(disp2, __shfl_xor_sync(0xffffffffu, disp2, o));
    if ((threadIdx.x & 31) == 0)
    {
        const unsigned long long bits = (unsigned long long)__double_as_longlong(disp2);
        if (bits > *(volatile unsigned long long *)dmax2) atomicMax(dmax2, bits);
    }
}

// bead ids of the local slots, for getLocalBeads / getState
__global__ void k_ddc_list_locals(int nIon, const double4 *__restrict__ pos, int *__restrict__ counter, int *__restrict__ beads)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nIon) return;
    const unsigned long long w = (unsigned long long)__double_as_longlong(pos[s].w);
    if (w & DDC_GHOST_BIT) return;
    beads[atomicAdd(counter, 1)] = (int)((w >> 32) & 0x7fffffffull);
}
#endif

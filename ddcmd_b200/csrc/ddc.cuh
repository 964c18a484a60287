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
//   * Every rank holds the STATIC bead tables (species, gid, molecule and bonded-term tables); the dynamic state lives only
//     on the owner (locals) and on the ranks whose bricks it borders (ghost positions).  A re-domain step (every DDC
//     updateRate steps, together with the list rebuild) moves emigrants and rebuilds the ghost lists with two grouped
//     ncclSend/ncclRecv exchanges whose sizes come from two small all-gathers; all kernels run over this rank's beads only.
//   * Pair rows are full (both directions), so a cross-boundary pair is evaluated by both owners and ddcUpdateForce's
//     force back-communication has no message at all; per step the only exchange is ghost positions (24 B per ghost), on
//     its own stream, hidden behind the pair work of the rows that touch no ghost.
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

// ---- re-domain (every DDC updateRate steps, together with the list rebuild) ---------------------------------------------
// Two exchanges between the ranks, as ddcAssignment + ddcSendRecvTables (src/ddcAssignment.c:64-107, src/ddcSendRecv.c:41-277):
//   A. migration: every local bead goes to the brick of its molecule's ownership bead; emigrants travel as records
//      {bead, r, v[, random state]} straight to their new owner;
//   B. ghosts: every rank tells the others the bounding box of its new locals, and each owner sends every local bead that is
//      within the list radius of a peer's box to that peer.  The order of those records IS the order of the per-step halo
//      messages until the next re-domain, on both sides.
// Every kernel runs over this rank's beads only; nothing is sized by the global bead count.
struct DdcWork                    // device scratch of one re-domain (zeroed before each)
{
    int sendMig[DDC_MAXRANKS];    // emigrants per destination
    int sendGhost[DDC_MAXRANKS];  // ghost copies per peer
    int fillMig[DDC_MAXRANKS], fillGhost[DDC_MAXRANKS];
    int nStay;
    int error;                    // bit 0: a molecule is not whole on this rank (its ownership bead is not local)
    unsigned long long boxEnc[6]; // bounding box of the new locals about the brick centre, order-encoded min[3] max[3]
};
struct DdcOffsets
{
    int off[DDC_MAXRANKS];
};
#define DDC_ROW 32                // ints per rank in the gathered count rows: [0,16) counts, [16] error flags, [17] locals

__device__ __forceinline__ void warpCountByRank(int r, bool on, int nranks, int *counts)
{
    const int lane = threadIdx.x & 31;
    for (int p = 0; p < nranks; p++)
    {
        const unsigned m = __ballot_sync(0xffffffffu, on && r == p);
        if (m && lane == (__ffs(m) - 1)) atomicAdd(&counts[p], __popc(m));
    }
}

// A1. destination of every local bead = brick of its molecule's ownership bead (ddcRuleMolecule, src/ddcRuleMolecule.c:43)
__global__ void __launch_bounds__(256)
k_rd_dest(int nIon, const double4 *__restrict__ pos, const int *__restrict__ slotOfBead, const int *__restrict__ ownerBead, DdcGeom g,
          int *__restrict__ dest, DdcWork *__restrict__ work)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    int r = -1;
    if (s < nIon)
    {
        const double4 p = pos[s];
        const unsigned long long w = (unsigned long long)__double_as_longlong(p.w);
        if (!(w & DDC_GHOST_BIT))
        {
            const int ob = ownerBead[(int)((w >> 32) & 0x7fffffffull)];
            const int so = slotOfBead[ob];
            if (so < 0 || isGhostW(pos[so].w))
            {
                atomicOr(&work->error, 1);
                r = g.me;
            }
            else
            {
                const double4 po = pos[so];
                r = ddcBrickOf(po.x, po.y, po.z, g);
            }
        }
        dest[s] = r;
    }
    warpCountByRank(r, r >= 0 && r != g.me, g.nranks, work->sendMig);
}

// A4. emigrant records: {bead, x, y, z, vx, vy, vz[, random state]}; the order inside a message is arbitrary (the cell sort
// that follows orders the slots by (cell, sub-cell, bead) whatever the order here)
__global__ void k_rd_pack(int nIon, const double4 *__restrict__ pos, const double *__restrict__ vx, const double *__restrict__ vy,
                          const double *__restrict__ vz, const int *__restrict__ dest, int me, DdcWork *__restrict__ work, DdcOffsets off,
                          int rec, const uint64_t *__restrict__ rngState, double *__restrict__ buf)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nIon) return;
    const int d = dest[s];
    if (d < 0 || d == me) return;
    const int k = off.off[d] + atomicAdd(&work->fillMig[d], 1);
    const double4 p = pos[s];
    const unsigned long long w = (unsigned long long)__double_as_longlong(p.w);
    const int bead = (int)((w >> 32) & 0x7fffffffull);
    double *o = buf + (size_t)k * rec;
    o[0] = (double)bead;
    o[1] = p.x; o[2] = p.y; o[3] = p.z;
    o[4] = vx[s]; o[5] = vy[s]; o[6] = vz[s];
    if (rec > 7) o[7] = __longlong_as_double((long long)rngState[bead]);
}

__global__ void k_rd_clear(int nIon, const int *__restrict__ beadOfSlot, int *__restrict__ slotOfBead)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < nIon) slotOfBead[beadOfSlot[s]] = -1;
}

// A6. the beads that stay, compacted into the other buffer set
__global__ void __launch_bounds__(256)
k_rd_compact(int nIon, const double4 *__restrict__ pos, const double *__restrict__ vx, const double *__restrict__ vy,
             const double *__restrict__ vz, const int *__restrict__ bead, const int *__restrict__ dest, int me, DdcWork *__restrict__ work,
             double4 *__restrict__ posN, double *__restrict__ vxN, double *__restrict__ vyN, double *__restrict__ vzN,
             int *__restrict__ beadN, int *__restrict__ slotOfBead)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    const bool stay = s < nIon && dest[s] == me;
    const unsigned m = __ballot_sync(0xffffffffu, stay);
    const int lane = threadIdx.x & 31;
    int base = 0;
    if (m && lane == (__ffs(m) - 1)) base = atomicAdd(&work->nStay, __popc(m));
    base = __shfl_sync(0xffffffffu, base, m ? __ffs(m) - 1 : 0);
    if (!stay) return;
    const int k = base + __popc(m & ((1u << lane) - 1u));
    posN[k] = pos[s];
    vxN[k] = vx[s]; vyN[k] = vy[s]; vzN[k] = vz[s];
    const int b = bead[s];
    beadN[k] = b;
    slotOfBead[b] = k;
}

__global__ void k_rd_unpack(int n, const double *__restrict__ buf, int rec, int base, const uint64_t *__restrict__ wOfBead,
                            double4 *__restrict__ pos, double *__restrict__ vx, double *__restrict__ vy, double *__restrict__ vz,
                            int *__restrict__ beadOfSlot, int *__restrict__ slotOfBead, uint64_t *__restrict__ rngState)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double *o = buf + (size_t)i * rec;
    const int b = (int)o[0];
    const int k = base + i;
    pos[k] = make_double4(o[1], o[2], o[3], __longlong_as_double((long long)wOfBead[b]));
    vx[k] = o[4]; vy[k] = o[5]; vz[k] = o[6];
    beadOfSlot[k] = b;
    slotOfBead[b] = k;
    if (rec > 7) rngState[b] = (uint64_t)__double_as_longlong(o[7]);
}

// B1. bounding box of my new locals about my brick centre (nearest image), for the peers' ghost test
__global__ void __launch_bounds__(256)
k_rd_bbox(int n, const double4 *__restrict__ pos, DdcGeom g, DdcWork *__restrict__ work)
{
    __shared__ unsigned long long sEnc[6];
    if (threadIdx.x < 6) sEnc[threadIdx.x] = threadIdx.x < 3 ? ~0ull : 0ull;
    __syncthreads();
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < n)
    {
        double c[3];
        ddcBrickCentre(g.me, g, c);
        const double4 p = pos[s];
        const double q[3] = {p.x, p.y, p.z};
        for (int a = 0; a < 3; a++)
        {
            const unsigned long long e = encOrd(ddcMinImg(q[a] - c[a], g.L[a], g.hL[a]));
            atomicMin(&sEnc[a], e);
            atomicMax(&sEnc[3 + a], e);
        }
    }
    __syncthreads();
    if (threadIdx.x < 3) atomicMin(&work->boxEnc[threadIdx.x], sEnc[threadIdx.x]);
    else if (threadIdx.x < 6) atomicMax(&work->boxEnc[threadIdx.x], sEnc[threadIdx.x]);
}

// counts (and the box) of this rank as one row for the all-gather
__global__ void k_rd_row(const DdcWork *__restrict__ work, int phase, int nLocal, int *__restrict__ row, double *__restrict__ box6)
{
    const int k = threadIdx.x;
    if (k < DDC_MAXRANKS) row[k] = phase == 0 ? work->sendMig[k] : work->sendGhost[k];
    if (k == 16) row[16] = work->error;
    if (k == 17) row[17] = nLocal;
    if (k > 17 && k < DDC_ROW) row[k] = 0;
    if (box6 && k < 6)
    {
        const unsigned long long e = work->boxEnc[k];
        box6[k] = k < 3 ? ((e == ~0ull) ? 1e300 : decOrd(e)) : ((e == 0ull) ? -1e300 : decOrd(e));   // empty rank: nothing is near it
    }
}

__global__ void k_rd_boxes(int nranks, const double *__restrict__ all6, DdcBoxes *__restrict__ bx)
{
    const int k = threadIdx.x;
    if (k >= nranks * 6) return;
    const int r = k / 6, a = k % 6;
    if (a < 3) bx->lo[r][a] = all6[k];
    else bx->hi[r][a - 3] = all6[k];
}

// B2. which peers need which of my locals as ghosts (domain_possibleRemote, src/domain.c:101-124, against the peers' boxes)
__global__ void __launch_bounds__(256)
k_rd_ghostmask(int nLocal, const double4 *__restrict__ pos, DdcGeom g, const DdcBoxes *__restrict__ bxp, uint32_t *__restrict__ mask,
               DdcWork *__restrict__ work)
{
    __shared__ DdcBoxes bx;
    for (int k = threadIdx.x; k < (int)(sizeof(DdcBoxes) / sizeof(double)); k += blockDim.x) ((double *)&bx)[k] = ((const double *)bxp)[k];
    __syncthreads();
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t m = 0u;
    if (s < nLocal)
    {
        const double4 p = pos[s];
        for (int q = 0; q < g.nranks; q++)
            if (q != g.me && ddcNear(p.x, p.y, p.z, q, g, bx)) m |= 1u << q;
        mask[s] = m;
    }
    const int lane = threadIdx.x & 31;
    for (int q = 0; q < g.nranks; q++)
    {
        const unsigned v = __ballot_sync(0xffffffffu, (m >> q) & 1u);
        if (v && lane == (__ffs(v) - 1)) atomicAdd(&work->sendGhost[q], __popc(v));
    }
}

// B4. ghost records {bead, x, y, z} and the send list (bead ids) in the same order
__global__ void k_rd_ghostpack(int nLocal, const double4 *__restrict__ pos, const uint32_t *__restrict__ mask, int nranks,
                               DdcWork *__restrict__ work, DdcOffsets off, int *__restrict__ sendBead, double *__restrict__ buf)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nLocal) return;
    uint32_t m = mask[s];
    if (!m) return;
    const double4 p = pos[s];
    const int bead = (int)((((unsigned long long)__double_as_longlong(p.w)) >> 32) & 0x7fffffffull);
    while (m)
    {
        const int q = __ffs(m) - 1;
        m &= m - 1;
        const int k = off.off[q] + atomicAdd(&work->fillGhost[q], 1);
        sendBead[k] = bead;
        double *o = buf + 4 * (size_t)k;
        o[0] = (double)bead;
        o[1] = p.x; o[2] = p.y; o[3] = p.z;
    }
}

__global__ void k_rd_ghostunpack(int n, const double *__restrict__ buf, int base, const uint64_t *__restrict__ wOfBead,
                                 double4 *__restrict__ pos, double *__restrict__ vx, double *__restrict__ vy, double *__restrict__ vz,
                                 int *__restrict__ beadOfSlot, int *__restrict__ slotOfBead, int *__restrict__ recvBead)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double *o = buf + 4 * (size_t)i;
    const int b = (int)o[0];
    const int k = base + i;
    pos[k] = make_double4(o[1], o[2], o[3], __longlong_as_double((long long)(wOfBead[b] | DDC_GHOST_BIT)));
    vx[k] = 0.0; vy[k] = 0.0; vz[k] = 0.0;      // velocities only exist for local beads
    beadOfSlot[k] = b;
    slotOfBead[b] = k;
    recvBead[i] = b;
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
                              unsigned long long *__restrict__ dmax2, unsigned long long *__restrict__ cellDmax,
                              const int *__restrict__ cellOfSlot)
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
        trackCellDisp(cellDmax, cellOfSlot, s, disp2);      // the ghost part of the bead's cell
    }
    for (int o = 16; o > 0; o >>= 1) disp2 = fmax(disp2, __shfl_xor_sync(0xffffffffu, disp2, o));
    if ((threadIdx.x & 31) == 0)
    {
        const unsigned long long bits = (unsigned long long)__double_as_longlong(disp2);
        if (bits > *(volatile unsigned long long *)(dmax2 + 1)) atomicMax(dmax2 + 1, bits);      // [1]: ghosts (engine.cuh)
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

// pair.cuh - Martini non-bonded pair force: shifted Lennard-Jones + reaction-field Coulomb,
// with the reaction-field-only treatment of pruned (excluded) intramolecular pairs.
//
// Replaces martiniNonBond + martiniIntraMoleReaction (src/bioMartini.c:989-1208) and
// nlistGPU.cu's evalList5.  One thread per bead walks its full (both-direction) list, so
// forces need no atomics and the summation order is fixed: results are bitwise
// reproducible run to run.  Energies and the virial are halved per visit.
//
// fp64 throughout.  The summation order differs from the CPU path's linked-list order, so
// parity is to tolerance (forces 1e-6, energy 1e-9 relative), not bitwise.
#pragma once
#include "engine.cuh"

// One 256-bit load per j-bead: pos4 records are 32-byte aligned, so the whole record is one
// L1/L2 sector and one LSU request (LDG.E.256, sm_100 only) instead of two 128-bit requests.
__device__ __forceinline__ double4 ldPos(const double4 *p)
{
#ifdef DDCB200_EMU
    return *p;
#else
    double4 r;
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p));
    return r;
#endif
}

// A row streams through once: its loads do not allocate in L1, which is left to the gathered positions that neighbouring beads
// re-read (measured: -3 % on the walk; evict-last on the position loads made no difference - profiles/r02o_prune_ab.txt)
__device__ __forceinline__ uint32_t ldRow(const uint32_t *p)
{
#ifdef DDCB200_EMU
    return *p;
#else
    uint32_t v;
    asm("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
#endif
}

#define PF 4   // gathers in flight per thread

template <bool ENERGY>
__global__ void __launch_bounds__(TILE)
k_pair(int nIon, int nPad, const int *__restrict__ tileOrder, int tileBase, const double4 *__restrict__ pos, const uint32_t *__restrict__ nbr,
       const uint16_t *__restrict__ cum, const unsigned long long *__restrict__ dmax2, int withGhosts, const float *__restrict__ dispOfSlot,
       const double2 *__restrict__ ljTab, const double *__restrict__ shiftTab, const double *__restrict__ qTab, PairConst pc,
       double *__restrict__ fx, double *__restrict__ fy, double *__restrict__ fz, double *__restrict__ accPartial)
{
    EXTERN_SHARED(double2, sLJ);               // ntypes*ntypes {c6,c12}
    double *sQ = (double *)(sLJ + pc.ntypes * pc.ntypes);   // 256 charges
    double *sShift = sQ + 256;                              // ntypes*ntypes, ENERGY only
    for (int k = threadIdx.x; k < pc.ntypes * pc.ntypes; k += blockDim.x)
    {
        sLJ[k] = ljTab[k];
        if (ENERGY) sShift[k] = shiftTab[k];
    }
    for (int k = threadIdx.x; k < 256; k += blockDim.x) sQ[k] = qTab[k];
    __syncthreads();

    // several ranks: the launch covers a range of the tile order (rows without / with ghost entries)
    const int tile = tileOrder ? tileOrder[tileBase + blockIdx.x] : (int)blockIdx.x;
    const int i = tile * TILE + threadIdx.x;
    const int ii = i < nIon ? i : 0;
    const double4 pi = ldPos(pos + ii);
    const uint64_t wi = (uint64_t)__double_as_longlong(pi.w);
    const bool live = i < nIon && !(wi >> 63);   // ghost slots (bit 63 of w) own no row and receive no force here
    const int ti = (int)(wi & 0xff);
    const double qi = sQ[(wi >> 8) & 0xff];
    const double kqi = pc.keR * qi;
    const double2 *ljRow = sLJ + ti * pc.ntypes;
    // Rows are ordered by build-time distance bin.  A pair (i, j) listed at distance r_build can only be inside the
    // cutoff now if r_build - d_i - d_j - listSlack < rmax; d_i = this bead's own displacement since the build (rounded
    // up, written by k_integrate / k_nglfc), d_j <= dmax = the largest displacement of any resident bead, listSlack =
    // change of the box edges since the build (0 without a barostat): bins that start beyond that are not even loaded.
    // Exact, not a heuristic - the skipped entries would have added exact zeros, so the forces are bit-for-bit the same.
    int binLimit = 0;
    {
        // dmax2[0]: local beads (complete when this kernel starts), dmax2[1]: ghosts (complete once the halo has arrived, which
        // the launch over the rows with ghost entries waits for; rows without ghost entries only have local partners)
        unsigned long long db = dmax2[0];
        if (withGhosts) db = max(db, dmax2[1]);
        const double dmax = sqrt(__longlong_as_double((long long)db));
        // dispOfSlot == nullptr (DDCB200_WALK=global): every bead takes the global bound, d_i := dmax
        const double di = (live && dispOfSlot) ? fmin((double)dispOfSlot[ii], dmax) : dmax;
        const double lim = (pc.rmax + pc.listSlack + dmax + di) * (1.0 + 1e-12);
#pragma unroll
        for (int e = 0; e < NBINS - 1; e++) binLimit += (pc.binEdge[e] < lim) ? 1 : 0;
    }
    int n = live ? (int)cum[(size_t)binLimit * nPad + ii] : 0;
    int nmax = n;
    for (int o = 16; o > 0; o >>= 1) nmax = max(nmax, __shfl_xor_sync(0xffffffffu, nmax, o));

    double fxi = 0.0, fyi = 0.0, fzi = 0.0;
    double eLJ = 0.0, eEle = 0.0, vxx = 0.0, vyy = 0.0, vzz = 0.0, vxy = 0.0, vxz = 0.0, vyz = 0.0;

    // The walk is latency-bound on the dependent chain entry -> j position (ncu: >60% of stall samples on the first
    // use of the gathered position), so PF gathers are issued back to back before any is consumed, and the entries
    // of the next chunk are already in flight.
    const uint32_t *row = nbr + ii;
    uint32_t eNext[PF];
#pragma unroll
    for (int u = 0; u < PF; u++) eNext[u] = (u < n) ? row[(size_t)u * nPad] : (uint32_t)ii;
    for (int k0 = 0; k0 < nmax; k0 += PF)
    {
        uint32_t eCur[PF];
        double4 pCur[PF];
#pragma unroll
        for (int u = 0; u < PF; u++)
        {
            eCur[u] = eNext[u];
            // lanes past the end of their own row issue no load at all (a dummy gather would still cost an L1 tag lookup)
            pCur[u] = pi;
            if (k0 + u < n) pCur[u] = ldPos(pos + (eCur[u] & 0x07ffffffu));
        }
#pragma unroll
        for (int u = 0; u < PF; u++) eNext[u] = (k0 + PF + u < n) ? row[(size_t)(k0 + PF + u) * nPad] : (uint32_t)ii;
#pragma unroll
        for (int u = 0; u < PF; u++)
        {
        const int k = k0 + u;
        const uint32_t e = eCur[u];
        const double4 pj = pCur[u];
        const bool valid = k < n;
        double x = pi.x - pj.x, y = pi.y - pj.y, z = pi.z - pj.z;
        double r2 = x * x + y * y + z * z;
        if (r2 > pc.R2cut)
        {
            // nearestImage_fast: one lattice reduction per component (src/preduce.c:147-160)
            if (x > pc.hhx) x -= pc.hxx;
            if (x < -pc.hhx) x += pc.hxx;
            if (y > pc.hhy) y -= pc.hyy;
            if (y < -pc.hhy) y += pc.hyy;
            if (z > pc.hhz) z -= pc.hzz;
            if (z < -pc.hhz) z += pc.hzz;
            r2 = x * x + y * y + z * z;
        }
        const bool in = valid && (r2 < pc.rc2);
        if (__any_sync(0xffffffffu, in))
        {
            const uint64_t wj = (uint64_t)__double_as_longlong(pj.w);
            const bool excl = (e & EXCL_BIT) != 0u;
            const double kqij = kqi * sQ[(wj >> 8) & 0xff];
            const double r2s = in ? r2 : 1.0;
            double dvdr, vlj = 0.0, vele = 0.0;
            // one reciprocal square root serves both terms (the reference takes sqrt(1/r2), src/bioMartini.c:1068)
            const double ir1 = rsqrt(r2s);
            const double ir2 = ir1 * ir1;
            {
                // Lennard-Jones: 4 eps (s12 - s6) + shift ; dvdr = 24 eps (s6 - 2 s12)/r^2 (src/bioMartini.c:1073-1080)
                const double2 cc = ljRow[wj & 0xff];
                const double ir6 = ir2 * ir2 * ir2;
                const double a6 = excl ? 0.0 : cc.x * ir6;
                const double a12 = excl ? 0.0 : cc.y * ir6 * ir6;
                dvdr = 6.0 * (a6 - 2.0 * a12) * ir2;
                if (ENERGY) vlj = excl ? 0.0 : (a12 - a6) + sShift[ti * pc.ntypes + (int)(wj & 0xff)];
            }
            if (__any_sync(0xffffffffu, in && kqij != 0.0))
            {
                // reaction field (src/bioMartini.c:1082-1085); pruned pairs keep only krf r^2 - crf (:1172-1174)
                const double ir = excl ? 0.0 : ir1;
                dvdr += kqij * (2.0 * pc.krf - ir2 * ir);
                if (ENERGY) vele = kqij * (ir + pc.krf * r2s - pc.crf);
            }
            if (!in)
            {
                dvdr = 0.0;
                vlj = 0.0;
                vele = 0.0;
            }
            const double fxij = -dvdr * x, fyij = -dvdr * y, fzij = -dvdr * z;
            fxi += fxij;
            fyi += fyij;
            fzi += fzij;
            if (ENERGY)
            {
                eLJ += vlj;
                eEle += vele;
                vxx += fxij * x;
                vyy += fyij * y;
                vzz += fzij * z;
                vxy += fxij * y;
                vxz += fxij * z;
                vyz += fyij * z;
            }
        }
        }
    }
    if (live)
    {
        fx[i] = fxi;
        fy[i] = fyi;
        fz[i] = fzi;
    }
    if (ENERGY)
    {
        // every pair is visited from both ends: halve.  Self term -0.5 q_i^2 keR crf (src/bioMartini.c:1031-1035)
        double v[8] = {0.5 * eLJ, 0.5 * eEle + (live ? -0.5 * qi * qi * pc.keR * pc.crf : 0.0),
                       0.5 * vxx, 0.5 * vyy, 0.5 * vzz, 0.5 * vxy, 0.5 * vxz, 0.5 * vyz};
        __shared__ double red[8][TILE / 32];
#pragma unroll
        for (int a = 0; a < 8; a++)
        {
            double t = v[a];
            for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
            if ((threadIdx.x & 31) == 0) red[a][threadIdx.x >> 5] = t;
        }
        __syncthreads();
        if (threadIdx.x < 8)
        {
            double t = 0.0;
            for (int w = 0; w < TILE / 32; w++) t += red[threadIdx.x][w];
            accPartial[(size_t)tile * 8 + threadIdx.x] = t;
        }
    }
}


// ---- k_pair2: the same walk with fewer instructions per entry -------------------------------------------------------------
// ncu on k_pair (profiles/r02a_k_pair_ncu_full.txt): 68 warp instructions per walked entry of which 18 are FP64, issue slots 50 %
// busy with 4.7 warps per scheduler (83 registers), stalls split between the L1 scoreboard and fixed-latency FP64 chains.  So the
// walk is bound by instruction issue and latency, not by a pipe.  Here:
//   * the in-cutoff work sits behind a real branch (lanes outside the cutoff skip it) instead of warp votes + 64-bit selects;
//   * the Coulomb part is only entered for charged i beads (most Martini beads carry no charge);
//   * the reciprocal square root is the hardware approximation + two Newton steps (no special-case handling: r2 is a finite
//     positive distance), the LJ table holds 6 c6 and 12 c12 so the force needs no extra factor;
//   * PF (gathers in flight per thread) and the register cap (MINB CTAs per SM) are template parameters, A/B-ed on the B200.
__device__ __forceinline__ double rsqrtFast(double x)
{
#ifdef DDCB200_EMU
    return 1.0 / sqrt(x);
#else
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));      // MUFU.RSQ64H: about 2^-22 relative
    const double h = 0.5 * x;
    y = y * fma(-h * y, y, 1.5);
    y = y * fma(-h * y, y, 1.5);
    return y;
#endif
}

// The pruned rows (MODE): the list radius is rmax + deltaR for a rebuild every updateRate steps, but over a few steps the beads move
// a fraction of deltaR only.  Every pruneEvery steps the walk therefore also WRITES a second, shorter row per bead (MODE 1): the
// entries now closer than rmax + margin, in the order of the full row, and the positions of that step become the reference of the
// displacement bounds (k_rebase).  On the steps in between (MODE 2) a bead walks its pruned row as long as its own displacement
// plus the largest in its stencil cells since the prune stays within the margin - every skipped entry is then provably outside
// the cutoff and would have added an exact zero, so forces and energies are bitwise those of the full walk - and its full row
// otherwise.  The same idea as the rolling pruning of a dual pair list; here it is exact by construction, not by a drift estimate.
// The distance (at the time the reference positions were taken: the build, or the last prune) beyond which an entry of bead ii's
// row cannot be inside the cutoff now: rmax + its own displacement + the largest displacement of a possible partner.
__device__ __forceinline__ double pairWalkLim(int ii, bool live, const unsigned long long *__restrict__ nbrDmax, const int *__restrict__ cellOfSlot,
                                              const unsigned long long *__restrict__ dmax2, int withGhosts, const float *__restrict__ dispOfSlot,
                                              const PairConst &pc)
{
    // d_j: the largest displacement in this bead's stencil cells (k_nbr_dmax; the launch over the rows with ghost entries
    // gets the table that includes the ghosts), or - DDCB200_WALK=bead / global - of any resident bead
    unsigned long long db;
    if (nbrDmax) db = nbrDmax[cellOfSlot[ii]];
    else
    {
        db = dmax2[0];
        if (withGhosts) db = max(db, dmax2[1]);
    }
    const double dmax = sqrt(__longlong_as_double((long long)db));
    double di = dmax;
    if (live && dispOfSlot)
    {
        // this bead's own displacement; never more than the global maximum, which the DDCB200_WALK=bead bound is capped with
        unsigned long long dg = dmax2[0];
        if (withGhosts) dg = max(dg, dmax2[1]);
        di = fmin((double)dispOfSlot[ii], sqrt(__longlong_as_double((long long)dg)));
    }
    return (pc.rmax + pc.listSlack + dmax + di) * (1.0 + 1e-12);
}

struct PruneArgs
{
    uint32_t *rows;      // pruned rows, transposed like the full rows (the candidate buffer of the list build, idle between builds)
    uint16_t *count;     // entries per pruned row
    double keep2;        // MODE 1: (rmax + margin)^2
    double walkLim;      // MODE 1: the full row is walked up to this build-time distance (rmax + margin right after a build, else all)
    double useLim;       // MODE 2: a bead may use its pruned row while rmax + its displacement bound <= rmax + margin
};

template <bool ENERGY, int NPF, int MINB, int MODE>
__global__ void __launch_bounds__(TILE, MINB)
k_pair2(int nIon, int nPad, const int *__restrict__ tileOrder, int tileBase, const double4 *__restrict__ pos, const uint32_t *__restrict__ nbr,
        const uint16_t *__restrict__ cum, const unsigned long long *__restrict__ dmax2, int withGhosts, const float *__restrict__ dispOfSlot,
        const double2 *__restrict__ ljTab, const double *__restrict__ shiftTab, const double *__restrict__ qTab, PairConst pc,
        double *__restrict__ fx, double *__restrict__ fy, double *__restrict__ fz, double *__restrict__ accPartial,
        const unsigned long long *__restrict__ nbrDmax, const int *__restrict__ cellOfSlot, PruneArgs pr)
{
    EXTERN_SHARED(double2, sLJ);               // ntypes*ntypes {6 c6, 12 c12}
    double *sQ = (double *)(sLJ + pc.ntypes * pc.ntypes);   // 256 charges
    double *sShift = sQ + 256;                              // ntypes*ntypes, ENERGY only
    for (int k = threadIdx.x; k < pc.ntypes * pc.ntypes; k += blockDim.x)
    {
        const double2 c = ljTab[k];
        sLJ[k] = make_double2(6.0 * c.x, 12.0 * c.y);
        if (ENERGY) sShift[k] = shiftTab[k];
    }
    for (int k = threadIdx.x; k < 256; k += blockDim.x) sQ[k] = qTab[k];
    __syncthreads();

    const int tile = tileOrder ? tileOrder[tileBase + blockIdx.x] : (int)blockIdx.x;
    const int i = tile * TILE + threadIdx.x;
    const int ii = i < nIon ? i : 0;
    const double4 pi = ldPos(pos + ii);
    const uint64_t wi = (uint64_t)__double_as_longlong(pi.w);
    const bool live = i < nIon && !(wi >> 63);
    const int ti = (int)(wi & 0xff);
    const double qi = sQ[(wi >> 8) & 0xff];
    const double kqi = pc.keR * qi;
    const bool charged = kqi != 0.0;
    const double2 *ljRow = sLJ + ti * pc.ntypes;
    int binLimit = 0;
    bool usePruned = false;
    {
        double lim = pairWalkLim(ii, live, nbrDmax, cellOfSlot, dmax2, withGhosts, dispOfSlot, pc);
        if (MODE == 1) lim = pr.walkLim;
        if (MODE == 2)
        {
            usePruned = lim <= pr.useLim;
            lim = 1e300;      // the displacements are those since the prune, not since the build: a full row is walked to its end
        }
#pragma unroll
        for (int e = 0; e < NBINS - 1; e++) binLimit += (pc.binEdge[e] < lim) ? 1 : 0;
    }
    int n = 0;
    if (live) n = (MODE == 2 && usePruned) ? (int)pr.count[ii] : (int)cum[(size_t)binLimit * nPad + ii];
    int nmax = n;
    for (int o = 16; o > 0; o >>= 1) nmax = max(nmax, __shfl_xor_sync(0xffffffffu, nmax, o));

    double fxi = 0.0, fyi = 0.0, fzi = 0.0;
    double eLJ = 0.0, eEle = 0.0, vxx = 0.0, vyy = 0.0, vzz = 0.0, vxy = 0.0, vxz = 0.0, vyz = 0.0;
    const double twoKrf = 2.0 * pc.krf;

    const uint32_t *row = ((MODE == 2 && usePruned) ? pr.rows : nbr) + ii;
    int nKept = 0;      // MODE 1: entries written to the pruned row so far
    uint32_t eNext[NPF];
#pragma unroll
    for (int u = 0; u < NPF; u++) eNext[u] = (u < n) ? ldRow(row + (size_t)u * nPad) : 0u;
    for (int k0 = 0; k0 < nmax; k0 += NPF)
    {
        uint32_t eCur[NPF];
        double4 pCur[NPF];      // only read where the lane still has an entry (k0 + u < n)
#pragma unroll
        for (int u = 0; u < NPF; u++)
        {
            eCur[u] = eNext[u];
            // lanes past the end of their own row issue no load at all (a dummy gather would still cost an L1 tag lookup)
            if (k0 + u < n) pCur[u] = ldPos(pos + (eCur[u] & 0x07ffffffu));
        }
#pragma unroll
        for (int u = 0; u < NPF; u++) eNext[u] = (k0 + NPF + u < n) ? ldRow(row + (size_t)(k0 + NPF + u) * nPad) : 0u;
#pragma unroll
        for (int u = 0; u < NPF; u++)
        {
            const bool have = k0 + u < n;
            const double4 pj = pCur[u];
            double x = pi.x - pj.x, y = pi.y - pj.y, z = pi.z - pj.z;
            double r2 = x * x + y * y + z * z;
            if (have && r2 > pc.R2cut)
            {
                // nearestImage_fast: one lattice reduction per component (src/preduce.c:147-160)
                if (x > pc.hhx) x -= pc.hxx;
                if (x < -pc.hhx) x += pc.hxx;
                if (y > pc.hhy) y -= pc.hyy;
                if (y < -pc.hhy) y += pc.hyy;
                if (z > pc.hhz) z -= pc.hzz;
                if (z < -pc.hhz) z += pc.hzz;
                r2 = x * x + y * y + z * z;
            }
            if (MODE == 1 && have && r2 < pr.keep2)
            {
                pr.rows[(size_t)nKept * nPad + ii] = eCur[u];
                nKept++;
            }
            if (have && r2 < pc.rc2)
            {
                const uint64_t wj = (uint64_t)__double_as_longlong(pj.w);
                const bool excl = (eCur[u] & EXCL_BIT) != 0u;
                const double ir1 = rsqrtFast(r2);
                const double ir2 = ir1 * ir1;
                double dvdr = 0.0;
                if (!excl)
                {
                    // Lennard-Jones: 4 eps (s12 - s6) + shift ; dvdr = 24 eps (s6 - 2 s12)/r^2 (src/bioMartini.c:1073-1080)
                    const double2 cc = ljRow[wj & 0xff];
                    const double ir6 = ir2 * ir2 * ir2;
                    const double a6 = cc.x * ir6, a12 = cc.y * ir6 * ir6;
                    dvdr = (a6 - a12) * ir2;
                    if (ENERGY) eLJ += (a12 * (1.0 / 12.0) - a6 * (1.0 / 6.0)) + sShift[ti * pc.ntypes + (int)(wj & 0xff)];
                }
                if (charged)
                {
                    const double kqij = kqi * sQ[(wj >> 8) & 0xff];
                    // reaction field (src/bioMartini.c:1082-1085); pruned pairs keep only krf r^2 - crf (:1172-1174)
                    const double ir = excl ? 0.0 : ir1;
                    dvdr += kqij * (twoKrf - ir2 * ir);
                    if (ENERGY) eEle += kqij * (ir + pc.krf * r2 - pc.crf);
                }
                const double fxij = -dvdr * x, fyij = -dvdr * y, fzij = -dvdr * z;
                fxi += fxij;
                fyi += fyij;
                fzi += fzij;
                if (ENERGY)
                {
                    vxx += fxij * x;
                    vyy += fyij * y;
                    vzz += fzij * z;
                    vxy += fxij * y;
                    vxz += fxij * z;
                    vyz += fyij * z;
                }
            }
        }
    }
    if (live)
    {
        fx[i] = fxi;
        fy[i] = fyi;
        fz[i] = fzi;
        if (MODE == 1) pr.count[i] = (uint16_t)nKept;
    }
    if (ENERGY)
    {
        double v[8] = {0.5 * eLJ, 0.5 * eEle + (live ? -0.5 * qi * qi * pc.keR * pc.crf : 0.0),
                       0.5 * vxx, 0.5 * vyy, 0.5 * vzz, 0.5 * vxy, 0.5 * vxz, 0.5 * vyz};
        __shared__ double red[8][TILE / 32];
#pragma unroll
        for (int a = 0; a < 8; a++)
        {
            double t = v[a];
            for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
            if ((threadIdx.x & 31) == 0) red[a][threadIdx.x >> 5] = t;
        }
        __syncthreads();
        if (threadIdx.x < 8)
        {
            double t = 0.0;
            for (int w = 0; w < TILE / 32; w++) t += red[threadIdx.x][w];
            accPartial[(size_t)tile * 8 + threadIdx.x] = t;
        }
    }
}


// what the walk of the next force evaluation would visit (ddcb200_pruneInfo; a measurement hook, not part of a step):
// out[0] = entries of the rows the beads would walk, out[1] = beads that would walk their pruned row, out[2] = entries of all pruned rows
__global__ void __launch_bounds__(TILE)
k_prune_stats(int nIon, int nPad, const double4 *__restrict__ pos, const uint16_t *__restrict__ cum, const unsigned long long *__restrict__ dmax2,
              int withGhosts, const float *__restrict__ dispOfSlot, PairConst pc, const unsigned long long *__restrict__ nbrDmax,
              const int *__restrict__ cellOfSlot, PruneArgs pr, unsigned long long *__restrict__ out)
{
    const int i = blockIdx.x * TILE + threadIdx.x;
    unsigned long long walked = 0ull, pruned = 0ull, kept = 0ull;
    if (i < nIon && !(((uint64_t)__double_as_longlong(pos[i].w)) >> 63))
    {
        const double lim = pairWalkLim(i, true, nbrDmax, cellOfSlot, dmax2, withGhosts, dispOfSlot, pc);
        const bool usePruned = lim <= pr.useLim;
        walked = usePruned ? pr.count[i] : cum[(size_t)(NBINS - 1) * nPad + i];
        pruned = usePruned ? 1ull : 0ull;
        kept = pr.count[i];
    }
    for (int o = 16; o > 0; o >>= 1)
    {
        walked += __shfl_xor_sync(0xffffffffu, walked, o);
        pruned += __shfl_xor_sync(0xffffffffu, pruned, o);
        kept += __shfl_xor_sync(0xffffffffu, kept, o);
    }
    if ((threadIdx.x & 31) == 0)
    {
        atomicAdd(out, walked);
        atomicAdd(out + 1, pruned);
        atomicAdd(out + 2, kept);
    }
}


// ---- k_pair3: the tile's partners staged in shared memory --------------------------------------------------------------------
// ncu on k_pair2 mid-cycle (profiles/r02d_k_pair2_ncu_full.txt): l1tex data-pipe wavefronts 83 % of peak - every gathered partner
// is one 32-byte sector of its own, one wavefront each, whatever the hit rate.  Here the partners of a tile's rows come from
// shared memory instead: the tile's WINDOW (the slot runs of the stencil cells of the tile's cells, k_tile_window) is copied in
// with coalesced loads, split into {x, y} and {z, w} arrays so that a warp-wide gather of 16-byte halves spreads over all 32
// banks, and the rows hold window offsets.  Two threads per bead walk the even and the odd entries of the row (twice the
// warps for the same window) and are added with one shuffle, in a fixed order.  Tiles whose window did not fit keep slot
// entries and gather from global memory as k_pair2 does.
struct PairAcc
{
    double fx, fy, fz, eLJ, eEle, vxx, vyy, vzz, vxy, vxz, vyz;
};

template <bool ENERGY, bool WIN>
__device__ __forceinline__ void pairWalk3(PairAcc &A, const double4 pi, int ti, double kqi, int n, int nmax, int h, const uint32_t *__restrict__ row,
                                          int nPad, const double4 *__restrict__ pos, const double2 *__restrict__ sA, const double2 *__restrict__ sB,
                                          const double2 *__restrict__ ljRow, const double *__restrict__ sQ, const double *__restrict__ sShiftRow,
                                          const PairConst &pc)
{
    const bool charged = kqi != 0.0;
    const double twoKrf = 2.0 * pc.krf;
    // this thread's entries: k = h, h + 2, ...: two per trip, and the entries of the next trip are already in flight (the rows
    // stream from HBM with no reuse; without the prefetch the walk waits a memory latency per trip)
    uint32_t eNext[2];
#pragma unroll
    for (int u = 0; u < 2; u++) eNext[u] = (h + 2 * u < n) ? row[(size_t)(h + 2 * u) * nPad] : 0u;
    for (int k0 = h; k0 < nmax; k0 += 4)
    {
        uint32_t e[2];
        double4 pj[2];
#pragma unroll
        for (int u = 0; u < 2; u++)
        {
            e[u] = eNext[u];
            if (k0 + 2 * u < n)
            {
                const uint32_t idx = e[u] & 0x07ffffffu;
                if (WIN)
                {
                    const double2 a = sA[idx], b = sB[idx];
                    pj[u] = make_double4(a.x, a.y, b.x, b.y);
                }
                else pj[u] = ldPos(pos + idx);
            }
        }
#pragma unroll
        for (int u = 0; u < 2; u++) eNext[u] = (k0 + 4 + 2 * u < n) ? row[(size_t)(k0 + 4 + 2 * u) * nPad] : 0u;
#pragma unroll
        for (int u = 0; u < 2; u++)
        {
            const bool have = k0 + 2 * u < n;
            double x = pi.x - pj[u].x, y = pi.y - pj[u].y, z = pi.z - pj[u].z;
            double r2 = x * x + y * y + z * z;
            if (have && r2 > pc.R2cut)
            {
                // nearestImage_fast: one lattice reduction per component (src/preduce.c:147-160)
                if (x > pc.hhx) x -= pc.hxx;
                if (x < -pc.hhx) x += pc.hxx;
                if (y > pc.hhy) y -= pc.hyy;
                if (y < -pc.hhy) y += pc.hyy;
                if (z > pc.hhz) z -= pc.hzz;
                if (z < -pc.hhz) z += pc.hzz;
                r2 = x * x + y * y + z * z;
            }
            if (have && r2 < pc.rc2)
            {
                const uint64_t wj = (uint64_t)__double_as_longlong(pj[u].w);
                const bool excl = (e[u] & EXCL_BIT) != 0u;
                const double ir1 = rsqrtFast(r2);
                const double ir2 = ir1 * ir1;
                double dvdr = 0.0;
                if (!excl)
                {
                    const double2 cc = ljRow[wj & 0xff];
                    const double ir6 = ir2 * ir2 * ir2;
                    const double a6 = cc.x * ir6, a12 = cc.y * ir6 * ir6;
                    dvdr = (a6 - a12) * ir2;
                    if (ENERGY) A.eLJ += (a12 * (1.0 / 12.0) - a6 * (1.0 / 6.0)) + sShiftRow[wj & 0xff];
                }
                if (charged)
                {
                    const double kqij = kqi * sQ[(wj >> 8) & 0xff];
                    const double ir = excl ? 0.0 : ir1;
                    dvdr += kqij * (twoKrf - ir2 * ir);
                    if (ENERGY) A.eEle += kqij * (ir + pc.krf * r2 - pc.crf);
                }
                const double fxij = -dvdr * x, fyij = -dvdr * y, fzij = -dvdr * z;
                A.fx += fxij;
                A.fy += fyij;
                A.fz += fzij;
                if (ENERGY)
                {
                    A.vxx += fxij * x;
                    A.vyy += fyij * y;
                    A.vzz += fzij * z;
                    A.vxy += fxij * y;
                    A.vxz += fxij * z;
                    A.vyz += fyij * z;
                }
            }
        }
    }
    (void)ti;
}

#define PAIR3_THREADS (2 * TILE)
template <bool ENERGY>
__global__ void __launch_bounds__(PAIR3_THREADS, 2)
k_pair3(int nIon, int nPad, const int *__restrict__ tileOrder, int tileBase, const double4 *__restrict__ pos, const uint32_t *__restrict__ nbr,
        const uint16_t *__restrict__ cum, const unsigned long long *__restrict__ dmax2, int withGhosts, const float *__restrict__ dispOfSlot,
        const TileWin *__restrict__ tileWin, int wcap, const double2 *__restrict__ ljTab, const double *__restrict__ shiftTab,
        const double *__restrict__ qTab, PairConst pc, double *__restrict__ fx, double *__restrict__ fy, double *__restrict__ fz,
        double *__restrict__ accPartial)
{
    EXTERN_SHARED(double2, sLJ);               // ntypes*ntypes {6 c6, 12 c12}
    const int nt2 = pc.ntypes * pc.ntypes;
    double *sQ = (double *)(sLJ + nt2);        // 256 charges
    double *sShift = sQ + 256;                 // ntypes*ntypes (read by the ENERGY instantiation only)
    double2 *sA = (double2 *)(sShift + nt2 + (nt2 & 1));      // window {x, y}, 16-byte aligned
    double2 *sB = sA + wcap;                                  // window {z, w}
    __shared__ TileWin sWin;
    const int tile = tileOrder ? tileOrder[tileBase + blockIdx.x] : (int)blockIdx.x;
    for (int k = threadIdx.x; k < (int)(sizeof(TileWin) / sizeof(int)); k += blockDim.x) ((int *)&sWin)[k] = ((const int *)(tileWin + tile))[k];
    for (int k = threadIdx.x; k < nt2; k += blockDim.x)
    {
        const double2 c = ljTab[k];
        sLJ[k] = make_double2(6.0 * c.x, 12.0 * c.y);
        sShift[k] = shiftTab[k];
    }
    for (int k = threadIdx.x; k < 256; k += blockDim.x) sQ[k] = qTab[k];
    __syncthreads();
    const bool windowed = sWin.nRuns > 0;
    if (windowed)
    {
        for (int r = 0; r < sWin.nRuns; r++)
        {
            const int lo = sWin.lo[r], o = sWin.off[r], cnt = sWin.off[r + 1] - o;
            // four independent loads per thread in flight (plain loads: the compiler may batch them)
            const double2 *src = (const double2 *)(pos + lo);
#pragma unroll 4
            for (int q = threadIdx.x; q < cnt; q += blockDim.x)
            {
                const double2 a = src[2 * q], b = src[2 * q + 1];
                sA[o + q] = a;
                sB[o + q] = b;
            }
        }
        __syncthreads();
    }

    const int h = threadIdx.x & 1;
    const int i = tile * TILE + (threadIdx.x >> 1);
    const int ii = i < nIon ? i : 0;
    const double4 pi = ldPos(pos + ii);
    const uint64_t wi = (uint64_t)__double_as_longlong(pi.w);
    const bool live = i < nIon && !(wi >> 63);
    const int ti = (int)(wi & 0xff);
    const double qi = sQ[(wi >> 8) & 0xff];
    const double kqi = pc.keR * qi;
    int binLimit = 0;
    {
        unsigned long long db = dmax2[0];
        if (withGhosts) db = max(db, dmax2[1]);
        const double dmax = sqrt(__longlong_as_double((long long)db));
        const double di = (live && dispOfSlot) ? fmin((double)dispOfSlot[ii], dmax) : dmax;
        const double lim = (pc.rmax + pc.listSlack + dmax + di) * (1.0 + 1e-12);
#pragma unroll
        for (int e = 0; e < NBINS - 1; e++) binLimit += (pc.binEdge[e] < lim) ? 1 : 0;
    }
    const int n = live ? (int)cum[(size_t)binLimit * nPad + ii] : 0;
    int nmax = n;
    for (int o = 16; o > 0; o >>= 1) nmax = max(nmax, __shfl_xor_sync(0xffffffffu, nmax, o));

    PairAcc A = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    if (windowed)
        pairWalk3<ENERGY, true>(A, pi, ti, kqi, n, nmax, h, nbr + ii, nPad, pos, sA, sB, sLJ + ti * pc.ntypes, sQ, sShift + ti * pc.ntypes, pc);
    else
        pairWalk3<ENERGY, false>(A, pi, ti, kqi, n, nmax, h, nbr + ii, nPad, pos, sA, sB, sLJ + ti * pc.ntypes, sQ, sShift + ti * pc.ntypes, pc);
    // the two halves of a row, added in a fixed order: even entries + odd entries
    A.fx += __shfl_xor_sync(0xffffffffu, A.fx, 1);
    A.fy += __shfl_xor_sync(0xffffffffu, A.fy, 1);
    A.fz += __shfl_xor_sync(0xffffffffu, A.fz, 1);
    if (live && h == 0)
    {
        fx[i] = A.fx;
        fy[i] = A.fy;
        fz[i] = A.fz;
    }
    if (ENERGY)
    {
        // every pair is visited from both ends: halve.  Self term -0.5 q_i^2 keR crf (src/bioMartini.c:1031-1035), once per bead
        double v[8] = {0.5 * A.eLJ, 0.5 * A.eEle + ((live && h == 0) ? -0.5 * qi * qi * pc.keR * pc.crf : 0.0),
                       0.5 * A.vxx, 0.5 * A.vyy, 0.5 * A.vzz, 0.5 * A.vxy, 0.5 * A.vxz, 0.5 * A.vyz};
        __shared__ double red[8][PAIR3_THREADS / 32];
#pragma unroll
        for (int a = 0; a < 8; a++)
        {
            double t = v[a];
            for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
            if ((threadIdx.x & 31) == 0) red[a][threadIdx.x >> 5] = t;
        }
        __syncthreads();
        if (threadIdx.x < 8)
        {
            double t = 0.0;
            for (int w = 0; w < PAIR3_THREADS / 32; w++) t += red[threadIdx.x][w];
            accPartial[(size_t)tile * 8 + threadIdx.x] = t;
        }
    }
}

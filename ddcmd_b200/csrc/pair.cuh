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

// ---- k_pair2 -------------------------------------------------------------------------------------------------------------------
// The round-1 kernel k_pair (votes + selects around the in-cutoff work, 83 registers; removed, see the history before round 2's
// last commits) ran 68 warp instructions per walked entry of which 18 FP64, issue slots 50 % busy with 4.7 warps per scheduler
// (profiles/r02a_k_pair_ncu_full.txt).  The shared-memory window variant k_pair3 (north_star's tiles) was built, measured slower
// and removed too (profiles/r02h_k_pair3_ncu_full.txt, r02h_variants.txt).  Here:
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
    const double h = __dmul_rn(0.5, x);
    y = __dmul_rn(y, __fma_rn(__dmul_rn(-h, y), y, 1.5));
    y = __dmul_rn(y, __fma_rn(__dmul_rn(-h, y), y, 1.5));
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
    int farTop;          // full rows in two segments (k_nbr_exact2): entry k >= nNear of a row sits at farTop - (k - nNear); -1: rows run forward
};

// The bonded forces of the step (k_bonded has staged them, one entry per term and endpoint) are added up per bead, in the bead's
// fixed entry order, by the thread that owns the bead's pair force, and the total goes out in one store: no pass of its own over
// the force arrays.  start == nullptr: no bonded terms.
struct BondAdd
{
    const int *start, *count, *stageIdx;     // per local slot: its run of entries; per entry: where the force on this bead is staged (-1: none)
    const double *stage;                     // 3 doubles per (term, endpoint)
};

template <bool ENERGY, int NPF, int MINB, int MODE>
__global__ void __launch_bounds__(TILE, MINB)
k_pair2(int nIon, int nPad, const int *__restrict__ tileOrder, int tileBase, const double4 *__restrict__ pos, const uint32_t *__restrict__ nbr,
        const uint16_t *__restrict__ cum, const unsigned long long *__restrict__ dmax2, int withGhosts, const float *__restrict__ dispOfSlot,
        const double2 *__restrict__ ljTab, const double *__restrict__ shiftTab, const double *__restrict__ qTab, PairConst pc,
        double *__restrict__ fx, double *__restrict__ fy, double *__restrict__ fz, double *__restrict__ accPartial,
        const unsigned long long *__restrict__ nbrDmax, const int *__restrict__ cellOfSlot, PruneArgs pr, BondAdd ba)
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
    // a full row of the one-pass build is two segments: the entries listed closer than the first bin edge from the front, the others
    // from the end of the row's allocation backwards (cum[0] = the length of the first segment)
    int nFront = 0x7fffffff, farBase = 0;
    if (pr.farTop >= 0 && !(MODE == 2 && usePruned) && binLimit > 0 && live)
    {
        nFront = (int)cum[ii];
        farBase = pr.farTop + nFront;
    }
#define ROWAT(k) ldRow(row + (size_t)((k) < nFront ? (k) : farBase - (k)) * nPad)
    int nKept = 0;      // MODE 1: entries written to the pruned row so far
    uint32_t eNext[NPF];
#pragma unroll
    for (int u = 0; u < NPF; u++) eNext[u] = (u < n) ? ROWAT(u) : 0u;
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
        for (int u = 0; u < NPF; u++) eNext[u] = (k0 + NPF + u < n) ? ROWAT(k0 + NPF + u) : 0u;
#pragma unroll
        for (int u = 0; u < NPF; u++)
        {
            const bool have = k0 + u < n;
            const double4 pj = pCur[u];
            // The arithmetic of a pair is spelled out (no contraction left to the compiler): the instantiations of this kernel - with
            // and without energies, over full and pruned rows - must give a pair the same force bit for bit, whichever of them meets it
            double x = pi.x - pj.x, y = pi.y - pj.y, z = pi.z - pj.z;
            double r2 = __fma_rn(z, z, __fma_rn(y, y, __dmul_rn(x, x)));
            if (have && r2 > pc.R2cut)
            {
                // nearestImage_fast: one lattice reduction per component (src/preduce.c:147-160)
                if (x > pc.hhx) x -= pc.hxx;
                if (x < -pc.hhx) x += pc.hxx;
                if (y > pc.hhy) y -= pc.hyy;
                if (y < -pc.hhy) y += pc.hyy;
                if (z > pc.hhz) z -= pc.hzz;
                if (z < -pc.hhz) z += pc.hzz;
                r2 = __fma_rn(z, z, __fma_rn(y, y, __dmul_rn(x, x)));
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
                const double ir2 = __dmul_rn(ir1, ir1);
                double dvdr = 0.0;
                if (!excl)
                {
                    // Lennard-Jones: 4 eps (s12 - s6) + shift ; dvdr = 24 eps (s6 - 2 s12)/r^2 (src/bioMartini.c:1073-1080)
                    const double2 cc = ljRow[wj & 0xff];
                    const double ir6 = __dmul_rn(__dmul_rn(ir2, ir2), ir2);
                    const double a6 = __dmul_rn(cc.x, ir6), a12 = __dmul_rn(__dmul_rn(cc.y, ir6), ir6);
                    dvdr = __dmul_rn(__dadd_rn(a6, -a12), ir2);
                    if (ENERGY) eLJ += (a12 * (1.0 / 12.0) - a6 * (1.0 / 6.0)) + sShift[ti * pc.ntypes + (int)(wj & 0xff)];
                }
                if (charged)
                {
                    const double kqij = __dmul_rn(kqi, sQ[(wj >> 8) & 0xff]);
                    // reaction field (src/bioMartini.c:1082-1085); pruned pairs keep only krf r^2 - crf (:1172-1174)
                    const double ir = excl ? 0.0 : ir1;
                    dvdr = __fma_rn(kqij, __fma_rn(-ir2, ir, twoKrf), dvdr);
                    if (ENERGY) eEle += kqij * (ir + pc.krf * r2 - pc.crf);
                }
                const double ndv = -dvdr;
                fxi = __fma_rn(ndv, x, fxi);
                fyi = __fma_rn(ndv, y, fyi);
                fzi = __fma_rn(ndv, z, fzi);
                if (ENERGY)
                {
                    const double fxij = __dmul_rn(ndv, x), fyij = __dmul_rn(ndv, y), fzij = __dmul_rn(ndv, z);
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
#undef ROWAT
    if (live)
    {
        double bfx = 0.0, bfy = 0.0, bfz = 0.0;
        if (ba.start)
        {
            const int nb = ba.count[i], lo = ba.start[i];
            for (int q = 0; q < nb; q++)
            {
                const int k = ba.stageIdx[lo + q];
                if (k < 0) continue;
                bfx += ba.stage[3 * (size_t)k];
                bfy += ba.stage[3 * (size_t)k + 1];
                bfz += ba.stage[3 * (size_t)k + 2];
            }
        }
        fx[i] = fxi + bfx;
        fy[i] = fyi + bfy;
        fz[i] = fzi + bfz;
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

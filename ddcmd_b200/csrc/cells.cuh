// cells.cuh - cell assignment, cell sort and neighbor-list build.
//
// Replaces GeomBox / geomMethodDefault / GeomDenseBox (src/geom.c:311-454,537-619),
// pairlist1 (src/pairlist.c:205-314) and reOrgPairs (src/bioMartini.c:1392-1485) of the
// reference CPU path, and nlistGPU.cu's buildList.  Cell index and list membership are
// bit-exact with the CPU path: every floating-point operation that feeds an integer
// decision is written with explicit round-to-nearest intrinsics (no FMA contraction),
// in the reference's operation order.
#pragma once
#include "engine.cuh"

// ---- exact helpers -------------------------------------------------------------------
__device__ __forceinline__ void wrapOnce(double &x, double &y, double &z, const BoxConst &b)
{
    // PreduceOrthorhombicB7_OneLatticeReduction, src/preduce.c:147-160
    if (x > b.hhx) x = __dadd_rn(x, -b.hxx);
    if (x < -b.hhx) x = __dadd_rn(x, b.hxx);
    if (y > b.hhy) y = __dadd_rn(y, -b.hyy);
    if (y < -b.hhy) y = __dadd_rn(y, b.hyy);
    if (z > b.hhz) z = __dadd_rn(z, -b.hzz);
    if (z < -b.hhz) z = __dadd_rn(z, b.hzz);
}

__device__ __forceinline__ void normCoord(const double4 p, const BoxConst &b, double &ux, double &uy, double &uz)
{
    // GeomBox first loop, src/geom.c:335-343: r - center, backInBox_fast, hinv * r
    double x = __dadd_rn(p.x, -b.cx), y = __dadd_rn(p.y, -b.cy), z = __dadd_rn(p.z, -b.cz);
    wrapOnce(x, y, z, b);
    ux = __dadd_rn(__dadd_rn(__dmul_rn(b.hinv[0], x), __dmul_rn(b.hinv[1], y)), __dmul_rn(b.hinv[2], z));
    uy = __dadd_rn(__dadd_rn(__dmul_rn(b.hinv[3], x), __dmul_rn(b.hinv[4], y)), __dmul_rn(b.hinv[5], z));
    uz = __dadd_rn(__dadd_rn(__dmul_rn(b.hinv[6], x), __dmul_rn(b.hinv[7], y)), __dmul_rn(b.hinv[8], z));
}

__device__ __forceinline__ double exactR2(double x, double y, double z)
{
    return __dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), __dmul_rn(z, z));
}

// ---- 1. min/max of normalised coordinates -------------------------------------------
__global__ void k_minmax_partial(const double4 *__restrict__ pos, int n, BoxConst b, double *__restrict__ partial)
{
    double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        double u[3];
        normCoord(pos[i], b, u[0], u[1], u[2]);
#pragma unroll
        for (int a = 0; a < 3; a++)
        {
            mn[a] = fmin(mn[a], u[a]);
            mx[a] = fmax(mx[a], u[a]);
        }
    }
    __shared__ double s[6][32];
#pragma unroll
    for (int a = 0; a < 3; a++)
    {
        for (int o = 16; o > 0; o >>= 1)
        {
            mn[a] = fmin(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], o));
            mx[a] = fmax(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], o));
        }
    }
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0)
        for (int a = 0; a < 3; a++)
        {
            s[a][w] = mn[a];
            s[3 + a][w] = mx[a];
        }
    __syncthreads();
    if (threadIdx.x == 0)
    {
        int nw = blockDim.x >> 5;
        for (int a = 0; a < 3; a++)
        {
            double m0 = s[a][0], m1 = s[3 + a][0];
            for (int k = 1; k < nw; k++)
            {
                m0 = fmin(m0, s[a][k]);
                m1 = fmax(m1, s[3 + a][k]);
            }
            partial[blockIdx.x * 6 + a] = m0;
            partial[blockIdx.x * 6 + 3 + a] = m1;
        }
    }
}

// ---- 2. grid parameters (geomMethodDefault, src/geom.c:537-583), one thread ------------
__global__ void k_grid_setup(const double *__restrict__ partial, int nblocks, int nion, BoxConst b, GridDev *g, int maxCells)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double mn[3], mx[3];
    for (int a = 0; a < 3; a++)
    {
        mn[a] = partial[a];
        mx[a] = partial[3 + a];
    }
    for (int k = 1; k < nblocks; k++)
        for (int a = 0; a < 3; a++)
        {
            mn[a] = fmin(mn[a], partial[k * 6 + a]);
            mx[a] = fmax(mx[a], partial[k * 6 + 3 + a]);
        }
    const double span[3] = {b.spanx, b.spany, b.spanz};
    double ext[3], nn[3];
    for (int a = 0; a < 3; a++)
    {
        ext[a] = __dadd_rn(mx[a], -mn[a]);
        double d = __ddiv_rn(b.rcutGeom, span[a]);
        double v = floor(__ddiv_rn(ext[a], d));
        nn[a] = v > 1.0 ? v : 1.0;
    }
    const double coarsen = 1.2599;  // src/geom.c:66
    while (__dadd_rn(__dmul_rn(__dmul_rn(nn[0], nn[1]), nn[2]), -1.0) > (double)nion)
        for (int a = 0; a < 3; a++)
        {
            double v = floor(__ddiv_rn(nn[a], coarsen));
            nn[a] = v > 1.0 ? v : 1.0;
        }
    for (int a = 0; a < 3; a++)
    {
        g->mn[a] = mn[a];
        g->mx[a] = mx[a];
        g->d[a] = __ddiv_rn(ext[a], nn[a]);
        g->n[a] = (int)nn[a];
    }
    g->ncell = g->n[0] * g->n[1] * g->n[2];
    if (g->ncell > maxCells) g->error |= 2;
    g->maxCount = 0;
    g->totalEntries = 0ull;
}

__device__ __forceinline__ int cellIndexOf(const double4 p, const BoxConst &b, const GridDev &g)
{
    // GeomDenseBox second loop, src/geom.c:428-441
    double ux, uy, uz;
    normCoord(p, b, ux, uy, uz);
    int ix = (int)__ddiv_rn(__dadd_rn(ux, -g.mn[0]), g.d[0]);
    int iy = (int)__ddiv_rn(__dadd_rn(uy, -g.mn[1]), g.d[1]);
    int iz = (int)__ddiv_rn(__dadd_rn(uz, -g.mn[2]), g.d[2]);
    ix = max(min(ix, g.n[0] - 1), 0);
    iy = max(min(iy, g.n[1] - 1), 0);
    iz = max(min(iz, g.n[2] - 1), 0);
    return ix + g.n[0] * (iy + g.n[1] * iz);
}

// ---- 3. count beads per cell -----------------------------------------------------------
__global__ void k_cell_count(const double4 *__restrict__ pos, int n, BoxConst b, const GridDev *__restrict__ gp,
                             int *__restrict__ cellOf, int *__restrict__ rank0, int *__restrict__ cellCount)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    GridDev g = *gp;
    if (g.error & 2) return;
    int c = cellIndexOf(pos[i], b, g);
    cellOf[i] = c;
    rank0[i] = atomicAdd(&cellCount[c], 1);
}

// ---- 4. exclusive scan of the cell counts (one block; ncell ~ nion/25) -----------------
__global__ void k_cell_scan(const int *__restrict__ cnt, int *__restrict__ start, const GridDev *__restrict__ gp)
{
    __shared__ int sums[1024];
    const int n = gp->ncell;
    const int per = (n + blockDim.x - 1) / blockDim.x;
    const int lo = min(n, (int)threadIdx.x * per), hi = min(n, lo + per);
    int s = 0;
    for (int i = lo; i < hi; i++) s += cnt[i];
    sums[threadIdx.x] = s;
    __syncthreads();
    // Hillis-Steele inclusive scan over blockDim.x partial sums
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
        start[i] = run;
        run += cnt[i];
    }
    if (threadIdx.x == blockDim.x - 1) start[n] = sums[threadIdx.x];
}

// ---- 5/6. deterministic order inside a cell: by input (bead) index ---------------------
__global__ void k_cell_scatter(int n, const int *__restrict__ cellOf, const int *__restrict__ rank0,
                               const int *__restrict__ cellStart, int *__restrict__ member)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    member[cellStart[cellOf[i]] + rank0[i]] = i;
}

__global__ void k_cell_rank(int n, const int *__restrict__ cellOf, const int *__restrict__ beadOfSlot,
                            const int *__restrict__ cellStart, const int *__restrict__ member, int *__restrict__ perm)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int c = cellOf[i];
    const int lo = cellStart[c], hi = cellStart[c + 1];
    const int me = beadOfSlot[i];
    int r = 0;
    for (int k = lo; k < hi; k++) r += (beadOfSlot[member[k]] < me) ? 1 : 0;
    perm[lo + r] = i;
}

// ---- 7. gather dynamic state into the new slot order -----------------------------------
__global__ void k_gather(int n, const int *__restrict__ perm, const int *__restrict__ cellOld, int *__restrict__ cellNew,
                         const double4 *__restrict__ posOld, double4 *__restrict__ posNew,
                         const double *__restrict__ vxo, const double *__restrict__ vyo, const double *__restrict__ vzo,
                         double *__restrict__ vxn, double *__restrict__ vyn, double *__restrict__ vzn,
                         const int *__restrict__ beadOld, int *__restrict__ beadNew, int *__restrict__ slotOfBead, int nLocal)
{
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const int o = perm[s];
    posNew[s] = posOld[o];
    cellNew[s] = cellOld[o];
    const int bead = beadOld[o];
    beadNew[s] = bead;
    slotOfBead[bead] = s;
    // velocities only exist for local beads; ghosts carry zeros
    vxn[s] = vxo[o];
    vyn[s] = vyo[o];
    vzn[s] = vzo[o];
    (void)nLocal;
}

// ---- 8. raw neighbor pass ---------------------------------------------------------------
// One thread per slot; visits the <=27 periodic neighbour cells (deduplicated when a grid
// dimension has fewer than 3 cells) and applies pairlist1's test bit for bit.
__device__ __forceinline__ bool isPruned(int si, int sj, const int *__restrict__ beadOfSlot, const uint64_t *__restrict__ gid,
                                         const int *__restrict__ molTypeOfBead, const int *__restrict__ molTypeSingle,
                                         const int *__restrict__ bpairOffset, const uint32_t *__restrict__ bpairKey)
{
    // reOrgPairs, src/bioMartini.c:1426-1464
    const int bi = beadOfSlot[si], bj = beadOfSlot[sj];
    const uint64_t gi = gid[bi], gj = gid[bj];
    if ((gi >> 32) != (gj >> 32)) return false;
    // the reference takes the molecule type of the bead that owns the pair (smaller gid)
    const int mt = molTypeOfBead[(gi < gj) ? bi : bj];
    if (mt < 0) return false;
    if (molTypeSingle[mt]) return true;
    const uint32_t a = (uint32_t)(gi & 0xffffull), c = (uint32_t)(gj & 0xffffull);
    const uint32_t key = (min(a, c) << 16) | max(a, c);
    int lo = bpairOffset[mt], hi = bpairOffset[mt + 1];
    while (lo < hi)
    {
        int mid = (lo + hi) >> 1;
        uint32_t v = bpairKey[mid];
        if (v == key) return true;
        if (v < key) lo = mid + 1;
        else hi = mid;
    }
    return false;
}

__global__ void __launch_bounds__(128)
k_nbr_raw(int nLocal, int nIon, int nPad, const double4 *__restrict__ pos, const int *__restrict__ cellOf,
          const int *__restrict__ cellStart, BoxConst b, GridDev *gp, int cap, uint32_t *__restrict__ raw,
          int *__restrict__ count, const int *__restrict__ beadOfSlot, const uint64_t *__restrict__ gid,
          const int *__restrict__ molTypeOfBead, const int *__restrict__ molTypeSingle, const int *__restrict__ bpairOffset,
          const uint32_t *__restrict__ bpairKey, int haveExcl)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nIon) return;
    const int nx = gp->n[0], ny = gp->n[1], nz = gp->n[2];
    const double4 pi = pos[i];
    const int c = cellOf[i];
    const int cx = c % nx, cy = (c / nx) % ny, cz = c / (nx * ny);
    // only local beads own list rows; ghost rows stay empty (multi-GPU)
    int cnt = 0;
    if (i < nLocal)
    {
        const int lx = nx >= 3 ? -1 : 0, hx = nx >= 2 ? 1 : 0;
        const int ly = ny >= 3 ? -1 : 0, hy = ny >= 2 ? 1 : 0;
        const int lz = nz >= 3 ? -1 : 0, hz = nz >= 2 ? 1 : 0;
        for (int dz = lz; dz <= hz; dz++)
            for (int dy = ly; dy <= hy; dy++)
                for (int dx = lx; dx <= hx; dx++)
                {
                    int ax = cx + dx, ay = cy + dy, az = cz + dz;
                    ax = ax < 0 ? ax + nx : (ax >= nx ? ax - nx : ax);
                    ay = ay < 0 ? ay + ny : (ay >= ny ? ay - ny : ay);
                    az = az < 0 ? az + nz : (az >= nz ? az - nz : az);
                    const int cc = ax + nx * (ay + ny * az);
                    const int lo = cellStart[cc], hi = cellStart[cc + 1];
                    for (int j = lo; j < hi; j++)
                    {
                        if (j == i) continue;
                        const double4 pj = pos[j];
                        // pairlist1, src/pairlist.c:280-288
                        double x = __dadd_rn(pi.x, -pj.x), y = __dadd_rn(pi.y, -pj.y), z = __dadd_rn(pi.z, -pj.z);
                        double r2 = exactR2(x, y, z);
                        if (r2 > b.R2cut)
                        {
                            wrapOnce(x, y, z, b);
                            r2 = exactR2(x, y, z);
                        }
                        if (r2 < b.rlist2)
                        {
                            int bin = 0;
#pragma unroll
                            for (int e = 0; e < NBINS - 1; e++) bin += (r2 >= b.binEdge2[e]) ? 1 : 0;
                            uint32_t ent = (uint32_t)j | ((uint32_t)bin << 27);
                            if (haveExcl && isPruned(i, j, beadOfSlot, gid, molTypeOfBead, molTypeSingle, bpairOffset, bpairKey))
                                ent |= EXCL_BIT;
                            if (cnt < cap) raw[(size_t)cnt * nPad + i] = ent;
                            cnt++;
                        }
                    }
                }
    }
    count[i] = cnt;
    // statistics + overflow flag
    int m = cnt;
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    unsigned long long t = (unsigned long long)cnt;
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if ((threadIdx.x & 31) == 0)
    {
        atomicMax(&gp->maxCount, m);
        atomicAdd(&gp->totalEntries, t);
        if (m > cap) atomicOr(&gp->error, 1);
    }
}

// ---- 9. order every row by build-time distance bin (stable within a bin) ----------------
__global__ void __launch_bounds__(128)
k_nbr_order(int nLocal, int nPad, int cap, const uint32_t *__restrict__ raw, const int *__restrict__ count, uint32_t *__restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nLocal) return;
    const int n = min(count[i], cap);
    int off[NBINS];
#pragma unroll
    for (int k = 0; k < NBINS; k++) off[k] = 0;
    for (int k = 0; k < n; k++)
    {
        const uint32_t e = raw[(size_t)k * nPad + i];
        const int bin = (e >> 27) & 7;
#pragma unroll
        for (int q = 0; q < NBINS; q++) off[q] += (q > bin) ? 1 : 0;   // exclusive prefix, branch-free
    }
    for (int k = 0; k < n; k++)
    {
        const uint32_t e = raw[(size_t)k * nPad + i];
        const int bin = (e >> 27) & 7;
        int dst = 0;
#pragma unroll
        for (int q = 0; q < NBINS; q++)
            if (q == bin) dst = off[q]++;
        out[(size_t)dst * nPad + i] = (e & 0x07ffffffu) | (e & EXCL_BIT);
    }
}

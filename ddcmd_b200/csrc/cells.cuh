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

// ---- 2. grid parameters (geomMethodDefault, src/geom.c:537-583), one warp ---------------
__global__ void k_grid_setup(const double *__restrict__ partial, int nblocks, int nion, BoxConst b, GridDev *g, int maxCells)
{
    double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
    for (int k = threadIdx.x; k < nblocks; k += 32)
        for (int a = 0; a < 3; a++)
        {
            mn[a] = fmin(mn[a], partial[k * 6 + a]);
            mx[a] = fmax(mx[a], partial[k * 6 + 3 + a]);
        }
    for (int a = 0; a < 3; a++)
        for (int o = 16; o > 0; o >>= 1)
        {
            mn[a] = fmin(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], o));
            mx[a] = fmax(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], o));
        }
    if (threadIdx.x != 0) return;
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
    g->maxRaw = 0;
    g->totalEntries = 0ull;
}

__device__ __forceinline__ int spread2(int v) { return (v & 1) | ((v & 2) << 2); }   // bits 0,1 -> bits 0,3

__device__ __forceinline__ int cellIndexOf(const double4 p, const BoxConst &b, const GridDev &g, int &subKey)
{
    // GeomDenseBox second loop, src/geom.c:428-441
    double ux, uy, uz;
    normCoord(p, b, ux, uy, uz);
    const double qx = __ddiv_rn(__dadd_rn(ux, -g.mn[0]), g.d[0]);
    const double qy = __ddiv_rn(__dadd_rn(uy, -g.mn[1]), g.d[1]);
    const double qz = __ddiv_rn(__dadd_rn(uz, -g.mn[2]), g.d[2]);
    int ix = (int)qx, iy = (int)qy, iz = (int)qz;
    ix = max(min(ix, g.n[0] - 1), 0);
    iy = max(min(iy, g.n[1] - 1), 0);
    iz = max(min(iz, g.n[2] - 1), 0);
    // 4x4x4 sub-cells in Morton order: only the slot order inside a cell depends on it (locality of the
    // per-step gathers), never a parity-visible quantity
    const int sx = max(min((int)((qx - ix) * 4.0), 3), 0), sy = max(min((int)((qy - iy) * 4.0), 3), 0),
              sz = max(min((int)((qz - iz) * 4.0), 3), 0);
    subKey = spread2(sx) | (spread2(sy) << 1) | (spread2(sz) << 2);
    return ix + g.n[0] * (iy + g.n[1] * iz);
}

// ---- 3. count beads per cell -----------------------------------------------------------
__global__ void k_cell_count(const double4 *__restrict__ pos, int n, BoxConst b, const GridDev *__restrict__ gp,
                             int *__restrict__ cellOf, int *__restrict__ rank0, int *__restrict__ cellCount,
                             const int *__restrict__ beadOfSlot, uint64_t *__restrict__ orderKey)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    GridDev g = *gp;
    if (g.error & 2) return;
    int sub;
    const double4 p = pos[i];
    int c = cellIndexOf(p, b, g, sub);
    // slots are ordered by (local before ghost, cell, sub-cell, bead): the sort key is the cell index, plus the number of cells
    // for a ghost, so [0, nLocal) are exactly the local slots and every kernel over local beads runs over a dense range
    if ((((unsigned long long)__double_as_longlong(p.w)) >> 63) != 0ull) c += g.ncell;
    cellOf[i] = c;
    orderKey[i] = ((uint64_t)sub << 32) | (uint32_t)beadOfSlot[i];
    rank0[i] = atomicAdd(&cellCount[c], 1);
}

// ---- 4. exclusive scan of the cell counts (one block; ncell ~ nion/25; local cells, then ghost cells) -----------------
__global__ void k_cell_scan(const int *__restrict__ cnt, int *__restrict__ start, const GridDev *__restrict__ gp)
{
    __shared__ int sums[1024];
    const int n = 2 * gp->ncell;
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

// ---- 5/6. deterministic order inside a cell: by (sub-cell Morton key, input bead index) -----
__global__ void k_cell_scatter(int n, const int *__restrict__ cellOf, const int *__restrict__ rank0,
                               const int *__restrict__ cellStart, int *__restrict__ member)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    member[cellStart[cellOf[i]] + rank0[i]] = i;
}

__global__ void k_cell_rank(int n, const int *__restrict__ cellOf, const uint64_t *__restrict__ orderKey,
                            const int *__restrict__ cellStart, const int *__restrict__ member, int *__restrict__ perm)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int c = cellOf[i];
    const int lo = cellStart[c], hi = cellStart[c + 1];
    const uint64_t me = orderKey[i];
    int r = 0;
    for (int k = lo; k < hi; k++) r += (orderKey[member[k]] < me) ? 1 : 0;
    perm[lo + r] = i;
}

// ---- 7. gather dynamic state into the new slot order -----------------------------------
__global__ void k_gather(int n, const int *__restrict__ perm, const int *__restrict__ cellOld, int *__restrict__ cellNew,
                         const double4 *__restrict__ posOld, double4 *__restrict__ posNew,
                         const double *__restrict__ vxo, const double *__restrict__ vyo, const double *__restrict__ vzo,
                         double *__restrict__ vxn, double *__restrict__ vyn, double *__restrict__ vzn,
                         const int *__restrict__ beadOld, int *__restrict__ beadNew, int *__restrict__ slotOfBead, int nLocal,
                         float4 *__restrict__ pos32, double *__restrict__ bx, double *__restrict__ by, double *__restrict__ bz)
{
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const int o = perm[s];
    const double4 p = posOld[o];
    posNew[s] = p;
    // candidate filter of the list build; w = 1 marks a ghost slot (no row of its own)
    pos32[s] = make_float4((float)p.x, (float)p.y, (float)p.z, (((unsigned long long)__double_as_longlong(p.w)) >> 63) ? 1.0f : 0.0f);
    bx[s] = p.x;                                                        // build-time positions: displacement bound
    by[s] = p.y;
    bz[s] = p.z;
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

// ---- 8. candidate pass (fp32, conservative) ----------------------------------------------
// One thread per local slot walks the <=27 periodic neighbour cells and keeps every j whose
// single-precision distance is below (rcut+skin)^2 times a safety margin covering the fp32
// rounding of the coordinates.  It only shrinks the work of the exact pass below; every list
// decision is taken there in fp64 with the reference's own arithmetic.
// The x-adjacent cells of a stencil row are one run of consecutive slots, so a bead scans 9 long runs (plus the rare cell that
// wraps around) instead of 27 short ones, and the nearest-image arithmetic of boxes with fewer than three cells along an axis
// lives in a loop of its own: the common loop is a load, seven flops, two compares.
// One run of consecutive candidate slots, 32 at a time: the distance tests of a chunk only set bits of a mask (eight loads in flight,
// no store and no branch between them), then the set bits are turned into row entries.  Interleaving test and store cost 24
// instructions per candidate, 10 of them the store path that some lane of the warp takes in nearly every iteration.
template <bool IMAGE>
__device__ __forceinline__ void filterRun(int lo, int hi, int i, float bx, float by, float bz, bool px, bool py, bool pz, float Lx, float Ly, float Lz,
                                          float rl2f, const float4 *__restrict__ pos32, int nPad, int cap, uint32_t *__restrict__ raw, int &cnt)
{
    const float hx2 = 0.5f * Lx, hy2 = 0.5f * Ly, hz2 = 0.5f * Lz;
    auto inside = [&](int j) {
        const float4 pj = pos32[j];
        float x = bx - pj.x, y = by - pj.y, z = bz - pj.z;
        if (IMAGE)
        {
            if (px) { if (x > hx2) x -= Lx; if (x < -hx2) x += Lx; }
            if (py) { if (y > hy2) y -= Ly; if (y < -hy2) y += Ly; }
            if (pz) { if (z > hz2) z -= Lz; if (z < -hz2) z += Lz; }
        }
        return x * x + y * y + z * z < rl2f;
    };
    for (int j0 = lo; j0 < hi; j0 += 32)
    {
        // one code path for whole and partial chunks (slots past the end of the run re-test its last candidate and are masked off):
        // the lanes of a warp sit in two cells with different runs, and two paths would be executed one after the other
        unsigned mask = 0u;
        const int last = hi - 1;
#pragma unroll
        for (int q = 0; q < 32; q++) mask |= inside(min(j0 + q, last)) ? (1u << q) : 0u;      // bit positions are compile-time constants
        if (hi - j0 < 32) mask &= (1u << (hi - j0)) - 1u;
        if ((unsigned)(i - j0) < 32u) mask &= ~(1u << (i - j0));      // the bead itself
        while (mask)
        {
            const int q = __ffs(mask) - 1;
            mask &= mask - 1u;
            if (cnt < cap) raw[(size_t)cnt * nPad + i] = (uint32_t)(j0 + q);
            cnt++;
        }
    }
}

__global__ void __launch_bounds__(128)
k_nbr_filter(int nIon, int nPad, const float4 *__restrict__ pos32, const int *__restrict__ cellOf,
             const int *__restrict__ cellStart, BoxConst b, float rl2f, GridDev *gp, int cap, uint32_t *__restrict__ raw,
             int *__restrict__ rawCount)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    int cnt = 0;
    const float4 pi = pos32[i < nIon ? i : 0];
    if (i < nIon && pi.w == 0.0f)
    {
        const int nx = gp->n[0], ny = gp->n[1], nz = gp->n[2], ncell = nx * ny * nz;
        const float Lx = (float)b.hxx, Ly = (float)b.hyy, Lz = (float)b.hzz;
        // with >= 3 cells along an axis a wrapped stencil cell has ONE possible image: shift it;
        // with fewer the stencil is deduplicated and each pair takes its nearest image
        const bool px = nx < 3, py = ny < 3, pz = nz < 3;
        const bool anyImage = px || py || pz;
        const int c = cellOf[i];         // a local slot: the plain cell index
        const int cx = c % nx, cy = (c / nx) % ny, cz = c / (nx * ny);
        const int ly = ny >= 3 ? -1 : 0, hy = ny >= 2 ? 1 : 0;
        const int lz = nz >= 3 ? -1 : 0, hz = nz >= 2 ? 1 : 0;
        for (int dz = lz; dz <= hz; dz++)
        {
            int az = cz + dz;
            float sz = 0.0f;
            if (az < 0) { az += nz; sz = -Lz; }
            else if (az >= nz) { az -= nz; sz = Lz; }
            const float bz = pz ? pi.z : pi.z - sz;
            for (int dy = ly; dy <= hy; dy++)
            {
                int ay = cy + dy;
                float sy = 0.0f;
                if (ay < 0) { ay += ny; sy = -Ly; }
                else if (ay >= ny) { ay -= ny; sy = Ly; }
                const float by = py ? pi.y : pi.y - sy;
                const int rowBase = nx * (ay + ny * az);
                // runs along x.  nx >= 3: the cells cx-1 .. cx+1 clipped to the row, then the one cell that wraps around (most beads have
                // none); nx < 3: the bead's own cell, then the other one (nx = 2), pairs take their nearest image
                for (int sg = 0; sg < 2; sg++)
                {
                    int c0, c1;      // cells c0 .. c1 of the row; c1 < c0: nothing
                    float sx = 0.0f;
                    if (!px)
                    {
                        if (sg == 0) { c0 = max(cx - 1, 0); c1 = min(cx + 1, nx - 1); }
                        else if (cx == 0) { c0 = c1 = nx - 1; sx = -Lx; }
                        else if (cx == nx - 1) { c0 = c1 = 0; sx = Lx; }
                        else { c0 = 1; c1 = 0; }
                    }
                    else
                    {
                        if (sg == 0) c0 = c1 = cx;
                        else if (nx == 2) c0 = c1 = 1 - cx;
                        else { c0 = 1; c1 = 0; }
                    }
                    if (c1 < c0) continue;
                    const float bx = px ? pi.x : pi.x - sx;
                    // the cells' local beads, then (several ranks) their ghosts
                    for (int part = 0; part < 2; part++)
                    {
                        const int lo = cellStart[part * ncell + rowBase + c0], hi = cellStart[part * ncell + rowBase + c1 + 1];
                        if (anyImage) filterRun<true>(lo, hi, i, bx, by, bz, px, py, pz, Lx, Ly, Lz, rl2f, pos32, nPad, cap, raw, cnt);
                        else filterRun<false>(lo, hi, i, bx, by, bz, px, py, pz, Lx, Ly, Lz, rl2f, pos32, nPad, cap, raw, cnt);
                    }
                }
            }
        }
    }
    if (i < nIon) rawCount[i] = cnt;
    int m = cnt;
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0)
    {
        atomicMax(&gp->maxRaw, m);
        if (m > cap) atomicOr(&gp->error, 1);
    }
}

// ---- 9. exact pass: pairlist1's test bit for bit, reOrgPairs' pruning ----
__device__ __forceinline__ bool isPruned(int bi, int bj, const uint64_t *__restrict__ gid,
                                         const int *__restrict__ molTypeOfBead, const int *__restrict__ molTypeSingle,
                                         const int *__restrict__ bpairOffset, const uint32_t *__restrict__ bpairKey)
{
    // reOrgPairs, src/bioMartini.c:1426-1464
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

__device__ __forceinline__ double4 ldPos256(const double4 *p)
{
#ifdef DDCB200_EMU
    return *p;
#else
    double4 r;
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p));
    return r;
#endif
}

// ---- 9b. exact pass in one sweep, rows in two segments ----------------------------------------------------------------------------
// pairlist1's decision for every candidate (src/pairlist.c:280-288, the reference's arithmetic without FMA: bit-exact membership) and
// reOrgPairs' pruning flag; each entry is written once, straight to its place: the entries listed closer than nearEdge go to the
// front of the bead's row, the others from the end of the row's allocation backwards - two segments in candidate order.  (Until
// round 2 an exact pass k_nbr_exact sorted each row into eight distance bins with a counting sweep and a placing sweep: 1.56 ms and
// 2.6 GB moved for 0.9 GB of rows and candidates against 0.84 ms here; profiles/r02c_k_nbr_exact_ncu_full.txt, r02t_*.)  The lanes of a warp accept nearly every candidate (the fp32
// filter is tight), so their cursors advance together and a warp's stores fall into a few lines.  The pair walk visits the front
// segment first: nearly all of it is inside the cutoff and right after a build it is all the pruned rows need, while the far
// segment is nearly all outside - the lanes of a warp agree on whether the force block runs, which is what the bins were for.
// Two one-warp-per-bead builds that fused the candidate scan into this pass were measured and lost (1.5 G and 3.2 G warp
// instructions against 1.2 G for the two passes; profiles/r02q_k_nbr_build_ncu_full.txt, r02s_k_nbr_tile_ncu_full.txt).
__global__ void __launch_bounds__(128)
k_nbr_exact2(int nIon, int nPad, int cap, const double4 *__restrict__ pos, BoxConst b, double near2, GridDev *gp, const uint32_t *__restrict__ raw,
             const int *__restrict__ rawCount, uint32_t *__restrict__ out, int *__restrict__ count, uint16_t *__restrict__ cum,
             const uint64_t *__restrict__ gid, const int *__restrict__ molTypeOfBead, const int *__restrict__ molTypeSingle,
             const int *__restrict__ bpairOffset, const uint32_t *__restrict__ bpairKey, int haveExcl, int *__restrict__ tileGhost)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    int nNear = 0, nFar = 0;
    bool ghostEntry = false;      // some entry of this row is a ghost slot (several ranks): the row waits for the halo
    if (i < nIon)
    {
        const int n = min(rawCount[i], cap);
        const double4 pi = pos[i];
        const uint64_t wi = (uint64_t)__double_as_longlong(pi.w);
        // two candidates in flight: the gathers of a pair of candidates are issued before either is tested
        uint32_t jn0 = (0 < n) ? raw[i] : 0u, jn1 = (1 < n) ? raw[(size_t)nPad + i] : 0u;
        for (int k = 0; k < n; k += 2)
        {
            const uint32_t j0 = jn0, j1 = jn1;
            const bool has1 = k + 1 < n;
            const double4 p0 = ldPos256(pos + j0);
            double4 p1 = p0;
            if (has1) p1 = ldPos256(pos + j1);
            if (k + 2 < n) jn0 = raw[(size_t)(k + 2) * nPad + i];
            if (k + 3 < n) jn1 = raw[(size_t)(k + 3) * nPad + i];
#pragma unroll
            for (int u = 0; u < 2; u++)
            {
                if (u == 1 && !has1) break;
                const uint32_t j = u ? j1 : j0;
                const double4 pj = u ? p1 : p0;
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
                    uint32_t ent = j;
                    const uint64_t wj = (uint64_t)__double_as_longlong(pj.w);
                    ghostEntry |= (wj >> 63) != 0ull;
                    if (haveExcl)
                    {
                        // same molecule? bits 16..31 of w carry the low 16 bits of gid>>32: cheap reject before the gid gathers
                        if (((wi ^ wj) & 0xffff0000ull) == 0ull &&
                            isPruned((int)((wi >> 32) & 0x7fffffffull), (int)((wj >> 32) & 0x7fffffffull), gid, molTypeOfBead, molTypeSingle, bpairOffset, bpairKey))
                            ent |= EXCL_BIT;
                    }
                    if (r2 < near2)
                    {
                        out[(size_t)nNear * nPad + i] = ent;
                        nNear++;
                    }
                    else
                    {
                        out[(size_t)(cap - 1 - nFar) * nPad + i] = ent;
                        nFar++;
                    }
                }
            }
        }
        // cum[0] = the front segment, cum[1..] = the whole row: the pair walk's two "bins" (every bin edge is nearEdge)
        const int total = nNear + nFar;
        cum[i] = (uint16_t)nNear;
#pragma unroll
        for (int bnd = 1; bnd < NBINS; bnd++) cum[(size_t)bnd * nPad + i] = (uint16_t)total;
        count[i] = total;
    }
    // statistics
    const int total = nNear + nFar;
    int m = total;
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    unsigned long long t = (unsigned long long)total;
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if ((threadIdx.x & 31) == 0 && m > 0)
    {
        atomicMax(&gp->maxCount, m);
        atomicAdd(&gp->totalEntries, t);
    }
    if (tileGhost)
    {
        // one tile of k_pair = this block (TILE threads): does any of its rows read a ghost position?
        __shared__ int anyGhost;
        if (threadIdx.x == 0) anyGhost = 0;
        __syncthreads();
        if (ghostEntry) atomicOr(&anyGhost, 1);
        __syncthreads();
        if (threadIdx.x == 0) tileGhost[blockIdx.x] = anyGhost;
    }
}

// Order of the k_pair tiles on several ranks: the tiles whose rows touch no ghost first (they run while the halo is in
// flight), then the others; ascending inside each group.  One block; nInterior goes to the grid record.
__global__ void __launch_bounds__(1024)
k_tile_order(int nTiles, const int *__restrict__ tileGhost, int *__restrict__ order, GridDev *gp)
{
    __shared__ int sums[1024];
    __shared__ int totalInterior;
    const int per = (nTiles + blockDim.x - 1) / blockDim.x;
    const int lo = min(nTiles, (int)threadIdx.x * per), hi = min(nTiles, lo + per);
    int s = 0;
    for (int t = lo; t < hi; t++) s += tileGhost[t] ? 0 : 1;
    sums[threadIdx.x] = s;
    __syncthreads();
    for (int o = 1; o < (int)blockDim.x; o <<= 1)
    {
        int v = ((int)threadIdx.x >= o) ? sums[threadIdx.x - o] : 0;
        __syncthreads();
        sums[threadIdx.x] += v;
        __syncthreads();
    }
    if (threadIdx.x == blockDim.x - 1) totalInterior = sums[threadIdx.x];
    __syncthreads();
    int ni = sums[threadIdx.x] - s;            // interior tiles before my chunk
    int nb = lo - ni;                          // boundary tiles before my chunk
    for (int t = lo; t < hi; t++)
    {
        if (tileGhost[t]) order[totalInterior + nb++] = t;
        else order[ni++] = t;
    }
    if (threadIdx.x == 0) gp->nInterior = totalInterior;
}


// ---- at a prune step: the current positions become the reference of the displacement bounds ------------------------------------------
// (k_pair2 MODE 1 writes the pruned rows from the same positions.)  dmax2[2] keeps a bound of the displacement since the BUILD
// for the callers that still need one (the pair correlation's cell walk): the sum of the maxima of the intervals between prunes.
__global__ void __launch_bounds__(256)
k_rebase(int nIon, const double4 *__restrict__ pos, double *__restrict__ bx, double *__restrict__ by, double *__restrict__ bz,
         float *__restrict__ dispOfSlot, unsigned long long *__restrict__ dmax2)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < nIon)
    {
        const double4 p = pos[s];
        bx[s] = p.x;
        by[s] = p.y;
        bz[s] = p.z;
        dispOfSlot[s] = 0.0f;
    }
    if (s == 0)
    {
        const unsigned long long m = max(dmax2[0], dmax2[1]);
        const double base = __longlong_as_double((long long)dmax2[2]) + sqrt(__longlong_as_double((long long)m)) * (1.0 + 1e-12);
        dmax2[2] = (unsigned long long)__double_as_longlong(base);
        dmax2[0] = 0ull;
        dmax2[1] = 0ull;
    }
}

// ---- per step: the displacement bound of every cell's neighbourhood ---------------------------------------------------------------
// out[c] = largest squared displacement since the build of any bead in the stencil cells of cell c (its local beads, and with
// withGhosts also its ghosts).  A partner j of a bead of cell c was in one of those cells at the build (that is how the list
// is made), so d_j is at most the square root of this: k_pair's walk bound no longer pays for the fastest bead of the whole
// system (the maximum over a million beads is about 5.5 sigma, over the ~750 of a neighbourhood about 4 sigma).
__global__ void __launch_bounds__(128)
k_nbr_dmax(const GridDev *__restrict__ gp, const unsigned long long *__restrict__ cellDmax, int withGhosts, unsigned long long *__restrict__ out)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int nx = gp->n[0], ny = gp->n[1], nz = gp->n[2], ncell = nx * ny * nz;
    if (c >= ncell) return;
    const int cx = c % nx, cy = (c / nx) % ny, cz = c / (nx * ny);
    const int lx = nx >= 3 ? -1 : 0, hx = nx >= 2 ? 1 : 0;
    const int ly = ny >= 3 ? -1 : 0, hy = ny >= 2 ? 1 : 0;
    const int lz = nz >= 3 ? -1 : 0, hz = nz >= 2 ? 1 : 0;
    unsigned long long m = 0ull;
    for (int dz = lz; dz <= hz; dz++)
    {
        int az = cz + dz;
        if (az < 0) az += nz; else if (az >= nz) az -= nz;
        for (int dy = ly; dy <= hy; dy++)
        {
            int ay = cy + dy;
            if (ay < 0) ay += ny; else if (ay >= ny) ay -= ny;
            for (int dx = lx; dx <= hx; dx++)
            {
                int ax = cx + dx;
                if (ax < 0) ax += nx; else if (ax >= nx) ax -= nx;
                const int cc = ax + nx * (ay + ny * az);
                m = max(m, cellDmax[cc]);
                if (withGhosts) m = max(m, cellDmax[cc + ncell]);
            }
        }
    }
    out[c] = m;
}

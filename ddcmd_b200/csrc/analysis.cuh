// analysis.cuh - pair correlation counts on the device.
//
// Replaces the pair loops of paircorrelation_eval_{geom,grid,nbrList} (src/paircorrelation.c:174-420): every pair with
// gid_i < gid_j and r < rmax is counted in (species pair, distance bin), 2 for a same-species pair and 1 otherwise.  The
// three methods of the reference find the same pairs; here they come from a walk over the cell structure of the last
// build, widened by the displacement since then.  Counts are integers, so the atomics make the result independent of the
// order of the threads.
#pragma once
#include "engine.cuh"
#include "cells.cuh"

__host__ __device__ inline int comboIndexOf(int i, int j, int ns)
{
    // comboIndex, src/paircorrelation.c:516-521
    const int mx = i > j ? i : j, mn = i > j ? j : i;
    return (mx - mn) + ns * mn - (mn * (mn - 1)) / 2;
}

__global__ void __launch_bounds__(TILE)
k_paircorr(int nIon, const double4 *__restrict__ pos, const int *__restrict__ cellOfSlot, const int *__restrict__ cellStart,
           const GridDev *__restrict__ gp, int reach, const uint64_t *__restrict__ gid, const int *__restrict__ speciesOfBead, BoxConst b,
           double rmax2, double rmin, double delta, int logScale, double log10rmin, int nBins, int ns,
           unsigned long long *__restrict__ hist, unsigned long long *__restrict__ nAtoms)
{
    // One thread per local bead walks the cells within `reach` cells of the cell the bead was sorted into at the last build.
    // Slots keep their build-time cell between builds, and the host chose reach so that reach x (smallest cell edge) covers
    // rmax + twice the largest displacement since the build: every pair with r < rmax now is among the candidates.
    const int i = blockIdx.x * TILE + threadIdx.x;
    if (i >= nIon) return;
    const double4 pi = pos[i];
    const uint64_t wi = (uint64_t)__double_as_longlong(pi.w);
    if (wi >> 63) return;                                  // ghosts are counted by their owner
    const int bi = (int)((wi >> 32) & 0x7fffffffull);
    const int si = speciesOfBead[bi];
    const uint64_t gi = gid[bi];
    atomicAdd(&nAtoms[si], 1ull);
    const int nx = gp->n[0], ny = gp->n[1], nz = gp->n[2];
    const int c = cellOfSlot[i];
    const int cx = c % nx, cy = (c / nx) % ny, cz = c / (nx * ny);
    // along an axis with fewer than 2 reach + 1 cells every cell is visited exactly once
    const int wx = min(2 * reach + 1, nx), wy = min(2 * reach + 1, ny), wz = min(2 * reach + 1, nz);
    const int x0 = wx == nx ? 0 : cx - reach, y0 = wy == ny ? 0 : cy - reach, z0 = wz == nz ? 0 : cz - reach;
    for (int dz = 0; dz < wz; dz++)
    {
        const int az = ((z0 + dz) % nz + nz) % nz;
        for (int dy = 0; dy < wy; dy++)
        {
            const int ay = ((y0 + dy) % ny + ny) % ny;
            for (int dx = 0; dx < wx; dx++)
            {
                const int ax = ((x0 + dx) % nx + nx) % nx;
                const int cc = ax + nx * (ay + ny * az);
                for (int part = 0; part < 2; part++)      // the cell's local beads, then its ghosts
                for (int j = cellStart[cc + part * nx * ny * nz]; j < cellStart[cc + part * nx * ny * nz + 1]; j++)
                {
                    const double4 pj = pos[j];
                    const int bj = (int)((((uint64_t)__double_as_longlong(pj.w)) >> 32) & 0x7fffffffull);
                    if (!(gi < gid[bj])) continue;
                    double x = __dadd_rn(pi.x, -pj.x), y = __dadd_rn(pi.y, -pj.y), z = __dadd_rn(pi.z, -pj.z);
                    double r2 = exactR2(x, y, z);
                    if (r2 > b.R2cut)
                    {
                        wrapOnce(x, y, z, b);
                        r2 = exactR2(x, y, z);
                    }
                    if (!(r2 < rmax2)) continue;
                    const double r = sqrt(r2);
                    // linearBins / logBins, src/paircorrelation.c:60-68
                    const double q = logScale ? __ddiv_rn(__dadd_rn(log10(r), -log10rmin), delta) : __ddiv_rn(__dadd_rn(r, -rmin), delta);
                    const int bin = (int)q;
                    if (q < 0.0 || bin >= nBins) continue;
                    const int sj = speciesOfBead[bj];
                    atomicAdd(&hist[(size_t)bin + (size_t)nBins * comboIndexOf(si, sj, ns)], si == sj ? 2ull : 1ull);
                }
            }
        }
    }
}

// ---- per-group / per-species kinetic terms and the thermal flux ---------------------------------------------------------
// kinetic_terms (src/energy.c:48-163) also files rk, mass, number and the kinetic stress under the bead's GROUP and SPECIES and
// sums the thermal flux J = sum (K + U) v - S v / 2 (U and S, the per-particle energy and stress, are not kept by the Martini
// path: they stay zero, so J = sum K v).  One CTA per class (group or species) strides over the slots in a fixed thread -> slot
// assignment and reduces in a fixed tree: deterministic.  out[class][13] = rk, mass, number, m v_a v_b (xx yy zz xy xz yz), K v (xyz), pad.
__global__ void __launch_bounds__(256)
k_kinetic_classes(int nIon, const double4 *__restrict__ pos, const double *__restrict__ vx, const double *__restrict__ vy, const double *__restrict__ vz,
                  const double *__restrict__ massOfBead, const int *__restrict__ classOfBead, double *__restrict__ out)
{
    const int cls = blockIdx.x;
    double a[12];
#pragma unroll
    for (int k = 0; k < 12; k++) a[k] = 0.0;
    for (int i = threadIdx.x; i < nIon; i += blockDim.x)
    {
        const uint64_t w = (uint64_t)__double_as_longlong(pos[i].w);
        if (w >> 63) continue;
        const int bead = (int)((w >> 32) & 0x7fffffffull);
        if (classOfBead[bead] != cls) continue;
        const double m = massOfBead[bead], x = vx[i], y = vy[i], z = vz[i];
        const double K = 0.5 * m * (x * x + y * y + z * z);
        a[0] += K; a[1] += m; a[2] += 1.0;
        a[3] += m * x * x; a[4] += m * y * y; a[5] += m * z * z; a[6] += m * x * y; a[7] += m * x * z; a[8] += m * y * z;
        a[9] += K * x; a[10] += K * y; a[11] += K * z;
    }
    __shared__ double red[12][8];
#pragma unroll
    for (int k = 0; k < 12; k++)
    {
        double t = a[k];
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = t;
    }
    __syncthreads();
    if (threadIdx.x < 12)
    {
        double t = 0.0;
        for (int w = 0; w < 8; w++) t += red[threadIdx.x][w];
        out[(size_t)cls * 12 + threadIdx.x] = t;
    }
}


// ---- parity hook: order-independent hash of the pair set ------------------------------------------------------------------
// sum and xor (mod 2^64) of a 64-bit mix of (smaller gid, larger gid) over the pairs this rank owns (smaller gid local, as
// pairlist1 reports them, src/pairlist.c:244,279), separately for the interacting and the pruned list: what the million-bead
// parity test compares with the same numbers from the reference (oracle/ref_dump.c "hash").  out: count, sum, xor per list.
__device__ __forceinline__ unsigned long long mix64(unsigned long long x)
{
    x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull;
    x ^= x >> 27; x *= 0x94d049bb133111ebull;
    x ^= x >> 31;
    return x;
}

__global__ void k_pair_hash(int nIon, int nPad, const uint32_t *__restrict__ nbr, const int *__restrict__ count, const int *__restrict__ beadOfSlot,
                            const uint64_t *__restrict__ gid, const uint16_t *__restrict__ cum, int farTop, unsigned long long *__restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long h[6] = {0, 0, 0, 0, 0, 0};
    if (i < nIon)
    {
        const int n = count[i];
        const int nFront = farTop >= 0 ? (int)cum[i] : n;      // rows in two segments (k_nbr_exact2): the second one runs down from farTop
        const uint64_t gi = gid[beadOfSlot[i]];
        for (int k = 0; k < n; k++)
        {
            const uint32_t e = nbr[(size_t)(k < nFront ? k : farTop - (k - nFront)) * nPad + i];
            const int j = (int)(e & 0x07ffffffu);
            const uint64_t gj = gid[beadOfSlot[j]];
            if (gi < gj)
            {
                const unsigned long long v = mix64(mix64(gi) + 0x9e3779b97f4a7c15ull * gj);
                const int l = (e & EXCL_BIT) ? 3 : 0;
                h[l]++;
                h[l + 1] += v;
                h[l + 2] ^= v;
            }
        }
    }
#pragma unroll
    for (int a = 0; a < 6; a++)
    {
        unsigned long long v = h[a];
        for (int o = 16; o > 0; o >>= 1)
        {
            const unsigned long long w = __shfl_xor_sync(0xffffffffu, v, o);
            v = (a % 3 == 2) ? (v ^ w) : (v + w);
        }
        if ((threadIdx.x & 31) == 0)
        {
            if (a % 3 == 2) atomicXor(&out[a], v);
            else atomicAdd(&out[a], v);
        }
    }
}

// nbrcheck.cuh - displacement-triggered list rebuild (DDC updateRate = 0).
//
// Replaces neighborRef + neighborCheck (src/neighbor.c:209-246, :117-208) as called by check4updateNeighbor /
// evalUpdateFlag (src/ddcUpdateAll.c:48-71): with updateRate == 0 the list is rebuilt at the first step where
//     |1 - h0 hinv u|_max (rcut + deltaR)  +  2 sqrt(max_i |(r_i - rbar) - (r0_i - rbar0)|^2)  >=  deltaR,
// rbar = mean of the local beads' positions (image nearest to the domain centre), r0/rbar0 = the same at the build.
// Only launched in that mode; the fixed-rate path (updateRate > 0) never runs these kernels.
#pragma once
#include "engine.cuh"

// per-CTA partial sums of nearestImage(r - c) + c over the local (non-ghost) beads: columns x y z
__global__ void __launch_bounds__(TILE)
k_nbr_rbar_partial(int nIon, const double4 *__restrict__ pos, BoxConst b, double *__restrict__ partial)
{
    const int i = blockIdx.x * TILE + threadIdx.x;
    double s[3] = {0.0, 0.0, 0.0};
    if (i < nIon)
    {
        const double4 p = pos[i];
        if (!((((uint64_t)__double_as_longlong(p.w)) >> 63)))
        {
            // Preduce, orthorhombic pbc 7 (src/preduce.c:282-340): r += h * (-rint(hinv r))
            double x = p.x - b.cx, y = p.y - b.cy, z = p.z - b.cz;
            x += b.hxx * -rint(b.hinv[0] * x);
            y += b.hyy * -rint(b.hinv[4] * y);
            z += b.hzz * -rint(b.hinv[8] * z);
            s[0] = x + b.cx;
            s[1] = y + b.cy;
            s[2] = z + b.cz;
        }
    }
    __shared__ double red[3][TILE / 32];
#pragma unroll
    for (int a = 0; a < 3; a++)
    {
        double t = s[a];
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if ((threadIdx.x & 31) == 0) red[a][threadIdx.x >> 5] = t;
    }
    __syncthreads();
    if (threadIdx.x < 3)
    {
        double t = 0.0;
        for (int w = 0; w < TILE / 32; w++) t += red[threadIdx.x][w];
        partial[(size_t)blockIdx.x * 3 + threadIdx.x] = t;
    }
}

// chk layout (doubles): [0..2] sum of local positions now, [3..5] the same at the build, [6] bits of max d^2
// max over all resident beads (locals and ghosts, as the reference loops over number_particles) of the squared
// displacement since the build, each position taken relative to the mean local position of its own time
__global__ void __launch_bounds__(TILE)
k_nbr_check(int nIon, int nLocal, const double4 *__restrict__ pos, const double *__restrict__ bx, const double *__restrict__ by,
            const double *__restrict__ bz, PairConst pc, double *__restrict__ chk)
{
    const int i = blockIdx.x * TILE + threadIdx.x;
    double d2 = 0.0;
    if (i < nIon)
    {
        const double inv = (double)nLocal;
        const double r1x = chk[0] / inv, r1y = chk[1] / inv, r1z = chk[2] / inv;
        const double r0x = chk[3] / inv, r0y = chk[4] / inv, r0z = chk[5] / inv;
        const double4 p = pos[i];
        double ax = p.x - r1x, ay = p.y - r1y, az = p.z - r1z;
        double cx = bx[i] - r0x, cy = by[i] - r0y, cz = bz[i] - r0z;
#define FAST1(v, hh, h) { if (v > hh) v -= h; if (v < -hh) v += h; }
        FAST1(ax, pc.hhx, pc.hxx) FAST1(ay, pc.hhy, pc.hyy) FAST1(az, pc.hhz, pc.hzz)
        FAST1(cx, pc.hhx, pc.hxx) FAST1(cy, pc.hhy, pc.hyy) FAST1(cz, pc.hhz, pc.hzz)
        double x = ax - cx, y = ay - cy, z = az - cz;
        FAST1(x, pc.hhx, pc.hxx) FAST1(y, pc.hhy, pc.hyy) FAST1(z, pc.hhz, pc.hzz)
#undef FAST1
        d2 = x * x + y * y + z * z;
    }
    for (int o = 16; o > 0; o >>= 1) d2 = fmax(d2, __shfl_xor_sync(0xffffffffu, d2, o));
    if ((threadIdx.x & 31) == 0 && d2 > 0.0)
        atomicMax((unsigned long long *)(chk + 6), (unsigned long long)__double_as_longlong(d2));
}

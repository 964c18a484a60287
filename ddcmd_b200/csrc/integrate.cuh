// integrate.cuh - nglf velocity-Verlet fused with the kinetic-energy / kinetic-stress block
// reductions and the per-step position wrap.
//
// Replaces nglf (src/nglf.c:67-112) with the FREE group update (src/free.c:13-28),
// backInBox_fast (src/preduce.c:147-160), kinetic_terms (src/energy.c:48-163) and
// nglfGPU.cu's freeVelocityUpdate/freePositionUpdate + kineticGPU.cu.
//
// One launch does, per bead:   [kick2 of step n]  [KE of step n]  [kick1 + drift + wrap of step n+1]
// so consecutive steps cost one pass over r, v, f instead of three.
#pragma once
#include "engine.cuh"

#define INT_KICK2 1
#define INT_KE 2
#define INT_KICK1_DRIFT 4

// largest squared displacement since the build of any bead of a cell (k_pair's walk bound uses the maximum over a bead's stencil
// cells, k_nbr_dmax): non-negative doubles order like their bit patterns; the running maximum is read first, so atomics are rare
__device__ __forceinline__ void trackCellDisp(unsigned long long *__restrict__ cellDmax, const int *__restrict__ cellOfSlot, int slot, double disp2)
{
    if (!cellDmax) return;
    unsigned long long *p = cellDmax + cellOfSlot[slot];
    const unsigned long long bits = (unsigned long long)__double_as_longlong(disp2);
    if (bits > *(volatile unsigned long long *)p) atomicMax(p, bits);
}

template <int MODE>
__global__ void __launch_bounds__(TILE)
k_integrate(int nIon, double4 *__restrict__ pos, double *__restrict__ vx, double *__restrict__ vy, double *__restrict__ vz,
            const double *__restrict__ fx, const double *__restrict__ fy, const double *__restrict__ fz,
            const double *__restrict__ massOfBead, double halfDt2, double halfDt1, double dt, PairConst pc, double *__restrict__ partial,
            const double *__restrict__ bx, const double *__restrict__ by, const double *__restrict__ bz, unsigned long long *__restrict__ dmax2,
            float *__restrict__ dispOfSlot, unsigned long long *__restrict__ cellDmax, const int *__restrict__ cellOfSlot)
{
    const int i = blockIdx.x * TILE + threadIdx.x;
    double ke[7] = {0, 0, 0, 0, 0, 0, 0};
    double disp2 = 0.0;
    double4 p = pos[i < nIon ? i : 0];
    if (i < nIon && !((((uint64_t)__double_as_longlong(p.w)) >> 63)))   // ghosts are moved by their owner
    {
        const uint32_t bead = (uint32_t)((((uint64_t)__double_as_longlong(p.w)) >> 32) & 0x7fffffffull);
        const double mass = massOfBead[bead];
        double v0 = vx[i], v1 = vy[i], v2 = vz[i];
        const double f0 = fx[i], f1 = fy[i], f2 = fz[i];
        if (MODE & INT_KICK2)
        {
            // free_velocityUpdate: a = dt/mass ; v += a*f   (src/free.c:24-27)
            const double a = halfDt2 / mass;
            v0 += a * f0;
            v1 += a * f1;
            v2 += a * f2;
        }
        if (MODE & INT_KE)
        {
            // kinetic_terms (src/energy.c:92-112)
            ke[0] = 0.5 * mass * (v0 * v0 + v1 * v1 + v2 * v2);
            ke[1] = mass * v0 * v0;
            ke[2] = mass * v1 * v1;
            ke[3] = mass * v2 * v2;
            ke[4] = mass * v0 * v1;
            ke[5] = mass * v0 * v2;
            ke[6] = mass * v1 * v2;
        }
        if (MODE & INT_KICK1_DRIFT)
        {
            const double a = halfDt1 / mass;
            v0 += a * f0;
            v1 += a * f1;
            v2 += a * f2;
            p.x += dt * v0;
            p.y += dt * v1;
            p.z += dt * v2;
            // backInBox_fast
            if (p.x > pc.hhx) p.x -= pc.hxx;
            if (p.x < -pc.hhx) p.x += pc.hxx;
            if (p.y > pc.hhy) p.y -= pc.hyy;
            if (p.y < -pc.hhy) p.y += pc.hyy;
            if (p.z > pc.hhz) p.z -= pc.hzz;
            if (p.z < -pc.hhz) p.z += pc.hzz;
            pos[i] = p;
            // displacement since the list build (nearest image): bounds which list bins k_pair must visit
            double dx = p.x - bx[i], dy = p.y - by[i], dz = p.z - bz[i];
            if (dx > pc.hhx) dx -= pc.hxx;
            if (dx < -pc.hhx) dx += pc.hxx;
            if (dy > pc.hhy) dy -= pc.hyy;
            if (dy < -pc.hhy) dy += pc.hyy;
            if (dz > pc.hhz) dz -= pc.hzz;
            if (dz < -pc.hhz) dz += pc.hzz;
            disp2 = dx * dx + dy * dy + dz * dz;
            dispOfSlot[i] = __double2float_ru(sqrt(disp2));     // this bead's own displacement, rounded up: k_pair's per-bead walk bound
            trackCellDisp(cellDmax, cellOfSlot, i, disp2);
        }
        if (MODE & (INT_KICK2 | INT_KICK1_DRIFT))
        {
            vx[i] = v0;
            vy[i] = v1;
            vz[i] = v2;
        }
    }
    if (MODE & INT_KICK1_DRIFT)
    {
        // non-negative doubles order like their bit patterns: one integer atomicMax per warp
        // (the running maximum is read first, so almost every warp skips the same-address atomic)
        for (int o = 16; o > 0; o >>= 1) disp2 = fmax(disp2, __shfl_xor_sync(0xffffffffu, disp2, o));
        if ((threadIdx.x & 31) == 0)
        {
            const unsigned long long bits = (unsigned long long)__double_as_longlong(disp2);
            if (bits > *(volatile unsigned long long *)dmax2) atomicMax(dmax2, bits);
        }
    }
    if (MODE & INT_KE)
    {
        __shared__ double red[7][TILE / 32];
#pragma unroll
        for (int a = 0; a < 7; a++)
        {
            double t = ke[a];
            for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
            if ((threadIdx.x & 31) == 0) red[a][threadIdx.x >> 5] = t;
        }
        __syncthreads();
        if (threadIdx.x < 7)
        {
            double t = 0.0;
            for (int w = 0; w < TILE / 32; w++) t += red[threadIdx.x][w];
            partial[(size_t)blockIdx.x * 7 + threadIdx.x] = t;
        }
    }
}

// Deterministic final reduction: out[dstIdx[c]] (+)= sum over blocks of partial[b*ncol + c].
// One CTA per column, fixed tree.
__global__ void __launch_bounds__(256)
k_reduce_cols(const double *__restrict__ partial, int nblocks, int ncol, const int *__restrict__ dstIdx, double *__restrict__ out, int accumulate)
{
    const int c = blockIdx.x;
    double t = 0.0;
    for (int b = threadIdx.x; b < nblocks; b += blockDim.x) t += partial[(size_t)b * ncol + c];
    __shared__ double red[8];
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = t;
    __syncthreads();
    if (threadIdx.x == 0)
    {
        double s = 0.0;
        for (int w = 0; w < 8; w++) s += red[w];
        const int d = dstIdx[c];
        out[d] = accumulate ? out[d] + s : s;
    }
}

// Molecular virial correction (molecularVirial, src/molecularPressure.c:22-55): one thread
// per multi-bead molecule; diagonal only.
__global__ void k_mol_virial(int64_t nMol, const int64_t *__restrict__ molOffset, const int *__restrict__ molBeads,
                             const int *__restrict__ slotOfBead, const double4 *__restrict__ pos,
                             const double *__restrict__ fx, const double *__restrict__ fy, const double *__restrict__ fz,
                             const double *__restrict__ massOfBead, PairConst pc, double *__restrict__ out3)
{
    const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double c[3] = {0, 0, 0};
    if (m < nMol)
    {
        const int64_t lo = molOffset[m], hi = molOffset[m + 1];
        const int s0 = slotOfBead[molBeads[lo]];   // first listed bead = ownership bead
        // the molecule is summed on the rank that owns it: its ownership bead is resident AND not a ghost (there every
        // bead of the molecule is local, src/ddcRuleMolecule.c:43; on other ranks some of its beads may be ghosts or absent)
        if (s0 >= 0 && !((((unsigned long long)__double_as_longlong(pos[s0].w)) >> 63)))
        {
            const double4 p0 = pos[s0];
            double M = 0, Rx = 0, Ry = 0, Rz = 0;
            for (int64_t a = lo; a < hi; a++)
            {
                const int b = molBeads[a];
                const double4 p = pos[slotOfBead[b]];
                const double mass = massOfBead[b];
                double dx = p.x - p0.x, dy = p.y - p0.y, dz = p.z - p0.z;
                dx -= pc.hxx * rint(dx / pc.hxx);
                dy -= pc.hyy * rint(dy / pc.hyy);
                dz -= pc.hzz * rint(dz / pc.hzz);
                Rx += mass * dx;
                Ry += mass * dy;
                Rz += mass * dz;
                M += mass;
            }
            Rx /= M;
            Ry /= M;
            Rz /= M;
            for (int64_t a = lo; a < hi; a++)
            {
                const int s = slotOfBead[molBeads[a]];
                const double4 p = pos[s];
                double dx = p.x - p0.x, dy = p.y - p0.y, dz = p.z - p0.z;
                dx -= pc.hxx * rint(dx / pc.hxx);
                dy -= pc.hyy * rint(dy / pc.hyy);
                dz -= pc.hzz * rint(dz / pc.hzz);
                c[0] -= (dx - Rx) * fx[s];
                c[1] -= (dy - Ry) * fy[s];
                c[2] -= (dz - Rz) * fz[s];
            }
        }
    }
    // block reduce then one atomic per block (order varies run to run at the 1e-16 level; print-time only)
    __shared__ double red[3][8];
    for (int a = 0; a < 3; a++)
    {
        double t = c[a];
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if ((threadIdx.x & 31) == 0) red[a][threadIdx.x >> 5] = t;
    }
    __syncthreads();
    if (threadIdx.x < 3)
    {
        double s = 0;
        for (int w = 0; w < (blockDim.x >> 5); w++) s += red[threadIdx.x][w];
        atomicAdd(out3 + threadIdx.x, s);
    }
}

// pack / unpack helpers between caller ("input") order and slot order
__global__ void k_upload_state(int n, const int *__restrict__ bead, const double *__restrict__ rx, const double *__restrict__ ry,
                               const double *__restrict__ rz, const double *__restrict__ vxi, const double *__restrict__ vyi,
                               const double *__restrict__ vzi, const uint64_t *__restrict__ wOfBead, double4 *__restrict__ pos,
                               double *__restrict__ vx, double *__restrict__ vy, double *__restrict__ vz,
                               int *__restrict__ beadOfSlot, int *__restrict__ slotOfBead)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int b = bead ? bead[i] : i;
    pos[i] = make_double4(rx[i], ry[i], rz[i], __longlong_as_double((long long)wOfBead[b]));
    vx[i] = vxi[i];
    vy[i] = vyi[i];
    vz[i] = vzi[i];
    beadOfSlot[i] = b;
    slotOfBead[b] = i;
}

// new positions and velocities for beads that are already resident (slots, cells and the neighbor list are kept): the
// displacement bookkeeping of the list walk is refreshed from the build-time positions, as the integrator kernels do
__global__ void k_update_state(int n, const int *__restrict__ bead, const int *__restrict__ slotOfBead, const double *__restrict__ rx,
                               const double *__restrict__ ry, const double *__restrict__ rz, const double *__restrict__ vxi,
                               const double *__restrict__ vyi, const double *__restrict__ vzi, double4 *__restrict__ pos,
                               double *__restrict__ vx, double *__restrict__ vy, double *__restrict__ vz, PairConst pc,
                               const double *__restrict__ bx, const double *__restrict__ by, const double *__restrict__ bz,
                               unsigned long long *__restrict__ dmax2, float *__restrict__ dispOfSlot, unsigned long long *__restrict__ cellDmax,
                               const int *__restrict__ cellOfSlot)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double disp2 = 0.0;
    if (i < n)
    {
        const int s = slotOfBead[bead ? bead[i] : i];
        double4 p = pos[s];
        p.x = rx[i];
        p.y = ry[i];
        p.z = rz[i];
        pos[s] = p;
        vx[s] = vxi[i];
        vy[s] = vyi[i];
        vz[s] = vzi[i];
        double dx = p.x - bx[s], dy = p.y - by[s], dz = p.z - bz[s];
        if (dx > pc.hhx) dx -= pc.hxx;
        if (dx < -pc.hhx) dx += pc.hxx;
        if (dy > pc.hhy) dy -= pc.hyy;
        if (dy < -pc.hhy) dy += pc.hyy;
        if (dz > pc.hhz) dz -= pc.hzz;
        if (dz < -pc.hhz) dz += pc.hzz;
        disp2 = dx * dx + dy * dy + dz * dz;
        dispOfSlot[s] = __double2float_ru(sqrt(disp2));
        trackCellDisp(cellDmax, cellOfSlot, s, disp2);
    }
    // one atomicMax per warp, and only when it can raise the running maximum (as k_integrate)
    for (int o = 16; o > 0; o >>= 1) disp2 = fmax(disp2, __shfl_xor_sync(0xffffffffu, disp2, o));
    if ((threadIdx.x & 31) == 0)
    {
        const unsigned long long bits = (unsigned long long)__double_as_longlong(disp2);
        if (bits > *(volatile unsigned long long *)dmax2) atomicMax(dmax2, bits);
    }
}

__global__ void k_download_state(int n, const int *__restrict__ bead, const int *__restrict__ slotOfBead, const double4 *__restrict__ pos,
                                 const double *__restrict__ vx, const double *__restrict__ vy, const double *__restrict__ vz,
                                 const double *__restrict__ fx, const double *__restrict__ fy, const double *__restrict__ fz,
                                 double *__restrict__ out)   // out: 9 arrays of n
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int s = slotOfBead[bead ? bead[i] : i];
    const double4 p = pos[s];
    out[0 * (size_t)n + i] = p.x;
    out[1 * (size_t)n + i] = p.y;
    out[2 * (size_t)n + i] = p.z;
    out[3 * (size_t)n + i] = vx[s];
    out[4 * (size_t)n + i] = vy[s];
    out[5 * (size_t)n + i] = vz[s];
    out[6 * (size_t)n + i] = fx[s];
    out[7 * (size_t)n + i] = fy[s];
    out[8 * (size_t)n + i] = fz[s];
}

// nglfcons.cuh - the NGLFCONSTRAINT integrator: velocity-Verlet with per-GROUP velocity updates
// (FREE or LANGEVIN thermostat with the per-bead LCG64 stream), pair-distance constraints solved on
// velocities, and the Berendsen molecular-pressure barostat's position scaling.
//
// Replaces nglfconstraint (src/nglfconstraint.c:510-574) with velocityConstraintOld / resMoveConsOld
// (:177-270, :439-457), changeVolume + adjustPosn (:46-84), langevin_velocityUpdate (src/langevin.c:92-128),
// free_velocityUpdate (src/free.c:13-28), gasdev3d (src/random.c) over lcg64_2 (src/lcg64.c:131-141), and the
// GPU analogue nglfconstraintGPU.cu.  SURVEY.md section 8(f) N1.
//
// k_nglfc<MODE> is k_integrate's sibling: one pass over r, v, f does
//   [BACK velocity update of step n] [kinetic terms] [barostat scaling] [FRONT update of step n+1] [drift + wrap]
// with the parts selected at compile time; systems without constraints run one launch per step, systems with
// constraints split the pass around k_constraint.
#pragma once
#include "engine.cuh"

#define NC_BACK 1
#define NC_KE 2
#define NC_SCALE 4
#define NC_FRONT 8
#define NC_DRIFT 16

#define MAXGROUPS 8
#define GROUP_FREE 0
#define GROUP_LANGEVIN 1

struct GroupTab
{
    int n;
    int type[MAXGROUPS];
    double kBT[MAXGROUPS], tau[MAXGROUPS];
    double vcx[MAXGROUPS], vcy[MAXGROUPS], vcz[MAXGROUPS];
};

// lcg64 (src/lcg64.c:122-130): state = MULT[multID]*state + prime ; r = state * 2^-64
__device__ __forceinline__ double lcg64Next(uint64_t &state, uint64_t mult, uint64_t prime)
{
    state = mult * state + prime;
    return __ull2double_rn(state) * 5.4210108624275222e-20;
}

// one accepted point of the polar method (the do-while of gasdev3d, src/random.c): returns fac, point in (px, py)
__device__ __forceinline__ double polarPoint(uint64_t &state, uint64_t mult, uint64_t prime, double &px, double &py)
{
    double rsq;
    do
    {
        const double ux = lcg64Next(state, mult, prime);
        const double uy = lcg64Next(state, mult, prime);
        px = __dadd_rn(__dmul_rn(2.0, ux), -1.0);
        py = __dadd_rn(__dmul_rn(2.0, uy), -1.0);
        rsq = __dadd_rn(__dmul_rn(px, px), __dmul_rn(py, py));
    } while (rsq >= 1.0 || rsq == 0.0);
    return sqrt(-2.0 * log(rsq) / rsq);
}

// gasdev3d: x, y from the first accepted point, z from the x of a second one
__device__ __forceinline__ void gasdev3d(uint64_t &state, uint64_t mult, uint64_t prime, double &gx, double &gy, double &gz)
{
    double px, py;
    double fac = polarPoint(state, mult, prime, px, py);
    gx = px * fac;
    gy = py * fac;
    fac = polarPoint(state, mult, prime, px, py);
    gz = px * fac;
}

__device__ __forceinline__ uint64_t lcg64Mult(uint32_t multID)
{
    return multID == 0 ? 0x27bb2ee687b0b0fdull : (multID == 1 ? 0x2c6fe96ee78b6955ull : 0x369dea0f31a53f85ull);
}

template <int MODE>
__global__ void __launch_bounds__(TILE)
k_nglfc(int nIon, double4 *__restrict__ pos, double *__restrict__ vx, double *__restrict__ vy, double *__restrict__ vz,
        const double *__restrict__ fx, const double *__restrict__ fy, const double *__restrict__ fz,
        const double *__restrict__ massOfBead, const unsigned char *__restrict__ groupOfBead, uint64_t *__restrict__ rngState,
        const uint2 *__restrict__ rngMP, GroupTab g, double halfDt, double dt, double sx, double sy, double sz, PairConst pc,
        double *__restrict__ partial, const double *__restrict__ bx, const double *__restrict__ by, const double *__restrict__ bz,
        unsigned long long *__restrict__ dmax2, float *__restrict__ dispOfSlot, unsigned long long *__restrict__ cellDmax,
        const int *__restrict__ cellOfSlot)
{
    const int i = blockIdx.x * TILE + threadIdx.x;
    double ke[7] = {0, 0, 0, 0, 0, 0, 0};
    double disp2 = 0.0;
    double4 p = pos[i < nIon ? i : 0];
    if (i < nIon && !((((uint64_t)__double_as_longlong(p.w)) >> 63)))
    {
        const uint32_t bead = (uint32_t)((((uint64_t)__double_as_longlong(p.w)) >> 32) & 0x7fffffffull);
        const double mass = massOfBead[bead];
        double v0 = vx[i], v1 = vy[i], v2 = vz[i];
        double f0 = 0.0, f1 = 0.0, f2 = 0.0;
        int gi = 0;
        bool lang = false;
        uint64_t state = 0, mult = 0, prime = 0;
        double la = 0.0, lc = 0.0, ld = 0.0;
        if (MODE & (NC_BACK | NC_FRONT))
        {
            f0 = fx[i];
            f1 = fy[i];
            f2 = fz[i];
            gi = groupOfBead ? (int)groupOfBead[bead] : 0;
            lang = g.type[gi] == GROUP_LANGEVIN;
            if (lang)
            {
                // langevin_velocityUpdate (src/langevin.c:105-111); dt there is the half step
                state = rngState[bead];
                const uint2 mp = rngMP[bead];
                mult = lcg64Mult(mp.x);
                prime = (uint64_t)mp.y;
                la = exp(-halfDt / g.tau[gi]);
                lc = halfDt / mass;
                ld = sqrt(2.0 * halfDt * g.kBT[gi] / (mass * g.tau[gi]));
            }
        }
        if (MODE & NC_BACK)
        {
            if (lang)
            {
                double gx, gy, gz;
                gasdev3d(state, mult, prime, gx, gy, gz);
                // BACK_TIMESTEP: v = vcm + a*((v - vcm) + c*f + d*g)   (src/langevin.c:121-125)
                v0 = g.vcx[gi] + la * ((v0 - g.vcx[gi]) + lc * f0 + ld * gx);
                v1 = g.vcy[gi] + la * ((v1 - g.vcy[gi]) + lc * f1 + ld * gy);
                v2 = g.vcz[gi] + la * ((v2 - g.vcz[gi]) + lc * f2 + ld * gz);
            }
            else
            {
                const double a = halfDt / mass;   // free_velocityUpdate (src/free.c:24-27)
                v0 += a * f0;
                v1 += a * f1;
                v2 += a * f2;
            }
        }
        if (MODE & NC_KE)
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
        if (MODE & NC_SCALE)
        {
            // adjustPosn: r = hfac r, hfac = h_new h_old^-1 (diagonal here) (src/nglfconstraint.c:46-58)
            p.x = sx * p.x;
            p.y = sy * p.y;
            p.z = sz * p.z;
        }
        if (MODE & NC_FRONT)
        {
            if (lang)
            {
                double gx, gy, gz;
                gasdev3d(state, mult, prime, gx, gy, gz);
                // FRONT_TIMESTEP: v = vcm + a*(v - vcm) + c*f + d*g   (src/langevin.c:116-120)
                v0 = g.vcx[gi] + la * (v0 - g.vcx[gi]) + lc * f0 + ld * gx;
                v1 = g.vcy[gi] + la * (v1 - g.vcy[gi]) + lc * f1 + ld * gy;
                v2 = g.vcz[gi] + la * (v2 - g.vcz[gi]) + lc * f2 + ld * gz;
            }
            else
            {
                const double a = halfDt / mass;
                v0 += a * f0;
                v1 += a * f1;
                v2 += a * f2;
            }
        }
        if ((MODE & (NC_BACK | NC_FRONT)) && lang) rngState[bead] = state;
        if (MODE & NC_DRIFT)
        {
            p.x += dt * v0;
            p.y += dt * v1;
            p.z += dt * v2;
            // backInBox_fast (src/preduce.c:147-160) with the current box
            if (p.x > pc.hhx) p.x -= pc.hxx;
            if (p.x < -pc.hhx) p.x += pc.hxx;
            if (p.y > pc.hhy) p.y -= pc.hyy;
            if (p.y < -pc.hhy) p.y += pc.hyy;
            if (p.z > pc.hhz) p.z -= pc.hzz;
            if (p.z < -pc.hhz) p.z += pc.hzz;
            // displacement since the list build, nearest image in the current box; barostat scaling since the build is
            // part of it, the image mismatch between the two boxes is covered by PairConst::listSlack
            double dx = p.x - bx[i], dy = p.y - by[i], dz = p.z - bz[i];
            if (dx > pc.hhx) dx -= pc.hxx;
            if (dx < -pc.hhx) dx += pc.hxx;
            if (dy > pc.hhy) dy -= pc.hyy;
            if (dy < -pc.hhy) dy += pc.hyy;
            if (dz > pc.hhz) dz -= pc.hzz;
            if (dz < -pc.hhz) dz += pc.hzz;
            disp2 = dx * dx + dy * dy + dz * dz;
            dispOfSlot[i] = __double2float_ru(sqrt(disp2));
            trackCellDisp(cellDmax, cellOfSlot, i, disp2);
        }
        if (MODE & (NC_SCALE | NC_DRIFT)) pos[i] = p;
        if (MODE & (NC_BACK | NC_FRONT))
        {
            vx[i] = v0;
            vy[i] = v1;
            vz[i] = v2;
        }
    }
    if (MODE & NC_DRIFT)
    {
        for (int o = 16; o > 0; o >>= 1) disp2 = fmax(disp2, __shfl_xor_sync(0xffffffffu, disp2, o));
        if ((threadIdx.x & 31) == 0)
        {
            const unsigned long long bits = (unsigned long long)__double_as_longlong(disp2);
            if (bits > *(volatile unsigned long long *)dmax2) atomicMax(dmax2, bits);
        }
    }
    if (MODE & NC_KE)
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

// ---- constraints ------------------------------------------------------------------------------------------------
// One thread per constraint cluster (a CONSTRAINT of genConstraint, src/bioMartini.c:445-565: the atoms and pairs of
// one CONSLISTPARMS of one residue instance).  Clusters are disjoint, so threads never share a bead; inside a cluster
// the Gauss-Seidel sweep runs in the reference's pair order (resMoveConsOld, src/nglfconstraint.c:177-270).
#define CONS_MAXATOM 32
#define CONS_MAXPAIR 48
#define CONS_TOL 1.0e-12
#define CONS_MAXIT 500

template <bool FRONT>
__global__ void __launch_bounds__(64)
k_constraint(int nCons, const int *__restrict__ atomOffset, const int *__restrict__ atomBead, const int *__restrict__ pairOffset,
             const int *__restrict__ pairA, const int *__restrict__ pairB, const double *__restrict__ pairDist,
             const int *__restrict__ slotOfBead, const double4 *__restrict__ pos, double *__restrict__ vx, double *__restrict__ vy,
             double *__restrict__ vz, const double *__restrict__ massOfBead, double dt, PairConst pc, double hix, double hiy, double hiz,
             int *__restrict__ notConverged)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nCons) return;
    const int a0 = atomOffset[c], na = atomOffset[c + 1] - a0;
    const int p0 = pairOffset[c], np = pairOffset[c + 1] - p0;
    if (np == 0) return;
    {
        // several ranks: a cluster is solved where its molecule is local (molecules are whole on their owner); elsewhere its
        // beads are absent or ghosts
        const int s0 = slotOfBead[atomBead[a0]];
        if (s0 < 0 || ((((unsigned long long)__double_as_longlong(pos[s0].w)) >> 63) != 0ull)) return;
    }
    int slot[CONS_MAXATOM];
    double rMass[CONS_MAXATOM], r[CONS_MAXATOM][3], v[CONS_MAXATOM][3];
    double rab[CONS_MAXPAIR][3];
    for (int j = 0; j < na; j++)
    {
        const int b = atomBead[a0 + j];
        const int s = slotOfBead[b];
        slot[j] = s;
        rMass[j] = 1.0 / massOfBead[b];
        const double4 p = pos[s];
        r[j][0] = p.x; r[j][1] = p.y; r[j][2] = p.z;
        v[j][0] = vx[s]; v[j][1] = vy[s]; v[j][2] = vz[s];
    }
    for (int ab = 0; ab < np; ab++)
    {
        const int a = pairA[p0 + ab], b = pairB[p0 + ab];
        double x = r[a][0] - r[b][0], y = r[a][1] - r[b][1], z = r[a][2] - r[b][2];
        // nearestImage = Preduce for an orthorhombic box (src/preduce.c:466, case 7 of dpreduce)
        x += pc.hxx * (-rint(hix * x));
        y += pc.hyy * (-rint(hiy * y));
        z += pc.hzz * (-rint(hiz * z));
        rab[ab][0] = x; rab[ab][1] = y; rab[ab][2] = z;
    }
    int it = 0;
    for (; it < CONS_MAXIT; it++)
    {
        double errMax = 0.0;
        for (int ab = 0; ab < np; ab++)
        {
            const int a = pairA[p0 + ab], b = pairB[p0 + ab];
            const double d = pairDist[p0 + ab];
            const double dist2 = d * d;
            const double vabx = v[a][0] - v[b][0], vaby = v[a][1] - v[b][1], vabz = v[a][2] - v[b][2];
            double fn;
            if (FRONT)
            {
                // frontFunc: ((rab + dt vab)^2 - d^2) / (2 dt)   (src/nglfconstraint.c:117-126)
                const double px = rab[ab][0] + dt * vabx, py = rab[ab][1] + dt * vaby, pz = rab[ab][2] + dt * vabz;
                fn = ((px * px + py * py + pz * pz) - dist2) / (2 * dt);
            }
            else
                fn = rab[ab][0] * vabx + rab[ab][1] * vaby + rab[ab][2] * vabz;   // backFunc
            const double rvab = fn / dist2;
            const double rma = rMass[a], rmb = rMass[b];
            const double gab = -rvab / (rma + rmb);
            const double err = fabs(rvab * dt);
            if (err > errMax) errMax = err;
            const double ca = rma * gab, cb = rmb * gab;
            v[a][0] += ca * rab[ab][0]; v[a][1] += ca * rab[ab][1]; v[a][2] += ca * rab[ab][2];
            v[b][0] -= cb * rab[ab][0]; v[b][1] -= cb * rab[ab][1]; v[b][2] -= cb * rab[ab][2];
        }
        if (errMax < CONS_TOL) break;
    }
    if (it == CONS_MAXIT) atomicAdd(notConverged, 1);   // the reference prints "too many contraint iterations" and goes on
    for (int j = 0; j < na; j++)
    {
        const int s = slot[j];
        vx[s] = v[j][0];
        vy[s] = v[j][1];
        vz[s] = v[j][2];
    }
}

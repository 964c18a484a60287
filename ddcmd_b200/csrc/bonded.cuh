// bonded.cuh - all Martini bonded terms and position restraints in ONE kernel over a
// kind-sorted term list: harmonic bonds, harmonic / cosine / restricted-bending angles,
// proper torsions, harmonic impropers.
//
// Replaces charmmConvalent + connectiveEnergy + res*Sorted (src/bioCharmmCovalent.c:95-251,
// src/bioCharmmCovalentEnergies.c:266-351,754-794, src/bioCharmmCovalentEnergiesSorted.c)
// and bondedGPU.cu's seven kernels, and restraint() (src/restraint.c:259-361).
// One thread per local term evaluates it once and stages the forces on its beads; the pair kernel adds each bead's staged forces
// in the bead's fixed entry order (no atomics: bitwise reproducible).  The local terms are numbered kind by kind at every list
// build, so the lanes of a warp mostly run the same formula.
#pragma once
#include "engine.cuh"

struct V3
{
    double x, y, z;
};
__device__ __forceinline__ V3 vsub(const double4 a, const double4 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ double vdot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ V3 vcross(V3 a, V3 b) { return V3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
__device__ __forceinline__ V3 vscale(V3 a, double s) { return V3{a.x * s, a.y * s, a.z * s}; }
__device__ __forceinline__ V3 vaxpy(double s, V3 a, V3 b) { return V3{s * a.x + b.x, s * a.y + b.y, s * a.z + b.z}; }

__device__ __forceinline__ V3 minImage(V3 d, const PairConst &pc)
{
    // nearestImage == Preduce for an orthorhombic box (src/preduce.c:449-470): d -= h*rint(d/h).  The quotient only picks the
    // lattice vector (an integer), so it is taken with the reciprocal edge: three multiplications instead of three divisions
    d.x -= pc.hxx * rint(d.x * pc.ihx);
    d.y -= pc.hyy * rint(d.y * pc.ihy);
    d.z -= pc.hzz * rint(d.z * pc.ihz);
    return d;
}

// bioDihedralFast (src/bioCharmmCovalentEnergies.c:266-351): angle, sin, d(cos)/dr and the
// geometric virial factor.
__device__ __forceinline__ void dihedral(V3 vij, V3 vjk, V3 vkl, double &ang, double &sinX, V3 &dI, V3 &dJ, V3 &dK, V3 &dL, double vir[6])
{
    const double eps = 1e-12;
    const double a2 = vdot(vij, vij), b2 = vdot(vjk, vjk), c2 = vdot(vkl, vkl);
    const double ab = vdot(vij, vjk), bc = vdot(vjk, vkl), ac = vdot(vij, vkl);
    const double f = ab * bc - ac * b2;
    const double g1 = a2 * b2 - ab * ab + eps;
    const double g2 = b2 * c2 - bc * bc + eps;
    const double y = 1.0 / sqrt(g1 * g2);
    double x = y * f;
    const double xab = y * bc + x / g1 * ab;
    const double xbc = y * ab + x / g2 * bc;
    const double xac = -y * b2;
    const double xaa = -0.5 * x * b2 / g1;
    const double xcc = -0.5 * x * b2 / g2;
    const double xbb = -y * ac - 0.5 * x * (a2 / g1 + c2 / g2);
    V3 ca = vscale(vjk, xab);
    ca = vaxpy(xac, vkl, ca);
    ca = vaxpy(2.0 * xaa, vij, ca);
    V3 cb = vscale(vij, xab);
    cb = vaxpy(xbc, vkl, cb);
    cb = vaxpy(2.0 * xbb, vjk, cb);
    V3 cc = vscale(vjk, xbc);
    cc = vaxpy(xac, vij, cc);
    cc = vaxpy(2.0 * xcc, vkl, cc);
    const V3 m = vcross(vij, vjk), n = vcross(vjk, vkl), mxn = vcross(m, n);
    const double sign = (vdot(vjk, mxn) < 0.0) ? -1.0 : 1.0;
    x = fmax(fmin(x, 1.0), -1.0);
    ang = sign * acos(x);
    sinX = sin(ang);
    dI = ca;
    dJ = V3{cb.x - ca.x, cb.y - ca.y, cb.z - ca.z};
    dK = V3{cc.x - cb.x, cc.y - cb.y, cc.z - cb.z};
    dL = V3{-cc.x, -cc.y, -cc.z};
    vir[0] = -(ca.x * vij.x + cb.x * vjk.x + cc.x * vkl.x);   // xx
    vir[1] = -(ca.y * vij.y + cb.y * vjk.y + cc.y * vkl.y);   // yy
    vir[2] = -(ca.z * vij.z + cb.z * vjk.z + cc.z * vkl.z);   // zz
    vir[3] = -(ca.x * vij.y + cb.x * vjk.y + cc.x * vkl.y);   // xy
    vir[4] = -(ca.x * vij.z + cb.x * vjk.z + cc.x * vkl.z);   // xz
    vir[5] = -(ca.y * vij.z + cb.y * vjk.z + cc.y * vkl.z);   // yz
}

#define BONDED_THREADS 128
#define BONDED_ACC 11   // 0..5 virial (xx yy zz xy xz yz), 6 bond, 7 angle, 8 torsion, 9 improper, 10 restraint

// No atomics, fixed summation order: every resident local bead has the (static) list of bonded terms it takes part in - entry =
// (term << 2 | role of this bead in the term), ascending term order.  Every term is evaluated once, by the thread of its
// role-0 entry's term record, which stages the forces on all of the term's beads; a second kernel adds up each bead's
// entries from the stage in their fixed order.  Forces and energies are bitwise reproducible run to run (the reference
// accumulates in owner order too, src/bioCharmmCovalent.c:95-251).  Term records (endpoint slots + parameters) and the
// per-bead stage indices are rebuilt at every list build for the local beads only.  A term with an endpoint that is not
// resident here is skipped, and so is a bead's entry whose term has its role-0 bead on another rank (molecules are whole
// on their owner rank, src/ddcRuleMolecule.c:43, so neither happens for a Martini deck).
// ---- at every list build: the terms and the per-bead contribution lists of the resident local beads, in slot order -----------------
// cnt[s] = entries of slot s (its contributions); cntK[g * nIon + s] = the terms this bead "owns" (its role-0 entries: each term
// has exactly one) of kind group g - the local terms are numbered group by group (bonds, the three angle forms, torsions,
// impropers, restraints) and by slot inside a group, so that the threads of a warp of k_bonded evaluate the same kind of term
#define BOND_GROUPS 7
__global__ void k_bond_count(int nIon, const double4 *__restrict__ pos, const int *__restrict__ csrOff, const uint32_t *__restrict__ ent,
                             int64_t nTerms, const Term *__restrict__ terms, int *__restrict__ cnt, int *__restrict__ cntK)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nIon) return;
    const unsigned long long w = (unsigned long long)__double_as_longlong(pos[s].w);
    int n = 0, nk[BOND_GROUPS];
#pragma unroll
    for (int g = 0; g < BOND_GROUPS; g++) nk[g] = 0;
    if (!(w >> 63))
    {
        const int b = (int)((w >> 32) & 0x7fffffffull);
        const int lo = csrOff[b];
        n = csrOff[b + 1] - lo;
        for (int q = 0; q < n; q++)
        {
            const uint32_t e = ent[lo + q];
            if ((e & 3u) != 0u) continue;
            const int64_t t = (int64_t)(e >> 2);
            const int kind = t >= nTerms ? 6 : terms[t].kind;
#pragma unroll
            for (int g = 0; g < BOND_GROUPS; g++) nk[g] += (kind == g) ? 1 : 0;
        }
    }
    cnt[s] = n;
#pragma unroll
    for (int g = 0; g < BOND_GROUPS; g++) cntK[(size_t)g * nIon + s] = nk[g];
}

// exclusive scan of n ints (n = resident beads): per-block scans, a scan of the block totals by one block, then the offsets
#define SCAN_BLOCK 1024
__global__ void __launch_bounds__(SCAN_BLOCK)
k_scan_local(int n, const int *__restrict__ in, int *__restrict__ out, int *__restrict__ blockSum)
{
    __shared__ int sums[SCAN_BLOCK / 32];
    const int i = blockIdx.x * SCAN_BLOCK + threadIdx.x;
    const int v = i < n ? in[i] : 0;
    int s = v;
    for (int o = 1; o < 32; o <<= 1)
    {
        const int t = __shfl_up_sync(0xffffffffu, s, o);
        if ((int)(threadIdx.x & 31) >= o) s += t;
    }
    if ((threadIdx.x & 31) == 31) sums[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32)
    {
        int w = sums[threadIdx.x];
        for (int o = 1; o < 32; o <<= 1)
        {
            const int t = __shfl_up_sync(0xffffffffu, w, o);
            if ((int)threadIdx.x >= o) w += t;
        }
        sums[threadIdx.x] = w;
    }
    __syncthreads();
    const int before = (threadIdx.x >> 5) ? sums[(threadIdx.x >> 5) - 1] : 0;
    if (i < n) out[i] = before + s - v;
    if (threadIdx.x == SCAN_BLOCK - 1) blockSum[blockIdx.x] = before + s;
}

// one block: exclusive scan of the block totals in place (a few thousand at most), grand total to *total
__global__ void __launch_bounds__(1024)
k_scan_blocks(int nb, int *__restrict__ blockSum, int *__restrict__ total)
{
    __shared__ int sums[1024];
    const int per = (nb + blockDim.x - 1) / blockDim.x;
    const int lo = min(nb, (int)threadIdx.x * per), hi = min(nb, lo + per);
    int s = 0;
    for (int i = lo; i < hi; i++) s += blockSum[i];
    sums[threadIdx.x] = s;
    __syncthreads();
    for (int o = 1; o < (int)blockDim.x; o <<= 1)
    {
        int v = ((int)threadIdx.x >= o) ? sums[threadIdx.x - o] : 0;
        __syncthreads();
        sums[threadIdx.x] += v;
        __syncthreads();
    }
    int run = sums[threadIdx.x] - s;
    for (int i = lo; i < hi; i++)
    {
        const int v = blockSum[i];
        blockSum[i] = run;
        run += v;
    }
    if (threadIdx.x == blockDim.x - 1) *total = sums[threadIdx.x];
}

__global__ void __launch_bounds__(SCAN_BLOCK)
k_scan_add(int n, int *__restrict__ out, const int *__restrict__ blockSum)
{
    const int i = blockIdx.x * SCAN_BLOCK + threadIdx.x;
    if (i < n) out[i] += blockSum[blockIdx.x];
}

// the local terms: one record per term, written by the slot of its role-0 bead at the term's place in the kind-grouped numbering;
// termMap[t] = index of term t here
__global__ void k_bond_resolve_terms(int nIon, const double4 *__restrict__ pos, const int *__restrict__ csrOff, const uint32_t *__restrict__ ent,
                                     int64_t nTerms, const Term *__restrict__ terms, const int *__restrict__ slotOfBead,
                                     const int *__restrict__ startK, BondRec *__restrict__ recs, int *__restrict__ termMap)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nIon) return;
    const unsigned long long w = (unsigned long long)__double_as_longlong(pos[s].w);
    if (w >> 63) return;
    const int b = (int)((w >> 32) & 0x7fffffffull);
    const int elo = csrOff[b], n = csrOff[b + 1] - elo;
    int next[BOND_GROUPS];      // where this bead's next term of each kind group goes
#pragma unroll
    for (int g = 0; g < BOND_GROUPS; g++) next[g] = startK[(size_t)g * nIon + s];
    for (int q = 0; q < n; q++)
    {
        const uint32_t e = ent[elo + q];
        if ((e & 3u) != 0u) continue;
        const int64_t t = (int64_t)(e >> 2);
        const int group = t >= nTerms ? 6 : terms[t].kind;
        int lt = 0;
#pragma unroll
        for (int g = 0; g < BOND_GROUPS; g++)
            if (group == g) lt = next[g]++;
        BondRec r;
        r.role = 0;
        r.q = 0;
        if (t >= nTerms)
        {
            // restraint (src/restraint.c:287-357): its 7 parameters stay in the table, the record keeps its index in s[1]
            r.kind = 6;
            r.n = 1;
            r.s[0] = s; r.s[1] = (int)(t - nTerms); r.s[2] = 0; r.s[3] = 0;
            r.p0 = r.p1 = r.p2 = 0.0;
        }
        else
        {
            const Term tm = terms[t];
            r.s[0] = slotOfBead[tm.i];
            r.s[1] = slotOfBead[tm.j];
            r.s[2] = tm.k >= 0 ? slotOfBead[tm.k] : 0;
            r.s[3] = tm.l >= 0 ? slotOfBead[tm.l] : 0;
            r.p0 = tm.p0; r.p1 = tm.p1; r.p2 = tm.p2;
            r.kind = (short)(((r.s[0] | r.s[1] | r.s[2] | r.s[3]) < 0) ? -1 : tm.kind);      // an endpoint is not resident on this rank
            r.n = (unsigned short)(tm.kind == 0 ? 2 : (tm.kind <= 3 ? 3 : 4));
        }
        recs[lt] = r;
        termMap[t] = lt;
    }
}

// the contributions of every local bead: where the force of (term, role) is staged, in the bead's fixed entry order
__global__ void k_bond_resolve_beads(int nIon, const double4 *__restrict__ pos, const int *__restrict__ csrOff, const uint32_t *__restrict__ ent,
                                     const int *__restrict__ termMap, const int *__restrict__ start, const int *__restrict__ cnt,
                                     int *__restrict__ stageIdx)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nIon) return;
    const int n = cnt[s], lo = start[s];
    if (n == 0) return;
    const unsigned long long w = (unsigned long long)__double_as_longlong(pos[s].w);
    const int elo = csrOff[(int)((w >> 32) & 0x7fffffffull)];
    for (int q = 0; q < n; q++)
    {
        const uint32_t e = ent[elo + q];
        const int lt = termMap[e >> 2];      // -1: the term's role-0 bead is not local here (cannot happen for whole molecules)
        stageIdx[lo + q] = lt >= 0 ? 4 * lt + (int)(e & 3u) : -1;
    }
}

// forces of one term on its (up to four) beads, and its energy and virial into acc
template <bool ENERGY>
__device__ __forceinline__ void bondedEval(const BondRec &tm, const double *__restrict__ restrParm, int restrOrigin, const double4 *__restrict__ pos,
                                           const PairConst &pc, double *acc, V3 fo[4])
{
            const bool count = ENERGY;
            fo[0] = fo[1] = fo[2] = fo[3] = V3{0.0, 0.0, 0.0};
            if (tm.kind < 0) return;
            if (tm.kind == 6)
            {
                // restraint (src/restraint.c:287-357)
                const double *p = restrParm + 7 * (size_t)tm.s[1];
                double x0 = p[0] * pc.hxx, y0 = p[1] * pc.hyy, z0 = p[2] * pc.hzz;
                if (restrOrigin == 0)
                {
                    x0 -= pc.hxx / 2.0;
                    y0 -= pc.hyy / 2.0;
                    z0 -= pc.hzz / 2.0;
                }
                const double kb = p[3];
                const double4 ps = pos[tm.s[0]];
                V3 d = V3{ps.x - x0, ps.y - y0, ps.z - z0};
                if ((p[4] > 0 && fabs(d.x) > pc.hhx) || (p[5] > 0 && fabs(d.y) > pc.hhy) || (p[6] > 0 && fabs(d.z) > pc.hhz)) d = minImage(d, pc);
                const V3 cd = V3{p[4] * d.x, p[5] * d.y, p[6] * d.z};
                const V3 f = vscale(cd, -2.0 * kb);
                if (ENERGY)
                {
                    acc[10] += kb * (cd.x * d.x + cd.y * d.y + cd.z * d.z);
                    acc[0] += f.x * cd.x; acc[1] += f.y * cd.y; acc[2] += f.z * cd.z;
                    acc[3] += f.x * cd.y; acc[4] += f.x * cd.z; acc[5] += f.y * cd.z;
                }
                fo[0] = f;
                return;
            }
            const int si = tm.s[0], sj = tm.s[1], sk = tm.s[2], sl = tm.s[3];
            if (tm.kind == 0)
            {
                // resBondSorted (src/bioCharmmCovalentEnergiesSorted.c:18-116)
                const V3 b = minImage(vsub(pos[si], pos[sj]), pc);
                const double len = sqrt(vdot(b, b));
                const double dl = len - tm.p1;
                const double kf = -2.0 * tm.p0 * dl / len;
                const V3 fi = vscale(b, kf);
                fo[0] = fi;
                fo[1] = V3{-fi.x, -fi.y, -fi.z};
                if (count)
                {
                    acc[6] += tm.p0 * dl * dl;
                    acc[0] += fi.x * b.x; acc[1] += fi.y * b.y; acc[2] += fi.z * b.z;
                    acc[3] += fi.x * b.y; acc[4] += fi.x * b.z; acc[5] += fi.y * b.z;
                }
            }
            else if (tm.kind <= 3)
            {
                // resAngleSorted / resAngleCosineSorted / resAngleRestrainSorted (:118-487)
                const double4 pj = pos[sj];
                const V3 vij = minImage(vsub(pos[si], pj), pc), vkj = minImage(vsub(pos[sk], pj), pc);
                const double bij = sqrt(vdot(vij, vij)), bkj = sqrt(vdot(vkj, vkj));
                const double ibij = 1.0 / bij, ibkj = 1.0 / bkj;
                const V3 uij = vscale(vij, ibij), ukj = vscale(vkj, ibkj);
                const double c = vdot(uij, ukj);
                double coef, en;
                if (tm.kind == 1)
                {
                    const double a = acos(c), da = a - tm.p1;
                    en = tm.p0 * da * da;
                    coef = 2.0 * tm.p0 * da / sin(a);
                }
                else if (tm.kind == 2)
                {
                    const double da = c - tm.p1;
                    en = tm.p0 * da * da;
                    coef = -2.0 * tm.p0 * da;
                }
                else
                {
                    const double s2 = 1.0 - c * c, da = c - tm.p1;
                    en = tm.p0 * da * da / s2;
                    coef = -2.0 * tm.p0 * da * (1.0 - c * tm.p1) / (s2 * s2);
                }
                const double ci = coef * ibij, ck = coef * ibkj;
                const V3 fi = V3{ci * (ukj.x - uij.x * c), ci * (ukj.y - uij.y * c), ci * (ukj.z - uij.z * c)};
                const V3 fk = V3{ck * (uij.x - ukj.x * c), ck * (uij.y - ukj.y * c), ck * (uij.z - ukj.z * c)};
                fo[0] = fi;
                fo[1] = V3{-(fi.x + fk.x), -(fi.y + fk.y), -(fi.z + fk.z)};
                fo[2] = fk;
                if (count)
                {
                    acc[7] += en;
                    acc[0] += fi.x * vij.x + fk.x * vkj.x; acc[1] += fi.y * vij.y + fk.y * vkj.y; acc[2] += fi.z * vij.z + fk.z * vkj.z;
                    acc[3] += fi.x * vij.y + fk.x * vkj.y; acc[4] += fi.x * vij.z + fk.x * vkj.z; acc[5] += fi.y * vij.z + fk.y * vkj.z;
                }
            }
            else
            {
                // resTorsionSorted / resImproperSorted (:577-848)
                const double4 pI = pos[si], pJ = pos[sj], pK = pos[sk], pL = pos[sl];
                const V3 vij = minImage(vsub(pI, pJ), pc), vjk = minImage(vsub(pJ, pK), pc), vkl = minImage(vsub(pK, pL), pc);
                double ang, sinX, vir[6];
                V3 dI, dJ, dK, dL;
                dihedral(vij, vjk, vkl, ang, sinX, dI, dJ, dK, dL, vir);
                double kf, en;
                if (tm.kind == 4)
                {
                    const double kchi = tm.p0, delta = tm.p1, n = tm.p2;
                    en = kchi * (1.0 + cos(n * ang - delta));
                    if (fabs(sinX) > 1e-8) kf = kchi * n * sin(n * ang - delta) / sinX;
                    else
                    {
                        const double nX2 = (n * ang) * (n * ang), X2 = ang * ang;
                        const double num = 1 - nX2 / 6 + nX2 * nX2 / 120 - nX2 * nX2 * nX2 / 5040 + nX2 * nX2 * nX2 * nX2 / 362880 - nX2 * nX2 * nX2 * nX2 * nX2 / 39916800;
                        const double den = 1 - X2 / 6 + X2 * X2 / 120 - X2 * X2 * X2 / 5040 + X2 * X2 * X2 * X2 / 362880 - X2 * X2 * X2 * X2 * X2 / 39916800;
                        const double ratio = n * num / den;
                        kf = (delta > 3.12413936106985) ? -kchi * n * ratio : kchi * n * ratio;
                    }
                }
                else
                {
                    const double kpsi = tm.p0, psi0 = tm.p1;
                    double d = ang - psi0;
                    if (d < -M_PI) d += 2 * M_PI;
                    else if (d > M_PI) d -= 2 * M_PI;
                    en = kpsi * d * d;
                    if (fabs(sinX) > 1e-8) kf = -2.0 * kpsi * d / sinX;
                    else
                    {
                        const double X2 = ang * ang;
                        kf = -2.0 * kpsi / (1 - X2 / 6 + X2 * X2 / 120 - X2 * X2 * X2 / 5040 + X2 * X2 * X2 * X2 / 362880 - X2 * X2 * X2 * X2 * X2 / 39916800);
                    }
                }
                fo[0] = vscale(dI, -kf);
                fo[1] = vscale(dJ, -kf);
                fo[2] = vscale(dK, -kf);
                fo[3] = vscale(dL, -kf);
                if (count)
                {
                    acc[tm.kind == 4 ? 8 : 9] += en;
#pragma unroll
                    for (int a = 0; a < 6; a++) acc[a] += vir[a] * kf;
                }
            }
}

// Every term is evaluated ONCE, by one thread, which stages the forces on the term's beads (stage[4 term + role], coalesced);
// the pair kernel, which runs after this one, adds per local bead the bead's contributions in their fixed order to the bead's
// pair force before it stores it (BondAdd, pair.cuh) - no atomic, one writer per slot, the same order every run.
template <bool ENERGY, int MINB>
__global__ void __launch_bounds__(BONDED_THREADS, MINB)
k_bonded(int nTermsLocal, const BondRec *__restrict__ recs, const double *__restrict__ restrParm, int restrOrigin, const double4 *__restrict__ pos,
         PairConst pc, V3 *__restrict__ stage, double *__restrict__ partial)
{
    const int lt = blockIdx.x * blockDim.x + threadIdx.x;
    double acc[BONDED_ACC];
#pragma unroll
    for (int a = 0; a < BONDED_ACC; a++) acc[a] = 0.0;
    if (lt < nTermsLocal)
    {
        const BondRec tm = recs[lt];
        V3 fo[4];
        bondedEval<ENERGY>(tm, restrParm, restrOrigin, pos, pc, acc, fo);
        const int need = (int)tm.n;
#pragma unroll
        for (int r = 0; r < 4; r++)
            if (r < need) stage[4 * (size_t)lt + r] = fo[r];
    }
    if (ENERGY)
    {
        __shared__ double red[BONDED_ACC][BONDED_THREADS / 32];
#pragma unroll
        for (int a = 0; a < BONDED_ACC; a++)
        {
            double v = acc[a];
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if ((threadIdx.x & 31) == 0) red[a][threadIdx.x >> 5] = v;
        }
        __syncthreads();
        if (threadIdx.x < BONDED_ACC)
        {
            double v = 0.0;
            for (int w = 0; w < BONDED_THREADS / 32; w++) v += red[threadIdx.x][w];
            partial[(size_t)blockIdx.x * BONDED_ACC + threadIdx.x] = v;
        }
    }
}

// api.cu - C-ABI of the B200-native Martini MD step (include/ddcmd_b200.h): host
// orchestration of the kernels in cells.cuh / pair.cuh / bonded.cuh / integrate.cuh.
// There is no CPU path: every compute entry point requires a CUDA device.
#include "engine.cuh"
#include "cells.cuh"
#include "pair.cuh"
#include "bonded.cuh"
#include "integrate.cuh"
#include "nbrcheck.cuh"
#include "nglfcons.cuh"
#include "ddc.cuh"
#include "analysis.cuh"
#include <nccl.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <map>

static thread_local std::string g_err;
static int fail(int code, const std::string &msg)
{
    g_err = msg;
    return code;
}
#define CK(call)                                                                                         \
    do                                                                                                   \
    {                                                                                                    \
        cudaError_t e__ = (call);                                                                        \
        if (e__ != cudaSuccess)                                                                          \
            return fail(DDCB200_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));          \
    } while (0)
#define CKN(call)                                                                                        \
    do                                                                                                   \
    {                                                                                                    \
        ncclResult_t r__ = (call);                                                                       \
        if (r__ != ncclSuccess)                                                                          \
            return fail(DDCB200_ERR_NCCL, std::string(#call) + ": " + ncclGetErrorString(r__));          \
    } while (0)
#define CKL(what)                                                                                        \
    do                                                                                                   \
    {                                                                                                    \
        c->kernelLaunches += launchesOf(what);                                                           \
        cudaError_t e__ = cudaGetLastError();                                                            \
        if (e__ != cudaSuccess)                                                                          \
            return fail(DDCB200_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e__));           \
    } while (0)

// number of kernels launched before each CKL() checkpoint (for the bench's gpu_launches count)
static int launchesOf(const char *what) { return strcmp(what, "cell sort") == 0 ? 7 : 1; }   // minmax, grid, count, scan, scatter, rank, gather

extern "C" const char *ddcb200_lastError(void) { return g_err.c_str(); }

extern "C" int ddcb200_deviceCount(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess)
    {
        cudaGetLastError();
        return 0;
    }
    return n;
}

// ---- profiling helpers ----------------------------------------------------------------
struct ProfScope
{
    ddcb200_ctx *c;
    int slot;
    cudaEvent_t a = nullptr, b = nullptr;
    cudaStream_t st;
    ProfScope(ddcb200_ctx *ctx, int s, cudaStream_t stream = nullptr) : c(ctx), slot(s), st(stream ? stream : ctx->stream)
    {
        c->profLaunch[slot]++;
        if (!c->prof) return;
        auto get = [&]() {
            cudaEvent_t e;
            if (!c->evPool.empty())
            {
                e = c->evPool.back();
                c->evPool.pop_back();
            }
            else
                cudaEventCreate(&e);
            return e;
        };
        a = get();
        b = get();
        cudaEventRecord(a, st);
    }
    ~ProfScope()
    {
        if (!a) return;
        cudaEventRecord(b, st);
        c->pending.push_back({a, b, slot});
    }
};

// ---- box constants, in the reference's operation order ---------------------------------
static int setupBox(ddcb200_ctx *c)
{
    const ddcb200_params &p = c->prm;
    const double *h = p.h;
    const double eps = 1e-10;   // orthorhombicBox(), src/preduce.c:483-494
    if (fabs(h[1]) > eps || fabs(h[2]) > eps || fabs(h[3]) > eps || fabs(h[5]) > eps || fabs(h[6]) > eps || fabs(h[7]) > eps)
        return fail(DDCB200_ERR_ARG, "only ORTHORHOMBIC boxes are supported");
    if (p.pbc != 7) return fail(DDCB200_ERR_ARG, "only pbc=7 is supported");
    BoxConst &b = c->box;
    const double xx = h[0], xy = h[1], xz = h[2], yx = h[3], yy = h[4], yz = h[5], zx = h[6], zy = h[7], zz = h[8];
    // matinv, src/three_algebra.c:37-64
    const double d00 = yy * zz - yz * zy, d11 = zz * xx - zx * xz, d22 = xx * yy - xy * yx;
    const double d01 = yz * zx - yx * zz, d12 = zx * xy - zy * xx, d20 = xy * yz - xz * yy;
    const double d02 = yx * zy - zx * yy, d10 = zy * xz - xy * zz, d21 = xz * yx - yz * xx;
    const double det = xx * d00 + xy * d01 + xz * d02;
    b.hinv[0] = d00 / det; b.hinv[4] = d11 / det; b.hinv[8] = d22 / det;
    b.hinv[1] = d10 / det; b.hinv[3] = d01 / det; b.hinv[2] = d20 / det;
    b.hinv[6] = d02 / det; b.hinv[5] = d21 / det; b.hinv[7] = d12 / det;
    b.volume = det;
    b.hxx = xx; b.hyy = yy; b.hzz = zz;
    b.hhx = 0.5 * xx; b.hhy = 0.5 * yy; b.hhz = 0.5 * zz;
    b.cx = p.center[0]; b.cy = p.center[1]; b.cz = p.center[2];
    // box_get_minspan, src/box.c:213-227
    static const double lv[13][3] = {{0, 0, 1}, {0, 1, 0}, {1, 0, 0}, {0, 1, 1}, {1, 0, 1}, {1, 1, 0}, {1, 1, 1},
                                     {0, 1, -1}, {1, 0, -1}, {1, -1, 0}, {1, 1, -1}, {1, -1, 1}, {-1, 1, 0}};
    double r2min = 0;
    for (int i = 0; i < 13; i++)
    {
        const double rx = xx * lv[i][0] + xy * lv[i][1] + xz * lv[i][2];
        const double ry = yx * lv[i][0] + yy * lv[i][1] + yz * lv[i][2];
        const double rz = zx * lv[i][0] + zy * lv[i][1] + zz * lv[i][2];
        const double r2 = rx * rx + ry * ry + rz * rz;
        if (i == 0 || r2 < r2min) r2min = r2;
    }
    const double minspan = sqrt(r2min);
    b.R2cut = 0.25 * minspan * minspan;
    const double rl = p.rmax + p.deltaR;
    b.rlist2 = rl * rl;                     // src/pairlist.c:226
    b.rc2 = p.rmax * p.rmax;                // SQ(parms->rmax), src/bioMartini.c:1002
    b.rcutGeom = rl > p.minBoxSide ? rl : p.minBoxSide;
    // computeBoxSpan, src/geom.c:478-511
    {
        const double a0[3] = {xx, yx, zx}, a1[3] = {xy, yy, zy}, a2[3] = {xz, yz, zz};
        auto dot = [](const double *u, const double *v) { return u[0] * v[0] + u[1] * v[1] + u[2] * v[2]; };
        auto cross = [](const double *u, const double *v, double *w) {
            w[0] = u[1] * v[2] - u[2] * v[1];
            w[1] = u[2] * v[0] - u[0] * v[2];
            w[2] = u[0] * v[1] - u[1] * v[0];
        };
        const double l0 = sqrt(dot(a0, a0)), l1 = sqrt(dot(a1, a1)), l2 = sqrt(dot(a2, a2));
        double n0[3], n1[3], n2[3];
        cross(a1, a2, n0);
        cross(a2, a0, n1);
        cross(a0, a1, n2);
        const double s0 = 1.0 / (l1 * l2), s1 = 1.0 / (l2 * l0), s2 = 1.0 / (l0 * l1);
        for (int k = 0; k < 3; k++)
        {
            n0[k] *= s0;
            n1[k] *= s1;
            n2[k] *= s2;
        }
        const double bd[3] = {a0[0] + a1[0] + a2[0], a0[1] + a1[1] + a2[1], a0[2] + a1[2] + a2[2]};
        b.spanx = fabs(dot(n0, bd));
        b.spany = fabs(dot(n1, bd));
        b.spanz = fabs(dot(n2, bd));
    }
    // ordering bins (not part of parity): edges relative to cutoff and skin
    {
        const double *f = c->binFrac;
        for (int k = 0; k < NBINS - 1; k++)
        {
            const double r = p.rmax + f[k] * p.deltaR;
            b.binEdge2[k] = r * r;
            c->pc.binEdge[k] = r * (1.0 - 1e-12);   // r_build >= sqrt(binEdge2) >= this
        }
    }
    PairConst &pc = c->pc;
    pc.rc2 = b.rc2; pc.R2cut = b.R2cut;
    pc.hxx = b.hxx; pc.hyy = b.hyy; pc.hzz = b.hzz;
    pc.hhx = b.hhx; pc.hhy = b.hhy; pc.hhz = b.hhz;
    pc.ihx = 1.0 / b.hxx; pc.ihy = 1.0 / b.hyy; pc.ihz = 1.0 / b.hzz;
    pc.keR = p.keR; pc.krf = p.krf; pc.crf = p.crf;
    pc.rmax = p.rmax;
    pc.ntypes = c->ntypes;
    // barostat: the box may have changed since the list was built (hBuild = 0 before the first build)
    {
        const double ex = c->hBuild[0] != 0.0 ? xx - c->hBuild[0] : 0.0, ey = c->hBuild[1] != 0.0 ? yy - c->hBuild[1] : 0.0,
                     ez = c->hBuild[2] != 0.0 ? zz - c->hBuild[2] : 0.0;
        pc.listSlack = sqrt(ex * ex + ey * ey + ez * ez);
    }
    return DDCB200_OK;
}

// The k_pair2 instantiations (gathers in flight per thread, CTAs per SM the register allocation is capped for): A/B-ed on the
// B200 through DDCB200_PAIR=<pf>,<minb> | old
typedef void (*PairKernel)(int, int, const int *, int, const double4 *, const uint32_t *, const uint16_t *, const unsigned long long *, int,
                           const float *, const double2 *, const double *, const double *, PairConst, double *, double *, double *, double *,
                           const unsigned long long *, const int *, PruneArgs, BondAdd);
struct PairVariant
{
    int pf, minb, eminb;
    PairKernel force[3], energy[3];      // [MODE]: 0 plain, 1 walk + write the pruned rows, 2 walk the pruned rows
};
// the energy instantiation carries eight more accumulators: its register cap (E CTAs per SM, 1 = none) is chosen apart
#define PV(P, M, E) {P, M, E, {k_pair2<false, P, M, 0>, k_pair2<false, P, M, 1>, k_pair2<false, P, M, 2>}, \
                              {k_pair2<true, P, E, 0>, k_pair2<true, P, E, 1>, k_pair2<true, P, E, 2>}}
// measured: profiles/r02d_pair_variants.txt (force kernel), r02w_e2e_energy_caps.jsonl (energy kernel: 96 registers uncapped = 5 CTAs
// per SM; capped for 6 = 80 registers, 8 bytes spilled: the per-step energy evaluation of the end-to-end loop 12 % faster)
static const PairVariant g_pairVariants[] = {PV(1, 1, 1), PV(2, 1, 1), PV(2, 8, 6), PV(3, 8, 1), PV(4, 1, 1), PV(2, 8, 1), PV(2, 8, 5)};
#undef PV
static const int g_nPairVariants = (int)(sizeof(g_pairVariants) / sizeof(g_pairVariants[0]));

// margin of the pruned rows as a fraction of deltaR (DDCB200_PRUNE; default 1.4 x every / updateRate: a bead and its fastest
// neighbour together move about 0.07 deltaR per step at 310 K with deltaR = 4 A)
static double pruneFracOf(const ddcb200_ctx *c)
{
    // (updateRate = 0, rebuilds triggered by the displacements: the margin of a 20-step schedule)
    // (a deck that rebuilds rarely for its skin would get a margin no bead can keep: at least a fifth of the skin)
    const double frac = c->pruneMargin > 0.0 ? c->pruneMargin : std::max(0.2, 1.4 * c->pruneEvery / (c->prm.updateRate > 0 ? c->prm.updateRate : 20));
    return std::min(frac, 1.0);
}

// everything of ddcb200_create that can fail after the context exists: a failure is unwound by ddcb200_destroy
static int createInit(ddcb200_ctx *c)
{
    if (const char *pe = getenv("DDCB200_PRUNE"))
    {
        // pruned rows: <every>[,<margin>] - rewritten every <every> force evaluations from the entries closer than rmax + <margin> x deltaR
        // (default margin 1.4 x every / updateRate: a bead and its fastest neighbour move about 0.07 deltaR per step at 310 K);
        // 0 = off.  Results are bitwise those of the full walk whatever the values (pair.cuh)
        int every = 0;
        double margin = 0.0;
        const int got = sscanf(pe, "%d,%lf", &every, &margin);
        if (got < 1 || every < 0 || every > 1000 || (got == 2 && !(margin > 0.0 && margin <= 1.0)))
            return fail(DDCB200_ERR_ARG, "DDCB200_PRUNE must be <every>[,<margin as a fraction of deltaR>]");
        c->pruneEvery = every;
        c->pruneMargin = got == 2 ? margin : 0.0;
    }
    {
        // a row is two segments around one edge, rmax + 0.3 deltaR (DDCB200_NEAR): just beyond the default margin of the pruned rows, so
        // that the evaluation after a build finds all it needs in the first segment, and independent of that margin, so that the
        // order of a row (the order of the force sums) does not depend on the pruning
        double nearFrac = 0.3;
        if (const char *nf = getenv("DDCB200_NEAR"))
        {
            nearFrac = atof(nf);
            if (!(nearFrac > 0.0 && nearFrac <= 1.0)) return fail(DDCB200_ERR_ARG, "DDCB200_NEAR must be a fraction of deltaR in (0, 1]");
        }
        for (int k = 0; k < NBINS - 1; k++) c->binFrac[k] = nearFrac;
    }
    int rc = setupBox(c);
    if (rc != DDCB200_OK) return rc;
    CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    if (const char *wk = getenv("DDCB200_WALK"))
    {
        // list-walk bound rmax + d_i + d_j: "cell" (default) = this bead's own displacement + the largest in its stencil cells,
        // "bead" = own + the largest of any resident bead, "global" = twice the latter.  All exact, bitwise equal forces; the
        // looser ones walk more entries (A/B knob)
        if (strcmp(wk, "global") == 0) c->walkPerBead = c->walkPerCell = false;
        else if (strcmp(wk, "bead") == 0) c->walkPerCell = false;
        else if (strcmp(wk, "cell") != 0) return fail(DDCB200_ERR_ARG, "DDCB200_WALK must be cell, bead or global");
    }
    CK(cudaMalloc((void **)&c->grid, sizeof(GridDev)));
    CK(cudaMemset(c->grid, 0, sizeof(GridDev)));
    CK(cudaMallocHost((void **)&c->gridHost, sizeof(GridDev)));
    CK(cudaMalloc((void **)&c->acc, ACC_N * sizeof(double)));
    CK(cudaMemset(c->acc, 0, ACC_N * sizeof(double)));
    CK(cudaMallocHost((void **)&c->accHost, (ACC_N + 8) * sizeof(double)));
    CK(cudaMallocHost((void **)&c->ddcHost, 64 * sizeof(int)));
    CK(cudaMalloc((void **)&c->ddcCounters, 8 * sizeof(int)));
    CK(cudaMalloc((void **)&c->dmax2, 4 * sizeof(unsigned long long)));
    CK(cudaMemset(c->dmax2, 0, 4 * sizeof(unsigned long long)));
    if (const char *pv = getenv("DDCB200_PAIR"))
    {
        int pf = 0, mb = 0, eb = 0;
        if (sscanf(pv, "%d,%d,%d", &pf, &mb, &eb) >= 2)
        {
            // without the third number: the first built combination with that force kernel
            c->pairVariant = -2;
            for (int v = g_nPairVariants - 1; v >= 0; v--)
                if (g_pairVariants[v].pf == pf && g_pairVariants[v].minb == mb && (eb == 0 || g_pairVariants[v].eminb == eb)) c->pairVariant = v;
        }
        else c->pairVariant = -2;
        if (c->pairVariant == -2) return fail(DDCB200_ERR_ARG, "DDCB200_PAIR must be one of the built <pf>,<minb>[,<minb of the energy kernel>] combinations");
    }
    if (const char *bm = getenv("DDCB200_BONDED"))
    {
        // A/B: CTAs per SM the registers of k_bonded are capped for.  With one thread per (term, endpoint) record the 40-register
        // build (12) won: 0.073 ms against 0.115 uncapped (profiles/r02h_variants.txt).  Since a thread evaluates a whole term and
        // stages up to four forces the caps only spill: uncapped (1, the default) 0.037 ms, 8: 0.040, 12: 0.046
        // (profiles/r02aa_bonded_filter_ab.jsonl)
        c->bondedCap = atoi(bm);
        if (c->bondedCap != 1 && c->bondedCap != 8 && c->bondedCap != 12) return fail(DDCB200_ERR_ARG, "DDCB200_BONDED must be 1, 8 or 12");
    }
    if (const char *hm = getenv("DDCB200_HALO"))
    {
        // several ranks: "overlap" (default) = the ghost halo runs on its own stream beside the pair rows that read no ghost,
        // "inline" = on the compute stream before the pair kernel.  Same results (A/B knob)
        if (strcmp(hm, "inline") == 0) c->haloOverlap = false;
        else if (strcmp(hm, "overlap") != 0) return fail(DDCB200_ERR_ARG, "DDCB200_HALO must be overlap or inline");
    }
    return DDCB200_OK;
}


extern "C" int ddcb200_create(const ddcb200_params *p, ddcb200_ctx **out)
{
    if (!p || !out) return fail(DDCB200_ERR_ARG, "null argument");
    int ndev = ddcb200_deviceCount();
    if (ndev <= 0) return fail(DDCB200_ERR_NODEVICE, "no CUDA device: ddcmd_b200 has no CPU path");
    if (p->device < 0 || p->device >= ndev) return fail(DDCB200_ERR_ARG, "bad device ordinal");
    CK(cudaSetDevice(p->device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, p->device));
    if (prop.major < 10) return fail(DDCB200_ERR_NODEVICE, std::string("device ") + prop.name + " is not sm_100 class");
    ddcb200_ctx *c = new ddcb200_ctx();
    c->prm = *p;
    c->device = p->device;
    c->numSM = prop.multiProcessorCount;
    const int rc = createInit(c);
    if (rc != DDCB200_OK)
    {
        const std::string msg = g_err;      // ddcb200_destroy must not lose the reason
        ddcb200_destroy(c);
        return fail(rc, msg);
    }
    *out = c;
    return DDCB200_OK;
}

extern "C" void ddcb200_destroy(ddcb200_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    c->ljTab.release(); c->shiftTab.release(); c->qTab.release(); c->massOfBead.release(); c->wOfBead.release();
    c->gidOfBead.release(); c->molTypeOfBead.release(); c->molTypeSingle.release(); c->bpairOffset.release();
    c->bpairKey.release(); c->termsBead.release(); c->restrBead.release(); c->bondCsrOff.release(); c->bondEnt.release(); c->bondRec.release(); c->bondCount.release(); c->bondStart.release(); c->scanBlocks.release(); c->bondCount0.release(); c->bondStart0.release();
    c->termMap.release(); c->bondStageIdx.release(); c->bondStage.release();
    c->restrParm.release(); c->molOffset.release(); c->molBeads.release();
    for (int k = 0; k < 2; k++)
    {
        c->pos4[k].release();
        c->beadOfSlot[k].release();
        for (int a = 0; a < 3; a++) c->vel[k][a].release();
    }
    for (int a = 0; a < 3; a++) c->frc[a].release();
    c->slotOfBead.release(); c->cellOfSlot[0].release(); c->cellOfSlot[1].release(); c->rank0.release(); c->cellCount.release();
    c->cellStart.release(); c->member.release(); c->perm.release(); c->mmPartial.release(); c->nbrRaw.release();
    c->nbr.release(); c->nbrCount.release(); c->pairPartial.release(); c->bondPartial.release(); c->kinPartial.release();
    c->colMap.release(); c->stage.release(); c->stageI.release();
    c->orderKey.release(); c->pos32.release(); c->nbrRawCount.release(); c->nbrCum.release();
    for (int a = 0; a < 3; a++) c->posBuild[a].release();
    for (int a = 0; a < 3; a++) c->posCheck[a].release();
    c->dispOfSlot.release(); c->pruneCount.release();
    if (c->dmax2) cudaFree(c->dmax2);
    c->groupOfBead.release(); c->rngState.release(); c->rngMP.release(); c->consAtomOff.release(); c->consAtomBead.release();
    c->consPairOff.release(); c->consPairA.release(); c->consPairB.release(); c->consPairDist.release();
    if (c->consFlag) cudaFree(c->consFlag);
    c->chk.release(); c->chkPartial.release();
    if (c->chkHost) cudaFreeHost(c->chkHost);
    c->ownerBead.release(); c->ddcDest.release(); c->ddcMask.release();
    c->ddcList.release(); c->sendSlot.release(); c->recvSlot.release();
    c->sendBuf.release(); c->recvBuf.release(); c->accG.release();
    if (c->streamH) { cudaStreamSynchronize(c->streamH); cudaStreamDestroy(c->streamH); }
    if (c->streamB) { cudaStreamSynchronize(c->streamB); cudaStreamDestroy(c->streamB); }
    if (c->evBoundary) cudaEventDestroy(c->evBoundary);
    if (c->evBonded) cudaEventDestroy(c->evBonded);
    if (c->evPos) cudaEventDestroy(c->evPos);
    if (c->evHalo) cudaEventDestroy(c->evHalo);
    c->tileGhost.release(); c->tileOrder.release(); c->cellDmax.release(); c->nbrDmax.release();
    if (c->ddcWork) cudaFree(c->ddcWork);
    if (c->ddcWorkInit) cudaFreeHost(c->ddcWorkInit);
    if (c->ddcRow) cudaFree(c->ddcRow);
    if (c->ddcRowAll) cudaFree(c->ddcRowAll);
    if (c->ddcRowHost) cudaFreeHost(c->ddcRowHost);
    if (c->ddcBox6) cudaFree(c->ddcBox6);
    if (c->ddcBoxAll) cudaFree(c->ddcBoxAll);
    if (c->boxes) cudaFree(c->boxes);
    if (c->ddcCounters) cudaFree(c->ddcCounters);
    if (c->ddcHost) cudaFreeHost(c->ddcHost);
    if (c->nccl) ncclCommDestroy((ncclComm_t)c->nccl);
    for (auto &pe : c->pending)
    {
        cudaEventDestroy(pe.a);
        cudaEventDestroy(pe.b);
    }
    for (auto e : c->evPool) cudaEventDestroy(e);
    for (int k = 0; k < 2; k++)
        if (c->evList[k]) cudaEventDestroy(c->evList[k]);
    for (int k = 0; k < 4; k++)
        if (c->timer[k]) cudaEventDestroy(c->timer[k]);
    if (c->grid) cudaFree(c->grid);
    if (c->gridHost) cudaFreeHost(c->gridHost);
    if (c->acc) cudaFree(c->acc);
    if (c->accHost) cudaFreeHost(c->accHost);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

extern "C" int ddcb200_sync(ddcb200_ctx *c)
{
    if (!c) return fail(DDCB200_ERR_ARG, "null ctx");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    return DDCB200_OK;
}

static size_t pairSmemBytes(int ntypes) { return (size_t)ntypes * ntypes * (sizeof(double2) + sizeof(double)) + 256 * sizeof(double); }

extern "C" int ddcb200_martiniNonBondParms(ddcb200_ctx *c, int ntypes, const double *eps, const double *sigma, const double *shift)
{
    if (!c || ntypes <= 0 || ntypes > 255 || !eps || !sigma || !shift) return fail(DDCB200_ERR_ARG, "bad LJ table");
    CK(cudaSetDevice(c->device));
    std::vector<double2> t((size_t)ntypes * ntypes);
    for (int k = 0; k < ntypes * ntypes; k++)
    {
        const double s2 = sigma[k] * sigma[k], s6 = s2 * s2 * s2;
        t[k].x = 4.0 * eps[k] * s6;
        t[k].y = 4.0 * eps[k] * s6 * s6;
    }
    CK(c->ljTab.ensure(t.size()));
    CK(c->shiftTab.ensure(t.size()));
    CK(cudaMemcpy(c->ljTab.p, t.data(), t.size() * sizeof(double2), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c->shiftTab.p, shift, t.size() * sizeof(double), cudaMemcpyHostToDevice));
    // k_pair keeps both tables and the 256 charges in dynamic shared memory: opt in beyond the 48 KB default, refuse what cannot fit
    const size_t smem = pairSmemBytes(ntypes);
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, c->device));
    if (smem + 1024 > prop.sharedMemPerBlockOptin)
        return fail(DDCB200_ERR_CAPACITY, "too many LJ atom types for the shared-memory tables of the pair kernel (about 96 at most)");
    for (int v = 0; v < g_nPairVariants; v++)
    {
        for (int m = 0; m < 3; m++)
        {
            CK(cudaFuncSetAttribute(g_pairVariants[v].force[m], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            CK(cudaFuncSetAttribute(g_pairVariants[v].energy[m], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        }
    }
    c->ntypes = ntypes;
    c->pc.ntypes = ntypes;
    return DDCB200_OK;
}

extern "C" int ddcb200_setSpecies(ddcb200_ctx *c, int nspecies, const int *ljType, const double *charge, const double *mass)
{
    if (!c || nspecies <= 0 || !ljType || !charge || !mass) return fail(DDCB200_ERR_ARG, "bad species table");
    c->nspecies = nspecies;
    c->hSpecLJ.assign(ljType, ljType + nspecies);
    c->hSpecQ.assign(charge, charge + nspecies);
    c->hSpecMass.assign(mass, mass + nspecies);
    c->hQtab.clear();
    c->hSpecQi.resize(nspecies);
    for (int s = 0; s < nspecies; s++)
    {
        if (ljType[s] < 0 || ljType[s] > 254) return fail(DDCB200_ERR_ARG, "LJ type out of range");
        size_t k = 0;
        for (; k < c->hQtab.size(); k++)
            if (c->hQtab[k] == charge[s]) break;
        if (k == c->hQtab.size()) c->hQtab.push_back(charge[s]);
        if (k > 255) return fail(DDCB200_ERR_CAPACITY, "more than 256 distinct charges");
        c->hSpecQi[s] = (int)k;
    }
    CK(cudaSetDevice(c->device));
    std::vector<double> q(256, 0.0);
    std::copy(c->hQtab.begin(), c->hQtab.end(), q.begin());
    CK(c->qTab.ensure(256));
    CK(cudaMemcpy(c->qTab.p, q.data(), 256 * sizeof(double), cudaMemcpyHostToDevice));
    return DDCB200_OK;
}

static int uploadBeadTables(ddcb200_ctx *c)
{
    const int64_t n = c->nGlobal;
    std::vector<uint64_t> w((size_t)n);
    std::vector<double> m((size_t)n);
    std::vector<int> mt((size_t)n, -1);
    for (int64_t i = 0; i < n; i++)
    {
        const int s = c->hSpecies[(size_t)i];
        if (s < 0 || s >= c->nspecies) return fail(DDCB200_ERR_ARG, "species index out of range");
        w[(size_t)i] = packW(c->hSpecLJ[s], c->hSpecQi[s], (uint32_t)(c->hGid[(size_t)i] >> 32), (uint32_t)i);
        m[(size_t)i] = c->hSpecMass[s];
        if (!c->hMolTypeOfSpecies.empty()) mt[(size_t)i] = c->hMolTypeOfSpecies[s];
    }
    CK(c->wOfBead.ensure((size_t)n));
    CK(c->massOfBead.ensure((size_t)n));
    CK(c->gidOfBead.ensure((size_t)n));
    CK(c->molTypeOfBead.ensure((size_t)n));
    CK(c->slotOfBead.ensure((size_t)n));
    CK(cudaMemcpy(c->wOfBead.p, w.data(), n * sizeof(uint64_t), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c->massOfBead.p, m.data(), n * sizeof(double), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c->gidOfBead.p, c->hGid.data(), n * sizeof(uint64_t), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c->molTypeOfBead.p, mt.data(), n * sizeof(int), cudaMemcpyHostToDevice));
    CK(cudaMemset(c->slotOfBead.p, 0xff, n * sizeof(int)));
    return DDCB200_OK;
}

extern "C" int ddcb200_setBeads(ddcb200_ctx *c, int64_t nGlobal, const uint64_t *gid, const int *species)
{
    if (!c || nGlobal <= 0 || nGlobal >= (1ll << 27) || !gid || !species) return fail(DDCB200_ERR_ARG, "bad bead table");
    if (c->nspecies == 0) return fail(DDCB200_ERR_STATE, "setSpecies must precede setBeads");
    CK(cudaSetDevice(c->device));
    c->nGlobal = nGlobal;
    c->hGid.assign(gid, gid + nGlobal);
    c->hSpecies.assign(species, species + nGlobal);
    c->bondCsrDirty = true;
    return uploadBeadTables(c);
}

extern "C" int ddcb200_setExclusions(ddcb200_ctx *c, int nMolTypes, const int *molTypeOfSpecies, const int *molTypeNSpecies,
                                     const int *bpairOffset, const int *bpairI, const int *bpairJ)
{
    if (!c || nMolTypes <= 0 || !molTypeOfSpecies || !molTypeNSpecies || !bpairOffset) return fail(DDCB200_ERR_ARG, "bad exclusion tables");
    if (c->nspecies == 0) return fail(DDCB200_ERR_STATE, "setSpecies must precede setExclusions");
    CK(cudaSetDevice(c->device));
    c->nMolTypes = nMolTypes;
    c->hMolTypeOfSpecies.assign(molTypeOfSpecies, molTypeOfSpecies + c->nspecies);
    c->hMolTypeNSpecies.assign(molTypeNSpecies, molTypeNSpecies + nMolTypes);
    std::vector<int> single(nMolTypes), off(nMolTypes + 1, 0);
    std::vector<uint32_t> keys;
    for (int m = 0; m < nMolTypes; m++)
    {
        single[m] = molTypeNSpecies[m] <= 1 ? 1 : 0;
        std::vector<uint32_t> k;
        for (int e = bpairOffset[m]; e < bpairOffset[m + 1]; e++)
        {
            const uint32_t a = (uint32_t)bpairI[e] & 0xffffu, b = (uint32_t)bpairJ[e] & 0xffffu;
            k.push_back((std::min(a, b) << 16) | std::max(a, b));
        }
        std::sort(k.begin(), k.end());
        k.erase(std::unique(k.begin(), k.end()), k.end());
        off[m] = (int)keys.size();
        keys.insert(keys.end(), k.begin(), k.end());
    }
    off[nMolTypes] = (int)keys.size();
    CK(c->molTypeSingle.ensure(nMolTypes));
    CK(c->bpairOffset.ensure(nMolTypes + 1));
    CK(c->bpairKey.ensure(keys.size() + 1));
    CK(cudaMemcpy(c->molTypeSingle.p, single.data(), nMolTypes * sizeof(int), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c->bpairOffset.p, off.data(), (nMolTypes + 1) * sizeof(int), cudaMemcpyHostToDevice));
    if (!keys.empty()) CK(cudaMemcpy(c->bpairKey.p, keys.data(), keys.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    c->haveExcl = true;
    if (c->nGlobal > 0) return uploadBeadTables(c);
    return DDCB200_OK;
}

extern "C" int ddcb200_martiniBondParms(ddcb200_ctx *c, int64_t nTerms, const int *kind, const int *idx, const double *parm)
{
    if (!c || nTerms < 0 || (nTerms > 0 && (!kind || !idx || !parm))) return fail(DDCB200_ERR_ARG, "bad bonded terms");
    CK(cudaSetDevice(c->device));
    std::vector<Term> t((size_t)nTerms);
    for (int64_t k = 0; k < nTerms; k++)
    {
        Term &tm = t[(size_t)k];
        tm.i = idx[4 * k]; tm.j = idx[4 * k + 1]; tm.k = idx[4 * k + 2]; tm.l = idx[4 * k + 3];
        tm.p0 = parm[3 * k]; tm.p1 = parm[3 * k + 1]; tm.p2 = parm[3 * k + 2];
        tm.kind = kind[k]; tm.pad = 0;
        if (tm.kind < 0 || tm.kind > 5) return fail(DDCB200_ERR_ARG, "unknown bonded term kind");
        const int need = tm.kind == 0 ? 2 : (tm.kind <= 3 ? 3 : 4);
        const int ids[4] = {tm.i, tm.j, tm.k, tm.l};
        for (int a = 0; a < need; a++)
            if (ids[a] < 0 || (c->nGlobal > 0 && ids[a] >= c->nGlobal)) return fail(DDCB200_ERR_ARG, "bonded term bead index out of range");
        if (need < 3) tm.k = -1;
        if (need < 4) tm.l = -1;
    }
    // one formula per warp: sort by kind (stable, keeps the caller's order inside a kind)
    std::stable_sort(t.begin(), t.end(), [](const Term &a, const Term &b) { return a.kind < b.kind; });
    if (nTerms >= (1ll << 29)) return fail(DDCB200_ERR_CAPACITY, "too many bonded terms");
    c->nTerms = nTerms;
    CK(c->termsBead.ensure((size_t)nTerms + 1));
    if (nTerms) CK(cudaMemcpy(c->termsBead.p, t.data(), nTerms * sizeof(Term), cudaMemcpyHostToDevice));
    c->hTerms.swap(t);
    c->bondCsrDirty = true;
    c->listValid = false;
    return DDCB200_OK;
}

extern "C" int ddcb200_setRestraints(ddcb200_ctx *c, int64_t n, const int *bead, const double *frac0, const double *kb,
                                     const double *fc, int origin)
{
    if (!c || n < 0 || (n > 0 && (!bead || !frac0 || !kb || !fc))) return fail(DDCB200_ERR_ARG, "bad restraints");
    CK(cudaSetDevice(c->device));
    std::vector<double> p((size_t)n * 7);
    for (int64_t k = 0; k < n; k++)
    {
        for (int a = 0; a < 3; a++)
        {
            p[7 * k + a] = frac0[3 * k + a];
            p[7 * k + 4 + a] = fc[3 * k + a];
        }
        p[7 * k + 3] = kb[k];
    }
    c->nRestr = n;
    c->restrOrigin = origin;
    for (int64_t k = 0; k < n; k++)
        if (bead[k] < 0 || (c->nGlobal > 0 && bead[k] >= c->nGlobal)) return fail(DDCB200_ERR_ARG, "restraint bead index out of range");
    c->hRestrBead.assign(bead, bead + n);
    c->bondCsrDirty = true;
    CK(c->restrBead.ensure((size_t)n + 1));
    CK(c->restrParm.ensure((size_t)n * 7 + 1));
    if (n)
    {
        CK(cudaMemcpy(c->restrBead.p, bead, n * sizeof(int), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(c->restrParm.p, p.data(), n * 7 * sizeof(double), cudaMemcpyHostToDevice));
    }
    c->listValid = false;
    return DDCB200_OK;
}

extern "C" int ddcb200_setMolecules(ddcb200_ctx *c, int64_t nMol, const int64_t *molOffset, const int *molBeads, int64_t nMolTotal)
{
    if (!c || nMol < 0 || (nMol > 0 && (!molOffset || !molBeads))) return fail(DDCB200_ERR_ARG, "bad molecule table");
    CK(cudaSetDevice(c->device));
    c->nMol = nMol;
    c->nMolTotal = nMolTotal;
    c->hMolOffset.assign(molOffset, molOffset + (nMol ? nMol + 1 : 0));
    c->hMolBeads.assign(molBeads, molBeads + (nMol ? molOffset[nMol] : 0));
    c->ownerBeadValid = false;
    if (nMol)
    {
        CK(c->molOffset.ensure((size_t)nMol + 1));
        CK(c->molBeads.ensure((size_t)molOffset[nMol] + 1));
        CK(cudaMemcpy(c->molOffset.p, molOffset, (nMol + 1) * sizeof(int64_t), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(c->molBeads.p, molBeads, molOffset[nMol] * sizeof(int), cudaMemcpyHostToDevice));
    }
    return DDCB200_OK;
}

static int64_t padTile(int64_t n) { return ((n + TILE - 1) / TILE) * TILE; }

// the slot arrays of one buffer set; keep = what they hold must survive a reallocation
static int ensureSlots(ddcb200_ctx *c, int k, int64_t nIon, bool keep)
{
    const size_t nPad = (size_t)padTile(nIon);
    if (keep)
    {
        CK(c->pos4[k].grow(nPad, c->stream));
        CK(c->beadOfSlot[k].grow(nPad, c->stream));
        for (int a = 0; a < 3; a++) CK(c->vel[k][a].grow(nPad, c->stream));
    }
    else
    {
        CK(c->pos4[k].ensure(nPad));
        CK(c->beadOfSlot[k].ensure(nPad));
        for (int a = 0; a < 3; a++) CK(c->vel[k][a].ensure(nPad));
    }
    CK(c->cellOfSlot[k].ensure(nPad));
    return DDCB200_OK;
}

// everything else that is sized by the resident bead count and rebuilt by the list build
static int ensureAux(ddcb200_ctx *c, int64_t nIon)
{
    const int64_t nPad = padTile(nIon);
    for (int a = 0; a < 3; a++) CK(c->frc[a].ensure((size_t)nPad));
    CK(c->rank0.ensure((size_t)nPad));
    CK(c->member.ensure((size_t)nPad));
    CK(c->perm.ensure((size_t)nPad));
    CK(c->cellCount.ensure(2 * (size_t)nIon + 16));      // local cells, then ghost cells
    CK(c->cellStart.ensure(2 * (size_t)nIon + 16));
    CK(c->nbrCount.ensure((size_t)nPad));
    CK(c->nbrRawCount.ensure((size_t)nPad));
    CK(c->nbrCum.ensure((size_t)nPad * NBINS));
    CK(c->orderKey.ensure((size_t)nPad));
    CK(c->pos32.ensure((size_t)nPad));
    for (int a = 0; a < 3; a++) CK(c->posBuild[a].ensure((size_t)nPad));
    CK(c->mmPartial.ensure(6 * 1024));
    const size_t tiles = (size_t)(nPad / TILE);
    CK(c->pairPartial.ensure(tiles * 8 + 8));
    CK(c->kinPartial.ensure(tiles * 7 + 8));
    c->nPad = nPad;
    return DDCB200_OK;
}

static int ensureState(ddcb200_ctx *c, int64_t nIon)
{
    for (int k = 0; k < 2; k++)
    {
        const int rc = ensureSlots(c, k, nIon, false);
        if (rc) return rc;
    }
    return ensureAux(c, nIon);
}

extern "C" int ddcb200_sendState(ddcb200_ctx *c, int64_t nLocal, const int *bead, const double *rx, const double *ry, const double *rz,
                                 const double *vx, const double *vy, const double *vz, int64_t loop, double time)
{
    if (!c || nLocal <= 0 || !rx || !ry || !rz || !vx || !vy || !vz) return fail(DDCB200_ERR_ARG, "bad state");
    if (c->nGlobal == 0) return fail(DDCB200_ERR_STATE, "setBeads must precede sendState");
    if (nLocal > c->nGlobal) return fail(DDCB200_ERR_ARG, "nLocal exceeds the bead table");
    CK(cudaSetDevice(c->device));
    int rc = ensureState(c, nLocal);
    if (rc) return rc;
    c->nLocal = nLocal;
    c->nIon = nLocal;
    c->hLocalBeads.resize((size_t)nLocal);
    for (int64_t i = 0; i < nLocal; i++) c->hLocalBeads[(size_t)i] = bead ? bead[i] : (int)i;
    CK(c->stage.ensure((size_t)nLocal * 9));
    CK(c->stageI.ensure((size_t)nLocal));
    const double *src[6] = {rx, ry, rz, vx, vy, vz};
    for (int a = 0; a < 6; a++)
        CK(cudaMemcpyAsync(c->stage.p + (size_t)a * nLocal, src[a], nLocal * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->stageI.p, c->hLocalBeads.data(), nLocal * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemsetAsync(c->slotOfBead.p, 0xff, c->nGlobal * sizeof(int), c->stream));
    const int cur = c->cur;
    const int blocks = (int)((nLocal + 255) / 256);
    const double *s = c->stage.p;
    LAUNCH(k_upload_state, blocks, 256, 0, c->stream)((int)nLocal, c->stageI.p, s, s + nLocal, s + 2 * nLocal, s + 3 * nLocal, s + 4 * nLocal,
                                                  s + 5 * nLocal, c->wOfBead.p, c->pos4[cur].p, c->vel[cur][0].p, c->vel[cur][1].p,
                                                  c->vel[cur][2].p, c->beadOfSlot[cur].p, c->slotOfBead.p);
    CKL("k_upload_state");
    c->loop = loop;
    c->time = time;
    c->listValid = false;
    c->nCellsBuilt = 0;       // the cells of the last build say nothing about the new slots
    c->forcesValid = false;
    c->energyValid = false;
    c->haloDirty = false;
    c->localsDirty = false;
    return DDCB200_OK;
}

// New positions / velocities for the beads that are already resident, e.g. from a host-side integrator between two force
// evaluations (the reference's per-step sendPosnToGPU, src/gpuMemUtils.h): unlike sendState the cell order and the neighbor list
// are kept, so the list is rebuilt on the DDC schedule exactly as when the whole step runs on the device.
extern "C" int ddcb200_updateState(ddcb200_ctx *c, int64_t nLocal, const int *bead, const double *rx, const double *ry, const double *rz,
                                   const double *vx, const double *vy, const double *vz, int64_t loop, double time)
{
    if (!c || nLocal <= 0 || !rx || !ry || !rz || !vx || !vy || !vz) return fail(DDCB200_ERR_ARG, "bad state");
    if (!c->listValid || c->nranks > 1 || nLocal != c->nLocal) return ddcb200_sendState(c, nLocal, bead, rx, ry, rz, vx, vy, vz, loop, time);
    CK(cudaSetDevice(c->device));
    CK(c->stage.ensure((size_t)nLocal * 9));
    CK(c->stageI.ensure((size_t)nLocal));
    const double *src[6] = {rx, ry, rz, vx, vy, vz};
    for (int a = 0; a < 6; a++)
        CK(cudaMemcpyAsync(c->stage.p + (size_t)a * nLocal, src[a], nLocal * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    if (bead)
    {
        c->hLocalBeads.assign(bead, bead + nLocal);
        CK(cudaMemcpyAsync(c->stageI.p, c->hLocalBeads.data(), nLocal * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    }
    else
    {
        // the identity order replaces whatever bead list the previous upload used: getState reads through stageI
        for (int64_t i = 0; i < nLocal; i++) c->hLocalBeads[(size_t)i] = (int)i;
        CK(cudaMemcpyAsync(c->stageI.p, c->hLocalBeads.data(), nLocal * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    }
    const int cur = c->cur;
    const double *s = c->stage.p;
    LAUNCH(k_update_state, (int)((nLocal + 255) / 256), 256, 0, c->stream)((int)nLocal, bead ? c->stageI.p : nullptr, c->slotOfBead.p, s, s + nLocal,
                                                                       s + 2 * nLocal, s + 3 * nLocal, s + 4 * nLocal, s + 5 * nLocal, c->pos4[cur].p,
                                                                       c->vel[cur][0].p, c->vel[cur][1].p, c->vel[cur][2].p, c->pc, c->posBuild[0].p,
                                                                       c->posBuild[1].p, c->posBuild[2].p, c->dmax2, c->dispOfSlot.p,
                                                                       c->nCellsBuilt > 0 ? c->cellDmax.p : nullptr, c->cellOfSlot[cur].p);
    CKL("k_update_state");
    c->movedSinceRef = true;
    c->loop = loop;
    c->time = time;
    c->forcesValid = false;
    c->energyValid = false;
    c->kineticValid = false;
    c->pendingKick2 = false;
    return DDCB200_OK;
}

// After a re-domain the set of local beads has changed: list them (slot order is not meaningful to callers).
static int refreshLocals(ddcb200_ctx *c)
{
    if (!c->localsDirty) return DDCB200_OK;
    cudaStream_t st = c->stream;
    CK(c->stageI.ensure((size_t)c->nIon));
    CK(cudaMemsetAsync(c->ddcCounters + 4, 0, sizeof(int), st));
    LAUNCH(k_ddc_list_locals, (int)((c->nIon + 255) / 256), 256, 0, st)((int)c->nIon, c->pos4[c->cur].p, c->ddcCounters + 4, c->stageI.p);
    CKL("k_ddc_list_locals");
    CK(cudaMemcpyAsync(c->ddcHost + 4, c->ddcCounters + 4, sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    c->nLocal = c->ddcHost[4];
    c->hLocalBeads.resize((size_t)c->nLocal);
    if (c->nLocal) CK(cudaMemcpy(c->hLocalBeads.data(), c->stageI.p, c->nLocal * sizeof(int), cudaMemcpyDeviceToHost));
    c->localsDirty = false;
    return DDCB200_OK;
}

extern "C" int64_t ddcb200_numLocal(ddcb200_ctx *c)
{
    if (!c) return 0;
    if (c->localsDirty && (cudaSetDevice(c->device) != cudaSuccess || refreshLocals(c) != DDCB200_OK)) return -1;
    return c->nLocal;
}

extern "C" int ddcb200_getLocalBeads(ddcb200_ctx *c, int *bead)
{
    if (!c || !bead) return fail(DDCB200_ERR_ARG, "null argument");
    CK(cudaSetDevice(c->device));
    int rcl = refreshLocals(c);
    if (rcl) return rcl;
    std::copy(c->hLocalBeads.begin(), c->hLocalBeads.end(), bead);
    return DDCB200_OK;
}

extern "C" int ddcb200_getState(ddcb200_ctx *c, double *rx, double *ry, double *rz, double *vx, double *vy, double *vz,
                                double *fx, double *fy, double *fz)
{
    if (!c) return fail(DDCB200_ERR_ARG, "null ctx");
    if (c->nLocal == 0) return fail(DDCB200_ERR_STATE, "no state");
    CK(cudaSetDevice(c->device));
    int rcl = refreshLocals(c);
    if (rcl) return rcl;
    const int64_t n = c->nLocal;
    const int cur = c->cur;
    CK(c->stage.ensure((size_t)n * 9));
    const int blocks = (int)((n + 255) / 256);
    LAUNCH(k_download_state, blocks, 256, 0, c->stream)((int)n, c->stageI.p, c->slotOfBead.p, c->pos4[cur].p, c->vel[cur][0].p,
                                                    c->vel[cur][1].p, c->vel[cur][2].p, c->frc[0].p, c->frc[1].p, c->frc[2].p, c->stage.p);
    CKL("k_download_state");
    double *dst[9] = {rx, ry, rz, vx, vy, vz, fx, fy, fz};
    for (int a = 0; a < 9; a++)
        if (dst[a]) CK(cudaMemcpyAsync(dst[a], c->stage.p + (size_t)a * n, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return DDCB200_OK;
}

// ---- multi-GPU: re-domain and halo (ddc.cuh) ----------------------------------------------
static DdcGeom ddcGeomOf(const ddcb200_ctx *c)
{
    DdcGeom g;
    g.nranks = c->nranks;
    g.me = c->rank;
    for (int a = 0; a < 3; a++) g.lat[a] = c->lat[a];
    g.L[0] = c->box.hxx; g.L[1] = c->box.hyy; g.L[2] = c->box.hzz;
    for (int a = 0; a < 3; a++) g.hL[a] = 0.5 * g.L[a];
    g.rlist2 = c->box.rlist2 * (1.0 + 1e-9);
    return g;
}

static void buildOwnerBead(const std::vector<int64_t> &molOffset, const std::vector<int> &molBeads, int64_t nGlobal, std::vector<int> &ob)
{
    // ddcRuleMolecule / bioMartiniRule: a molecule follows its ownership bead (first listed bead); beads of
    // unlisted (single-bead) molecules own themselves
    ob.resize((size_t)nGlobal);
    for (int64_t b = 0; b < nGlobal; b++) ob[(size_t)b] = (int)b;
    const int64_t nMol = molOffset.empty() ? 0 : (int64_t)molOffset.size() - 1;
    for (int64_t m = 0; m < nMol; m++)
        for (int64_t k = molOffset[m]; k < molOffset[m + 1]; k++) ob[(size_t)molBeads[(size_t)k]] = molBeads[(size_t)molOffset[m]];
}

// every rank's count row (DDC_ROW ints) to every rank, then to pinned host memory; returns with the stream idle
static int gatherRows(ddcb200_ctx *c, int phase, int nLocal, double *box6)
{
    cudaStream_t st = c->stream;
    LAUNCH(k_rd_row, 1, DDC_ROW, 0, st)((const DdcWork *)c->ddcWork, phase, nLocal, c->ddcRow, box6);
    CKL("k_rd_row");
    CKN(ncclAllGather(c->ddcRow, c->ddcRowAll, DDC_ROW, ncclInt32, (ncclComm_t)c->nccl, st));
    CK(cudaMemcpyAsync(c->ddcRowHost, c->ddcRowAll, (size_t)c->nranks * DDC_ROW * sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return DDCB200_OK;
}

// grouped point-to-point exchange of `rec` doubles per item (ddcSendRecvTables' MPI_Isend/Irecv pairs, src/ddcSendRecv.c:216-262)
static int exchange(ddcb200_ctx *c, int rec, const std::vector<int> &sendCount, const std::vector<int> &sendOff, const std::vector<int> &recvCount,
                    const std::vector<int> &recvOff)
{
    ncclComm_t comm = (ncclComm_t)c->nccl;
    CKN(ncclGroupStart());
    for (int pp = 0; pp < c->nranks; pp++)
    {
        if (pp == c->rank) continue;
        if (sendCount[pp]) CKN(ncclSend(c->sendBuf.p + (size_t)rec * sendOff[pp], (size_t)rec * sendCount[pp], ncclDouble, pp, comm, c->stream));
        if (recvCount[pp]) CKN(ncclRecv(c->recvBuf.p + (size_t)rec * recvOff[pp], (size_t)rec * recvCount[pp], ncclDouble, pp, comm, c->stream));
    }
    CKN(ncclGroupEnd());
    return DDCB200_OK;
}

// ddcAssignment + ddcSendRecvTables (src/ddcAssignment.c:64-107, src/ddcSendRecv.c:41-277): see ddc.cuh
static int redomain(ddcb200_ctx *c)
{
    cudaStream_t st = c->stream;
    const DdcGeom g = ddcGeomOf(c);
    const int nr = c->nranks, me = c->rank;
    const int cur = c->cur, nxt = cur ^ 1;
    const int nIonOld = (int)c->nIon, nLocalOld = (int)c->nLocal;
    if (!c->ownerBeadValid)
    {
        std::vector<int> ob;
        buildOwnerBead(c->hMolOffset, c->hMolBeads, c->nGlobal, ob);
        CK(c->ownerBead.ensure((size_t)c->nGlobal));
        CK(cudaMemcpy(c->ownerBead.p, ob.data(), c->nGlobal * sizeof(int), cudaMemcpyHostToDevice));
        c->ownerBeadValid = true;
    }
    DdcWork *work = (DdcWork *)c->ddcWork;
    const int nbOld = (nIonOld + 255) / 256;
    // ---- A. migration ----
    CK(c->ddcDest.ensure((size_t)nIonOld + 1));
    CK(cudaMemcpyAsync(work, c->ddcWorkInit, sizeof(DdcWork), cudaMemcpyHostToDevice, st));
    LAUNCH(k_rd_dest, nbOld, 256, 0, st)(nIonOld, c->pos4[cur].p, c->slotOfBead.p, c->ownerBead.p, g, c->ddcDest.p, work);
    CKL("k_rd_dest");
    int rc = gatherRows(c, 0, nLocalOld, nullptr);
    if (rc) return rc;
    const int *rows = c->ddcRowHost;
    std::vector<int> sendCount(nr, 0), recvCount(nr, 0), sendOff(nr, 0), recvOff(nr, 0);
    int nSend = 0, nRecv = 0;
    for (int r = 0; r < nr; r++)
    {
        // every rank sees every row, so every rank takes the same decision here (nobody is left waiting in a collective)
        if (rows[r * DDC_ROW + 16] & 1)
            return fail(DDCB200_ERR_STATE, "a molecule is split between ranks: the beads given to sendState on one rank must be whole molecules (ddcRuleMolecule)");
        int64_t nNew = rows[r * DDC_ROW + 17];
        for (int q = 0; q < nr; q++) nNew += rows[q * DDC_ROW + r] - rows[r * DDC_ROW + q];
        if (nNew <= 0) return fail(DDCB200_ERR_STATE, "a rank owns no beads after the domain assignment (fewer bricks than the system can fill)");
    }
    for (int q = 0; q < nr; q++)
    {
        if (q == me) continue;
        sendCount[q] = rows[me * DDC_ROW + q];
        recvCount[q] = rows[q * DDC_ROW + me];
        sendOff[q] = nSend;
        recvOff[q] = nRecv;
        nSend += sendCount[q];
        nRecv += recvCount[q];
    }
    const int nStay = nLocalOld - nSend, nLocal = nStay + nRecv;
    const int rec = c->haveRandom ? 8 : 7;
    CK(c->sendBuf.ensure((size_t)nSend * rec + 8));
    CK(c->recvBuf.ensure((size_t)nRecv * rec + 8));
    if (nSend)
    {
        DdcOffsets off;
        for (int q = 0; q < DDC_MAXRANKS; q++) off.off[q] = q < nr ? sendOff[q] : 0;
        LAUNCH(k_rd_pack, nbOld, 256, 0, st)(nIonOld, c->pos4[cur].p, c->vel[cur][0].p, c->vel[cur][1].p, c->vel[cur][2].p, c->ddcDest.p, me, work, off,
                                         rec, c->rngState.p, c->sendBuf.p);
        CKL("k_rd_pack");
    }
    rc = exchange(c, rec, sendCount, sendOff, recvCount, recvOff);
    if (rc) return rc;
    rc = ensureSlots(c, nxt, nLocal, false);
    if (rc) return rc;
    LAUNCH(k_rd_clear, nbOld, 256, 0, st)(nIonOld, c->beadOfSlot[cur].p, c->slotOfBead.p);
    CKL("k_rd_clear");
    LAUNCH(k_rd_compact, nbOld, 256, 0, st)(nIonOld, c->pos4[cur].p, c->vel[cur][0].p, c->vel[cur][1].p, c->vel[cur][2].p, c->beadOfSlot[cur].p,
                                        c->ddcDest.p, me, work, c->pos4[nxt].p, c->vel[nxt][0].p, c->vel[nxt][1].p, c->vel[nxt][2].p,
                                        c->beadOfSlot[nxt].p, c->slotOfBead.p);
    CKL("k_rd_compact");
    if (nRecv)
    {
        LAUNCH(k_rd_unpack, (nRecv + 255) / 256, 256, 0, st)(nRecv, c->recvBuf.p, rec, nStay, c->wOfBead.p, c->pos4[nxt].p, c->vel[nxt][0].p,
                                                         c->vel[nxt][1].p, c->vel[nxt][2].p, c->beadOfSlot[nxt].p, c->slotOfBead.p, c->rngState.p);
        CKL("k_rd_unpack");
    }
    // ---- B. ghosts ----
    const int nbNew = (nLocal + 255) / 256;
    LAUNCH(k_rd_bbox, nbNew, 256, 0, st)(nLocal, c->pos4[nxt].p, g, work);
    CKL("k_rd_bbox");
    LAUNCH(k_rd_row, 1, DDC_ROW, 0, st)(work, 1, nLocal, c->ddcRow, c->ddcBox6);
    CKL("k_rd_row");
    CKN(ncclAllGather(c->ddcBox6, c->ddcBoxAll, 6, ncclDouble, (ncclComm_t)c->nccl, st));
    LAUNCH(k_rd_boxes, 1, 128, 0, st)(nr, c->ddcBoxAll, (DdcBoxes *)c->boxes);
    CKL("k_rd_boxes");
    CK(c->ddcMask.ensure((size_t)nLocal + 1));
    LAUNCH(k_rd_ghostmask, nbNew, 256, 0, st)(nLocal, c->pos4[nxt].p, g, (const DdcBoxes *)c->boxes, c->ddcMask.p, work);
    CKL("k_rd_ghostmask");
    rc = gatherRows(c, 1, nLocal, nullptr);
    if (rc) return rc;
    c->hSendCount.assign(nr, 0); c->hRecvCount.assign(nr, 0);
    c->hSendOff.assign(nr, 0); c->hRecvOff.assign(nr, 0);
    int nSendG = 0, nRecvG = 0;
    for (int q = 0; q < nr; q++)
    {
        if (q == me) continue;
        c->hSendCount[q] = rows[me * DDC_ROW + q];
        c->hRecvCount[q] = rows[q * DDC_ROW + me];
        c->hSendOff[q] = nSendG;
        c->hRecvOff[q] = nRecvG;
        nSendG += c->hSendCount[q];
        nRecvG += c->hRecvCount[q];
    }
    c->nSendTot = nSendG;
    c->nRecvTot = nRecvG;
    const int nIon = nLocal + nRecvG;
    rc = ensureSlots(c, nxt, nIon, true);      // holds the new locals
    if (rc) return rc;
    rc = ensureSlots(c, cur, nIon, false);     // the old set is dead from here on
    if (rc) return rc;
    rc = ensureAux(c, nIon);
    if (rc) return rc;
    CK(c->ddcList.ensure((size_t)nSendG + nRecvG + 1));
    CK(c->sendSlot.ensure((size_t)nSendG + 1));
    CK(c->recvSlot.ensure((size_t)nRecvG + 1));
    CK(c->sendBuf.ensure((size_t)nSendG * 4 + 8));
    CK(c->recvBuf.ensure((size_t)nRecvG * 4 + 8));
    if (nSendG)
    {
        DdcOffsets off;
        for (int q = 0; q < DDC_MAXRANKS; q++) off.off[q] = q < nr ? c->hSendOff[q] : 0;
        LAUNCH(k_rd_ghostpack, nbNew, 256, 0, st)(nLocal, c->pos4[nxt].p, c->ddcMask.p, nr, work, off, c->ddcList.p, c->sendBuf.p);
        CKL("k_rd_ghostpack");
    }
    rc = exchange(c, 4, c->hSendCount, c->hSendOff, c->hRecvCount, c->hRecvOff);
    if (rc) return rc;
    if (nRecvG)
    {
        LAUNCH(k_rd_ghostunpack, (nRecvG + 255) / 256, 256, 0, st)(nRecvG, c->recvBuf.p, nLocal, c->wOfBead.p, c->pos4[nxt].p, c->vel[nxt][0].p,
                                                               c->vel[nxt][1].p, c->vel[nxt][2].p, c->beadOfSlot[nxt].p, c->slotOfBead.p,
                                                               c->ddcList.p + nSendG);
        CKL("k_rd_ghostunpack");
    }
    c->cur = nxt;
    c->nLocal = nLocal;
    c->nIon = nIon;
    c->localsDirty = true;
    return DDCB200_OK;
}

// ddcUpdate (src/ddcUpdate.c:40-88): owners send the new positions of the beads their neighbours hold as ghosts
static int haloExchange(ddcb200_ctx *c, cudaStream_t st)
{
    ProfScope ps(c, PROF_HALO, st);
    ncclComm_t comm = (ncclComm_t)c->nccl;
    const int cur = c->cur;
    if (c->nSendTot)
    {
        LAUNCH(k_halo_pack, (c->nSendTot + 255) / 256, 256, 0, st)(c->nSendTot, c->sendSlot.p, c->pos4[cur].p, c->sendBuf.p);
        CKL("k_halo_pack");
    }
    CKN(ncclGroupStart());
    for (int pp = 0; pp < c->nranks; pp++)
    {
        if (pp == c->rank) continue;
        if (c->hSendCount[pp]) CKN(ncclSend(c->sendBuf.p + 3 * (size_t)c->hSendOff[pp], 3 * (size_t)c->hSendCount[pp], ncclDouble, pp, comm, st));
        if (c->hRecvCount[pp]) CKN(ncclRecv(c->recvBuf.p + 3 * (size_t)c->hRecvOff[pp], 3 * (size_t)c->hRecvCount[pp], ncclDouble, pp, comm, st));
    }
    CKN(ncclGroupEnd());
    if (c->nRecvTot)
    {
        LAUNCH(k_halo_unpack, (c->nRecvTot + 255) / 256, 256, 0, st)(c->nRecvTot, c->recvSlot.p, c->recvBuf.p, c->pos4[cur].p, c->posBuild[0].p,
                                                                 c->posBuild[1].p, c->posBuild[2].p, c->pc, c->dmax2,
                                                                 c->nCellsBuilt > 0 ? c->cellDmax.p : nullptr, c->cellOfSlot[cur].p);
        CKL("k_halo_unpack");
    }
    c->haloDirty = false;
    return DDCB200_OK;
}

// ---- list build -------------------------------------------------------------------------
// per-bead gather lists of the bonded kernel: for every bead the (term, role) pairs it takes part in, ascending term order
// (terms are kind-sorted), restraints last.  Static; rebuilt only when the term tables change.
static int ensureBondCsr(ddcb200_ctx *c)
{
    if (!c->bondCsrDirty) return DDCB200_OK;
    const int64_t nG = c->nGlobal;
    std::vector<int> off((size_t)nG + 1, 0);
    auto need = [](const Term &t) { return t.kind == 0 ? 2 : (t.kind <= 3 ? 3 : 4); };
    for (const Term &t : c->hTerms)
    {
        const int ids[4] = {t.i, t.j, t.k, t.l};
        for (int a = 0; a < need(t); a++)
        {
            if (ids[a] < 0 || ids[a] >= nG) return fail(DDCB200_ERR_ARG, "bonded term bead index out of range");
            off[(size_t)ids[a] + 1]++;
        }
    }
    for (int b : c->hRestrBead)
    {
        if (b < 0 || b >= nG) return fail(DDCB200_ERR_ARG, "restraint bead index out of range");
        off[(size_t)b + 1]++;
    }
    for (int64_t b = 0; b < nG; b++) off[(size_t)b + 1] += off[(size_t)b];
    if ((int64_t)off[(size_t)nG] < 0) return fail(DDCB200_ERR_CAPACITY, "too many bonded term entries");
    std::vector<uint32_t> ent((size_t)off[(size_t)nG] + 1);
    std::vector<int> fill(off.begin(), off.end() - 1);
    for (size_t t = 0; t < c->hTerms.size(); t++)
    {
        const Term &tm = c->hTerms[t];
        const int ids[4] = {tm.i, tm.j, tm.k, tm.l};
        for (int a = 0; a < need(tm); a++) ent[(size_t)fill[(size_t)ids[a]]++] = ((uint32_t)t << 2) | (uint32_t)a;
    }
    for (size_t r = 0; r < c->hRestrBead.size(); r++)
        ent[(size_t)fill[(size_t)c->hRestrBead[r]]++] = (uint32_t)(c->hTerms.size() + r) << 2;
    CK(c->bondCsrOff.ensure((size_t)nG + 1));
    CK(c->bondEnt.ensure(ent.size()));
    CK(cudaMemcpy(c->bondCsrOff.p, off.data(), ((size_t)nG + 1) * sizeof(int), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c->bondEnt.p, ent.data(), ent.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    c->bondCsrDirty = false;
    return DDCB200_OK;
}

static int localSums(ddcb200_ctx *c, int atBuild);
extern "C" int ddcb200_constructList(ddcb200_ctx *c)
{
    if (!c) return fail(DDCB200_ERR_ARG, "null ctx");
    if (c->nLocal == 0) return fail(DDCB200_ERR_STATE, "no state");
    if (c->ntypes == 0) return fail(DDCB200_ERR_STATE, "martiniNonBondParms not called");
    CK(cudaSetDevice(c->device));
    ProfScope ps(c, PROF_LIST);
    if (c->nranks > 1)
    {
        int rcd = redomain(c);
        if (rcd) return rcd;
    }
    const int nIon = (int)c->nIon, nLocal = (int)c->nLocal, nPad = (int)c->nPad;
    const int cur = c->cur, nxt = cur ^ 1;
    cudaStream_t st = c->stream;
    const int maxCells = nIon + 4;
    const int mmBlocks = std::min(1024, (nIon + 255) / 256);
    LAUNCH(k_minmax_partial, mmBlocks, 256, 0, st)(c->pos4[cur].p, nIon, c->box, c->mmPartial.p);
    LAUNCH(k_grid_setup, 1, 32, 0, st)(c->mmPartial.p, mmBlocks, nIon, c->box, c->grid, maxCells);
    CK(cudaMemsetAsync(c->cellCount.p, 0, (2 * (size_t)nIon + 16) * sizeof(int), st));
    const int nb = (nIon + 255) / 256;
    LAUNCH(k_cell_count, nb, 256, 0, st)(c->pos4[cur].p, nIon, c->box, c->grid, c->cellOfSlot[cur].p, c->rank0.p, c->cellCount.p,
                                     c->beadOfSlot[cur].p, c->orderKey.p);
    LAUNCH(k_cell_scan, 1, 1024, 0, st)(c->cellCount.p, c->cellStart.p, c->grid);
    LAUNCH(k_cell_scatter, nb, 256, 0, st)(nIon, c->cellOfSlot[cur].p, c->rank0.p, c->cellStart.p, c->member.p);
    LAUNCH(k_cell_rank, nb, 256, 0, st)(nIon, c->cellOfSlot[cur].p, c->orderKey.p, c->cellStart.p, c->member.p, c->perm.p);
    LAUNCH(k_gather, nb, 256, 0, st)(nIon, c->perm.p, c->cellOfSlot[cur].p, c->cellOfSlot[nxt].p, c->pos4[cur].p, c->pos4[nxt].p,
                                 c->vel[cur][0].p, c->vel[cur][1].p, c->vel[cur][2].p, c->vel[nxt][0].p, c->vel[nxt][1].p,
                                 c->vel[nxt][2].p, c->beadOfSlot[cur].p, c->beadOfSlot[nxt].p, c->slotOfBead.p, nLocal,
                                 c->pos32.p, c->posBuild[0].p, c->posBuild[1].p, c->posBuild[2].p);
    CKL("cell sort");
    c->cur = nxt;
    CK(cudaMemsetAsync(c->dmax2, 0, 4 * sizeof(unsigned long long), st));
    c->pruneValid = false;      // the candidate buffer the pruned rows live in is overwritten by the build
    c->rebased = false;
    c->movedSinceRef = false;
    c->sincePrune = 0;
    CK(c->dispOfSlot.ensure((size_t)nPad));
    CK(cudaMemsetAsync(c->dispOfSlot.p, 0, (size_t)nPad * sizeof(float), st));

    // fp32 candidate filter: margin covers the rounding of box-sized coordinates (and the image shift) to fp32
    const double Lmax = std::max(c->box.hxx, std::max(c->box.hyy, c->box.hzz));
    const double rl = sqrt(c->box.rlist2);
    const double margin = std::max(1e-3, 64.0 * 1.1920929e-7 * 1.5 * Lmax / rl);
    const float rl2f = (float)(c->box.rlist2 * (1.0 + margin));
    if (c->nbrCap == 0)
    {
        // entries per row allocated at first; a build that overflows it regrows from the measured maximum and repeats
        const char *cp = getenv("DDCB200_NBRCAP");
        c->nbrCap = cp ? std::max(8, atoi(cp)) : 176;
    }
    if (!c->evList[0])
    {
        CK(cudaEventCreate(&c->evList[0]));
        CK(cudaEventCreate(&c->evList[1]));
    }
    // bonded records of the resident local beads: counted and scanned here so that the total comes back with the grid record
    const bool haveBonded = c->nTerms + c->nRestr > 0;
    if (haveBonded)
    {
        int rcb = ensureBondCsr(c);
        if (rcb) return rcb;
        CK(c->bondCount.ensure((size_t)nPad));
        CK(c->bondStart.ensure((size_t)nPad));
        CK(c->bondCount0.ensure((size_t)BOND_GROUPS * nPad));
        CK(c->bondStart0.ensure((size_t)BOND_GROUPS * nPad));
        LAUNCH(k_bond_count, (nLocal + 255) / 256, 256, 0, st)(nLocal, c->pos4[nxt].p, c->bondCsrOff.p, c->bondEnt.p, c->nTerms, c->termsBead.p,
                                                           c->bondCount.p, c->bondCount0.p);
        CKL("k_bond_count");
        for (int which = 0; which < 2; which++)
        {
            // entries per bead -> where the bead's run starts; owned terms per (kind group, bead) -> the place of the bead's first term
            // of that group in the kind-grouped numbering of the local terms
            const int n = which ? BOND_GROUPS * nLocal : nLocal;
            const int nsb = (n + SCAN_BLOCK - 1) / SCAN_BLOCK;
            CK(c->scanBlocks.ensure((size_t)nsb + 1));
            const int *cnt = which ? c->bondCount0.p : c->bondCount.p;
            int *start = which ? c->bondStart0.p : c->bondStart.p;
            LAUNCH(k_scan_local, nsb, SCAN_BLOCK, 0, st)(n, cnt, start, c->scanBlocks.p);
            CKL("k_scan_local");
            LAUNCH(k_scan_blocks, 1, 1024, 0, st)(nsb, c->scanBlocks.p, which ? &c->grid->bondTerms : &c->grid->bondTotal);
            CKL("k_scan_blocks");
            LAUNCH(k_scan_add, nsb, SCAN_BLOCK, 0, st)(n, start, c->scanBlocks.p);
            CKL("k_scan_add");
        }
    }
    // after the cell sort the local beads are slots [0, nLocal): rows, tiles and every per-bead kernel cover that range only
    const int tilesL = (nLocal + TILE - 1) / TILE;
    if (c->nranks > 1)
    {
        CK(c->tileGhost.ensure((size_t)tilesL + 1));
        CK(c->tileOrder.ensure((size_t)tilesL + 1));
    }
    for (int attempt = 0; attempt < 4; attempt++)
    {
        CK(c->nbr.ensure((size_t)c->nbrCap * nPad));
        CK(c->nbrRaw.ensure((size_t)c->nbrCap * nPad));
        CK(cudaEventRecord(c->evList[0], st));
        LAUNCH(k_nbr_filter, tilesL, 128, 0, st)(nLocal, nPad, c->pos32.p, c->cellOfSlot[nxt].p, c->cellStart.p, c->box, rl2f, c->grid,
                                            c->nbrCap, c->nbrRaw.p, c->nbrRawCount.p);
        CKL("k_nbr_filter");
        // one sweep; rows in two segments around the bin edge
        LAUNCH(k_nbr_exact2, tilesL, 128, 0, st)(nLocal, nPad, c->nbrCap, c->pos4[nxt].p, c->box, c->box.binEdge2[0], c->grid, c->nbrRaw.p,
                                                c->nbrRawCount.p, c->nbr.p, c->nbrCount.p, c->nbrCum.p, c->gidOfBead.p, c->molTypeOfBead.p,
                                                c->molTypeSingle.p, c->bpairOffset.p, c->bpairKey.p, c->haveExcl ? 1 : 0,
                                                c->nranks > 1 ? c->tileGhost.p : nullptr);
        CKL("k_nbr_exact2");
        CK(cudaEventRecord(c->evList[1], st));
        if (c->nranks > 1)
        {
            LAUNCH(k_tile_order, 1, 1024, 0, st)(tilesL, c->tileGhost.p, c->tileOrder.p, c->grid);
            CKL("k_tile_order");
        }
        CK(cudaMemcpyAsync(c->gridHost, c->grid, sizeof(GridDev), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        if (c->gridHost->error & 2) return fail(DDCB200_ERR_CAPACITY, "cell grid larger than the bead count bound");
        if (!(c->gridHost->error & 1))
        {
            CK(cudaEventElapsedTime(&c->listBuildMs, c->evList[0], c->evList[1]));
            break;
        }
        // a candidate row overflowed: grow and redo both passes
        if (attempt == 3) return fail(DDCB200_ERR_CAPACITY, "neighbor list capacity could not be satisfied");
        c->nbrCap = (int)(c->gridHost->maxRaw * 1.25) + 8;
        CK(cudaMemsetAsync(&c->grid->error, 0, sizeof(int), st));
        CK(cudaMemsetAsync(&c->grid->maxCount, 0, sizeof(int), st));
        CK(cudaMemsetAsync(&c->grid->maxRaw, 0, sizeof(int), st));
        CK(cudaMemsetAsync(&c->grid->totalEntries, 0, sizeof(unsigned long long), st));
    }
    if (haveBonded)
    {
        c->nBondRec = c->gridHost->bondTotal;
        c->nBondTerms = c->gridHost->bondTerms;
        const size_t nAll = (size_t)(c->nTerms + c->nRestr);
        CK(c->bondRec.ensure((size_t)c->nBondTerms + 1));
        CK(c->bondStageIdx.ensure((size_t)c->nBondRec + 1));
        CK(c->bondStage.ensure(12 * (size_t)c->nBondTerms + 12));
        CK(c->termMap.ensure(nAll + 1));
        CK(cudaMemsetAsync(c->termMap.p, 0xff, nAll * sizeof(int), st));
        LAUNCH(k_bond_resolve_terms, (nLocal + 255) / 256, 256, 0, st)(nLocal, c->pos4[nxt].p, c->bondCsrOff.p, c->bondEnt.p, c->nTerms, c->termsBead.p,
                                                                   c->slotOfBead.p, c->bondStart0.p, c->bondRec.p, c->termMap.p);
        CKL("k_bond_resolve_terms");
        LAUNCH(k_bond_resolve_beads, (nLocal + 255) / 256, 256, 0, st)(nLocal, c->pos4[nxt].p, c->bondCsrOff.p, c->bondEnt.p, c->termMap.p,
                                                                   c->bondStart.p, c->bondCount.p, c->bondStageIdx.p);
        CKL("k_bond_resolve_beads");
    }
    if (c->nranks > 1)
    {
        if (c->nSendTot)
        {
            LAUNCH(k_ddc_toslots, (c->nSendTot + 255) / 256, 256, 0, st)(c->nSendTot, c->ddcList.p, c->slotOfBead.p, c->sendSlot.p);
            CKL("k_ddc_toslots");
        }
        if (c->nRecvTot)
        {
            LAUNCH(k_ddc_toslots, (c->nRecvTot + 255) / 256, 256, 0, st)(c->nRecvTot, c->ddcList.p + c->nSendTot, c->slotOfBead.p, c->recvSlot.p);
            CKL("k_ddc_toslots");
        }
        c->haloDirty = false;   // the replicated state carried the current positions of every ghost
    }
    if (c->prm.updateRate == 0)
    {
        // neighborCheck compares with the positions of the build; posBuild is re-based by the prunes of the pair walk
        for (int a = 0; a < 3; a++)
        {
            CK(c->posCheck[a].ensure((size_t)nPad));
            CK(cudaMemcpyAsync(c->posCheck[a].p, c->posBuild[a].p, (size_t)nIon * sizeof(double), cudaMemcpyDeviceToDevice, st));
        }
        int rcs = localSums(c, 1);   // neighborRef: rbar at the build
        if (rcs) return rcs;
    }
    c->listValid = true;
    c->hBuild[0] = c->box.hxx; c->hBuild[1] = c->box.hyy; c->hBuild[2] = c->box.hzz;
    for (int a = 0; a < 3; a++) c->hCheck[a] = c->hBuild[a];
    c->pc.listSlack = 0.0;
    c->lastBuildLoop = c->loop;
    c->totalEntries = (int64_t)c->gridHost->totalEntries;
    c->nPairsListed = (int64_t)(c->gridHost->totalEntries / 2);
    c->nTilesInterior = c->nranks > 1 ? c->gridHost->nInterior : tilesL;
    // per-cell displacement maxima start at zero with the new list
    c->nCellsBuilt = c->gridHost->ncell;
    CK(c->cellDmax.ensure(2 * (size_t)c->nCellsBuilt + 2));
    CK(c->nbrDmax.ensure(2 * (size_t)c->nCellsBuilt + 2));
    CK(cudaMemsetAsync(c->cellDmax.p, 0, 2 * (size_t)c->nCellsBuilt * sizeof(unsigned long long), st));
    return DDCB200_OK;
}

// ---- force evaluation -------------------------------------------------------------------
// the pruned rows of this context (pair.cuh): where they live and the margin they are cut with
static double pruneArgsOf(const ddcb200_ctx *c, PruneArgs &pr)
{
    const double deltaR = sqrt(c->box.rlist2) - c->pc.rmax;
    const double margin = deltaR * pruneFracOf(c);
    pr.rows = c->nbrRaw.p;
    pr.count = c->pruneCount.p;
    pr.keep2 = (c->pc.rmax + margin) * (c->pc.rmax + margin) * (1.0 + 1e-12);
    pr.walkLim = 1e300;
    pr.useLim = c->pc.rmax + margin;
    pr.farTop = c->nbrCap - 1;
    return margin;
}

static int reduceCols(ddcb200_ctx *c, const double *partial, int nblocks, int ncol, const int *map)
{
    // column map lives in a small device table: [0..7] pair, [8..18] bonded, [19..25] kinetic
    LAUNCH(k_reduce_cols, ncol, 256, 0, c->stream)(partial, nblocks, ncol, map, c->acc, 1);
    CKL("k_reduce_cols");
    return DDCB200_OK;
}

static int ensureColMap(ddcb200_ctx *c)
{
    if (c->colMap.p) return DDCB200_OK;
    const int m[32] = {ACC_ELJ, ACC_EELE, ACC_VXX, ACC_VYY, ACC_VZZ, ACC_VXY, ACC_VXZ, ACC_VYZ,
                       ACC_VXX, ACC_VYY, ACC_VZZ, ACC_VXY, ACC_VXZ, ACC_VYZ, ACC_EBOND, ACC_EANGLE, ACC_ETORS, ACC_EIMPR, ACC_EREST,
                       ACC_RK, ACC_TXX, ACC_TYY, ACC_TZZ, ACC_TXY, ACC_TXZ, ACC_TYZ,
                       0, 1, 2, 3, 4, 5};   // [26..31]: columns of the neighbor-check sums (chk, nbrcheck.cuh)
    CK(c->colMap.ensure(32));
    CK(cudaMemcpy(c->colMap.p, m, sizeof(m), cudaMemcpyHostToDevice));
    return DDCB200_OK;
}


// ---- displacement-triggered rebuild (updateRate == 0): neighborRef / neighborCheck ------------
static BoxConst checkBox(const ddcb200_ctx *c)
{
    // positions are taken at the image nearest to the domain centre (src/neighbor.c:220-230): the brick centre on
    // several ranks, the GeomBox centre otherwise
    BoxConst b = c->box;
    if (c->nranks > 1)
    {
        double cc[3];
        ddcBrickCentre(c->rank, ddcGeomOf(c), cc);
        b.cx = cc[0]; b.cy = cc[1]; b.cz = cc[2];
    }
    return b;
}

static int localSums(ddcb200_ctx *c, int atBuild)
{
    const int tiles = (int)(c->nPad / TILE);
    CK(c->chk.ensure(8));
    CK(c->chkPartial.ensure((size_t)tiles * 3 + 8));
    if (!c->chkHost) CK(cudaMallocHost((void **)&c->chkHost, 8 * sizeof(double)));
    int rc = ensureColMap(c);
    if (rc) return rc;
    LAUNCH(k_nbr_rbar_partial, tiles, TILE, 0, c->stream)((int)c->nIon, c->pos4[c->cur].p, checkBox(c), c->chkPartial.p);
    CKL("k_nbr_rbar_partial");
    LAUNCH(k_reduce_cols, 3, 256, 0, c->stream)(c->chkPartial.p, tiles, 3, c->colMap.p + (atBuild ? 29 : 26), c->chk.p, 0);
    CKL("k_reduce_cols");
    return DDCB200_OK;
}

// neighborCheck (src/neighbor.c:117-208): 1 = rebuild now
static int neighborCheck(ddcb200_ctx *c, bool *update)
{
    ProfScope ps(c, PROF_LIST);
    cudaStream_t st = c->stream;
    int rc = localSums(c, 0);
    if (rc) return rc;
    CK(cudaMemsetAsync(c->chk.p + 6, 0, sizeof(double), st));
    LAUNCH(k_nbr_check, (int)(c->nPad / TILE), TILE, 0, st)((int)c->nIon, (int)c->nLocal, c->pos4[c->cur].p, c->posCheck[0].p,
                                                             c->posCheck[1].p, c->posCheck[2].p, c->pc, c->chk.p);
    CKL("k_nbr_check");
    if (c->nranks > 1)   // check4updateNeighbor's MPI_Allreduce of the flags (src/ddcUpdateAll.c:48-62) = max of d^2
        CKN(ncclAllReduce(c->chk.p + 6, c->chk.p + 6, 1, ncclDouble, ncclMax, (ncclComm_t)c->nccl, st));
    CK(cudaMemcpyAsync(c->chkHost, c->chk.p + 6, sizeof(double), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    // box-strain term |1 - h0 hinv u| (rcut + deltaR): h0 is the box at the build (it only differs under the barostat)
    const double *hi = c->box.hinv;
    double h[9];
    memcpy(h, c->prm.h, sizeof(h));
    if (c->hCheck[0] != 0.0) { h[0] = c->hCheck[0]; h[4] = c->hCheck[1]; h[8] = c->hCheck[2]; }
    const double ux = hi[0] + hi[1] + hi[2], uy = hi[3] + hi[4] + hi[5], uz = hi[6] + hi[7] + hi[8];
    const double sx = fabs(1.0 - (h[0] * ux + h[1] * uy + h[2] * uz)), sy = fabs(1.0 - (h[3] * ux + h[4] * uy + h[5] * uz)),
                 sz = fabs(1.0 - (h[6] * ux + h[7] * uy + h[8] * uz));
    double dmax = sx;
    if (dmax < sy) dmax = sy;
    if (dmax < sz) dmax = sz;
    dmax *= c->prm.rmax + c->prm.deltaR;   // nbr->rcut[0].value + nbr->deltaR
    dmax += 2.0 * sqrt(c->chkHost[0]);
    *update = !(dmax < c->prm.deltaR);
    return DDCB200_OK;
}

extern "C" int ddcb200_ddcenergy(ddcb200_ctx *c, int withEnergy)
{
    if (!c) return fail(DDCB200_ERR_ARG, "null ctx");
    if (c->nLocal == 0) return fail(DDCB200_ERR_STATE, "no state");
    if (c->ntypes == 0) return fail(DDCB200_ERR_STATE, "martiniNonBondParms not called");
    CK(cudaSetDevice(c->device));
    int rc = ensureColMap(c);
    if (rc) return rc;
    // evalUpdateFlag, src/ddcUpdateAll.c:64-71
    bool due = c->prm.updateRate > 0 && (c->loop % c->prm.updateRate) == 0 && c->lastBuildLoop != c->loop;
    // ddcUpdate (src/ddcUpdate.c:40-88).  With a fixed rebuild rate the halo goes to its own stream: the pair rows that read no
    // ghost position run meanwhile, the others wait for it.  (updateRate = 0: neighborCheck needs the ghosts first)
    bool overlapped = false;
    // pruned rows (pair.cuh): this evaluation writes them if it follows a build or the last prune is pruneEvery evaluations old
    const bool pruneCfg = c->pruneEvery > 0 && pruneFracOf(c) < 0.8;      // (a margin near the whole skin prunes nothing: e.g. a deck that rebuilds every 5 steps)
    bool pruneStep = pruneCfg && (!c->listValid || due || !c->pruneValid || c->sincePrune + 1 >= c->pruneEvery);
    if (c->listValid && !due && c->nranks > 1 && c->haloDirty)
    {
        // (a prune step takes the reference positions of the ghosts after the halo: in line)
        if (c->haloOverlap && c->prm.updateRate > 0 && !pruneStep)
        {
            CK(cudaEventRecord(c->evPos, c->stream));
            CK(cudaStreamWaitEvent(c->streamH, c->evPos, 0));
            rc = haloExchange(c, c->streamH);
            if (rc) return rc;
            if (c->walkPerCell && c->walkPerBead && c->nCellsBuilt > 0)
            {
                // the neighbourhood bounds that include the ghosts' displacements, for the rows that wait for the halo.  (The
                // local parts of cellDmax are complete: this stream waited for the integrator's event.)
                LAUNCH(k_nbr_dmax, (c->nCellsBuilt + 127) / 128, 128, 0, c->streamH)(c->grid, c->cellDmax.p, 1, c->nbrDmax.p + c->nCellsBuilt);
                CKL("k_nbr_dmax");
            }
            CK(cudaEventRecord(c->evHalo, c->streamH));
            overlapped = true;
        }
        else
        {
            rc = haloExchange(c, c->stream);
            if (rc) return rc;
        }
    }
    if (c->listValid && c->prm.updateRate == 0 && c->lastBuildLoop != c->loop)
    {
        rc = neighborCheck(c, &due);
        if (rc) return rc;
    }
    if (!c->listValid || due)
    {
        rc = ddcb200_constructList(c);
        if (rc) return rc;
    }
    pruneStep = pruneCfg && (pruneStep || !c->pruneValid);      // (updateRate = 0: neighborCheck may just have asked for the build)
    cudaStream_t st = c->stream;
    const int cur = c->cur;
    const int nLocal = (int)c->nLocal, nPad = (int)c->nPad;      // local beads = slots [0, nLocal)
    const int tiles = (nLocal + TILE - 1) / TILE;
    if (withEnergy) CK(cudaMemsetAsync(c->acc, 0, ACC_N * sizeof(double), st));
    // the bonded terms first: every term once, its forces staged; the pair kernel adds each bead's staged forces to its pair force
    int bBlocks = 0;
    BondAdd ba = {nullptr, nullptr, nullptr, nullptr};
    if (c->nTerms + c->nRestr > 0 && c->nBondTerms > 0)
    {
        ba.start = c->bondStart.p;
        ba.count = c->bondCount.p;
        ba.stageIdx = c->bondStageIdx.p;
        ba.stage = c->bondStage.p;
        ProfScope ps(c, PROF_BONDED);
        bBlocks = (c->nBondTerms + BONDED_THREADS - 1) / BONDED_THREADS;
        CK(c->bondPartial.ensure((size_t)bBlocks * BONDED_ACC + 8));
        V3 *stage = (V3 *)c->bondStage.p;
        if (withEnergy)
            LAUNCH((k_bonded<true, 1>), bBlocks, BONDED_THREADS, 0, st)(c->nBondTerms, c->bondRec.p, c->restrParm.p, c->restrOrigin, c->pos4[cur].p, c->pc,
                                                                    stage, c->bondPartial.p);
        else if (c->bondedCap == 8)
            LAUNCH((k_bonded<false, 8>), bBlocks, BONDED_THREADS, 0, st)(c->nBondTerms, c->bondRec.p, c->restrParm.p, c->restrOrigin, c->pos4[cur].p, c->pc,
                                                                     stage, c->bondPartial.p);
        else if (c->bondedCap == 12)
            LAUNCH((k_bonded<false, 12>), bBlocks, BONDED_THREADS, 0, st)(c->nBondTerms, c->bondRec.p, c->restrParm.p, c->restrOrigin, c->pos4[cur].p, c->pc,
                                                                      stage, c->bondPartial.p);
        else
            LAUNCH((k_bonded<false, 1>), bBlocks, BONDED_THREADS, 0, st)(c->nBondTerms, c->bondRec.p, c->restrParm.p, c->restrOrigin, c->pos4[cur].p, c->pc,
                                                                     stage, c->bondPartial.p);
        CKL("k_bonded");
    }
    int pruneMode = 0;
    PruneArgs pr = {nullptr, nullptr, 0.0, 0.0, 0.0, c->nbrCap - 1};
    PairConst pcl = c->pc;
    if (pruneCfg)
    {
        CK(c->pruneCount.ensure((size_t)nPad));
        const double margin = pruneArgsOf(c, pr);
        if (pruneStep)
        {
            if (c->movedSinceRef)
            {
                // the positions of this evaluation (ghosts included: the halo ran in line) become the reference of the displacement bounds
                const int nIon = (int)c->nIon;
                LAUNCH(k_rebase, (nIon + 255) / 256, 256, 0, st)(nIon, c->pos4[cur].p, c->posBuild[0].p, c->posBuild[1].p, c->posBuild[2].p,
                                                              c->dispOfSlot.p, c->dmax2);
                CKL("k_rebase");
                if (c->nCellsBuilt > 0) CK(cudaMemsetAsync(c->cellDmax.p, 0, 2 * (size_t)c->nCellsBuilt * sizeof(unsigned long long), st));
                c->rebased = true;
                c->movedSinceRef = false;
                // the box of the reference positions: a barostat's change of the box edges counts from here (PairConst::listSlack)
                c->hBuild[0] = c->box.hxx; c->hBuild[1] = c->box.hyy; c->hBuild[2] = c->box.hzz;
                c->pc.listSlack = 0.0;
                pcl.listSlack = 0.0;
            }
            // right after a build the rows are ordered by the distances of these very positions: the walk can stop at rmax + margin
            if (!c->rebased) pr.walkLim = (c->pc.rmax + margin) * (1.0 + 1e-12);
            pruneMode = 1;
        }
        else pruneMode = 2;
    }
    else if (c->rebased) pcl.listSlack = 1e30;      // no bin-limited walk against reference positions that are not the build's
    {
        const size_t smem = pairSmemBytes(c->ntypes);
        const float *disp = c->walkPerBead ? c->dispOfSlot.p : nullptr;
        cudaStream_t pst = st;      // the stream of the next pair launch
        auto launchPair = [&](int nTiles, const int *order, int base, int withGhosts) -> int {
            if (nTiles <= 0) return DDCB200_OK;
            cudaStream_t st = pst;
            ProfScope ps(c, pruneMode == 1 ? PROF_PAIR_PRUNE : PROF_PAIR, st);
            {
                const PairVariant &pv = g_pairVariants[c->pairVariant];
                const PairKernel kern = withEnergy ? pv.energy[pruneMode] : pv.force[pruneMode];
                const bool cellWalk = c->walkPerCell && c->walkPerBead && c->nCellsBuilt > 0;
                LAUNCH(kern, nTiles, TILE, smem, st)(nLocal, nPad, order, base, c->pos4[cur].p, c->nbr.p, c->nbrCum.p, c->dmax2, withGhosts, disp, c->ljTab.p,
                                                     c->shiftTab.p, c->qTab.p, pcl, c->frc[0].p, c->frc[1].p, c->frc[2].p, c->pairPartial.p,
                                                     cellWalk ? c->nbrDmax.p + (withGhosts ? c->nCellsBuilt : 0) : nullptr, c->cellOfSlot[cur].p, pr, ba);
            }
            CKL("k_pair");
            return DDCB200_OK;
        };
        const bool cellWalk = c->walkPerCell && c->walkPerBead && c->nCellsBuilt > 0;
        if (cellWalk)
        {
            // the displacement bound of every cell's neighbourhood, from the local beads (complete since the integrator ran)
            LAUNCH(k_nbr_dmax, (c->nCellsBuilt + 127) / 128, 128, 0, st)(c->grid, c->cellDmax.p, 0, c->nbrDmax.p);
            CKL("k_nbr_dmax");
            if (c->nranks > 1 && !overlapped)
            {
                LAUNCH(k_nbr_dmax, (c->nCellsBuilt + 127) / 128, 128, 0, st)(c->grid, c->cellDmax.p, 1, c->nbrDmax.p + c->nCellsBuilt);
                CKL("k_nbr_dmax");
            }
        }
        if (c->nranks > 1 && overlapped)
        {
            // the rows that read no ghost run now; the others wait for the halo on a stream of their own, so that their CTAs
            // fill the slots the first launch leaves idle in its last wave (two launches in one stream would each pay a tail)
            rc = launchPair(c->nTilesInterior, c->tileOrder.p, 0, 0);
            if (rc) return rc;
            if (tiles - c->nTilesInterior > 0)
            {
                CK(cudaStreamWaitEvent(c->streamB, c->evHalo, 0));
                if (bBlocks)
                {
                    // the boundary rows add the staged bonded forces too: their stream waits for k_bonded
                    CK(cudaEventRecord(c->evBonded, st));
                    CK(cudaStreamWaitEvent(c->streamB, c->evBonded, 0));
                }
                pst = c->streamB;
                rc = launchPair(tiles - c->nTilesInterior, c->tileOrder.p, c->nTilesInterior, 1);
                if (rc) return rc;
                CK(cudaEventRecord(c->evBoundary, c->streamB));
                CK(cudaStreamWaitEvent(st, c->evBoundary, 0));
            }
            else
                CK(cudaStreamWaitEvent(st, c->evHalo, 0));
        }
        else
            rc = launchPair(tiles, nullptr, 0, c->nranks > 1 ? 1 : 0);      // one launch over every row (ghost positions are in place)
        if (rc) return rc;
        if (pruneMode == 1)
        {
            c->pruneValid = true;
            c->sincePrune = 0;
        }
        else if (pruneMode == 2) c->sincePrune++;
    }
    if (withEnergy)
    {
        ProfScope ps(c, PROF_REDUCE);
        rc = reduceCols(c, c->pairPartial.p, tiles, 8, c->colMap.p);
        if (rc) return rc;
        if (bBlocks)
        {
            rc = reduceCols(c, c->bondPartial.p, bBlocks, BONDED_ACC, c->colMap.p + 8);
            if (rc) return rc;
        }
    }
    c->forcesValid = true;
    c->energyValid = withEnergy != 0;
    c->kineticValid = false;
    return DDCB200_OK;
}

template <int MODE>
static int launchIntegrate(ddcb200_ctx *c, double halfDt2, double halfDt1, double dt)
{
    ProfScope ps(c, PROF_INTEGRATE);
    const int cur = c->cur;
    const int tiles = (int)((c->nLocal + TILE - 1) / TILE);      // local beads are slots [0, nLocal)
    if (MODE & INT_KICK1_DRIFT) c->haloDirty = c->movedSinceRef = true;
    LAUNCH(k_integrate<MODE>, tiles, TILE, 0, c->stream)((int)c->nLocal, c->pos4[cur].p, c->vel[cur][0].p, c->vel[cur][1].p, c->vel[cur][2].p,
                                                     c->frc[0].p, c->frc[1].p, c->frc[2].p, c->massOfBead.p, halfDt2, halfDt1, dt, c->pc,
                                                     c->kinPartial.p, c->posBuild[0].p, c->posBuild[1].p, c->posBuild[2].p, c->dmax2, c->dispOfSlot.p,
                                                     c->nCellsBuilt > 0 ? c->cellDmax.p : nullptr, c->cellOfSlot[cur].p);
    CKL("k_integrate");
    return DDCB200_OK;
}

static int kineticTerms(ddcb200_ctx *c)
{
    // kinetic_terms alone (src/energy.c:48-163), e.g. after the first energy call
    int rc = launchIntegrate<INT_KE>(c, 0, 0, 0);
    if (rc) return rc;
    ProfScope ps(c, PROF_REDUCE);
    return reduceCols(c, c->kinPartial.p, (int)((c->nLocal + TILE - 1) / TILE), 7, c->colMap.p + 19);
}

static int nglfcSteps(ddcb200_ctx *c, int nsteps, double dt, const bool baro, const bool cons);
extern "C" int ddcb200_nglf(ddcb200_ctx *c, int nsteps, double dt)
{
    if (!c || nsteps < 0) return fail(DDCB200_ERR_ARG, "bad arguments");
    if (c->nLocal == 0) return fail(DDCB200_ERR_STATE, "no state");
    // nglf calls every bead's group->velocityUpdate (src/nglf.c:75,104): with a LANGEVIN group the step is the
    // NGLFCONSTRAINT pass without barostat and constraints
    if (c->anyLangevin) return nglfcSteps(c, nsteps, dt, false, false);
    CK(cudaSetDevice(c->device));
    int rc;
    if (!c->forcesValid)
    {
        // firstEnergyCall (src/masters.c:579-620)
        rc = ddcb200_ddcenergy(c, nsteps == 0 ? 1 : 0);
        if (rc) return rc;
    }
    const double half = 0.5 * dt;
    for (int s = 0; s < nsteps; s++)
    {
        // kick2 of the previous step is fused with kick1+drift of this one
        if (s == 0) rc = launchIntegrate<INT_KICK1_DRIFT>(c, 0.0, half, dt);
        else rc = launchIntegrate<INT_KICK2 | INT_KICK1_DRIFT>(c, half, half, dt);
        if (rc) return rc;
        c->loop++;
        c->time += dt;
        rc = ddcb200_ddcenergy(c, s == nsteps - 1 ? 1 : 0);
        if (rc) return rc;
    }
    if (nsteps > 0)
    {
        rc = launchIntegrate<INT_KICK2 | INT_KE>(c, half, 0.0, 0.0);
        if (rc) return rc;
        ProfScope ps(c, PROF_REDUCE);
        rc = reduceCols(c, c->kinPartial.p, (int)((c->nLocal + TILE - 1) / TILE), 7, c->colMap.p + 19);
        if (rc) return rc;
        c->kineticValid = true;
    }
    return DDCB200_OK;
}

// molecular virial correction + (several ranks) the all-reduce + the copy of the accumulators to pinned host memory
static int fetchAccumulators(ddcb200_ctx *c)
{
    cudaStream_t st = c->stream;
    CK(cudaMemsetAsync(c->acc + ACC_MVX, 0, 3 * sizeof(double), st));
    if (c->nMol > 0)
    {
        const int cur = c->cur;
        LAUNCH(k_mol_virial, (int)((c->nMol + 255) / 256), 256, 0, st)(c->nMol, c->molOffset.p, c->molBeads.p, c->slotOfBead.p, c->pos4[cur].p,
                                                                   c->frc[0].p, c->frc[1].p, c->frc[2].p, c->massOfBead.p, c->pc,
                                                                   c->acc + ACC_MVX);
        CKL("k_mol_virial");
    }
    c->accHost[ACC_N] = (double)c->totalEntries;
    CK(cudaMemcpyAsync(c->acc + ACC_NENTRIES, c->accHost + ACC_N, sizeof(double), cudaMemcpyHostToDevice, st));
    const double *accSrc = c->acc;
    if (c->nranks > 1)
    {
        // eval_energyInfo's MPI_Allreduce (src/energyInfo.c:9-63)
        CK(c->accG.ensure(ACC_N));
        CKN(ncclAllReduce(c->acc, c->accG.p, ACC_N, ncclDouble, ncclSum, (ncclComm_t)c->nccl, st));
        accSrc = c->accG.p;
    }
    CK(cudaMemcpyAsync(c->accHost, accSrc, ACC_N * sizeof(double), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return DDCB200_OK;
}

extern "C" int ddcb200_energyInfo(ddcb200_ctx *c, double kB, ddcb200_etype *out)
{
    if (!c || !out) return fail(DDCB200_ERR_ARG, "null argument");
    if (!c->energyValid) return fail(DDCB200_ERR_STATE, "no energy evaluation is current (call ddcenergy(1) or nglf)");
    CK(cudaSetDevice(c->device));
    int rc;
    if (!c->kineticValid)
    {
        rc = kineticTerms(c);
        if (rc) return rc;
        c->kineticValid = true;
    }
    rc = fetchAccumulators(c);
    if (rc) return rc;
    const double *a = c->accHost;
    memset(out, 0, sizeof(*out));
    out->eLJ = a[ACC_ELJ]; out->eEle = a[ACC_EELE];
    out->eBond = a[ACC_EBOND]; out->eAngle = a[ACC_EANGLE]; out->eTorsion = a[ACC_ETORS]; out->eImproper = a[ACC_EIMPR];
    out->eRestraint = a[ACC_EREST];
    out->eion = (a[ACC_ELJ] + a[ACC_EELE]) + (a[ACC_EBOND] + a[ACC_EANGLE] + a[ACC_ETORS] + a[ACC_EIMPR]) + a[ACC_EREST];
    out->rk = a[ACC_RK];
    for (int k = 0; k < 6; k++)
    {
        out->virial[k] = a[ACC_VXX + k];
        out->tion[k] = a[ACC_TXX + k];
    }
    // eval_energyInfo, src/energyInfo.c:75-113
    out->number = (double)(c->nranks > 1 ? c->nGlobal : c->nLocal);
    out->volume = c->box.volume;
    for (int k = 0; k < 6; k++) out->sion[k] = -(out->virial[k] + out->tion[k]) / out->volume;
    out->pion = -(out->sion[0] + out->sion[1] + out->sion[2]) / 3.0;
    out->temperature = 2.0 * out->rk / (3.0 * out->number - c->prm.nConstraints);
    // molecularPressure, src/molecularPressure.c:57-67
    for (int k = 0; k < 3; k++)
    {
        out->molVirial[k] = out->virial[k] + a[ACC_MVX + k];
        out->molPressure[k] = (out->molVirial[k] + (double)c->nMolTotal * kB * out->temperature) / out->volume;
    }
    out->pMolecular = (out->molPressure[0] + out->molPressure[1] + out->molPressure[2]) / 3.0;
    out->loop = c->loop;
    out->time = c->time;
    out->nMolecules = c->nMolTotal;
    out->nPairsListed = (int64_t)(a[ACC_NENTRIES] / 2.0 + 0.25);
    return DDCB200_OK;
}

// ---- NGLFCONSTRAINT: groups, per-bead random streams, constraints, barostat (nglfcons.cuh) ---------------------
extern "C" int ddcb200_setGroups(ddcb200_ctx *c, int ngroups, const int *type, const double *kBT, const double *tau, const double *vcm,
                                 int64_t nGlobal, const unsigned char *groupOfBead)
{
    if (!c || ngroups < 1 || !type) return fail(DDCB200_ERR_ARG, "bad arguments");
    if (ngroups > MAXGROUPS) return fail(DDCB200_ERR_CAPACITY, "more than 8 GROUP objects");
    CK(cudaSetDevice(c->device));
    c->anyLangevin = false;
    for (int g = 0; g < ngroups; g++)
    {
        if (type[g] != GROUP_FREE && type[g] != GROUP_LANGEVIN) return fail(DDCB200_ERR_ARG, "GROUP type must be FREE (0) or LANGEVIN (1)");
        c->groupType[g] = type[g];
        c->groupKBT[g] = kBT ? kBT[g] : 0.0;
        c->groupTau[g] = tau ? tau[g] : 1.0;
        for (int a = 0; a < 3; a++) c->groupVcm[g][a] = vcm ? vcm[3 * g + a] : 0.0;
        if (type[g] == GROUP_LANGEVIN)
        {
            if (!(c->groupTau[g] > 0.0)) return fail(DDCB200_ERR_ARG, "LANGEVIN group needs tau > 0");
            c->anyLangevin = true;
        }
    }
    c->nGroups = ngroups;
    if (groupOfBead)
    {
        if (c->nGlobal == 0 || nGlobal != c->nGlobal) return fail(DDCB200_ERR_STATE, "setGroups: call setBeads first (bead count differs)");
        for (int64_t i = 0; i < nGlobal; i++)
            if (groupOfBead[i] >= ngroups) return fail(DDCB200_ERR_ARG, "group index out of range");
        CK(c->groupOfBead.ensure((size_t)nGlobal));
        CK(cudaMemcpy(c->groupOfBead.p, groupOfBead, (size_t)nGlobal, cudaMemcpyHostToDevice));
    }
    else
        c->groupOfBead.release();
    return DDCB200_OK;
}

extern "C" int ddcb200_setRandom(ddcb200_ctx *c, int64_t nGlobal, const uint64_t *state, const uint32_t *multID, const uint32_t *prime)
{
    if (!c || !state || !multID || !prime) return fail(DDCB200_ERR_ARG, "null argument");
    if (c->nGlobal == 0 || nGlobal != c->nGlobal) return fail(DDCB200_ERR_STATE, "setRandom: call setBeads first (bead count differs)");
    CK(cudaSetDevice(c->device));
    std::vector<uint2> mp((size_t)nGlobal);
    for (int64_t i = 0; i < nGlobal; i++)
    {
        // lcg64_checkValue (src/lcg64.c:110-120)
        if (multID[i] > 2 || state[i] == 0 || (prime[i] & 1u) == 0) return fail(DDCB200_ERR_ARG, "bad LCG64 state (multID > 2, state 0 or even prime)");
        mp[(size_t)i] = make_uint2(multID[i], prime[i]);
    }
    CK(c->rngState.ensure((size_t)nGlobal));
    CK(c->rngMP.ensure((size_t)nGlobal));
    CK(cudaMemcpy(c->rngState.p, state, (size_t)nGlobal * sizeof(uint64_t), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c->rngMP.p, mp.data(), (size_t)nGlobal * sizeof(uint2), cudaMemcpyHostToDevice));
    c->haveRandom = true;
    return DDCB200_OK;
}

extern "C" int ddcb200_getRandom(ddcb200_ctx *c, int64_t nGlobal, uint64_t *state)
{
    if (!c || !state) return fail(DDCB200_ERR_ARG, "null argument");
    if (!c->haveRandom || nGlobal != c->nGlobal) return fail(DDCB200_ERR_STATE, "no random state of that size");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaMemcpy(state, c->rngState.p, (size_t)nGlobal * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    return DDCB200_OK;
}

extern "C" int ddcb200_setConstraints(ddcb200_ctx *c, int64_t nCons, const int64_t *atomOffset, const int *atomBead, const int64_t *pairOffset,
                                      const int *pairA, const int *pairB, const double *pairDist)
{
    if (!c || nCons < 0) return fail(DDCB200_ERR_ARG, "bad arguments");
    CK(cudaSetDevice(c->device));
    c->nCons = 0;
    if (nCons == 0) return DDCB200_OK;
    if (!atomOffset || !atomBead || !pairOffset || !pairA || !pairB || !pairDist) return fail(DDCB200_ERR_ARG, "null argument");
    if (c->nGlobal == 0) return fail(DDCB200_ERR_STATE, "setConstraints: call setBeads first");
    if (nCons > 0x7fffffff || atomOffset[nCons] > 0x7fffffff || pairOffset[nCons] > 0x7fffffff) return fail(DDCB200_ERR_CAPACITY, "too many constraints");
    std::vector<int> ao((size_t)nCons + 1), po((size_t)nCons + 1);
    std::vector<char> seen((size_t)c->nGlobal, 0);
    for (int64_t k = 0; k <= nCons; k++)
    {
        ao[(size_t)k] = (int)atomOffset[k];
        po[(size_t)k] = (int)pairOffset[k];
    }
    for (int64_t k = 0; k < nCons; k++)
    {
        const int na = ao[(size_t)k + 1] - ao[(size_t)k], np = po[(size_t)k + 1] - po[(size_t)k];
        if (na < 0 || np < 0) return fail(DDCB200_ERR_ARG, "constraint offsets must not decrease");
        if (na > CONS_MAXATOM || np > CONS_MAXPAIR) return fail(DDCB200_ERR_CAPACITY, "constraint cluster larger than 32 atoms / 48 pairs");
        for (int a = ao[(size_t)k]; a < ao[(size_t)k + 1]; a++)
        {
            const int b = atomBead[a];
            if (b < 0 || b >= c->nGlobal) return fail(DDCB200_ERR_ARG, "constraint bead index out of range");
            if (seen[(size_t)b]) return fail(DDCB200_ERR_ARG, "constraint clusters must be disjoint (a bead is in two clusters)");
            seen[(size_t)b] = 1;
        }
        for (int q = po[(size_t)k]; q < po[(size_t)k + 1]; q++)
        {
            if (pairA[q] < 0 || pairA[q] >= na || pairB[q] < 0 || pairB[q] >= na || pairA[q] == pairB[q])
                return fail(DDCB200_ERR_ARG, "constraint pair index outside its cluster");
            if (!(pairDist[q] > 0.0)) return fail(DDCB200_ERR_ARG, "constraint distance must be positive");
        }
    }
    const size_t nA = (size_t)ao[(size_t)nCons], nP = (size_t)po[(size_t)nCons];
    CK(c->consAtomOff.ensure((size_t)nCons + 1));
    CK(c->consPairOff.ensure((size_t)nCons + 1));
    CK(c->consAtomBead.ensure(nA + 1));
    CK(c->consPairA.ensure(nP + 1));
    CK(c->consPairB.ensure(nP + 1));
    CK(c->consPairDist.ensure(nP + 1));
    CK(cudaMemcpy(c->consAtomOff.p, ao.data(), ((size_t)nCons + 1) * sizeof(int), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c->consPairOff.p, po.data(), ((size_t)nCons + 1) * sizeof(int), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c->consAtomBead.p, atomBead, nA * sizeof(int), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c->consPairA.p, pairA, nP * sizeof(int), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c->consPairB.p, pairB, nP * sizeof(int), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c->consPairDist.p, pairDist, nP * sizeof(double), cudaMemcpyHostToDevice));
    if (!c->consFlag)
    {
        CK(cudaMalloc((void **)&c->consFlag, sizeof(int)));
        CK(cudaMemset(c->consFlag, 0, sizeof(int)));
    }
    c->nCons = (int)nCons;
    return DDCB200_OK;
}

extern "C" int ddcb200_nglfconstraintParms(ddcb200_ctx *c, double kBT, double P0, double beta, double tauBarostat)
{
    if (!c) return fail(DDCB200_ERR_ARG, "null ctx");
    if (beta > 0.0 && !(tauBarostat > 0.0)) return fail(DDCB200_ERR_ARG, "barostat (beta > 0) needs tauBarostat > 0");
    c->ncKBT = kBT;
    c->ncP0 = P0;
    c->ncBeta = beta;
    c->ncTau = tauBarostat;
    return DDCB200_OK;
}

// box_put(NULL, HO, &h) (src/box.c:130-150) from the host side, e.g. a barostat that runs in the caller: positions sent afterwards are
// in the new box; the neighbor list survives, its walk bound grows by the change of the box edges since the build (listSlack)
extern "C" int ddcb200_setBox(ddcb200_ctx *c, const double h[9])
{
    if (!c || !h) return fail(DDCB200_ERR_ARG, "null argument");
    double old[9];
    memcpy(old, c->prm.h, sizeof old);
    memcpy(c->prm.h, h, 9 * sizeof(double));
    const int rc = setupBox(c);
    if (rc != DDCB200_OK)
    {
        memcpy(c->prm.h, old, sizeof old);
        setupBox(c);
        return rc;
    }
    c->forcesValid = false;
    c->energyValid = false;
    return DDCB200_OK;
}

extern "C" int ddcb200_getBox(ddcb200_ctx *c, double h[9])
{
    if (!c || !h) return fail(DDCB200_ERR_ARG, "null argument");
    memcpy(h, c->prm.h, 9 * sizeof(double));
    return DDCB200_OK;
}

static GroupTab groupTabOf(const ddcb200_ctx *c)
{
    GroupTab g;
    memset(&g, 0, sizeof(g));
    g.n = c->nGroups;
    for (int k = 0; k < MAXGROUPS; k++)
    {
        g.type[k] = c->groupType[k];
        g.kBT[k] = c->groupKBT[k];
        g.tau[k] = c->groupTau[k] > 0.0 ? c->groupTau[k] : 1.0;
        g.vcx[k] = c->groupVcm[k][0];
        g.vcy[k] = c->groupVcm[k][1];
        g.vcz[k] = c->groupVcm[k][2];
    }
    return g;
}

template <int MODE>
static int launchNglfc(ddcb200_ctx *c, double halfDt, double dt, const double scale[3])
{
    ProfScope ps(c, PROF_INTEGRATE);
    const int cur = c->cur;
    const int tiles = (int)((c->nLocal + TILE - 1) / TILE);
    if (MODE & (NC_DRIFT | NC_SCALE)) c->haloDirty = c->movedSinceRef = true;
    LAUNCH(k_nglfc<MODE>, tiles, TILE, 0, c->stream)((int)c->nLocal, c->pos4[cur].p, c->vel[cur][0].p, c->vel[cur][1].p, c->vel[cur][2].p,
                                                 c->frc[0].p, c->frc[1].p, c->frc[2].p, c->massOfBead.p, c->groupOfBead.p, c->rngState.p,
                                                 c->rngMP.p, groupTabOf(c), halfDt, dt, scale[0], scale[1], scale[2], c->pc, c->kinPartial.p,
                                                 c->posBuild[0].p, c->posBuild[1].p, c->posBuild[2].p, c->dmax2, c->dispOfSlot.p,
                                                 c->nCellsBuilt > 0 ? c->cellDmax.p : nullptr, c->cellOfSlot[cur].p);
    CKL("k_nglfc");
    return DDCB200_OK;
}

template <bool FRONT>
static int launchConstraint(ddcb200_ctx *c, double dt)
{
    ProfScope ps(c, PROF_INTEGRATE);
    const int cur = c->cur;
    LAUNCH(k_constraint<FRONT>, (c->nCons + 63) / 64, 64, 0, c->stream)(c->nCons, c->consAtomOff.p, c->consAtomBead.p, c->consPairOff.p,
                                                                    c->consPairA.p, c->consPairB.p, c->consPairDist.p, c->slotOfBead.p,
                                                                    c->pos4[cur].p, c->vel[cur][0].p, c->vel[cur][1].p, c->vel[cur][2].p,
                                                                    c->massOfBead.p, dt, c->pc, c->box.hinv[0], c->box.hinv[4], c->box.hinv[8],
                                                                    c->consFlag);
    CKL("k_constraint");
    return DDCB200_OK;
}

// changeVolume (src/nglfconstraint.c:64-84) on molecularPressure (src/molecularPressure.c:57-67): new box, and the
// diagonal of hfac = h_new h_old^-1 that adjustPosn applies to the positions (updateH0, src/box.c:40-48)
static int barostat(ddcb200_ctx *c, double dt, double scale[3])
{
    int rc = fetchAccumulators(c);
    if (rc) return rc;
    const double *a = c->accHost;
    const double vol = c->box.volume;
    const double nkT = (double)c->nMolTotal * c->ncKBT;
    double pxx = (a[ACC_VXX] + a[ACC_MVX]) + nkT, pyy = (a[ACC_VYY] + a[ACC_MVY]) + nkT, pzz = (a[ACC_VZZ] + a[ACC_MVZ]) + nkT;
    pxx /= vol; pyy /= vol; pzz /= vol;       // SMATNORM
    pxx -= c->ncP0; pyy -= c->ncP0; pzz -= c->ncP0;
    const double btt = c->ncBeta * dt / c->ncTau;
    const double Pxx = 0.5 * (pxx + pyy), Pzz = pzz;     // semi-isotropic: x and y share the mean in-plane pressure
    const double lx = cbrt(1.0 + Pxx * btt), ly = cbrt(1.0 + Pxx * btt), lz = cbrt(1.0 + Pzz * btt);
    const double hi[3] = {c->box.hinv[0], c->box.hinv[4], c->box.hinv[8]};   // inverse of the old box
    double *h = c->prm.h;
    const double hn[3] = {lx * h[0], ly * h[4], lz * h[8]};
    scale[0] = hn[0] * hi[0]; scale[1] = hn[1] * hi[1]; scale[2] = hn[2] * hi[2];
    if (fabs(scale[0] - 1.0) <= 1e-14 && fabs(scale[1] - 1.0) <= 1e-14 && fabs(scale[2] - 1.0) <= 1e-14)
        scale[0] = scale[1] = scale[2] = 1.0;     // matrix_equal_tol(hfac, I, 1e-14)
    h[0] = hn[0]; h[4] = hn[1]; h[8] = hn[2];
    return setupBox(c);
}

static int nglfcSteps(ddcb200_ctx *c, int nsteps, double dt, const bool baro, const bool cons)
{
    if (!c || nsteps < 0) return fail(DDCB200_ERR_ARG, "bad arguments");
    if (c->nLocal == 0) return fail(DDCB200_ERR_STATE, "no state");
    // several ranks (src/nglfconstraint.c:510-574 runs on any task count): the per-bead random state travels with a migrating bead
    // (k_rd_pack), the barostat works on the all-reduced virial so every rank scales its box and its beads alike, constraint
    // clusters are solved on the rank that owns their molecule
    if (c->anyLangevin && !c->haveRandom) return fail(DDCB200_ERR_STATE, "LANGEVIN groups need the per-bead random state (setRandom)");
    CK(cudaSetDevice(c->device));
    int rc;
    if (!c->forcesValid || (baro && nsteps > 0 && !c->energyValid))
    {
        // firstEnergyCall (src/masters.c:579-620); the barostat needs the virial of the current forces
        rc = ddcb200_ddcenergy(c, 1);
        if (rc) return rc;
    }
    const double half = 0.5 * dt;
    double scale[3] = {1.0, 1.0, 1.0};
    for (int s = 0; s < nsteps; s++)
    {
        if (baro)
        {
            rc = barostat(c, dt, scale);
            if (rc) return rc;
        }
        if (!cons)
        {
            // one pass: BACK update of the previous step (inside this call), barostat scaling, FRONT update, drift
            if (s == 0) rc = baro ? launchNglfc<NC_SCALE | NC_FRONT | NC_DRIFT>(c, half, dt, scale) : launchNglfc<NC_FRONT | NC_DRIFT>(c, half, dt, scale);
            else rc = baro ? launchNglfc<NC_BACK | NC_SCALE | NC_FRONT | NC_DRIFT>(c, half, dt, scale)
                           : launchNglfc<NC_BACK | NC_FRONT | NC_DRIFT>(c, half, dt, scale);
            if (rc) return rc;
        }
        else
        {
            rc = baro ? launchNglfc<NC_SCALE | NC_FRONT>(c, half, dt, scale) : launchNglfc<NC_FRONT>(c, half, dt, scale);
            if (rc) return rc;
            rc = launchConstraint<true>(c, dt);
            if (rc) return rc;
            rc = launchNglfc<NC_DRIFT>(c, half, dt, scale);
            if (rc) return rc;
        }
        c->loop++;
        c->time += dt;
        const bool last = s == nsteps - 1;
        rc = ddcb200_ddcenergy(c, (baro || last) ? 1 : 0);
        if (rc) return rc;
        if (cons)
        {
            rc = launchNglfc<NC_BACK>(c, half, dt, scale);
            if (rc) return rc;
            rc = launchConstraint<false>(c, dt);
            if (rc) return rc;
        }
        else if (last)
        {
            rc = launchNglfc<NC_BACK | NC_KE>(c, half, dt, scale);
            if (rc) return rc;
        }
    }
    if (nsteps > 0)
    {
        if (cons)
        {
            rc = launchNglfc<NC_KE>(c, half, dt, scale);
            if (rc) return rc;
        }
        ProfScope ps(c, PROF_REDUCE);
        rc = reduceCols(c, c->kinPartial.p, (int)((c->nLocal + TILE - 1) / TILE), 7, c->colMap.p + 19);
        if (rc) return rc;
        c->kineticValid = true;
    }
    return DDCB200_OK;
}

extern "C" int ddcb200_nglfconstraint(ddcb200_ctx *c, int nsteps, double dt)
{
    if (!c) return fail(DDCB200_ERR_ARG, "null ctx");
    return nglfcSteps(c, nsteps, dt, c->ncBeta > 0.0, c->nCons > 0);
}

// number of constraint clusters that hit the iteration cap since the context was created (the reference prints
// "too many contraint iterations" and continues, src/nglfconstraint.c:246)
extern "C" int64_t ddcb200_constraintFailures(ddcb200_ctx *c)
{
    if (!c || !c->consFlag) return 0;
    int v = 0;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    cudaMemcpy(&v, c->consFlag, sizeof(int), cudaMemcpyDeviceToHost);
    return v;
}

// ---- parity hooks -------------------------------------------------------------------------
extern "C" int ddcb200_getCells(ddcb200_ctx *c, int *cellOfBead, int dims[3], double geom[9])
{
    if (!c || !cellOfBead) return fail(DDCB200_ERR_ARG, "null argument");
    if (!c->listValid) return fail(DDCB200_ERR_STATE, "no list built");
    CK(cudaSetDevice(c->device));
    int rcl = refreshLocals(c);
    if (rcl) return rcl;
    const int n = (int)c->nIon;
    std::vector<int> cell(n), bead(n);
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaMemcpy(cell.data(), c->cellOfSlot[c->cur].p, n * sizeof(int), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(bead.data(), c->beadOfSlot[c->cur].p, n * sizeof(int), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(c->gridHost, c->grid, sizeof(GridDev), cudaMemcpyDeviceToHost));
    std::map<int, int> pos;
    for (size_t i = 0; i < c->hLocalBeads.size(); i++) pos[c->hLocalBeads[i]] = (int)i;
    for (int s = 0; s < n; s++)
    {
        auto it = pos.find(bead[s]);
        if (it != pos.end()) cellOfBead[it->second] = cell[s];
    }
    if (dims)
        for (int a = 0; a < 3; a++) dims[a] = c->gridHost->n[a];
    if (geom)
        for (int a = 0; a < 3; a++)
        {
            geom[a] = c->gridHost->mn[a];
            geom[3 + a] = c->gridHost->mx[a];
            geom[6 + a] = c->gridHost->d[a];
        }
    return DDCB200_OK;
}

extern "C" int64_t ddcb200_getPairs(ddcb200_ctx *c, int64_t capacity, int *beadI, int *beadJ, int *pruned)
{
    if (!c) return fail(DDCB200_ERR_ARG, "null ctx");
    if (!c->listValid) return fail(DDCB200_ERR_STATE, "no list built");
    if (cudaSetDevice(c->device) != cudaSuccess) return fail(DDCB200_ERR_CUDA, "cudaSetDevice");
    const int n = (int)c->nLocal, nPad = (int)c->nPad;   // rows exist for the local slots [0, nLocal)
    std::vector<int> cnt(n), bead((size_t)c->nIon);
    cudaStreamSynchronize(c->stream);
    if (cudaMemcpy(cnt.data(), c->nbrCount.p, n * sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess ||
        cudaMemcpy(bead.data(), c->beadOfSlot[c->cur].p, c->nIon * sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess)
        return fail(DDCB200_ERR_CUDA, "getPairs copy");
    int64_t np = 0;
    auto emit = [&](int i, int j, bool excl) {
        const int bi = bead[i], bj = bead[j];
        if (c->hGid[bi] < c->hGid[bj])
        {
            if (np < capacity && beadI && beadJ)
            {
                beadI[np] = bi;
                beadJ[np] = bj;
                if (pruned) pruned[np] = excl ? 1 : 0;
            }
            np++;
        }
    };
    // rows of the one-pass build are two segments: the first cum[0] entries from the front, the others from the end of the row's
    // allocation backwards (k_nbr_exact2); rows of k_nbr_exact run forward
    const int farTop = c->nbrCap - 1;
    int maxc = 0;
    for (int i = 0; i < n; i++) maxc = std::max(maxc, cnt[i]);
    const size_t nrows = farTop >= 0 ? (size_t)c->nbrCap : (size_t)maxc;
    std::vector<uint32_t> rows(nrows * nPad);
    std::vector<uint16_t> nearN;
    if (maxc && cudaMemcpy(rows.data(), c->nbr.p, rows.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost) != cudaSuccess)
        return fail(DDCB200_ERR_CUDA, "getPairs copy");
    if (farTop >= 0)
    {
        nearN.resize((size_t)n);
        if (cudaMemcpy(nearN.data(), c->nbrCum.p, (size_t)n * sizeof(uint16_t), cudaMemcpyDeviceToHost) != cudaSuccess)
            return fail(DDCB200_ERR_CUDA, "getPairs copy");
    }
    for (int i = 0; i < n; i++)
        for (int k = 0; k < cnt[i]; k++)
        {
            const int at = (farTop >= 0 && k >= (int)nearN[(size_t)i]) ? farTop - (k - (int)nearN[(size_t)i]) : k;
            const uint32_t e = rows[(size_t)at * nPad + i];
            emit(i, (int)(e & 0x07ffffffu), (e & EXCL_BIT) != 0u);
        }
    return np;
}

extern "C" int ddcb200_pairSetHash(ddcb200_ctx *c, uint64_t out[6])
{
    if (!c || !out) return fail(DDCB200_ERR_ARG, "null argument");
    if (!c->listValid) return fail(DDCB200_ERR_STATE, "no list built");
    CK(cudaSetDevice(c->device));
    struct Scratch
    {
        DevBuf<unsigned long long> h;
        ~Scratch() { h.release(); }
    } sc;
    CK(sc.h.ensure(8));
    CK(cudaMemsetAsync(sc.h.p, 0, 8 * sizeof(unsigned long long), c->stream));
    LAUNCH(k_pair_hash, (int)((c->nLocal + 255) / 256), 256, 0, c->stream)((int)c->nLocal, (int)c->nPad, c->nbr.p, c->nbrCount.p, c->beadOfSlot[c->cur].p,
                                                                     c->gidOfBead.p, c->nbrCum.p, c->nbrCap - 1, sc.h.p);
    CKL("k_pair_hash");
    CK(cudaMemcpyAsync(out, sc.h.p, 6 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return DDCB200_OK;
}

extern "C" int ddcb200_timerRecord(ddcb200_ctx *c, int which)
{
    if (!c || which < 0 || which > 3) return fail(DDCB200_ERR_ARG, "bad timer slot");
    CK(cudaSetDevice(c->device));
    if (!c->timer[which]) CK(cudaEventCreate(&c->timer[which]));
    CK(cudaEventRecord(c->timer[which], c->stream));
    return DDCB200_OK;
}

extern "C" int ddcb200_timerElapsed(ddcb200_ctx *c, int from, int to, double *ms)
{
    if (!c || !ms || from < 0 || from > 3 || to < 0 || to > 3 || !c->timer[from] || !c->timer[to]) return fail(DDCB200_ERR_ARG, "bad timer slot");
    CK(cudaSetDevice(c->device));
    CK(cudaEventSynchronize(c->timer[to]));
    float t = 0;
    CK(cudaEventElapsedTime(&t, c->timer[from], c->timer[to]));
    *ms = t;
    return DDCB200_OK;
}

extern "C" int64_t ddcb200_kernelLaunches(ddcb200_ctx *c) { return c ? c->kernelLaunches : 0; }

extern "C" int64_t ddcb200_lastListBuild(ddcb200_ctx *c) { return c ? c->lastBuildLoop : -1; }

extern "C" int ddcb200_kineticByClass(ddcb200_ctx *c, int bySpecies, int nClasses, double *out12)
{
    if (!c || !out12 || nClasses < 1) return fail(DDCB200_ERR_ARG, "kineticByClass: bad arguments");
    if (c->nLocal == 0) return fail(DDCB200_ERR_STATE, "no state");
    if (bySpecies ? nClasses != c->nspecies : nClasses != std::max(1, c->nGroups)) return fail(DDCB200_ERR_ARG, "kineticByClass: class count does not match the deck");
    CK(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    if (c->pendingKick2) return fail(DDCB200_ERR_STATE, "kineticByClass: velocities are mid-step");
    struct Scratch
    {
        DevBuf<int> cls;
        DevBuf<double> out;
        ~Scratch() { cls.release(); out.release(); }
    } sc;
    std::vector<int> h((size_t)c->nGlobal);
    if (bySpecies) h = c->hSpecies;
    else if (c->groupOfBead.p)
    {
        std::vector<unsigned char> g((size_t)c->nGlobal);
        CK(cudaMemcpy(g.data(), c->groupOfBead.p, (size_t)c->nGlobal, cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < g.size(); i++) h[i] = g[i];
    }
    else std::fill(h.begin(), h.end(), 0);
    CK(sc.cls.ensure((size_t)c->nGlobal + 1));
    CK(sc.out.ensure((size_t)nClasses * 12));
    CK(cudaMemcpyAsync(sc.cls.p, h.data(), (size_t)c->nGlobal * sizeof(int), cudaMemcpyHostToDevice, st));
    const int cur = c->cur;
    LAUNCH(k_kinetic_classes, nClasses, 256, 0, st)((int)c->nIon, c->pos4[cur].p, c->vel[cur][0].p, c->vel[cur][1].p, c->vel[cur][2].p, c->massOfBead.p,
                                                sc.cls.p, sc.out.p);
    CKL("k_kinetic_classes");
    CK(cudaMemcpyAsync(out12, sc.out.p, (size_t)nClasses * 12 * sizeof(double), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return DDCB200_OK;
}

extern "C" int ddcb200_pairCorrelation(ddcb200_ctx *c, int nBins, double rmin, double delta, int logScale, double rmax,
                                       unsigned long long *counts, unsigned long long *nAtoms)
{
    if (!c || !counts || !nAtoms || nBins < 1 || !(delta > 0.0) || !(rmax > 0.0) || (logScale && !(rmin > 0.0)))
        return fail(DDCB200_ERR_ARG, "pairCorrelation: bad arguments");
    if (c->nLocal == 0) return fail(DDCB200_ERR_STATE, "no state");
    CK(cudaSetDevice(c->device));
    int rc;
    if (!c->listValid)
    {
        rc = ddcb200_constructList(c);
        if (rc) return rc;
    }
    if (c->nranks > 1 && c->haloDirty)
    {
        rc = haloExchange(c, c->stream);
        if (rc) return rc;
    }
    cudaStream_t st = c->stream;
    // candidates: the cells within `reach` cells of a bead's build-time cell.  A pair with r < rmax now was within
    // rmax + 2 dmax + box slack at the build, and a cell is at least d_min wide: reach = ceil of that over d_min
    unsigned long long dbits2[3] = {0, 0, 0};      // [2]: a bound of the displacement between the build and the last prune (k_rebase)
    CK(cudaMemcpyAsync(dbits2, c->dmax2, sizeof dbits2, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(c->gridHost, c->grid, sizeof(GridDev), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    const unsigned long long dbits = std::max(dbits2[0], dbits2[1]);     // locals and ghosts
    double d2, dBase;
    memcpy(&d2, &dbits, sizeof d2);
    memcpy(&dBase, &dbits2[2], sizeof dBase);
    const double need = (rmax + 2.0 * (sqrt(d2) + dBase) + c->pc.listSlack) * (1.0 + 1e-9);
    // cell edges in length units: d (fraction of the box span) x span
    const double edge[3] = {c->gridHost->d[0] * c->box.spanx, c->gridHost->d[1] * c->box.spany, c->gridHost->d[2] * c->box.spanz};
    const double dmin = std::min(edge[0], std::min(edge[1], edge[2]));
    if (!(dmin > 0.0)) return fail(DDCB200_ERR_STATE, "pairCorrelation: no cell grid");
    if (c->nranks > 1 && need > sqrt(c->box.rlist2))
        return fail(DDCB200_ERR_STATE, "pairCorrelation on several ranks: rmax plus twice the largest displacement exceeds the ghost shell (list radius)");
    const int reach = std::max(1, (int)ceil(need / dmin));
    if (rmax > 0.5 * std::min(c->box.hxx, std::min(c->box.hyy, c->box.hzz)))
        return fail(DDCB200_ERR_ARG, "pairCorrelation: rmax exceeds half the shortest box edge");
    const int ns = c->nspecies;
    const size_t nh = (size_t)nBins * (size_t)(ns * (ns + 1) / 2);
    struct Scratch
    {
        DevBuf<unsigned long long> hist;
        DevBuf<int> spec;
        ~Scratch() { hist.release(); spec.release(); }      // also on the early returns of CK
    } scratch;
    DevBuf<unsigned long long> &hist = scratch.hist;
    DevBuf<int> &spec = scratch.spec;
    CK(hist.ensure(nh + (size_t)ns));
    CK(spec.ensure((size_t)c->nGlobal + 1));
    CK(cudaMemsetAsync(hist.p, 0, (nh + (size_t)ns) * sizeof(unsigned long long), st));
    CK(cudaMemcpyAsync(spec.p, c->hSpecies.data(), (size_t)c->nGlobal * sizeof(int), cudaMemcpyHostToDevice, st));
    const int cur = c->cur;
    LAUNCH(k_paircorr, (int)(c->nPad / TILE), TILE, 0, st)((int)c->nIon, c->pos4[cur].p, c->cellOfSlot[cur].p, c->cellStart.p, c->grid, reach,
                                                       c->gidOfBead.p, spec.p, c->box, rmax * rmax, rmin, delta, logScale,
                                                       logScale ? log10(rmin) : 0.0, nBins, ns, hist.p, hist.p + nh);
    CKL("k_paircorr");
    CK(cudaMemcpyAsync(counts, hist.p, nh * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(nAtoms, hist.p + nh, (size_t)ns * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return DDCB200_OK;
}

// measurement hook: the state of the pruned rows and what the walk of the next force evaluation at the current positions would visit
extern "C" int ddcb200_pruneInfo(ddcb200_ctx *c, int64_t info[6])
{
    if (!c || !info) return fail(DDCB200_ERR_ARG, "bad arguments");
    info[0] = c->pruneEvery;
    info[1] = c->pruneValid ? c->sincePrune : -1;
    info[2] = info[3] = info[4] = 0;
    info[5] = c->totalEntries;
    if (!c->pruneValid || !c->listValid || c->pruneEvery <= 0) return DDCB200_OK;
    CK(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    PruneArgs pr;
    pruneArgsOf(c, pr);
    unsigned long long *d = nullptr, h[3] = {0, 0, 0};
    CK(cudaMalloc((void **)&d, sizeof h));
    cudaError_t e = cudaMemsetAsync(d, 0, sizeof h, st);
    if (e == cudaSuccess)
    {
        const int nLocal = (int)c->nLocal, cur = c->cur;
        const bool cellWalk = c->walkPerCell && c->walkPerBead && c->nCellsBuilt > 0;
        const int withGhosts = c->nranks > 1 ? 1 : 0;
        if (cellWalk)
        {
            LAUNCH(k_nbr_dmax, (c->nCellsBuilt + 127) / 128, 128, 0, st)(c->grid, c->cellDmax.p, 0, c->nbrDmax.p);
            if (withGhosts) LAUNCH(k_nbr_dmax, (c->nCellsBuilt + 127) / 128, 128, 0, st)(c->grid, c->cellDmax.p, 1, c->nbrDmax.p + c->nCellsBuilt);
        }
        LAUNCH(k_prune_stats, (nLocal + TILE - 1) / TILE, TILE, 0, st)(nLocal, (int)c->nPad, c->pos4[cur].p, c->nbrCum.p, c->dmax2, withGhosts,
                                                                 c->walkPerBead ? c->dispOfSlot.p : nullptr, c->pc,
                                                                 cellWalk ? c->nbrDmax.p + (withGhosts ? c->nCellsBuilt : 0) : nullptr,
                                                                 c->cellOfSlot[cur].p, pr, d);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(h, d, sizeof h, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(d);
    CK(e);
    info[2] = (int64_t)h[0];
    info[3] = (int64_t)h[1];
    info[4] = (int64_t)h[2];
    return DDCB200_OK;
}

extern "C" int ddcb200_listBuildInfo(ddcb200_ctx *c, int *variant, double ms[2])
{
    if (!c) return fail(DDCB200_ERR_ARG, "null ctx");
    if (variant) *variant = 1;      // one build: fp32 candidate filter + exact fp64 pass
    if (ms)
    {
        ms[0] = c->listBuildMs;     // device time of the last build
        ms[1] = 0.0;
    }
    return DDCB200_OK;
}

extern "C" int ddcb200_profile(ddcb200_ctx *c, int enable)
{
    if (!c) return fail(DDCB200_ERR_ARG, "null ctx");
    c->prof = enable != 0;
    return DDCB200_OK;
}

extern "C" int ddcb200_profileRead(ddcb200_ctx *c, double ms[8], int64_t launches[8], int reset)
{
    if (!c) return fail(DDCB200_ERR_ARG, "null ctx");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    for (auto &pe : c->pending)
    {
        float t = 0;
        cudaEventElapsedTime(&t, pe.a, pe.b);
        c->profMs[pe.slot] += t;
        c->evPool.push_back(pe.a);
        c->evPool.push_back(pe.b);
    }
    c->pending.clear();
    for (int k = 0; k < PROF_N; k++)
    {
        if (ms) ms[k] = c->profMs[k];
        if (launches) launches[k] = c->profLaunch[k];
        if (reset)
        {
            c->profMs[k] = 0;
            c->profLaunch[k] = 0;
        }
    }
    return DDCB200_OK;
}

extern "C" int ddcb200_ncclUniqueId(unsigned char id[128])
{
    if (!id) return fail(DDCB200_ERR_ARG, "null argument");
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId u;
    CKN(ncclGetUniqueId(&u));
    memcpy(id, &u, 128);
    return DDCB200_OK;
}

// ddc_init (src/ddc.c:61-117: lattice lx ly lz of domain centres) + the communicator
extern "C" int ddcb200_ddcInit(ddcb200_ctx *c, int rank, int nranks, int lx, int ly, int lz, const unsigned char id[128])
{
    if (!c || !id || nranks < 1 || rank < 0 || rank >= nranks) return fail(DDCB200_ERR_ARG, "bad rank arguments");
    if (nranks > DDC_MAXRANKS) return fail(DDCB200_ERR_ARG, "at most 16 ranks (one NVSwitch box)");
    if (lx < 1 || ly < 1 || lz < 1 || lx * ly * lz != nranks) return fail(DDCB200_ERR_ARG, "DDC lx*ly*lz must equal the number of ranks");
    if (c->nccl) return fail(DDCB200_ERR_STATE, "ddcInit called twice");
    CK(cudaSetDevice(c->device));
    c->rank = rank;
    c->nranks = nranks;
    c->lat[0] = lx; c->lat[1] = ly; c->lat[2] = lz;
    if (nranks == 1) return DDCB200_OK;
    ncclUniqueId u;
    memcpy(&u, id, 128);
    ncclComm_t comm;
    CKN(ncclCommInitRank(&comm, nranks, u, rank));
    c->nccl = (void *)comm;
    {
        // the halo stream outranks the compute stream: its small kernels (pack, NCCL's send/recv, unpack) take the next free
        // CTA slots instead of queueing behind the pair kernel's grid, so the exchange really runs beside the interior rows
        int least = 0, greatest = 0;
        CK(cudaDeviceGetStreamPriorityRange(&least, &greatest));
        CK(cudaStreamCreateWithPriority(&c->streamH, cudaStreamNonBlocking, greatest));
    }
    CK(cudaEventCreateWithFlags(&c->evPos, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&c->evHalo, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&c->evBoundary, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&c->evBonded, cudaEventDisableTiming));
    CK(cudaStreamCreateWithFlags(&c->streamB, cudaStreamNonBlocking));
    CK(cudaMalloc((void **)&c->boxes, sizeof(DdcBoxes)));
    CK(cudaMalloc((void **)&c->ddcWork, sizeof(DdcWork)));
    CK(cudaMallocHost((void **)&c->ddcWorkInit, sizeof(DdcWork)));
    memset(c->ddcWorkInit, 0, sizeof(DdcWork));
    for (int k = 0; k < 3; k++) ((DdcWork *)c->ddcWorkInit)->boxEnc[k] = ~0ull;     // running minima start at the top
    CK(cudaMalloc((void **)&c->ddcRow, DDC_ROW * sizeof(int)));
    CK(cudaMalloc((void **)&c->ddcRowAll, DDC_MAXRANKS * DDC_ROW * sizeof(int)));
    CK(cudaMallocHost((void **)&c->ddcRowHost, DDC_MAXRANKS * DDC_ROW * sizeof(int)));
    CK(cudaMalloc((void **)&c->ddcBox6, 6 * sizeof(double)));
    CK(cudaMalloc((void **)&c->ddcBoxAll, DDC_MAXRANKS * 6 * sizeof(double)));
    c->listValid = false;
    return DDCB200_OK;
}

// CPU restatement of the domain classification, running the same __host__ __device__ predicates as the kernels
// (k_rd_dest, k_rd_bbox, k_rd_ghostmask) on a replicated copy of the positions.  Used by the CPU tests; not part of the step.
extern "C" int ddcb200_ddcPlan(const double h[9], int lx, int ly, int lz, double rlist, int64_t nGlobal, const double *rx, const double *ry,
                               const double *rz, const int *ownerBead, int rank, int *owner, uint32_t *mask)
{
    if (!h || !rx || !ry || !rz || !owner || !mask || nGlobal <= 0) return fail(DDCB200_ERR_ARG, "null argument");
    DdcGeom g;
    g.nranks = lx * ly * lz;
    g.me = rank;
    if (g.nranks < 1 || g.nranks > DDC_MAXRANKS || rank < 0 || rank >= g.nranks) return fail(DDCB200_ERR_ARG, "bad lattice or rank");
    g.lat[0] = lx; g.lat[1] = ly; g.lat[2] = lz;
    g.L[0] = h[0]; g.L[1] = h[4]; g.L[2] = h[8];
    for (int a = 0; a < 3; a++) g.hL[a] = 0.5 * g.L[a];
    g.rlist2 = rlist * rlist * (1.0 + 1e-9);
    DdcBoxes bx;
    for (int r = 0; r < DDC_MAXRANKS; r++)
        for (int a = 0; a < 3; a++) { bx.lo[r][a] = 1e300; bx.hi[r][a] = -1e300; }
    for (int64_t b = 0; b < nGlobal; b++)
    {
        const int ob = ownerBead ? ownerBead[b] : (int)b;
        const int r = ddcBrickOf(rx[ob], ry[ob], rz[ob], g);
        owner[b] = r;
        double cc[3];
        ddcBrickCentre(r, g, cc);
        const double p[3] = {rx[b], ry[b], rz[b]};
        for (int a = 0; a < 3; a++)
        {
            const double d = ddcMinImg(p[a] - cc[a], g.L[a], g.hL[a]);
            if (d < bx.lo[r][a]) bx.lo[r][a] = d;
            if (d > bx.hi[r][a]) bx.hi[r][a] = d;
        }
    }
    for (int64_t b = 0; b < nGlobal; b++)
    {
        uint32_t m = 0u;
        if (owner[b] == rank)
        {
            m = 0x80000000u;
            for (int pp = 0; pp < g.nranks; pp++)
                if (pp != rank && ddcNear(rx[b], ry[b], rz[b], pp, g, bx)) m |= 1u << pp;
        }
        else if (ddcNear(rx[b], ry[b], rz[b], rank, g, bx)) m = 0x40000000u | (1u << owner[b]);
        mask[b] = m;
    }
    return DDCB200_OK;
}

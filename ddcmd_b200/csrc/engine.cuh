// engine.cuh - device data model of the B200-native Martini MD step.
//
// Layout in HBM (see DESIGN.md "Data layout"):
//   * "slot" order = beads sorted by (reference GeomBox cell, input index); rebuilt with the
//     neighbor list every DDC updateRate steps.  All per-step streams are in slot order so a
//     warp touches consecutive memory.
//   * pos4[slot] = {x, y, z, w}: 32-byte records so one j-bead is one DRAM/L2 sector; w carries
//     packed ints (LJ type, charge index, input index) reinterpreted as a double.
//   * vel / frc are SoA (3 arrays each).
//   * neighbor list: full (both directions) per-bead list, transposed (entry k of slot i at
//     k*nPad + i) so a warp reads 32 consecutive ints; entries are ordered by build-time
//     distance so the lanes of a warp leave the cutoff sphere together.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>
#include "../../include/ddcmd_b200.h"

// Kernel-launch and dynamic-shared-memory spellings.  The no-GPU test build (tests/cpu_emu, test infrastructure only)
// compiles these sources with g++ against a shim that defines both macros itself; the product always takes this branch.
#ifndef DDCB200_EMU
#define LAUNCH(k, ...) k<<<__VA_ARGS__>>>
#define EXTERN_SHARED(type, name) extern __shared__ type name[]
#endif

#define TILE 128           // beads per CTA in the per-step kernels
#define NBINS 2            // segments of a row: listed closer / farther than the near edge (PairConst::binEdge)
#define EXCL_BIT 0x80000000u

struct BoxConst
{
    double hxx, hyy, hzz;      // box edges
    double hhx, hhy, hhz;      // 0.5*h (PreduceOrthorhombicB7_OneLatticeReduction, src/preduce.c:147-160)
    double hinv[9];            // matinv cofactor inverse (src/three_algebra.c:37-64)
    double cx, cy, cz;         // GeomBox centre
    double R2cut;              // 0.25*minspan^2 (src/pairlist.c:227)
    double rlist2;             // (rcut0+deltaR)^2 (src/pairlist.c:226)
    double rc2;                // rmax^2
    double rcutGeom;           // max(rcut0+deltaR, minBoxSide)
    double spanx, spany, spanz;  // computeBoxSpan (src/geom.c:478-511)
    double binEdge2[NBINS - 1];  // r^2 edges of the ordering bins
    double volume;
};

struct GridDev
{
    double mn[3], mx[3], d[3];
    int n[3];
    int ncell;
    int error;       // sticky error flags (bit0: neighbor capacity overflow, bit1: cell overflow)
    int maxCount;    // max full-list length over beads
    int maxRaw;      // max candidate count over beads (fp32 filter pass)
    int nInterior;   // several ranks: k_pair tiles whose rows touch no ghost slot
    int bondTotal;   // bonded (term, endpoint) entries of the resident local beads
    int bondTerms;   // bonded terms (and restraints) whose role-0 bead is local
    unsigned long long totalEntries;
};

struct PairConst
{
    double rc2, R2cut, hxx, hyy, hzz, hhx, hhy, hhz;
    double ihx, ihy, ihz;        // reciprocal box edges
    double keR, krf, crf;
    double rmax;
    double binEdge[NBINS - 1];   // r edges of the build-time distance bins
    double listSlack;            // |h - h_build| (barostat): added to the displacement bound of the list walk; 0 for a fixed box
    int ntypes;
};

// packed w of pos4: [0,8) LJ type, [8,16) charge index, [16,32) low 16 bits of the molecule id (gid >> 32),
// [32,64) input (bead) index
__host__ __device__ inline uint64_t packW(int lj, int qi, uint32_t mol, uint32_t bead)
{
    return (uint64_t)(lj & 0xff) | ((uint64_t)(qi & 0xff) << 8) | ((uint64_t)(mol & 0xffff) << 16) | ((uint64_t)bead << 32);
}

struct Term
{
    int i, j, k, l;      // bead (input) indices; k, l = -1 when unused
    double p0, p1, p2;
    int kind, pad;
};

struct alignas(16) BondRec   // one local bonded term, endpoints resolved to slots at the list build
{
    int s[4];            // slots of the term's beads (unused: 0)
    double p0, p1, p2;
    short kind;          // term kind 0..5, 6 = restraint, -1 = an endpoint is not resident: skip
    short role;          // 0
    unsigned short q, n; // n = number of beads of the term (forces staged)
};

enum { PROF_INTEGRATE = 0, PROF_PAIR, PROF_BONDED, PROF_LIST, PROF_REDUCE, PROF_HALO, PROF_PAIR_PRUNE, PROF_N = 8 };      // PAIR_PRUNE: the pair launches that also write the pruned rows
enum { ACC_ELJ = 0, ACC_EELE, ACC_VXX, ACC_VYY, ACC_VZZ, ACC_VXY, ACC_VXZ, ACC_VYZ,
       ACC_EBOND, ACC_EANGLE, ACC_ETORS, ACC_EIMPR, ACC_EREST, ACC_RK,
       ACC_TXX, ACC_TYY, ACC_TZZ, ACC_TXY, ACC_TXZ, ACC_TYZ, ACC_MVX, ACC_MVY, ACC_MVZ, ACC_NENTRIES, ACC_N = 24 };

template <typename T>
struct DevBuf
{
    T *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t n)
    {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = n + n / 4 + 256;      // a quarter of headroom: bead and ghost counts drift from one re-domain to the next
        cudaError_t e = cudaMalloc((void **)&p, want * sizeof(T));
        if (e == cudaSuccess) cap = want;
        return e;
    }
    // like ensure(), but what the buffer holds survives the reallocation
    cudaError_t grow(size_t n, cudaStream_t st)
    {
        if (n <= cap) return cudaSuccess;
        const size_t want = n + n / 4 + 256;
        T *q = nullptr;
        cudaError_t e = cudaMalloc((void **)&q, want * sizeof(T));
        if (e != cudaSuccess) return e;
        if (p)
        {
            e = cudaMemcpyAsync(q, p, cap * sizeof(T), cudaMemcpyDeviceToDevice, st);
            if (e == cudaSuccess) e = cudaStreamSynchronize(st);
            cudaFree(p);
            if (e != cudaSuccess)
            {
                cudaFree(q);
                p = nullptr;
                cap = 0;
                return e;
            }
        }
        p = q;
        cap = want;
        return cudaSuccess;
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

struct ddcb200_ctx
{
    ddcb200_params prm;
    BoxConst box;
    PairConst pc;
    int device = 0;
    cudaStream_t stream = nullptr;
    int numSM = 148;

    // static tables
    int ntypes = 0;
    DevBuf<double2> ljTab;        // {c6, c12} = {4 eps sigma^6, 4 eps sigma^12}
    DevBuf<double> shiftTab;
    int nspecies = 0;
    std::vector<int> hSpecLJ, hSpecQi;
    std::vector<double> hSpecQ, hSpecMass, hQtab;
    DevBuf<double> qTab;          // unique charges
    DevBuf<double> massOfBead;    // by input index  (mass)
    DevBuf<uint64_t> wOfBead;     // packed w by input index
    DevBuf<uint64_t> gidOfBead;
    std::vector<uint64_t> hGid;
    std::vector<int> hSpecies;
    int64_t nGlobal = 0;

    // exclusions
    int nMolTypes = 0;
    std::vector<int> hMolTypeOfSpecies, hMolTypeNSpecies;
    DevBuf<int> molTypeOfBead;    // -1: none
    DevBuf<int> molTypeSingle;    // per mol type: 1 if nSpecies == 1
    DevBuf<int> bpairOffset;
    DevBuf<uint32_t> bpairKey;    // sorted (min<<16|max) keys per mol type
    bool haveExcl = false;

    // bonded terms (bead-index form) and the per-bead gather lists of k_bonded (bonded.cuh)
    int64_t nTerms = 0;
    DevBuf<Term> termsBead;
    std::vector<Term> hTerms;
    int64_t nRestr = 0;
    DevBuf<int> restrBead;
    std::vector<int> hRestrBead;
    bool bondCsrDirty = true;
    DevBuf<int> bondCsrOff;       // nGlobal + 1: entries of bead b = bondEnt[bondCsrOff[b] .. bondCsrOff[b + 1])
    DevBuf<uint32_t> bondEnt;     // (term << 2) | role of the bead in the term; term >= nTerms = restraint term - nTerms
    DevBuf<BondRec> bondRec;      // the local terms in the slot order of their role-0 beads, refreshed at every list build
    DevBuf<int> bondCount, bondStart, bondCount0, bondStart0, scanBlocks;
    DevBuf<int> termMap;          // term (restraint: nTerms + index) -> local term index, -1 elsewhere
    DevBuf<int> bondStageIdx;     // per local bead entry: 4 * local term + role
    DevBuf<double> bondStage;     // forces of every local term on its (up to 4) beads: 12 doubles per term
    int nBondRec = 0, nBondTerms = 0;   // entries / terms of the resident local beads (set at the list build)
    DevBuf<double> restrParm;     // 7 doubles: frac0[3], kb, fc[3]
    int restrOrigin = 0;

    // molecules (molecular virial)
    int64_t nMol = 0, nMolTotal = 0;
    DevBuf<int64_t> molOffset;
    DevBuf<int> molBeads;

    // dynamic state, slot order (double buffered for the re-sort)
    int64_t nLocal = 0, nIon = 0, nPad = 0;
    DevBuf<double4> pos4[2];
    DevBuf<double> vel[2][3];
    DevBuf<double> frc[3];
    DevBuf<int> beadOfSlot[2];
    int cur = 0;
    DevBuf<int> slotOfBead;       // input index -> slot (-1 if absent)
    std::vector<int> hLocalBeads; // order of sendState

    // cells
    DevBuf<int> cellOfSlot[2], rank0, cellCount, cellStart, member, perm;
    DevBuf<uint64_t> orderKey;    // (sub-cell Morton key, bead) : slot order inside a cell
    DevBuf<float4> pos32;         // fp32 copy of the build-time positions (candidate filter)
    DevBuf<double> posBuild[3];   // build-time positions, slot order (displacement bound)
    unsigned long long *dmax2 = nullptr;   // device: [0] bits of the max squared displacement of a LOCAL bead since the build (or the last
                                           // prune), [1] of a ghost, [2] bits of a bound of the displacement between the build and the last prune
    // pruned rows (k_pair2 MODE 1 / 2, DDCB200_PRUNE=<every>[,<margin>]; 0 = off)
    int pruneEvery = 4;           // steps between prunes (measured best with its default margin: profiles/r02p_prune_ab.jsonl)
    double pruneMargin = 0.0;     // entries closer than rmax + pruneMargin x deltaR are kept (0: 1.4 x pruneEvery / updateRate)
    int sincePrune = 0;           // force evaluations since the pruned rows were written
    bool pruneValid = false;      // the pruned rows belong to the current list and reference positions
    bool rebased = false;         // the reference positions are no longer those of the build (bin-limited walks are off until the next build)
    bool movedSinceRef = false;   // positions changed since the reference positions were taken
    DevBuf<uint16_t> pruneCount;  // entries per pruned row
    bool walkPerBead = true;      // DDCB200_WALK
    bool walkPerCell = true;      // DDCB200_WALK=cell (default): d_j bounded by the bead's stencil cells instead of the whole system
    DevBuf<unsigned long long> cellDmax;   // [2 ncell]: largest squared displacement since the build per cell (local part, ghost part)
    DevBuf<unsigned long long> nbrDmax;    // [2][ncell]: the maximum over each cell's stencil, without / with the ghost parts
    int nCellsBuilt = 0;
    int pairVariant = 2;          // DDCB200_PAIR: index into the k_pair2 instantiations of api.cu
    int bondedCap = 1;            // DDCB200_BONDED: register cap of k_bonded as CTAs per SM; 1 = none (84 registers): best since a thread evaluates a whole term
    DevBuf<float> dispOfSlot;     // each local bead's own displacement since the build, rounded up (0 right after a build)
    DevBuf<double> mmPartial;
    GridDev *grid = nullptr;      // device
    GridDev *gridHost = nullptr;  // pinned

    // neighbor list
    DevBuf<uint32_t> nbrRaw, nbr;
    DevBuf<int> nbrCount, nbrRawCount;
    DevBuf<uint16_t> nbrCum;      // [NBINS][nPad] cumulative entries at every distance-bin boundary
    int nbrCap = 0;               // entries per bead allocated
    bool listValid = false;
    int64_t lastBuildLoop = -1;
    double binFrac[NBINS - 1] = {0.3};   // the edge between the two segments of a row, as a fraction of deltaR beyond rmax (DDCB200_NEAR)
    float listBuildMs = 0.f;         // device time of the last list build (filter + exact pass), for ddcb200_listBuildInfo
    cudaEvent_t evList[2] = {nullptr, nullptr};
    // displacement-triggered rebuild (updateRate == 0, nbrcheck.cuh)
    DevBuf<double> chk, chkPartial;   // chk: 3 sums now, 3 sums at the build, bits of max d^2
    double *chkHost = nullptr;        // pinned

    // accumulators
    DevBuf<double> pairPartial, bondPartial, kinPartial;   // per-CTA partial sums
    DevBuf<int> colMap;           // partial column -> acc slot
    double *acc = nullptr;        // device ACC_N
    double *accHost = nullptr;    // pinned
    bool energyValid = false, kineticValid = false;
    int64_t nPairsListed = 0, totalEntries = 0;
    DevBuf<double> stage;         // H2D / D2H staging in caller order
    DevBuf<int> stageI;

    int64_t loop = 0;
    double time = 0.0;
    bool forcesValid = false;
    bool pendingKick2 = false;
    double pendingDt = 0.0;

    // profiling
    bool prof = false;
    double profMs[PROF_N] = {0};
    int64_t profLaunch[PROF_N] = {0};
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    struct PendingEv { cudaEvent_t a, b; int slot; };
    std::vector<PendingEv> pending;
    std::vector<cudaEvent_t> evPool;

    cudaEvent_t timer[4] = {nullptr, nullptr, nullptr, nullptr};
    int64_t kernelLaunches = 0;

    // multi-GPU (ddc.cuh)
    int rank = 0, nranks = 1;
    int lat[3] = {1, 1, 1};
    void *nccl = nullptr;             // ncclComm_t
    std::vector<int64_t> hMolOffset;  // host copies of the molecule table (ownership beads)
    std::vector<int> hMolBeads;
    DevBuf<int> ownerBead;            // static: bead index of the ownership bead of each bead's molecule
    bool ownerBeadValid = false;
    DevBuf<int> ddcDest;              // re-domain: destination rank of every slot (-1: ghost)
    DevBuf<uint32_t> ddcMask;         // re-domain: bit p = this local bead is a ghost of rank p
    DevBuf<int> ddcList, sendSlot, recvSlot;   // ddcList = bead ids [send lists by peer | recv lists by peer]
    DevBuf<double> sendBuf, recvBuf;
    void *ddcWork = nullptr;                // device DdcWork
    void *ddcWorkInit = nullptr;            // pinned DdcWork: the cleared pattern
    int *ddcRow = nullptr, *ddcRowAll = nullptr;   // device: this rank's count row, and every rank's (all-gather)
    int *ddcRowHost = nullptr;              // pinned copy of ddcRowAll
    double *ddcBox6 = nullptr, *ddcBoxAll = nullptr;   // device: my bounding box, every rank's
    void *boxes = nullptr;                  // device DdcBoxes
    int *ddcCounters = nullptr;             // device, 8 ints
    int *ddcHost = nullptr;                 // pinned, 64 ints
    std::vector<int> hSendCount, hRecvCount, hSendOff, hRecvOff;
    int nSendTot = 0, nRecvTot = 0;
    bool haloDirty = false, localsDirty = false;
    cudaStream_t streamH = nullptr;         // the halo runs here, beside the pair work of the rows that read no ghost
    cudaStream_t streamB = nullptr;         // the pair rows that wait for the halo: beside the tail of the other rows' launch
    cudaEvent_t evPos = nullptr, evHalo = nullptr, evBoundary = nullptr, evBonded = nullptr;
    DevBuf<int> tileGhost, tileOrder;
    int nTilesInterior = 0;
    bool haloOverlap = true;                // DDCB200_HALO=overlap|inline
    DevBuf<double> accG;              // all-reduced accumulators

    // NGLFCONSTRAINT (nglfcons.cuh): groups, per-bead LCG64 streams, constraint clusters, barostat
    int nGroups = 1;
    int groupType[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    double groupKBT[8] = {0}, groupTau[8] = {0}, groupVcm[8][3] = {{0}};
    bool anyLangevin = false;
    DevBuf<unsigned char> groupOfBead;   // by input index; null = every bead in group 0
    DevBuf<uint64_t> rngState;           // by input index
    DevBuf<uint2> rngMP;                 // {multID, prime}
    bool haveRandom = false;
    int nCons = 0;
    DevBuf<int> consAtomOff, consAtomBead, consPairOff, consPairA, consPairB;
    DevBuf<double> consPairDist;
    int *consFlag = nullptr;             // device: clusters that hit the iteration cap
    double ncKBT = 0.0, ncP0 = 0.0, ncBeta = 0.0, ncTau = 0.0;   // nglfconstraint_parms: kB*T, P0, beta, tauBarostat
    double hBuild[3] = {0, 0, 0};        // box edges when the reference positions of the displacement bounds were taken (the last build or prune)
    double hCheck[3] = {0, 0, 0};        // box edges at the last list build (nbr->h0, src/neighbor.c:231): neighborCheck
    DevBuf<double> posCheck[3];          // updateRate = 0: the positions at the last list build (neighborCheck; posBuild moves on with the prunes)
};

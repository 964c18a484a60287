/* ddcmd_b200.h - C-ABI of the B200-native Martini MD step (drop-in for ddcMD's GPU seam).
 *
 * Every entry point is extern "C", takes plain pointers and sizes, and returns an int
 * status (0 = ok, <0 = error; text via ddcb200_lastError()).  No CPU fallback exists:
 * every compute entry point fails with DDCB200_ERR_NODEVICE when no sm_100 device is
 * usable.  All floating point is fp64 in ddcMD internal units (bohr, fs, Rydberg-ish
 * energy, src/ddcMD.c:71); host arrays are in the caller's bead
 * order ("input order", the order of STATE arrays, src/state.h:7-27).
 *
 * Each function cites the reference interface it replaces.  INTEGRATION.md shows the
 * few lines a ddcMD maintainer adds to martini_parms()/nglf_parms() to bind them.
 */
#ifndef DDCMD_B200_H
#define DDCMD_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DDCB200_OK 0
#define DDCB200_ERR_ARG (-1)
#define DDCB200_ERR_NODEVICE (-2)
#define DDCB200_ERR_CUDA (-3)
#define DDCB200_ERR_STATE (-4)
#define DDCB200_ERR_CAPACITY (-5)
#define DDCB200_ERR_NCCL (-6)

typedef struct ddcb200_ctx ddcb200_ctx;

/* Box, neighbor and non-bonded constants.
 * Replaces: BOX{h,pbc} (src/box.c:50-87), NEIGHBOR{deltaR,minBoxSide} (src/neighbor.c:50-54),
 * DDC{updateRate} (src/ddc.c:61-117), and the CHARMMPOT_PARMS scalars filled by
 * martini_parms (src/bioMartini.c:1210-1245): rmax, krf, crf, ke/epsilon_r. */
typedef struct ddcb200_params
{
    double h[9];          /* box matrix, row major xx xy xz yx .. zz; must be orthorhombic */
    int pbc;              /* boundary bits, only 7 (xyz periodic) is supported */
    int updateRate;       /* DDC updateRate: rebuild cells+lists when loop % updateRate == 0; 0 = when displacements demand it */
    double rmax;          /* potential cutoff (POTENTIAL cutoff, 11 Angstrom) */
    double deltaR;        /* NEIGHBOR deltaR (skin) */
    double minBoxSide;    /* NEIGHBOR minBoxSide */
    double keR;           /* ke / epsilon_r */
    double krf, crf;      /* reaction-field constants */
    double center[3];     /* domain centre used by GeomBox (src/geom.c:335-340); 0 for one domain */
    int nConstraints;     /* SYSTEM nConstraints (temperature denominator, src/energyInfo.c:112) */
    int device;           /* CUDA device ordinal */
} ddcb200_params;

/* ETYPE subset (src/energyInfo.h:18-40) after eval_energyInfo (src/energyInfo.c:75-148),
 * plus bioEnergies-style term sums and the molecular-pressure tensor
 * (src/molecularPressure.c:22-67, src/printinfo.c:233-240). Symmetric tensors: xx yy zz xy xz yz. */
typedef struct ddcb200_etype
{
    double eion;            /* potential energy */
    double rk;              /* kinetic energy */
    double virial[6];
    double tion[6];         /* sum m v_a v_b */
    double sion[6];         /* -(virial+tion)/V */
    double pion;            /* -tr(sion)/3 */
    double temperature;     /* 2 rk / (3 N - nConstraints), internal units */
    double number;          /* beads */
    double volume;
    double eLJ, eEle, eBond, eAngle, eTorsion, eImproper, eRestraint;
    double molVirial[3];    /* virial diag after molecular correction */
    double molPressure[3];  /* (molVirial + Nmol kB T)/V, diag */
    double pMolecular;      /* trace/3 of the above */
    int64_t loop;
    double time;
    int64_t nMolecules;
    int64_t nPairsListed;   /* half-list pair count, == nbr->npairs of the reference */
} ddcb200_etype;

const char *ddcb200_lastError(void);
int ddcb200_deviceCount(void);

/* Replaces accelerator_init + allocSendGPUState/allocGPUBoxInfo (src/accelerator.c:21,
 * src/system.c:183-184). */
int ddcb200_create(const ddcb200_params *p, ddcb200_ctx **out);
void ddcb200_destroy(ddcb200_ctx *ctx);
int ddcb200_sync(ddcb200_ctx *ctx);

/* Replaces martiniNonBondGPUParms (src/bioMartiniGPU.h:8; table built by martiniLJ_parms,
 * src/bioMartini.c:868-950).  eps/sigma/shift are ntypes*ntypes, symmetric. */
int ddcb200_martiniNonBondParms(ddcb200_ctx *ctx, int ntypes, const double *eps, const double *sigma, const double *shift);

/* Per-species constants (SPECIES objects + getCGLJindexbySpecie, src/bioMartini.c:952-987):
 * LJ type, charge, mass, and the molecule type of each species. */
int ddcb200_setSpecies(ddcb200_ctx *ctx, int nspecies, const int *ljType, const double *charge, const double *mass);

/* Static per-bead identity, in input order: gid label (src/bioGid.h:13-23) and species
 * index.  nGlobal beads in total (all ranks hold the full static table). */
int ddcb200_setBeads(ddcb200_ctx *ctx, int64_t nGlobal, const uint64_t *gid, const int *species);

/* Exclusion ("bpair") tables consumed by the list build, replacing reOrgPairs
 * (src/bioMartini.c:1392-1485): for every bead the index of its molecule type, and per
 * molecule type either "single species" (all intra-molecule pairs pruned) or a list of
 * (atomI, atomJ) keys built by genMartiniBondPair (src/bioMartini.c:135-282).
 * bpairOffset has nMolTypes+1 entries into bpairI/bpairJ. */
int ddcb200_setExclusions(ddcb200_ctx *ctx, int nMolTypes, const int *molTypeOfSpecies, const int *molTypeNSpecies,
                          const int *bpairOffset, const int *bpairI, const int *bpairJ);

/* Replaces martiniBondGPUParms (src/bioMartiniGPU.h:9): flattened bonded terms, bead
 * indices in input order.  kind: 0 bond kb(b-b0)^2 (resBondSorted), 1 harmonic angle,
 * 2 cosine angle, 3 restricted-bending angle, 4 proper torsion, 5 improper
 * (src/bioCharmmCovalentEnergiesSorted.c).  idx is 4 ints per term (unused = -1);
 * parm is 3 doubles per term: (k, x0, n). */
int ddcb200_martiniBondParms(ddcb200_ctx *ctx, int64_t nTerms, const int *kind, const int *idx, const double *parm);

/* Position restraints (src/restraint.c:259-361): bead index, fractional reference point
 * (x0,y0,z0 of RESTRAINTPARMS, scaled by the box edge at evaluation time; origin==0 shifts by
 * -L/2 as the reference does), kb, and per-axis factors fc (3 per restraint). */
int ddcb200_setRestraints(ddcb200_ctx *ctx, int64_t n, const int *bead, const double *frac0, const double *kb, const double *fc,
                          int origin);

/* Molecule membership for the molecular virial (moleculeScanState, src/molecule.c:118-211):
 * molOffset has nMol+1 entries into molBeads (bead indices, input order); only
 * multi-bead molecules need to be listed.  nMolTotal counts every molecule. */
int ddcb200_setMolecules(ddcb200_ctx *ctx, int64_t nMol, const int64_t *molOffset, const int *molBeads, int64_t nMolTotal);

/* Replaces sendGPUState + sendForceVelocityToGPU (src/gpuMemUtils.h:19-29): positions and
 * velocities of the nLocal beads this context owns; bead[] gives their input-order index
 * (NULL = 0..nLocal-1). */
int ddcb200_sendState(ddcb200_ctx *ctx, int64_t nLocal, const int *bead, const double *rx, const double *ry, const double *rz,
                      const double *vx, const double *vy, const double *vz, int64_t loop, double time);

/* Replaces sendPosnToHost / sendForceVelocityToHost.  Any pointer may be NULL.  Output is
 * ordered like the bead[] array returned by ddcb200_getLocalBeads (for one context that
 * never migrates beads: the order given to sendState). */
/* The per-step upload of a host-side integrator (the reference's sendPosnToGPU / sendForceVelocityToGPU between force
 * evaluations): new positions and velocities of the beads that are already resident, same set as the last sendState.  The
 * cell order and the neighbor list are kept and rebuilt on the DDC schedule, exactly as when the step runs on the device;
 * sendState, in contrast, starts a new run (the next force evaluation rebuilds).  Falls back to sendState when no list exists. */
int ddcb200_updateState(ddcb200_ctx *ctx, int64_t nLocal, const int *bead, const double *rx, const double *ry, const double *rz,
                        const double *vx, const double *vy, const double *vz, int64_t loop, double time);
int64_t ddcb200_numLocal(ddcb200_ctx *ctx);
int ddcb200_getLocalBeads(ddcb200_ctx *ctx, int *bead);
int ddcb200_getState(ddcb200_ctx *ctx, double *rx, double *ry, double *rz, double *vx, double *vy, double *vz,
                     double *fx, double *fy, double *fz);

/* Replaces constructList(SYSTEM*, double rcut) (src/nlistGPU.h:184) = GeomBox + pairlist1 +
 * reOrgPairs of the CPU path (src/geom.c:311-383, src/pairlist.c:205-314). */
int ddcb200_constructList(ddcb200_ctx *ctx);

/* Replaces ddcenergy(ddc, sys, e_eval_flag) with eval_potential = martiniGPU1
 * (src/ddcenergy.c:160-238, src/bioMartini.cu:146): rebuilds the list when due, zeroes and
 * evaluates non-bonded + bonded (+restraint) forces at the current positions.
 * withEnergy != 0 also accumulates energies and the virial. */
int ddcb200_ddcenergy(ddcb200_ctx *ctx, int withEnergy);

/* Replaces nglfGPU / nglf (src/nglfGPU.h:16, src/nglf.c:67-112) called nsteps times:
 * velocity-Verlet with the FREE group update (src/free.c:13-28); the last step is an
 * energy step so ddcb200_energyInfo is valid afterwards. */
int ddcb200_nglf(ddcb200_ctx *ctx, int nsteps, double dt);

/* Replaces sendForceEnergyToHost + kinetic_terms + eval_energyInfo
 * (src/pairProcessGPU.cu:1556, src/energy.c:48-163, src/energyInfo.c:75-148) and
 * molecularPressure (src/molecularPressure.c:57-67). kB in internal units. */
int ddcb200_energyInfo(ddcb200_ctx *ctx, double kB, ddcb200_etype *out);

/* ---- NGLFCONSTRAINT integrator ----------------------------------------------
 * Replaces nglfconstraint_parms + nglfconstraint (src/nglfconstraint.c:86-115, :510-574) and the GPU analogue
 * nglfconstraintGPU (src/nglfconstraintGPU.cu:1255-1365).  One GPU in this version. */

/* GROUP objects (src/group.c:78-82) and the group of every bead (the "group" column of the atoms file, input order;
 * NULL = every bead in group 0).  type: 0 FREE (free_velocityUpdate, src/free.c:13-28), 1 LANGEVIN
 * (langevin_velocityUpdate, src/langevin.c:92-128) with kBT = kB*Teq, tau and vcm[3] per group, internal units. */
int ddcb200_setGroups(ddcb200_ctx *ctx, int ngroups, const int *type, const double *kBT, const double *tau, const double *vcm,
                      int64_t nGlobal, const unsigned char *groupOfBead);

/* Per-bead LCG64 streams, input order: LCG64_PARM {state, multID, prime} (src/lcg64.h:8-12), as read from the atoms file
 * or made by lcg64_default (src/lcg64.c:96-109).  getRandom returns the current states (for restart files). */
int ddcb200_setRandom(ddcb200_ctx *ctx, int64_t nGlobal, const uint64_t *state, const uint32_t *multID, const uint32_t *prime);
int ddcb200_getRandom(ddcb200_ctx *ctx, int64_t nGlobal, uint64_t *state);

/* Constraint clusters = the CONSTRAINT objects of genConstraint (src/bioMartini.c:445-565), one per residue instance
 * and CONSLISTPARMS: atomOffset/pairOffset have nCons+1 entries; atomBead = bead indices (input order) in
 * atomIDList order; pairA/pairB index into the cluster's own atom list (atomIindex/atomJindex); pairDist = r0.
 * Clusters must be disjoint; at most 32 atoms and 48 pairs each. */
int ddcb200_setConstraints(ddcb200_ctx *ctx, int64_t nCons, const int64_t *atomOffset, const int *atomBead, const int64_t *pairOffset,
                           const int *pairA, const int *pairB, const double *pairDist);

/* INTEGRATOR NGLFCONSTRAINT keys: kBT = kB*T, P0, beta (0 = no barostat), tauBarostat, internal units. */
int ddcb200_nglfconstraintParms(ddcb200_ctx *ctx, double kBT, double P0, double beta, double tauBarostat);

/* nglfconstraint called nsteps times: [barostat: molecularPressure -> changeVolume -> adjustPosn] -> group FRONT
 * velocity update -> velocity constraints -> drift + wrap -> ddcenergy -> group BACK update -> velocity constraints ->
 * kinetic_terms.  The last step is an energy step. */
int ddcb200_nglfconstraint(ddcb200_ctx *ctx, int nsteps, double dt);

/* Current box matrix (the barostat changes it; box_get_h, src/box.c:173-176). */
int ddcb200_getBox(ddcb200_ctx *ctx, double h[9]);
/* box_put(NULL, HO, &h) from the caller's side (a barostat that runs in the host code): orthorhombic h, positions sent afterwards
 * are in the new box, the neighbor list is kept (its walk bound grows by the change of the box edges since the build). */
int ddcb200_setBox(ddcb200_ctx *ctx, const double h[9]);

/* Constraint clusters that reached the 500-iteration cap so far (the reference prints a warning and goes on). */
int64_t ddcb200_constraintFailures(ddcb200_ctx *ctx);

/* Parity hooks: cell index of every local bead in the reference's GeomBox numbering
 * (src/geom.c:386-454), the grid {nx,ny,nz}, geom = {min[3], max[3], d[3]} in normalised
 * coordinates; and the half list as (beadI, beadJ, pruned) triples with gid_I < gid_J. */
int ddcb200_getCells(ddcb200_ctx *ctx, int *cellOfBead, int dims[3], double geom[9]);
int64_t ddcb200_getPairs(ddcb200_ctx *ctx, int64_t capacity, int *beadI, int *beadJ, int *pruned);

/* Timing aid for bench.py: device time (ms) spent in each kernel family since the last
 * reset, measured with CUDA events on the launching stream when profiling is enabled.
 * slots: 0 integrate, 1 pair, 2 bonded, 3 list build, 4 reductions, 5 halo. */
int ddcb200_profile(ddcb200_ctx *ctx, int enable);
int ddcb200_profileRead(ddcb200_ctx *ctx, double ms[8], int64_t launches[8], int reset);

/* CUDA-event stopwatch on the context's own stream (torch.cuda.Event would only see torch's
 * stream): record slot 0..3, elapsed ms between two recorded slots; and the number of kernels
 * this context has launched so far. */
int ddcb200_timerRecord(ddcb200_ctx *ctx, int which);
int ddcb200_timerElapsed(ddcb200_ctx *ctx, int from, int to, double *ms);
int64_t ddcb200_kernelLaunches(ddcb200_ctx *ctx);

/* Loop index of the last cell sort + list build = sys->neighbor->lastUpdate (src/ddcUpdateAll.c:135).  With DDC
 * updateRate = 0 the build is displacement-triggered: neighborRef + neighborCheck (src/neighbor.c:117-246) through
 * check4updateNeighbor / evalUpdateFlag (src/ddcUpdateAll.c:48-71). */
int64_t ddcb200_lastListBuild(ddcb200_ctx *ctx);

/* The per-GROUP (bySpecies = 0) or per-SPECIES (1) copies kinetic_terms files (src/energy.c:116-143) and each class's share of the
 * thermal flux (src/energy.c:104-106; the Martini path keeps no per-particle energy or stress, so J = sum K v), from the local beads'
 * current velocities.  out12[class * 12 + k]: k = 0 rk, 1 mass, 2 number, 3-8 sum m v_a v_b (xx yy zz xy xz yz), 9-11 sum K v.
 * nClasses = the deck's group count (at least 1) or species count.  Sums are taken in a fixed order: reproducible. */
int ddcb200_kineticByClass(ddcb200_ctx *ctx, int bySpecies, int nClasses, double *out12);

/* Replaces the pair loops of paircorrelation_eval (src/paircorrelation.c:158-420, methods geom / grid / neighborList alike):
 * counts[bin + nBins * comboIndex(si, sj)] over the local beads' pairs with gid_i < gid_j and r < rmax, 2 per same-species pair
 * and 1 otherwise; bin = (int)((r - rmin) / delta), or (int)((log10 r - log10 rmin) / delta) with logScale; nAtoms[species] =
 * local beads per species.  np = ns (ns + 1) / 2 species pairs in the reference's comboIndex order.  rmax may be any length up
 * to half the shortest box edge: the walk over the cells of the last build is widened by the displacement since then. */
int ddcb200_pairCorrelation(ddcb200_ctx *ctx, int nBins, double rmin, double delta, int logScale, double rmax,
                            unsigned long long *counts, unsigned long long *nAtoms);

/* The list build (measurement hook): *variant = 1 (there is one build: fp32 candidate pass, then the exact pairlist1 test over the
 * candidates in one sweep); ms[0] = device time of the last build's two passes, ms[1] = 0. */
int ddcb200_listBuildInfo(ddcb200_ctx *ctx, int *variant, double ms[2]);

/* Pruned rows of the pair walk (measurement hook, no reference counterpart; DDCB200_PRUNE=<every>[,<margin>]): every <every> force
 * evaluations the pair kernel also writes, per bead, the entries now closer than rmax + margin x deltaR; in between a bead walks that
 * shorter row while its displacement bounds since the prune stay within the margin (bitwise the results of the full walk).
 * info[0] = <every> (0: off), [1] = evaluations since the rows were written (-1: none valid), [2] = entries the next evaluation
 * would walk at the current positions, [3] = local beads that would walk their pruned row, [4] = entries of all pruned rows,
 * [5] = entries of all full rows. */
int ddcb200_pruneInfo(ddcb200_ctx *ctx, int64_t info[6]);

/* ---- multi-GPU: ddc-style spatial decomposition over the GPUs of one box, one process per GPU ----
 * Replaces ddc_init + ddcAssignment + ddcSendRecvTables + ddcUpdate (src/ddc.c:61-117,
 * src/ddcAssignment.c:64-107, src/ddcSendRecv.c:41-277, src/ddcUpdate.c:40-88) and the reduction of
 * eval_energyInfo (src/energyInfo.c:9-63).  Domains are the bricks of the DDC lattice lx*ly*lz (Voronoi
 * cells of a regular lattice of centres); a molecule follows its ownership bead (ddcRuleMolecule).
 * Call order on every rank: create, [setters], ddcInit, sendState (any partition of the beads among the
 * ranks), then ddcenergy / nglf / energyInfo collectively.  Rank 0 makes the id with ddcb200_ncclUniqueId
 * and the caller distributes the 128 bytes (MPI_Bcast in ddcMD, torch.distributed in bench.py).
 * After a re-domain step the local beads of a rank change: query numLocal / getLocalBeads / getState. */
int ddcb200_ncclUniqueId(unsigned char id[128]);
int ddcb200_ddcInit(ddcb200_ctx *ctx, int rank, int nranks, int lx, int ly, int lz, const unsigned char id[128]);

/* Parity hook for decks too large to copy the pair list out: order-independent hash of the pairs this rank owns (smaller gid
 * local, src/pairlist.c:244,279).  out = {count, sum, xor} of the interacting list, then of the pruned list, of
 * mix64(mix64(gid_lo) + 0x9e3779b97f4a7c15 * gid_hi) (mod 2^64); oracle/ref_dump.c computes the same from ifirst[0] / ifirst[1]. */
int ddcb200_pairSetHash(ddcb200_ctx *ctx, uint64_t out[6]);

/* Host restatement of the domain classification (same predicates as the kernels), for tests and tools:
 * owner[b] = rank owning bead b; mask[b] for `rank`: bit 31 = mine, and then bit p (p < 16) = a ghost on rank p;
 * bit 30 = owned elsewhere and a ghost here, and then bit p = its owner.  ownerBead[b] = ownership bead of b's molecule (NULL = b). */
int ddcb200_ddcPlan(const double h[9], int lx, int ly, int lz, double rlist, int64_t nGlobal, const double *rx, const double *ry,
                    const double *rz, const int *ownerBead, int rank, int *owner, uint32_t *mask);

#ifdef __cplusplus
}
#endif
#endif

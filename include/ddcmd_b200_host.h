/* ddcmd_b200_host.h - host side of the drop-in: ddcMD's object-database front end for the
 * Martini MD step, in plain C, above the CUDA C-ABI of ddcmd_b200.h.
 *
 * It reads the same decks ddcMD reads (object.data + restart + martini.data + restraint.data +
 * atoms#NNNNNN, reference src/objectSetup.c:35-41, src/collection_read.c:86-170) and mirrors
 * the reference's init chain: simulate_init -> system_init -> {species, molecule, box,
 * collection, potential(MARTINI|RESTRAINT), neighbor} -> integrator(NGLF|NGLFCONSTRAINT) -> ddc
 * (SURVEY.md section 3.1).  The result is a flat, POD description (ddcb200_deck) that
 * ddcb200_simulateBind() pushes through the C-ABI setters.
 */
#ifndef DDCMD_B200_HOST_H
#define DDCMD_B200_HOST_H
#include "ddcmd_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ddcb200_deck
{
    /* SIMULATE (src/simulate.c:151-169) */
    double dt, time;
    int64_t loop, maxloop;
    int printrate, deltaloop, snapshotrate, checkpointrate;
    /* BOX / NEIGHBOR / DDC / POTENTIAL scalars, ready for ddcb200_create */
    ddcb200_params params;
    int ddc_lx, ddc_ly, ddc_lz;
    double rcoulomb, epsilon_r, epsilon_rf, rmax4all;
    int excludePotentialTerm, potentialShift;
    /* unit system (src/units.c:450-486) */
    double kB, ke;
    double lengthPerAngstrom, energyPerKJmol, massPerAmu, pressurePerBar, timePerFs;
    /* PRINTINFO */
    int printMolecularPressure;
    /* species (creation order = SPECIES index, src/molecule.c:212-241) */
    int nspecies;
    char **speciesName;
    int *specLJ;        /* LJ type (getCGLJindexbySpecie, src/bioMartini.c:952-987) */
    double *specCharge, *specMass;
    int *specMolType;   /* molecule type index */
    int *specResidue;   /* index into the MMFF resiParms list */
    int *specAtom;      /* atom offset inside the residue */
    /* LJ table, ntypes x ntypes (martiniLJ_parms, src/bioMartini.c:868-950) */
    int ntypes;
    double *ljEps, *ljSigma, *ljShift;
    /* molecule types + bpair exclusion keys (genMartiniBondPair, src/bioMartini.c:135-282) */
    int nMolTypes;
    int *molTypeNSpecies, *molTypeResidue, *molTypeOwnerOffset;
    int *bpairOffset, *bpairI, *bpairJ;
    /* beads, input (file) order */
    int64_t n;
    uint64_t *gid;
    int *species;
    double *rx, *ry, *rz, *vx, *vy, *vz;
    /* flattened bonded terms (see ddcb200_martiniBondParms) */
    int64_t nTerms;
    int *termKind, *termIdx;
    double *termParm;
    /* restraints */
    int64_t nRestraints;
    int *restrBead;
    double *restrFrac0, *restrKb, *restrFc;
    int restrOrigin;
    /* multi-bead molecules for the molecular virial; first bead of each = ownership bead */
    int64_t nMol, nMolTotal;
    int64_t *molOffset;
    int *molBeads;
    /* INTEGRATOR (src/integrator.c:59-83): 0 = NGLF, 1 = NGLFCONSTRAINT with its keys T, P0, beta, tauBarostat,
     * isotropic (src/nglfconstraint.c:86-95), internal units */
    int integratorType;
    double ncT, ncP0, ncBeta, ncTauBarostat;
    int ncIsotropic;
    /* GROUP objects of SYSTEM groups (src/group.c:78-82): 0 = FREE, 1 = LANGEVIN {Teq, tau, vcm} (src/langevin.c:63-91,
     * 130-170); groupOfBead = index of the group named in each atoms record */
    int nGroups;
    char **groupName;
    int *groupType;
    double *groupTeq, *groupTau, *groupVcm;
    unsigned char *groupOfBead;
    /* RANDOM LCG64 (src/random.c:47-72): per-bead LCG64_PARM read from the atoms records or made by lcg64_default
     * (src/lcg64.c:96-109; src/collection.c:96-110) */
    int haveRandom;
    uint64_t randomSeed;
    uint64_t *rngState;
    uint32_t *rngMult, *rngPrime;
    /* constraint clusters (genConstraint, src/bioMartini.c:445-565), see ddcb200_setConstraints */
    int64_t nCons;
    int64_t *consAtomOffset, *consPairOffset;
    int *consAtomBead, *consPairA, *consPairB;
    double *consPairDist;
} ddcb200_deck;

/* object_compilefile(object.data) + object_compilefile(restart) + the init chain.
 * restartFile may be NULL (then "restart" next to objectFile is used when present).
 * simulateName NULL = "simulate".  Errors: returns <0, text in ddcb200_lastHostError(). */
int ddcb200_deckLoad(const char *objectFile, const char *restartFile, const char *simulateName, ddcb200_deck **out);
void ddcb200_deckFree(ddcb200_deck *deck);
const char *ddcb200_lastHostError(void);

/* simulate_init tail + firstEnergyCall prerequisites: create the device context and push
 * parameters, topology and state (src/simulate.c:104-297, src/masters.c:579-620). */
int ddcb200_simulateBind(const ddcb200_deck *deck, int device, ddcb200_ctx **out);

/* The same for rank `rank` of `nranks` processes (one per GPU): DDC lattice lx*ly*lz = nranks (the DDC
 * object's lx ly lz, src/ddc.c:72-84), ncclId = the 128 bytes of ddcb200_ncclUniqueId made on rank 0. */
int ddcb200_simulateBindRank(const ddcb200_deck *deck, int device, int rank, int nranks, int lx, int ly, int lz,
                             const unsigned char *ncclId, ddcb200_ctx **out);

/* One line of the reference's `data` file (printinfoA, src/printinfo.c:125-232): loop, time(ns),
 * Etotal, Ekin, Epot (kJ/mol per bead), T (K), P (bar; molecular if printMolecularPressure),
 * volume per bead (Ang^3), lx ly lz (Ang).  Returns the number of characters written. */
int ddcb200_printinfoLine(const ddcb200_deck *deck, const ddcb200_etype *e, char *buf, size_t len);

/* unit conversion exposed for tests: units_convert(value, from, to), NULL = internal. */
double ddcb200_unitsConvert(double value, const char *from, const char *to);

#ifdef __cplusplus
}
#endif
#endif

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

/* ANALYSIS type = subsetWrite (src/subsetWrite.c:62-168): which beads go into the periodic subset files */
typedef struct ddcb200_subset
{
    char *name, *filename, *lengthUnit, *parmsInfo;
    int evalRate, outputRate;            /* ANALYSIS eval_rate / outputrate (src/analysis.c:152-153) */
    int modulus, odd;
    uint64_t idMin, idMax;
    int64_t nIdList;
    uint64_t *idList;                    /* sorted */
    double lo[3], hi[3], vlo[3], vhi[3]; /* internal units */
    int *includeSpecies;                 /* per SPECIES index */
} ddcb200_subset;

/* ANALYSIS type = PAIRCORRELATION (src/paircorrelation.c:71-145): g(r) per species pair, linear or logarithmic bins */
typedef struct ddcb200_paircorr
{
    char *name, *filename, *miscInfo;
    int evalRate, outputRate, nBins, logScale;
    double rmin, deltaR, logDelta, rmax;     /* internal units */
} ddcb200_paircorr;

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
    /* what the writers need (writeRestart src/io.c:58-113, printinfoA src/printinfo.c:125-232): object names, the
     * directory the deck lives in (the reference's cwd), SIMULATE snapshotRootDir / gidFormat / nLoopDigits / run_id
     * (src/simulate.c:159-183,209), the SPECIES type strings, and the PRINTINFO units (src/printinfo.c:35-36,60-77)
     * in the order LENGTH TIME TEMPERATURE ENERGY PRESSURE VOLUME with their factors from internal units */
    char *runDir, *simulateName, *boxName, *collectionName, *atomsdir;
    int nLoopDigits, gidFormatHex;
    unsigned runId;
    char **speciesType;
    char *printUnit[6];
    double printConvert[6];
    double reducedCorner[3];
    /* SIMULATE checkpointmode = ASCII | BINARY, checkpointprecision = FULL | BRIEF (src/simulate.c:166-196): writeRestart
     * writes FIXRECORDBINARY records (collection_writeBLOCK_binary, src/collection_write.c:188-336) when checkpointBinary */
    int checkpointBinary, checkpointBrief;
    /* SIMULATE analysis = ... : the ANALYSIS objects of type subsetWrite with format = binaryCharmm (the positions feed of the
     * MuMMI workflow) and of type PAIRCORRELATION; any other ANALYSIS type or format is an error at load time */
    int nSubsets;
    ddcb200_subset *subsets;
    int nPairCorr;
    ddcb200_paircorr *pairCorr;
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

/* The header line printinfoA writes once at the top of `data` (src/printinfo.c:153-185). */
int ddcb200_printinfoHeader(const ddcb200_deck *deck, char *buf, size_t len);

/* One ddcMD-format snapshot: writeRestart (src/io.c:58-113) = CreateSnapshotdir (src/io.c:115-143) +
 * collection_writeBLOCK (src/collection_write.c:57-186: FIXRECORDASCII records with a CRC32 per record, pio FILEHEADER
 * of write_fileheader src/io.c:352-407; or collection_writeBLOCK_binary, :188-336, when the deck says
 * checkpointmode=BINARY) into <snapshotdir>/atoms#000000 + the `restart` object file (SIMULATE loop/time,
 * BOX h, LANGEVIN groups' Teq, COLLECTION size/files), and the ./restart link when restartLink != 0.
 * dirname NULL = "snapshot.<loop>" under SIMULATE snapshotRootDir; paths are relative to the deck's directory.
 * State arrays are in the deck's bead order, internal units; rngState NULL = the deck's LCG64 states.
 * snapshotdirOut (may be NULL) receives the directory written.  Returns 0, or <0 with ddcb200_lastHostError(). */
int ddcb200_writeRestart(const ddcb200_deck *deck, const char *dirname, int64_t loop, double time, const double h[9],
                         const double *rx, const double *ry, const double *rz, const double *vx, const double *vy,
                         const double *vz, const uint64_t *rngState, int restartLink, char *snapshotdirOut, size_t len);

/* writeBXYZ (src/io.c:144-155, collection_writeBXYZ mode 1 src/collection_write.c:338-465): <snapshotdir>/bxyz#000000, single-precision
 * positions and velocities with CRC32 per record; ddcb200_simulateMaster writes it every SIMULATE snapshotrate loops. */
int ddcb200_writeBXYZ(const ddcb200_deck *deck, const char *dirname, int64_t loop, double time, const double h[9],
                      const double *rx, const double *ry, const double *rz, const double *vx, const double *vy, const double *vz);

/* subsetWriteBinaryCharmm (src/subsetWrite.c:409-522): <snapshotdir>/<filename>#000000 with one 24-byte record
 * {id u8, pinfo u4, rx ry rz f4 relative to the box corner, in lengthUnit} per bead that passes the subset's filters
 * (rejectParticle, :532-564).  snapshotdir as in ddcb200_writeRestart (dirname NULL = snapshot.<loop>).  Returns the number
 * of records written, or <0. */
int64_t ddcb200_subsetWrite(const ddcb200_deck *deck, int which, const char *dirname, int64_t loop, double time, const double h[9],
                            const double *rx, const double *ry, const double *rz, const double *vx, const double *vy, const double *vz);

/* paircorrelation_output (src/paircorrelation.c:448-513): <snapshotdir>/<filename> with one row per bin (bin centre in Ang, then
 * g for every species pair) from the accumulated g[nBins * np] (sum over the samples of counts / (N_i N_j)) and the sample count. */
int ddcb200_pairCorrelationWrite(const ddcb200_deck *deck, int which, const char *dirname, int64_t loop, double volume, const double *g, int nsample);

/* readCMDS (src/readCmds.c:20-57): commands left in <runDir>/ddcMD_CMDS ("checkpoint", "kill", "exit", "profile", "hpm",
 * "analysis"), one per line; the file is truncated after reading.  Returns the OR of the DDCB200_CMD_* flags. */
#define DDCB200_CMD_CHECKPOINT 1
#define DDCB200_CMD_STOP 2
#define DDCB200_CMD_DUMP_PROFILE 4
#define DDCB200_CMD_HPM_PRINT 8
#define DDCB200_CMD_DO_ANALYSIS 16
#define DDCB200_CMD_NEW_OBJECT 32
int ddcb200_readCMDS(const char *filename);

/* simulateMaster (src/masters.c:383-559) for a Martini deck on one GPU: simulate_init, firstEnergyCall, then the MD loop
 * with the reference's cadence - eval_integrator up to the next printrate / snapshotrate / checkpointrate loop
 * (findEndLoop, src/masters.c:263-281), a `data` line (and stdout line) every printrate loops, ddcMD_CMDS polled on print
 * loops, writeRestart every checkpointrate loops or on a "checkpoint"/"exit" command, the subsetWrite analyses every
 * outputrate loops (doAnalysis, src/masters.c:295-302), a final data line when maxloop is
 * not a print loop.  Files are written into the deck's directory.  Returns 0, or <0 with ddcb200_lastHostError(). */
int ddcb200_simulateMaster(const char *objectFile, const char *restartFile, const char *simulateName, int device);

/* unit conversion exposed for tests: units_convert(value, from, to), NULL = internal. */
double ddcb200_unitsConvert(double value, const char *from, const char *to);

#ifdef __cplusplus
}
#endif
#endif

/* ref_dump: drives the UNMODIFIED reference CPU path (LLNL/ddcMD) and dumps
 * per-bead state so the C restatement in oracle/martini_oracle.c and the CUDA
 * path can be pinned against the real reference.
 *
 * TEST INFRASTRUCTURE ONLY - never linked into the product library.
 *
 * Linked against the reference's own objects (built from /root/reference/src by
 * oracle/build_ref.sh with -DTESTEXE=1 on ddcMD.c, which drops main(),
 * src/ddcMD.c:62).  Mirrors main() (src/ddcMD.c:66-88) and the head of
 * simulateMaster() (src/masters.c:369-404): simulate_init -> adjustBox ->
 * firstEnergyCall, then calls eval_integrator nsteps times.
 *
 * Usage (cwd = deck directory with object.data + restart):
 *     ref_dump <out.bin> [nsteps] [full_dump_every] [light|hash]
 * "light" (4th argument) skips the per-bead and pair-list records: timing runs.  "hash" keeps the per-bead records and the
 * cell ids but replaces the two pair lists by order-independent hashes of their (gid, gid) pairs: million-bead decks.
 *
 * Output: sequence of records  name[32] | dtype char ('d','q','i') | pad[7] |
 * count u64 | payload.  Read by tests/refdump.py.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <mpi.h>

#include "commandLineOptions.h"
#include "masters.h"
#include "routineManager.h"
#include "simulate.h"
#include "system.h"
#include "state.h"
#include "neighbor.h"
#include "geom.h"
#include "units.h"
#include "codata.h"
#include "primes.h"
void objectSetup(void *parms, MPI_Comm comm);
#include "utilities.h"
#include "bioCharmm.h"
#include "bioCharmmParms.h"
#include "energyInfo.h"
#include "box.h"
#include "ddcenergy.h"
#include "preduce.h"
#include "random.h"
#include "lcg64.h"

void mpiStartUp(int argc, char *argv[]);
void commons_init(void);
void version_init(int argc, char *argv[]);
void adjustBox(SIMULATE *simulate);
void firstEnergyCall(SIMULATE *simulate);
void kinetic_terms(SYSTEM *sys, int flag);
PARTICLESET *getParticleSet(void);
double units_convert(double value, const char *from, const char *to);

static FILE *out;
static int nsteps = 0;
static int dump_every = 0;
static int light = 0;
static int hashPairs = 0;

/* order-independent hash of a pair set: sum and xor (mod 2^64) of a 64-bit mix of (smaller gid, larger gid).  The CUDA side
 * computes the same numbers on the device (ddcb200_pairSetHash) */
static uint64_t mix64(uint64_t x)
{
    x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull;
    x ^= x >> 27; x *= 0x94d049bb133111ebull;
    x ^= x >> 31;
    return x;
}
static uint64_t pairHash(uint64_t a, uint64_t b)
{
    const uint64_t lo = a < b ? a : b, hi = a < b ? b : a;
    return mix64(mix64(lo) + 0x9e3779b97f4a7c15ull * hi);
}

static void rec(const char *name, char dtype, const void *data, uint64_t count)
{
    char hdr[40];
    memset(hdr, 0, sizeof hdr);
    strncpy(hdr, name, 31);
    hdr[32] = dtype;
    fwrite(hdr, 1, 40, out);
    fwrite(&count, 8, 1, out);
    size_t sz = (dtype == 'i') ? 4 : 8;
    if (count) fwrite(data, sz, count, out);
}
static void rec_d(const char *name, double v) { rec(name, 'd', &v, 1); }
static void rec_i(const char *name, int v) { rec(name, 'i', &v, 1); }

static void dump_energy(const char *prefix, SYSTEM *sys)
{
    char nm[32];
    ETYPE *e = &sys->energyInfo;
    double v[16] = {e->eion, e->rk, e->pion, e->temperature, e->number, e->mass,
                    e->virial.xx, e->virial.yy, e->virial.zz, e->virial.xy, e->virial.xz, e->virial.yz,
                    sys->energy, 0, 0, 0};
    snprintf(nm, sizeof nm, "%senergy", prefix);
    rec(nm, 'd', v, 13);
    double s[12] = {e->sion.xx, e->sion.yy, e->sion.zz, e->sion.xy, e->sion.xz, e->sion.yz,
                    e->tion.xx, e->tion.yy, e->tion.zz, e->tion.xy, e->tion.xz, e->tion.yz};
    snprintf(nm, sizeof nm, "%sstress", prefix);
    rec(nm, 'd', s, 12);
}

static void dump_state(const char *prefix, SYSTEM *sys)
{
    char nm[32];
    STATE *st = sys->collection->state;
    unsigned n = sys->nion;
#define R(field, ty, ptr) snprintf(nm, sizeof nm, "%s" field, prefix); rec(nm, ty, ptr, n)
    R("label", 'q', st->label);
    R("rx", 'd', st->rx); R("ry", 'd', st->ry); R("rz", 'd', st->rz);
    R("vx", 'd', st->vx); R("vy", 'd', st->vy); R("vz", 'd', st->vz);
    R("fx", 'd', st->fx); R("fy", 'd', st->fy); R("fz", 'd', st->fz);
    R("q", 'd', st->q);
#undef R
    int *sp = malloc(sizeof(int) * (n + 1));
    for (unsigned i = 0; i < n; i++) sp[i] = st->species[i]->index;
    snprintf(nm, sizeof nm, "%sspecies", prefix);
    rec(nm, 'i', sp, n);
    free(sp);
}

/* per-bead LCG64 state (LANGEVIN groups; src/lcg64.h:8-12), when the SYSTEM has a RANDOM object */
static void dump_rng(const char *prefix, SYSTEM *sys)
{
    RANDOM *random = system_getRandom(sys);
    if (random == NULL || random->itype != LCG64) return;
    unsigned n = sys->nlocal;
    uint64_t *st = malloc(8 * (n + 1));
    int *mp = malloc(sizeof(int) * 2 * (n + 1));
    for (unsigned i = 0; i < n; i++)
    {
        LCG64_PARM *p = (LCG64_PARM *)random_getParms(random, i);
        st[i] = p->state;
        mp[2 * i] = (int)p->multID;
        mp[2 * i + 1] = (int)p->prime;
    }
    char nm[32];
    snprintf(nm, sizeof nm, "%srng_state", prefix);
    rec(nm, 'q', st, n);
    snprintf(nm, sizeof nm, "%srng_mp", prefix);
    rec(nm, 'i', mp, 2 * (uint64_t)n);
    free(st);
    free(mp);
}

static void dump_neighbor(SYSTEM *sys)
{
    NBR *nbr = sys->neighbor;
    GEOM *g = nbr->geom;
    unsigned n = sys->nion, nlocal = sys->nlocal;
    int *cell = malloc(sizeof(int) * (n + 1));
    for (unsigned i = 0; i < n; i++) cell[i] = (int)(g->pinfo[i].box - g->box);
    rec("cell", 'i', cell, n);
    free(cell);
    int dims[5] = {g->nx, g->ny, g->nz, g->nbox, (int)g->method};
    rec("geom_dims", 'i', dims, 5);
    double gp[14] = {g->min.x, g->min.y, g->min.z, g->max.x, g->max.y, g->max.z,
                     g->d.x, g->d.y, g->d.z, g->rcut, g->minBoxSide,
                     getParticleSet()->center->x, getParticleSet()->center->y, getParticleSet()->center->z};
    rec("geom_parms", 'd', gp, 14);
    /* pair lists: ifirst[0] = interacting, ifirst[1] = pruned (reOrgPairs). */
    if (hashPairs)
    {
        STATE *st = sys->collection->state;
        uint64_t h[6] = {0, 0, 0, 0, 0, 0};      /* per list: count, sum, xor */
        for (int l = 0; l < 2; l++)
            for (unsigned i = 0; i < nlocal; i++)
                for (PAIRS *p = nbr->particles[i].ifirst[l]; p; p = p->ilink)
                {
                    const uint64_t v = pairHash(st->label[i], st->label[p->j]);
                    h[3 * l]++;
                    h[3 * l + 1] += v;
                    h[3 * l + 2] ^= v;
                }
        rec("pairhash", 'q', h, 6);
    }
    else
    for (int l = 0; l < 2; l++)
    {
        uint64_t cnt = 0;
        for (unsigned i = 0; i < nlocal; i++)
            for (PAIRS *p = nbr->particles[i].ifirst[l]; p; p = p->ilink) cnt++;
        int *ij = malloc(sizeof(int) * (2 * cnt + 2));
        uint64_t k = 0;
        for (unsigned i = 0; i < nlocal; i++)
            for (PAIRS *p = nbr->particles[i].ifirst[l]; p; p = p->ilink)
            {
                ij[2 * k] = (int)i;
                ij[2 * k + 1] = p->j;
                k++;
            }
        rec(l == 0 ? "pairs0" : "pairs1", 'i', ij, 2 * cnt);
        free(ij);
    }
    rec_i("nSearch", (int)nbr->nSearch);
    rec_i("npairs", (int)nbr->npairs);
}

static void dumpMaster(void *parms, MPI_Comm comm)
{
    SIMULATEMASTERPARMS *smParms = (SIMULATEMASTERPARMS *)parms;
    SIMULATE *simulate = simulate_init(NULL, smParms->common.simulateName, comm);
    SYSTEM *sys = simulate->system;
    adjustBox(simulate);
    firstEnergyCall(simulate);

    rec_i("nlocal", (int)sys->nlocal);
    rec_i("nion", (int)sys->nion);
    rec_i("nspecies", sys->nspecies);
    rec_d("dt", simulate->dt);
    THREE_MATRIX h = box_get_h(sys->box);
    rec("h", 'd', &h, 9);
    rec("hinv", 'd', &sys->box->hinv, 9);
    double un[8] = {units_convert(1.0, NULL, "Angstrom"), units_convert(1.0, NULL, "kJ/mol"),
                    units_convert(1.0, NULL, "amu"), units_convert(1.0, NULL, "bar"),
                    units_convert(1.0, NULL, "K"), ke, kB, units_convert(1.0, NULL, "fs")};
    rec("units", 'd', un, 8);
    for (int i = 0; i < sys->npotential; i++)
        if (strcmp(sys->potential[i]->type, "MARTINI") == 0)
        {
            CHARMMPOT_PARMS *p = (CHARMMPOT_PARMS *)sys->potential[i]->parms;
            double mp[6] = {p->rmax, p->rcoulomb, p->epsilon_r, p->epsilon_rf, p->krf, p->crf};
            rec("martini_parms", 'd', mp, 6);
            rec("bioEnergies", 'd', &p->bioEnergies, sizeof(BIOENERGIES) / sizeof(double));
        }
    double *mass = malloc(sizeof(double) * sys->nspecies);
    for (int i = 0; i < sys->nspecies; i++) mass[i] = ((ATOMTYPE_PARMS *)(sys->species[i]->parm))->mass;
    rec("species_mass", 'd', mass, sys->nspecies);
    free(mass);
    {
        char names[1 << 16];
        names[0] = 0;
        for (int i = 0; i < sys->nspecies; i++)
        {
            strcat(names, sys->species[i]->name);
            strcat(names, " ");
        }
        size_t len = strlen(names);
        size_t nq = (len + 8) / 8;
        char *buf = calloc(nq, 8);
        memcpy(buf, names, len);
        rec("species_names", 'q', buf, nq);
        free(buf);
    }
    if (!light) dump_state("s0_", sys);
    if (!light) dump_rng("s0_", sys);
    dump_energy("s0_", sys);
    if (!light) dump_neighbor(sys);
    rec_i("npairs0", (int)sys->neighbor->npairs);

    if (nsteps > 0)
    {
        double *trace = malloc(sizeof(double) * 16 * (size_t)nsteps);
        double *wall = malloc(sizeof(double) * (size_t)nsteps);
        double *boxtrace = malloc(sizeof(double) * 3 * (size_t)nsteps);   /* barostat: box edges after every step */
        const double t0 = MPI_Wtime();
        for (int s = 0; s < nsteps; s++)
        {
            simulate->integrator->eval_integrator(simulate->ddc, simulate, simulate->integrator->parms);
            kinetic_terms(sys, 1);
            eval_energyInfo(sys);
            ETYPE *e = &sys->energyInfo;
            double *t = trace + 16 * (size_t)s;
            t[0] = (double)simulate->loop; t[1] = e->eion; t[2] = e->rk; t[3] = e->pion; t[4] = e->temperature;
            t[5] = e->virial.xx; t[6] = e->virial.yy; t[7] = e->virial.zz;
            t[8] = e->virial.xy; t[9] = e->virial.xz; t[10] = e->virial.yz;
            t[11] = e->tion.xx; t[12] = e->tion.yy; t[13] = e->tion.zz;
            t[14] = (double)sys->neighbor->npairs; t[15] = (double)sys->neighbor->lastUpdate;
            wall[s] = MPI_Wtime() - t0;
            {
                THREE_MATRIX hs = box_get_h(sys->box);
                boxtrace[3 * s] = hs.xx; boxtrace[3 * s + 1] = hs.yy; boxtrace[3 * s + 2] = hs.zz;
            }
            if (dump_every > 0 && (s + 1) % dump_every == 0 && s + 1 < nsteps)
            {
                char pre[32];
                snprintf(pre, sizeof pre, "s%d_", s + 1);
                dump_state(pre, sys);
            }
        }
        rec("trace", 'd', trace, 16 * (uint64_t)nsteps);
        rec("wall", 'd', wall, (uint64_t)nsteps);
        rec("boxtrace", 'd', boxtrace, 3 * (uint64_t)nsteps);
        free(boxtrace);
        free(trace);
        free(wall);
        if (!light) dump_state("sN_", sys);
        if (!light) dump_rng("sN_", sys);
        dump_energy("sN_", sys);
    }
    fclose(out);
}

int main(int argc, char *argv[])
{
    if (argc < 2)
    {
        fprintf(stderr, "usage: ref_dump out.bin [nsteps] [dump_every]\n");
        return 2;
    }
    out = fopen(argv[1], "wb");
    if (!out) { perror(argv[1]); return 2; }
    if (argc > 2) nsteps = atoi(argv[2]);
    if (argc > 3) dump_every = atoi(argv[3]);
    if (argc > 4)
    {
        if (strcmp(argv[4], "hash") == 0) hashPairs = 1;
        else light = 1;
    }
    char *fake_argv[2] = {argv[0], NULL};
    int fake_argc = 1;
    mpiStartUp(fake_argc, fake_argv);
    COMMAND_LINE_OPTIONS opt = parseCommandLine(fake_argc, fake_argv);
    checkLimits();
    commons_init();
    prime_init(30000, 0, 1);
    units_internal(a0_MKS, Rinfhc_MKS * 1e-30 / (a0_MKS * a0_MKS), 1e-15, e_MKS / 1e-15, Rinfhc_eV / kB_eV, 1.0, 1.0);
    units_external(1e-10, u_MKS, 1e-15, e_MKS / 1e-15, 1.0, 1.0, 1.0);
    version_init(fake_argc, fake_argv);
    MASTER master = masterFactory(opt);
    objectSetup(master.parms, MPI_COMM_WORLD);
    ROUTINE *routine = routineManager_init(NULL, "routineManager", dumpMaster, master.parms);
    routine->fcn(routine->parms, routine->comm);
    MPI_Finalize();
    return 0;
}

/* ddcmd_shim.c - the binding of INTEGRATION.md section B as real code, compiled against ddcMD's own headers.
 *
 * It plugs libddcmd_b200 into a running ddcMD at the reference's plug-in seam, without touching a reference source file:
 *   POTENTIAL  MARTINI    ->eval_potential  = martiniB200   (the signature of martini(), src/potential.c:43, src/ddcenergy.c:212)
 *   POTENTIAL  RESTRAINT  ->eval_potential  = a no-op       (the library evaluates the restraints with the bonded terms)
 *   INTEGRATOR NGLF       ->eval_integrator = nglfB200      (mode 2 only; the signature of nglf(), src/nglf.h:12, src/masters.c:445)
 * Everything else - object database, simulate_init, ddcenergy's bookkeeping, kinetic_terms, eval_energyInfo, molecular
 * pressure, printinfo - stays ddcMD's.  The library reads the same deck files for its static tables (ddcb200_deckLoad) and
 * takes the dynamic state from ddcMD's STATE arrays, so whatever ddcMD did to them before (restart, thermalize, ...) holds.
 *
 * mode 1: ddcMD's own nglf integrates on the host; every force evaluation goes through the library (state up, forces down).
 * mode 2: the whole step runs on the device; positions, velocities, forces, eion and the virial come back after each step so
 *         that ddcMD's kinetic_terms / eval_energyInfo / printinfo / writeRestart see them.
 *
 * Built by oracle/build_ref.sh next to the reference objects (it needs ddcMD's headers); never part of libddcmd_b200.so.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "simulate.h"
#include "system.h"
#include "state.h"
#include "potential.h"
#include "integrator.h"
#include "energyInfo.h"
#include "error.h"
#include "ddc.h"
#include "codata.h"
#include "box.h"
#include "bioCharmmParms.h"
#include "neighbor.h"
void charmmResidues(SYSTEM *sys, CHARMMPOT_PARMS *parms);
#include "../include/ddcmd_b200_host.h"

static ddcb200_ctx *b200;
static ddcb200_deck *b200deck;
static int *beadOfLocal;            /* ddcMD local index -> bead index of the deck (file order), by gid */
static double *fbuf[3];
static unsigned nbuf;

static void ck(int rc, const char *where, const char *text)
{
    /* the reference's convention for plug-ins: abort from inside (src/bioMartini.c:209-212) */
    if (rc) error_action((char *)text, ERROR_IN((char *)where, ABORT));
}

typedef struct { gid_type gid; int bead; } GidBead;
static int cmpGidBead(const void *a, const void *b)
{
    const gid_type x = ((const GidBead *)a)->gid, y = ((const GidBead *)b)->gid;
    return x < y ? -1 : (x > y ? 1 : 0);
}

static void mapLocals(SYSTEM *sys)
{
    STATE *s = sys->collection->state;
    const unsigned n = sys->nlocal;
    if (n != (unsigned)b200deck->n) ck(-1, "ddcmd_shim", "the shim runs ddcMD on one task: every bead must be local");
    if (n > nbuf)
    {
        beadOfLocal = (int *)realloc(beadOfLocal, sizeof(int) * n);
        for (int a = 0; a < 3; a++) fbuf[a] = (double *)realloc(fbuf[a], sizeof(double) * n);
        nbuf = n;
    }
    GidBead *tab = (GidBead *)malloc(sizeof(GidBead) * n);
    for (unsigned i = 0; i < n; i++) { tab[i].gid = b200deck->gid[i]; tab[i].bead = (int)i; }
    qsort(tab, n, sizeof(GidBead), cmpGidBead);
    for (unsigned i = 0; i < n; i++)
    {
        GidBead key = {s->label[i], 0};
        GidBead *hit = (GidBead *)bsearch(&key, tab, n, sizeof(GidBead), cmpGidBead);
        if (!hit) ck(-1, "ddcmd_shim", "a bead of ddcMD's state is not in the deck");
        beadOfLocal[i] = hit->bead;
    }
    free(tab);
}

static void sendState(SYSTEM *sys)
{
    STATE *s = sys->collection->state;
    mapLocals(sys);                          /* ddcMD may reorder its locals inside ddcenergy */
    {
        /* a host-side barostat (nglfconstraint's changeVolume) may have changed the box since the last call */
        static THREE_MATRIX last;
        THREE_MATRIX h = box_get_h(sys->box);
        if (memcmp(&h, &last, sizeof h) != 0)
        {
            ck(ddcb200_setBox(b200, (const double *)&h), "setBox", ddcb200_lastError());
            last = h;
        }
    }
    /* the first call starts the run (sendState); later calls are the per-step upload that keeps the neighbor list */
    static int started;
    if (!started)
        ck(ddcb200_sendState(b200, sys->nlocal, beadOfLocal, s->rx, s->ry, s->rz, s->vx, s->vy, s->vz, sys->loop, sys->time), "sendState",
           ddcb200_lastError());
    else
        ck(ddcb200_updateState(b200, sys->nlocal, beadOfLocal, s->rx, s->ry, s->rz, s->vx, s->vy, s->vz, sys->loop, sys->time), "updateState",
           ddcb200_lastError());
    started = 1;
}

static void addEnergies(ETYPE *e, const ddcb200_etype *o)
{
    /* potentials accumulate into ETYPE after zeroAll */
    e->eion += o->eion;
    e->virial.xx += o->virial[0]; e->virial.yy += o->virial[1]; e->virial.zz += o->virial[2];
    e->virial.xy += o->virial[3]; e->virial.xz += o->virial[4]; e->virial.yz += o->virial[5];
}

/* eval_potential of POTENTIAL MARTINI (replaces martini(), src/bioMartini.c:1487-1510) */
static void martiniB200(void *sys_, void *parms, void *e_)
{
    SYSTEM *sys = (SYSTEM *)sys_;
    STATE *s = sys->collection->state;
    ddcb200_etype o;
    /* bookkeeping of martini() that other parts of ddcMD read: nglfconstraint takes the residue table and the gid order of the
     * local beads from the potential's parms (paddingCons / velocityConstraintOld, src/nglfconstraint.c:316-391,439-457), and it is
     * charmmConvalent that refreshes them on every call (charmmResidues, src/bioCharmmCovalent.c:48-93).  Host-side, O(n log n). */
    charmmResidues(sys, (CHARMMPOT_PARMS *)parms);
    sendState(sys);
    ck(ddcb200_ddcenergy(b200, 1), "martiniB200", ddcb200_lastError());
    ck(ddcb200_energyInfo(b200, kB, &o), "martiniB200", ddcb200_lastError());
    ck(ddcb200_getState(b200, NULL, NULL, NULL, NULL, NULL, NULL, fbuf[0], fbuf[1], fbuf[2]), "martiniB200", ddcb200_lastError());
    for (unsigned i = 0; i < sys->nlocal; i++)
    {
        s->fx[i] += fbuf[0][i];
        s->fy[i] += fbuf[1][i];
        s->fz[i] += fbuf[2][i];
    }
    addEnergies((ETYPE *)e_, &o);
}

static void noPotential(void *sys, void *parms, void *e) { (void)sys; (void)parms; (void)e; }

/* eval_integrator of INTEGRATOR NGLF (replaces nglf(), src/nglf.c:67-112): one whole step on the device */
static void nglfB200(void *ddc_, void *simulate_, void *parms)
{
    SIMULATE *simulate = (SIMULATE *)simulate_;
    SYSTEM *sys = simulate->system;
    STATE *s = sys->collection->state;
    ddcb200_etype o;
    (void)ddc_; (void)parms;
    ck(ddcb200_nglf(b200, 1, simulate->dt), "nglfB200", ddcb200_lastError());
    simulate->loop++;
    simulate->time += simulate->dt;
    sys->loop = simulate->loop;
    sys->time = simulate->time;
    ck(ddcb200_energyInfo(b200, kB, &o), "nglfB200", ddcb200_lastError());
    ck(ddcb200_getState(b200, s->rx, s->ry, s->rz, s->vx, s->vy, s->vz, s->fx, s->fy, s->fz), "nglfB200", ddcb200_lastError());
    sys->energyInfo.eion = 0.0;
    sys->energyInfo.virial = szero;
    addEnergies(&sys->energyInfo, &o);       /* kinetic_terms + eval_energyInfo of the caller finish the ETYPE from these */
}

int b200_install(SIMULATE *simulate, const char *objectFile, const char *restartFile, int mode, int device)
{
    SYSTEM *sys = simulate->system;
    if (ddcb200_deckLoad(objectFile, restartFile, simulate->name, &b200deck)) ck(-1, "b200_install", ddcb200_lastHostError());
    if (ddcb200_simulateBind(b200deck, device, &b200)) ck(-1, "b200_install", ddcb200_lastHostError());
    int found = 0;
    sys->neighborTableType = 0;
    for (int i = 0; i < sys->npotential; i++)
    {
        POTENTIAL *p = sys->potential[i];
        if (strcmp(p->type, "MARTINI") == 0)
        {
            /* what martini_parms() does when an accelerator is configured (src/bioMartini.c:1337-1345): the pair list lives on the
             * device, so ddcUpdateAll calls constructList() - a stub in a CPU build; the library rebuilds its own list on the same
             * schedule inside ddcb200_ddcenergy - instead of neighbors1(), and ddcMD no longer builds its CPU pair list next to ours */
            p->eval_potential = martiniB200;
            p->use_gpu_list = 1;
            p->neighborTableType = NEIGHBORTABLE_GPU;
            found = 1;
        }
        else if (strcmp(p->type, "RESTRAINT") == 0) p->eval_potential = noPotential;
        else ck(-1, "b200_install", "only MARTINI and RESTRAINT potentials can be bound");
        sys->neighborTableType |= p->neighborTableType;      /* system_init's rule, src/system.c:202-206 */
    }
    if (!found) ck(-1, "b200_install", "no MARTINI potential in the SYSTEM");
    if (mode == 2)
    {
        if (simulate->integrator->itype != NGLF) ck(-1, "b200_install", "mode 2 binds INTEGRATOR type = NGLF");
        simulate->integrator->eval_integrator = nglfB200;
        sendState(sys);                      /* the state ddcMD holds now is where the device run starts */
    }
    return 0;
}

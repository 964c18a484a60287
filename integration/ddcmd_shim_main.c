/* ddcmd_shim_main.c - ddcMD with libddcmd_b200 plugged in at the plug-in seam (integration/ddcmd_shim.c).
 *
 *     ddcMD_shim <mode 1|2> [device]          (cwd = a deck directory: object.data + restart, as for ddcMD itself)
 *
 * main() and the master below use ddcMD's own functions throughout - start-up as src/ddcMD.c:66-88, then the head and the
 * print cadence of simulateMaster (src/masters.c:369-420, 447-485): simulate_init, adjustBox, [b200_install], firstEnergyCall,
 * printinfo, and per step eval_integrator + kinetic_terms + eval_energyInfo + printinfoAll at printrate.  The `data` file it
 * leaves is therefore ddcMD's own output of a run whose forces (mode 1) or whole steps (mode 2) came from the library.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <mpi.h>
#include "commandLineOptions.h"
#include "masters.h"
#include "routineManager.h"
#include "simulate.h"
#include "system.h"
#include "units.h"
#include "codata.h"
#include "primes.h"
#include "utilities.h"
#include "energyInfo.h"
#include "printinfo.h"
#include "object.h"

void objectSetup(void *parms, MPI_Comm comm);
void mpiStartUp(int argc, char *argv[]);
void commons_init(void);
void version_init(int argc, char *argv[]);
void adjustBox(SIMULATE *simulate);
void firstEnergyCall(SIMULATE *simulate);
void kinetic_terms(SYSTEM *sys, int flag);
int b200_install(SIMULATE *simulate, const char *objectFile, const char *restartFile, int mode, int device);

static int shimMode = 1, shimDevice = 0;

static void shimMaster(void *parms, MPI_Comm comm)
{
    SIMULATEMASTERPARMS *smParms = (SIMULATEMASTERPARMS *)parms;
    SIMULATE *simulate = simulate_init(NULL, smParms->common.simulateName, comm);
    SYSTEM *sys = simulate->system;
    adjustBox(simulate);
    b200_install(simulate, "object.data", NULL, shimMode, shimDevice);
    firstEnergyCall(simulate);
    printinfo(simulate, &sys->energyInfo);
    printinfoAll(simulate, &sys->energyInfo);
    while (simulate->loop < simulate->maxloop)
    {
        simulate->integrator->eval_integrator(simulate->ddc, simulate, simulate->integrator->parms);
        kinetic_terms(sys, 1);
        eval_energyInfo(sys);
        if (TEST0(simulate->loop, simulate->printrate)) printinfoAll(simulate, &sys->energyInfo);
    }
    printinfo_close(simulate);
}

int main(int argc, char *argv[])
{
    if (argc < 2) { fprintf(stderr, "usage: ddcMD_shim <mode 1|2> [device]\n"); return 2; }
    shimMode = atoi(argv[1]);
    if (argc > 2) shimDevice = atoi(argv[2]);
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
    ROUTINE *routine = routineManager_init(NULL, "routineManager", shimMaster, master.parms);
    routine->fcn(routine->parms, routine->comm);
    MPI_Finalize();
    return 0;
}

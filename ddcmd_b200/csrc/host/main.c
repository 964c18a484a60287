/* main.c - stand-alone driver: the reference's `ddcMD [-o object.data] [-r restart] [-s simulate]` (src/ddcMD.c:66-88,
 * option parsing src/commandLineOptions.c:17-25,68-100) for Martini decks, on one B200.
 *
 *   ddcMD_b200 [simulate] [-o objectFile] [-r restartFile] [-s simulateName] [-d device]
 *
 * Runs ddcb200_simulateMaster: data lines, ddcMD_CMDS control and ddcMD-format restarts in the deck's directory.
 * There is no CPU path: without an sm_100 device the run stops with the library's error text.
 */
#include "../../../include/ddcmd_b200_host.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

int main(int argc, char **argv)
{
    const char *objectFile = "object.data", *restartFile = NULL, *simulateName = NULL;
    int device = 0;
    int i = 1;
    if (argc > 1 && argv[1][0] != '-')
    {
        /* the master name; only simulateMaster exists here */
        if (strcmp(argv[1], "simulate") != 0)
        {
            fprintf(stderr, "ddcMD_b200: master \"%s\" is not supported (only \"simulate\")\n", argv[1]);
            return 2;
        }
        i++;
    }
    for (; i < argc; i++)
    {
        const char *opt = argv[i], *val = (i + 1 < argc && argv[i + 1][0] != '-') ? argv[++i] : "";
        if (strcmp(opt, "-o") == 0) objectFile = val;
        else if (strcmp(opt, "-r") == 0) restartFile = val;
        else if (strcmp(opt, "-s") == 0) simulateName = val;
        else if (strcmp(opt, "-d") == 0) device = atoi(val);
        else
        {
            fprintf(stderr, "ddcMD_b200: unknown option %s\n", opt);
            return 2;
        }
    }
    if (ddcb200_simulateMaster(objectFile, restartFile, simulateName, device) != 0)
    {
        fprintf(stderr, "ddcMD_b200: %s\n", ddcb200_lastHostError());
        return 1;
    }
    return 0;
}

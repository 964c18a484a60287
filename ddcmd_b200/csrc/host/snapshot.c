/* snapshot.c - ddcMD-format output and the simulateMaster loop, host side (plain C).
 *
 * Mirrors, for a Martini deck on one rank:
 *   writeRestart            src/io.c:58-113        <snapshotdir>/atoms#000000 + <snapshotdir>/restart (+ ./restart link)
 *   CreateSnapshotdir       src/io.c:115-143       "<snapshotRootDir>/snapshot.<loop>"
 *   collection_writeBLOCK   src/collection_write.c:57-186   FIXRECORDASCII records, CRC32 per record, LCG64 field
 *   write_fileheader        src/io.c:352-407       the pio FILEHEADER object
 *   box_write               src/box.c:91-108
 *   langevin_write_dynamics src/langevin.c:25-30
 *   readCMDS                src/readCmds.c:20-57
 *   simulateMaster          src/masters.c:383-559  (findEndLoop :263-281, doCheckpoint :319-326)
 *
 * Records written here are byte-identical to the reference's for the same state (tests/test_snapshot_cpu.py compares
 * with a snapshot written by the unmodified reference); header lines that name the writer (create_time, run_id,
 * code_version, srcpath) differ by construction.
 */
#include "host.h"
#include "../../../include/ddcmd_b200_host.h"
#include <errno.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

#define MAXLREC 1024   /* src/collection_write.c:27 */

typedef struct { char *p; size_t n, cap; } SBuf;
static void sbCat(SBuf *b, const char *fmt, ...) __attribute__((format(printf, 2, 3)));
#include <stdarg.h>
static void sbCat(SBuf *b, const char *fmt, ...)
{
    va_list ap;
    for (;;)
    {
        va_start(ap, fmt);
        int k = vsnprintf(b->p ? b->p + b->n : NULL, b->p ? b->cap - b->n : 0, fmt, ap);
        va_end(ap);
        if (b->p && (size_t)k < b->cap - b->n) { b->n += (size_t)k; return; }
        b->cap = 2 * (b->cap + (size_t)k) + 64;
        b->p = (char *)realloc(b->p, b->cap);
    }
}

static char *pathJoin(const char *a, const char *b)
{
    if (b[0] == '/') return strdup(b);
    size_t n = strlen(a) + strlen(b) + 2;
    char *p = (char *)malloc(n);
    snprintf(p, n, "%s/%s", a, b);
    return p;
}

/* bFieldSize / bFieldPack (src/ioUtils.c:183-204): big-endian integer fields of the binary records */
static int bFieldSize(uint64_t i)
{
    int r = 0;
    do { i /= 256; r++; } while (i > 0);
    return r;
}
static void bFieldPack(unsigned char *buf, int size, uint64_t in)
{
    for (int k = 0; k < size; k++) { buf[(size - 1) - k] = (unsigned char)(in % 256); in /= 256; }
}

/* matinv of a diagonal box as src/three_algebra.c:37-64 evaluates it (cofactor / determinant): the factors Preduce uses */
static void diagInverse(const double h[9], double hi[3])
{
    const double d00 = h[4] * h[8] - h[5] * h[7], d11 = h[8] * h[0] - h[6] * h[2], d22 = h[0] * h[4] - h[1] * h[3];
    const double d01 = h[5] * h[6] - h[3] * h[8], d02 = h[3] * h[7] - h[6] * h[4];
    const double det = h[0] * d00 + h[1] * d01 + h[2] * d02;
    hi[0] = d00 / det;
    hi[1] = d11 / det;
    hi[2] = d22 / det;
}

/* CreateSnapshotdir's name (src/io.c:32-56): <atomsdir>/snapshot.<loopFormat>, or <atomsdir>/<dirname>, or an absolute dirname as it is;
 * relative to the deck's directory.  snprintf reports the untruncated length, so every piece is checked against the room left. */
static int snapshotRel(const ddcb200_deck *d, const char *dirname, int64_t loop, char *rel, size_t size)
{
    char loopFmt[16], num[64];
    snprintf(loopFmt, sizeof loopFmt, "%%%d.%dllu", d->nLoopDigits, d->nLoopDigits);
    snprintf(num, sizeof num, loopFmt, (unsigned long long)loop);
    int k;
    if (dirname && dirname[0] == '/') k = snprintf(rel, size, "%s", dirname);
    else if (dirname) k = snprintf(rel, size, "%s/%s", d->atomsdir, dirname);
    else k = snprintf(rel, size, "%s/snapshot.%s", d->atomsdir, num);
    if (k < 0 || (size_t)k >= size) return herr("snapshot directory name longer than %d characters", (int)size - 1);
    return 0;
}

int ddcb200_writeRestart(const ddcb200_deck *d, const char *dirname, int64_t loop, double time_, const double h[9],
                         const double *rx, const double *ry, const double *rz, const double *vx, const double *vy,
                         const double *vz, const uint64_t *rngState, int restartLink, char *snapshotdirOut, size_t len)
{
    if (!d || !h || !rx || !ry || !rz || !vx || !vy || !vz) return herr("writeRestart: null argument");
    for (int k = 0; k < 9; k++)
        if ((k % 4) != 0 && fabs(h[k]) > 1e-10) return herr("writeRestart: only orthorhombic boxes are supported");
    if (d->params.pbc != 7) return herr("writeRestart: only pbc = 7 is supported");
    if (!rngState) rngState = d->rngState;
    const int haveRandom = d->haveRandom && rngState && d->rngMult && d->rngPrime;

    /* CreateSnapshotdir: <atomsdir>/snapshot.<loopFormat> or <atomsdir>/<dirname>; relative to the deck's directory */
    char rel[1024];
    if (snapshotRel(d, dirname, loop, rel, sizeof rel) != 0) return -1;
    char *dir = pathJoin(d->runDir, rel);
    if (mkdir(dir, 0777) != 0 && errno != EEXIST) { herr("writeRestart: cannot create %s: %s", dir, strerror(errno)); free(dir); return -1; }
    if (snapshotdirOut) snprintf(snapshotdirOut, len, "%s", dir);

    /* ---- record format and length (src/collection_write.c:65-99) ---- */
    char fmt[160];
    snprintf(fmt, sizeof fmt, "%s %s %s", "%08x ", d->gidFormatHex ? "%16.16lx" : "%12.12lu", " %s %s %s %21.13e %21.13e %21.13e %21.13e %21.13e %21.13e");
    char line[MAXLREC + 64];
    int lrec = snprintf(line, MAXLREC, fmt, 0u, 0ul, " ", " ", " ", 0.0, 0.0, 0.0, 0.0, 0.0, 0.0);
    int maxType = 0, maxName = 0, maxGroup = 0;
    for (int s = 0; s < d->nspecies; s++)
    {
        int a = (int)strlen(d->speciesType[s]), b = (int)strlen(d->speciesName[s]);
        if (a > maxType) maxType = a;
        if (b > maxName) maxName = b;
    }
    const int nGroups = d->nGroups > 0 ? d->nGroups : 1;
    for (int g = 0; g < d->nGroups; g++)
    {
        int a = (int)strlen(d->groupName[g]);
        if (a > maxGroup) maxGroup = a;
    }
    if (d->nGroups <= 0) maxGroup = 5;   /* "group" */
    lrec += maxType - 1;
    lrec += maxName - 1;
    lrec += maxGroup - 1;
    int randomFieldSize = haveRandom ? 27 : 0;   /* strlen(lcg64_write(NULL)) = 16 + 1 + 1 + 1 + 8, src/lcg64.c:56-63 */
    lrec += randomFieldSize + 1;
    /* FREE and LANGEVIN groups have no per-particle write (GROUPMAXWRITELENGTH = 0) */
    lrec++;
    lrec = 8 * ((lrec + 7) / 8);
    /* binary records (collection_writeBLOCK_binary, src/collection_write.c:188-260): u4 crc | id | pinfo | 3 f8 | 3 f8 or f4 | LCG64 */
    const int binary = d->checkpointBinary != 0;
    int gidFieldSize = 1, pinfoFieldSize = 1, nTypes = 0;
    const int vsize = d->checkpointBrief ? 4 : 8;
    if (binary)
    {
        uint64_t gmax = 0;
        for (int64_t i = 0; i < d->n; i++)
            if (d->gid[i] > gmax) gmax = d->gid[i];
        for (int s = 0; s < d->nspecies; s++)
        {
            int seen = 0;
            for (int t = 0; t < s && !seen; t++) seen = strcmp(d->speciesType[t], d->speciesType[s]) == 0;
            nTypes += !seen;
        }
        if (nTypes != 1) { herr("writeRestart: binary restarts need a single SPECIES type (found %d)", nTypes); free(dir); return -1; }
        gidFieldSize = bFieldSize(gmax);                                                     /* bFieldSize(mpiMaxVal(label)) */
        pinfoFieldSize = bFieldSize((uint64_t)nGroups * (uint64_t)d->nspecies * (uint64_t)nTypes);   /* pinfoMaxIndex */
        randomFieldSize = haveRandom ? 16 : 0;                                               /* lcg64_bwrite, src/lcg64.c:75-84 */
        lrec = 4 + gidFieldSize + pinfoFieldSize + 24 + 3 * vsize + randomFieldSize;
    }
    if (lrec > MAXLREC) { herr("writeRestart: record length %d exceeds MAXLREC=%d", lrec, MAXLREC); free(dir); return -1; }

    const double cLen = hu_convert(1.0, NULL, "l"), cTime = hu_convert(1.0, NULL, "t"), cVel = cLen / cTime;

    /* ---- FILEHEADER (write_fileheader, src/io.c:352-407) ---- */
    SBuf hb = {0};
    {
        time_t now = time(NULL);
        char stamp[64];
        snprintf(stamp, sizeof stamp, "%s", ctime(&now));
        stamp[strcspn(stamp, "\n")] = 0;
        sbCat(&hb, "particle FILEHEADER {type=MULTILINE; datatype=%s; checksum=CRC32; create_time=%s; run_id=0x%08x;\n",
              binary ? "FIXRECORDBINARY" : "FIXRECORDASCII", stamp, d->runId);
        sbCat(&hb, "code_version=ddcmd_b200 (B200-native Martini step); srcpath=ddcmd_b200;\n");
        sbCat(&hb, "loop=%lld; time=%f fs;\n", (long long)loop, time_ * cTime);
        sbCat(&hb, "nfiles=1; nrecord=%llu; lrec=%d; nfields=%d; endian_key=%d;\n", (unsigned long long)d->n, lrec, binary ? 9 : 11, 875770417);
        if (binary)
        {
            /* src/collection_write.c:207-225,255-256: " %.1s%1d" per field */
            sbCat(&hb, "field_names=checksum id pinfo rx  ry  rz  vx  vy  vz;\n");
            sbCat(&hb, "field_types= u4 b%d b%d f8 f8 f8 f%d f%d f%d;\n", gidFieldSize, pinfoFieldSize, vsize, vsize, vsize);
        }
        else
        {
            sbCat(&hb, "field_names=checksum id class type group rx ry rz vx vy vz;\n");
            sbCat(&hb, "field_types=u u s s s f f f f f f;\n");
            sbCat(&hb, "field_units=1 1 1 1 1 Ang Ang Ang Ang/fs Ang/fs Ang/fs;\n");
            sbCat(&hb, "field_format=%s;\n", fmt);
        }
        sbCat(&hb, "reducedcorner=%21.14f %21.14f %21.14f;\n", d->reducedCorner[0], d->reducedCorner[1], d->reducedCorner[2]);
        sbCat(&hb, "h=%21.14f %21.14f %21.14f\n", h[0] * cLen, h[1] * cLen, h[2] * cLen);
        sbCat(&hb, "  %21.14f %21.14f %21.14f\n", h[3] * cLen, h[4] * cLen, h[5] * cLen);
        sbCat(&hb, "  %21.14f %21.14f %21.14f Ang;\n", h[6] * cLen, h[7] * cLen, h[8] * cLen);
        /* misc_info: every PioSet appends its string and one blank (src/pio.c:209-216) */
        sbCat(&hb, "random = %s ;\nrandomFieldSize = %d ;\ngroups = ", haveRandom ? "lcg64" : "NONE", randomFieldSize);
        if (d->nGroups > 0) for (int g = 0; g < d->nGroups; g++) sbCat(&hb, "%s ", d->groupName[g]);
        else sbCat(&hb, "group ");
        sbCat(&hb, ";\nspecies = ");
        for (int s = 0; s < d->nspecies; s++) sbCat(&hb, "%s ", d->speciesName[s]);
        sbCat(&hb, ";\ntypes = ");
        for (int s = 0; s < d->nspecies; s++)
        {
            int seen = 0;
            for (int t = 0; t < s && !seen; t++) seen = strcmp(d->speciesType[t], d->speciesType[s]) == 0;
            if (!seen) sbCat(&hb, "%s ", d->speciesType[s]);
        }
        sbCat(&hb, "; \n}\n \n\n");
    }

    char *apath = pathJoin(dir, "atoms#000000");
    FILE *f = fopen(apath, "wb");
    if (!f) { herr("writeRestart: cannot open %s: %s", apath, strerror(errno)); free(apath); free(dir); free(hb.p); return -1; }
    int rc = 0;
    if (fwrite(hb.p, 1, hb.n, f) != hb.n) rc = herr("writeRestart: short write to %s", apath);
    free(hb.p);

    /* ---- records (src/collection_write.c:150-184) ---- */
    double hi[3];
    diagInverse(h, hi);
    for (int64_t i = 0; i < d->n && rc == 0; i++)
    {
        /* backInBox = Preduce, pbc 7 (src/preduce.c:282-340): r += h * (-rint(hinv r)) */
        double x = rx[i], y = ry[i], z = rz[i];
        x += h[0] * -rint(hi[0] * x);
        y += h[4] * -rint(hi[1] * y);
        z += h[8] * -rint(hi[2] * z);
        const int s = d->species[i];
        if (binary)
        {
            unsigned char *b = (unsigned char *)line;
            memset(b, 0, (size_t)lrec);
            const int ig = (d->nGroups > 0 && d->groupOfBead) ? d->groupOfBead[i] : 0;
            int o = 4;
            bFieldPack(b + o, gidFieldSize, d->gid[i]);
            o += gidFieldSize;
            bFieldPack(b + o, pinfoFieldSize, (uint64_t)ig + (uint64_t)s * (uint64_t)nGroups);    /* pinfoEncode with one type */
            o += pinfoFieldSize;
            double f8[6] = {x * cLen, y * cLen, z * cLen, vx[i] * cVel, vy[i] * cVel, vz[i] * cVel};
            memcpy(b + o, f8, 24);
            o += 24;
            if (vsize == 8) memcpy(b + o, f8 + 3, 24);
            else
                for (int k = 0; k < 3; k++) { float f4 = (float)f8[3 + k]; memcpy(b + o + 4 * k, &f4, 4); }
            o += 3 * vsize;
            if (haveRandom)
            {
                memcpy(b + o, &rngState[i], 8);
                memcpy(b + o + 8, &d->rngMult[i], 4);
                memcpy(b + o + 12, &d->rngPrime[i], 4);
            }
            const uint32_t crc = hcrc32(b + 4, (size_t)lrec - 4);
            memcpy(b, &crc, 4);
            if (fwrite(b, 1, (size_t)lrec, f) != (size_t)lrec) rc = herr("writeRestart: short write to %s", apath);
            continue;
        }
        const char *gname = (d->nGroups > 0 && d->groupOfBead) ? d->groupName[d->groupOfBead[i]] : (d->nGroups > 0 ? d->groupName[0] : "group");
        int length = snprintf(line, MAXLREC, fmt, 0u, (unsigned long)d->gid[i], d->speciesType[s], d->speciesName[s], gname,
                              x * cLen, y * cLen, z * cLen, vx[i] * cVel, vy[i] * cVel, vz[i] * cVel);
        if (haveRandom)
            length += snprintf(line + length, (size_t)(MAXLREC - length), " %16.16llx %1u %8.8x", (unsigned long long)rngState[i], d->rngMult[i], d->rngPrime[i]);
        if (length > lrec - 1) { rc = herr("writeRestart: record of gid %llu is longer than lrec=%d", (unsigned long long)d->gid[i], lrec); break; }
        for (int l = length; l < lrec; l++) line[l] = ' ';
        line[lrec - 1] = '\n';
        char ck[16];
        snprintf(ck, sizeof ck, "%08x", hcrc32((const unsigned char *)line + 8, (size_t)lrec - 8));
        memcpy(line, ck, 8);
        if (fwrite(line, 1, (size_t)lrec, f) != (size_t)lrec) rc = herr("writeRestart: short write to %s", apath);
    }
    if (fclose(f) != 0 && rc == 0) rc = herr("writeRestart: closing %s failed", apath);
    free(apath);
    if (rc) { free(dir); return rc; }

    /* ---- restart (src/io.c:75-104) ---- */
    char *rpath = pathJoin(dir, "restart");
    f = fopen(rpath, "w");
    if (!f) { herr("writeRestart: cannot open %s: %s", rpath, strerror(errno)); free(rpath); free(dir); return -1; }
    fprintf(f, "%s SIMULATE { run_id=0x%08x; loop=%lld; time=%f fs;}\n", d->simulateName, d->runId, (long long)loop, time_ * cTime);
    fprintf(f, "%s BOX {\n h  = ", d->boxName);
    fprintf(f, "%21.14e %21.14e %21.14e\n      %21.14e %21.14e %21.14e\n      %21.14e %21.14e %21.14e;\n", h[0] * cLen, h[1] * cLen, h[2] * cLen,
            h[3] * cLen, h[4] * cLen, h[5] * cLen, h[6] * cLen, h[7] * cLen, h[8] * cLen);
    fprintf(f, "}\n");
    /* NGLF / NGLFCONSTRAINT write nothing (writeNULL, src/integrator.c:48); LANGEVIN groups with an explicit Teq write it */
    for (int g = 0; g < d->nGroups; g++)
        if (d->groupType[g] == 1) fprintf(f, "%s GROUP { Teq=%f ;}\n", d->groupName[g], hu_convert(d->groupTeq[g], NULL, "T"));
    fprintf(f, "%s COLLECTION { size=%llu; files=%s/atoms#;}\n", d->collectionName, (unsigned long long)d->n, rel);
    /* a truncated checkpoint (full disk) must not become ./restart */
    const int werr = ferror(f);
    if (fclose(f) != 0 || werr) { herr("writeRestart: writing %s failed", rpath); free(rpath); free(dir); return -1; }
    if (restartLink)
    {
        /* unlink("restart"); symlink(<snapshotdir>/restart, "restart") in the run directory */
        char *lpath = pathJoin(d->runDir, "restart");
        char target[1100];
        snprintf(target, sizeof target, "%s/restart", rel);
        unlink(lpath);
        if (symlink(target, lpath) != 0) { herr("writeRestart: cannot link %s: %s", lpath, strerror(errno)); rc = -1; }
        free(lpath);
    }
    free(rpath);
    free(dir);
    return rc;
}

int ddcb200_writeBXYZ(const ddcb200_deck *d, const char *dirname, int64_t loop, double time_, const double h[9],
                      const double *rx, const double *ry, const double *rz, const double *vx, const double *vy, const double *vz)
{
    /* writeBXYZ (src/io.c:144-155) + collection_writeBXYZ mode 1 (src/collection_write.c:338-465): crc u4 | id | pinfo | r f4 x3 |
     * v f4 x3 | energy f4 | virial f4.  The Martini path keeps no per-particle energy or virial (they stay at zeroAll's 0). */
    if (!d || !h || !rx || !ry || !rz || !vx || !vy || !vz) return herr("writeBXYZ: null argument");
    char rel[1024];
    if (snapshotRel(d, dirname, loop, rel, sizeof rel) != 0) return -1;
    char *dir = pathJoin(d->runDir, rel);
    if (mkdir(dir, 0777) != 0 && errno != EEXIST) { herr("writeBXYZ: cannot create %s: %s", dir, strerror(errno)); free(dir); return -1; }
    char *path = pathJoin(dir, "bxyz#000000");
    free(dir);
    const int nGroups = d->nGroups > 0 ? d->nGroups : 1;
    uint64_t gmax = 0;
    for (int64_t i = 0; i < d->n; i++)
        if (d->gid[i] > gmax) gmax = d->gid[i];
    const int gs = bFieldSize(gmax), ps = bFieldSize((uint64_t)nGroups * (uint64_t)d->nspecies);
    const int lrec = 9 * 4 + gs + ps;
    const double cLen = hu_convert(1.0, NULL, "l"), cTime = hu_convert(1.0, NULL, "t"), cVel = cLen / cTime;
    SBuf hb = {0};
    time_t now = time(NULL);
    char stamp[64];
    snprintf(stamp, sizeof stamp, "%s", ctime(&now));
    stamp[strcspn(stamp, "\n")] = 0;
    sbCat(&hb, "bxyz FILEHEADER {type=MULTILINE; datatype=FIXRECORDBINARY; checksum=CRC32; create_time=%s; run_id=0x%08x;\n", stamp, d->runId);
    sbCat(&hb, "code_version=ddcmd_b200 (B200-native Martini step); srcpath=ddcmd_b200;\n");
    sbCat(&hb, "loop=%lld; time=%f fs;\n", (long long)loop, time_ * cTime);
    sbCat(&hb, "nfiles=1; nrecord=%llu; lrec=%d; nfields=11; endian_key=%d;\n", (unsigned long long)d->n, lrec, 875770417);
    sbCat(&hb, "field_names=checksum id pinfo rx ry rz vx vy vz energy virial ;\n");
    sbCat(&hb, "field_types=u4 b%d b%d f4 f4 f4 f4 f4 f4 f4 f4 ;\n", gs, ps);
    sbCat(&hb, "reducedcorner=%21.14f %21.14f %21.14f;\n", d->reducedCorner[0], d->reducedCorner[1], d->reducedCorner[2]);
    sbCat(&hb, "h=%21.14f %21.14f %21.14f\n", h[0] * cLen, h[1] * cLen, h[2] * cLen);
    sbCat(&hb, "  %21.14f %21.14f %21.14f\n", h[3] * cLen, h[4] * cLen, h[5] * cLen);
    sbCat(&hb, "  %21.14f %21.14f %21.14f Ang;\n", h[6] * cLen, h[7] * cLen, h[8] * cLen);
    sbCat(&hb, "groups = ");
    if (d->nGroups > 0) for (int g = 0; g < d->nGroups; g++) sbCat(&hb, "%s ", d->groupName[g]);
    else sbCat(&hb, "group ");
    sbCat(&hb, ";\nspecies = ");
    for (int s = 0; s < d->nspecies; s++) sbCat(&hb, "%s ", d->speciesName[s]);
    sbCat(&hb, ";\ntypes = ");
    for (int s = 0; s < d->nspecies; s++)
    {
        int seen = 0;
        for (int t = 0; t < s && !seen; t++) seen = strcmp(d->speciesType[t], d->speciesType[s]) == 0;
        if (!seen) sbCat(&hb, "%s ", d->speciesType[s]);
    }
    sbCat(&hb, "; \n}\n \n\n");
    FILE *f = fopen(path, "wb");
    if (!f) { herr("writeBXYZ: cannot open %s: %s", path, strerror(errno)); free(path); free(hb.p); return -1; }
    fwrite(hb.p, 1, hb.n, f);
    free(hb.p);
    double hi[3];
    diagInverse(h, hi);
    int rc = 0;
    for (int64_t i = 0; i < d->n; i++)
    {
        unsigned char b[64];
        memset(b, 0, sizeof b);
        double x = rx[i], y = ry[i], z = rz[i];
        x += h[0] * -rint(hi[0] * x);
        y += h[4] * -rint(hi[1] * y);
        z += h[8] * -rint(hi[2] * z);
        const int ig = (d->nGroups > 0 && d->groupOfBead) ? d->groupOfBead[i] : 0;
        int o = 4;
        bFieldPack(b + o, gs, d->gid[i]);
        o += gs;
        bFieldPack(b + o, ps, (uint64_t)ig + (uint64_t)d->species[i] * (uint64_t)nGroups);
        o += ps;
        const float f4[6] = {(float)(x * cLen), (float)(y * cLen), (float)(z * cLen), (float)(vx[i] * cVel), (float)(vy[i] * cVel), (float)(vz[i] * cVel)};
        memcpy(b + o, f4, 24);
        const uint32_t crc = hcrc32(b + 4, (size_t)lrec - 4);
        memcpy(b, &crc, 4);
        if (fwrite(b, 1, (size_t)lrec, f) != (size_t)lrec) { rc = herr("writeBXYZ: short write to %s", path); break; }
    }
    fclose(f);
    free(path);
    return rc;
}

int64_t ddcb200_subsetWrite(const ddcb200_deck *d, int which, const char *dirname, int64_t loop, double time_, const double h[9],
                            const double *rx, const double *ry, const double *rz, const double *vx, const double *vy, const double *vz)
{
    if (!d || !h || !rx || !ry || !rz || !vx || !vy || !vz) return herr("subsetWrite: null argument");
    if (which < 0 || which >= d->nSubsets) return herr("subsetWrite: no such ANALYSIS (%d of %d)", which, d->nSubsets);
    const ddcb200_subset *q = &d->subsets[which];
    /* CreateSnapshotdir(simulate, NULL) + "<snapshotdir>/<filename>" (src/subsetWrite.c:441-444) */
    char rel[1024];
    if (snapshotRel(d, dirname, loop, rel, sizeof rel) != 0) return -1;
    char *dir = pathJoin(d->runDir, rel);
    if (mkdir(dir, 0777) != 0 && errno != EEXIST) { herr("subsetWrite: cannot create %s: %s", dir, strerror(errno)); free(dir); return -1; }
    char fname[300];
    snprintf(fname, sizeof fname, "%s#000000", q->filename);
    char *path = pathJoin(dir, fname);
    free(dir);

    /* rejectParticle (src/subsetWrite.c:532-564) on the positions as they are in the state (the integrator keeps them in the box) */
    const int nGroups = d->nGroups > 0 ? d->nGroups : 1;
    unsigned char *keep = (unsigned char *)malloc((size_t)d->n + 1);
    int64_t nrec = 0;
    for (int64_t i = 0; i < d->n; i++)
    {
        const uint64_t gid = d->gid[i];
        int rej = gid < q->idMin || gid > q->idMax || (gid % (uint64_t)q->modulus) != 0 || (q->odd && gid % 2 == 0) ||
                  rx[i] > q->hi[0] || ry[i] > q->hi[1] || rz[i] > q->hi[2] || rx[i] < q->lo[0] || ry[i] < q->lo[1] || rz[i] < q->lo[2] ||
                  vx[i] > q->vhi[0] || vy[i] > q->vhi[1] || vz[i] > q->vhi[2] || vx[i] < q->vlo[0] || vy[i] < q->vlo[1] || vz[i] < q->vlo[2] ||
                  q->includeSpecies[d->species[i]] == 0;
        if (!rej && q->idList)
        {
            int64_t lo = 0, hi = q->nIdList;
            while (lo < hi)
            {
                const int64_t mid = (lo + hi) / 2;
                if (q->idList[mid] < gid) lo = mid + 1;
                else hi = mid;
            }
            rej = !(lo < q->nIdList && q->idList[lo] == gid);
        }
        keep[i] = (unsigned char)!rej;
        nrec += !rej;
    }
    const double cLen = hu_convert(1.0, NULL, "l"), cTime = hu_convert(1.0, NULL, "t");
    SBuf hb = {0};
    time_t now = time(NULL);
    char stamp[64];
    snprintf(stamp, sizeof stamp, "%s", ctime(&now));
    stamp[strcspn(stamp, "\n")] = 0;
    sbCat(&hb, "subset FILEHEADER {type=MULTILINE; datatype=FIXRECORDBINARY; checksum=NONE; create_time=%s; run_id=0x%08x;\n", stamp, d->runId);
    sbCat(&hb, "code_version=ddcmd_b200 (B200-native Martini step); srcpath=ddcmd_b200;\n");
    sbCat(&hb, "loop=%lld; time=%f fs;\n", (long long)loop, time_ * cTime);
    sbCat(&hb, "nfiles=1; nrecord=%llu; lrec=24; nfields=5; endian_key=%d;\n", (unsigned long long)nrec, 875770417);
    sbCat(&hb, "field_names=id pinfo rx ry  rz;\nfield_types= u8 u4 f4 f4 f4;\nfield_units=1 1 %s %s %s;\n", q->lengthUnit, q->lengthUnit, q->lengthUnit);
    sbCat(&hb, "reducedcorner=%21.14f %21.14f %21.14f;\n", d->reducedCorner[0], d->reducedCorner[1], d->reducedCorner[2]);
    sbCat(&hb, "h=%21.14f %21.14f %21.14f\n", h[0] * cLen, h[1] * cLen, h[2] * cLen);
    sbCat(&hb, "  %21.14f %21.14f %21.14f\n", h[3] * cLen, h[4] * cLen, h[5] * cLen);
    sbCat(&hb, "  %21.14f %21.14f %21.14f Ang;\n", h[6] * cLen, h[7] * cLen, h[8] * cLen);
    sbCat(&hb, "random = NONE;\n nrandomFieldSize = 0;\n types = ");
    for (int s = 0; s < d->nspecies; s++)
    {
        int seen = 0;
        for (int t = 0; t < s && !seen; t++) seen = strcmp(d->speciesType[t], d->speciesType[s]) == 0;
        if (!seen) sbCat(&hb, "%s ", d->speciesType[s]);
    }
    sbCat(&hb, ";\n groups = ");
    if (d->nGroups > 0) for (int g = 0; g < d->nGroups; g++) sbCat(&hb, "%s ", d->groupName[g]);
    else sbCat(&hb, "group ");
    sbCat(&hb, ";\n species = ");
    for (int s = 0; s < d->nspecies; s++) sbCat(&hb, "%s ", d->speciesName[s]);
    sbCat(&hb, ";\n %s \n}\n \n\n", q->parmsInfo);
    FILE *f = fopen(path, "wb");
    if (!f) { herr("subsetWrite: cannot open %s: %s", path, strerror(errno)); free(path); free(keep); free(hb.p); return -1; }
    fwrite(hb.p, 1, hb.n, f);
    free(hb.p);
    /* box corner = h * reducedcorner (src/box.c:46); positions relative to it, in lengthUnit, as floats */
    const double corner[3] = {h[0] * d->reducedCorner[0], h[4] * d->reducedCorner[1], h[8] * d->reducedCorner[2]};
    const double cL = hu_convert(1.0, NULL, q->lengthUnit);
    int64_t rc = nrec;
    for (int64_t i = 0; i < d->n; i++)
    {
        if (!keep[i]) continue;
        unsigned char line[24];
        const uint64_t gid = d->gid[i];
        const int ig = (d->nGroups > 0 && d->groupOfBead) ? d->groupOfBead[i] : 0;
        const uint32_t index = (uint32_t)ig + (uint32_t)d->species[i] * (uint32_t)nGroups;     /* pinfoEncode, one type */
        const float f4[3] = {(float)((rx[i] - corner[0]) * cL), (float)((ry[i] - corner[1]) * cL), (float)((rz[i] - corner[2]) * cL)};
        memcpy(line, &gid, 8);
        memcpy(line + 8, &index, 4);
        memcpy(line + 12, f4, 12);
        if (fwrite(line, 1, 24, f) != 24) { rc = herr("subsetWrite: short write to %s", path); break; }
    }
    fclose(f);
    free(path);
    free(keep);
    return rc;
}

static int comboIndexHost(int i, int j, int ns)
{
    const int mx = i > j ? i : j, mn = i > j ? j : i;      /* comboIndex, src/paircorrelation.c:516-521 */
    return (mx - mn) + ns * mn - (mn * (mn - 1)) / 2;
}

int ddcb200_pairCorrelationWrite(const ddcb200_deck *d, int which, const char *dirname, int64_t loop, double volume, const double *gacc, int nsample)
{
    if (!d || !gacc) return herr("pairCorrelationWrite: null argument");
    if (which < 0 || which >= d->nPairCorr) return herr("pairCorrelationWrite: no such ANALYSIS (%d of %d)", which, d->nPairCorr);
    if (nsample <= 0) return 0;                              /* nothing sampled: the reference writes nothing */
    const ddcb200_paircorr *q = &d->pairCorr[which];
    char rel[1024];
    if (snapshotRel(d, dirname, loop, rel, sizeof rel) != 0) return -1;
    char *dir = pathJoin(d->runDir, rel);
    if (mkdir(dir, 0777) != 0 && errno != EEXIST) { herr("pairCorrelationWrite: cannot create %s: %s", dir, strerror(errno)); free(dir); return -1; }
    char *path = pathJoin(dir, q->filename);
    free(dir);
    FILE *f = fopen(path, "w");
    if (!f) { herr("pairCorrelationWrite: cannot open %s: %s", path, strerror(errno)); free(path); return -1; }
    const int ns = d->nspecies, np = ns * (ns + 1) / 2, nb = q->nBins;
    /* bin edges as paircorrelation_parms makes them (src/paircorrelation.c:100-131), scaling of :470-479 */
    const double s = volume / nsample;
    fprintf(f, "# %s\n# nsample = %d;\n# r(Ang) ", q->miscInfo, nsample);
    for (int l = 0; l < np; l++)
    {
        int ti = -1, tj = -1;
        for (int a = 0; a < ns && ti < 0; a++)
            for (int b = a; b < ns; b++)
                if (comboIndexHost(a, b, ns) == l) { ti = a; tj = b; break; }
        fprintf(f, "%s-%s ", d->speciesName[ti], d->speciesName[tj]);
    }
    fprintf(f, "\n");
    for (int kb = 0; kb < nb; kb++)
    {
        double r0, r1;
        if (q->logScale)
        {
            r0 = pow(10, log10(q->rmin) + kb * q->logDelta);
            r1 = kb == nb - 1 ? q->rmax : pow(10, log10(q->rmin) + (kb + 1) * q->logDelta);
        }
        else
        {
            r0 = q->rmin + kb * q->deltaR;
            r1 = q->rmin + (kb + 1) * q->deltaR;
        }
        const double dv = 4.0 * M_PI / 3.0 * (r1 * r1 * r1 - r0 * r0 * r0);
        fprintf(f, "%f ", hu_convert(0.5 * (r0 + r1), NULL, "Angstrom"));
        for (int l = 0; l < np; l++)
        {
            double g = gacc[kb + nb * l];
            g *= s / dv;
            fprintf(f, "%e ", g);
        }
        fprintf(f, "\n");
    }
    fclose(f);
    free(path);
    return 0;
}

int ddcb200_readCMDS(const char *filename)
{
    int flag = 0;
    FILE *file = fopen(filename, "r");
    if (!file) return 0;
    char line[256];
    line[0] = 0;
    if (fgets(line, 255, file))
    {
        if (strchr(line, '{')) flag |= DDCB200_CMD_NEW_OBJECT;   /* object text: reported, not applied (object_rescan is out of scope) */
        do
        {
            line[strcspn(line, "\n")] = 0;
            printf("Received command \"%s\" from ddcMD_CMDS\n", line);
            if (strcmp(line, "checkpoint") == 0) flag |= DDCB200_CMD_CHECKPOINT;
            if (strcmp(line, "kill") == 0) flag |= DDCB200_CMD_STOP;
            if (strcmp(line, "exit") == 0) flag |= DDCB200_CMD_STOP | DDCB200_CMD_CHECKPOINT;
            if (strcmp(line, "profile") == 0) flag |= DDCB200_CMD_DUMP_PROFILE;
            if (strcmp(line, "hpm") == 0) flag |= DDCB200_CMD_HPM_PRINT;
            if (strcmp(line, "analysis") == 0) flag |= DDCB200_CMD_DO_ANALYSIS;
        } while (fgets(line, 255, file));
    }
    fclose(file);
    if (truncate(filename, 0) != 0) { /* the reference ignores this too */ }
    return flag;
}

#define TEST0(A, B) ((B) != 0 && ((A) % (B)) == 0)   /* src/object.h:17 */

static int64_t findEndLoop(const ddcb200_deck *d, int64_t loop, int64_t maxloop)
{
    /* the next loop at which something is printed or written (src/masters.c:263-281) */
    for (int64_t l = loop + 1; l < maxloop; l++)
    {
        if (TEST0(l, d->printrate) || TEST0(l, d->snapshotrate) || TEST0(l, d->checkpointrate)) return l;
        for (int a = 0; a < d->nSubsets; a++)
            if (TEST0(l, d->subsets[a].evalRate) || TEST0(l, d->subsets[a].outputRate)) return l;
        for (int a = 0; a < d->nPairCorr; a++)
            if (TEST0(l, d->pairCorr[a].evalRate) || TEST0(l, d->pairCorr[a].outputRate)) return l;
    }
    return maxloop;
}

typedef struct { double *r[6]; uint64_t *rng; } HostState;

static int checkpoint(ddcb200_deck *d, ddcb200_ctx *c, HostState *hs, const ddcb200_etype *e)
{
    double h[9];
    if (ddcb200_getState(c, hs->r[0], hs->r[1], hs->r[2], hs->r[3], hs->r[4], hs->r[5], NULL, NULL, NULL)) return herr("getState: %s", ddcb200_lastError());
    if (ddcb200_getBox(c, h)) return herr("getBox: %s", ddcb200_lastError());
    if (d->haveRandom && ddcb200_getRandom(c, d->n, hs->rng)) return herr("getRandom: %s", ddcb200_lastError());
    char where[1024];
    int rc = ddcb200_writeRestart(d, NULL, e->loop, e->time, h, hs->r[0], hs->r[1], hs->r[2], hs->r[3], hs->r[4], hs->r[5],
                                  d->haveRandom ? hs->rng : NULL, 1, where, sizeof where);
    if (rc == 0) printf("Wrote restart %s\n", where);
    return rc;
}

int ddcb200_simulateMaster(const char *objectFile, const char *restartFile, const char *simulateName, int device)
{
    ddcb200_deck *d = NULL;
    ddcb200_ctx *c = NULL;
    HostState hs;
    memset(&hs, 0, sizeof hs);
    FILE *data = NULL;
    char *cmds = NULL;
    double *pcG[16] = {0};
    int pcSamples[16] = {0};
    int rc = ddcb200_deckLoad(objectFile, restartFile, simulateName, &d);
    if (rc) return rc;
    rc = ddcb200_simulateBind(d, device, &c);
    if (rc) { ddcb200_deckFree(d); return rc; }
#define CK(call, what) do { if ((call) != 0) { rc = herr(what ": %s", ddcb200_lastError()); goto done; } } while (0)
    for (int k = 0; k < 6; k++) hs.r[k] = (double *)malloc(sizeof(double) * (size_t)(d->n + 1));
    hs.rng = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)(d->n + 1));
    int64_t loop = d->loop, maxloop = d->maxloop;
    const int64_t startLoop = loop;
    for (int a = 0; a < d->nPairCorr && a < 16; a++)
        pcG[a] = (double *)calloc((size_t)d->pairCorr[a].nBins * (size_t)(d->nspecies * (d->nspecies + 1) / 2) + 1, sizeof(double));
    if (d->nPairCorr > 16) { rc = herr("more than 16 PAIRCORRELATION analyses"); goto done; }
    if (d->deltaloop > -1 && loop + d->deltaloop < maxloop) maxloop = loop + d->deltaloop;   /* src/simulate.c:242 */
    {
        char *dpath = pathJoin(d->runDir, "data");
        data = fopen(dpath, "a");
        if (!data) { rc = herr("cannot open %s", dpath); free(dpath); goto done; }
        free(dpath);
    }
    cmds = pathJoin(d->runDir, "ddcMD_CMDS");
    char buf[1024];
    ddcb200_etype e;
    /* firstEnergyCall (src/masters.c:579-620) + the first printinfoAll with the header */
    CK(ddcb200_ddcenergy(c, 1), "ddcenergy");
    CK(ddcb200_energyInfo(c, d->kB, &e), "energyInfo");
    ddcb200_printinfoHeader(d, buf, sizeof buf);
    printf("%s\n", buf);
    fprintf(data, "%s\n", buf);
#define PRINTLINE() do { ddcb200_getBox(c, d->params.h); ddcb200_printinfoLine(d, &e, buf, sizeof buf); printf("%s\n", buf); \
                         fprintf(data, "%s\n", buf); fflush(stdout); fflush(data); } while (0)
    PRINTLINE();
    while (loop < maxloop)
    {
        const int64_t endLoop = findEndLoop(d, loop, maxloop);
        const int n = (int)(endLoop - loop);
        if (d->integratorType == 1) CK(ddcb200_nglfconstraint(c, n, d->dt), "nglfconstraint");
        else CK(ddcb200_nglf(c, n, d->dt), "nglf");
        loop = endLoop;
        CK(ddcb200_energyInfo(c, d->kB, &e), "energyInfo");
        int flag = 0;
        if (!isfinite(e.eion))
        {
            /* src/masters.c:467-472 */
            flag = DDCB200_CMD_STOP;
            printf("eion = %e is bad. Simulation is being killed at loop = %lld\n", e.eion, (long long)loop);
            PRINTLINE();
            rc = herr("eion is not finite at loop %lld", (long long)loop);
            break;
        }
        if (TEST0(loop, d->printrate))
        {
            PRINTLINE();
            flag = ddcb200_readCMDS(cmds);
        }
        if (TEST0(loop, d->checkpointrate) || (flag & DDCB200_CMD_CHECKPOINT))
            if ((rc = checkpoint(d, c, &hs, &e)) != 0) break;
        int fetched = 0;
        /* doSnapshot (src/masters.c:340-352): bxyz every snapshotrate loops once the run has advanced */
        if (loop > startLoop && TEST0(loop, d->snapshotrate))
        {
            double hh[9];
            if (ddcb200_getState(c, hs.r[0], hs.r[1], hs.r[2], hs.r[3], hs.r[4], hs.r[5], NULL, NULL, NULL)) { rc = herr("getState: %s", ddcb200_lastError()); break; }
            fetched = 1;
            if (ddcb200_getBox(c, hh)) { rc = herr("getBox: %s", ddcb200_lastError()); break; }
            if ((rc = ddcb200_writeBXYZ(d, NULL, e.loop, e.time, hh, hs.r[0], hs.r[1], hs.r[2], hs.r[3], hs.r[4], hs.r[5])) != 0) break;
        }
        /* doAnalysis (src/masters.c:295-302) */
        for (int a = 0; a < d->nSubsets && rc == 0; a++)
        {
            if (!(TEST0(loop, d->subsets[a].outputRate) || (flag & DDCB200_CMD_DO_ANALYSIS))) continue;
            double hh[9];
            if (!fetched && ddcb200_getState(c, hs.r[0], hs.r[1], hs.r[2], hs.r[3], hs.r[4], hs.r[5], NULL, NULL, NULL)) { rc = herr("getState: %s", ddcb200_lastError()); break; }
            fetched = 1;
            if (ddcb200_getBox(c, hh)) { rc = herr("getBox: %s", ddcb200_lastError()); break; }
            if (ddcb200_subsetWrite(d, a, NULL, e.loop, e.time, hh, hs.r[0], hs.r[1], hs.r[2], hs.r[3], hs.r[4], hs.r[5]) < 0) rc = -1;
        }
        if (rc) break;
        for (int a = 0; a < d->nPairCorr && rc == 0; a++)
        {
            const ddcb200_paircorr *q = &d->pairCorr[a];
            const int ns = d->nspecies, np = ns * (ns + 1) / 2;
            const size_t nh = (size_t)q->nBins * (size_t)np;
            if (TEST0(loop, q->evalRate) || (flag & DDCB200_CMD_DO_ANALYSIS))
            {
                /* paircorrelation_eval: g += counts / (N_i N_j) (src/paircorrelation.c:222-235) */
                unsigned long long *cnt = (unsigned long long *)malloc(sizeof(unsigned long long) * (nh + (size_t)ns));
                if (ddcb200_pairCorrelation(c, q->nBins, q->rmin, q->logScale ? q->logDelta : q->deltaR, q->logScale, q->rmax, cnt, cnt + nh))
                {
                    rc = herr("ANALYSIS %s: %s", q->name, ddcb200_lastError());
                    free(cnt);
                    break;
                }
                for (int si = 0; si < ns; si++)
                    for (int sj = si; sj < ns; sj++)
                    {
                        const int l = comboIndexHost(si, sj, ns);
                        const double recip = 1.0 / ((double)cnt[nh + (size_t)si] * (double)cnt[nh + (size_t)sj]);
                        for (int kb = 0; kb < q->nBins; kb++)
                        {
                            double v = (double)cnt[(size_t)kb + (size_t)q->nBins * (size_t)l];
                            v *= recip;
                            pcG[a][(size_t)kb + (size_t)q->nBins * (size_t)l] += v;
                        }
                    }
                pcSamples[a]++;
                free(cnt);
            }
            if (TEST0(loop, q->outputRate) || (flag & DDCB200_CMD_DO_ANALYSIS))
            {
                double hh[9];
                if (ddcb200_getBox(c, hh)) { rc = herr("getBox: %s", ddcb200_lastError()); break; }
                rc = ddcb200_pairCorrelationWrite(d, a, NULL, e.loop, hh[0] * hh[4] * hh[8], pcG[a], pcSamples[a]);
                memset(pcG[a], 0, sizeof(double) * nh);       /* paircorrelation_clear */
                pcSamples[a] = 0;
            }
        }
        if (rc) break;
        if (flag & DDCB200_CMD_STOP) break;
    }
    if (rc == 0 && !TEST0(loop, d->printrate)) PRINTLINE();
done:
    free(cmds);
    for (int a = 0; a < 16; a++) free(pcG[a]);
    if (data) fclose(data);
    for (int k = 0; k < 6; k++) free(hs.r[k]);
    free(hs.rng);
    ddcb200_destroy(c);
    ddcb200_deckFree(d);
    return rc;
#undef CK
#undef PRINTLINE
}

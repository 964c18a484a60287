/* deck.c - ddcMD object-database front end for the Martini MD step (plain C host code).
 *
 * Mirrors, for the subset of objects a Martini deck uses, the reference's init chain:
 *   simulate_init (src/simulate.c:104-297) -> system_init (src/system.c:79-217)
 *   -> moleculeClassInit/moleculeInit/species_init (src/molecule.c:20-58,212-241, src/species.c:20-40)
 *   -> box_init (src/box.c:50-87) -> collection read (src/collection_read.c:86-170)
 *   -> martini_parms (src/bioMartini.c:1210-1353): mmff_init (src/bioMMFF.c:236-273),
 *      genMartiniConn (:567-838), genMartiniBondPair (:135-282), martiniLJ_parms (:868-950)
 *   -> restraint_parms (src/restraint.c:200-257) -> neighbor_init (src/neighbor.c:34-56)
 *   -> nglf_parms -> ddc_init (src/ddc.c:40-120)
 * and flattens the per-residue templates into per-bead term lists, which is what
 * charmmResidues + connectiveEnergy rebuild every step on the CPU (src/bioCharmmCovalent.c:48-93).
 */
#include "host.h"
#include "../../../include/ddcmd_b200_host.h"
#include <ctype.h>
#include <float.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <strings.h>
#include <time.h>

static char g_hostErr[1024];
const char *ddcb200_lastHostError(void) { return g_hostErr; }
int herr(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_hostErr, sizeof g_hostErr, fmt, ap);
    va_end(ap);
    return -1;
}

double ddcb200_unitsConvert(double value, const char *from, const char *to) { return hu_convert(value, from, to); }

/* ---- MMFF tree (reference src/bioMMFF.h:4-218) ----------------------------------------------- */
typedef struct { int atomI, atomJ, func; char typeI[32], typeJ[32]; double kb, b0; } H_BOND;
typedef struct { int atomI, atomJ, atomK, func; double ktheta, theta0; } H_ANGLE;
typedef struct { int atomI, atomJ, atomK, atomL, func, n; double kchi, delta; } H_TORS;
typedef struct { int atomI, atomJ, valid, list; double r0; } H_PAIR;   /* list, r0: constraints only */
typedef struct { char name[32], type[32]; int atomID, typeID; double charge; } H_ATOM;
typedef struct
{
    char objName[64], resName[32];
    int resID, nAtoms, nBonds, nAngles, nTors, nExcl, nCons, nConsLists;
    H_ATOM *atoms;
    H_BOND *bonds;
    H_ANGLE *angles;
    H_TORS *tors;
    H_PAIR *excl, *cons;
    int nBpair;
    int *bpI, *bpJ;
} H_RESI;

typedef struct
{
    int nResi, nTypes;
    H_RESI *resi;
    int *typeID;           /* atomTypes[i]->atomTypeID */
    char (*typeName)[32];
} H_MMFF;

#define NEED(o, what, name) do { if (!(o)) return herr("object %s %s not found", name, what); } while (0)

static int loadResi(ODB *db, const char *name, H_RESI *r)
{
    const ODB_OBJECT *o = odb_find(db, name, "RESIPARMS");
    NEED(o, "RESIPARMS", name);
    memset(r, 0, sizeof *r);
    snprintf(r->objName, sizeof r->objName, "%s", name);
    char *s;
    odb_getString(o, "resName", &s, "NoName");
    snprintf(r->resName, sizeof r->resName, "%s", s);
    free(s);
    odb_getInts(o, "resID", &r->resID, 1, "0");
    char **groups;
    int ng = odb_getStrings(o, "groupList", &groups, NULL);
    if (ng <= 0) return herr("RESIPARMS %s has no groupList", name);
    int cap = 0;
    for (int g = 0; g < ng; g++)
    {
        const ODB_OBJECT *go = odb_find(db, groups[g], "GROUPPARMS");
        NEED(go, "GROUPPARMS", groups[g]);
        char **an;
        int na = odb_getStrings(go, "atomList", &an, NULL);
        for (int a = 0; a < na; a++)
        {
            const ODB_OBJECT *ao = odb_find(db, an[a], "ATOMPARMS");
            NEED(ao, "ATOMPARMS", an[a]);
            if (r->nAtoms == cap)
            {
                cap = cap ? 2 * cap : 16;
                r->atoms = (H_ATOM *)realloc(r->atoms, cap * sizeof(H_ATOM));
            }
            H_ATOM *at = &r->atoms[r->nAtoms++];
            memset(at, 0, sizeof *at);
            odb_getString(ao, "atomName", &s, "NoName");
            snprintf(at->name, sizeof at->name, "%s", s);
            free(s);
            odb_getString(ao, "atomType", &s, "NoType");
            snprintf(at->type, sizeof at->type, "%s", s);
            free(s);
            odb_getInts(ao, "atomID", &at->atomID, 1, "0");
            odb_getInts(ao, "atomTypeID", &at->typeID, 1, "0");
            if (odb_getWithUnits(ao, "charge", &at->charge, 1, "0.0", "i*t", NULL) < 0) return herr("bad charge unit in %s", an[a]);
        }
        odb_freeStrings(an, na);
    }
    odb_freeStrings(groups, ng);

    char **names;
    int n = odb_getStrings(o, "bondList", &names, NULL);
    r->nBonds = n;
    r->bonds = (H_BOND *)calloc(n > 0 ? n : 1, sizeof(H_BOND));
    for (int i = 0; i < n; i++)
    {
        const ODB_OBJECT *b = odb_find(db, names[i], "BONDPARMS");
        NEED(b, "BONDPARMS", names[i]);
        H_BOND *h = &r->bonds[i];
        odb_getInts(b, "atomI", &h->atomI, 1, "0");
        odb_getInts(b, "atomJ", &h->atomJ, 1, "0");
        odb_getInts(b, "func", &h->func, 1, "1");
        odb_getString(b, "atomTypeI", &s, "NoType"); snprintf(h->typeI, 32, "%s", s); free(s);
        odb_getString(b, "atomTypeJ", &s, "NoType"); snprintf(h->typeJ, 32, "%s", s); free(s);
        if (odb_getWithUnits(b, "kb", &h->kb, 1, "0.0", "kJ*mol^-1*nm^-2", NULL) < 0) return herr("bad kb unit in %s", names[i]);
        if (odb_getWithUnits(b, "b0", &h->b0, 1, "0.0", "nm", NULL) < 0) return herr("bad b0 unit in %s", names[i]);
    }
    odb_freeStrings(names, n);

    n = odb_getStrings(o, "angleList", &names, NULL);
    r->nAngles = n;
    r->angles = (H_ANGLE *)calloc(n > 0 ? n : 1, sizeof(H_ANGLE));
    for (int i = 0; i < n; i++)
    {
        const ODB_OBJECT *b = odb_find(db, names[i], "ANGLEPARMS");
        NEED(b, "ANGLEPARMS", names[i]);
        H_ANGLE *h = &r->angles[i];
        odb_getInts(b, "atomI", &h->atomI, 1, "0");
        odb_getInts(b, "atomJ", &h->atomJ, 1, "0");
        odb_getInts(b, "atomK", &h->atomK, 1, "0");
        odb_getInts(b, "func", &h->func, 1, "1");
        if (odb_getWithUnits(b, "ktheta", &h->ktheta, 1, "0.0", "kJ*mol^-1", NULL) < 0) return herr("bad ktheta unit in %s", names[i]);
        odb_getDoubles(b, "theta0", &h->theta0, 1, "0");
    }
    odb_freeStrings(names, n);

    n = odb_getStrings(o, "dihedralList", &names, NULL);
    r->nTors = n;
    r->tors = (H_TORS *)calloc(n > 0 ? n : 1, sizeof(H_TORS));
    for (int i = 0; i < n; i++)
    {
        const ODB_OBJECT *b = odb_find(db, names[i], "TORSPARMS");
        NEED(b, "TORSPARMS", names[i]);
        H_TORS *h = &r->tors[i];
        odb_getInts(b, "atomI", &h->atomI, 1, "0");
        odb_getInts(b, "atomJ", &h->atomJ, 1, "0");
        odb_getInts(b, "atomK", &h->atomK, 1, "0");
        odb_getInts(b, "atomL", &h->atomL, 1, "0");
        odb_getInts(b, "func", &h->func, 1, "1");
        odb_getInts(b, "n", &h->n, 1, "1");
        if (odb_getWithUnits(b, "kchi", &h->kchi, 1, "0.0", "kJ*mol^-1", NULL) < 0) return herr("bad kchi unit in %s", names[i]);
        odb_getDoubles(b, "delta", &h->delta, 1, "0");
    }
    odb_freeStrings(names, n);

    n = odb_getStrings(o, "exclusionList", &names, NULL);
    r->nExcl = n;
    r->excl = (H_PAIR *)calloc(n > 0 ? n : 1, sizeof(H_PAIR));
    for (int i = 0; i < n; i++)
    {
        const ODB_OBJECT *b = odb_find(db, names[i], "EXCLUDEPARMS");
        NEED(b, "EXCLUDEPARMS", names[i]);
        odb_getInts(b, "atomI", &r->excl[i].atomI, 1, "0");
        odb_getInts(b, "atomJ", &r->excl[i].atomJ, 1, "0");
        r->excl[i].valid = 1;
    }
    odb_freeStrings(names, n);

    /* constraintList -> CONSLISTPARMS{constraintSubList} -> CONSPARMS (valid iff func==1, src/bioMMFF.c:66-82) */
    n = odb_getStrings(o, "constraintList", &names, NULL);
    int ccap = 0;
    for (int i = 0; i < n; i++)
    {
        const ODB_OBJECT *cl = odb_find(db, names[i], "CONSLISTPARMS");
        NEED(cl, "CONSLISTPARMS", names[i]);
        char **sub;
        int ns = odb_getStrings(cl, "constraintSubList", &sub, NULL);
        for (int k = 0; k < ns; k++)
        {
            const ODB_OBJECT *b = odb_find(db, sub[k], "CONSPARMS");
            NEED(b, "CONSPARMS", sub[k]);
            if (r->nCons == ccap)
            {
                ccap = ccap ? 2 * ccap : 16;
                r->cons = (H_PAIR *)realloc(r->cons, ccap * sizeof(H_PAIR));
            }
            H_PAIR *h = &r->cons[r->nCons++];
            int func;
            odb_getInts(b, "atomI", &h->atomI, 1, "0");
            odb_getInts(b, "atomJ", &h->atomJ, 1, "0");
            odb_getInts(b, "func", &func, 1, "1");
            h->valid = (func == 1);
            h->list = i;
            if (odb_getWithUnits(b, "r0", &h->r0, 1, "0.0", "nm", NULL) < 0) return herr("bad r0 unit in CONSPARMS %s", sub[k]);
        }
        odb_freeStrings(sub, ns);
    }
    r->nConsLists = n > 0 ? n : 0;
    odb_freeStrings(names, n);
    return 0;
}

static int samePair(int a, int b, int c, int d) { return (a == c && b == d) || (a == d && b == c); }

/* validateExclusions + genMartiniBondPair (src/bioMartini.c:54-282) */
static void buildBpairs(H_RESI *r)
{
    for (int e = 0; e < r->nExcl; e++)
        for (int i = 0; i < r->nBonds; i++)
            if (r->bonds[i].func == 1 && samePair(r->excl[e].atomI, r->excl[e].atomJ, r->bonds[i].atomI, r->bonds[i].atomJ))
            {
                r->excl[e].valid = 0;
                break;
            }
    for (int c = 0; c < r->nCons; c++)
    {
        if (!r->cons[c].valid) continue;
        for (int i = 0; i < r->nBonds; i++)
            if (r->bonds[i].func == 1 && samePair(r->cons[c].atomI, r->cons[c].atomJ, r->bonds[i].atomI, r->bonds[i].atomJ))
            {
                r->cons[c].valid = 0;
                break;
            }
        for (int e = 0; e < r->nExcl; e++)
            if (r->excl[e].valid && samePair(r->cons[c].atomI, r->cons[c].atomJ, r->excl[e].atomI, r->excl[e].atomJ))
            {
                r->cons[c].valid = 0;
                break;
            }
    }
    int n = 0;
    for (int i = 0; i < r->nBonds; i++) n += r->bonds[i].func == 1;
    for (int e = 0; e < r->nExcl; e++) n += r->excl[e].valid;
    for (int c = 0; c < r->nCons; c++) n += r->cons[c].valid;
    r->bpI = (int *)calloc(n > 0 ? n : 1, sizeof(int));
    r->bpJ = (int *)calloc(n > 0 ? n : 1, sizeof(int));
    r->nBpair = 0;
    for (int i = 0; i < r->nBonds; i++)
        if (r->bonds[i].func == 1) { r->bpI[r->nBpair] = r->bonds[i].atomI; r->bpJ[r->nBpair++] = r->bonds[i].atomJ; }
    for (int e = 0; e < r->nExcl; e++)
        if (r->excl[e].valid) { r->bpI[r->nBpair] = r->excl[e].atomI; r->bpJ[r->nBpair++] = r->excl[e].atomJ; }
    for (int c = 0; c < r->nCons; c++)
        if (r->cons[c].valid) { r->bpI[r->nBpair] = r->cons[c].atomI; r->bpJ[r->nBpair++] = r->cons[c].atomJ; }
}

static void freeMMFF(H_MMFF *m)
{
    for (int i = 0; i < m->nResi; i++)
    {
        H_RESI *r = &m->resi[i];
        free(r->atoms); free(r->bonds); free(r->angles); free(r->tors); free(r->excl); free(r->cons); free(r->bpI); free(r->bpJ);
    }
    free(m->resi); free(m->typeID); free(m->typeName);
}

/* CRC-32 (reflected 0x04c11db7, as checksum_crc32_table src/crc32.c:70-82) */
uint32_t hcrc32(const unsigned char *data, size_t len)
{
    static uint32_t lut[256];
    static int inited = 0;
    if (!inited)
    {
        for (uint32_t n = 0; n < 256; n++)
        {
            uint32_t c = n;
            for (int k = 0; k < 8; k++) c = (c & 1) ? 0xedb88320u ^ (c >> 1) : c >> 1;
            lut[n] = c;
        }
        inited = 1;
    }
    uint32_t crc = 0xffffffffu;
    while (len--) crc = (crc >> 8) ^ lut[(crc & 0xff) ^ *data++];
    return crc ^ 0xffffffffu;
}

/* ---- atoms file (VARRECORDASCII, src/collection_read.c:86-170) ---------------------------------- */
typedef struct { const char *name; int index; } NameIdx;
static int cmpName(const void *a, const void *b) { return strcmp(((const NameIdx *)a)->name, ((const NameIdx *)b)->name); }

/* FIXRECORDBINARY records (collection_readBINARY, src/collection_read.c:201-345; written by collection_writeBLOCK_binary,
 * src/collection_write.c:188-336): checksum u4 | id bN (big-endian) | pinfo bM | rx ry rz f8 | vx vy vz f8 or f4 | LCG64 16 B.
 * pinfo = group + species * nGroups over the header's `groups` / `species` lists (pinfoDecode, src/pinfo.c:129-146). */
static uint64_t beField(const unsigned char *b, int n)
{
    uint64_t r = 0;
    for (int i = 0; i < n; i++) r = r * 256 + b[i];
    return r;
}
static int readAtomsBinary(const char *path, const unsigned char *data, size_t ndata, const ODB_OBJECT *h, int64_t size, ddcb200_deck *d,
                           NameIdx *sortedSpecies, int nspecies, int64_t *filled, int *randomMissing)
{
    int lrec = 0, key = 0, rfs = 0;
    odb_getInts(h, "lrec", &lrec, 1, "0");
    odb_getInts(h, "endian_key", &key, 1, "0");
    odb_getInts(h, "randomFieldSize", &rfs, 1, "0");
    if (key != 875770417) return herr("atoms file %s: byte-swapped binary records are not supported (endian_key %d)", path, key);
    char **ft, **fn, **gn, **sn, **tn, *ck = NULL, *rnd = NULL;
    const int nt = odb_getStrings(h, "field_types", &ft, NULL), nf = odb_getStrings(h, "field_names", &fn, NULL);
    const int ng = odb_getStrings(h, "groups", &gn, NULL), ns = odb_getStrings(h, "species", &sn, NULL), nty = odb_getStrings(h, "types", &tn, NULL);
    odb_getString(h, "checksum", &ck, "NONE");
    odb_getString(h, "random", &rnd, "NotSet");
    int rc = 0;
    int off[16], len[16], role[16], fixed = 0;   /* role: 0 checksum 1 id 2 pinfo 3..8 rx..vz, -1 ignored */
    static const char *names[9] = {"checksum", "id", "pinfo", "rx", "ry", "rz", "vx", "vy", "vz"};
    int *spOf = NULL, *grOf = NULL;
    if (nt != nf || nf < 1 || nf > 16 || lrec <= 0 || ng < 1 || ns < 1 || nty != 1) { rc = herr("atoms file %s: malformed binary header", path); goto out; }
    for (int k = 0; k < nf; k++)
    {
        len[k] = atoi(ft[k] + 1);
        off[k] = fixed;
        fixed += len[k];
        role[k] = -1;
        for (int r = 0; r < 9; r++)
            if (strcmp(fn[k], names[r]) == 0) role[k] = r;
        if (len[k] <= 0 || len[k] > 8 || (role[k] >= 3 && ft[k][0] != 'f') || (role[k] >= 3 && len[k] != 8 && len[k] != 4))
        { rc = herr("atoms file %s: unsupported binary field %s %s", path, fn[k], ft[k]); goto out; }
    }
    const int haveRnd = strcmp(rnd, "NotSet") != 0 && strcmp(rnd, "NONE") != 0 && rfs == 16;
    if (!haveRnd) *randomMissing = 1;
    if (fixed + (haveRnd ? 16 : 0) > lrec) { rc = herr("atoms file %s: fields exceed lrec", path); goto out; }
    spOf = (int *)malloc(sizeof(int) * (size_t)ns);
    grOf = (int *)malloc(sizeof(int) * (size_t)ng);
    for (int k = 0; k < ns; k++)
    {
        NameIdx q = {sn[k], 0};
        NameIdx *hit = (NameIdx *)bsearch(&q, sortedSpecies, nspecies, sizeof(NameIdx), cmpName);
        spOf[k] = hit ? hit->index : -1;
    }
    for (int k = 0; k < ng; k++)
    {
        grOf[k] = -1;
        for (int g = 0; g < d->nGroups; g++)
            if (strcmp(d->groupName[g], gn[k]) == 0) grOf[k] = g;
    }
    const double lc = hu_convert(1.0, "l", NULL), tc = hu_convert(1.0, "t", NULL), vc = lc / tc;
    const int useCrc = strcmp(ck, "CRC32") == 0;
    int64_t i = *filled;
    const int64_t nrec = (int64_t)(ndata / (size_t)lrec);
    for (int64_t r = 0; r < nrec && i < size; r++, i++)
    {
        const unsigned char *b = data + (size_t)r * (size_t)lrec;
        double v[6] = {0, 0, 0, 0, 0, 0};
        for (int k = 0; k < nf; k++)
        {
            const unsigned char *q = b + off[k];
            if (role[k] == 0)
            {
                uint32_t want;
                memcpy(&want, q, 4);
                if (useCrc && len[k] == 4 && hcrc32(b + 4, (size_t)lrec - 4) != want) { rc = herr("atoms file %s: CRC32 mismatch in record %lld", path, (long long)i); goto out; }
            }
            else if (role[k] == 1) d->gid[i] = beField(q, len[k]);
            else if (role[k] == 2)
            {
                const uint64_t pin = beField(q, len[k]);
                const int ig = (int)(pin % (uint64_t)ng), is = (int)((pin / (uint64_t)ng) % (uint64_t)ns);
                if (pin >= (uint64_t)ng * (uint64_t)ns || spOf[is] < 0) { rc = herr("atoms file %s: record %lld names an unknown species", path, (long long)i); goto out; }
                d->species[i] = spOf[is];
                if (d->groupOfBead)
                {
                    if (grOf[ig] < 0) { rc = herr("atoms file %s: record %lld names GROUP %s, which SYSTEM groups does not list", path, (long long)i, gn[ig]); goto out; }
                    d->groupOfBead[i] = (unsigned char)grOf[ig];
                }
            }
            else if (role[k] >= 3)
            {
                if (len[k] == 8) memcpy(&v[role[k] - 3], q, 8);
                else { float f4; memcpy(&f4, q, 4); v[role[k] - 3] = f4; }
            }
        }
        d->rx[i] = lc * v[0]; d->ry[i] = lc * v[1]; d->rz[i] = lc * v[2];
        d->vx[i] = vc * v[3]; d->vy[i] = vc * v[4]; d->vz[i] = vc * v[5];
        if (haveRnd && d->rngState)
        {
            /* lcg64_bread (src/lcg64.c:65-74): state u8, multID u4, prime u4, native byte order */
            uint64_t st; uint32_t m, pr;
            memcpy(&st, b + fixed, 8); memcpy(&m, b + fixed + 8, 4); memcpy(&pr, b + fixed + 12, 4);
            if (pr == 0 || m > 2) *randomMissing = 1;
            else { d->rngState[i] = st; d->rngMult[i] = m; d->rngPrime[i] = pr; }
        }
    }
    *filled = i;
out:
    free(spOf); free(grOf); free(ck); free(rnd);
    odb_freeStrings(ft, nt); odb_freeStrings(fn, nf); odb_freeStrings(gn, ng); odb_freeStrings(sn, ns); odb_freeStrings(tn, nty);
    return rc;
}

static int readAtoms(const char *path, int64_t size, ddcb200_deck *d, NameIdx *sortedSpecies, int nspecies, int64_t *filled, int *randomMissing)
{
    FILE *f = fopen(path, "rb");
    if (!f) return herr("cannot open atoms file %s", path);
    fseek(f, 0, SEEK_END);
    long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    char *text = (char *)malloc((size_t)sz + 1);
    size_t got = fread(text, 1, (size_t)sz, f);
    fclose(f);
    text[got] = 0;
    int crcField = 0, gidBase = 10, lrec = 0;
    char *p = strchr(text, '}');
    if (!p) { free(text); return herr("atoms file %s has no FILEHEADER", path); }
    /* header checks: ASCII variable records with the standard field list */
    {
        ODB *hdb = odb_new();
        char save = p[1];
        p[1] = 0;
        odb_compileString(hdb, text);
        p[1] = save;
        if (hdb->n > 0)
        {
            char *dt = NULL;
            odb_getString(hdb->obj[0], "datatype", &dt, "VARRECORDASCII");
            if (strcmp(dt, "FIXRECORDBINARY") == 0)
            {
                /* the records start after the first blank line that follows the header (readPheader, src/pio.c:751-781) */
                free(dt);
                const char *q = p + 1;
                const char *end = text + got;
                while (q + 1 < end && !(q[0] == '\n' && q[1] == '\n')) q++;
                int rcb = q + 2 <= end ? readAtomsBinary(path, (const unsigned char *)q + 2, (size_t)(end - (q + 2)), hdb->obj[0], size, d, sortedSpecies,
                                                         nspecies, filled, randomMissing)
                                       : herr("atoms file %s: no records after the header", path);
                odb_free(hdb);
                free(text);
                return rcb;
            }
            int bad = strcmp(dt, "VARRECORDASCII") != 0 && strcmp(dt, "FIXRECORDASCII") != 0;
            free(dt);
            char **fn;
            int nf = odb_getStrings(hdb->obj[0], "field_names", &fn, "id class type group rx ry rz vx vy vz");
            const char *want[10] = {"id", "class", "type", "group", "rx", "ry", "rz", "vx", "vy", "vz"};
            /* restart files written by collection_writeBLOCK carry a CRC32 per record as the first field
             * (src/collection_write.c:104-108; checkRecord skips it, src/check_line.c:70-109) */
            crcField = nf > 0 && strcmp(fn[0], "checksum") == 0;
            if (nf < 10 + crcField) bad = 1;
            for (int i = 0; i < 10 && i + crcField < nf; i++) bad |= strcmp(fn[i + crcField], want[i]) != 0;
            odb_freeStrings(fn, nf);
            {
                char *ck = NULL;
                odb_getString(hdb->obj[0], "checksum", &ck, "NONE");
                if ((strcmp(ck, "CRC32") == 0) != crcField) bad = 1;
                free(ck);
                /* the id is hexadecimal when its field_format ends in 'x' (src/collection_read.c:113-127) */
                char **ff;
                int nff = odb_getStrings(hdb->obj[0], "field_format", &ff, NULL);
                if (nff > 1 && ff[1][0] && ff[1][strlen(ff[1]) - 1] == 'x') gidBase = 16;
                odb_freeStrings(ff, nff);
                odb_getInts(hdb->obj[0], "lrec", &lrec, 1, "0");
            }
            {
                char *rnd = NULL;
                odb_getString(hdb->obj[0], "random", &rnd, "NotSet");
                if (strcmp(rnd, "NONE") == 0) *randomMissing = 1;   /* hasRandomField == No (src/collection_read.c:103-109) */
                free(rnd);
            }
            if (bad) { odb_free(hdb); free(text); return herr("atoms file %s: only ASCII records 'id class type group rx ry rz vx vy vz' are supported", path); }
        }
        odb_free(hdb);
    }
    p++;
    const double lc = hu_convert(1.0, "l", NULL), tc = hu_convert(1.0, "t", NULL);
    const double vc = lc / tc;   /* src/collection_read.c:94-96 */
    int64_t i = *filled;
    while (i < size)
    {
        while (*p && isspace((unsigned char)*p)) p++;
        if (!*p) break;
        char *end;
        if (crcField)
        {
            /* 8 hex digits + ' ' ; the CRC covers the rest of the fixed-length record, newline included */
            unsigned long want = strtoul(p, &end, 16);
            if (end != p + 8) { free(text); return herr("atoms file %s: bad checksum field in record %lld", path, (long long)i); }
            if (lrec > 8 && (size_t)(p - text) + (size_t)lrec <= got && hcrc32((const unsigned char *)p + 8, (size_t)lrec - 8) != (uint32_t)want)
            {
                free(text);
                return herr("atoms file %s: CRC32 mismatch in record %lld", path, (long long)i);
            }
            p = end;
        }
        uint64_t gid = strtoull(p, &end, gidBase);
        if (end == p) { free(text); return herr("atoms file %s: bad record %lld", path, (long long)i); }
        p = end;
        /* class, species ("type") and group names: manual tokens (sscanf would strlen the whole file per record) */
        char *tok[3];
        for (int k = 0; k < 3; k++)
        {
            while (*p == ' ' || *p == '\t') p++;
            tok[k] = p;
            while (*p && !isspace((unsigned char)*p)) p++;
            if (p == tok[k]) { free(text); return herr("atoms file %s: short record %lld", path, (long long)i); }
            if (*p) *p++ = 0;
        }
        const char *type = tok[1];
        if (d->groupOfBead)
        {
            int gi = -1;
            for (int g = 0; g < d->nGroups; g++)
                if (strcmp(d->groupName[g], tok[2]) == 0) { gi = g; break; }
            if (gi < 0) { herr("atoms file %s: record %lld names GROUP %s, which SYSTEM groups does not list", path, (long long)i, tok[2]); free(text); return -1; }
            d->groupOfBead[i] = (unsigned char)gi;
        }
        NameIdx key = {type, 0};
        NameIdx *hit = (NameIdx *)bsearch(&key, sortedSpecies, nspecies, sizeof(NameIdx), cmpName);
        if (!hit) { free(text); return herr("atoms file %s: unknown species %s", path, type); }
        double v[6];
        for (int k = 0; k < 6; k++)
        {
            v[k] = strtod(p, &end);
            if (end == p) { free(text); return herr("atoms file %s: bad number in record %lld", path, (long long)i); }
            p = end;
        }
        if (d->rngState && !*randomMissing)
        {
            /* lcg64_parse (src/lcg64.c:87-94): "%llx %u %x"; one unparsable record makes every bead use the default */
            unsigned long long st;
            unsigned mult, prime;
            int used = 0;
            char *eol = p;
            while (*eol && *eol != '\n') eol++;
            const char save = *eol;
            *eol = 0;
            const int cnt = sscanf(p, "%llx %u %x%n", &st, &mult, &prime, &used);
            *eol = save;
            if (cnt != 3 || prime == 0 || mult > 2) *randomMissing = 1;
            else { d->rngState[i] = st; d->rngMult[i] = mult; d->rngPrime[i] = prime; }
        }
        while (*p && *p != '\n') p++;   /* ignore the remaining fields (group data) */
        d->gid[i] = gid;
        d->species[i] = hit->index;
        d->rx[i] = lc * v[0]; d->ry[i] = lc * v[1]; d->rz[i] = lc * v[2];
        d->vx[i] = vc * v[3]; d->vy[i] = vc * v[4]; d->vz[i] = vc * v[5];
        i++;
    }
    *filled = i;
    free(text);
    return 0;
}

/* ---- helpers ------------------------------------------------------------------------------------ */
static char *dirOf(const char *path)
{
    const char *s = strrchr(path, '/');
    if (!s) return strdup(".");
    size_t n = (size_t)(s - path);
    char *d = (char *)malloc(n + 1);
    memcpy(d, path, n);
    d[n] = 0;
    return d;
}
static char *joinPath(const char *dir, const char *file)
{
    if (file[0] == '/') return strdup(file);
    size_t n = strlen(dir) + strlen(file) + 2;
    char *p = (char *)malloc(n);
    snprintf(p, n, "%s/%s", dir, file);
    return p;
}
static double ljShift(double sigma, double eps, double rcut)
{
    /* CGLennardJones_setShift, src/bioMartini.c:840-848 */
    double sr = sigma / rcut, s2 = sr * sr, s4 = s2 * s2, s6 = s4 * s2, s12 = s6 * s6;
    return (-4.0 * eps * (s12 - s6));
}

typedef struct { uint64_t gid; int64_t idx; } GidIdx;
static int cmpGid(const void *a, const void *b)
{
    uint64_t x = ((const GidIdx *)a)->gid, y = ((const GidIdx *)b)->gid;
    return x < y ? -1 : (x > y ? 1 : 0);
}

/* ---- per-bead LCG64 default streams ----------------------------------------------------------- */
static uint64_t mulmod(uint64_t a, uint64_t b, uint64_t m) { return (uint64_t)(((__uint128_t)a * b) % m); }
static uint64_t powmod(uint64_t a, uint64_t e, uint64_t m)
{
    uint64_t r = 1;
    a %= m;
    while (e)
    {
        if (e & 1) r = mulmod(r, a, m);
        a = mulmod(a, a, m);
        e >>= 1;
    }
    return r;
}
/* primality as isPrime1 decides it (src/primes.c:128-160): Miller-Rabin with the bases 2..17, exact below 3.4e14 */
static int isPrime64(uint64_t n)
{
    static const uint64_t bases[7] = {2, 3, 5, 7, 11, 13, 17};
    if (n < 2) return 0;
    for (int i = 0; i < 7; i++)
    {
        if (n == bases[i]) return 1;
        if (n % bases[i] == 0) return 0;
    }
    uint64_t s = n - 1;
    int r = 0;
    while ((s & 1) == 0) { s >>= 1; r++; }
    for (int i = 0; i < 7; i++)
    {
        uint64_t x = powmod(bases[i], s, n);
        if (x == 1 || x == n - 1) continue;
        int comp = 1;
        for (int k = 1; k < r && comp; k++)
        {
            x = mulmod(x, x, n);
            if (x == n - 1) comp = 0;
        }
        if (comp) return 0;
    }
    return 1;
}
/* lcg64_default called for bead 0..n-1 in file order with seed = gid (src/collection.c:102-108, src/lcg64.c:96-109):
 * state = INIT_SEED ^ gid, multID cycles 0,1,2, a new prime every third bead.  nextPrime (src/primes.c:33-66) on one
 * task walks the odd numbers upwards from (2^31 + 1) - 30000 (prime_init(30000, 0, 1), src/ddcMD.c:70). */
static void lcg64Defaults(ddcb200_deck *d)
{
    const uint64_t INIT_SEED = 0x2bc6ffff8cfe166dull;
    uint64_t cand = ((2ull << 30) + 1ull) - 30000ull;
    if ((cand & 1) == 0) cand++;
    int first = 1;
    uint64_t prime = 0;
    unsigned mult = 0;
    for (int64_t i = 0; i < d->n; i++)
    {
        if (mult == 0)
        {
            if (!first) cand += 2;
            first = 0;
            while (!isPrime64(cand)) cand += 2;
            prime = cand;
        }
        d->rngState[i] = INIT_SEED ^ d->gid[i];
        d->rngMult[i] = mult;
        d->rngPrime[i] = (uint32_t)prime;
        mult = (mult + 1) % 3;
    }
}

void ddcb200_deckFree(ddcb200_deck *d)
{
    if (!d) return;
    for (int i = 0; i < d->nspecies; i++) free(d->speciesName[i]);
    free(d->speciesName); free(d->specLJ); free(d->specCharge); free(d->specMass); free(d->specMolType);
    free(d->specResidue); free(d->specAtom); free(d->ljEps); free(d->ljSigma); free(d->ljShift);
    free(d->molTypeNSpecies); free(d->molTypeResidue); free(d->molTypeOwnerOffset); free(d->bpairOffset); free(d->bpairI); free(d->bpairJ);
    free(d->gid); free(d->species); free(d->rx); free(d->ry); free(d->rz); free(d->vx); free(d->vy); free(d->vz);
    free(d->termKind); free(d->termIdx); free(d->termParm);
    free(d->restrBead); free(d->restrFrac0); free(d->restrKb); free(d->restrFc);
    free(d->molOffset); free(d->molBeads);
    for (int i = 0; i < d->nGroups; i++) free(d->groupName[i]);
    free(d->groupName); free(d->groupType); free(d->groupTeq); free(d->groupTau); free(d->groupVcm); free(d->groupOfBead);
    free(d->rngState); free(d->rngMult); free(d->rngPrime);
    free(d->consAtomOffset); free(d->consPairOffset); free(d->consAtomBead); free(d->consPairA); free(d->consPairB); free(d->consPairDist);
    for (int a = 0; a < d->nSubsets; a++)
    {
        ddcb200_subset *q = &d->subsets[a];
        free(q->name); free(q->filename); free(q->lengthUnit); free(q->parmsInfo); free(q->idList); free(q->includeSpecies);
    }
    free(d->subsets);
    for (int a = 0; a < d->nPairCorr; a++) { free(d->pairCorr[a].name); free(d->pairCorr[a].filename); free(d->pairCorr[a].miscInfo); }
    free(d->pairCorr);
    free(d->runDir); free(d->simulateName); free(d->boxName); free(d->collectionName); free(d->atomsdir);
    if (d->speciesType)
        for (int i = 0; i < d->nspecies; i++) free(d->speciesType[i]);
    free(d->speciesType);
    for (int i = 0; i < 6; i++) free(d->printUnit[i]);
    free(d);
}

#define FAIL(...) do { rc = herr(__VA_ARGS__); goto done; } while (0)

int ddcb200_deckLoad(const char *objectFile, const char *restartFile, const char *simulateName, ddcb200_deck **out)
{
    int rc = 0;
    hu_init();
    if (!objectFile || !out) return herr("null argument");
    ODB *db = odb_new();
    ddcb200_deck *d = (ddcb200_deck *)calloc(1, sizeof(ddcb200_deck));
    H_MMFF mm;
    memset(&mm, 0, sizeof mm);
    char *dir = dirOf(objectFile);
    NameIdx *sorted = NULL;
    GidIdx *order = NULL;
    char **molNames = NULL;
    int nMolNames = 0;
    /* everything `done:` frees is initialised before the first FAIL */
    char *sysName = NULL, *intName = NULL, *ddcName = NULL, *piName = NULL, *s = NULL;
    char *boxName = NULL, *nbrName = NULL, *colName = NULL, *mcName = NULL;

    if (odb_compileFile(db, objectFile)) FAIL("%s", db->err);
    {
        char *rpath = restartFile ? strdup(restartFile) : joinPath(dir, "restart");
        FILE *t = fopen(rpath, "rb");
        if (t)
        {
            fclose(t);
            if (odb_compileFile(db, rpath)) { free(rpath); FAIL("%s", db->err); }
        }
        else if (restartFile) { free(rpath); FAIL("cannot open restart file %s", restartFile); }
        free(rpath);
    }
    const ODB_OBJECT *sim = odb_find(db, simulateName ? simulateName : "simulate", "SIMULATE");
    if (!sim) FAIL("SIMULATE object %s not found", simulateName ? simulateName : "simulate");
    odb_getString(sim, "system", &sysName, NULL);
    odb_getString(sim, "integrator", &intName, NULL);
    odb_getString(sim, "ddc", &ddcName, "ddc");
    odb_getString(sim, "printinfo", &piName, "printinfo");
    if (!sysName || !intName) FAIL("SIMULATE needs system and integrator keywords");
    d->runDir = strdup(dir);
    d->simulateName = strdup(simulateName ? simulateName : "simulate");
    {
        /* snapshotRootDir -> atomsdir (atomsdirParse, src/simulate.c:38-57: first word, trailing '/' dropped), gidFormat,
         * nLoopDigits, run_id (src/simulate.c:159-183,209) */
        char *root = NULL, *gf = NULL;
        odb_getString(sim, "snapshotRootDir", &root, "./");
        char *sp = root;
        while (*sp && !isspace((unsigned char)*sp)) sp++;
        *sp = 0;
        size_t rl = strlen(root);
        if (rl > 0 && root[rl - 1] == '/') root[rl - 1] = 0;
        d->atomsdir = root;
        odb_getString(sim, "gidFormat", &gf, "decimal");
        d->gidFormatHex = strcasecmp(gf, "hex") == 0;
        free(gf);
        char *cm = NULL;
        odb_getString(sim, "checkpointmode", &cm, "ASCII");
        d->checkpointBinary = strcasecmp(cm, "BINARY") == 0;
        free(cm);
        cm = NULL;
        odb_getString(sim, "checkpointprecision", &cm, "FULL");
        d->checkpointBrief = strcasecmp(cm, "BRIEF") == 0;
        free(cm);
        odb_getInts(sim, "nLoopDigits", &d->nLoopDigits, 1, "8");
        if (d->nLoopDigits < 1 || d->nLoopDigits > 18) FAIL("SIMULATE nLoopDigits out of range");
        int64_t rid = 0;
        odb_getI64(sim, "run_id", &rid, "0");
        d->runId = rid ? (unsigned)rid : (unsigned)time(NULL);
    }
    odb_getI64(sim, "loop", &d->loop, "0");
    odb_getI64(sim, "maxloop", &d->maxloop, "0");
    odb_getInts(sim, "printrate", &d->printrate, 1, "5");
    odb_getInts(sim, "deltaloop", &d->deltaloop, 1, "-1");
    odb_getInts(sim, "snapshotrate", &d->snapshotrate, 1, "100");
    odb_getInts(sim, "checkpointrate", &d->checkpointrate, 1, "1000");
    if (odb_getWithUnits(sim, "time", &d->time, 1, "0.0", "t", NULL) < 0) FAIL("bad time unit");
    if (odb_getWithUnits(sim, "dt", &d->dt, 1, "1.0", "t", NULL) < 0) FAIL("bad dt unit");

    /* INTEGRATOR: NGLF or NGLFCONSTRAINT (src/integrator.c:59-83) */
    const ODB_OBJECT *integ = odb_find(db, intName, "INTEGRATOR");
    if (!integ) FAIL("INTEGRATOR %s not found", intName);
    odb_getString(integ, "type", &s, "");
    if (strcmp(s, "NGLF") == 0) d->integratorType = 0;
    else if (strcmp(s, "NGLFCONSTRAINT") == 0)
    {
        /* nglfconstraint_parms (src/nglfconstraint.c:86-95) */
        d->integratorType = 1;
        if (odb_getWithUnits(integ, "T", &d->ncT, 1, "310", "T", NULL) < 0 || odb_getWithUnits(integ, "P0", &d->ncP0, 1, "0.0", "pressure", NULL) < 0 ||
            odb_getWithUnits(integ, "beta", &d->ncBeta, 1, "0.0", "1/pressure", NULL) < 0 ||
            odb_getWithUnits(integ, "tauBarostat", &d->ncTauBarostat, 1, "0.0", "t", NULL) < 0)
        { free(s); s = NULL; FAIL("bad unit in INTEGRATOR %s", intName); }
        odb_getInts(integ, "isotropic", &d->ncIsotropic, 1, "0");
    }
    else { char t[64]; snprintf(t, 64, "%s", s); free(s); s = NULL; FAIL("INTEGRATOR type %s is not supported (NGLF, NGLFCONSTRAINT)", t); }
    free(s);
    s = NULL;

    const ODB_OBJECT *sys = odb_find(db, sysName, "SYSTEM");
    if (!sys) FAIL("SYSTEM %s not found", sysName);
    {
        /* GROUP objects: FREE (src/free.c) and LANGEVIN with a numeric tau = NORMAL Langevin, constant Teq (src/langevin.c:63-91,130-170) */
        char **gn;
        int ng = odb_getStrings(sys, "groups", &gn, NULL);
        if (ng > 8) { odb_freeStrings(gn, ng); FAIL("more than 8 GROUP objects are not supported"); }
        d->nGroups = ng > 0 ? ng : 0;
        d->groupName = (char **)calloc(ng > 0 ? ng : 1, sizeof(char *));
        d->groupType = (int *)calloc(ng > 0 ? ng : 1, sizeof(int));
        d->groupTeq = (double *)calloc(ng > 0 ? ng : 1, sizeof(double));
        d->groupTau = (double *)calloc(ng > 0 ? ng : 1, sizeof(double));
        d->groupVcm = (double *)calloc(ng > 0 ? 3 * ng : 1, sizeof(double));
        for (int g = 0; g < ng; g++) d->groupName[g] = strdup(gn[g]);
        for (int g = 0; g < ng; g++)
        {
            const ODB_OBJECT *go = odb_find(db, gn[g], "GROUP");
            if (!go) { odb_freeStrings(gn, ng); FAIL("GROUP %s not found", d->groupName[g]); }
            odb_getString(go, "type", &s, "");
            int ok = 1;
            if (strcmp(s, "FREE") == 0) d->groupType[g] = 0;
            else if (strcmp(s, "LANGEVIN") == 0)
            {
                d->groupType[g] = 1;
                char *tv = NULL, *end = NULL, *dyn = NULL;
                odb_getString(go, "tau", &tv, "NotDefined");
                strtod(tv, &end);
                const int numericTau = end != tv;
                free(tv);
                odb_getString(go, "Teq_dynamics", &dyn, "EXPLICIT_TIME");
                const int explicitTime = strcmp(dyn, "EXPLICIT_TIME") == 0;
                free(dyn);
                if (!numericTau || !explicitTime) { free(s); s = NULL; odb_freeStrings(gn, ng); FAIL("LANGEVIN group %s: only the NORMAL thermostat (numeric tau, Teq_dynamics=EXPLICIT_TIME) is supported", d->groupName[g]); }
                char *tq = NULL;
                odb_getString(go, "Teq", &tq, NULL);
                if (!tq) { free(s); s = NULL; odb_freeStrings(gn, ng); FAIL("LANGEVIN group %s needs Teq", d->groupName[g]); }
                /* Teq is an eq_parse expression of time in the reference; a constant "<number><unit>" is what Martini decks use */
                strtod(tq, &end);
                int constant = end != tq;
                for (const char *q = end; constant && *q; q++)
                    if (!(isalnum((unsigned char)*q) || *q == '_' || *q == ' ')) constant = 0;
                free(tq);
                if (!constant || odb_getWithUnits(go, "Teq", &d->groupTeq[g], 1, "0.0", "T", NULL) < 0)
                { free(s); s = NULL; odb_freeStrings(gn, ng); FAIL("LANGEVIN group %s: Teq must be a constant temperature", d->groupName[g]); }
                if (odb_getWithUnits(go, "tau", &d->groupTau[g], 1, "1.0", "t", NULL) < 0 ||
                    odb_getWithUnits(go, "vcm", &d->groupVcm[3 * g], 3, "0.0 0.0 0.0", "l/t", NULL) < 0)
                { free(s); s = NULL; odb_freeStrings(gn, ng); FAIL("bad unit in GROUP %s", d->groupName[g]); }
                if (!(d->groupTau[g] > 0.0)) { free(s); s = NULL; odb_freeStrings(gn, ng); FAIL("LANGEVIN group %s: tau must be positive", d->groupName[g]); }
            }
            else ok = 0;
            free(s);
            s = NULL;
            if (!ok) { odb_freeStrings(gn, ng); FAIL("GROUP %s: type other than FREE or LANGEVIN is not supported", d->groupName[g]); }
        }
        odb_freeStrings(gn, ng);
    }
    {
        /* RANDOM (src/random.c:47-72): only LCG64; randomizeSeed only matters for generators made after start-up, the per-bead
         * default streams are seeded by the gid (src/collection.c:96-110) */
        char *rname = NULL;
        odb_getString(sys, "random", &rname, "NotSet");
        const ODB_OBJECT *ro = odb_find(db, rname, "RANDOM");
        if (ro)
        {
            char *rt = NULL;
            odb_getString(ro, "type", &rt, "");
            const int lcg = strcmp(rt, "LCG64") == 0;
            free(rt);
            if (!lcg) { free(rname); FAIL("RANDOM type other than LCG64 is not supported"); }
            int64_t seed = 0;
            odb_getI64(ro, "seed", &seed, "0");
            d->randomSeed = (uint64_t)seed;
            d->haveRandom = 1;
        }
        free(rname);
        for (int g = 0; g < d->nGroups; g++)
            if (d->groupType[g] == 1 && !d->haveRandom) FAIL("LANGEVIN group %s needs a RANDOM object in SYSTEM (missingRandomError)", d->groupName[g]);
    }
    odb_getInts(sys, "nConstraints", &d->params.nConstraints, 1, "0");

    /* BOX */
    odb_getString(sys, "box", &boxName, NULL);
    odb_getString(sys, "neighbor", &nbrName, NULL);
    odb_getString(sys, "collection", &colName, NULL);
    odb_getString(sys, "moleculeClass", &mcName, "NONE");
    if (!boxName || !nbrName || !colName) FAIL("SYSTEM needs box, neighbor and collection");
    const ODB_OBJECT *box = odb_find(db, boxName, "BOX");
    if (!box) FAIL("BOX %s not found", boxName);
    d->boxName = strdup(boxName);
    d->collectionName = strdup(colName);
    odb_getDoubles(box, "reducedcorner", d->reducedCorner, 3, "-0.5 -0.5 -0.5");
    if (odb_getWithUnits(box, "h", d->params.h, 9, "1 0 0 0 1 0 0 0 1", "l", NULL) < 0) FAIL("bad box unit");
    odb_getInts(box, odb_has(box, "bndcdn") ? "bndcdn" : "pbc", &d->params.pbc, 1, "7");
    const ODB_OBJECT *nbr = odb_find(db, nbrName, "NEIGHBOR");
    if (!nbr) FAIL("NEIGHBOR %s not found", nbrName);
    if (odb_getWithUnits(nbr, "deltaR", &d->params.deltaR, 1, "0", "l", NULL) < 0) FAIL("bad deltaR unit");
    if (odb_getWithUnits(nbr, "minBoxSide", &d->params.minBoxSide, 1, "0", "l", NULL) < 0) FAIL("bad minBoxSide unit");
    const ODB_OBJECT *ddc = odb_find(db, ddcName, "DDC");
    if (ddc)
    {
        odb_getInts(ddc, "updateRate", &d->params.updateRate, 1, "0");
        odb_getInts(ddc, "lx", &d->ddc_lx, 1, "0");
        odb_getInts(ddc, "ly", &d->ddc_ly, 1, "0");
        odb_getInts(ddc, "lz", &d->ddc_lz, 1, "0");
    }
    if (d->params.updateRate < 0) FAIL("DDC updateRate must be >= 0");
    const ODB_OBJECT *pi = odb_find(db, piName, "PRINTINFO");
    if (pi) odb_getInts(pi, "printMolecularPressure", &d->printMolecularPressure, 1, "0");
    {
        /* PRINTINFO units (src/printinfo.c:35-36,60-77): value strings and factors from internal units */
        static const char *key[6] = {"LENGTH", "TIME", "TEMPERATURE", "ENERGY", "PRESSURE", "VOLUME"};
        static const char *dflt[6] = {"Ang", "fs", "K", "eV", "GPa", "Bohr^3"};
        for (int k = 0; k < 6; k++)
        {
            if (pi) odb_getString(pi, key[k], &d->printUnit[k], dflt[k]);
            else d->printUnit[k] = strdup(dflt[k]);
            d->printConvert[k] = hu_convert(1.0, NULL, d->printUnit[k]);
            if (!(d->printConvert[k] == d->printConvert[k]) || d->printConvert[k] == 0.0) FAIL("PRINTINFO %s = %s is not a unit", key[k], d->printUnit[k]);
        }
    }

    /* MOLECULECLASS -> MOLECULE -> SPECIES */
    if (strcmp(mcName, "NONE") == 0) FAIL("SYSTEM without moleculeClass is not supported (Martini decks define one)");
    const ODB_OBJECT *mc = odb_find(db, mcName, "MOLECULECLASS");
    if (!mc) FAIL("MOLECULECLASS %s not found", mcName);
    nMolNames = odb_getStrings(mc, "molecules", &molNames, NULL);
    if (nMolNames <= 0) FAIL("MOLECULECLASS has no molecules");
    d->nMolTypes = nMolNames;
    d->molTypeNSpecies = (int *)calloc(nMolNames, sizeof(int));
    d->molTypeResidue = (int *)calloc(nMolNames, sizeof(int));
    d->molTypeOwnerOffset = (int *)calloc(nMolNames, sizeof(int));
    int spCap = 0;
    for (int m = 0; m < nMolNames; m++)
    {
        const ODB_OBJECT *mo = odb_find(db, molNames[m], "MOLECULE");
        if (!mo) FAIL("MOLECULE %s not found", molNames[m]);
        char **sn, *owner = NULL;
        int ns = odb_getStrings(mo, "species", &sn, NULL);
        if (ns <= 0) FAIL("MOLECULE %s has no species", molNames[m]);
        odb_getString(mo, "ownershipSpecies", &owner, "$NONE$");
        d->molTypeNSpecies[m] = ns;
        d->molTypeOwnerOffset[m] = 0;
        for (int k = 0; k < ns; k++)
        {
            const ODB_OBJECT *so = odb_find(db, sn[k], "SPECIES");
            if (!so) FAIL("SPECIES %s not found", sn[k]);
            if (d->nspecies == spCap)
            {
                spCap = spCap ? 2 * spCap : 64;
                d->speciesName = (char **)realloc(d->speciesName, spCap * sizeof(char *));
                d->specLJ = (int *)realloc(d->specLJ, spCap * sizeof(int));
                d->specCharge = (double *)realloc(d->specCharge, spCap * sizeof(double));
                d->specMass = (double *)realloc(d->specMass, spCap * sizeof(double));
                d->specMolType = (int *)realloc(d->specMolType, spCap * sizeof(int));
                d->specResidue = (int *)realloc(d->specResidue, spCap * sizeof(int));
                d->specAtom = (int *)realloc(d->specAtom, spCap * sizeof(int));
                d->speciesType = (char **)realloc(d->speciesType, spCap * sizeof(char *));
            }
            const int idx = d->nspecies++;
            d->speciesName[idx] = strdup(sn[k]);
            d->speciesType[idx] = NULL;
            odb_getString(so, "type", &d->speciesType[idx], "ATOM");
            d->specMolType[idx] = m;
            if (odb_getWithUnits(so, "mass", &d->specMass[idx], 1, "1.0", "m", NULL) < 0) FAIL("bad mass unit in SPECIES %s", sn[k]);
            if (odb_getWithUnits(so, "charge", &d->specCharge[idx], 1, "0.0", "i*t", NULL) < 0) FAIL("bad charge unit in SPECIES %s", sn[k]);
            if (strcmp(sn[k], owner) == 0) d->molTypeOwnerOffset[m] = k;
        }
        free(owner);
        odb_freeStrings(sn, ns);
    }

    /* POTENTIALs */
    {
        char **pn;
        int np = odb_getStrings(sys, "potential", &pn, NULL);
        int haveMartini = 0;
        for (int k = 0; k < np; k++)
        {
            const ODB_OBJECT *po = odb_find(db, pn[k], "POTENTIAL");
            if (!po) { odb_freeStrings(pn, np); FAIL("POTENTIAL %s not found", pn[k]); }
            char *type = NULL, *parmfile = NULL;
            odb_getString(po, "type", &type, "");
            if (strcmp(type, "MARTINI") == 0)
            {
                haveMartini = 1;
                odb_getString(po, "parmfile", &parmfile, "martini.data");
                char *pp = joinPath(dir, parmfile);
                int e = odb_compileFile(db, pp);
                free(pp);
                if (e) { free(type); free(parmfile); odb_freeStrings(pn, np); FAIL("%s", db->err); }
                odb_getInts(po, "excludePotentialTerm", &d->excludePotentialTerm, 1, "0");
                odb_getInts(po, "potential-shift", &d->potentialShift, 1, "1");
                double cutoff;
                if (odb_getWithUnits(po, "cutoff", &cutoff, 1, "11.0", "Angstrom", NULL) < 0 ||
                    odb_getWithUnits(po, "rmax4all", &d->rmax4all, 1, "11.0", "Angstrom", NULL) < 0 ||
                    odb_getWithUnits(po, "rcoulomb", &d->rcoulomb, 1, "11.0", "Angstrom", NULL) < 0)
                { free(type); free(parmfile); odb_freeStrings(pn, np); FAIL("bad unit in POTENTIAL %s", pn[k]); }
                odb_getDoubles(po, "epsilon_r", &d->epsilon_r, 1, "15.0");
                odb_getDoubles(po, "epsilon_rf", &d->epsilon_rf, 1, "-1.0");
                d->params.rmax = cutoff;
                /* src/bioMartini.c:1234-1245 */
                const double irc = 1.0 / d->rcoulomb, irc3 = irc * irc * irc;
                if (d->epsilon_rf != -1.0)
                {
                    d->params.krf = (d->epsilon_rf - d->epsilon_r) / (2 * d->epsilon_rf + d->epsilon_r) * irc3;
                    d->params.crf = 3 * (d->epsilon_rf) / (2 * d->epsilon_rf + d->epsilon_r) * irc;
                }
                else
                {
                    d->params.krf = 0.5 * irc3;
                    d->params.crf = 1.5 * irc;
                }
                if (d->params.rmax > d->rmax4all) d->rmax4all = d->params.rmax;
                if (d->rcoulomb > d->rmax4all) d->rmax4all = d->rcoulomb;
                d->params.keR = hu_ke() / d->epsilon_r;
            }
            else if (strcmp(type, "RESTRAINT") == 0)
            {
                odb_getString(po, "parmfile", &parmfile, "restraint.data");
                char *pp = joinPath(dir, parmfile);
                int e = odb_compileFile(db, pp);
                free(pp);
                if (e) { free(type); free(parmfile); odb_freeStrings(pn, np); FAIL("%s", db->err); }
            }
            else
            {
                char t[64];
                snprintf(t, 64, "%s", type);
                free(type); free(parmfile); odb_freeStrings(pn, np);
                FAIL("POTENTIAL type %s is not supported (MARTINI, RESTRAINT)", t);
            }
            free(type);
            free(parmfile);
        }
        odb_freeStrings(pn, np);
        if (!haveMartini) FAIL("no POTENTIAL of type MARTINI in SYSTEM");
    }

    /* MMFF (mmff_init, src/bioMMFF.c:236-273) */
    {
        const ODB_OBJECT *mo = odb_find(db, "martini", "MMFF");
        if (!mo) FAIL("MMFF object 'martini' not found in the parmfile");
        char **rn, **tn, **ln;
        int nr = odb_getStrings(mo, "resiParms", &rn, NULL);
        int nt = odb_getStrings(mo, "atomTypeList", &tn, NULL);
        int nl = odb_getStrings(mo, "ljParms", &ln, NULL);
        if (nr <= 0 || nt <= 0) FAIL("MMFF needs resiParms and atomTypeList");
        mm.nResi = nr;
        mm.resi = (H_RESI *)calloc(nr, sizeof(H_RESI));
        for (int i = 0; i < nr; i++)
        {
            if (loadResi(db, rn[i], &mm.resi[i])) { rc = -1; goto done; }
            buildBpairs(&mm.resi[i]);
        }
        mm.nTypes = nt;
        mm.typeID = (int *)calloc(nt, sizeof(int));
        mm.typeName = (char(*)[32])calloc(nt, 32);
        for (int i = 0; i < nt; i++)
        {
            const ODB_OBJECT *to = odb_find(db, tn[i], "MASSPARMS");
            if (!to) FAIL("MASSPARMS %s not found", tn[i]);
            odb_getInts(to, "atomTypeID", &mm.typeID[i], 1, "0");
            odb_getString(to, "atomType", &s, "NoType");
            snprintf(mm.typeName[i], 32, "%s", s);
            free(s);
            if (mm.typeID[i] < 0 || mm.typeID[i] >= nt) FAIL("MASSPARMS %s: atomTypeID out of range", tn[i]);
        }
        d->ntypes = nt;
        d->ljEps = (double *)calloc((size_t)nt * nt, sizeof(double));
        d->ljSigma = (double *)calloc((size_t)nt * nt, sizeof(double));
        d->ljShift = (double *)calloc((size_t)nt * nt, sizeof(double));
        for (int k = 0; k < nt * nt; k++) d->ljSigma[k] = 1.0;
        for (int i = 0; i < nl; i++)
        {
            const ODB_OBJECT *lo = odb_find(db, ln[i], "LJPARMS");
            if (!lo) FAIL("LJPARMS %s not found", ln[i]);
            int a, b;
            double sigma, eps;
            odb_getInts(lo, "indexI", &a, 1, "0");
            odb_getInts(lo, "indexJ", &b, 1, "0");
            if (odb_getWithUnits(lo, "sigma", &sigma, 1, "1.0", "nm", NULL) < 0 || odb_getWithUnits(lo, "eps", &eps, 1, "0.0", "kJ*mol^-1", NULL) < 0)
                FAIL("bad unit in LJPARMS %s", ln[i]);
            if (a < 0 || a >= nt || b < 0 || b >= nt) FAIL("LJPARMS %s: index out of range", ln[i]);
            const double sh = d->potentialShift ? ljShift(sigma, eps, d->params.rmax) : 0.0;
            d->ljEps[a + b * nt] = d->ljEps[b + a * nt] = eps;
            d->ljSigma[a + b * nt] = d->ljSigma[b + a * nt] = sigma;
            d->ljShift[a + b * nt] = d->ljShift[b + a * nt] = sh;
        }
        odb_freeStrings(rn, nr);
        odb_freeStrings(tn, nt);
        odb_freeStrings(ln, nl);
    }

    /* species -> residue / atom / LJ type (src/bioMartini.c:1274-1308, :952-987) */
    for (int sp = 0; sp < d->nspecies; sp++)
    {
        const char *nm = d->speciesName[sp];
        const char *x = strchr(nm, 'x');
        if (!x) FAIL("species name %s is not of the form <RES>x<ATOM>", nm);
        char res[64];
        snprintf(res, sizeof res, "%.*s", (int)(x - nm), nm);
        int ri = -1, ai = -1;
        for (int r = 0; r < mm.nResi && ri < 0; r++)
            if (strcmp(mm.resi[r].resName, res) == 0)
                for (int a = 0; a < mm.resi[r].nAtoms; a++)
                    if (strcmp(mm.resi[r].atoms[a].name, x + 1) == 0) { ri = r; ai = a; break; }
        if (ri < 0) FAIL("species %s: no residue/atom in the MMFF", nm);
        d->specResidue[sp] = ri;
        d->specAtom[sp] = ai;
        const int tid = mm.resi[ri].atoms[ai].typeID;
        if (tid < 0 || tid >= mm.nTypes) FAIL("species %s: atomTypeID out of range", nm);
        d->specLJ[sp] = mm.typeID[tid];   /* massParms[atomTypeID]->atmTypeID, src/bioMartini.c:634,980 */
    }
    /* molecule type -> residue of its ownership species (reOrgPairs, src/bioMartini.c:1416-1423) */
    {
        int base = 0, total = 0;
        d->bpairOffset = (int *)calloc(d->nMolTypes + 1, sizeof(int));
        for (int m = 0; m < d->nMolTypes; m++)
        {
            d->molTypeResidue[m] = d->specResidue[base + d->molTypeOwnerOffset[m]];
            total += mm.resi[d->molTypeResidue[m]].nBpair;
            base += d->molTypeNSpecies[m];
        }
        d->bpairI = (int *)calloc(total > 0 ? total : 1, sizeof(int));
        d->bpairJ = (int *)calloc(total > 0 ? total : 1, sizeof(int));
        int k = 0;
        for (int m = 0; m < d->nMolTypes; m++)
        {
            d->bpairOffset[m] = k;
            const H_RESI *r = &mm.resi[d->molTypeResidue[m]];
            for (int b = 0; b < r->nBpair; b++) { d->bpairI[k] = r->bpI[b]; d->bpairJ[k++] = r->bpJ[b]; }
        }
        d->bpairOffset[d->nMolTypes] = k;
    }

    /* COLLECTION */
    {
        const ODB_OBJECT *co = odb_find(db, colName, "COLLECTION");
        if (!co) FAIL("COLLECTION %s not found", colName);
        int64_t size = 0;
        odb_getI64(co, "size", &size, "0");
        char *files = NULL;
        odb_getString(co, "files", &files, NULL);
        if (size <= 0 || !files) FAIL("COLLECTION needs size and files");
        d->n = size;
        d->gid = (uint64_t *)calloc(size, sizeof(uint64_t));
        d->species = (int *)calloc(size, sizeof(int));
        d->rx = (double *)calloc(size, sizeof(double)); d->ry = (double *)calloc(size, sizeof(double)); d->rz = (double *)calloc(size, sizeof(double));
        d->vx = (double *)calloc(size, sizeof(double)); d->vy = (double *)calloc(size, sizeof(double)); d->vz = (double *)calloc(size, sizeof(double));
        int randomMissing = 0;
        if (d->nGroups > 0) d->groupOfBead = (unsigned char *)calloc(size, 1);
        if (d->haveRandom)
        {
            d->rngState = (uint64_t *)calloc(size, sizeof(uint64_t));
            d->rngMult = (uint32_t *)calloc(size, sizeof(uint32_t));
            d->rngPrime = (uint32_t *)calloc(size, sizeof(uint32_t));
        }
        sorted = (NameIdx *)calloc(d->nspecies, sizeof(NameIdx));
        for (int i = 0; i < d->nspecies; i++) { sorted[i].name = d->speciesName[i]; sorted[i].index = i; }
        qsort(sorted, d->nspecies, sizeof(NameIdx), cmpName);
        int64_t filled = 0;
        for (int fi = 0; fi < 4096 && filled < size; fi++)
        {
            char fn[1024];
            snprintf(fn, sizeof fn, "%s%06d", files, fi);
            char *pp = joinPath(dir, fn);
            FILE *t = fopen(pp, "rb");
            if (!t) { free(pp); break; }
            fclose(t);
            int e = readAtoms(pp, size, d, sorted, d->nspecies, &filled, &randomMissing);
            free(pp);
            if (e) { free(files); rc = -1; goto done; }
        }
        free(files);
        if (filled != size) FAIL("COLLECTION size=%lld but %lld records were read", (long long)size, (long long)filled);
        if (d->haveRandom && randomMissing) lcg64Defaults(d);
    }

    /* flatten bonded terms: walk residues in gid order (charmmResidues, src/bioCharmmCovalent.c:48-93) */
    {
        const int64_t n = d->n;
        order = (GidIdx *)malloc((size_t)n * sizeof(GidIdx));
        for (int64_t i = 0; i < n; i++) { order[i].gid = d->gid[i]; order[i].idx = i; }
        qsort(order, (size_t)n, sizeof(GidIdx), cmpGid);
        const uint64_t molResMask = 0xffffffffffff0000ull;
        const int ex = d->excludePotentialTerm;
        for (int pass = 0; pass < 2; pass++)
        {
            int64_t nt = 0, nm = 0, nmb = 0, nmt = 0, nc = 0, nca = 0, ncp = 0;
            for (int64_t a = 0; a < n;)
            {
                int64_t b = a;
                while (b < n && (order[b].gid & molResMask) == (order[a].gid & molResMask)) b++;
                const int sp0 = d->species[order[a].idx];
                const int ri = d->specResidue[sp0];
                const H_RESI *r = &mm.resi[ri];
                if (b - a != r->nAtoms) FAIL("residue instance at gid %llu has %lld beads, template %s has %d (incomplete residues are not supported)",
                                             (unsigned long long)order[a].gid, (long long)(b - a), r->resName, r->nAtoms);
                for (int64_t k = a; k < b; k++)
                {
                    const int sp = d->species[order[k].idx];
                    if (d->specResidue[sp] != ri || d->specAtom[sp] != (int)(k - a))
                        FAIL("bead gid %llu: species %s is not atom %d of residue %s", (unsigned long long)order[k].gid, d->speciesName[sp], (int)(k - a), r->resName);
                    if ((int)(order[k].gid & 0xffffull) != (int)(k - a))
                        FAIL("bead gid %llu: low 16 bits must equal the atom offset %d in residue %s", (unsigned long long)order[k].gid, (int)(k - a), r->resName);
                }
#define BEAD(off) ((int)order[a + (off)].idx)
#define CHECKOFF(off) do { if ((off) < 0 || (off) >= r->nAtoms) FAIL("term atom offset %d outside residue %s", (off), r->resName); } while (0)
                if (!(ex & 1))
                    for (int t = 0; t < r->nBonds; t++)
                    {
                        /* every bond in bondList is evaluated (resBondSorted loops the whole sorted list); Martini constraints
                         * (func != 1 bonds are still harmonic here, as in genMartiniConn which copies all of bondList) */
                        CHECKOFF(r->bonds[t].atomI); CHECKOFF(r->bonds[t].atomJ);
                        if (pass)
                        {
                            /* sortBondList swaps so that atmI < atmJ (src/bioCharmmParms.c:2147-2190) */
                            int ia = r->bonds[t].atomI, ja = r->bonds[t].atomJ;
                            if (ia > ja) { int tmp = ia; ia = ja; ja = tmp; }
                            d->termKind[nt] = 0;
                            d->termIdx[4 * nt] = BEAD(ia); d->termIdx[4 * nt + 1] = BEAD(ja); d->termIdx[4 * nt + 2] = -1; d->termIdx[4 * nt + 3] = -1;
                            d->termParm[3 * nt] = r->bonds[t].kb; d->termParm[3 * nt + 1] = r->bonds[t].b0; d->termParm[3 * nt + 2] = 0;
                        }
                        nt++;
                    }
                for (int t = 0; t < r->nAngles; t++)
                {
                    const H_ANGLE *h = &r->angles[t];
                    int kind = h->func == 1 ? 1 : (h->func == 2 ? 2 : (h->func == 10 ? 3 : -1));
                    if (kind < 0) continue;   /* genMartiniConn keeps only func 1, 2, 10 (src/bioMartini.c:662-754) */
                    if ((kind == 1 && (ex & 2)) || (kind == 2 && (ex & 4)) || (kind == 3 && (ex & 256))) continue;
                    CHECKOFF(h->atomI); CHECKOFF(h->atomJ); CHECKOFF(h->atomK);
                    if (pass)
                    {
                        d->termKind[nt] = kind;
                        d->termIdx[4 * nt] = BEAD(h->atomI); d->termIdx[4 * nt + 1] = BEAD(h->atomJ); d->termIdx[4 * nt + 2] = BEAD(h->atomK); d->termIdx[4 * nt + 3] = -1;
                        d->termParm[3 * nt] = h->ktheta; d->termParm[3 * nt + 1] = h->theta0; d->termParm[3 * nt + 2] = 0;
                    }
                    nt++;
                }
                for (int t = 0; t < r->nTors; t++)
                {
                    const H_TORS *h = &r->tors[t];
                    int kind = h->func == 1 ? 4 : (h->func == 2 ? 5 : -1);
                    if (kind < 0) continue;
                    if ((kind == 4 && (ex & 16)) || (kind == 5 && (ex & 32))) continue;
                    CHECKOFF(h->atomI); CHECKOFF(h->atomJ); CHECKOFF(h->atomK); CHECKOFF(h->atomL);
                    if (pass)
                    {
                        d->termKind[nt] = kind;
                        d->termIdx[4 * nt] = BEAD(h->atomI); d->termIdx[4 * nt + 1] = BEAD(h->atomJ); d->termIdx[4 * nt + 2] = BEAD(h->atomK); d->termIdx[4 * nt + 3] = BEAD(h->atomL);
                        d->termParm[3 * nt] = h->kchi; d->termParm[3 * nt + 1] = h->delta; d->termParm[3 * nt + 2] = (double)h->n;
                    }
                    nt++;
                }
                /* constraint clusters (genConstraint, src/bioMartini.c:445-565): one per CONSLISTPARMS, in the order of
                 * the lowest atom that carries each list; an atom named by two lists belongs to the later one */
                if (r->nCons > 0)
                {
                    int consIndex[r->nAtoms], done[r->nConsLists];
                    for (int k = 0; k < r->nAtoms; k++) consIndex[k] = -1;
                    for (int k = 0; k < r->nConsLists; k++) done[k] = 0;
                    for (int q = 0; q < r->nCons; q++)
                    {
                        CHECKOFF(r->cons[q].atomI); CHECKOFF(r->cons[q].atomJ);
                        consIndex[r->cons[q].atomI] = r->cons[q].list;
                        consIndex[r->cons[q].atomJ] = r->cons[q].list;
                    }
                    for (int k = 0; k < r->nAtoms; k++)
                    {
                        const int cl = consIndex[k];
                        if (cl < 0 || done[cl]) continue;
                        done[cl] = 1;
                        int local[r->nAtoms], na = 0;
                        for (int a2 = 0; a2 < r->nAtoms; a2++)
                        {
                            local[a2] = -1;
                            if (consIndex[a2] == cl)
                            {
                                if (pass) d->consAtomBead[nca + na] = BEAD(a2);
                                local[a2] = na++;
                            }
                        }
                        int np = 0;
                        for (int q = 0; q < r->nCons; q++)
                        {
                            if (r->cons[q].list != cl) continue;
                            const int la = local[r->cons[q].atomI], lb = local[r->cons[q].atomJ];
                            if (la < 0 || lb < 0) FAIL("residue %s: constraint lists overlap (findIndexInConsGroup would abort)", r->resName);
                            if (pass)
                            {
                                d->consPairA[ncp + np] = la;
                                d->consPairB[ncp + np] = lb;
                                d->consPairDist[ncp + np] = r->cons[q].r0;
                            }
                            np++;
                        }
                        if (pass) { d->consAtomOffset[nc] = nca; d->consPairOffset[nc] = ncp; }
                        nc++;
                        nca += na;
                        ncp += np;
                    }
                }
                /* molecule bookkeeping (moleculeScanState, src/molecule.c:118-211): one molecule per (gid>>32) */
                {
                    const int mt = d->specMolType[sp0];
                    const int ownerOff = d->molTypeOwnerOffset[mt];
                    nmt++;   /* every residue instance holds its molecule's ownership bead in a Martini deck */
                    if (d->molTypeNSpecies[mt] > 1)
                    {
                        if (pass)
                        {
                            d->molOffset[nm] = nmb;
                            d->molBeads[nmb] = BEAD(ownerOff);
                            int64_t w = nmb + 1;
                            for (int k = 0; k < r->nAtoms; k++)
                                if (k != ownerOff) d->molBeads[w++] = BEAD(k);
                        }
                        nm++;
                        nmb += r->nAtoms;
                    }
                }
                a = b;
            }
            if (!pass)
            {
                d->nTerms = nt;
                d->termKind = (int *)calloc(nt > 0 ? nt : 1, sizeof(int));
                d->termIdx = (int *)calloc(nt > 0 ? 4 * nt : 1, sizeof(int));
                d->termParm = (double *)calloc(nt > 0 ? 3 * nt : 1, sizeof(double));
                d->nMol = nm;
                d->nMolTotal = nmt;
                d->molOffset = (int64_t *)calloc(nm + 1, sizeof(int64_t));
                d->molBeads = (int *)calloc(nmb > 0 ? nmb : 1, sizeof(int));
                d->nCons = nc;
                d->consAtomOffset = (int64_t *)calloc(nc + 1, sizeof(int64_t));
                d->consPairOffset = (int64_t *)calloc(nc + 1, sizeof(int64_t));
                d->consAtomBead = (int *)calloc(nca > 0 ? nca : 1, sizeof(int));
                d->consPairA = (int *)calloc(ncp > 0 ? ncp : 1, sizeof(int));
                d->consPairB = (int *)calloc(ncp > 0 ? ncp : 1, sizeof(int));
                d->consPairDist = (double *)calloc(ncp > 0 ? ncp : 1, sizeof(double));
            }
            else
            {
                d->molOffset[nm] = nmb;
                d->consAtomOffset[nc] = nca;
                d->consPairOffset[nc] = ncp;
            }
        }
    }

    /* RESTRAINTLIST (src/restraint.c:25-80,164-198) */
    {
        const ODB_OBJECT *ro = odb_find(db, "restraint", "RESTRAINTLIST");
        if (ro)
        {
            odb_getInts(ro, "origin", &d->restrOrigin, 1, "0");
            char **rn;
            int nr = odb_getStrings(ro, "restraintList", &rn, NULL);
            d->restrBead = (int *)calloc(nr > 0 ? nr : 1, sizeof(int));
            d->restrFrac0 = (double *)calloc(nr > 0 ? 3 * nr : 1, sizeof(double));
            d->restrKb = (double *)calloc(nr > 0 ? nr : 1, sizeof(double));
            d->restrFc = (double *)calloc(nr > 0 ? 3 * nr : 1, sizeof(double));
            int64_t k = 0;
            for (int i = 0; i < nr; i++)
            {
                const ODB_OBJECT *po = odb_find(db, rn[i], "RESTRAINTPARMS");
                if (!po) { odb_freeStrings(rn, nr); FAIL("RESTRAINTPARMS %s not found", rn[i]); }
                int64_t gid = 0;
                odb_getI64(po, "gid", &gid, "0");
                GidIdx key = {(uint64_t)gid, 0};
                GidIdx *hit = (GidIdx *)bsearch(&key, order, (size_t)d->n, sizeof(GidIdx), cmpGid);
                if (!hit) continue;   /* restraintMap stays -1: no bead with that gid */
                d->restrBead[k] = (int)hit->idx;
                int fc[3];
                odb_getInts(po, "fcx", &fc[0], 1, "0");
                odb_getInts(po, "fcy", &fc[1], 1, "0");
                odb_getInts(po, "fcz", &fc[2], 1, "0");
                odb_getDoubles(po, "x0", &d->restrFrac0[3 * k], 1, "0");
                odb_getDoubles(po, "y0", &d->restrFrac0[3 * k + 1], 1, "0");
                odb_getDoubles(po, "z0", &d->restrFrac0[3 * k + 2], 1, "0");
                for (int a = 0; a < 3; a++) d->restrFc[3 * k + a] = fc[a];
                if (odb_getWithUnits(po, "kb", &d->restrKb[k], 1, "0.0", "kJ*mol^-1*nm^-2", NULL) < 0) { odb_freeStrings(rn, nr); FAIL("bad kb unit in %s", rn[i]); }
                k++;
            }
            d->nRestraints = k;
            odb_freeStrings(rn, nr);
        }
    }

    d->kB = hu_kB();
    d->ke = hu_ke();
    /* ANALYSIS objects named by SIMULATE analysis (src/simulate.c:264-271, src/analysis.c:142-176): subsetWrite only */
    {
        char **an;
        const int na = odb_getStrings(sim, "analysis", &an, NULL);
        if (na > 0) d->subsets = (ddcb200_subset *)calloc((size_t)na, sizeof(ddcb200_subset));
        for (int a = 0; a < na; a++)
        {
            const ODB_OBJECT *ao = odb_find(db, an[a], "ANALYSIS");
            if (!ao) { herr("ANALYSIS %s not found", an[a]); odb_freeStrings(an, na); rc = -1; goto done; }
            char *type = NULL, *fmt = NULL;
            odb_getString(ao, "type", &type, "NONE");
            odb_getString(ao, "format", &fmt, "pio");
            const int isSubset = strcasecmp(type, "subsetWrite") == 0 || strcasecmp(type, "subset_write") == 0;
            if (strncasecmp(type, "PAIRCORRELATION", 15) == 0)
            {
                /* paircorrelation_parms, src/paircorrelation.c:71-145 (the three pair-finding methods give the same counts) */
                free(type); free(fmt);
                if (!d->pairCorr) d->pairCorr = (ddcb200_paircorr *)calloc((size_t)na, sizeof(ddcb200_paircorr));
                ddcb200_paircorr *g = &d->pairCorr[d->nPairCorr++];
                g->name = strdup(an[a]);
                odb_getInts(ao, "eval_rate", &g->evalRate, 1, "0");
                odb_getInts(ao, "outputrate", &g->outputRate, 1, "0");
                odb_getString(ao, "filename", &g->filename, "paircorrelation.dat");
                odb_getInts(ao, "length", &g->nBins, 1, "1");
                char *rs = NULL;
                odb_getString(ao, "rscale", &rs, "normal");
                g->logScale = strcasecmp(rs, "log") == 0;
                const int okScale = g->logScale || strcasecmp(rs, "normal") == 0;
                free(rs);
                if (odb_getWithUnits(ao, "delta_r", &g->deltaR, 1, "1", "l", NULL) < 0 || odb_getWithUnits(ao, "rmin", &g->rmin, 1, "0", "l", NULL) < 0 ||
                    !okScale || g->nBins < 1 || !(g->deltaR > 0.0) || (g->logScale && !(g->rmin > 0.0)))
                {
                    herr("ANALYSIS %s: bad PAIRCORRELATION parameters", an[a]);
                    odb_freeStrings(an, na);
                    rc = -1;
                    goto done;
                }
                g->rmax = g->rmin + g->nBins * g->deltaR;
                if (g->logScale) g->logDelta = (log10(g->rmax) - log10(g->rmin)) / (g->nBins * 1.0);
                char info[256];
                snprintf(info, sizeof info, "rmin = %f Ang; delta_r = %f Ang; length = %d; eval_rate = %d; outputrate = %d;",
                         hu_convert(g->rmin, NULL, "Angstrom"), hu_convert(g->deltaR, NULL, "Angstrom"), g->nBins, g->evalRate, g->outputRate);
                g->miscInfo = strdup(info);
                continue;
            }
            if (!isSubset || strcmp(fmt, "binaryCharmm") != 0)
            {
                herr("ANALYSIS %s: only type = PAIRCORRELATION, or subsetWrite with format = binaryCharmm, is supported (type %s, format %s)", an[a], type, fmt);
                free(type); free(fmt); odb_freeStrings(an, na);
                rc = -1;
                goto done;
            }
            free(type); free(fmt);
            ddcb200_subset *q = &d->subsets[d->nSubsets++];
            q->name = strdup(an[a]);
            odb_getInts(ao, "eval_rate", &q->evalRate, 1, "0");
            odb_getInts(ao, "outputrate", &q->outputRate, 1, "0");
            odb_getInts(ao, "modulus", &q->modulus, 1, "1");
            odb_getInts(ao, "odd", &q->odd, 1, "0");
            odb_getString(ao, "filename", &q->filename, "subset");
            odb_getString(ao, "lengthUnit", &q->lengthUnit, "Ang");
            char **t;
            int nt = odb_getStrings(ao, "idmin", &t, "0");
            q->idMin = nt > 0 ? strtoull(t[0], NULL, 0) : 0;
            odb_freeStrings(t, nt);
            nt = odb_getStrings(ao, "idmax", &t, "18446744073709551614");      /* gid_max */
            q->idMax = nt > 0 ? strtoull(t[0], NULL, 0) : 18446744073709551614ull;
            odb_freeStrings(t, nt);
            nt = odb_getStrings(ao, "idList", &t, NULL);
            q->nIdList = nt > 0 ? nt : 0;
            if (nt > 0)
            {
                q->idList = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)nt);
                for (int k = 0; k < nt; k++) q->idList[k] = strtoull(t[k], NULL, 0);
                for (int k = 1; k < nt; k++)               /* insertion sort: the lists are short */
                {
                    uint64_t v = q->idList[k];
                    int m = k - 1;
                    while (m >= 0 && q->idList[m] > v) { q->idList[m + 1] = q->idList[m]; m--; }
                    q->idList[m + 1] = v;
                }
            }
            odb_freeStrings(t, nt);
            q->includeSpecies = (int *)malloc(sizeof(int) * (size_t)(d->nspecies + 1));
            nt = odb_getStrings(ao, "species", &t, NULL);
            for (int k = 0; k < d->nspecies; k++) q->includeSpecies[k] = nt > 0 ? 0 : 1;
            for (int k = 0; k < nt; k++)
            {
                int hit = -1;
                for (int sidx = 0; sidx < d->nspecies; sidx++)
                    if (strcmp(d->speciesName[sidx], t[k]) == 0) hit = sidx;
                if (hit < 0) { herr("ANALYSIS %s: unknown species %s", an[a], t[k]); odb_freeStrings(t, nt); odb_freeStrings(an, na); rc = -1; goto done; }
                q->includeSpecies[hit] = 1;
            }
            odb_freeStrings(t, nt);
            /* bounds: defaults are -/+ the longest box edge, printed with %e as the reference does (src/subsetWrite.c:119-139) */
            double maxSize = fmax(fabs(d->params.h[0]), fmax(fabs(d->params.h[4]), fabs(d->params.h[8])));
            maxSize = hu_convert(maxSize, NULL, "l");
            char lo[32], hi[32], vlo[32], vhi[32];
            snprintf(lo, sizeof lo, "%e", -maxSize);
            snprintf(hi, sizeof hi, "%e", maxSize);
            snprintf(vlo, sizeof vlo, "%e", -DBL_MAX);
            snprintf(vhi, sizeof vhi, "%e", DBL_MAX);
            static const char *ax[3] = {"x", "y", "z"};
            int bad = 0;
            for (int k = 0; k < 3; k++)
            {
                char key[16];
                snprintf(key, sizeof key, "%smin", ax[k]);
                bad |= odb_getWithUnits(ao, key, &q->lo[k], 1, lo, "l", NULL) < 0;
                snprintf(key, sizeof key, "%smax", ax[k]);
                bad |= odb_getWithUnits(ao, key, &q->hi[k], 1, hi, "l", NULL) < 0;
                snprintf(key, sizeof key, "v%smin", ax[k]);
                bad |= odb_getWithUnits(ao, key, &q->vlo[k], 1, vlo, "l/t", NULL) < 0;
                snprintf(key, sizeof key, "v%smax", ax[k]);
                bad |= odb_getWithUnits(ao, key, &q->vhi[k], 1, vhi, "l/t", NULL) < 0;
            }
            if (bad) { herr("ANALYSIS %s: bad unit in a bound", an[a]); odb_freeStrings(an, na); rc = -1; goto done; }
            if (hu_convert(1.0, NULL, q->lengthUnit) != hu_convert(1.0, NULL, q->lengthUnit)) { herr("ANALYSIS %s: lengthUnit %s is not a unit", an[a], q->lengthUnit); odb_freeStrings(an, na); rc = -1; goto done; }
            const double lcv = hu_convert(1.0, NULL, "Angstrom"), vcv = hu_convert(1.0, NULL, "Angstrom/fs");
            char info[1024];
            snprintf(info, sizeof info, "idmin = %llu; idmax = %llu; modulus = %d; odd = %d;\n"
                     "xmin = %f Ang; xmax = %f Ang;\nymin = %f Ang; ymax = %f Ang;\nzmin = %f Ang; zmax = %f Ang;\n"
                     "vxmin = %f Ang/fs; vxmax = %f Ang/fs;\nvymin = %f Ang/fs; vymax = %f Ang/fs;\nvzmin = %f Ang/fs; vzmax = %f Ang/fs;\n",
                     (unsigned long long)q->idMin, (unsigned long long)q->idMax, q->modulus, q->odd,
                     q->lo[0] * lcv, q->hi[0] * lcv, q->lo[1] * lcv, q->hi[1] * lcv, q->lo[2] * lcv, q->hi[2] * lcv,
                     q->vlo[0] * vcv, q->vhi[0] * vcv, q->vlo[1] * vcv, q->vhi[1] * vcv, q->vlo[2] * vcv, q->vhi[2] * vcv);
            q->parmsInfo = strdup(info);
            if (q->modulus < 1) { herr("ANALYSIS %s: modulus must be >= 1", an[a]); odb_freeStrings(an, na); rc = -1; goto done; }
        }
        odb_freeStrings(an, na);
    }
    d->lengthPerAngstrom = hu_convert(1.0, "Angstrom", NULL);
    d->energyPerKJmol = hu_convert(1.0, "kJ/mol", NULL);
    d->massPerAmu = hu_convert(1.0, "amu", NULL);
    d->pressurePerBar = hu_convert(1.0, "bar", NULL);
    d->timePerFs = hu_convert(1.0, "fs", NULL);
    d->params.center[0] = d->params.center[1] = d->params.center[2] = 0.0;
    d->params.device = 0;

done:
    free(sysName); free(intName); free(ddcName); free(piName);
    free(boxName); free(nbrName); free(colName); free(mcName);
    odb_freeStrings(molNames, nMolNames);
    free(sorted);
    free(order);
    free(dir);
    freeMMFF(&mm);
    odb_free(db);
    if (rc)
    {
        ddcb200_deckFree(d);
        return rc;
    }
    *out = d;
    return 0;
}

/* One rank of nranks: push the parameters and static tables (every rank holds them), join the communicator,
 * then hand this rank a contiguous slice of the file-order beads; the first list build re-domains them
 * (ddcAssignment on the first call, src/ddcUpdateAll.c:81-119). */
int ddcb200_simulateBindRank(const ddcb200_deck *d, int device, int rank, int nranks, int lx, int ly, int lz, const unsigned char *ncclId,
                             ddcb200_ctx **out)
{
    if (!d || !out) return herr("null argument");
    if (nranks < 1 || rank < 0 || rank >= nranks) return herr("bad rank %d of %d", rank, nranks);
    ddcb200_params p = d->params;
    p.device = device;
    ddcb200_ctx *c = NULL;
    int rc = ddcb200_create(&p, &c);
    if (rc) return herr("ddcb200_create: %s", ddcb200_lastError());
#define TRY(call) do { rc = (call); if (rc) { herr(#call ": %s", ddcb200_lastError()); ddcb200_destroy(c); return rc; } } while (0)
    TRY(ddcb200_martiniNonBondParms(c, d->ntypes, d->ljEps, d->ljSigma, d->ljShift));
    TRY(ddcb200_setSpecies(c, d->nspecies, d->specLJ, d->specCharge, d->specMass));
    if ((d->excludePotentialTerm & 128) != 0) { herr("excludePotentialTerm nonBondMask is not supported"); ddcb200_destroy(c); return -1; }
    TRY(ddcb200_setExclusions(c, d->nMolTypes, d->specMolType, d->molTypeNSpecies, d->bpairOffset, d->bpairI, d->bpairJ));
    TRY(ddcb200_setBeads(c, d->n, d->gid, d->species));
    TRY(ddcb200_martiniBondParms(c, d->nTerms, d->termKind, d->termIdx, d->termParm));
    TRY(ddcb200_setRestraints(c, d->nRestraints, d->restrBead, d->restrFrac0, d->restrKb, d->restrFc, d->restrOrigin));
    TRY(ddcb200_setMolecules(c, d->nMol, d->molOffset, d->molBeads, d->nMolTotal));
    if (d->nGroups > 0)
    {
        double kBT[8];
        for (int g = 0; g < d->nGroups && g < 8; g++) kBT[g] = d->kB * d->groupTeq[g];   /* langevin_Update: kBT = kB*Teq */
        TRY(ddcb200_setGroups(c, d->nGroups, d->groupType, kBT, d->groupTau, d->groupVcm, d->n, d->groupOfBead));
    }
    if (d->haveRandom && d->rngState) TRY(ddcb200_setRandom(c, d->n, d->rngState, d->rngMult, d->rngPrime));
    if (d->integratorType == 1)
    {
        TRY(ddcb200_setConstraints(c, d->nCons, d->consAtomOffset, d->consAtomBead, d->consPairOffset, d->consPairA, d->consPairB, d->consPairDist));
        TRY(ddcb200_nglfconstraintParms(c, d->kB * d->ncT, d->ncP0, d->ncBeta, d->ncTauBarostat));
    }
    if (nranks > 1)
    {
        if (!ncclId) { herr("multi-rank bind needs the NCCL unique id"); ddcb200_destroy(c); return -1; }
        TRY(ddcb200_ddcInit(c, rank, nranks, lx, ly, lz, ncclId));
        /* the start-up partition only has to keep molecules whole (ddcRuleMolecule): every bead goes with its molecule's
         * ownership bead, and those are dealt out in equal index ranges; the first re-domain moves everything to its brick */
        int *ob = (int *)malloc(sizeof(int) * (size_t)(d->n + 1));
        for (int64_t i = 0; i < d->n; i++) ob[i] = (int)i;
        for (int64_t m = 0; m < d->nMol; m++)
            for (int64_t k = d->molOffset[m]; k < d->molOffset[m + 1]; k++) ob[d->molBeads[k]] = d->molBeads[d->molOffset[m]];
        int64_t cnt = 0;
        for (int64_t i = 0; i < d->n; i++) cnt += ((int64_t)ob[i] * nranks / d->n == rank);
        int *bead = (int *)malloc(sizeof(int) * (size_t)(cnt + 1));
        double *st = (double *)malloc(sizeof(double) * 6 * (size_t)(cnt + 1));
        int64_t k = 0;
        for (int64_t i = 0; i < d->n; i++)
            if ((int64_t)ob[i] * nranks / d->n == rank)
            {
                bead[k] = (int)i;
                st[k] = d->rx[i]; st[cnt + k] = d->ry[i]; st[2 * cnt + k] = d->rz[i];
                st[3 * cnt + k] = d->vx[i]; st[4 * cnt + k] = d->vy[i]; st[5 * cnt + k] = d->vz[i];
                k++;
            }
        free(ob);
        rc = cnt > 0 ? ddcb200_sendState(c, cnt, bead, st, st + cnt, st + 2 * cnt, st + 3 * cnt, st + 4 * cnt, st + 5 * cnt, d->loop, d->time) : -1;
        free(bead);
        free(st);
        if (cnt <= 0) { herr("rank %d of %d gets no beads at start-up", rank, nranks); ddcb200_destroy(c); return -1; }
    }
    else
        rc = ddcb200_sendState(c, d->n, NULL, d->rx, d->ry, d->rz, d->vx, d->vy, d->vz, d->loop, d->time);
    if (rc) { herr("ddcb200_sendState: %s", ddcb200_lastError()); ddcb200_destroy(c); return rc; }
#undef TRY
    *out = c;
    return 0;
}

int ddcb200_simulateBind(const ddcb200_deck *d, int device, ddcb200_ctx **out)
{
    return ddcb200_simulateBindRank(d, device, 0, 1, 1, 1, 1, NULL, out);
}

int ddcb200_printinfoLine(const ddcb200_deck *d, const ddcb200_etype *e, char *buf, size_t len)
{
    /* printinfoA, src/printinfo.c:125-232, in the PRINTINFO object's units.  The box edges are deck->params.h: a caller that
     * runs the barostat refreshes them from ddcb200_getBox before printing (ddcb200_simulateMaster does) */
    const double n = e->number;
    const double cL = d->printConvert[0], ct = d->printConvert[1], cT = d->printConvert[2], cE = d->printConvert[3], cP = d->printConvert[4],
                 cV = d->printConvert[5];
    const double pressure = d->printMolecularPressure ? e->pMolecular : e->pion;
    char loopFmt[16];
    snprintf(loopFmt, sizeof loopFmt, "%%%d.%dllu", d->nLoopDigits, d->nLoopDigits);   /* loopFormatInit, src/format.c:7-11 */
    int k = snprintf(buf, len, loopFmt, (unsigned long long)e->loop);
    if (k < 0 || (size_t)k >= len) return k;
    return k + snprintf(buf + k, len - (size_t)k, " %16.6f %18.12f %18.12f %18.12f %18.8f %18.12f %18.12f %15.8f %15.8f %15.8f",
                        ct * e->time, cE * ((e->eion + e->rk) / n), cE * (e->rk / n), cE * (e->eion / n),
                        cT * e->temperature, cP * pressure, cV * e->volume / n, cL * d->params.h[0], cL * d->params.h[4], cL * d->params.h[8]);
}

int ddcb200_printinfoHeader(const ddcb200_deck *d, char *buf, size_t len)
{
    /* src/printinfo.c:153-168: "#loop" left-justified in the loop width, then name(unit) columns */
    char col[10][96];
    static const char *name[10] = {"time", "Etotal", "Ekin", "Epot", "Temp", "Press", "Volume", "lx", "ly", "lz"};
    static const int unit[10] = {1, 3, 3, 3, 2, 4, 5, 0, 0, 0};
    for (int k = 0; k < 10; k++) snprintf(col[k], sizeof col[k], "%s(%s)", name[k], d->printUnit[unit[k]]);
    return snprintf(buf, len, "%-*s %16s %18s %18s %18s %18s %18s %18s %15s %15s %15s", d->nLoopDigits, "#loop", col[0], col[1], col[2], col[3],
                    col[4], col[5], col[6], col[7], col[8], col[9]);
}

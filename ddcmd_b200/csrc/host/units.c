/* units.c - unit algebra of ddcMD's object database, restated for the Martini decks.
 *
 * Follows units_internal/units_external/units_convert (reference src/units.c:450-551) with the
 * CODATA-2014 constants the reference compiles in (src/codata.h:58-89, selected at :11), in
 * the same operation order, because converted values feed bit-exact decisions (the list
 * cutoff, cell size and bead coordinates): value * from_mks / to_mks with
 * to_mks = prod_i pow(base_i, exponent_i).
 */
#include "host.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* CODATA 2014 (NIST), MKS */
#define a0_MKS 0.52917721067e-10
#define Rinfhc_MKS 2.179872325e-18
#define Rinfhc_eV 13.605693009
#define kB_eV 8.6173303e-5
#define kB_MKS 1.38064852e-23
#define e_MKS 1.6021766208e-19
#define u_MKS 1.660539040e-27
#define mp_MKS 1.672621898e-27
#define mn_MKS 1.674927471e-27
#define me_MKS 9.10938356e-31
#define Eh_MKS 4.359744650e-18
#define c_MKS 299792458.0
#define ke_MKS (c_MKS * c_MKS * (1e-7))
#define NA_MKS 6.022140857e23
#define cal_MKS 4.184

typedef struct
{
    const char *name;
    double dim[7]; /* length mass time current temperature amount luminous */
    double mks;
} SYMBOL;

enum { U_LENGTH, U_MASS, U_TIME, U_CURRENT, U_TEMPERATURE, U_AMOUNT, U_LUM, U_ENERGY, U_PRESSURE, U_VELOCITY, U_NBASE };

static SYMBOL table[] = {
    {"length_internal", {1, 0, 0, 0, 0, 0, 0}, 1.0},
    {"mass_internal", {0, 1, 0, 0, 0, 0, 0}, 1.0},
    {"time_internal", {0, 0, 1, 0, 0, 0, 0}, 1.0},
    {"current_internal", {0, 0, 0, 1, 0, 0, 0}, 1.0},
    {"temperature_internal", {0, 0, 0, 0, 1, 0, 0}, 1.0},
    {"amount_internal", {0, 0, 0, 0, 0, 1, 0}, 1.0},
    {"luminous_intensity_internal", {0, 0, 0, 0, 0, 0, 1}, 1.0},
    {"energy_internal", {2, 1, -2, 0, 0, 0, 0}, 1.0},
    {"pressure_internal", {-1, 1, -2, 0, 0, 0, 0}, 1.0},
    {"velocity_internal", {1, 0, -1, 0, 0, 0, 0}, 1.0},
    {"l", {1, 0, 0, 0, 0, 0, 0}, 1.0},
    {"m", {0, 1, 0, 0, 0, 0, 0}, 1.0},
    {"t", {0, 0, 1, 0, 0, 0, 0}, 1.0},
    {"i", {0, 0, 0, 1, 0, 0, 0}, 1.0},
    {"T", {0, 0, 0, 0, 1, 0, 0}, 1.0},
    {"n", {0, 0, 0, 0, 0, 1, 0}, 1.0},
    {"I", {0, 0, 0, 0, 0, 0, 1}, 1.0},
    {"energy", {2, 1, -2, 0, 0, 0, 0}, 1.0},
    {"pressure", {-1, 1, -2, 0, 0, 0, 0}, 1.0},
    {"velocity", {1, 0, -1, 0, 0, 0, 0}, 1.0},
    {"Ang", {1, 0, 0, 0, 0, 0, 0}, 1e-10},
    {"Angstrom", {1, 0, 0, 0, 0, 0, 0}, 1e-10},
    {"Bohr", {1, 0, 0, 0, 0, 0, 0}, a0_MKS},
    {"a0", {1, 0, 0, 0, 0, 0, 0}, a0_MKS},
    {"meter", {1, 0, 0, 0, 0, 0, 0}, 1.0},
    {"mm", {1, 0, 0, 0, 0, 0, 0}, 1.0e-3},
    {"um", {1, 0, 0, 0, 0, 0, 0}, 1.0e-6},
    {"nm", {1, 0, 0, 0, 0, 0, 0}, 1.0e-9},
    {"gram", {0, 1, 0, 0, 0, 0, 0}, 1.0e-3},
    {"g", {0, 1, 0, 0, 0, 0, 0}, 1.0e-3},
    {"kg", {0, 1, 0, 0, 0, 0, 0}, 1.0},
    {"eV", {2, 1, -2, 0, 0, 0, 0}, e_MKS * 1.0},
    {"keV", {2, 1, -2, 0, 0, 0, 0}, e_MKS * 1.0e+3},
    {"Hartree", {2, 1, -2, 0, 0, 0, 0}, Eh_MKS},
    {"Ry", {2, 1, -2, 0, 0, 0, 0}, Eh_MKS * 0.5},
    {"J", {2, 1, -2, 0, 0, 0, 0}, 1},
    {"kJ", {2, 1, -2, 0, 0, 0, 0}, 1.0e+3},
    {"cal", {2, 1, -2, 0, 0, 0, 0}, cal_MKS},
    {"kcal", {2, 1, -2, 0, 0, 0, 0}, cal_MKS * 1.0e+3},
    {"amu", {0, 1, 0, 0, 0, 0, 0}, u_MKS},
    {"second", {0, 0, 1, 0, 0, 0, 0}, 1},
    {"s", {0, 0, 1, 0, 0, 0, 0}, 1},
    {"ms", {0, 0, 1, 0, 0, 0, 0}, 1e-3},
    {"us", {0, 0, 1, 0, 0, 0, 0}, 1e-6},
    {"ns", {0, 0, 1, 0, 0, 0, 0}, 1e-9},
    {"ps", {0, 0, 1, 0, 0, 0, 0}, 1e-12},
    {"fs", {0, 0, 1, 0, 0, 0, 0}, 1e-15},
    {"cc", {3, 0, 0, 0, 0, 0, 0}, 1.0e-6},
    {"mol", {0, 0, 0, 0, 0, 1, 0}, NA_MKS},
    {"GPa", {-1, 1, -2, 0, 0, 0, 0}, 1e9},
    {"atm", {-1, 1, -2, 0, 0, 0, 0}, 1.01325e5},
    {"bar", {-1, 1, -2, 0, 0, 0, 0}, 1e5},
    {"Mbar", {-1, 1, -2, 0, 0, 0, 0}, 1e11},
    {"K", {0, 0, 0, 0, 1, 0, 0}, 1.0},
    {"coulomb", {0, 0, -1, 1, 0, 0, 0}, 1.0},
    {"C", {0, 0, -1, 1, 0, 0, 0}, 1.0},
    {"kB", {2, 1, -2, 0, -1, 0, 0}, kB_MKS},
    {"e", {0, 0, 1, 1, 0, 0, 0}, e_MKS},
    {"M_e", {0, 1, 0, 0, 0, 0, 0}, me_MKS},
    {"M_p", {0, 1, 0, 0, 0, 0, 0}, mp_MKS},
    {"M_n", {0, 1, 0, 0, 0, 0, 0}, mn_MKS},
    {NULL, {0, 0, 0, 0, 0, 0, 0}, 0.0}};

#define EXTERNAL_BASE 10 /* index of "l" */

static double g_kB = 0.0, g_ke = 0.0;
static int g_init = 0;

static void setBase(SYMBOL *s, double length, double mass, double time, double current, double temperature, double amount, double lum)
{
    s[U_LENGTH].mks = length;
    s[U_MASS].mks = mass;
    s[U_TIME].mks = time;
    s[U_CURRENT].mks = current;
    s[U_TEMPERATURE].mks = temperature;
    s[U_AMOUNT].mks = amount;
    s[U_LUM].mks = lum;
    s[U_ENERGY].mks = mass * length * length / (time * time);
    s[U_PRESSURE].mks = mass / (time * time * length);
    s[U_VELOCITY].mks = length / time;
}

void hu_init(void)
{
    if (g_init) return;
    /* units_internal(a0, Rinfhc*1e-30/a0^2, 1e-15, e/1e-15, Rinfhc_eV/kB_eV, 1, 1), reference src/ddcMD.c:71 */
    const double length = a0_MKS, mass = Rinfhc_MKS * 1e-30 / (a0_MKS * a0_MKS), time = 1e-15, current = e_MKS / 1e-15;
    const double temperature = Rinfhc_eV / kB_eV;
    setBase(table, length, mass, time, current, temperature, 1.0, 1.0);
    const double energy = mass * length * length / (time * time);
    const double charge = current * time;
    g_kB = kB_MKS * temperature / energy;              /* src/units.c:467 */
    g_ke = ke_MKS * charge * charge / (energy * length); /* src/units.c:470 */
    /* units_external(1e-10, u, 1e-15, e/1e-15, 1, 1, 1), src/ddcMD.c:72 */
    setBase(table + EXTERNAL_BASE, 1e-10, u_MKS, 1e-15, e_MKS / 1e-15, 1.0, 1.0, 1.0);
    g_init = 1;
}

double hu_kB(void) { hu_init(); return g_kB; }
double hu_ke(void) { hu_init(); return g_ke; }

/* ---- recursive-descent parser: expr := factor (('*'|'/') factor)* ; factor := atom ['^' number] */
static int parseExpr(const char **s, double dim[7], double *val);

static int parseAtom(const char **s, double dim[7], double *val)
{
    const char *p = *s;
    if (*p == '(')
    {
        p++;
        if (!parseExpr(&p, dim, val) || *p != ')') return 0;
        *s = p + 1;
        return 1;
    }
    if (*p == '1')
    {
        for (int i = 0; i < 7; i++) dim[i] = 0;
        *val = 1;
        *s = p + 1;
        return 1;
    }
    int len = 0;
    while ((p[len] >= 'a' && p[len] <= 'z') || (p[len] >= 'A' && p[len] <= 'Z') || p[len] == '_' || (len > 0 && p[len] >= '0' && p[len] <= '9')) len++;
    if (len == 0) return 0;
    /* the reference's names never end in digits except M_He3/M_He4/a0: try longest match first, then strip digits */
    for (int l = len; l > 0; l--)
    {
        for (SYMBOL *t = table; t->name; t++)
            if ((int)strlen(t->name) == l && strncmp(t->name, p, l) == 0)
            {
                for (int i = 0; i < 7; i++) dim[i] = t->dim[i];
                *val = t->mks;
                *s = p + l;
                return 1;
            }
        if (!(p[l - 1] >= '0' && p[l - 1] <= '9')) break;
    }
    return 0;
}

static int parseFactor(const char **s, double dim[7], double *val)
{
    if (!parseAtom(s, dim, val)) return 0;
    if (**s == '^')
    {
        char *end;
        double ex = strtod(*s + 1, &end);
        if (end == *s + 1) return 0;
        for (int i = 0; i < 7; i++) dim[i] *= ex;
        *val = pow(*val, ex);
        *s = end;
    }
    return 1;
}

static int parseExpr(const char **s, double dim[7], double *val)
{
    for (int i = 0; i < 7; i++) dim[i] = 0;
    *val = 1;
    int n = 0;
    for (;;)
    {
        double d[7], v;
        const char *save = *s;
        if (n == 0 && **s != '/')
        {
            if (!parseFactor(s, d, &v)) { *s = save; break; }
            for (int i = 0; i < 7; i++) dim[i] += d[i];
            *val *= v;
        }
        else if (**s == '*')
        {
            (*s)++;
            if (!parseFactor(s, d, &v)) { *s = save; break; }
            for (int i = 0; i < 7; i++) dim[i] += d[i];
            *val *= v;
        }
        else if (**s == '/')
        {
            (*s)++;
            if (!parseFactor(s, d, &v)) { *s = save; break; }
            for (int i = 0; i < 7; i++) dim[i] -= d[i];
            *val /= v;
        }
        else break;
        n++;
    }
    return n > 0;
}

static int unitsParse(const char *unit, double dim[7], double *mks)
{
    char buf[256];
    int j = 0;
    for (const char *p = unit; *p && j < 255; p++)
        if (*p != ' ' && *p != '\t') buf[j++] = *p;
    buf[j] = 0;
    const char *s = buf;
    if (!parseExpr(&s, dim, mks) || *s != 0) return 0;
    return 1;
}

/* units_convert(value, from, to); NULL = internal units.  Returns NaN on a parse or dimension error. */
double hu_convert(double value, const char *from, const char *to)
{
    hu_init();
    double fdim[7], tdim[7], fmks = 0, tmks = 0;
    if (!from && !to) return NAN;
    if (!from)
    {
        if (!unitsParse(to, tdim, &tmks)) return NAN;
        fmks = 1.0;
        for (int i = 0; i < 7; i++) fmks *= pow(table[i].mks, tdim[i]);
        return value * fmks / tmks;
    }
    if (!to)
    {
        if (!unitsParse(from, fdim, &fmks)) return NAN;
        tmks = 1.0;
        for (int i = 0; i < 7; i++) tmks *= pow(table[i].mks, fdim[i]);
        return value * fmks / tmks;
    }
    if (!unitsParse(to, tdim, &tmks) || !unitsParse(from, fdim, &fmks)) return NAN;
    double sum = 0;
    for (int i = 0; i < 7; i++) sum += fabs(fdim[i] - tdim[i]);
    if (sum > 1e-8) return NAN;
    return value * fmks / tmks;
}

/* host.h - internal declarations of the host-side (plain C) object-database layer. */
#ifndef DDCB200_HOST_INTERNAL_H
#define DDCB200_HOST_INTERNAL_H
#include <stddef.h>
#include <stdint.h>

/* units.c */
void hu_init(void);
double hu_kB(void);
double hu_ke(void);
double hu_convert(double value, const char *from, const char *to);

/* deck.c */
uint32_t hcrc32(const unsigned char *data, size_t len);
int herr(const char *fmt, ...);

/* objdb.c : "name CLASS { key=value; ... }" records (reference src/object.c) */
typedef struct
{
    char *name, *cls, *value; /* value = "key=v;key=v;" with comments stripped */
} ODB_OBJECT;
typedef struct
{
    ODB_OBJECT **obj; /* stable addresses: callers keep pointers while more files are compiled */
    int n, cap;
    char err[512];
} ODB;

ODB *odb_new(void);
void odb_free(ODB *db);
int odb_compileFile(ODB *db, const char *path);
int odb_compileString(ODB *db, const char *text);
const ODB_OBJECT *odb_find(const ODB *db, const char *name, const char *cls);
/* raw value of a key (malloc'd copy) or NULL; later assignments win, "+=" appends */
char *odb_value(const ODB_OBJECT *o, const char *key);
int odb_has(const ODB_OBJECT *o, const char *key);
/* typed getters; return number of elements found (dvalue used when the key is absent) */
int odb_getStrings(const ODB_OBJECT *o, const char *key, char ***out, const char *dvalue); /* caller frees each + array */
int odb_getString(const ODB_OBJECT *o, const char *key, char **out, const char *dvalue);
int odb_getInts(const ODB_OBJECT *o, const char *key, int *out, int max, const char *dvalue);
int odb_getI64(const ODB_OBJECT *o, const char *key, int64_t *out, const char *dvalue);
int odb_getDoubles(const ODB_OBJECT *o, const char *key, double *out, int max, const char *dvalue);
/* WITH_UNITS: numbers then an optional unit; converted from (unit or defUnit) to `to` (NULL = internal).
 * returns count, or -1 on a unit error */
int odb_getWithUnits(const ODB_OBJECT *o, const char *key, double *out, int max, const char *dvalue, const char *defUnit, const char *to);
void odb_freeStrings(char **s, int n);

#endif

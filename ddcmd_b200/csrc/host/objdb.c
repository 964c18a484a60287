/* objdb.c - reader for ddcMD's object database: text records
 *     name CLASS { keyword = value ; keyword += value ; ... }
 * with C and C++ comments (reference src/object.c:949-1107 object_read, :386-470
 * object_compilevalue, :471-705 object_parse).  Records with the same name and class are
 * merged in file order; for a repeated keyword the last "=" wins and "+=" appends, as the
 * reference does.  Only what the Martini decks use is implemented ($S/$B file slices and
 * FILETYPE values are not).
 */
#include "host.h"
#include <ctype.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static char *xstrdup(const char *s)
{
    size_t n = strlen(s) + 1;
    char *p = (char *)malloc(n);
    memcpy(p, s, n);
    return p;
}
static char *trimInPlace(char *s)
{
    while (*s && isspace((unsigned char)*s)) s++;
    size_t n = strlen(s);
    while (n > 0 && isspace((unsigned char)s[n - 1])) s[--n] = 0;
    return s;
}

ODB *odb_new(void)
{
    ODB *db = (ODB *)calloc(1, sizeof(ODB));
    return db;
}
void odb_free(ODB *db)
{
    if (!db) return;
    for (int i = 0; i < db->n; i++)
    {
        free(db->obj[i]->name);
        free(db->obj[i]->cls);
        free(db->obj[i]->value);
        free(db->obj[i]);
    }
    free(db->obj);
    free(db);
}

static int addObject(ODB *db, const char *name, const char *cls, const char *value)
{
    for (int i = 0; i < db->n; i++)
        if (strcmp(db->obj[i]->name, name) == 0 && strcmp(db->obj[i]->cls, cls) == 0)
        {
            size_t a = strlen(db->obj[i]->value), b = strlen(value);
            db->obj[i]->value = (char *)realloc(db->obj[i]->value, a + b + 2);
            memcpy(db->obj[i]->value + a, value, b + 1);
            return 0;
        }
    if (db->n == db->cap)
    {
        db->cap = db->cap ? 2 * db->cap : 64;
        db->obj = (ODB_OBJECT **)realloc(db->obj, db->cap * sizeof(ODB_OBJECT *));
    }
    db->obj[db->n] = (ODB_OBJECT *)malloc(sizeof(ODB_OBJECT));
    db->obj[db->n]->name = xstrdup(name);
    db->obj[db->n]->cls = xstrdup(cls);
    db->obj[db->n]->value = xstrdup(value);
    db->n++;
    return 0;
}

int odb_compileString(ODB *db, const char *text)
{
    /* 1. strip comments, fold whitespace, keep quoted strings */
    size_t n = strlen(text);
    char *buf = (char *)malloc(n + 2);
    size_t m = 0;
    for (size_t i = 0; i < n; i++)
    {
        char c = text[i];
        if (c == '"')
        {
            buf[m++] = c;
            for (i++; i < n && text[i] != '"'; i++) buf[m++] = text[i];
            if (i < n) buf[m++] = '"';
            continue;
        }
        if (c == '/' && i + 1 < n && text[i + 1] == '*')
        {
            for (i += 2; i + 1 < n && !(text[i] == '*' && text[i + 1] == '/'); i++) {}
            i++;
            buf[m++] = ' ';
            continue;
        }
        if (c == '/' && i + 1 < n && text[i + 1] == '/')
        {
            for (i += 2; i < n && text[i] != '\n'; i++) {}
            buf[m++] = ' ';
            continue;
        }
        if (c == '\n' || c == '\t' || c == '\r') c = ' ';
        buf[m++] = c;
    }
    buf[m] = 0;
    /* 2. split into records at '}' */
    char *p = buf;
    while (*p)
    {
        char *open = strchr(p, '{');
        if (!open) break;
        char *close = strchr(open, '}');
        if (!close)
        {
            snprintf(db->err, sizeof db->err, "object record without closing brace near: %.60s", p);
            free(buf);
            return -1;
        }
        *open = 0;
        *close = 0;
        char *head = trimInPlace(p);
        char name[256], cls[256];
        if (sscanf(head, "%255s %255s", name, cls) != 2)
        {
            snprintf(db->err, sizeof db->err, "bad object header: %.80s", head);
            free(buf);
            return -1;
        }
        char *val = trimInPlace(open + 1);
        size_t l = strlen(val);
        char *v2 = (char *)malloc(l + 2);
        memcpy(v2, val, l + 1);
        if (l > 0 && v2[l - 1] != ';')
        {
            v2[l] = ';';
            v2[l + 1] = 0;
        }
        addObject(db, name, cls, v2);
        free(v2);
        p = close + 1;
    }
    free(buf);
    return 0;
}

int odb_compileFile(ODB *db, const char *path)
{
    FILE *f = fopen(path, "rb");
    if (!f)
    {
        snprintf(db->err, sizeof db->err, "cannot open object file %s", path);
        return -1;
    }
    fseek(f, 0, SEEK_END);
    long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    char *text = (char *)malloc((size_t)sz + 1);
    size_t got = fread(text, 1, (size_t)sz, f);
    fclose(f);
    text[got] = 0;
    /* an atoms file starts with a FILEHEADER record followed by data: stop at the first blank line after '}' */
    int rc = odb_compileString(db, text);
    free(text);
    return rc;
}

const ODB_OBJECT *odb_find(const ODB *db, const char *name, const char *cls)
{
    for (int i = 0; i < db->n; i++)
        if (strcmp(db->obj[i]->name, name) == 0 && strcmp(db->obj[i]->cls, cls) == 0) return db->obj[i];
    return NULL;
}

char *odb_value(const ODB_OBJECT *o, const char *key)
{
    if (!o) return NULL;
    char *buf = xstrdup(o->value);
    char *result = NULL;
    char *save = NULL;
    for (char *tok = strtok_r(buf, ";", &save); tok; tok = strtok_r(NULL, ";", &save))
    {
        char *eq = strchr(tok, '=');
        if (!eq) continue;
        int append = (eq > tok && eq[-1] == '+');
        if (append) eq[-1] = 0;
        *eq = 0;
        char *k = trimInPlace(tok);
        if (strcmp(k, key) != 0) continue;
        char *v = trimInPlace(eq + 1);
        if (append && result)
        {
            size_t a = strlen(result), b = strlen(v);
            result = (char *)realloc(result, a + b + 2);
            result[a] = ' ';
            memcpy(result + a + 1, v, b + 1);
        }
        else
        {
            free(result);
            result = xstrdup(v);
        }
    }
    free(buf);
    return result;
}

int odb_has(const ODB_OBJECT *o, const char *key)
{
    char *v = odb_value(o, key);
    if (!v) return 0;
    free(v);
    return 1;
}

static int tokenize(char *v, char ***out)
{
    int n = 0, cap = 8;
    char **t = (char **)malloc(cap * sizeof(char *));
    char *p = v;
    while (*p)
    {
        while (*p && isspace((unsigned char)*p)) p++;
        if (!*p) break;
        char *start;
        if (*p == '"')
        {
            start = ++p;
            while (*p && *p != '"') p++;
        }
        else
        {
            start = p;
            while (*p && !isspace((unsigned char)*p)) p++;
        }
        char save = *p;
        *p = 0;
        if (n == cap)
        {
            cap *= 2;
            t = (char **)realloc(t, cap * sizeof(char *));
        }
        t[n++] = xstrdup(start);
        if (save) p++;
    }
    *out = t;
    return n;
}

void odb_freeStrings(char **s, int n)
{
    if (!s) return;
    for (int i = 0; i < n; i++) free(s[i]);
    free(s);
}

int odb_getStrings(const ODB_OBJECT *o, const char *key, char ***out, const char *dvalue)
{
    char *v = odb_value(o, key);
    if (!v)
    {
        if (!dvalue)
        {
            *out = NULL;
            return 0;
        }
        v = xstrdup(dvalue);
    }
    int n = tokenize(v, out);
    free(v);
    return n;
}

int odb_getString(const ODB_OBJECT *o, const char *key, char **out, const char *dvalue)
{
    char **t;
    int n = odb_getStrings(o, key, &t, dvalue);
    if (n == 0)
    {
        *out = NULL;
        odb_freeStrings(t, n);
        return 0;
    }
    *out = xstrdup(t[0]);
    odb_freeStrings(t, n);
    return n;
}

int odb_getInts(const ODB_OBJECT *o, const char *key, int *out, int max, const char *dvalue)
{
    char **t;
    int n = odb_getStrings(o, key, &t, dvalue);
    for (int i = 0; i < n && i < max; i++) out[i] = (int)strtol(t[i], NULL, 0);
    odb_freeStrings(t, n);
    return n;
}

int odb_getI64(const ODB_OBJECT *o, const char *key, int64_t *out, const char *dvalue)
{
    char **t;
    int n = odb_getStrings(o, key, &t, dvalue);
    if (n > 0) *out = (int64_t)strtoll(t[0], NULL, 0);
    odb_freeStrings(t, n);
    return n;
}

int odb_getDoubles(const ODB_OBJECT *o, const char *key, double *out, int max, const char *dvalue)
{
    char **t;
    int n = odb_getStrings(o, key, &t, dvalue);
    for (int i = 0; i < n && i < max; i++) out[i] = strtod(t[i], NULL);
    odb_freeStrings(t, n);
    return n;
}

int odb_getWithUnits(const ODB_OBJECT *o, const char *key, double *out, int max, const char *dvalue, const char *defUnit, const char *to)
{
    /* object_parse WITH_UNITS (reference src/object.c:607-612,693-701): every token is read with
     * strtod; whatever follows the number in the LAST token (or the last token itself when it is
     * not a number) is the unit. */
    char **t;
    int n = odb_getStrings(o, key, &t, dvalue);
    int nv = 0;
    char unit[128] = "";
    for (int i = 0; i < n; i++)
    {
        char *end;
        double d = strtod(t[i], &end);
        if (end == t[i])
        {
            snprintf(unit, sizeof unit, "%s", t[i]);
            break;
        }
        if (nv < max) out[nv] = d;
        nv++;
        if (*end) snprintf(unit, sizeof unit, "%s", end);
    }
    odb_freeStrings(t, n);
    const char *u = unit[0] ? unit : defUnit;
    double f = hu_convert(1.0, u, to);
    if (f != f) return -1;
    for (int i = 0; i < nv && i < max; i++) out[i] *= f;
    return nv;
}

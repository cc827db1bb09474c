/* A functional stand-in for the slice of R's runtime that r/src/sdegpu_shim.c touches (tests/mock_r/Rinternals.h).
 * Not R: no garbage collector (an arena freed by mock_reset), two attributes (names, dim), Rf_error and a pending
 * interrupt become longjmps to the harness.  What it does keep are the properties the shim's correctness depends on:
 * PROTECT/UNPROTECT balance is counted, Rf_error never returns, R_alloc memory lives until the call ends, unif_rand
 * is a reproducible stream between GetRNGstate/PutRNGstate, R_ToplevelExec reports FALSE when its callee jumps. */
#include "Rinternals.h"
#include "R_ext/Rdynload.h"

#include <math.h>
#include <setjmp.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

struct SEXPREC {
    int type;
    R_xlen_t len;
    void *data;            /* double[], int[], SEXP[] or char[] */
    SEXP names, dim;
};

static struct SEXPREC nil_rec = {NILSXP, 0, NULL, NULL, NULL}, names_sym = {NILSXP, 0, NULL, NULL, NULL}, dim_sym = {NILSXP, 0, NULL, NULL, NULL};
SEXP R_NilValue = &nil_rec, R_NamesSymbol = &names_sym, R_DimSymbol = &dim_sym;
double R_NaReal;

static void **arena = NULL;
static size_t arena_n = 0, arena_cap = 0;
static int protect_depth = 0, rng_open = 0, pending_interrupt = 0, interrupt_checks = 0, interrupt_after = -1;
static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static char last_error[1024];
static jmp_buf *jump_target = NULL;

static void *arena_alloc(size_t bytes)
{
    if (arena_n == arena_cap) {
        arena_cap = arena_cap ? 2 * arena_cap : 256;
        arena = (void **)realloc(arena, arena_cap * sizeof(void *));
    }
    void *p = calloc(bytes ? bytes : 1, 1);
    arena[arena_n++] = p;
    return p;
}

static size_t elt_size(int type)
{
    switch (type) {
    case REALSXP: return sizeof(double);
    case INTSXP: case LGLSXP: return sizeof(int);
    case STRSXP: case VECSXP: return sizeof(SEXP);
    case CHARSXP: return 1;
    default: return 0;
    }
}

SEXP Rf_allocVector(SEXPTYPE type, R_xlen_t n)
{
    if (n < 0) Rf_error("allocVector: negative length");
    SEXP s = (SEXP)arena_alloc(sizeof(struct SEXPREC));
    s->type = (int)type; s->len = n; s->names = R_NilValue; s->dim = R_NilValue;
    s->data = arena_alloc(elt_size((int)type) * (size_t)(n + (type == CHARSXP)));
    if (type == VECSXP || type == STRSXP) for (R_xlen_t i = 0; i < n; ++i) ((SEXP *)s->data)[i] = R_NilValue;
    return s;
}

static SEXP with_dim(SEXP s, int nd, const int *d)
{
    SEXP dim = Rf_allocVector(INTSXP, nd);
    for (int i = 0; i < nd; ++i) INTEGER(dim)[i] = d[i];
    s->dim = dim;
    return s;
}
SEXP Rf_allocMatrix(SEXPTYPE t, int nr, int nc)
{
    if (nr < 0 || nc < 0) Rf_error("allocMatrix: negative extents");
    const int d[2] = {nr, nc};
    return with_dim(Rf_allocVector(t, (R_xlen_t)nr * nc), 2, d);
}
SEXP Rf_alloc3DArray(SEXPTYPE t, int a, int b, int c)
{
    if (a < 0 || b < 0 || c < 0) Rf_error("alloc3DArray: negative extents");
    const int d[3] = {a, b, c};
    return with_dim(Rf_allocVector(t, (R_xlen_t)a * b * c), 3, d);
}
SEXP Rf_duplicate(SEXP x)
{
    if (x == R_NilValue) return x;
    SEXP s = Rf_allocVector((SEXPTYPE)x->type, x->len);
    memcpy(s->data, x->data, elt_size(x->type) * (size_t)x->len);
    s->names = x->names; s->dim = x->dim;
    return s;
}
double *REAL(SEXP x) { if (x->type != REALSXP) Rf_error("REAL() applied to a non-numeric object (type %d)", x->type); return (double *)x->data; }
int *INTEGER(SEXP x) { if (x->type != INTSXP && x->type != LGLSXP) Rf_error("INTEGER() applied to a non-integer object (type %d)", x->type); return (int *)x->data; }
R_xlen_t XLENGTH(SEXP x) { return x->len; }
int TYPEOF(SEXP x) { return x->type; }
const char *CHAR(SEXP x) { return (const char *)x->data; }
SEXP STRING_ELT(SEXP x, R_xlen_t i) { if (x->type != STRSXP || i < 0 || i >= x->len) Rf_error("STRING_ELT out of range"); return ((SEXP *)x->data)[i]; }
SEXP VECTOR_ELT(SEXP x, R_xlen_t i) { if (x->type != VECSXP || i < 0 || i >= x->len) Rf_error("VECTOR_ELT out of range"); return ((SEXP *)x->data)[i]; }
SEXP SET_VECTOR_ELT(SEXP x, R_xlen_t i, SEXP v) { if (x->type != VECSXP || i < 0 || i >= x->len) Rf_error("SET_VECTOR_ELT out of range"); ((SEXP *)x->data)[i] = v; return v; }
void SET_STRING_ELT(SEXP x, R_xlen_t i, SEXP v) { if (x->type != STRSXP || i < 0 || i >= x->len) Rf_error("SET_STRING_ELT out of range"); ((SEXP *)x->data)[i] = v; }
SEXP Rf_getAttrib(SEXP x, SEXP what) { return x == R_NilValue ? R_NilValue : (what == R_NamesSymbol ? x->names : (what == R_DimSymbol ? x->dim : R_NilValue)); }
SEXP Rf_setAttrib(SEXP x, SEXP what, SEXP v) { if (what == R_NamesSymbol) x->names = v; else if (what == R_DimSymbol) x->dim = v; return v; }
SEXP Rf_mkChar(const char *s)
{
    SEXP c = Rf_allocVector(CHARSXP, (R_xlen_t)strlen(s));
    memcpy(c->data, s, strlen(s) + 1);
    return c;
}
SEXP Rf_mkString(const char *s) { SEXP v = Rf_allocVector(STRSXP, 1); SET_STRING_ELT(v, 0, Rf_mkChar(s)); return v; }
SEXP Rf_ScalarReal(double v) { SEXP s = Rf_allocVector(REALSXP, 1); REAL(s)[0] = v; return s; }
SEXP Rf_ScalarInteger(int v) { SEXP s = Rf_allocVector(INTSXP, 1); INTEGER(s)[0] = v; return s; }
Rboolean Rf_isNull(SEXP x) { return x == R_NilValue || x->type == NILSXP; }
Rboolean Rf_isReal(SEXP x) { return x->type == REALSXP; }
int Rf_asInteger(SEXP x)
{
    if (x->len < 1) return NA_INTEGER;
    if (x->type == INTSXP || x->type == LGLSXP) return ((int *)x->data)[0];
    if (x->type == REALSXP) { const double v = ((double *)x->data)[0]; return (isnan(v) || v > 2147483647.0 || v <= -2147483648.0) ? NA_INTEGER : (int)v; }
    return NA_INTEGER;
}
double Rf_asReal(SEXP x)
{
    if (x->len < 1) return R_NaReal;
    if (x->type == REALSXP) return ((double *)x->data)[0];
    if (x->type == INTSXP || x->type == LGLSXP) { const int v = ((int *)x->data)[0]; return v == NA_INTEGER ? R_NaReal : (double)v; }
    return R_NaReal;
}
SEXP Rf_protect(SEXP s) { ++protect_depth; return s; }
void Rf_unprotect(int n) { protect_depth -= n; if (protect_depth < 0) { snprintf(last_error, sizeof last_error, "unprotect(): stack imbalance"); abort(); } }
char *R_alloc(size_t n, int size) { return (char *)arena_alloc(n * (size_t)size); }
int R_finite(double x) { return isfinite(x); }

void Rf_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(last_error, sizeof last_error, fmt, ap);
    va_end(ap);
    if (!jump_target) { fprintf(stderr, "Rf_error outside mock_call: %s\n", last_error); abort(); }
    longjmp(*jump_target, 1);
}

void GetRNGstate(void) { rng_open = 1; }
void PutRNGstate(void) { rng_open = 0; }
double unif_rand(void)
{
    if (!rng_open) Rf_error("unif_rand() outside GetRNGstate()/PutRNGstate()");
    rng_state = rng_state * 6364136223846793005ull + 1442695040888963407ull;
    return (double)((rng_state >> 11) + 1) * (1.0 / 9007199254740994.0);      /* in (0,1) like R */
}

void R_CheckUserInterrupt(void)
{
    ++interrupt_checks;
    if (interrupt_after >= 0 && interrupt_checks > interrupt_after) pending_interrupt = 1;
    if (pending_interrupt) {
        pending_interrupt = 0;
        snprintf(last_error, sizeof last_error, "interrupt");
        if (!jump_target) abort();
        longjmp(*jump_target, 2);
    }
}

Rboolean R_ToplevelExec(void (*fun)(void *), void *data)
{
    jmp_buf here, *outer = jump_target;
    const int depth = protect_depth;
    jump_target = &here;
    if (setjmp(here) == 0) {
        fun(data);
        jump_target = outer;
        return TRUE;
    }
    jump_target = outer;
    protect_depth = depth;
    return FALSE;
}

int R_registerRoutines(DllInfo *dll, const void *c, const R_CallMethodDef *call, const void *f, const void *e) { (void)dll; (void)c; (void)call; (void)f; (void)e; return 1; }
int R_useDynamicSymbols(DllInfo *dll, int v) { (void)dll; (void)v; return 0; }

/* ---- harness side (ctypes) ------------------------------------------------------------------------------------- */
typedef SEXP (*call_fn)();
/* .Call(fn, args...): 0 = returned (result in *res), 1 = Rf_error, 2 = interrupt; the PROTECT stack must be balanced
 * after a normal return (after a jump R unwinds it itself) */
int mock_call(call_fn fn, int nargs, SEXP *a, SEXP *res)
{
    jmp_buf here;
    jump_target = &here;
    protect_depth = 0;
    R_NaReal = NAN;
    const int code = setjmp(here);
    if (code == 0) {
        SEXP r = R_NilValue;
        switch (nargs) {
        case 0: r = fn(); break;
        case 1: r = fn(a[0]); break;
        case 2: r = fn(a[0], a[1]); break;
        case 3: r = fn(a[0], a[1], a[2]); break;
        case 4: r = fn(a[0], a[1], a[2], a[3]); break;
        case 5: r = fn(a[0], a[1], a[2], a[3], a[4]); break;
        case 6: r = fn(a[0], a[1], a[2], a[3], a[4], a[5]); break;
        case 7: r = fn(a[0], a[1], a[2], a[3], a[4], a[5], a[6]); break;
        case 8: r = fn(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7]); break;
        default: snprintf(last_error, sizeof last_error, "mock_call: too many arguments"); jump_target = NULL; return 1;
        }
        jump_target = NULL;
        if (protect_depth != 0) { snprintf(last_error, sizeof last_error, "PROTECT stack imbalance: depth %d at return", protect_depth); return 3; }
        *res = r;
        return 0;
    }
    jump_target = NULL;
    rng_open = 0;
    return code;
}
const char *mock_last_error(void) { return last_error; }
void mock_reset(void)
{
    for (size_t i = 0; i < arena_n; ++i) free(arena[i]);
    arena_n = 0;
    protect_depth = 0; pending_interrupt = 0; interrupt_checks = 0; interrupt_after = -1;
}
void mock_set_seed(uint64_t s) { rng_state = s * 0x9E3779B97F4A7C15ull + 1; }
void mock_interrupt_after(int checks) { interrupt_after = checks; interrupt_checks = 0; }   /* -1: never */
int mock_interrupt_checks(void) { return interrupt_checks; }
SEXP mock_nil(void) { return R_NilValue; }
SEXP mock_real(const double *v, R_xlen_t n) { SEXP s = Rf_allocVector(REALSXP, n); if (v && n) memcpy(s->data, v, sizeof(double) * (size_t)n); return s; }
SEXP mock_int(const int *v, R_xlen_t n) { SEXP s = Rf_allocVector(INTSXP, n); if (v && n) memcpy(s->data, v, sizeof(int) * (size_t)n); return s; }
SEXP mock_list(R_xlen_t n) { SEXP s = Rf_allocVector(VECSXP, n); s->names = Rf_allocVector(STRSXP, n); return s; }
void mock_list_set(SEXP l, R_xlen_t i, const char *name, SEXP v) { SET_VECTOR_ELT(l, i, v); SET_STRING_ELT(l->names, i, Rf_mkChar(name)); }
void mock_set_dim(SEXP s, int nd, const int *d) { with_dim(s, nd, d); }
int mock_type(SEXP s) { return s->type; }
R_xlen_t mock_length(SEXP s) { return s->len; }
void *mock_data(SEXP s) { return s->data; }
int mock_ndim(SEXP s) { return s->dim == R_NilValue ? 0 : (int)s->dim->len; }
int mock_dim(SEXP s, int k) { return ((int *)s->dim->data)[k]; }
SEXP mock_elt(SEXP l, R_xlen_t i) { return ((SEXP *)l->data)[i]; }
const char *mock_name(SEXP l, R_xlen_t i) { return l->names == R_NilValue ? "" : (const char *)((SEXP *)l->names->data)[i]->data; }

/* Stand-in for the part of R's C API that r/src/sdegpu_shim.c uses, for an image without R.  Signatures follow
 * R 4.x's Rinternals.h / Rdynload.h / Utils.h.  tests/mock_r/mock_runtime.c implements them (a small arena of typed
 * vectors with names/dim attributes, Rf_error as a longjmp to the test harness, a PROTECT depth counter), so the shim
 * is not only type-checked but executed: tests/test_r_shim.py drives its .Call entry points through ctypes. */
#ifndef MOCK_RINTERNALS_H
#define MOCK_RINTERNALS_H
#include <stddef.h>
typedef struct SEXPREC *SEXP;
typedef ptrdiff_t R_xlen_t;
typedef unsigned int SEXPTYPE;
#define NILSXP 0
#define CHARSXP 9
#define LGLSXP 10
#define INTSXP 13
#define REALSXP 14
#define STRSXP 16
#define VECSXP 19
#define FALSE 0
#define TRUE 1
typedef int Rboolean;
extern SEXP R_NilValue, R_NamesSymbol, R_DimSymbol;
extern double R_NaReal;
#define NA_REAL R_NaReal
#define NA_INTEGER (-2147483647 - 1)
double *REAL(SEXP);
int *INTEGER(SEXP);
R_xlen_t XLENGTH(SEXP);
int TYPEOF(SEXP);
const char *CHAR(SEXP);
SEXP STRING_ELT(SEXP, R_xlen_t);
SEXP VECTOR_ELT(SEXP, R_xlen_t);
SEXP SET_VECTOR_ELT(SEXP, R_xlen_t, SEXP);
void SET_STRING_ELT(SEXP, R_xlen_t, SEXP);
SEXP Rf_getAttrib(SEXP, SEXP);
SEXP Rf_setAttrib(SEXP, SEXP, SEXP);
SEXP Rf_allocVector(SEXPTYPE, R_xlen_t);
SEXP Rf_allocMatrix(SEXPTYPE, int, int);
SEXP Rf_alloc3DArray(SEXPTYPE, int, int, int);
SEXP Rf_duplicate(SEXP);
SEXP Rf_mkChar(const char *);
SEXP Rf_ScalarReal(double);
SEXP Rf_ScalarInteger(int);
SEXP Rf_mkString(const char *);
Rboolean Rf_isNull(SEXP);
Rboolean Rf_isReal(SEXP);
int Rf_asInteger(SEXP);
double Rf_asReal(SEXP);
SEXP Rf_protect(SEXP);
void Rf_unprotect(int);
#define PROTECT(s) Rf_protect(s)
#define UNPROTECT(n) Rf_unprotect(n)
void Rf_error(const char *, ...) __attribute__((noreturn));
char *R_alloc(size_t, int);
void GetRNGstate(void);
void PutRNGstate(void);
double unif_rand(void);
int R_finite(double);
void R_CheckUserInterrupt(void);
Rboolean R_ToplevelExec(void (*fun)(void *), void *data);
#endif

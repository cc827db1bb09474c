/* Minimal stand-in for R's C API, only so that r/src/sdegpu_shim.c can be type-checked in an image
 * without R.  Signatures follow R 4.x's Rinternals.h / Rdynload.h; nothing here is ever linked. */
#ifndef MOCK_RINTERNALS_H
#define MOCK_RINTERNALS_H
#include <stddef.h>
typedef struct SEXPREC *SEXP;
typedef ptrdiff_t R_xlen_t;
typedef unsigned int SEXPTYPE;
#define STRSXP 16
#define VECSXP 19
#define REALSXP 14
#define INTSXP 13
#define FALSE 0
#define TRUE 1
typedef int Rboolean;
extern SEXP R_NilValue, R_NamesSymbol, R_DimSymbol;
double *REAL(SEXP);
int *INTEGER(SEXP);
R_xlen_t XLENGTH(SEXP);
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
Rboolean Rf_isNull(SEXP);
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
#endif

/* tests/mini_r/include/Rinternals.h -- TEST INFRASTRUCTURE. A miniature stand-in for R's C API, just large enough
 * to COMPILE AND RUN r/shim.c (the `.Call` glue a sparseRGPs maintainer adds) in an image without R, so that the
 * drop-in boundary is exercised at the SEXP level: routine registration (R_init_sparseRGPs), argument coercion,
 * named-list lookup, the `matrix()` NA sentinel, result construction, Rf_error propagation, PROTECT balance.
 * Semantics follow "Writing R Extensions" for the calls used; nothing here is R source. */
#ifndef MINI_RINTERNALS_H
#define MINI_RINTERNALS_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef ptrdiff_t R_xlen_t;
typedef enum { FALSE = 0, TRUE } Rboolean;

typedef unsigned int SEXPTYPE;
#define NILSXP 0
#define CHARSXP 9
#define LGLSXP 10
#define INTSXP 13
#define REALSXP 14
#define STRSXP 16
#define VECSXP 19

typedef struct mini_sexp *SEXP;

extern SEXP R_NilValue;
extern SEXP R_NamesSymbol;
extern SEXP R_DimSymbol;
extern double R_NaReal;   /* NA_real_: a NaN with payload 1954 */
extern double R_NaN;
extern int R_NaInt;
#define NA_REAL R_NaReal
#define NA_INTEGER R_NaInt
#define NA_LOGICAL R_NaInt

int R_IsNA(double);
int R_IsNaN(double);
#define ISNA(x) R_IsNA(x)
#define ISNAN(x) ((x) != (x))

SEXP Rf_protect(SEXP);
void Rf_unprotect(int);
#define PROTECT(s) Rf_protect(s)
#define UNPROTECT(n) Rf_unprotect(n)

SEXP Rf_allocVector(SEXPTYPE, R_xlen_t);
SEXP Rf_allocMatrix(SEXPTYPE, int, int);
SEXP Rf_coerceVector(SEXP, SEXPTYPE);
SEXP Rf_duplicate(SEXP);
SEXP Rf_lengthgets(SEXP, int);
int Rf_length(SEXP);
R_xlen_t Rf_xlength(SEXP);
int Rf_nrows(SEXP);
int Rf_ncols(SEXP);
double Rf_asReal(SEXP);
int Rf_asInteger(SEXP);
int Rf_asLogical(SEXP);
SEXP Rf_ScalarReal(double);
SEXP Rf_ScalarInteger(int);
SEXP Rf_ScalarLogical(int);
SEXP Rf_mkChar(const char *);
SEXP Rf_mkString(const char *);
SEXP Rf_getAttrib(SEXP, SEXP);
SEXP Rf_setAttrib(SEXP, SEXP, SEXP);
int TYPEOF(SEXP);
double *REAL(SEXP);
int *INTEGER(SEXP);
int *LOGICAL(SEXP);
const char *CHAR(SEXP);
SEXP STRING_ELT(SEXP, R_xlen_t);
void SET_STRING_ELT(SEXP, R_xlen_t, SEXP);
SEXP VECTOR_ELT(SEXP, R_xlen_t);
SEXP SET_VECTOR_ELT(SEXP, R_xlen_t, SEXP);

void Rf_error(const char *, ...) __attribute__((noreturn, format(printf, 1, 2)));
void Rf_warning(const char *, ...) __attribute__((format(printf, 1, 2)));

#ifdef __cplusplus
}
#endif
#endif

/* Stub of R's <Rinternals.h>: see R.h in this directory. */
#ifndef MCB_STUB_RINTERNALS_H
#define MCB_STUB_RINTERNALS_H
#include <stddef.h>
#include "R.h"
typedef struct SEXPREC *SEXP;
typedef ptrdiff_t R_xlen_t;
typedef unsigned char Rbyte;
typedef unsigned int SEXPTYPE;
#define INTSXP 13
#define REALSXP 14
#define STRSXP 16
#define VECSXP 19
#define RAWSXP 24
extern SEXP R_NilValue;
void Rf_error(const char *, ...) __attribute__((noreturn));
SEXP Rf_allocVector(SEXPTYPE, R_xlen_t);
SEXP Rf_protect(SEXP);
void Rf_unprotect(int);
#define PROTECT(s) Rf_protect(s)
#define UNPROTECT(n) Rf_unprotect(n)
double Rf_asReal(SEXP);
int Rf_asInteger(SEXP);
int Rf_asLogical(SEXP);
Rboolean Rf_isNull(SEXP);
Rboolean Rf_isInteger(SEXP);
R_xlen_t XLENGTH(SEXP);
int *INTEGER(SEXP);
double *REAL(SEXP);
Rbyte *RAW(SEXP);
SEXP STRING_ELT(SEXP, R_xlen_t);
void SET_STRING_ELT(SEXP, R_xlen_t, SEXP);
SEXP SET_VECTOR_ELT(SEXP, R_xlen_t, SEXP);
const char *CHAR(SEXP);
SEXP Rf_mkChar(const char *);
SEXP Rf_ScalarInteger(int);
SEXP Rf_ScalarReal(double);
void *R_ExternalPtrAddr(SEXP);
void R_ClearExternalPtr(SEXP);
SEXP R_MakeExternalPtr(void *, SEXP, SEXP);
typedef void (*R_CFinalizer_t)(SEXP);
void R_RegisterCFinalizerEx(SEXP, R_CFinalizer_t, Rboolean);
#endif

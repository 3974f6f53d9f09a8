/* see R.h in this directory */
#ifndef MMS_RINTERNALS_STUB_H
#define MMS_RINTERNALS_STUB_H
#include "R.h"
#define REALSXP 14
#define INTSXP 13
#define STRSXP 16
#define VECSXP 19
extern SEXP R_NilValue, R_NamesSymbol, R_DimSymbol;
int TYPEOF(SEXP);
int Rf_isMatrix(SEXP);            /* Rboolean in R */
char *R_alloc(size_t, int);       /* R_ext/Memory.h */
double *REAL(SEXP);
int *INTEGER(SEXP);
SEXP Rf_allocVector(unsigned int, R_xlen_t);
SEXP Rf_allocMatrix(unsigned int, int, int);
SEXP Rf_protect(SEXP);
void Rf_unprotect(int);
#define PROTECT(s) Rf_protect(s)
#define UNPROTECT(n) Rf_unprotect(n)
void Rf_error(const char *, ...) __attribute__((noreturn));
SEXP Rf_getAttrib(SEXP, SEXP);
SEXP Rf_setAttrib(SEXP, SEXP, SEXP);
R_xlen_t Rf_xlength(SEXP);
int Rf_nrows(SEXP);
int Rf_ncols(SEXP);
int Rf_asInteger(SEXP);
double Rf_asReal(SEXP);
SEXP Rf_ScalarInteger(int);
SEXP Rf_mkChar(const char *);
SEXP STRING_ELT(SEXP, R_xlen_t);
SEXP VECTOR_ELT(SEXP, R_xlen_t);
SEXP SET_VECTOR_ELT(SEXP, R_xlen_t, SEXP);
void SET_STRING_ELT(SEXP, R_xlen_t, SEXP);
const char *CHAR(SEXP);
#endif

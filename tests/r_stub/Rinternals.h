/* tests/r_stub -- TEST INFRASTRUCTURE.  A small stand-in for the part of R's C API that
 * rsplash_b200/rglue/rglue.cpp uses, so that the glue can be compiled, linked against libsplash_cuda and
 * exercised from pytest in an image without R.  Semantics follow "Writing R Extensions" (vectors are typed,
 * length-carrying, column-major; names are an attribute; Rf_error does not return).  Memory is never freed
 * (the tests allocate a few MB); PROTECT/UNPROTECT only count.  Not shipped, not used by the product. */
#ifndef R_STUB_RINTERNALS_H
#define R_STUB_RINTERNALS_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef ptrdiff_t R_xlen_t;
typedef struct stub_sexp {
    int type;
    R_xlen_t len;
    int nrow, ncol;        /* 0, 0 for plain vectors */
    void* data;            /* double[], int[], SEXP[] or char[] */
    struct stub_sexp* names;
} * SEXP;
enum { NILSXP = 0, CHARSXP = 9, LGLSXP = 10, INTSXP = 13, REALSXP = 14, STRSXP = 16, VECSXP = 19 };
extern SEXP R_NilValue, R_NamesSymbol;
double* REAL(SEXP x);
int* INTEGER(SEXP x);
int* LOGICAL(SEXP x);
R_xlen_t XLENGTH(SEXP x);
int TYPEOF(SEXP x);
void R_PreserveObject(SEXP x);
void R_ReleaseObject(SEXP x);
SEXP VECTOR_ELT(SEXP x, R_xlen_t i);
SEXP SET_VECTOR_ELT(SEXP x, R_xlen_t i, SEXP v);
void SET_STRING_ELT(SEXP x, R_xlen_t i, SEXP v);
SEXP Rf_allocVector(int type, R_xlen_t n);
SEXP Rf_allocMatrix(int type, int nrow, int ncol);
SEXP Rf_mkChar(const char* s);
SEXP Rf_mkNamed(int type, const char** names); /* names terminated by "" */
SEXP Rf_setAttrib(SEXP x, SEXP name, SEXP val);
int Rf_asInteger(SEXP x);
int Rf_asLogical(SEXP x);
double Rf_asReal(SEXP x);
SEXP Rf_protect(SEXP x);
void Rf_unprotect(int n);
#define PROTECT(x) Rf_protect(x)
#define UNPROTECT(n) Rf_unprotect(n)
#ifdef __cplusplus
[[noreturn]]
#endif
void Rf_error(const char* fmt, ...);
/* --- helpers for the test driver (not part of R) --- */
const char* stub_name(SEXP list, R_xlen_t i); /* names(list)[i] */
int stub_protect_depth(void);
const char* stub_last_error(void);            /* message of the last Rf_error caught by stub_call */
#ifdef __cplusplus
}
#endif
#endif

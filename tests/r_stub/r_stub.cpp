// tests/r_stub -- TEST INFRASTRUCTURE (see Rinternals.h): the stand-in's implementation plus a tiny driver API
// (stub_*) that pytest uses through ctypes to build arguments, call a registered .Call routine and read results.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>

#include "R.h"
#include "Rinternals.h"
#include "R_ext/Rdynload.h"

static stub_sexp g_nil = {NILSXP, 0, 0, 0, nullptr, nullptr}, g_names_sym = {NILSXP, 0, 0, 0, nullptr, nullptr};
SEXP R_NilValue = &g_nil, R_NamesSymbol = &g_names_sym;
static int g_depth = 0;
static std::string g_err;

struct r_error : std::runtime_error {
    using std::runtime_error::runtime_error;
};

static void need(SEXP x, int type, const char* what) {
    if (!x || x->type != type) Rf_error("stub: %s applied to an object of another type", what);
}
extern "C" {
double* REAL(SEXP x) { need(x, REALSXP, "REAL()"); return (double*)x->data; }
int* INTEGER(SEXP x) { if (!x || (x->type != INTSXP && x->type != LGLSXP)) Rf_error("stub: INTEGER() on a non-integer"); return (int*)x->data; }
int* LOGICAL(SEXP x) { need(x, LGLSXP, "LOGICAL()"); return (int*)x->data; }
R_xlen_t XLENGTH(SEXP x) { return x->len; }
int TYPEOF(SEXP x) { return x ? x->type : NILSXP; }
static int g_preserved = 0;
void R_PreserveObject(SEXP) { ++g_preserved; }
void R_ReleaseObject(SEXP) { --g_preserved; }
int stub_preserved(void) { return g_preserved; }
SEXP VECTOR_ELT(SEXP x, R_xlen_t i) { need(x, VECSXP, "VECTOR_ELT()"); return ((SEXP*)x->data)[i]; }
SEXP SET_VECTOR_ELT(SEXP x, R_xlen_t i, SEXP v) {
    need(x, VECSXP, "SET_VECTOR_ELT()");
    if (i < 0 || i >= x->len) Rf_error("stub: SET_VECTOR_ELT index out of range");
    ((SEXP*)x->data)[i] = v;
    return v;
}
void SET_STRING_ELT(SEXP x, R_xlen_t i, SEXP v) {
    need(x, STRSXP, "SET_STRING_ELT()");
    if (i < 0 || i >= x->len) Rf_error("stub: SET_STRING_ELT index out of range");
    ((SEXP*)x->data)[i] = v;
}
SEXP Rf_allocVector(int type, R_xlen_t n) {
    SEXP s = (SEXP)calloc(1, sizeof(stub_sexp));
    s->type = type;
    s->len = n;
    const size_t el = (type == REALSXP) ? 8 : (type == INTSXP || type == LGLSXP) ? 4 : (type == CHARSXP) ? 1 : sizeof(SEXP);
    s->data = calloc((size_t)(n > 0 ? n : 1) + (type == CHARSXP), el);
    if (type == VECSXP || type == STRSXP)
        for (R_xlen_t i = 0; i < n; ++i) ((SEXP*)s->data)[i] = R_NilValue;
    return s;
}
SEXP Rf_allocMatrix(int type, int nrow, int ncol) {
    SEXP s = Rf_allocVector(type, (R_xlen_t)nrow * ncol);
    s->nrow = nrow;
    s->ncol = ncol;
    return s;
}
SEXP Rf_mkChar(const char* str) {
    SEXP s = Rf_allocVector(CHARSXP, (R_xlen_t)strlen(str));
    memcpy(s->data, str, strlen(str) + 1);
    return s;
}
SEXP Rf_mkNamed(int type, const char** names) {
    R_xlen_t n = 0;
    while (names[n][0]) ++n;
    SEXP s = Rf_allocVector(type, n), nm = Rf_allocVector(STRSXP, n);
    for (R_xlen_t i = 0; i < n; ++i) SET_STRING_ELT(nm, i, Rf_mkChar(names[i]));
    s->names = nm;
    return s;
}
SEXP Rf_setAttrib(SEXP x, SEXP name, SEXP val) {
    if (name != R_NamesSymbol) Rf_error("stub: only the names attribute is supported");
    x->names = val;
    return val;
}
int Rf_asInteger(SEXP x) {
    if (x->len < 1) Rf_error("stub: asInteger of a zero-length object");
    return x->type == REALSXP ? (int)REAL(x)[0] : INTEGER(x)[0];
}
int Rf_asLogical(SEXP x) { return Rf_asInteger(x) != 0; }
double Rf_asReal(SEXP x) {
    if (x->len < 1) Rf_error("stub: asReal of a zero-length object");
    return x->type == REALSXP ? REAL(x)[0] : (double)INTEGER(x)[0];
}
SEXP Rf_protect(SEXP x) { ++g_depth; return x; }
void Rf_unprotect(int n) {
    g_depth -= n;
    if (g_depth < 0) { fprintf(stderr, "stub: UNPROTECT below zero\n"); abort(); }
}
void Rf_error(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    throw r_error(buf);  // R would longjmp to the top level; the protect stack is reset there
}
static const R_CallMethodDef* g_call = nullptr;
int R_registerRoutines(DllInfo*, const void*, const R_CallMethodDef* call_routines, const void*, const void*) {
    g_call = call_routines;
    return 1;
}

// ---- driver API --------------------------------------------------------------------------------------------
void splash_cuda_register(DllInfo* dll);  // the glue's registration hook (rglue.cpp)
SEXP stub_real(const double* v, R_xlen_t n, int nrow, int ncol) {
    SEXP s = nrow ? Rf_allocMatrix(REALSXP, nrow, ncol) : Rf_allocVector(REALSXP, n);
    memcpy(s->data, v, (size_t)n * 8);
    return s;
}
SEXP stub_int(const int* v, R_xlen_t n) {
    SEXP s = Rf_allocVector(INTSXP, n);
    memcpy(s->data, v, (size_t)n * 4);
    return s;
}
SEXP stub_lgl(int v) {
    SEXP s = Rf_allocVector(LGLSXP, 1);
    ((int*)s->data)[0] = v;
    return s;
}
const char* stub_name(SEXP list, R_xlen_t i) { return (list->names && i < list->names->len) ? (const char*)((SEXP*)list->names->data)[i]->data : ""; }
int stub_protect_depth(void) { return g_depth; }
const char* stub_last_error(void) { return g_err.c_str(); }
int stub_n_routines(void) {
    if (!g_call) splash_cuda_register(nullptr);
    int n = 0;
    while (g_call[n].name) ++n;
    return n;
}
const char* stub_routine_name(int i) { return g_call[i].name; }
int stub_routine_nargs(int i) { return g_call[i].numArgs; }
// .Call(name, args...) through the registration table; NULL + stub_last_error() if the routine raised an R error
SEXP stub_call(const char* name, int nargs, SEXP* a) {
    if (!g_call) splash_cuda_register(nullptr);
    g_err.clear();
    for (int i = 0; g_call[i].name; ++i) {
        if (strcmp(g_call[i].name, name)) continue;
        if (g_call[i].numArgs != nargs) { g_err = "wrong number of arguments"; return nullptr; }
        try {
            DL_FUNC f = g_call[i].fun;
            switch (nargs) {
                case 0: return ((SEXP(*)(void))f)();
                case 1: return ((SEXP(*)(SEXP))f)(a[0]);
                case 4: return ((SEXP(*)(SEXP, SEXP, SEXP, SEXP))f)(a[0], a[1], a[2], a[3]);
                case 16: return ((SEXP(*)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP))f)(
                    a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8], a[9], a[10], a[11], a[12], a[13], a[14], a[15]);
                case 15: return ((SEXP(*)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP))f)(
                    a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8], a[9], a[10], a[11], a[12], a[13], a[14]);
                default: g_err = "stub_call: unsupported arity"; return nullptr;
            }
        } catch (const r_error& e) {
            g_err = e.what();
            g_depth = 0;
            return nullptr;
        }
    }
    g_err = "no such routine";
    return nullptr;
}
int stub_type(SEXP s) { return s->type; }
R_xlen_t stub_len(SEXP s) { return s->len; }
int stub_nrow(SEXP s) { return s->nrow; }
int stub_ncol(SEXP s) { return s->ncol; }
void* stub_data(SEXP s) { return s->data; }
SEXP stub_elt(SEXP s, R_xlen_t i) { return VECTOR_ELT(s, i); }
}

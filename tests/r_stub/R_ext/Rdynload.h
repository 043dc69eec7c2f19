/* tests/r_stub -- TEST INFRASTRUCTURE (see Rinternals.h) */
#ifndef R_STUB_RDYNLOAD_H
#define R_STUB_RDYNLOAD_H
#ifdef __cplusplus
extern "C" {
#endif
typedef void* (*DL_FUNC)(void);
typedef struct { const char* name; DL_FUNC fun; int numArgs; } R_CallMethodDef;
typedef struct stub_dllinfo { const R_CallMethodDef* call; } DllInfo;
int R_registerRoutines(DllInfo* info, const void* c_routines, const R_CallMethodDef* call_routines, const void* fortran_routines,
                       const void* external_routines);
#ifdef __cplusplus
}
#endif
#endif

/* tests/mini_r/include/R_ext/Rdynload.h -- TEST INFRASTRUCTURE (see ../Rinternals.h). */
#ifndef MINI_RDYNLOAD_H
#define MINI_RDYNLOAD_H
#include <Rinternals.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef void *(*DL_FUNC)(void);
typedef struct {
    const char *name;
    DL_FUNC fun;
    int numArgs;
} R_CallMethodDef;
typedef R_CallMethodDef R_ExternalMethodDef;
typedef struct { const char *name; DL_FUNC fun; int numArgs; void *types; } R_CMethodDef;
typedef R_CMethodDef R_FortranMethodDef;
typedef struct mini_dllinfo {
    const R_CallMethodDef *call_methods;
    int n_call;
    int use_dynamic_symbols;
} DllInfo;
int R_registerRoutines(DllInfo *info, const R_CMethodDef *const c, const R_CallMethodDef *const call,
                       const R_FortranMethodDef *const f, const R_ExternalMethodDef *const ext);
Rboolean R_useDynamicSymbols(DllInfo *info, Rboolean value);
#ifdef __cplusplus
}
#endif
#endif

/* tests/mini_r/mini_r.c -- TEST INFRASTRUCTURE. Runtime behind the miniature R C API in include/ (see the header
 * comment of include/Rinternals.h). Objects live in an arena that `mr_reset()` frees; nothing is garbage collected,
 * but PROTECT/UNPROTECT are counted so that `mr_call` can report an unbalanced protection stack, which real R
 * reports as "stack imbalance in .Call". Rf_error longjmps back into `mr_call`, like R's error -> top level. */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>

#include <setjmp.h>
#include <stdarg.h>
#include <stdint.h>

struct mini_sexp {
    SEXPTYPE type;
    R_xlen_t length;
    void *data;      /* double[], int[], SEXP[], or char[] for CHARSXP */
    SEXP names;      /* names attribute (STRSXP) or NULL */
    SEXP dim;        /* dim attribute (INTSXP of length 2) or NULL */
};

static struct mini_sexp nil_obj = {NILSXP, 0, NULL, NULL, NULL};
static struct mini_sexp names_sym = {NILSXP, 0, NULL, NULL, NULL};
static struct mini_sexp dim_sym = {NILSXP, 0, NULL, NULL, NULL};
SEXP R_NilValue = &nil_obj;
SEXP R_NamesSymbol = &names_sym;
SEXP R_DimSymbol = &dim_sym;
double R_NaReal, R_NaN;
int R_NaInt = INT32_MIN;

/* ---- arena ------------------------------------------------------------------------------------------------ */
typedef struct blk { struct blk *next; } blk;
static blk *arena = NULL;
static void *arena_alloc(size_t n)
{
    blk *b = (blk *)calloc(1, sizeof(blk) + 16 + n);
    if (!b) { fprintf(stderr, "mini_r: out of memory\n"); abort(); }
    b->next = arena;
    arena = b;
    return (char *)b + ((sizeof(blk) + 15) & ~(size_t)15);
}
void mr_reset(void)
{
    while (arena) { blk *n = arena->next; free(arena); arena = n; }
}
char *R_alloc(size_t n, int size) { return (char *)arena_alloc(n * (size_t)size + 1); }

__attribute__((constructor)) static void mr_init(void)
{
    union { double d; uint64_t u; } na;
    na.u = 0x7FF00000000007A2ull;   /* payload 1954 */
    R_NaReal = na.d;
    R_NaN = NAN;
}
int R_IsNA(double x)
{
    union { double d; uint64_t u; } v;
    v.d = x;
    return x != x && (uint32_t)(v.u & 0xFFFFFFFFu) == 1954u;
}
int R_IsNaN(double x) { return x != x && !R_IsNA(x); }

/* ---- errors, protection ----------------------------------------------------------------------------------- */
static jmp_buf *err_jmp = NULL;
static char err_msg[1024];
static int protect_depth = 0;

void Rf_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(err_msg, sizeof err_msg, fmt, ap);
    va_end(ap);
    if (!err_jmp) { fprintf(stderr, "mini_r: Rf_error outside mr_call: %s\n", err_msg); abort(); }
    longjmp(*err_jmp, 1);
}
void Rf_warning(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    fprintf(stderr, "Warning: ");
    vfprintf(stderr, fmt, ap);
    fprintf(stderr, "\n");
    va_end(ap);
}
void Rprintf(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vfprintf(stdout, fmt, ap);
    va_end(ap);
}
void REprintf(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vfprintf(stderr, fmt, ap);
    va_end(ap);
    fflush(stderr);
}
SEXP Rf_protect(SEXP s) { protect_depth++; return s; }
void Rf_unprotect(int n)
{
    protect_depth -= n;
    if (protect_depth < 0) Rf_error("unprotect(): only %d protected items", protect_depth + n);
}

/* ---- objects ---------------------------------------------------------------------------------------------- */
static size_t elt_size(SEXPTYPE t)
{
    switch (t) {
    case REALSXP: return sizeof(double);
    case INTSXP: case LGLSXP: return sizeof(int);
    case STRSXP: case VECSXP: return sizeof(SEXP);
    case CHARSXP: return 1;
    default: Rf_error("mini_r: unsupported SEXPTYPE %u", t);
    }
}
SEXP Rf_allocVector(SEXPTYPE t, R_xlen_t n)
{
    if (n < 0) Rf_error("negative length vectors are not allowed");
    SEXP s = (SEXP)arena_alloc(sizeof(struct mini_sexp));
    s->type = t;
    s->length = n;
    s->data = arena_alloc(elt_size(t) * (size_t)(n + (t == CHARSXP)));
    if (t == STRSXP || t == VECSXP)
        for (R_xlen_t i = 0; i < n; i++) ((SEXP *)s->data)[i] = R_NilValue;
    return s;
}
SEXP Rf_allocMatrix(SEXPTYPE t, int nrow, int ncol)
{
    SEXP s = Rf_allocVector(t, (R_xlen_t)nrow * ncol);
    s->dim = Rf_allocVector(INTSXP, 2);
    INTEGER(s->dim)[0] = nrow;
    INTEGER(s->dim)[1] = ncol;
    return s;
}
int TYPEOF(SEXP s) { return (int)s->type; }
R_xlen_t Rf_xlength(SEXP s) { return s->length; }
int Rf_length(SEXP s) { return (int)s->length; }
int Rf_nrows(SEXP s)
{
    if (s->dim) return INTEGER(s->dim)[0];
    if (s->type == NILSXP) Rf_error("object is not a matrix");
    return (int)s->length;              /* a plain vector counts as one column */
}
int Rf_ncols(SEXP s)
{
    if (s->dim) return INTEGER(s->dim)[1];
    if (s->type == NILSXP) Rf_error("object is not a matrix");
    return 1;
}
static void need(SEXP s, SEXPTYPE t, const char *who)
{
    if (s->type != t) Rf_error("%s() can only be applied to a '%s', not type %u", who,
                               t == REALSXP ? "numeric" : t == INTSXP ? "integer" : t == LGLSXP ? "logical"
                               : t == STRSXP ? "character" : t == VECSXP ? "list" : "CHARSXP", s->type);
}
double *REAL(SEXP s) { need(s, REALSXP, "REAL"); return (double *)s->data; }
int *INTEGER(SEXP s)
{
    if (s->type != INTSXP && s->type != LGLSXP) need(s, INTSXP, "INTEGER");
    return (int *)s->data;
}
int *LOGICAL(SEXP s) { need(s, LGLSXP, "LOGICAL"); return (int *)s->data; }
const char *CHAR(SEXP s) { need(s, CHARSXP, "CHAR"); return (const char *)s->data; }
SEXP STRING_ELT(SEXP s, R_xlen_t i)
{
    need(s, STRSXP, "STRING_ELT");
    if (i < 0 || i >= s->length) Rf_error("attempt to access index %ld/%ld in STRING_ELT", (long)i, (long)s->length);
    return ((SEXP *)s->data)[i];
}
void SET_STRING_ELT(SEXP s, R_xlen_t i, SEXP v)
{
    need(s, STRSXP, "SET_STRING_ELT");
    need(v, CHARSXP, "SET_STRING_ELT value");
    if (i < 0 || i >= s->length) Rf_error("attempt to set index %ld/%ld in SET_STRING_ELT", (long)i, (long)s->length);
    ((SEXP *)s->data)[i] = v;
}
SEXP VECTOR_ELT(SEXP s, R_xlen_t i)
{
    need(s, VECSXP, "VECTOR_ELT");
    if (i < 0 || i >= s->length) Rf_error("attempt to access index %ld/%ld in VECTOR_ELT", (long)i, (long)s->length);
    return ((SEXP *)s->data)[i];
}
SEXP SET_VECTOR_ELT(SEXP s, R_xlen_t i, SEXP v)
{
    need(s, VECSXP, "SET_VECTOR_ELT");
    if (i < 0 || i >= s->length) Rf_error("attempt to set index %ld/%ld in SET_VECTOR_ELT", (long)i, (long)s->length);
    ((SEXP *)s->data)[i] = v;
    return v;
}
SEXP Rf_mkChar(const char *c)
{
    SEXP s = Rf_allocVector(CHARSXP, (R_xlen_t)strlen(c));
    memcpy(s->data, c, strlen(c) + 1);
    return s;
}
SEXP Rf_mkString(const char *c)
{
    SEXP s = Rf_allocVector(STRSXP, 1);
    SET_STRING_ELT(s, 0, Rf_mkChar(c));
    return s;
}
SEXP Rf_ScalarReal(double x) { SEXP s = Rf_allocVector(REALSXP, 1); REAL(s)[0] = x; return s; }
SEXP Rf_ScalarInteger(int x) { SEXP s = Rf_allocVector(INTSXP, 1); INTEGER(s)[0] = x; return s; }
SEXP Rf_ScalarLogical(int x) { SEXP s = Rf_allocVector(LGLSXP, 1); LOGICAL(s)[0] = x; return s; }

SEXP Rf_getAttrib(SEXP s, SEXP name)
{
    if (name == R_NamesSymbol) return s->names ? s->names : R_NilValue;
    if (name == R_DimSymbol) return s->dim ? s->dim : R_NilValue;
    return R_NilValue;
}
SEXP Rf_setAttrib(SEXP s, SEXP name, SEXP val)
{
    if (name == R_NamesSymbol) {
        if (val != R_NilValue && (val->type != STRSXP || val->length != s->length))
            Rf_error("'names' attribute [%ld] must be the same length as the vector [%ld]", (long)val->length, (long)s->length);
        s->names = val == R_NilValue ? NULL : val;
    } else if (name == R_DimSymbol) {
        s->dim = val == R_NilValue ? NULL : val;
    } else {
        Rf_error("mini_r: only names and dim attributes are modelled");
    }
    return val;
}

static double int_to_real(int v) { return v == R_NaInt ? R_NaReal : (double)v; }
static int real_to_int(double v) { return (v != v || v >= 2147483648.0 || v <= -2147483649.0) ? R_NaInt : (int)v; }

SEXP Rf_coerceVector(SEXP s, SEXPTYPE t)
{
    if (s->type == t) return s;                        /* R returns the object itself */
    if (s->type != REALSXP && s->type != INTSXP && s->type != LGLSXP)
        Rf_error("cannot coerce type %u to vector of type %u", s->type, t);
    if (t != REALSXP && t != INTSXP && t != LGLSXP) Rf_error("cannot coerce type %u to vector of type %u", s->type, t);
    SEXP r = Rf_allocVector(t, s->length);
    for (R_xlen_t i = 0; i < s->length; i++) {
        if (t == REALSXP) REAL(r)[i] = int_to_real(((int *)s->data)[i]);
        else if (s->type == REALSXP) ((int *)r->data)[i] = t == LGLSXP
                ? (REAL(s)[i] != REAL(s)[i] ? R_NaInt : REAL(s)[i] != 0.0) : real_to_int(REAL(s)[i]);
        else ((int *)r->data)[i] = (t == LGLSXP && ((int *)s->data)[i] != R_NaInt) ? ((int *)s->data)[i] != 0
                                                                                   : ((int *)s->data)[i];
    }
    r->names = s->names;                               /* coerceVector keeps dim and names */
    r->dim = s->dim;
    return r;
}
SEXP Rf_duplicate(SEXP s)
{
    if (s == R_NilValue || s->type == CHARSXP) return s;
    SEXP r = Rf_allocVector(s->type, s->length);
    if (s->type == VECSXP)
        for (R_xlen_t i = 0; i < s->length; i++) ((SEXP *)r->data)[i] = Rf_duplicate(((SEXP *)s->data)[i]);
    else
        memcpy(r->data, s->data, elt_size(s->type) * (size_t)s->length);
    r->names = s->names ? Rf_duplicate(s->names) : NULL;
    r->dim = s->dim ? Rf_duplicate(s->dim) : NULL;
    return r;
}
SEXP Rf_lengthgets(SEXP s, int n)
{
    SEXP r = Rf_allocVector(s->type, n);
    for (R_xlen_t i = 0; i < n; i++) {
        if (i < s->length) memcpy((char *)r->data + elt_size(s->type) * i, (char *)s->data + elt_size(s->type) * i, elt_size(s->type));
        else if (s->type == REALSXP) REAL(r)[i] = R_NaReal;
        else if (s->type == INTSXP || s->type == LGLSXP) ((int *)r->data)[i] = R_NaInt;
    }
    return r;
}
double Rf_asReal(SEXP s)
{
    if (s->length < 1) return R_NaReal;
    if (s->type == REALSXP) return REAL(s)[0];
    if (s->type == INTSXP || s->type == LGLSXP) return int_to_real(((int *)s->data)[0]);
    return R_NaReal;
}
int Rf_asInteger(SEXP s)
{
    if (s->length < 1) return R_NaInt;
    if (s->type == REALSXP) return real_to_int(REAL(s)[0]);
    if (s->type == INTSXP || s->type == LGLSXP) return ((int *)s->data)[0];
    return R_NaInt;
}
int Rf_asLogical(SEXP s)
{
    if (s->length < 1) return R_NaInt;
    if (s->type == REALSXP) return REAL(s)[0] != REAL(s)[0] ? R_NaInt : REAL(s)[0] != 0.0;
    if (s->type == INTSXP || s->type == LGLSXP) return ((int *)s->data)[0] == R_NaInt ? R_NaInt : ((int *)s->data)[0] != 0;
    return R_NaInt;
}

/* ---- dynamic loading -------------------------------------------------------------------------------------- */
int R_registerRoutines(DllInfo *info, const R_CMethodDef *const c, const R_CallMethodDef *const call,
                       const R_FortranMethodDef *const f, const R_ExternalMethodDef *const ext)
{
    (void)c; (void)f; (void)ext;
    info->call_methods = call;
    info->n_call = 0;
    while (call && call[info->n_call].name) info->n_call++;
    return 1;
}
Rboolean R_useDynamicSymbols(DllInfo *info, Rboolean value)
{
    Rboolean old = info->use_dynamic_symbols ? TRUE : FALSE;
    info->use_dynamic_symbols = value;
    return old;
}

/* ---- host-side helpers for the Python test driver ----------------------------------------------------------- */
static DllInfo the_dll = {NULL, 0, 1};
DllInfo *mr_dll(void) { return &the_dll; }
int mr_n_routines(void) { return the_dll.n_call; }
const char *mr_routine_name(int i) { return the_dll.call_methods[i].name; }
int mr_routine_nargs(int i) { return the_dll.call_methods[i].numArgs; }
int mr_dynamic_symbols(void) { return the_dll.use_dynamic_symbols; }
const char *mr_last_error(void) { return err_msg; }

typedef SEXP (*f1)(SEXP);
#define A(i) a[i]
/* .Call(name, ...): registered routines only, arity checked as R does. Returns NULL on an R error. */
SEXP mr_call(const char *name, int nargs, SEXP *a)
{
    const R_CallMethodDef *volatile m = NULL;
    for (int i = 0; i < the_dll.n_call; i++)
        if (!strcmp(the_dll.call_methods[i].name, name)) m = &the_dll.call_methods[i];
    err_msg[0] = 0;
    if (!m) { snprintf(err_msg, sizeof err_msg, "\"%s\" not available for .Call() for package \"sparseRGPs\"", name); return NULL; }
    if (m->numArgs != nargs) {
        snprintf(err_msg, sizeof err_msg, "Incorrect number of arguments (%d), expecting %d for '%s'", nargs, m->numArgs, name);
        return NULL;
    }
    jmp_buf jb;
    SEXP volatile r = NULL;
    const int depth0 = protect_depth;
    err_jmp = &jb;
    if (setjmp(jb) == 0) {
        DL_FUNC f = m->fun;
        switch (nargs) {
#define CALLN(N, ...) case N: r = ((SEXP(*)())f)(__VA_ARGS__); break;
        CALLN(1, A(0))
        CALLN(2, A(0), A(1))
        CALLN(3, A(0), A(1), A(2))
        CALLN(4, A(0), A(1), A(2), A(3))
        CALLN(5, A(0), A(1), A(2), A(3), A(4))
        CALLN(6, A(0), A(1), A(2), A(3), A(4), A(5))
        CALLN(7, A(0), A(1), A(2), A(3), A(4), A(5), A(6))
        CALLN(8, A(0), A(1), A(2), A(3), A(4), A(5), A(6), A(7))
        CALLN(9, A(0), A(1), A(2), A(3), A(4), A(5), A(6), A(7), A(8))
        CALLN(10, A(0), A(1), A(2), A(3), A(4), A(5), A(6), A(7), A(8), A(9))
        CALLN(11, A(0), A(1), A(2), A(3), A(4), A(5), A(6), A(7), A(8), A(9), A(10))
        CALLN(12, A(0), A(1), A(2), A(3), A(4), A(5), A(6), A(7), A(8), A(9), A(10), A(11))
        CALLN(13, A(0), A(1), A(2), A(3), A(4), A(5), A(6), A(7), A(8), A(9), A(10), A(11), A(12))
        CALLN(14, A(0), A(1), A(2), A(3), A(4), A(5), A(6), A(7), A(8), A(9), A(10), A(11), A(12), A(13))
        CALLN(15, A(0), A(1), A(2), A(3), A(4), A(5), A(6), A(7), A(8), A(9), A(10), A(11), A(12), A(13), A(14))
        CALLN(16, A(0), A(1), A(2), A(3), A(4), A(5), A(6), A(7), A(8), A(9), A(10), A(11), A(12), A(13), A(14), A(15))
        CALLN(17, A(0), A(1), A(2), A(3), A(4), A(5), A(6), A(7), A(8), A(9), A(10), A(11), A(12), A(13), A(14), A(15), A(16))
        CALLN(18, A(0), A(1), A(2), A(3), A(4), A(5), A(6), A(7), A(8), A(9), A(10), A(11), A(12), A(13), A(14), A(15), A(16), A(17))
        default: snprintf(err_msg, sizeof err_msg, "mini_r: %d arguments not supported", nargs); r = NULL;
        }
        err_jmp = NULL;
        if (r && protect_depth != depth0) {
            snprintf(err_msg, sizeof err_msg, "stack imbalance in '.Call', %d then %d", depth0, protect_depth);
            protect_depth = depth0;
            return NULL;
        }
        return r;
    }
    err_jmp = NULL;                 /* arrived here through Rf_error: R unwinds the protection stack itself */
    protect_depth = depth0;
    return NULL;
}

/* constructors / accessors with plain C types (ctypes-friendly) */
SEXP mr_real(const double *v, R_xlen_t n, int nrow, int ncol)
{
    SEXP s = nrow >= 0 ? Rf_allocMatrix(REALSXP, nrow, ncol) : Rf_allocVector(REALSXP, n);
    if (n) memcpy(REAL(s), v, sizeof(double) * (size_t)n);
    return s;
}
SEXP mr_int(const int *v, R_xlen_t n, int nrow, int ncol, int logical)
{
    SEXP s = nrow >= 0 ? Rf_allocMatrix(logical ? LGLSXP : INTSXP, nrow, ncol) : Rf_allocVector(logical ? LGLSXP : INTSXP, n);
    if (n) memcpy(s->data, v, sizeof(int) * (size_t)n);
    return s;
}
SEXP mr_na_matrix(void)            /* R's `matrix()`: 1 x 1 logical NA */
{
    SEXP s = Rf_allocMatrix(LGLSXP, 1, 1);
    LOGICAL(s)[0] = R_NaInt;
    return s;
}
SEXP mr_strings(const char *const *v, int n)
{
    SEXP s = Rf_allocVector(STRSXP, n);
    for (int i = 0; i < n; i++) SET_STRING_ELT(s, i, Rf_mkChar(v[i]));
    return s;
}
SEXP mr_list(const char *const *names, SEXP *vals, int n)
{
    SEXP s = Rf_allocVector(VECSXP, n);
    for (int i = 0; i < n; i++) SET_VECTOR_ELT(s, i, vals[i]);
    if (names) Rf_setAttrib(s, R_NamesSymbol, mr_strings(names, n));
    return s;
}
SEXP mr_nil(void) { return R_NilValue; }
int mr_type(SEXP s) { return (int)s->type; }
R_xlen_t mr_len(SEXP s) { return s->length; }
int mr_has_dim(SEXP s) { return s->dim != NULL; }
int mr_dim(SEXP s, int i) { return INTEGER(s->dim)[i]; }
int mr_has_names(SEXP s) { return s->names != NULL; }
const char *mr_name(SEXP s, int i) { return CHAR(STRING_ELT(s->names, i)); }
const void *mr_data(SEXP s) { return s->data; }
SEXP mr_elt(SEXP s, int i) { return VECTOR_ELT(s, i); }
const char *mr_string(SEXP s, int i) { return CHAR(STRING_ELT(s, i)); }

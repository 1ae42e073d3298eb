/* TEST INFRASTRUCTURE ONLY -- the runtime behind oracle/rstub/Rinternals.h plus the harness entry points tests/test_rlayer.py
 * drives through ctypes: object constructors / accessors and rstub_call, which runs one of the reference's .Call entry points
 * (registered in /root/reference/src/init.c:46-82) and catches error() like R's top level does.  Never linked into the product. */
#include "Rinternals.h"
#include <setjmp.h>
#include <stdarg.h>
#include <stdlib.h>

struct rstub_attr { SEXP name, value; struct rstub_attr* next; };
struct rstub_sexp {
    int type;
    R_xlen_t length, truelength;
    void* data;                      /* payload: bytes / ints / doubles / SEXP array / C string / external address */
    struct rstub_attr* attrib;
    R_CFinalizer_t finalizer;
};

static struct rstub_sexp nil_obj = {NILSXP, 0, 0, NULL, NULL, NULL};
SEXP R_NilValue = &nil_obj;
static struct rstub_sexp names_sym = {SYMSXP, 0, 0, (void*)"names", NULL, NULL}, class_sym = {SYMSXP, 0, 0, (void*)"class", NULL, NULL};
SEXP R_NamesSymbol = &names_sym, R_ClassSymbol = &class_sym;

static jmp_buf* g_top = NULL;
static char g_error[1024], g_warning[1024], g_printed[4096];
static int g_nwarn = 0;

static SEXP new_obj(int type, R_xlen_t n, size_t elem)
{
    SEXP s = (SEXP)calloc(1, sizeof(struct rstub_sexp));
    s->type = type; s->length = n; s->truelength = n;
    s->data = calloc((size_t)(n > 0 ? n : 1) + 1, elem ? elem : 1);
    return s;
}
int TYPEOF(SEXP x) { return x->type; }
R_xlen_t Rf_xlength(SEXP x) { return x->length; }
int Rf_length(SEXP x) { return (int)x->length; }
SEXP Rf_allocVector(SEXPTYPE type, R_xlen_t n)
{
    switch (type) {
    case RAWSXP: return new_obj(RAWSXP, n, 1);
    case LGLSXP: case INTSXP: return new_obj((int)type, n, sizeof(int));
    case REALSXP: return new_obj(REALSXP, n, sizeof(double));
    case STRSXP: { SEXP s = new_obj(STRSXP, n, sizeof(SEXP)); for (R_xlen_t i = 0; i < n; i++) ((SEXP*)s->data)[i] = Rf_mkChar(""); return s; }
    case VECSXP: { SEXP s = new_obj(VECSXP, n, sizeof(SEXP)); for (R_xlen_t i = 0; i < n; i++) ((SEXP*)s->data)[i] = R_NilValue; return s; }
    default: Rf_error("rstub: allocVector of type %u is not provided", type);
    }
}
SEXP Rf_protect(SEXP x) { return x; }
void Rf_unprotect(int n) { (void)n; }
Rbyte* RAW(SEXP x) { return (Rbyte*)x->data; }
int* INTEGER(SEXP x) { return (int*)x->data; }
int* LOGICAL(SEXP x) { return (int*)x->data; }
double* REAL(SEXP x) { return (double*)x->data; }
SEXP STRING_ELT(SEXP x, R_xlen_t i) { return ((SEXP*)x->data)[i]; }
void SET_STRING_ELT(SEXP x, R_xlen_t i, SEXP v) { ((SEXP*)x->data)[i] = v; }
SEXP VECTOR_ELT(SEXP x, R_xlen_t i) { return ((SEXP*)x->data)[i]; }
SEXP SET_VECTOR_ELT(SEXP x, R_xlen_t i, SEXP v) { ((SEXP*)x->data)[i] = v; return v; }
const char* CHAR(SEXP x) { return (const char*)x->data; }
SEXP Rf_mkChar(const char* s)
{
    const size_t n = strlen(s);
    SEXP c = new_obj(CHARSXP, (R_xlen_t)n, 1);
    memcpy(c->data, s, n);
    return c;
}
SEXP Rf_mkString(const char* s) { SEXP v = Rf_allocVector(STRSXP, 1); SET_STRING_ELT(v, 0, Rf_mkChar(s)); return v; }
SEXP Rf_ScalarInteger(int v) { SEXP s = Rf_allocVector(INTSXP, 1); INTEGER(s)[0] = v; return s; }
SEXP Rf_ScalarLogical(int v) { SEXP s = Rf_allocVector(LGLSXP, 1); LOGICAL(s)[0] = v ? 1 : 0; return s; }
SEXP Rf_ScalarReal(double v) { SEXP s = Rf_allocVector(REALSXP, 1); REAL(s)[0] = v; return s; }
int Rf_asInteger(SEXP x)
{
    if (x->length < 1) return NA_INTEGER;
    if (x->type == INTSXP || x->type == LGLSXP) return INTEGER(x)[0];
    if (x->type == REALSXP) return (int)REAL(x)[0];
    return NA_INTEGER;
}
int Rf_asLogical(SEXP x)
{
    if (x->length < 1) return NA_LOGICAL;
    if (x->type == LGLSXP || x->type == INTSXP) return INTEGER(x)[0] == NA_INTEGER ? NA_LOGICAL : INTEGER(x)[0] != 0;
    if (x->type == REALSXP) return REAL(x)[0] != 0.0;
    return NA_LOGICAL;
}
double Rf_asReal(SEXP x)
{
    if (x->length < 1) return 0.0;
    if (x->type == REALSXP) return REAL(x)[0];
    if (x->type == INTSXP || x->type == LGLSXP) return (double)INTEGER(x)[0];
    return 0.0;
}
Rboolean Rf_isNull(SEXP x) { return x->type == NILSXP; }
Rboolean Rf_isNewList(SEXP x) { return x->type == NILSXP || x->type == VECSXP; }
Rboolean Rf_isString(SEXP x) { return x->type == STRSXP; }
SEXP Rf_install(const char* name)
{
    if (!strcmp(name, "names")) return R_NamesSymbol;
    if (!strcmp(name, "class")) return R_ClassSymbol;
    SEXP s = new_obj(SYMSXP, 0, 1);
    free(s->data); s->data = strdup(name);
    return s;
}
SEXP Rf_setAttrib(SEXP x, SEXP name, SEXP value)
{
    for (struct rstub_attr* a = x->attrib; a; a = a->next) if (!strcmp((const char*)a->name->data, (const char*)name->data)) { a->value = value; return value; }
    struct rstub_attr* a = (struct rstub_attr*)calloc(1, sizeof(*a));
    a->name = name; a->value = value; a->next = x->attrib; x->attrib = a;
    return value;
}
SEXP Rf_getAttrib(SEXP x, SEXP name)
{
    for (struct rstub_attr* a = x->attrib; a; a = a->next) if (!strcmp((const char*)a->name->data, (const char*)name->data)) return a->value;
    return R_NilValue;
}
Rboolean Rf_inherits(SEXP x, const char* what)
{
    SEXP k = Rf_getAttrib(x, R_ClassSymbol);
    if (k->type != STRSXP) return 0;
    for (R_xlen_t i = 0; i < k->length; i++) if (!strcmp(CHAR(STRING_ELT(k, i)), what)) return 1;
    return 0;
}
void SETLENGTH(SEXP x, R_xlen_t n) { x->length = n; }
void SET_TRUELENGTH(SEXP x, R_xlen_t n) { x->truelength = n; }
void SET_GROWABLE_BIT(SEXP x) { (void)x; }
SEXP R_MakeExternalPtr(void* p, SEXP tag, SEXP prot) { (void)tag; (void)prot; SEXP s = new_obj(EXTPTRSXP, 0, 1); free(s->data); s->data = p; return s; }
void* R_ExternalPtrAddr(SEXP s) { return s->type == EXTPTRSXP ? s->data : NULL; }
void R_ClearExternalPtr(SEXP s) { s->data = NULL; }
void R_RegisterCFinalizer(SEXP s, R_CFinalizer_t fun) { s->finalizer = fun; }
void R_RegisterCFinalizerEx(SEXP s, R_CFinalizer_t fun, Rboolean onexit) { (void)onexit; s->finalizer = fun; }
void Rf_error(const char* fmt, ...)
{
    va_list ap; va_start(ap, fmt); vsnprintf(g_error, sizeof(g_error), fmt, ap); va_end(ap);
    if (g_top) longjmp(*g_top, 1);
    fprintf(stderr, "rstub: error() outside rstub_call: %s\n", g_error);
    abort();
}
void Rf_warning(const char* fmt, ...) { va_list ap; va_start(ap, fmt); vsnprintf(g_warning, sizeof(g_warning), fmt, ap); va_end(ap); g_nwarn++; }
void Rprintf(const char* fmt, ...) { va_list ap; va_start(ap, fmt); const size_t n = strlen(g_printed); vsnprintf(g_printed + n, sizeof(g_printed) - n, fmt, ap); va_end(ap); }
void REprintf(const char* fmt, ...) { va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap); }
char* R_alloc(size_t n, int size) { return (char*)calloc(n ? n : 1, (size_t)(size > 0 ? size : 1)); }
void R_CheckUserInterrupt(void) {}

/* ---- harness entry points (ctypes) ------------------------------------------------------------------------------------ */
SEXP rstub_nil(void) { return R_NilValue; }
SEXP rstub_raw(const void* p, size_t n) { SEXP s = Rf_allocVector(RAWSXP, (R_xlen_t)n); if (n) memcpy(RAW(s), p, n); return s; }
SEXP rstub_str(const char* s) { return Rf_mkString(s); }
SEXP rstub_int(int v) { return Rf_ScalarInteger(v); }
SEXP rstub_lgl(int v) { return Rf_ScalarLogical(v); }
SEXP rstub_real(double v) { return Rf_ScalarReal(v); }
SEXP rstub_list(int n) { SEXP l = Rf_allocVector(VECSXP, n); Rf_setAttrib(l, R_NamesSymbol, Rf_allocVector(STRSXP, n)); return l; }
void rstub_list_set(SEXP l, int i, const char* name, SEXP v) { SET_VECTOR_ELT(l, i, v); SET_STRING_ELT(Rf_getAttrib(l, R_NamesSymbol), i, Rf_mkChar(name ? name : "")); }
int rstub_type(SEXP s) { return s->type; }
size_t rstub_len(SEXP s) { return (size_t)s->length; }
const void* rstub_data(SEXP s) { return s->type == STRSXP ? (const void*)CHAR(STRING_ELT(s, 0)) : s->data; }
SEXP rstub_elt(SEXP l, int i) { return VECTOR_ELT(l, i); }
const char* rstub_name(SEXP l, int i) { SEXP n = Rf_getAttrib(l, R_NamesSymbol); return n->type == STRSXP && i < n->length ? CHAR(STRING_ELT(n, i)) : ""; }
const char* rstub_class(SEXP s) { SEXP k = Rf_getAttrib(s, R_ClassSymbol); return k->type == STRSXP && k->length ? CHAR(STRING_ELT(k, 0)) : ""; }
const char* rstub_last_error(void) { return g_error; }
const char* rstub_last_warning(void) { return g_warning; }
int rstub_warning_count(void) { return g_nwarn; }
const char* rstub_printed(void) { return g_printed; }
void rstub_reset_messages(void) { g_error[0] = g_warning[0] = g_printed[0] = 0; g_nwarn = 0; }
void rstub_finalize(SEXP s) { if (s->type == EXTPTRSXP && s->finalizer) { R_CFinalizer_t f = s->finalizer; s->finalizer = NULL; f(s); } }   /* what the collector does eventually */
/* run fn(a[0..n)) under an error handler: returns the result, or NULL after error() (message: rstub_last_error) */
SEXP rstub_call(void* fn, int n, SEXP* a)
{
    jmp_buf top; jmp_buf* prev = g_top;
    SEXP r = NULL;
    g_error[0] = 0;
    g_top = &top;
    if (!setjmp(top)) {
        switch (n) {
        case 0: r = ((SEXP (*)(void))fn)(); break;
        case 1: r = ((SEXP (*)(SEXP))fn)(a[0]); break;
        case 2: r = ((SEXP (*)(SEXP, SEXP))fn)(a[0], a[1]); break;
        case 3: r = ((SEXP (*)(SEXP, SEXP, SEXP))fn)(a[0], a[1], a[2]); break;
        case 4: r = ((SEXP (*)(SEXP, SEXP, SEXP, SEXP))fn)(a[0], a[1], a[2], a[3]); break;
        case 5: r = ((SEXP (*)(SEXP, SEXP, SEXP, SEXP, SEXP))fn)(a[0], a[1], a[2], a[3], a[4]); break;
        case 6: r = ((SEXP (*)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP))fn)(a[0], a[1], a[2], a[3], a[4], a[5]); break;
        default: snprintf(g_error, sizeof(g_error), "rstub_call: %d arguments", n);
        }
    }
    g_top = prev;
    return r;
}

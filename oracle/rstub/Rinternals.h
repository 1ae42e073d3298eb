/* TEST INFRASTRUCTURE ONLY -- a miniature of R's C API, just large enough to compile the reference's package C layer
 * (/root/reference/src/{cctx,dctx,raw-file,raw-file-in,raw-file-out,dictionaries,zstd-info,utils}.c) UNMODIFIED outside R
 * (R, Rscript and R's headers are absent from the image).  Objects are plain heap structs that are never
 * collected; error() unwinds to the harness with longjmp, as R's does.  Written from R's documented API ("Writing R
 * Extensions", section 5); nothing here comes from R's sources.  See oracle/rstub/rstub.c and tests/test_rlayer.py. */
#ifndef ZL_RSTUB_RINTERNALS_H
#define ZL_RSTUB_RINTERNALS_H
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef ptrdiff_t R_xlen_t;
typedef unsigned char Rbyte;
typedef int Rboolean;
#ifndef TRUE
#define TRUE 1
#define FALSE 0
#endif
typedef struct rstub_sexp* SEXP;
typedef unsigned int SEXPTYPE;

#define NILSXP 0
#define SYMSXP 1
#define CHARSXP 9
#define LGLSXP 10
#define INTSXP 13
#define REALSXP 14
#define STRSXP 16
#define VECSXP 19
#define EXTPTRSXP 22
#define RAWSXP 24

#define NA_INTEGER (-2147483647 - 1)
#define NA_LOGICAL (-2147483647 - 1)

extern SEXP R_NilValue, R_NamesSymbol, R_ClassSymbol;

int TYPEOF(SEXP x);
R_xlen_t Rf_xlength(SEXP x);
int Rf_length(SEXP x);
SEXP Rf_allocVector(SEXPTYPE type, R_xlen_t n);
SEXP Rf_protect(SEXP x);
void Rf_unprotect(int n);
Rbyte* RAW(SEXP x);
int* INTEGER(SEXP x);
int* LOGICAL(SEXP x);
double* REAL(SEXP x);
SEXP STRING_ELT(SEXP x, R_xlen_t i);
void SET_STRING_ELT(SEXP x, R_xlen_t i, SEXP v);
SEXP VECTOR_ELT(SEXP x, R_xlen_t i);
SEXP SET_VECTOR_ELT(SEXP x, R_xlen_t i, SEXP v);
const char* CHAR(SEXP x);
SEXP Rf_mkChar(const char* s);
SEXP Rf_mkString(const char* s);
SEXP Rf_ScalarInteger(int v);
SEXP Rf_ScalarLogical(int v);
SEXP Rf_ScalarReal(double v);
int Rf_asInteger(SEXP x);
int Rf_asLogical(SEXP x);
double Rf_asReal(SEXP x);
Rboolean Rf_isNull(SEXP x);
Rboolean Rf_isNewList(SEXP x);
Rboolean Rf_isString(SEXP x);
Rboolean Rf_inherits(SEXP x, const char* what);
SEXP Rf_install(const char* name);
SEXP Rf_setAttrib(SEXP x, SEXP name, SEXP value);
SEXP Rf_getAttrib(SEXP x, SEXP name);
void SETLENGTH(SEXP x, R_xlen_t n);
void SET_TRUELENGTH(SEXP x, R_xlen_t n);
void SET_GROWABLE_BIT(SEXP x);
SEXP R_MakeExternalPtr(void* p, SEXP tag, SEXP prot);
void* R_ExternalPtrAddr(SEXP s);
void R_ClearExternalPtr(SEXP s);
typedef void (*R_CFinalizer_t)(SEXP);
void R_RegisterCFinalizer(SEXP s, R_CFinalizer_t fun);
void R_RegisterCFinalizerEx(SEXP s, R_CFinalizer_t fun, Rboolean onexit);
void Rf_error(const char* fmt, ...) __attribute__((noreturn, format(printf, 1, 2)));
void Rf_warning(const char* fmt, ...) __attribute__((format(printf, 1, 2)));
void Rprintf(const char* fmt, ...) __attribute__((format(printf, 1, 2)));
void REprintf(const char* fmt, ...) __attribute__((format(printf, 1, 2)));
char* R_alloc(size_t n, int size);
void R_CheckUserInterrupt(void);

/* serialization stream types: only named by prototypes in buffer-static.h / calc-size-robust.h (never called here) */
typedef struct R_outpstream_st* R_outpstream_t;
typedef struct R_inpstream_st* R_inpstream_t;
typedef void* R_pstream_data_t;
typedef enum { R_pstream_any_format, R_pstream_ascii_format, R_pstream_binary_format, R_pstream_xdr_format, R_pstream_asciihex_format } R_pstream_format_t;
struct R_outpstream_st { R_pstream_data_t data; R_pstream_format_t type; int version;
    void (*OutChar)(R_outpstream_t, int); void (*OutBytes)(R_outpstream_t, void*, int); SEXP (*OutPersistHookFunc)(SEXP, SEXP); SEXP OutPersistHookData; };
struct R_inpstream_st { R_pstream_data_t data; R_pstream_format_t type;
    int (*InChar)(R_inpstream_t); void (*InBytes)(R_inpstream_t, void*, int); SEXP (*InPersistHookFunc)(SEXP, SEXP); SEXP InPersistHookData;
    char native_encoding[64]; void* nat2nat_obj; void* nat2utf8_obj; };

/* the unprefixed names the package sources use */
#define length Rf_length
#define xlength Rf_xlength
#define allocVector Rf_allocVector
#define PROTECT(x) Rf_protect(x)
#define UNPROTECT(n) Rf_unprotect(n)
#define mkChar Rf_mkChar
#define mkString Rf_mkString
#define ScalarInteger Rf_ScalarInteger
#define ScalarLogical Rf_ScalarLogical
#define ScalarReal Rf_ScalarReal
#define asInteger Rf_asInteger
#define asLogical Rf_asLogical
#define asReal Rf_asReal
#define isNull Rf_isNull
#define isNewList Rf_isNewList
#define isString Rf_isString
#define inherits Rf_inherits
#define install Rf_install
#define setAttrib Rf_setAttrib
#define getAttrib Rf_getAttrib
#define error Rf_error
#define warning Rf_warning

#define R_Version(v, p, s) (((v) * 65536) + ((p) * 256) + (s))
#define R_VERSION R_Version(4, 4, 1)

#ifdef __cplusplus
}
#endif
#endif

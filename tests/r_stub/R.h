/* Minimal stand-in for R's C API headers: only what r-binding/src/mmsgpu_shim.c uses, so that
 * the shim can be syntax- and prototype-checked in an image without R.  TEST INFRASTRUCTURE. */
#ifndef MMS_R_STUB_H
#define MMS_R_STUB_H
#include <stddef.h>
typedef struct SEXPREC *SEXP;
typedef ptrdiff_t R_xlen_t;
#endif

/* Stub of R's <R.h> for type-checking r/src/mcstate_b200_r.c where R is not installed
 * (tests/test_r_binding.py: gcc -fsyntax-only).  Declarations follow R's public C API
 * (Writing R Extensions, "The R API"); nothing here is linked or shipped. */
#ifndef MCB_STUB_R_H
#define MCB_STUB_R_H
#include <stddef.h>
typedef enum { FALSE = 0, TRUE } Rboolean;
char *R_alloc(size_t n, int size);
#endif

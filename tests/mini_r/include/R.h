/* tests/mini_r/include/R.h -- TEST INFRASTRUCTURE (see Rinternals.h in this directory). */
#ifndef MINI_R_H
#define MINI_R_H
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef __cplusplus
extern "C" {
#endif
char *R_alloc(size_t n, int size);            /* transient storage, reclaimed at the end of the .Call */
void Rprintf(const char *, ...) __attribute__((format(printf, 1, 2)));
void REprintf(const char *, ...) __attribute__((format(printf, 1, 2)));
#ifdef __cplusplus
}
#endif
#endif

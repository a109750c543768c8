/* TEST INFRASTRUCTURE (oracle build only).
 * Maps the ATLAS clapack names the reference binds at
 * bayesian_quadrature/linalg_c.pyx:42-45 onto scipy's bundled LAPACK (Fortran
 * symbols with a scipy_ prefix).  The reference always passes CblasColMajor /
 * CblasLower, so order/uplo are ignored and "L" is used. */
#pragma once
#include <stddef.h>
#include <stdint.h>
void scipy_dpotrf_(const char *, const int *, double *, const int *, int *, size_t);
void scipy_dpotrs_(const char *, const int *, const int *, const double *, const int *, double *,
                   const int *, int *, size_t);
static inline int32_t clapack_dpotrf(int32_t order, int32_t uplo, int32_t n, double *a, int32_t lda) {
    int info;
    (void)order; (void)uplo;
    scipy_dpotrf_("L", &n, a, &lda, &info, 1);
    return info;
}
static inline int32_t clapack_dpotrs(int32_t order, int32_t uplo, int32_t n, int32_t nrhs, double *a,
                                     int32_t lda, double *b, int32_t ldb) {
    int info;
    (void)order; (void)uplo;
    scipy_dpotrs_("L", &n, &nrhs, a, &lda, b, &ldb, &info, 1);
    return info;
}

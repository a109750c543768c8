/* TEST INFRASTRUCTURE (oracle build only).
 * Maps the ATLAS C-interface name the reference binds at
 * bayesian_quadrature/linalg_c.pyx:14-23 onto the LP64 OpenBLAS that ships inside
 * scipy (exported with a scipy_ prefix). */
#pragma once
#include <stdint.h>
enum CBLAS_ORDER { CblasRowMajor = 101, CblasColMajor = 102 };
enum CBLAS_UPLO { CblasUpper = 121, CblasLower = 122 };
double scipy_cblas_ddot(int n, const double *x, int incx, const double *y, int incy);
static inline double cblas_ddot(int32_t n, double *x, int32_t incx, double *y, int32_t incy) {
    return scipy_cblas_ddot(n, x, incx, y, incy);
}

/*
 * lapacke_dgesv.c -- LAPACKE_dgesv for the oracle/_ref build (test infrastructure only).
 *
 * The reference's encoder solves L x1 = z and U x2 = x1 with LAPACKE_dgesv
 * (lib/ldpc_encoder_bc_impl.cc:180-223) from the system LAPACK (version unpinned,
 * lib/CMakeLists.txt:41), which this image does not have.  This is the published algorithm of
 * DGESV: DGETRF (LU with partial pivoting, pivot = first entry of maximum magnitude in the
 * column, as IDAMAX picks it; info = k if U(k,k) is exactly zero, in which case no solve is
 * done) followed by DGETRS (row interchanges on B, unit-lower forward substitution, upper back
 * substitution).
 *
 * Why any conforming DGESV gives the same bits on this path: the reference only passes 0/1
 * matrices that are lower triangular (L) or upper triangular (U).  For a triangular matrix with
 * a non-zero diagonal the diagonal entry is always the first maximum of its column, the
 * elimination updates touch only zeros, so no row is ever interchanged and the solve is plain
 * substitution on integers -- exact in fp64 (below 2^53) whatever the blocking or the order of
 * the updates inside a particular LAPACK build.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "lapacke/lapacke.h"

lapack_int LAPACKE_dgesv(int layout, lapack_int n, lapack_int nrhs, double *a, lapack_int lda,
                         lapack_int *ipiv, double *b, lapack_int ldb)
{
    if (layout != LAPACK_ROW_MAJOR && layout != LAPACK_COL_MAJOR) return -1;
    if (n < 0) return -2;
    if (nrhs < 0) return -3;
    if (lda < (n > 1 ? n : 1)) return -5;
    if (layout == LAPACK_ROW_MAJOR ? ldb < (nrhs > 1 ? nrhs : 1) : ldb < (n > 1 ? n : 1)) return -8;

#define A_(i, j) (layout == LAPACK_ROW_MAJOR ? a[(size_t)(i) * lda + (j)] : a[(size_t)(j) * lda + (i)])
#define B_(i, j) (layout == LAPACK_ROW_MAJOR ? b[(size_t)(i) * ldb + (j)] : b[(size_t)(j) * ldb + (i)])
#define AP(i, j) (layout == LAPACK_ROW_MAJOR ? &a[(size_t)(i) * lda + (j)] : &a[(size_t)(j) * lda + (i)])
#define BP(i, j) (layout == LAPACK_ROW_MAJOR ? &b[(size_t)(i) * ldb + (j)] : &b[(size_t)(j) * ldb + (i)])

    lapack_int info = 0;
    for (lapack_int k = 0; k < n; k++) {                      /* DGETRF, unblocked */
        lapack_int p = k;
        double best = fabs(A_(k, k));
        for (lapack_int i = k + 1; i < n; i++) {
            const double v = fabs(A_(i, k));
            if (v > best) { best = v; p = i; }
        }
        ipiv[k] = p + 1;
        if (A_(p, k) != 0.0) {
            if (p != k) {
                for (lapack_int j = 0; j < n; j++) {
                    const double t = A_(k, j);
                    *AP(k, j) = A_(p, j);
                    *AP(p, j) = t;
                }
            }
            const double d = A_(k, k);
            for (lapack_int i = k + 1; i < n; i++) *AP(i, k) = A_(i, k) / d;
        } else if (info == 0) {
            info = k + 1;
        }
        for (lapack_int i = k + 1; i < n; i++) {
            const double l = A_(i, k);
            if (l == 0.0) continue;
            for (lapack_int j = k + 1; j < n; j++) *AP(i, j) = A_(i, j) - l * A_(k, j);
        }
    }
    if (info != 0) return info;

    for (lapack_int c = 0; c < nrhs; c++) {                   /* DGETRS, no transpose */
        for (lapack_int k = 0; k < n; k++) {
            const lapack_int p = ipiv[k] - 1;
            if (p != k) {
                const double t = B_(k, c);
                *BP(k, c) = B_(p, c);
                *BP(p, c) = t;
            }
        }
        for (lapack_int k = 0; k < n; k++) {                  /* L y = P b, unit diagonal */
            const double y = B_(k, c);
            if (y == 0.0) continue;
            for (lapack_int i = k + 1; i < n; i++) *BP(i, c) = B_(i, c) - y * A_(i, k);
        }
        for (lapack_int k = n - 1; k >= 0; k--) {             /* U x = y */
            if (B_(k, c) == 0.0) continue;
            *BP(k, c) = B_(k, c) / A_(k, k);
            const double x = B_(k, c);
            for (lapack_int i = 0; i < k; i++) *BP(i, c) = B_(i, c) - x * A_(i, k);
        }
    }
    return 0;
}

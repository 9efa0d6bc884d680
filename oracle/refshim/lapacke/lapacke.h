/* Stand-in for <lapacke/lapacke.h> (oracle/_ref build only; LAPACK is not in this image).
 * Declares the one routine the reference's encoder calls (lib/ldpc_encoder_bc_impl.cc:205);
 * refshim/lapacke_dgesv.c implements it. */
#pragma once
#ifdef __cplusplus
extern "C" {
#endif
#define LAPACK_ROW_MAJOR 101
#define LAPACK_COL_MAJOR 102
typedef int lapack_int;
lapack_int LAPACKE_dgesv(int matrix_layout, lapack_int n, lapack_int nrhs, double *a,
                         lapack_int lda, lapack_int *ipiv, double *b, lapack_int ldb);
#ifdef __cplusplus
}
#endif

/*
 * ldpc_oracle.c -- CPU restatement of the reference's LDPC encode/decode path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing the product ships links, imports or calls
 * this file: only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline /
 * `--impl reference` legs may use it, and only as the checker / the CPU number.
 *
 * Parity pinning, two anchors:
 *  (1) the ONLY golden vectors the reference holds for the path: the two QA known-answer
 *      tests on the 8x16 code (python/qa_ldpc_encoder_bc.py:23-46,
 *      python/qa_ldpc_decoder_cb.py:20-43; tests/test_oracle_kat.py);
 *  (2) the reference's OWN CODE run here: oracle/_ref/libldpc_ref.so is the reference's
 *      lib/ldpc_decoder_cb_impl.cc + lib/ldpc_encoder_bc_impl.cc compiled unmodified from
 *      /root/reference against small stand-ins for their dependencies (oracle/refshim/: uBLAS
 *      containers, gr::block base class, LAPACKE_dgesv -- GNU Radio 3.7, Boost and LAPACK are
 *      not in this image; oracle/ref_driver.cc, oracle/Makefile).  tests/test_reference_build.py
 *      checks every function of this file against it -- tables, parity, all four decode
 *      methods on all five reference matrices, both general_work bodies call by call with the
 *      sync machine -- live and through tests/golden/ref_build_golden.npz (outputs of that
 *      library, tools/gen_ref_golden.py).  What the stand-ins cannot pin: the real uBLAS /
 *      LAPACK binaries (see DESIGN.md "Oracle" for why they cannot change the values).
 *
 * Every function follows the reference loop for loop (dense H scans, fp64,
 * ascending accumulation order, per-call allocation of the message matrices)
 * and cites the lines it restates.  Knobs the reference does not have
 * (early_stop = 0, iteration count read-back, message dumps) default to the
 * reference behaviour.
 *
 * Plain C99, no dependencies:  gcc -O3 -DNDEBUG -shared -fPIC (see Makefile).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define H_(j, i) H[(size_t)(j) * n + (i)]

/* apps/ldpc_lapack.cpp:102-115, lib/ldpc_decoder_cb_impl.cc:580-595 */
static int mod2i(long long v)
{
    int r = (int)(v % 2);
    if (r < 0) r += 2;
    return r;
}

/*
 * reorderHMatrix -- apps/ldpc_lapack.cpp:163-211
 * (== lib/ldpc_decoder_cb_impl.cc:255-307, lib/ldpc_encoder_bc_impl.cc:225-273).
 * H is M x N, column-permuted in place; L and U are M x (N-M), zero-filled by the
 * caller like the constructors do (lib/ldpc_decoder_cb_impl.cc:104-105).
 * chosen[i] (may be NULL) records the pivot column of step i.
 * "First" strategy: first non-zero at/after the diagonal, else column 0.
 */
void orc_reorder_h(int *H, int M, int N, int *L, int *U, int *chosen)
{
    const int n = N, w = N - M;
    int *F = (int *)malloc(sizeof(int) * (size_t)M * N);
    memcpy(F, H, sizeof(int) * (size_t)M * N);

    for (int i = 0; i < M; i++) {
        int chosenCol = 0;
        for (int j = i; j < N; j++) {
            if (F[(size_t)i * n + j] != 0) { chosenCol = j; break; }
        }
        if (chosen) chosen[i] = chosenCol;

        for (int r = 0; r < M; r++) {           /* swap columns i <-> chosenCol in F and H */
            int t = F[(size_t)r * n + i];
            F[(size_t)r * n + i] = F[(size_t)r * n + chosenCol];
            F[(size_t)r * n + chosenCol] = t;
            t = H_(r, i); H_(r, i) = H_(r, chosenCol); H_(r, chosenCol) = t;
        }
        if (i < w) {                            /* L, U have N-M columns (== M for rate 1/2) */
            for (int r = i; r < M; r++) L[(size_t)r * w + i] = F[(size_t)r * n + i];
            for (int r = 0; r <= i; r++) U[(size_t)r * w + i] = F[(size_t)r * n + i];
        }
        if (i < M - 1) {
            for (int k = i + 1; k < M; k++) {
                if (F[(size_t)k * n + i] != 0) {
                    for (int c = 0; c < N; c++)
                        F[(size_t)k * n + c] =
                            mod2i((long long)F[(size_t)k * n + c] + F[(size_t)i * n + c]);
                }
            }
        }
    }
    free(F);
}

/*
 * solve -- apps/ldpc_lapack.cpp:121-161 / lib/ldpc_encoder_bc_impl.cc:180-223.
 * The reference calls LAPACKE_dgesv (LAPACK, system library, version unpinned) on
 * a 0/1 UNIT-TRIANGULAR matrix.  dgesv = LU with partial pivoting then two
 * triangular solves; on a unit lower-triangular A the pivot of every column is
 * its diagonal 1 (first max-|.| entry) and elimination touches only zeros, on a
 * unit upper-triangular A there is nothing to eliminate -- so dgesv reduces to
 * plain real-valued forward / back substitution in fp64.  That is restated here
 * (real arithmetic, result truncated double->int like `x_vec(i) = b[i]`).
 * lower != 0: A is lower-triangular, else upper.  Returns 1 if a diagonal is 0
 * (the reference prints to cerr and returns garbage there).
 */
int orc_solve_real(const int *A, const int *b, int m, int lower, int *x)
{
    double *y = (double *)malloc(sizeof(double) * (size_t)m);
    int singular = 0;
    if (lower) {
        for (int i = 0; i < m; i++) {
            double s = b[i];
            for (int k = 0; k < i; k++) s -= (double)A[(size_t)i * m + k] * y[k];
            if (A[(size_t)i * m + i] == 0) { singular = 1; y[i] = 0; }
            else y[i] = s / (double)A[(size_t)i * m + i];
        }
    } else {
        for (int i = m - 1; i >= 0; i--) {
            double s = b[i];
            for (int k = i + 1; k < m; k++) s -= (double)A[(size_t)i * m + k] * y[k];
            if (A[(size_t)i * m + i] == 0) { singular = 1; y[i] = 0; }
            else y[i] = s / (double)A[(size_t)i * m + i];
        }
    }
    for (int i = 0; i < m; i++) x[i] = (int)y[i];
    free(y);
    return singular;
}

/* The same solve over GF(2): what the real-valued detour equals mod 2 while the
 * intermediates stay below 2^53; the definition for large codes. */
int orc_solve_gf2(const int *A, const int *b, int m, int lower, int *x)
{
    int singular = 0;
    if (lower) {
        for (int i = 0; i < m; i++) {
            int s = b[i] & 1;
            for (int k = 0; k < i; k++) s ^= (A[(size_t)i * m + k] & x[k]);
            if (A[(size_t)i * m + i] == 0) singular = 1;
            x[i] = s & 1;
        }
    } else {
        for (int i = m - 1; i >= 0; i--) {
            int s = b[i] & 1;
            for (int k = i + 1; k < m; k++) s ^= (A[(size_t)i * m + k] & x[k]);
            if (A[(size_t)i * m + i] == 0) singular = 1;
            x[i] = s & 1;
        }
    }
    return singular;
}

/*
 * makeParityCheck -- apps/ldpc_lapack.cpp:213-242 / lib/ldpc_encoder_bc_impl.cc:275-294.
 * z = mod2(H[:, N-M:N] * d); x1 = solve(L, z); x2 = solve(U, x1); c = mod2(x2).
 * gf2 != 0 uses GF(2) substitution instead of the real-valued one.
 */
int orc_make_parity_check(const int *d, const int *H, const int *L, const int *U,
                          int M, int N, int gf2, int *c)
{
    const int n = N, K = N - M;
    int *z = (int *)malloc(sizeof(int) * (size_t)M * 3);
    int *x1 = z + M, *x2 = z + 2 * M;
    for (int j = 0; j < M; j++) {
        long long s = 0;
        for (int k = 0; k < K; k++) s += (long long)H_(j, N - M + k) * d[k];
        z[j] = mod2i(s);
    }
    int bad;
    if (gf2) {
        bad = orc_solve_gf2(L, z, M, 1, x1);
        bad |= orc_solve_gf2(U, x1, M, 0, x2);
    } else {
        bad = orc_solve_real(L, z, M, 1, x1);
        bad |= orc_solve_real(U, x1, M, 0, x2);
    }
    for (int j = 0; j < M; j++) c[j] = mod2i(x2[j]);
    free(z);
    return bad;
}

/* checkFrame -- apps/ldpc_lapack.cpp:244-258 / lib/ldpc_decoder_cb_impl.cc:236-253.
 * Number of unsatisfied checks, early-out once it exceeds threshold. */
int orc_check_frame(const int *u, const int *H, int M, int N, int threshold)
{
    const int n = N;
    int sNotZero = 0;
    for (int k = 0; k < M; k++) {
        long long ip = 0;
        for (int i = 0; i < N; i++) ip += (long long)u[i] * H_(k, i);
        if ((int)(ip % 2) != 0) sNotZero++;
        if (sNotZero > threshold) break;
    }
    return sNotZero;
}

/*
 * decodeSumProductSoft -- apps/ldpc_lapack.cpp:260-335 /
 * lib/ldpc_decoder_cb_impl.cc:478-557.  Dense scans, fp64, message matrices
 * allocated per call, loop order preserved.
 *   early_stop != 0 : reference behaviour (test the syndrome every iteration,
 *                     the last included, and stop when it is zero)
 *   iters_run       : 1-based index of the iteration whose test passed, else
 *                     `iterations` (may be NULL)
 *   dbgL/dbgE/dbgM  : optional dumps after the loop ends: L[n] (last "Test"
 *                     sums), E and M as dense m x n (zeros off the edges)
 */
void orc_decode_spa(const double *rx, const int *H, int m_, int n_, int iterations,
                    int early_stop, int *vHat, int *iters_run,
                    double *dbgL, double *dbgE, double *dbgM)
{
    const int m = m_, n = n_;
    double *r = (double *)malloc(sizeof(double) * (size_t)n);
    double *Mm = (double *)calloc((size_t)m * n, sizeof(double));
    double *E = (double *)malloc(sizeof(double) * (size_t)m * n);
    double *Lv = (double *)calloc((size_t)n, sizeof(double));
    int run = iterations;

    for (int i = 0; i < n; i++) r[i] = -rx[i];
    for (int j = 0; j < m; j++)
        for (int i = 0; i < n; i++) {
            E[(size_t)j * n + i] = 0.0;      /* uBLAS leaves E uninitialised; only edges are read */
            if (H_(j, i) != 0) Mm[(size_t)j * n + i] = r[i];
        }
    for (int i = 0; i < n; i++) vHat[i] = 0;

    for (int h = 0; h < iterations; h++) {
        /* Step 1: check messages */
        for (int j = 0; j < m; j++)
            for (int i = 0; i < n; i++)
                if (H_(j, i) != 0) {
                    double T = 1.0;
                    for (int k = 0; k < n; k++)
                        if (H_(j, k) != 0 && k != i) T *= tanh(Mm[(size_t)j * n + k] / 2.0);
                    E[(size_t)j * n + i] = log((1.0 + T) / (1.0 - T));
                }
        /* Test */
        for (int i = 0; i < n; i++) {
            double L = 0.0;
            for (int j = 0; j < m; j++)
                if (H_(j, i) != 0) L += E[(size_t)j * n + i] + r[i];
            Lv[i] = L;
            vHat[i] = (L <= 0) ? 1 : 0;
        }
        /* Finished? */
        if (early_stop && orc_check_frame(vHat, H, m, n, 0) == 0) { run = h + 1; break; }
        /* Step 2: bit messages */
        for (int j = 0; j < m; j++)
            for (int i = 0; i < n; i++)
                if (H_(j, i) != 0) {
                    double T = 0.0;
                    for (int k = 0; k < m; k++)
                        if (k != j && H_(k, i) != 0) T += E[(size_t)k * n + i] + r[i];
                    Mm[(size_t)j * n + i] = T;
                }
    }
    if (iters_run) *iters_run = run;
    if (dbgL) memcpy(dbgL, Lv, sizeof(double) * (size_t)n);
    if (dbgE) memcpy(dbgE, E, sizeof(double) * (size_t)m * n);
    if (dbgM) memcpy(dbgM, Mm, sizeof(double) * (size_t)m * n);
    free(r); free(Mm); free(E); free(Lv);
}

/*
 * Same arithmetic as orc_decode_spa, in the same accumulation order, but walking
 * sparse adjacency lists instead of scanning dense H -- so tests on the n = 8192
 * code finish in seconds.  NOT the reference's cost model (never timed as the
 * CPU baseline); tests check it is bit-identical to the dense restatement.
 * row_ptr/col_idx: CSR of H with columns ascending per row.
 * col_ptr/edge_of_col: for each column the edge ids (CSR positions) of its ones,
 * rows ascending.  dbgE/dbgM are per-edge (CSR order) here.
 */
void orc_decode_spa_sparse(const double *rx, int m, int n, const int *row_ptr,
                           const int *col_idx, const int *col_ptr, const int *edge_of_col,
                           int iterations, int early_stop, int *vHat, int *iters_run,
                           double *dbgL, double *dbgE, double *dbgM)
{
    const int ne = row_ptr[m];
    double *r = (double *)malloc(sizeof(double) * (size_t)n);
    double *Mm = (double *)malloc(sizeof(double) * (size_t)ne);
    double *E = (double *)calloc((size_t)ne, sizeof(double));
    double *t = (double *)malloc(sizeof(double) * (size_t)ne);
    double *Lv = (double *)calloc((size_t)n, sizeof(double));
    int run = iterations;

    for (int i = 0; i < n; i++) { r[i] = -rx[i]; vHat[i] = 0; }
    for (int j = 0; j < m; j++)
        for (int e = row_ptr[j]; e < row_ptr[j + 1]; e++) Mm[e] = r[col_idx[e]];

    for (int h = 0; h < iterations; h++) {
        for (int e = 0; e < ne; e++) t[e] = tanh(Mm[e] / 2.0);
        for (int j = 0; j < m; j++)
            for (int e = row_ptr[j]; e < row_ptr[j + 1]; e++) {
                double T = 1.0;
                for (int f = row_ptr[j]; f < row_ptr[j + 1]; f++)
                    if (f != e) T *= t[f];
                E[e] = log((1.0 + T) / (1.0 - T));
            }
        for (int i = 0; i < n; i++) {
            double L = 0.0;
            for (int q = col_ptr[i]; q < col_ptr[i + 1]; q++) L += E[edge_of_col[q]] + r[i];
            Lv[i] = L;
            vHat[i] = (L <= 0) ? 1 : 0;
        }
        if (early_stop) {
            int bad = 0;
            for (int j = 0; j < m && !bad; j++) {
                int s = 0;
                for (int e = row_ptr[j]; e < row_ptr[j + 1]; e++) s += vHat[col_idx[e]];
                bad = s & 1;
            }
            if (!bad) { run = h + 1; break; }
        }
        for (int i = 0; i < n; i++)
            for (int q = col_ptr[i]; q < col_ptr[i + 1]; q++) {
                double T = 0.0;
                for (int p = col_ptr[i]; p < col_ptr[i + 1]; p++)
                    if (p != q) T += E[edge_of_col[p]] + r[i];
                Mm[edge_of_col[q]] = T;
            }
    }
    if (iters_run) *iters_run = run;
    if (dbgL) memcpy(dbgL, Lv, sizeof(double) * (size_t)n);
    if (dbgE) memcpy(dbgE, E, sizeof(double) * (size_t)ne);
    if (dbgM) memcpy(dbgM, Mm, sizeof(double) * (size_t)ne);
    free(r); free(Mm); free(E); free(t); free(Lv);
}

static int sgn(double v) { return (v > 0) - (v < 0); }   /* lib/ldpc_decoder_cb_impl.cc:574-578 */

/*
 * decodeLogDomainSimple (min-sum) -- lib/ldpc_decoder_cb_impl.cc:309-412
 * (apps/ldpc_lapack.cpp:337-441).  Strict `< 0` decision; syndrome test skipped on
 * the last iteration (`n + 1 < iterations`).
 */
void orc_decode_minsum(const double *rx, const int *H, int M, int N, int iterations,
                       int early_stop, int *vHat, int *iters_run)
{
    const int n = N;
    double *Lci = (double *)malloc(sizeof(double) * (size_t)N);
    double *Lrji = (double *)calloc((size_t)M * N, sizeof(double));
    double *Lqij = (double *)malloc(sizeof(double) * (size_t)M * N);
    int *alpha = (int *)malloc(sizeof(int) * (size_t)M * N);
    double *beta = (double *)malloc(sizeof(double) * (size_t)M * N);
    int run = iterations;

    for (int i = 0; i < N; i++) { Lci[i] = -rx[i]; vHat[i] = 0; }
    for (int i = 0; i < M; i++)
        for (int j = 0; j < N; j++) Lqij[(size_t)i * n + j] = H_(i, j) * Lci[j];

    for (int it = 0; it < iterations; it++) {
        for (int i = 0; i < M; i++)
            for (int j = 0; j < N; j++) {
                alpha[(size_t)i * n + j] = sgn(Lqij[(size_t)i * n + j]);
                beta[(size_t)i * n + j] = fabs(Lqij[(size_t)i * n + j]);
            }
        for (int i = 0; i < M; i++) {
            int prod = 1;
            for (int j = 0; j < N; j++)
                if (H_(i, j) != 0) prod *= alpha[(size_t)i * n + j];
            for (int j = 0; j < N; j++)
                if (H_(i, j) != 0) {
                    double mn = 1.7976931348623157e308;   /* numeric_limits<double>::max() */
                    for (int k = 0; k < N; k++)
                        if (j != k && H_(i, k) != 0 && beta[(size_t)i * n + k] < mn)
                            mn = beta[(size_t)i * n + k];
                    Lrji[(size_t)i * n + j] = prod * alpha[(size_t)i * n + j] * mn;
                }
        }
        for (int j = 0; j < N; j++) {
            double sum = 0.0;
            for (int i = 0; i < M; i++)
                if (H_(i, j) != 0) sum += Lrji[(size_t)i * n + j];
            for (int i = 0; i < M; i++)
                if (H_(i, j) != 0) Lqij[(size_t)i * n + j] = Lci[j] + sum - Lrji[(size_t)i * n + j];
            double LQ = Lci[j] + sum;
            vHat[j] = (LQ < 0) ? 1 : 0;
        }
        if (early_stop && it + 1 < iterations && orc_check_frame(vHat, H, M, N, 0) == 0) {
            run = it + 1;
            break;
        }
    }
    if (iters_run) *iters_run = run;
    free(Lci); free(Lrji); free(Lqij); free(alpha); free(beta);
}

/* decodeHard -- lib/ldpc_decoder_cb_impl.cc:559-572: rx < 0 -> 0 else 1. */
void orc_decode_hard(const double *rx, int N, int *vHat)
{
    for (int i = 0; i < N; i++) vHat[i] = (rx[i] < 0) ? 0 : 1;
}

/*
 * decodeBitFlipping -- lib/ldpc_decoder_cb_impl.cc:414-476.  A bit flips only when
 * more than M/2 of its checks disagree (:464); E(i,j) is formed for every (i,j)
 * but compared only on edges.
 */
void orc_decode_bitflip(const double *rx, const int *H, int M, int N, int iterations,
                        int *ci)
{
    const int n = N;
    int *y = (int *)malloc(sizeof(int) * (size_t)N);
    int *flip = (int *)malloc(sizeof(int) * (size_t)N);
    int *rowsum = (int *)malloc(sizeof(int) * (size_t)M);
    for (int i = 0; i < N; i++) { y[i] = (rx[i] < 0.0) ? 0 : 1; ci[i] = y[i]; }
    for (int it = 0; it < iterations; it++) {
        /* The reference fills the whole E matrix from the pre-sweep ci before it
         * compares anything (:441-452): E(i,j) = (sum_{k != j, H(i,k)} ci(k)) % 2.
         * Row sums of the pre-sweep ci give the same values on the edges. */
        for (int i = 0; i < M; i++) {
            int s = 0;
            for (int k = 0; k < N; k++) if (H_(i, k) != 0) s += ci[k];
            rowsum[i] = s;
        }
        for (int j = 0; j < N; j++) {
            int disagreements = 0;
            for (int i = 0; i < M; i++)
                if (H_(i, j) != 0 && (rowsum[i] - ci[j]) % 2 != y[j]) disagreements++;
            flip[j] = disagreements > M / 2;             /* :464 */
        }
        for (int j = 0; j < N; j++)
            if (flip[j]) ci[j] = (y[j] + 1) % 2;         /* :465 */
        if (it + 1 < iterations && orc_check_frame(ci, H, M, N, 0) == 0) break;
    }
    free(y); free(flip); free(rowsum);
}

/* ------------------------------------------------------------------------- */
/* Block level                                                               */
/* ------------------------------------------------------------------------- */

/*
 * ldpc_encoder_bc_impl::general_work -- lib/ldpc_encoder_bc_impl.cc:118-178.
 * in: bytes; out: interleaved complex float (re, im).  H/L/U already re-ordered.
 * Returns items produced; *consumed = input bytes consumed.
 */
int orc_encoder_work(const int *H, const int *L, const int *U, int M, int N, int gf2,
                     int noutput_items, int ninput_items, const uint8_t *in, float *out,
                     int *consumed)
{
    const int min_in = M / 8, min_out = N;
    int ic = 0, op = 0;
    int *data = (int *)malloc(sizeof(int) * (size_t)(2 * M));
    int *c = data + M;
    while ((noutput_items - op) >= min_out && (ninput_items - ic) >= min_in) {
        for (int i = 0; i < min_in; i++) {
            for (int j = 0; j < 8; j++) data[i * 8 + j] = (in[ic + i] & (1 << (7 - j))) ? 1 : 0;
        }
        orc_make_parity_check(data, H, L, U, M, N, gf2, c);
        for (int i = 0; i < M; i++) { out[2 * (op + i)] = c[i] == 1 ? 1.f : -1.f; out[2 * (op + i) + 1] = 0.f; }
        for (int i = 0; i < M; i++) {
            out[2 * (op + M + i)] = data[i] == 1 ? 1.f : -1.f;
            out[2 * (op + M + i) + 1] = 0.f;
        }
        ic += min_in;
        op += N;
    }
    free(data);
    *consumed = ic;
    return op;
}

typedef struct {
    int method;        /* 0 LogDomain(min-sum), 1 SumProduct, 2 BitFlip, 3 Hard; else LogDomain */
    int state;         /* 0 OUT_OF_SYNC, 1 IN_SYNC, 2 IN_SYNC_INVERTED */
    unsigned errors;
    int iterations;    /* reference: 5 */
} orc_decoder_state;

static void decode_by_method(const orc_decoder_state *st, const double *tx, const int *H,
                             int M, int N, int *vhat)
{
    if (st->method == 3) orc_decode_hard(tx, N, vhat);
    else if (st->method == 2) orc_decode_bitflip(tx, H, M, N, st->iterations, vhat);
    else if (st->method == 1) orc_decode_spa(tx, H, M, N, st->iterations, 1, vhat, 0, 0, 0, 0);
    else orc_decode_minsum(tx, H, M, N, st->iterations, 1, vhat, 0);
}

/*
 * ldpc_decoder_cb_impl::general_work -- lib/ldpc_decoder_cb_impl.cc:132-234,
 * sync state machine included.  in: interleaved complex float; out: bytes.
 * events (may be NULL, capacity max_events): 1 "IN SYNC", 2 "IN SYNC; PHASE
 * INVERTED", 3 "MAX ERRORS; OUT OF SYNC", in order of occurrence (the reference
 * prints these to std::cout).
 */
int orc_decoder_work(orc_decoder_state *st, const int *H, int M, int N,
                     int noutput_items, int ninput_items, const float *in, uint8_t *out,
                     int *consumed, int *events, int max_events, int *n_events)
{
    const int min_out = M / 8, thr = M / 8;
    int ic = 0, op = 0, ne = 0;
    double *tx = (double *)malloc(sizeof(double) * (size_t)N * 2);
    double *ntx = tx + N;
    int *vhat = (int *)malloc(sizeof(int) * (size_t)N);

    while ((ninput_items - ic) >= N && (noutput_items - op) >= min_out) {
        for (int i = 0; i < N; i++)
            tx[i] = (double)in[2 * (ic + i)] * (st->state == 2 ? -1 : 1);
        decode_by_method(st, tx, H, M, N, vhat);
        int sNotZero = orc_check_frame(vhat, H, M, N, thr);

        if (sNotZero > thr) {
            if (st->state == 1 || st->state == 2) {
                st->errors++;
                if (st->errors > 10) {
                    st->errors = 0;
                    st->state = 0;
                    if (events && ne < max_events) events[ne] = 3;
                    ne++;
                }
            }
            if (st->state == 0) {
                for (int i = 0; i < N; i++) ntx[i] = -tx[i];
                decode_by_method(st, ntx, H, M, N, vhat);
                if ((sNotZero = orc_check_frame(vhat, H, M, N, thr)) <= thr) {
                    if (events && ne < max_events) events[ne] = 2;
                    ne++;
                    st->state = 2;
                    st->errors = 0;
                } else {
                    ic += 1;                    /* slide one symbol */
                }
            }
        } else if (st->state == 0) {
            if (events && ne < max_events) events[ne] = 1;
            ne++;
            st->state = 1;
            st->errors = 0;
        }

        if (st->state == 1 || st->state == 2) {
            for (int i = 0; i < min_out; i++) {
                uint8_t b = 0;
                for (int j = 0; j < 8; j++)
                    if (vhat[M + i * 8 + j] == 1) b |= (uint8_t)(1 << (7 - j));
                out[op + i] = b;
            }
            ic += N;
            op += min_out;
        }
    }
    free(tx); free(vhat);
    *consumed = ic;
    if (n_events) *n_events = ne;
    return op;
}

/* ------------------------------------------------------------------------- */
/* Batch drivers (CPU baseline timing; one call per worker thread)           */
/* ------------------------------------------------------------------------- */

/* Decode n_cw aligned frames (stride N complex symbols) with the dense SPA
 * restatement; writes packed data bytes (M/8 per frame), iteration counts and
 * the saturating syndrome weight of lib/ldpc_decoder_cb_impl.cc:166. */
void orc_decode_frames(const float *sym, long n_cw, const int *H, int M, int N, int method,
                       int iterations, int early_stop, int thr, uint8_t *out_bytes,
                       uint8_t *out_iters, uint8_t *out_synd)
{
    double *tx = (double *)malloc(sizeof(double) * (size_t)N);
    int *vhat = (int *)malloc(sizeof(int) * (size_t)N);
    const int nb = (N - M) / 8;
    for (long c = 0; c < n_cw; c++) {
        int run = 0;
        for (int i = 0; i < N; i++) tx[i] = (double)sym[2 * ((size_t)c * N + i)];
        if (method == 3) orc_decode_hard(tx, N, vhat);
        else if (method == 2) orc_decode_bitflip(tx, H, M, N, iterations, vhat);
        else if (method == 1) orc_decode_spa(tx, H, M, N, iterations, early_stop, vhat, &run, 0, 0, 0);
        else orc_decode_minsum(tx, H, M, N, iterations, early_stop, vhat, &run);
        for (int i = 0; i < nb; i++) {
            uint8_t b = 0;
            for (int j = 0; j < 8; j++)
                if (vhat[M + i * 8 + j] == 1) b |= (uint8_t)(1 << (7 - j));
            out_bytes[(size_t)c * nb + i] = b;
        }
        if (out_iters) out_iters[c] = (uint8_t)(run > 255 ? 255 : run);
        if (out_synd) out_synd[c] = (uint8_t)orc_check_frame(vhat, H, M, N, thr);
    }
    free(tx); free(vhat);
}

/* Plain-C restatement of the reference's density-fitted Fock build.
 *
 * TEST INFRASTRUCTURE ONLY -- an independent second statement of the same
 * arithmetic as oracle/df_fock_oracle.py (no BLAS, no NumPy), used to
 * cross-check the NumPy oracle and, through it, the CUDA engine.  Nothing in
 * the product package links or loads this file.
 *
 * PARITY UNPINNED: the Fortran reference cannot be built in this image and its
 * own tests hold no J/K/F element-level golden vectors.
 *
 * Follows, loop for loop,
 *   /root/reference/backends/libcint/mqc_libcint_rhf.f90:1576-1646  (build_fock_df)
 * with `pic_gemm(A,B,C[,transb][,alpha][,beta])` == C = alpha*A*op(B) + beta*C.
 *
 * Layout is the reference's: every matrix column-major; b is (n*n, naux) with
 * element (mu,nu) of slab p at b[(mu + n*nu) + n*n*p]
 * (mqc_libcint_integrals.F90:1425-1437).
 *
 * Build:  make -C oracle        (gcc -O2 -shared -fPIC; scalar, no OpenMP runtime in this image)
 */
#include <stdlib.h>
#include <string.h>

/* C(m x n) = alpha * A(m x k) * B(k x n) + beta * C, column-major, no transposes */
static void gemm_nn(int m, int n, int k, double alpha, const double *a, int lda,
                    const double *b, int ldb, double beta, double *c, int ldc)
{
    #pragma omp parallel for schedule(static)
    for (int j = 0; j < n; ++j) {
        for (int i = 0; i < m; ++i) c[i + (size_t)ldc * j] *= beta;
        for (int l = 0; l < k; ++l) {
            const double blj = alpha * b[l + (size_t)ldb * j];
            const double *acol = a + (size_t)lda * l;
            double *ccol = c + (size_t)ldc * j;
            for (int i = 0; i < m; ++i) ccol[i] += acol[i] * blj;
        }
    }
}

/* C(m x n) = alpha * A(m x k) * B(n x k)^T + beta * C */
static void gemm_nt(int m, int n, int k, double alpha, const double *a, int lda,
                    const double *b, int ldb, double beta, double *c, int ldc)
{
    #pragma omp parallel for schedule(static)
    for (int j = 0; j < n; ++j) {
        for (int i = 0; i < m; ++i) c[i + (size_t)ldc * j] *= beta;
        for (int l = 0; l < k; ++l) {
            const double bjl = alpha * b[j + (size_t)ldb * l];
            const double *acol = a + (size_t)lda * l;
            double *ccol = c + (size_t)ldc * j;
            for (int i = 0; i < m; ++i) ccol[i] += acol[i] * bjl;
        }
    }
}

/* J, K and c of build_fock_df before the final scaling.
 * k_factor is the alpha of the second gemm: 2.0 in the reference's RHF build
 * (rhf.f90:1637); 1.0 gives the per-spin K_sigma of the two-spin convention
 * (backends/cuest/backend/mqc_cuest_scf.f90:48-57).
 * Returns 0 on success, 1 on allocation failure. */
int df_ref_jk(int n, int naux, int n_occ, const double *b, const double *density,
              const double *coeff, int ldc, double k_factor,
              double *j, double *k, double *cvec)
{
    const size_t nn = (size_t)n * n;
    double *w = (double *)malloc(sizeof(double) * (size_t)n * (n_occ > 0 ? n_occ : 1));
    double *c_occ = (double *)malloc(sizeof(double) * (size_t)n * (n_occ > 0 ? n_occ : 1));
    if (!w || !c_occ) { free(w); free(c_occ); return 1; }

    for (int i = 0; i < n_occ; ++i)                         /* c_occ = coeff(:,1:n_occ)   :1618 */
        memcpy(c_occ + (size_t)n * i, coeff + (size_t)ldc * i, sizeof(double) * n);

    for (int p = 0; p < naux; ++p) {                        /* :1620-1622 */
        const double *bp = b + nn * p;
        double s = 0.0;
        for (size_t e = 0; e < nn; ++e) s += bp[e] * density[e];
        cvec[p] = s;
    }

    for (size_t e = 0; e < nn; ++e) j[e] = 0.0;             /* :1624 */
    for (int p = 0; p < naux; ++p) {                        /* :1625-1627 */
        const double *bp = b + nn * p;
        const double cp = cvec[p];
        for (size_t e = 0; e < nn; ++e) j[e] = j[e] + cp * bp[e];
    }

    for (size_t e = 0; e < nn; ++e) k[e] = 0.0;             /* :1629 */
    for (int p = 0; p < naux && n_occ > 0; ++p) {           /* :1630-1639 */
        const double *bp = b + nn * p;
        memset(w, 0, sizeof(double) * (size_t)n * n_occ);   /* w = 0            :1635 */
        gemm_nn(n, n_occ, n, 1.0, bp, n, c_occ, n, 0.0, w, n);      /* :1636 */
        gemm_nt(n, n, n_occ, k_factor, w, n, w, n, 1.0, k, n);      /* :1637 */
    }
    free(w); free(c_occ);
    return 0;
}

/* fock = h + jf*j - kf*k with kf = 0.5*k_scale, jf = j_scale     :1641-1645 */
int df_ref_build_fock(int n, int naux, int n_occ, const double *h, const double *b,
                      const double *density, const double *coeff, int ldc,
                      double k_scale, double j_scale, double *fock)
{
    const size_t nn = (size_t)n * n;
    double *j = (double *)malloc(sizeof(double) * nn);
    double *k = (double *)malloc(sizeof(double) * nn);
    double *c = (double *)malloc(sizeof(double) * (naux > 0 ? naux : 1));
    if (!j || !k || !c) { free(j); free(k); free(c); return 1; }
    int rc = df_ref_jk(n, naux, n_occ, b, density, coeff, ldc, 2.0, j, k, c);
    if (rc == 0) {
        const double kf = 0.5 * k_scale, jf = j_scale;
        for (size_t e = 0; e < nn; ++e) fock[e] = h[e] + jf * j[e] - kf * k[e];
    }
    free(j); free(k); free(c);
    return rc;
}

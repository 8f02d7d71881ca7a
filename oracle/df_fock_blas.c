/* df_fock_blas.c -- TEST INFRASTRUCTURE ONLY: the CPU baseline of bench.py.
 *
 * A plain-C restatement of the reference's build_fock_df
 * (backends/libcint/mqc_libcint_rhf.f90:1576-1646) with the SAME structure the Fortran has:
 *
 *   c(p) = sum(reshape(b(:,p),[n,n]) * density)                 :1620-1622   serial loops
 *   j    = j + c(p) * reshape(b(:,p),[n,n])                     :1624-1627   serial loops
 *   w = 0; pic_gemm(B_p, c_occ, w)                              :1635-1636   dgemm  (n x n) . (n x o)
 *   pic_gemm(w, w, k, transb='T', alpha=2, beta=1)              :1637        dgemm  (n x o) . (o x n), FULL, not syrk
 *   fock = h + jf*j - kf*k                                      :1645
 *
 * pic_gemm is the reference's thin wrapper over the vendor dgemm (pic-blas v0.3.1, not in the
 * tree), so the dgemm here is passed in by the caller: bench.py hands over the Fortran-ABI
 * `dgemm` of the OpenBLAS that SciPy ships (scipy.linalg.cython_blas), with the BLAS thread
 * count set explicitly (1 = the reference's default sequential-BLAS build,
 * CMakeLists.txt:38-46; all cores = a threaded-BLAS build).  The J loops have no OpenMP in the
 * reference and have none here.  Three passes over the unpacked B, as in the reference.
 *
 * Never linked into the product; only tests/ and bench.py's CPU legs load it.
 */
#include <stddef.h>
#include <string.h>

typedef void (*dgemm_fn)(const char *transa, const char *transb, const int *m, const int *n, const int *k,
                         const double *alpha, const double *a, const int *lda, const double *b, const int *ldb,
                         const double *beta, double *c, const int *ldc);

/* J and K of build_fock_df before scaling.  b: (n*n, naux) column-major; density: n x n;
 * coeff: n x (>= n_occ), leading dimension ldc; k_alpha: 2.0 (restricted, :1637) or 1.0 (per
 * spin, mqc_cuest_scf.f90:48-57).  j / k may be NULL to skip that half.  work: n*n_occ doubles
 * for w plus n*n_occ for the contiguous copy of c_occ (:1618).  Returns 0. */
int df_jk_blas(dgemm_fn dgemm, int n, int naux, const double *b, const double *density, const double *coeff,
               int ldc, int n_occ, double k_alpha, double *j, double *k, double *cvec, double *work) {
  const size_t nn = (size_t)n * (size_t)n;
  if (j) {
    for (int p = 0; p < naux; ++p) {                         /* :1620-1622 */
      const double *bp = b + (size_t)p * nn;
      double s = 0.0;
      for (size_t e = 0; e < nn; ++e) s += bp[e] * density[e];
      cvec[p] = s;
    }
    memset(j, 0, nn * sizeof(double));                        /* :1624 */
    for (int p = 0; p < naux; ++p) {                         /* :1625-1627 */
      const double *bp = b + (size_t)p * nn;
      const double c = cvec[p];
      for (size_t e = 0; e < nn; ++e) j[e] += c * bp[e];
    }
  }
  if (k) {
    memset(k, 0, nn * sizeof(double));                        /* :1629 */
    if (n_occ > 0) {
      double *w = work;
      double *c_occ = work + (size_t)n * (size_t)n_occ;
      for (int i = 0; i < n_occ; ++i)                         /* c_occ = coeff(:, 1:n_occ)   :1618 */
        memcpy(c_occ + (size_t)i * n, coeff + (size_t)i * ldc, (size_t)n * sizeof(double));
      const double one = 1.0, zero = 0.0;
      for (int p = 0; p < naux; ++p) {                       /* :1630-1639 */
        const double *bp = b + (size_t)p * nn;
        memset(w, 0, (size_t)n * (size_t)n_occ * sizeof(double));                   /* w = 0      :1635 */
        dgemm("N", "N", &n, &n_occ, &n, &one, bp, &n, c_occ, &n, &zero, w, &n);    /* :1636 */
        dgemm("N", "T", &n, &n, &n_occ, &k_alpha, w, &n, w, &n, &one, k, &n);      /* :1637 */
      }
    }
  }
  return 0;
}

/* fock = h + jf*j - kf*k   (:1641-1645); j or k may be NULL (== zero). */
void df_assemble(int n, const double *h, const double *j, const double *k, double jf, double kf, double *fock) {
  const size_t nn = (size_t)n * (size_t)n;
  for (size_t e = 0; e < nn; ++e) fock[e] = h[e] + (j ? jf * j[e] : 0.0) - (k ? kf * k[e] : 0.0);
}

// scf_kernels.cu -- the SCF step around the Fock build, on the device, for fragment-sized
// problems (n <= 80): so that a fragment's SCF iterates on the GPU and only scalars cross
// PCIe per iteration (the reference's GPU backend does the same around its
// closed-source Fock build, backends/cuest/backend/mqc_cuest_scf.f90:444-553).
//
// One CTA does everything that `run_libcint_rhf` does between two Fock builds
// (backends/libcint/mqc_libcint_rhf.f90:610-636):
//
//   commutator      e = X^T (F D S - S D F) X                      rhf.f90:1326-1352
//   DIIS            push, scaled B matrix, pivoted elimination      src/methods/mqc_diis.f90:94-273
//   diagonalize     F' = X^T F X, eigendecomposition, C = X C'      rhf.f90:1464-1489
//   density         D = 2 C_occ C_occ^T                             src/scf/mqc_scf_common.f90:84-96
//   convergence     iter > 1 and |dE| < e_tol and rms(dD) < d_tol   rhf.f90:626-636
//
// plus the set-up pieces (orthogonaliser X = U s^-1/2 over eigenvalues > 1e-7,
// mqc_scf_common.f90:42-82; GWH starting Fock, rhf.f90:1354-1380).  LAPACK's dsyev is replaced
// by a cyclic two-sided Jacobi eigensolver in shared memory (parallel "round-robin" ordering:
// n/2 disjoint rotations per round); eigenvalues are sorted ascending afterwards, so the occupied
// space -- the only thing the density depends on -- is the one dsyev gives.  All reductions are in
// fixed order: a fragment's SCF is bit-reproducible.
#include "common.cuh"
#include "kernels.cuh"

namespace mqcb200 {

constexpr int SCF_THREADS = 512;
constexpr int SCF_MAX_N = 80;
constexpr int SCF_LD = 81;                 // odd leading dimension in shared memory: conflict-free rows AND columns
constexpr double SCF_LINDEP_TOL = 1.0e-7;  // LINEAR_DEPENDENCE_TOL, mqc_scf_common.f90:27
constexpr double SCF_GWH_K = 1.75;         // mqc_scf_common.f90:33
constexpr double SCF_PIVOT_FLOOR = 1.0e-14;   // mqc_diis.f90:35

// C(M x N, ldc) = alpha * op(A) * op(B); op(A) is M x K.  Column-major; ta/tb = transpose flags.
// Block-cooperative, 2x2 register tiles; operands may have been written by this block (plain loads).
__device__ void blk_gemm(int M, int N, int K, double alpha, const double *A, int lda, bool ta, const double *B, int ldb,
                         bool tb, double *C, int ldc) {
  const int tm = (M + 1) / 2, tn = (N + 1) / 2;
  for (int t = threadIdx.x; t < tm * tn; t += blockDim.x) {
    const int i0 = 2 * (t % tm), j0 = 2 * (t / tm);
    const bool i1 = i0 + 1 < M, j1 = j0 + 1 < N;
    double c00 = 0.0, c01 = 0.0, c10 = 0.0, c11 = 0.0;
    for (int k = 0; k < K; ++k) {
      const double a0 = ta ? A[(size_t)k + (size_t)lda * i0] : A[(size_t)i0 + (size_t)lda * k];
      const double a1 = i1 ? (ta ? A[(size_t)k + (size_t)lda * (i0 + 1)] : A[(size_t)i0 + 1 + (size_t)lda * k]) : 0.0;
      const double b0 = tb ? B[(size_t)j0 + (size_t)ldb * k] : B[(size_t)k + (size_t)ldb * j0];
      const double b1 = j1 ? (tb ? B[(size_t)j0 + 1 + (size_t)ldb * k] : B[(size_t)k + (size_t)ldb * (j0 + 1)]) : 0.0;
      c00 = fma(a0, b0, c00); c01 = fma(a0, b1, c01);
      c10 = fma(a1, b0, c10); c11 = fma(a1, b1, c11);
    }
    C[(size_t)i0 + (size_t)ldc * j0] = alpha * c00;
    if (j1) C[(size_t)i0 + (size_t)ldc * (j0 + 1)] = alpha * c01;
    if (i1) C[(size_t)i0 + 1 + (size_t)ldc * j0] = alpha * c10;
    if (i1 && j1) C[(size_t)i0 + 1 + (size_t)ldc * (j0 + 1)] = alpha * c11;
  }
  __syncthreads();
}

// Fixed-order block sum of one value per thread; result broadcast to every thread.
__device__ double blk_sum(double v, double *red /*[SCF_THREADS/32 + 1]*/) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
    red[SCF_THREADS / 32] = s;
  }
  __syncthreads();
  return red[SCF_THREADS / 32];
}

// Symmetric eigendecomposition of the m x m matrix in a_s (shared, leading dimension SCF_LD):
// on return a_s's diagonal holds the eigenvalues and v_s the eigenvectors (columns), both
// UNSORTED.  Cyclic Jacobi, round-robin ordering: round r pairs position i with position
// m_e-1-i of the sequence [0, 1+((k-1+r) mod (m_e-1))], so the m_e/2 rotations of a round touch
// disjoint rows/columns and are applied in parallel: columns of A and V, then rows of A.
__device__ void jacobi_eigh(double *a_s, double *v_s, int m, double *red, int *pq_s /*[2*40]*/, double *cs_s /*[2*40]*/,
                            int *rotated_s) {
  const int tid = threadIdx.x, nth = blockDim.x;
  const int m_e = (m + 1) & ~1, half = m_e / 2;
  for (int e = tid; e < m * m; e += nth) {
    const int i = e % m, j = e / m;
    v_s[i + SCF_LD * j] = i == j ? 1.0 : 0.0;
  }
  double fro = 0.0;
  for (int e = tid; e < m * m; e += nth) { const double x = a_s[(e % m) + SCF_LD * (e / m)]; fro = fma(x, x, fro); }
  fro = blk_sum(fro, red);
  // an off-diagonal element below 1e-17 |A|_F is left alone; a sweep that rotates nothing ends the iteration
  const double skip = 1.0e-17 * sqrt(fro);
  for (int sweep = 0; sweep < 40; ++sweep) {
    if (tid == 0) *rotated_s = 0;
    __syncthreads();
    for (int r = 0; r < m_e - 1; ++r) {
      if (tid < half) {
        const int pa = tid == 0 ? 0 : 1 + (tid - 1 + r) % (m_e - 1);
        const int pb = 1 + (m_e - 1 - tid - 1 + r) % (m_e - 1);
        int p = pa < pb ? pa : pb, q = pa < pb ? pb : pa;
        double c = 1.0, s = 0.0;
        if (q < m) {
          const double apq = a_s[p + SCF_LD * q];
          const double app = a_s[p + SCF_LD * p], aqq = a_s[q + SCF_LD * q];
          if (fabs(apq) > skip) {
            const double tau = (aqq - app) / (2.0 * apq);
            const double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
            c = 1.0 / sqrt(1.0 + t * t);
            s = t * c;
            *rotated_s = 1;
          }
        } else {
          q = p;                                           // the padding index of an odd m: no rotation
        }
        pq_s[2 * tid] = p; pq_s[2 * tid + 1] = q;
        cs_s[2 * tid] = c; cs_s[2 * tid + 1] = s;
      }
      __syncthreads();
      // columns: (A[:,p], A[:,q]) <- (c A[:,p] - s A[:,q], s A[:,p] + c A[:,q]); same for V
      for (int e = tid; e < half * m * 2; e += nth) {
        const int which = e / (half * m);                  // 0: A, 1: V
        const int f = e - which * half * m;
        const int k = f / m, i = f - k * m;
        const int p = pq_s[2 * k], q = pq_s[2 * k + 1];
        if (p == q) continue;
        const double c = cs_s[2 * k], s = cs_s[2 * k + 1];
        double *mat = which ? v_s : a_s;
        const double xp = mat[i + SCF_LD * p], xq = mat[i + SCF_LD * q];
        mat[i + SCF_LD * p] = c * xp - s * xq;
        mat[i + SCF_LD * q] = s * xp + c * xq;
      }
      __syncthreads();
      // rows of A
      for (int e = tid; e < half * m; e += nth) {
        const int k = e / m, j = e - k * m;
        const int p = pq_s[2 * k], q = pq_s[2 * k + 1];
        if (p == q) continue;
        const double c = cs_s[2 * k], s = cs_s[2 * k + 1];
        const double xp = a_s[p + SCF_LD * j], xq = a_s[q + SCF_LD * j];
        a_s[p + SCF_LD * j] = c * xp - s * xq;
        a_s[q + SCF_LD * j] = s * xp + c * xq;
      }
      __syncthreads();
    }
    const int rotated = *rotated_s;
    __syncthreads();                                       // everyone has read it before the next sweep resets it
    if (rotated == 0) break;
  }
}

// Ascending order of the m diagonal entries of a_s: order_s[k] = index of the k-th smallest
// (ties by index, so the order is deterministic).  Rank counting, one thread per entry.
__device__ void sort_eigenvalues(const double *a_s, int m, int *order_s) {
  for (int i = threadIdx.x; i < m; i += blockDim.x) {
    const double wi = a_s[i + SCF_LD * i];
    int rank = 0;
    for (int j = 0; j < m; ++j) {
      const double wj = a_s[j + SCF_LD * j];
      rank += (wj < wi || (wj == wi && j < i)) ? 1 : 0;
    }
    order_s[rank] = i;
  }
  __syncthreads();
}

struct ScfShared {
  double a[SCF_MAX_N * SCF_LD];
  double v[SCF_MAX_N * SCF_LD];
  double red[SCF_THREADS / 32 + 2];
  double cs[2 * (SCF_MAX_N / 2)];
  double coef[16];
  int pq[2 * (SCF_MAX_N / 2)];
  int order[SCF_MAX_N];
  int flag;
  int rotated;
};

// X = U s^-1/2 over the eigenvalues of S above 1e-7 (ascending order kept): x (n x n_mo), *n_mo_out.
__global__ void __launch_bounds__(SCF_THREADS, 1) scf_orthogonalizer_kernel(const double *__restrict__ s, int n,
                                                                            double *__restrict__ x,
                                                                            int *__restrict__ n_mo_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  ScfShared &sh = *reinterpret_cast<ScfShared *>(smem_raw);
  for (int e = threadIdx.x; e < n * n; e += blockDim.x) sh.a[(e % n) + SCF_LD * (e / n)] = s[e];
  __syncthreads();
  jacobi_eigh(sh.a, sh.v, n, sh.red, sh.pq, sh.cs, &sh.rotated);
  sort_eigenvalues(sh.a, n, sh.order);
  if (threadIdx.x == 0) {
    int dropped = 0;
    for (int k = 0; k < n; ++k) dropped += sh.a[sh.order[k] + SCF_LD * sh.order[k]] > SCF_LINDEP_TOL ? 0 : 1;
    sh.flag = dropped;
    *n_mo_out = n - dropped;
  }
  __syncthreads();
  const int dropped = sh.flag, n_mo = n - dropped;
  for (int e = threadIdx.x; e < n * n_mo; e += blockDim.x) {
    const int i = e % n, k = e / n;
    const int col = sh.order[dropped + k];               // ascending: the discarded ones lead (mqc_scf_common.f90:66)
    x[e] = sh.v[i + SCF_LD * col] / sqrt(sh.a[col + SCF_LD * col]);
  }
}

struct ScfStepArgs {
  int n, n_mo, n_occ, diis_max;
  int mode;                 // 0: guess step (diagonalise the starting Fock, build D; no DIIS, no test)
                            // 1: SCF iteration
  int guess;                // mode 0: 0 = core (F = H), 1 = GWH
  const double *h, *s, *x;  // n x n, n x n, n x n_mo
  double *fock;             // in (mode 1): F[D] from the build; mode 0: written with the guess
  double *density, *coeff, *eps;
  double *work;             // 4 * n * n doubles
  double *diis_f, *diis_e;  // [diis_max][n*n], [diis_max][n_mo*n_mo]
  double *diis_b;           // [diis_max*diis_max] overlaps in slot coordinates
  int *state;               // [0] n_stored, [1] newest (1-based slot, 0 = none), [2] iterations done, [3] converged
  double *scalars;          // [0] e_elec of this build (in), [1] e_old, [2] |dE|, [3] rms(dD), [4] DIIS extrapolated (0/1)
  double energy_tol, density_tol;
};

__global__ void __launch_bounds__(SCF_THREADS, 1) scf_step_kernel(ScfStepArgs p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  ScfShared &sh = *reinterpret_cast<ScfShared *>(smem_raw);
  const int tid = threadIdx.x, nth = blockDim.x;
  const int n = p.n, m = p.n_mo, nn = n * n, mm = m * m;
  if (p.mode == 1 && p.state[3] != 0) return;            // already converged: later queued iterations are no-ops
  double *w0 = p.work, *w1 = w0 + nn, *w2 = w1 + nn, *w3 = w2 + nn;
  const double *f_use = p.fock;

  if (p.mode == 0) {
    // ---- starting Fock: core (F = H) or generalised Wolfsberg-Helmholz (rhf.f90:1354-1380)
    for (int e = tid; e < nn; e += nth) {
      const int i = e % n, j = e / n;
      double v = p.h[e];
      if (p.guess == 1 && i != j) v = 0.5 * SCF_GWH_K * p.s[e] * (p.h[(size_t)i * n + i] + p.h[(size_t)j * n + j]);
      p.fock[e] = v;
    }
    __syncthreads();
  } else {
    // ---- commutator e = X^T (F D S - S D F) X; S D F = (F D S)^T for the symmetric F, D, S
    blk_gemm(n, n, n, 1.0, p.fock, n, false, p.density, n, false, w0, n);        // F D
    blk_gemm(n, n, n, 1.0, w0, n, false, p.s, n, false, w1, n);                  // F D S
    for (int e = tid; e < nn; e += nth) { const int i = e % n, j = e / n; w0[e] = w1[e] - w1[(size_t)j + (size_t)n * i]; }
    __syncthreads();
    blk_gemm(n, m, n, 1.0, w0, n, false, p.x, n, false, w2, n);                  // (..) X
    // ---- DIIS push (mqc_diis.f90:94-119): the ring slot after the newest; overlaps of the new entry
    int n_stored = p.state[0], newest = p.state[1];
    const int dmax = p.diis_max;
    if (dmax > 0) {
      newest = newest % dmax + 1;
      if (n_stored < dmax) n_stored += 1;
      const int slot = newest - 1;
      blk_gemm(m, m, n, 1.0, p.x, n, true, w2, n, false, p.diis_e + (size_t)slot * mm, m);   // X^T (..)
      for (int e = tid; e < nn; e += nth) p.diis_f[(size_t)slot * nn + e] = p.fock[e];
      __syncthreads();
      for (int age = 1; age <= n_stored; ++age) {
        const int other = ((newest - n_stored + age - 1) % dmax + dmax) % dmax;          // diis_slot_of_age - 1
        double part = 0.0;
        for (int e = tid; e < mm; e += nth) part = fma(p.diis_e[(size_t)slot * mm + e], p.diis_e[(size_t)other * mm + e], part);
        const double dot = blk_sum(part, sh.red);
        if (tid == 0) { p.diis_b[slot * dmax + other] = dot; p.diis_b[other * dmax + slot] = dot; }
      }
      __syncthreads();
      // ---- coefficients (mqc_diis.f90:146-273): ages oldest -> newest, error block scaled to O(1)
      if (tid == 0) {
        int ok = 0;
        if (n_stored >= 2) {
          const int nb = n_stored + 1;
          double aug[9][10];
          double scale = 0.0;
          for (int i = 0; i < n_stored; ++i)
            for (int j = 0; j < n_stored; ++j) {
              const int si = ((newest - n_stored + i) % dmax + dmax) % dmax, sj = ((newest - n_stored + j) % dmax + dmax) % dmax;
              aug[i][j] = p.diis_b[si * dmax + sj];
              scale = fmax(scale, fabs(aug[i][j]));
            }
          for (int i = 0; i < n_stored; ++i)
            for (int j = 0; j < n_stored; ++j) if (scale > 0.0) aug[i][j] /= scale;
          for (int i = 0; i < n_stored; ++i) { aug[i][n_stored] = -1.0; aug[n_stored][i] = -1.0; aug[i][nb] = 0.0; }
          aug[n_stored][n_stored] = 0.0;
          aug[n_stored][nb] = -1.0;
          ok = 1;
          for (int i = 0; i < nb && ok; ++i) {
            int piv = i;
            for (int j = i + 1; j < nb; ++j) if (fabs(aug[j][i]) > fabs(aug[piv][i])) piv = j;
            if (piv != i) for (int c = 0; c <= nb; ++c) { const double t = aug[i][c]; aug[i][c] = aug[piv][c]; aug[piv][c] = t; }
            const double pivot = aug[i][i];
            if (fabs(pivot) < SCF_PIVOT_FLOOR) { ok = 0; break; }
            for (int j = i + 1; j < nb; ++j) {
              const double factor = aug[j][i] / pivot;
              for (int c = i; c <= nb; ++c) aug[j][c] -= factor * aug[i][c];
            }
          }
          if (ok) {
            double coef[9];
            for (int i = nb - 1; i >= 0; --i) {
              double sum = 0.0;
              for (int c = i + 1; c < nb; ++c) sum += aug[i][c] * coef[c];
              coef[i] = (aug[i][nb] - sum) / aug[i][i];
            }
            for (int i = 0; i < n_stored; ++i) sh.coef[i] = coef[i];
          }
        }
        sh.flag = ok;
        p.state[0] = n_stored;
        p.state[1] = newest;
        p.scalars[4] = ok ? 1.0 : 0.0;
      }
      __syncthreads();
      if (sh.flag) {                                     // fock = sum_i c_i F_i, oldest first (mqc_diis.f90:131-137)
        for (int e = tid; e < nn; e += nth) {
          double acc = 0.0;
          for (int i = 0; i < n_stored; ++i) {
            const int si = ((newest - n_stored + i) % dmax + dmax) % dmax;
            acc = acc + sh.coef[i] * p.diis_f[(size_t)si * nn + e];
          }
          w3[e] = acc;
        }
        __syncthreads();
        f_use = w3;
      }
    }
  }

  // ---- diagonalize (rhf.f90:1464-1489): F' = X^T F X, eigenvectors, C = X C'
  blk_gemm(n, m, n, 1.0, f_use, n, false, p.x, n, false, w0, n);                 // F X
  blk_gemm(m, m, n, 1.0, p.x, n, true, w0, n, false, sh.a, SCF_LD);              // X^T F X, into shared memory
  jacobi_eigh(sh.a, sh.v, m, sh.red, sh.pq, sh.cs, &sh.rotated);
  sort_eigenvalues(sh.a, m, sh.order);
  for (int k = tid; k < m; k += nth) p.eps[k] = sh.a[sh.order[k] + SCF_LD * sh.order[k]];
  // C' with sorted columns, into w1 (m x m), then C = X C'
  for (int e = tid; e < mm; e += nth) { const int i = e % m, k = e / m; w1[e] = sh.v[i + SCF_LD * sh.order[k]]; }
  __syncthreads();
  blk_gemm(n, m, m, 1.0, p.x, n, false, w1, m, false, p.coeff, n);
  // ---- density D = 2 C_occ C_occ^T and rms(dD)
  blk_gemm(n, n, p.n_occ, 2.0, p.coeff, n, false, p.coeff, n, true, w2, n);
  double part = 0.0;
  for (int e = tid; e < nn; e += nth) { const double d = w2[e] - p.density[e]; part = fma(d, d, part); }
  const double ss = blk_sum(part, sh.red);
  for (int e = tid; e < nn; e += nth) p.density[e] = w2[e];
  if (tid == 0 && p.mode == 1) {
    const double e_elec = p.scalars[0];
    const double de = fabs(e_elec - p.scalars[1]);
    const double drms = sqrt(ss / (double)nn);
    p.scalars[1] = e_elec;
    p.scalars[2] = de;
    p.scalars[3] = drms;
    const int iter = p.state[2] + 1;
    p.state[2] = iter;
    if (iter > 1 && de < p.energy_tol && drms < p.density_tol) p.state[3] = 1;
  }
}

size_t scf_shared_bytes() { return sizeof(ScfShared); }

void configure_scf_kernels() {
  cudaFuncSetAttribute(scf_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ScfShared));
  cudaFuncSetAttribute(scf_orthogonalizer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ScfShared));
}

bool scf_path_applies(int n) { return n >= 1 && n <= SCF_MAX_N; }

void launch_scf_orthogonalizer(const double *d_s, int n, double *d_x, int *d_n_mo, cudaStream_t s) {
  scf_orthogonalizer_kernel<<<1, SCF_THREADS, sizeof(ScfShared), s>>>(d_s, n, d_x, d_n_mo);
}

void launch_scf_step(const ScfStepLaunch &a, cudaStream_t s) {
  ScfStepArgs p;
  p.n = a.n; p.n_mo = a.n_mo; p.n_occ = a.n_occ; p.diis_max = a.diis_max; p.mode = a.mode; p.guess = a.guess;
  p.h = a.h; p.s = a.s; p.x = a.x; p.fock = a.fock; p.density = a.density; p.coeff = a.coeff; p.eps = a.eps;
  p.work = a.work; p.diis_f = a.diis_f; p.diis_e = a.diis_e; p.diis_b = a.diis_b; p.state = a.state; p.scalars = a.scalars;
  p.energy_tol = a.energy_tol; p.density_tol = a.density_tol;
  scf_step_kernel<<<1, SCF_THREADS, sizeof(ScfShared), s>>>(p);
}

}  // namespace mqcb200

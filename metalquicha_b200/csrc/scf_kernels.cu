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
//
// Measured (B200, n = 72): a Jacobi round costs ~2700 clocks, two thirds of it the shared-memory
// traffic of rotating A and V (every round reads and writes both matrices once: 250 KB at 128 B/clk
// of ONE SM), so an iteration's step is ~0.4 ms once the warm start has cut the sweeps to 2-4;
// the whole iteration (build + step) is ~0.84 ms against ~1.0 ms for the host-driven loop.
#include "common.cuh"
#include "kernels.cuh"

namespace mqcb200 {

constexpr int SCF_THREADS = 512;
constexpr int SCF_MAX_N = 80;
constexpr int SCF_LD = 81;                 // odd leading dimension in shared memory: conflict-free rows AND columns
constexpr double SCF_LINDEP_TOL = 1.0e-7;  // LINEAR_DEPENDENCE_TOL, mqc_scf_common.f90:27
constexpr double SCF_GWH_K = 1.75;         // mqc_scf_common.f90:33
constexpr double SCF_PIVOT_FLOOR = 1.0e-14;   // mqc_diis.f90:35

// C(M x N, ldc) = alpha * op(A) * op(B); op(A) is M x K, all dimensions <= 80.  Column-major;
// ta/tb = transpose flags.  Block-cooperative: op(A) and op(B) are first staged in shared memory
// (sa[i + LD k], sb[k + LD j]; the global reads are coalesced whichever way the operand lies, the
// odd leading dimension keeps the transposing writes conflict-free), then every thread owns a
// strided 4x4 set of outputs -- rows i, i+tm, i+2tm, i+3tm -- so that the lanes of a warp read
// consecutive words of sa and broadcast words of sb.  Operands may have been written by this
// block (plain loads).  C may live in global or shared memory, but not in sa/sb.
__device__ void blk_gemm(int M, int N, int K, double alpha, const double *A, int lda, bool ta, const double *B, int ldb,
                         bool tb, double *C, int ldc, double *sa, double *sb) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n_warps = blockDim.x >> 5;
  // Staging: a warp copies whole columns of the operand as it lies in memory (coalesced), every
  // lane fetching all of its elements of up to five columns BEFORE the first store -- the loads
  // would otherwise be serialised behind the stores (the compiler must assume sa aliases A).
  auto stage = [&](const double *src, int ld, int rows, int cols, double *dst, bool transpose_into) {
    // element (r, c) of src goes to dst[r + LD c] (or dst[c + LD r] when transposing)
    for (int c0 = warp; c0 < cols; c0 += 5 * n_warps) {
      double v[5][3];
#pragma unroll
      for (int u = 0; u < 5; ++u) {
        const int c = c0 + u * n_warps;
#pragma unroll
        for (int w = 0; w < 3; ++w) {
          const int r = lane + 32 * w;
          if (c < cols && r < rows) v[u][w] = src[(size_t)r + (size_t)ld * c];
        }
      }
#pragma unroll
      for (int u = 0; u < 5; ++u) {
        const int c = c0 + u * n_warps;
#pragma unroll
        for (int w = 0; w < 3; ++w) {
          const int r = lane + 32 * w;
          if (c < cols && r < rows) dst[transpose_into ? c + SCF_LD * r : r + SCF_LD * c] = v[u][w];
        }
      }
    }
  };
  // sa[i + LD k] = op(A)(i, k): A is M x K as stored when !ta, K x M when ta
  if (!ta) stage(A, lda, M, K, sa, false); else stage(A, lda, K, M, sa, true);
  // sb[k + LD j] = op(B)(k, j): B is K x N as stored when !tb, N x K when tb
  if (!tb) stage(B, ldb, K, N, sb, false); else stage(B, ldb, N, K, sb, true);
  __syncthreads();
  const int tm = (M + 3) / 4, tn = (N + 3) / 4;
  for (int t = threadIdx.x; t < tm * tn; t += blockDim.x) {
    const int ti = t % tm, tj = t / tm;
    int ri[4], cj[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      ri[r] = ti + r * tm < M ? ti + r * tm : M - 1;       // ragged edge: computed and dropped
      cj[r] = tj + r * tn < N ? tj + r * tn : N - 1;
    }
    double acc[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[r][c] = 0.0;
    for (int k = 0; k < K; ++k) {
      double av[4], bv[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) { av[r] = sa[ri[r] + SCF_LD * k]; bv[r] = sb[k + SCF_LD * cj[r]]; }
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = fma(av[r], bv[c], acc[r][c]);
    }
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c)
        if (ti + r * tm < M && tj + c * tn < N) C[(size_t)(ti + r * tm) + (size_t)ldc * (tj + c * tn)] = alpha * acc[r][c];
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
// disjoint rows/columns and are applied in parallel.  A warp owns up to three rotations of a
// round: it derives their (c, s) itself (every lane the same arithmetic -- they only read
// columns p and q, which nobody else writes in this phase), rotates those columns of A and V,
// and after ONE barrier rotates the rows p and q of A with the (c, s) still in its registers.
// Returns the number of sweeps.
constexpr int SCF_JW = 3;      // rotations per warp and round: ceil(40 / 16 warps)

__device__ int jacobi_eigh(double *a_s, double *v_s, int m, double *red, int *rotated_s) {
  const int tid = threadIdx.x, nth = blockDim.x;
  const int m_e = (m + 1) & ~1, half = m_e / 2;
  for (int e = tid; e < m * m; e += nth) {
    const int i = e % m, j = e / m;
    v_s[i + SCF_LD * j] = i == j ? 1.0 : 0.0;
  }
  double fro = 0.0;
  for (int e = tid; e < m * m; e += nth) { const double x = a_s[(e % m) + SCF_LD * (e / m)]; fro = fma(x, x, fro); }
  fro = blk_sum(fro, red);
  // an off-diagonal element below 2e-16 |A|_F is left alone (eigenvectors then carry errors of that size
  // over the gap); a sweep that rotates nothing ends the iteration
  const double skip = 2.0e-16 * sqrt(fro);
  const int warp = tid >> 5, lane = tid & 31, n_warps = nth >> 5;
  int sweeps = 0;
  for (int sweep = 0; sweep < 40; ++sweep) {
    ++sweeps;
    if (tid == 0) *rotated_s = 0;
    __syncthreads();
    int any = 0;
    int slot[SCF_JW];                                     // position of this warp's pairs in the round-robin sequence
#pragma unroll
    for (int u = 0; u < SCF_JW; ++u) slot[u] = warp + n_warps * u;
    // sequence position -> index: pos 0 is index 0, pos k >= 1 is 1 + (k - 1 + r) mod (m_e - 1); kept incrementally
    int ia[SCF_JW], ib[SCF_JW];
#pragma unroll
    for (int u = 0; u < SCF_JW; ++u) {
      ia[u] = slot[u] == 0 ? 0 : 1 + (slot[u] - 1) % (m_e - 1);
      ib[u] = 1 + (m_e - 2 - slot[u] + (m_e - 1)) % (m_e - 1);
    }
    for (int r = 0; r < m_e - 1; ++r) {
      int pp[SCF_JW], qq[SCF_JW];
      double cc[SCF_JW], ss[SCF_JW];
#pragma unroll
      for (int u = 0; u < SCF_JW; ++u) {
        const int p = ia[u] < ib[u] ? ia[u] : ib[u], q = ia[u] < ib[u] ? ib[u] : ia[u];
        pp[u] = p; qq[u] = q; cc[u] = 1.0; ss[u] = 0.0;
        if (slot[u] < half && q < m) {
          const double apq = a_s[p + SCF_LD * q];
          if (fabs(apq) > skip) {
            // t = sgn(tau) / (|tau| + sqrt(1 + tau^2)), tau = (aqq - app) / (2 apq), written with one
            // square root and one division; c = (1 + t^2)^-1/2
            const double d = a_s[q + SCF_LD * q] - a_s[p + SCF_LD * p];
            const double num = d >= 0.0 ? 2.0 * apq : -2.0 * apq;
            const double t = num / (fabs(d) + sqrt(fma(d, d, 4.0 * apq * apq)));
            const double c = rsqrt(fma(t, t, 1.0));
            cc[u] = c; ss[u] = t * c;
            any = 1;
          }
        }
      }
      __syncwarp();                                        // every lane has read a_pp, a_qq, a_pq before any lane rotates
      // columns: (A[:,p], A[:,q]) <- (c A[:,p] - s A[:,q], s A[:,p] + c A[:,q]); same for V
#pragma unroll
      for (int u = 0; u < SCF_JW; ++u) {
        if (ss[u] == 0.0) continue;
        const int p = pp[u], q = qq[u];
        const double c = cc[u], s = ss[u];
        double xp[3], xq[3], vp[3], vq[3];
#pragma unroll
        for (int w = 0; w < 3; ++w) {
          const int i = lane + 32 * w;
          if (i < m) {
            xp[w] = a_s[i + SCF_LD * p]; xq[w] = a_s[i + SCF_LD * q];
            vp[w] = v_s[i + SCF_LD * p]; vq[w] = v_s[i + SCF_LD * q];
          }
        }
#pragma unroll
        for (int w = 0; w < 3; ++w) {
          const int i = lane + 32 * w;
          if (i < m) {
            a_s[i + SCF_LD * p] = c * xp[w] - s * xq[w];
            a_s[i + SCF_LD * q] = s * xp[w] + c * xq[w];
            v_s[i + SCF_LD * p] = c * vp[w] - s * vq[w];
            v_s[i + SCF_LD * q] = s * vp[w] + c * vq[w];
          }
        }
      }
      __syncthreads();
      // rows of A (stride SCF_LD is odd: the lanes hit distinct banks)
#pragma unroll
      for (int u = 0; u < SCF_JW; ++u) {
        if (ss[u] == 0.0) continue;
        const int p = pp[u], q = qq[u];
        const double c = cc[u], s = ss[u];
        double xp[3], xq[3];
#pragma unroll
        for (int w = 0; w < 3; ++w) {
          const int j = lane + 32 * w;
          if (j < m) { xp[w] = a_s[p + SCF_LD * j]; xq[w] = a_s[q + SCF_LD * j]; }
        }
#pragma unroll
        for (int w = 0; w < 3; ++w) {
          const int j = lane + 32 * w;
          if (j < m) {
            a_s[p + SCF_LD * j] = c * xp[w] - s * xq[w];
            a_s[q + SCF_LD * j] = s * xp[w] + c * xq[w];
          }
        }
      }
      __syncthreads();
      // next round: every position but 0 moves on by one
#pragma unroll
      for (int u = 0; u < SCF_JW; ++u) {
        if (slot[u] != 0) ia[u] = ia[u] == m_e - 1 ? 1 : ia[u] + 1;
        ib[u] = ib[u] == m_e - 1 ? 1 : ib[u] + 1;
      }
    }
    if (any && lane == 0) *rotated_s = 1;
    __syncthreads();
    const int rotated = *rotated_s;
    __syncthreads();                                       // everyone has read it before the next sweep resets it
    if (rotated == 0) break;
  }
  return sweeps;
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
  double ga[SCF_MAX_N * SCF_LD];       // operand staging of blk_gemm
  double gb[SCF_MAX_N * SCF_LD];
  double red[SCF_THREADS / 32 + 2];
  double coef[16];
  int order[SCF_MAX_N];
  int flag;
  int rotated;
};

// X = U s^-1/2 over the eigenvalues of S above 1e-7 (ascending order kept): x (n x n_mo), *n_mo_out.
__global__ void __launch_bounds__(SCF_THREADS, 1) scf_orthogonalizer_kernel(const double *__restrict__ s, int n,
                                                                            double *__restrict__ x,
                                                                            int *__restrict__ n_mo_out, size_t block_stride) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  ScfShared &sh = *reinterpret_cast<ScfShared *>(smem_raw);
  s += (size_t)blockIdx.x * block_stride;            // one CTA per fragment of a batch
  x += (size_t)blockIdx.x * block_stride;
  n_mo_out += (size_t)blockIdx.x * block_stride * 2; // (ints inside a block of doubles)
  for (int e = threadIdx.x; e < n * n; e += blockDim.x) sh.a[(e % n) + SCF_LD * (e / n)] = s[e];
  __syncthreads();
  jacobi_eigh(sh.a, sh.v, n, sh.red, &sh.rotated);
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
  size_t block_stride;      // doubles between the state blocks of consecutive fragments (one CTA each)
  const int *n_mo_dev;      // non-null: read n_mo from the device (per fragment) instead of the argument
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
  if (blockIdx.x != 0) {                               // fragment of a batch: its own state block
    const size_t o = (size_t)blockIdx.x * p.block_stride;
    p.h += o; p.s += o; p.x += o; p.fock += o; p.density += o; p.coeff += o; p.eps += o; p.work += o;
    p.diis_f += o; p.diis_e += o; p.diis_b += o; p.scalars += o;
    p.state += 2 * o;
    if (p.n_mo_dev) p.n_mo_dev += 2 * o;
  }
  if (p.n_mo_dev) p.n_mo = *p.n_mo_dev;
  const int n = p.n, m = p.n_mo, nn = n * n, mm = m * m;
  if (m <= 0 || p.n_occ > m) {                         // batch: a fragment whose basis collapsed is reported, not iterated
    if (tid == 0 && p.mode == 1) p.state[3] = -1;
    return;
  }
  if (p.mode == 1 && p.state[3] != 0) return;            // already converged: later queued iterations are no-ops
  double *w0 = p.work, *w1 = w0 + nn, *w2 = w1 + nn, *w3 = w2 + nn;
  const double *f_use = p.fock;
  long long t_mark = clock64();
  auto lap = [&](int slot) {               // development instrumentation: clocks per stage, scalars[8 + slot]
    if (tid == 0) { const long long now = clock64(); p.scalars[8 + slot] = (double)(now - t_mark); t_mark = now; }
  };

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
    blk_gemm(n, n, n, 1.0, p.fock, n, false, p.density, n, false, w0, n, sh.ga, sh.gb);        // F D
    blk_gemm(n, n, n, 1.0, w0, n, false, p.s, n, false, w1, n, sh.ga, sh.gb);                  // F D S
    for (int e = tid; e < nn; e += nth) { const int i = e % n, j = e / n; w0[e] = w1[e] - w1[(size_t)j + (size_t)n * i]; }
    __syncthreads();
    blk_gemm(n, m, n, 1.0, w0, n, false, p.x, n, false, w2, n, sh.ga, sh.gb);                  // (..) X
    // ---- DIIS push (mqc_diis.f90:94-119): the ring slot after the newest; overlaps of the new entry
    int n_stored = p.state[0], newest = p.state[1];
    const int dmax = p.diis_max;
    if (dmax > 0) {
      newest = newest % dmax + 1;
      if (n_stored < dmax) n_stored += 1;
      const int slot = newest - 1;
      blk_gemm(m, m, n, 1.0, p.x, n, true, w2, n, false, p.diis_e + (size_t)slot * mm, m, sh.ga, sh.gb);   // X^T (..)
      for (int e = tid; e < nn; e += nth) p.diis_f[(size_t)slot * nn + e] = p.fock[e];
      __syncthreads();
      {
        // overlaps of the new error vector with every stored one, all in one pass over the new vector
        double part[8];
#pragma unroll
        for (int g = 0; g < 8; ++g) part[g] = 0.0;
        for (int e = tid; e < mm; e += nth) {
          const double en = p.diis_e[(size_t)slot * mm + e];
#pragma unroll
          for (int g = 0; g < 8; ++g)
            if (g < n_stored) part[g] = fma(en, p.diis_e[(size_t)(((newest - n_stored + g) % dmax + dmax) % dmax) * mm + e], part[g]);
        }
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          if (g >= n_stored) break;                                                          // uniform
          const int other = ((newest - n_stored + g) % dmax + dmax) % dmax;                  // diis_slot_of_age(g+1) - 1
          const double dot = blk_sum(part[g], sh.red);
          if (tid == 0) { p.diis_b[slot * dmax + other] = dot; p.diis_b[other * dmax + slot] = dot; }
        }
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

  lap(0);
  // ---- diagonalize (rhf.f90:1464-1489): F' = Y^T F Y, eigenvectors V, C = Y V.  The reference takes
  // Y = X every time; any S-orthonormal basis of the same space gives the same C up to the
  // eigensolver's rounding, so an iteration takes Y = the PREVIOUS orbitals: F' is then nearly
  // diagonal and the Jacobi iteration needs two or three sweeps instead of nine.  (The guess step
  // has no previous orbitals and takes Y = X.)
  const double *y = p.mode == 0 ? p.x : p.coeff;
  blk_gemm(n, m, n, 1.0, f_use, n, false, y, n, false, w0, n, sh.ga, sh.gb);                   // F Y
  blk_gemm(m, m, n, 1.0, y, n, true, w0, n, false, sh.a, SCF_LD, sh.ga, sh.gb);                // Y^T F Y, into shared memory
  for (int e = tid; e < mm; e += nth) {                                          // exactly symmetric input
    const int i = e % m, j = e / m;
    if (i > j) { const double v = 0.5 * (sh.a[i + SCF_LD * j] + sh.a[j + SCF_LD * i]); sh.a[i + SCF_LD * j] = v; sh.a[j + SCF_LD * i] = v; }
  }
  __syncthreads();
  lap(1);
  const int sweeps = jacobi_eigh(sh.a, sh.v, m, sh.red, &sh.rotated);
  lap(2);
  if (tid == 0) p.scalars[5] = (double)sweeps;
  sort_eigenvalues(sh.a, m, sh.order);
  for (int k = tid; k < m; k += nth) p.eps[k] = sh.a[sh.order[k] + SCF_LD * sh.order[k]];
  // V with sorted columns, into w1 (m x m), then C = Y V (via w3: Y may be p.coeff itself)
  for (int e = tid; e < mm; e += nth) { const int i = e % m, k = e / m; w1[e] = sh.v[i + SCF_LD * sh.order[k]]; }
  __syncthreads();
  blk_gemm(n, m, m, 1.0, y, n, false, w1, m, false, w0, n, sh.ga, sh.gb);
  for (int e = tid; e < n * m; e += nth) p.coeff[e] = w0[e];
  __syncthreads();
  // ---- density D = 2 C_occ C_occ^T and rms(dD)
  blk_gemm(n, n, p.n_occ, 2.0, p.coeff, n, false, p.coeff, n, true, w2, n, sh.ga, sh.gb);
  double part = 0.0;
  for (int e = tid; e < nn; e += nth) { const double d = w2[e] - p.density[e]; part = fma(d, d, part); }
  const double ss = blk_sum(part, sh.red);
  for (int e = tid; e < nn; e += nth) p.density[e] = w2[e];
  lap(3);
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

void launch_scf_orthogonalizer(const double *d_s, int n, double *d_x, int *d_n_mo, cudaStream_t s, int batch,
                               size_t block_stride) {
  scf_orthogonalizer_kernel<<<batch, SCF_THREADS, sizeof(ScfShared), s>>>(d_s, n, d_x, d_n_mo, block_stride);
}

void launch_scf_step(const ScfStepLaunch &a, cudaStream_t s) {
  ScfStepArgs p;
  p.block_stride = a.block_stride; p.n_mo_dev = a.n_mo_dev;
  p.n = a.n; p.n_mo = a.n_mo; p.n_occ = a.n_occ; p.diis_max = a.diis_max; p.mode = a.mode; p.guess = a.guess;
  p.h = a.h; p.s = a.s; p.x = a.x; p.fock = a.fock; p.density = a.density; p.coeff = a.coeff; p.eps = a.eps;
  p.work = a.work; p.diis_f = a.diis_f; p.diis_e = a.diis_e; p.diis_b = a.diis_b; p.state = a.state; p.scalars = a.scalars;
  p.energy_tol = a.energy_tol; p.density_tol = a.density_tol;
  scf_step_kernel<<<a.batch, SCF_THREADS, sizeof(ScfShared), s>>>(p);
}

}  // namespace mqcb200

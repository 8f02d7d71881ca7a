// whiten_kernels.cu -- B = (mu nu|P) . J^(-1/2) on the device, straight into the packed
// layout (last stage of the reference's build_df_tensor: `pic_gemm(three, half, b)`,
// backends/libcint/mqc_libcint_integrals.F90:985-986).
//
// Only the packed lower-triangular positions of each slab are formed, so the GEMM is
//     Bp[Q'][p] = sum_P half[Q'][P] * Tp[P][p],      M = naux, N = L, K = naux,
// 2*naux^2*L flops instead of the reference's 2*naux^2*n^2 (about half).  Tp is the
// three-centre tensor packed exactly like Bp; the order of p inside a row is irrelevant
// to a contraction over P.  Same machinery as the exchange kernels: FP64 DMMA, operands
// staged by 1-D bulk TMA on an mbarrier ring, 8 consumer warps + a producer warpgroup.
//   A = half in fragment order [P/16][Q'/8][4][32]  -> one contiguous 16 KiB panel per stage
//   B = 16 rows of Tp, 128 columns each, row stride padded to 132 doubles in shared
//       memory so the k-major m8n8k4 B-fragment (t*132 + g) is bank-conflict-free.
//
// The GEMM runs SLAB BY SLAB over the packed columns (a slab = the tile-columns of a range of
// nu): the caller streams (mu nu|P) by nu-slab, so neither the full host tensor nor a second
// device copy of the packed tensor is ever needed.  With the build sharded over GPUs by
// auxiliary index, a rank whitens ITS slabs for ALL rows Q' and the kernel's epilogue stores each
// row straight into the packed tensor of the rank that owns it -- the all-to-all from mu-nu slabs
// to Q-slabs rides on the GEMM's own stores over NVLink peer memory (WhitenDst).
//
// metric^(-1/2) (metric_inverse_sqrt, mqc_libcint_integrals.F90:992-1038) is formed on the device
// too: one-sided (Hestenes) Jacobi on the columns of the symmetric metric -- column operations
// only, so every access is coalesced -- then half = (V s^-1/2) V^T with the general batched GEMM.
#include "common.cuh"
#include "kernels.cuh"

#include <cmath>
#include <cooperative_groups.h>

namespace mqcb200 {

struct WhitenCfg {
  static constexpr int kStages = 4;
  static constexpr int kAElems = 16 * 128;        // 16 m-blocks x (4 k-subs x 32)
  static constexpr int kBStride = 132;            // doubles; == 4 (mod 16)
  static constexpr int kBElems = 16 * kBStride;
  static constexpr int kStageElems = kAElems + kBElems;
  static constexpr size_t kSmemBytes = (size_t)kStages * kStageElems * sizeof(double) + 2 * kStages * sizeof(uint64_t);
};

// half (naux x naux, symmetric, column-major) -> Af[kc][mb][ks][g*4+t] = half[Q' = 8mb+g][P = 16kc+4ks+t]
__global__ void __launch_bounds__(256) pack_half_kernel(const double *__restrict__ half, int naux, int row0, int m_rows,
                                                        int nmb, int nkc, double *__restrict__ af) {
  const size_t total = (size_t)nkc * nmb * 128;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const int t = e & 3, g = (e >> 2) & 7, ks = (e >> 5) & 3;
    const size_t blk = e >> 7;
    const int mb = (int)(blk % nmb), kc = (int)(blk / nmb);
    const int q = mb * 8 + g, p = kc * 16 + ks * 4 + t;
    af[e] = (q < m_rows && p < naux) ? half[(size_t)p + (size_t)naux * (row0 + q)] : 0.0;   // contiguous in p
  }
}

__global__ void __launch_bounds__(K_THREADS, 1)
whiten_gemm_kernel(const double *__restrict__ af, int nmb, int nkc, const double *__restrict__ tp, long long ld_tp,
                   int naux, int m_rows, WhitenDst dst) {
  using Cfg = WhitenCfg;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double *stage_base = reinterpret_cast<double *>(smem_raw);
  uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem_raw + (size_t)Cfg::kStages * Cfg::kStageElems * sizeof(double));
  uint64_t *empty_bar = full_bar + Cfg::kStages;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mt = blockIdx.x;                       // 128-row tile of Q' (fastest: neighbours share the Tp columns in L2)
  const long long p0 = (long long)blockIdx.y * 128;
  const int valid_a = nmb - 16 * mt < 16 ? nmb - 16 * mt : 16;

  for (int e = threadIdx.x; e < Cfg::kStages * Cfg::kStageElems; e += K_THREADS) stage_base[e] = 0.0;
  if (threadIdx.x == 0) {
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], K_CONSUMER_WARPS);
    }
    fence_mbar_init();
  }
  fence_proxy_async();
  __syncthreads();

  if (warp >= K_CONSUMER_WARPS) {
    reg_dealloc_producer();
    if (warp != K_CONSUMER_WARPS) return;
    // lanes 0..15 copy one Tp row each, lane 16 the half panel
    for (int kc = 0; kc < nkc; ++kc) {
      const int st = kc % Cfg::kStages;
      const uint32_t ph = (kc / Cfg::kStages) & 1;
      mbar_wait(&empty_bar[st], ph ^ 1);
      double *a_s = stage_base + (size_t)st * Cfg::kStageElems;
      const int rows = naux - 16 * kc < 16 ? naux - 16 * kc : 16;
      if (lane == 0)
        mbar_arrive_expect_tx(&full_bar[st], (uint32_t)(valid_a * 128 + rows * 128) * sizeof(double));
      __syncwarp();
      if (lane < rows) {
        tma_load_1d(a_s + Cfg::kAElems + lane * Cfg::kBStride, tp + (size_t)(16 * kc + lane) * ld_tp + p0,
                    128 * sizeof(double), &full_bar[st]);
      } else if (lane == 16) {
        tma_load_1d(a_s, af + ((size_t)kc * nmb + 16 * mt) * 128, (uint32_t)valid_a * 128 * sizeof(double),
                    &full_bar[st]);
      }
    }
    return;
  }

  reg_alloc_consumer();
  const int wm = warp & 3, wn = warp >> 2;
  const int g = lane >> 2, t = lane & 3;
  double acc[4][8][2];
#pragma unroll
  for (int m = 0; m < 4; ++m)
#pragma unroll
    for (int j = 0; j < 8; ++j) { acc[m][j][0] = 0.0; acc[m][j][1] = 0.0; }

  double fa[2][4], fb[2][8];
  auto load_frags = [&](double (&a)[4], double (&b)[8], int st, int ks) {
    const double *a_s = stage_base + (size_t)st * Cfg::kStageElems + (4 * wm) * 128;
    const double *b_s = stage_base + (size_t)st * Cfg::kStageElems + Cfg::kAElems + 64 * wn;
#pragma unroll
    for (int m = 0; m < 4; ++m) a[m] = a_s[(m * 4 + ks) * 32 + lane];
#pragma unroll
    for (int j = 0; j < 8; ++j) b[j] = b_s[(4 * ks + t) * Cfg::kBStride + 8 * j + g];
  };
  if (nkc > 0) {
    mbar_wait(&full_bar[0], 0);
    load_frags(fa[0], fb[0], 0, 0);
  }
  for (int kc = 0; kc < nkc; ++kc) {
    const int st = kc % Cfg::kStages;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      if (ks < 3) {
        load_frags(fa[(ks + 1) & 1], fb[(ks + 1) & 1], st, ks + 1);
      } else if (kc + 1 < nkc) {
        const int st1 = (kc + 1) % Cfg::kStages;
        mbar_wait(&full_bar[st1], ((kc + 1) / Cfg::kStages) & 1);
        load_frags(fa[0], fb[0], st1, 0);
      }
#pragma unroll
      for (int m = 0; m < 4; ++m)
#pragma unroll
        for (int j = 0; j < 8; ++j) dmma884(acc[m][j][0], acc[m][j][1], fa[ks & 1][m], fb[ks & 1][j]);
    }
    release_stage(&empty_bar[st], lane);
  }

  // epilogue: row q goes to the packed tensor of the rank that owns it (possibly over NVLink)
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    const int q = 128 * mt + 32 * wm + 8 * m + g;
    if (q >= m_rows) continue;
    int owner = 0;
    while (owner + 1 < dst.n_ranks && q >= dst.q_begin[owner + 1]) ++owner;
    double *row = dst.base[owner] + (size_t)(q - dst.q_begin[owner]) * dst.ld + dst.col_off;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const long long c = p0 + 64 * wn + 8 * j + 2 * t;
      *reinterpret_cast<double2 *>(row + c) = make_double2(acc[m][j][0], acc[m][j][1]);
    }
  }
}

void configure_whiten_kernels() {
  cudaFuncSetAttribute(whiten_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WhitenCfg::kSmemBytes);
}

size_t whiten_half_elems(int m_rows, int naux) {
  const int nmb = (m_rows + 7) / 8, nkc = (naux + 15) / 16;
  return (size_t)nkc * nmb * 128;
}

// af = rows [row0, row0 + m_rows) of half (naux x naux) in fragment order.
void launch_pack_half(const double *d_half, int naux, int row0, int m_rows, double *d_af, cudaStream_t s) {
  const int nmb = (m_rows + 7) / 8, nkc = (naux + 15) / 16;
  const size_t total = (size_t)nkc * nmb * 128;
  unsigned blocks = (unsigned)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  pack_half_kernel<<<blocks, 256, 0, s>>>(d_half, naux, row0, m_rows, nmb, nkc, d_af);
}

// Bp rows [0, m_rows) x packed columns [0, n_cols) = af . Tp_slab, each row stored at its owner (dst).
void launch_whiten_slab(const double *d_af, int m_rows, int naux, const double *d_tp, long long ld_tp, long long n_cols,
                        const WhitenDst &dst, cudaStream_t s) {
  const int nmb = (m_rows + 7) / 8, nkc = (naux + 15) / 16;
  const unsigned mtiles = (unsigned)((m_rows + 127) / 128);
  const long long ntiles = n_cols / 128;
  for (long long y0 = 0; y0 < ntiles; y0 += 65535) {
    const unsigned ny = (unsigned)(ntiles - y0 < 65535 ? ntiles - y0 : 65535);
    dim3 grid(mtiles, ny);
    WhitenDst d = dst;
    d.col_off += y0 * 128;
    whiten_gemm_kernel<<<grid, K_THREADS, WhitenCfg::kSmemBytes, s>>>(d_af, nmb, nkc, d_tp + y0 * 128, ld_tp, naux, m_rows, d);
  }
}

// A nu-slab of (mu nu|P) as it was shipped from the host -- only the rows the packed layout reads,
// mu >= nu_begin: d_slab[P][(mu - nu_begin) + (n - nu_begin) (nu - nu_begin)], P = 0..naux-1 --
// into packed fragment order: d_tp[P][tile columns tc0..tc1 of the packed row].
__global__ void __launch_bounds__(256) pack_slab_kernel(const double *__restrict__ slab, int n, int nt, int nu_begin,
                                                        int nu_count, int tc0, long long tile0, long long l_slab,
                                                        double *__restrict__ tp) {
  const int p = blockIdx.y;
  const long long tile = tile0 + blockIdx.x;             // global tile id in the packed row
  // tile id -> (tr, tc): columns tc0.. only
  int tc = tc0;
  long long off = tile0;
  while (tc + 1 < nt && off + (nt - tc) <= tile) { off += nt - tc; ++tc; }
  const int tr = tc + (int)(tile - off);
  const int r = threadIdx.x & 15, c = threadIdx.x >> 4;
  int mu = tr * TILE + r, nu = tc * TILE + c;
  double v = 0.0;
  if (mu < n && nu < n) {
    if (mu < nu) { const int tmp = mu; mu = nu; nu = tmp; }   // diagonal tile: mirror the lower triangle
    const size_t rows = (size_t)(n - nu_begin);
    v = slab[(size_t)p * rows * nu_count + (size_t)(mu - nu_begin) + rows * (nu - nu_begin)];
  }
  tp[(size_t)p * l_slab + (size_t)blockIdx.x * TILE_ELEMS + in_tile_offset(r, c)] = v;
}

void launch_pack_slab(const double *d_slab, int n, int naux, int nu_begin, int nu_count, double *d_tp, cudaStream_t s) {
  const int nt = num_tiles(n);
  const int tc0 = nu_begin / TILE, tc1 = (nu_begin + nu_count + TILE - 1) / TILE;
  const long long tile0 = tile_index(tc0, tc0, nt);
  const long long tile1 = tc1 >= nt ? num_lower_tiles(nt) : tile_index(tc1, tc1, nt);
  const long long n_tiles = tile1 - tile0;
  for (int p0 = 0; p0 < naux; p0 += 65535) {
    const int pc = naux - p0 < 65535 ? naux - p0 : 65535;
    dim3 grid((unsigned)n_tiles, (unsigned)pc);
    pack_slab_kernel<<<grid, 256, 0, s>>>(d_slab + (size_t)p0 * (size_t)(n - nu_begin) * nu_count, n, nt, nu_begin, nu_count, tc0, tile0,
                                          n_tiles * TILE_ELEMS, d_tp + (size_t)p0 * n_tiles * TILE_ELEMS);
  }
}

// ------------------------------------------------------------------------------------------------
// metric^(-1/2): one-sided Jacobi on the columns of the symmetric metric
// ------------------------------------------------------------------------------------------------
// One CTA per pair (p, q) of the round (round-robin ordering, as in scf_kernels.cu): alpha = |g_p|^2,
// beta = |g_q|^2, gamma = g_p . g_q in fixed order; if |gamma| > tol sqrt(alpha beta) the two columns
// of G and of V are rotated so that g_p . g_q = 0.  tol = 4 * 2^-53 * sqrt(n): the rounding error of an
// n-term dot product -- a smaller threshold only chases that noise (18 sweeps instead of ~11 at n = 1800).  At convergence G = V diag(lambda): the columns of V
// are the eigenvectors, lambda_i = v_i . g_i.
__global__ void __launch_bounds__(256) hestenes_round_kernel(double *__restrict__ g, double *__restrict__ v, int n,
                                                             int round, int *__restrict__ rotated, double tol) {
  const int n_e = (n + 1) & ~1;
  const int k = blockIdx.x;
  const int pa = k == 0 ? 0 : 1 + (k - 1 + round) % (n_e - 1);
  const int pb = 1 + (n_e - 2 - k + round) % (n_e - 1);
  const int p = pa < pb ? pa : pb, q = pa < pb ? pb : pa;
  if (q >= n) return;
  double *gp = g + (size_t)p * n, *gq = g + (size_t)q * n;
  double a = 0.0, b = 0.0, c = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) {
    const double x = gp[i], y = gq[i];
    a = fma(x, x, a); b = fma(y, y, b); c = fma(x, y, c);
  }
  __shared__ double red[3][8];
  __shared__ double cs[2];
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
    c += __shfl_xor_sync(0xffffffffu, c, o);
  }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = a; red[1][threadIdx.x >> 5] = b; red[2][threadIdx.x >> 5] = c; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double al = 0.0, be = 0.0, ga = 0.0;
    for (int w = 0; w < 8; ++w) { al += red[0][w]; be += red[1][w]; ga += red[2][w]; }
    double cc = 1.0, ss = 0.0;
    if (fabs(ga) > tol * sqrt(al * be) && fabs(ga) > 0.0) {
      const double d = be - al;
      const double num = d >= 0.0 ? 2.0 * ga : -2.0 * ga;
      const double t = num / (fabs(d) + sqrt(fma(d, d, 4.0 * ga * ga)));
      cc = rsqrt(fma(t, t, 1.0));
      ss = t * cc;
      *rotated = 1;
    }
    cs[0] = cc; cs[1] = ss;
  }
  __syncthreads();
  const double cc = cs[0], ss = cs[1];
  if (ss == 0.0) return;
  double *vp = v + (size_t)p * n, *vq = v + (size_t)q * n;
  for (int i = threadIdx.x; i < n; i += 256) {
    const double x = gp[i], y = gq[i];
    gp[i] = cc * x - ss * y;
    gq[i] = ss * x + cc * y;
    const double u = vp[i], w = vq[i];
    vp[i] = cc * u - ss * w;
    vq[i] = ss * u + cc * w;
  }
}

// The whole iteration in ONE cooperative launch: every round is followed by a grid-wide barrier instead
// of a kernel boundary, and the "did this sweep rotate anything?" test stays on the device -- no host
// synchronisation inside an eigensolve.  A CTA takes the pairs k = blockIdx.x, += gridDim.x.
// state[0]: rotation flag of the current sweep, state[1]: sweeps done (written at the end).
__global__ void __launch_bounds__(256) hestenes_solve_kernel(double *__restrict__ g, double *__restrict__ v, int n,
                                                             double tol, int max_sweeps, int *__restrict__ state) {
  namespace cg = cooperative_groups;
  cg::grid_group grid = cg::this_grid();
  const int n_e = (n + 1) & ~1, half = n_e / 2;
  __shared__ double red[3][8];
  __shared__ double cs[2];
  int sweeps = 0;
  for (int sweep = 0; sweep < max_sweeps; ++sweep) {
    ++sweeps;
    if (blockIdx.x == 0 && threadIdx.x == 0) state[0] = 0;
    grid.sync();
    for (int round = 0; round < n_e - 1; ++round) {
      for (int k = blockIdx.x; k < half; k += gridDim.x) {
        const int pa = k == 0 ? 0 : 1 + (k - 1 + round) % (n_e - 1);
        const int pb = 1 + (n_e - 2 - k + round) % (n_e - 1);
        const int p = pa < pb ? pa : pb, q = pa < pb ? pb : pa;
        if (q >= n) continue;                                   // (uniform per CTA)
        double *gp = g + (size_t)p * n, *gq = g + (size_t)q * n;
        double a = 0.0, b = 0.0, c = 0.0;
        for (int i = threadIdx.x; i < n; i += 256) {
          const double x = gp[i], y = gq[i];
          a = fma(x, x, a); b = fma(y, y, b); c = fma(x, y, c);
        }
        for (int o = 16; o > 0; o >>= 1) {
          a += __shfl_xor_sync(0xffffffffu, a, o);
          b += __shfl_xor_sync(0xffffffffu, b, o);
          c += __shfl_xor_sync(0xffffffffu, c, o);
        }
        __syncthreads();                                        // the previous pair's readers of red/cs are done
        if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = a; red[1][threadIdx.x >> 5] = b; red[2][threadIdx.x >> 5] = c; }
        __syncthreads();
        if (threadIdx.x == 0) {
          double al = 0.0, be = 0.0, ga = 0.0;
          for (int w = 0; w < 8; ++w) { al += red[0][w]; be += red[1][w]; ga += red[2][w]; }
          double cc = 1.0, ss = 0.0;
          if (fabs(ga) > tol * sqrt(al * be) && fabs(ga) > 0.0) {
            const double d = be - al;
            const double num = d >= 0.0 ? 2.0 * ga : -2.0 * ga;
            const double t = num / (fabs(d) + sqrt(fma(d, d, 4.0 * ga * ga)));
            cc = rsqrt(fma(t, t, 1.0));
            ss = t * cc;
            state[0] = 1;
          }
          cs[0] = cc; cs[1] = ss;
        }
        __syncthreads();
        const double cc = cs[0], ss = cs[1];
        if (ss != 0.0) {
          double *vp = v + (size_t)p * n, *vq = v + (size_t)q * n;
          for (int i = threadIdx.x; i < n; i += 256) {
            const double x = gp[i], y = gq[i];
            gp[i] = cc * x - ss * y;
            gq[i] = ss * x + cc * y;
            const double u = vp[i], w = vq[i];
            vp[i] = cc * u - ss * w;
            vq[i] = ss * u + cc * w;
          }
        }
      }
      grid.sync();
    }
    const int rotated = *reinterpret_cast<volatile int *>(state);
    grid.sync();                                                // everyone has read the flag before the next sweep clears it
    if (!rotated) break;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) state[1] = sweeps;
}

__global__ void __launch_bounds__(256) set_identity_kernel(double *__restrict__ v, int n) {
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < (size_t)n * n; e += (size_t)gridDim.x * blockDim.x)
    v[e] = (e % n) == (e / n) ? 1.0 : 0.0;
}

// lambda_i = v_i . g_i; scaled(:, i) = v_i / sqrt(lambda_i) if lambda_i > threshold else 0  (integrals.F90:1027-1033)
__global__ void __launch_bounds__(256) metric_scale_kernel(const double *__restrict__ g, const double *__restrict__ v, int n,
                                                           double threshold, double *__restrict__ scaled,
                                                           double *__restrict__ lambda, int *__restrict__ n_kept) {
  const int i = blockIdx.x;
  double s = 0.0;
  for (int r = threadIdx.x; r < n; r += 256) s = fma(v[(size_t)i * n + r], g[(size_t)i * n + r], s);
  __shared__ double red[8];
  __shared__ double lam;
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += red[w];
    lam = t;
    lambda[i] = t;
    if (t > threshold) atomicAdd(n_kept, 1);
  }
  __syncthreads();
  const double f = lam > threshold ? 1.0 / sqrt(lam) : 0.0;
  for (int r = threadIdx.x; r < n; r += 256) scaled[(size_t)i * n + r] = f * v[(size_t)i * n + r];
}

void launch_hestenes_round(double *d_g, double *d_v, int n, int round, int *d_rotated, cudaStream_t s) {
  const int n_e = (n + 1) & ~1;
  const double tol = 4.0 * 1.1102230246251565e-16 * sqrt((double)(n > 16 ? n : 16));
  hestenes_round_kernel<<<n_e / 2, 256, 0, s>>>(d_g, d_v, n, round, d_rotated, tol);
}

// Returns false when a cooperative launch is not possible (the caller then falls back to one kernel per round).
bool launch_hestenes_solve(double *d_g, double *d_v, int n, int max_sweeps, int *d_state, cudaStream_t s) {
  static int max_blocks = -1;
  if (max_blocks < 0) {
    int dev = 0, coop = 0, sms = 0, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, hestenes_solve_kernel, 256, 0);
    max_blocks = coop ? sms * per_sm : 0;
  }
  if (max_blocks <= 0) return false;
  const int n_e = (n + 1) & ~1;
  // One CTA per pair.  Measured (B200): the grid-wide barrier costs more the more CTAs take part -- at 344 pairs
  // (n = 688) a round is ~7 us against ~9 us for a kernel per round, at 900 pairs (n = 1800) ~18 us against ~14 us;
  // a warp per pair (57 CTAs) was twice slower still (the pair's column walk is then one warp's latency chain).
  // So: one cooperative launch up to 512 pairs, one kernel per round beyond.
  if (n_e / 2 > 512 || n_e / 2 > max_blocks) return false;
  int blocks = n_e / 2;
  if (blocks < 1) blocks = 1;
  double tol = 4.0 * 1.1102230246251565e-16 * sqrt((double)(n > 16 ? n : 16));
  void *args[] = {&d_g, &d_v, &n, &tol, &max_sweeps, &d_state};
  return cudaLaunchCooperativeKernel(reinterpret_cast<void *>(hestenes_solve_kernel), dim3((unsigned)blocks), dim3(256), args, 0, s) ==
         cudaSuccess;
}

void launch_set_identity(double *d_v, int n, cudaStream_t s) {
  unsigned blocks = (unsigned)(((size_t)n * n + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  set_identity_kernel<<<blocks, 256, 0, s>>>(d_v, n);
}

void launch_metric_scale(const double *d_g, const double *d_v, int n, double threshold, double *d_scaled, double *d_lambda,
                         int *d_n_kept, cudaStream_t s) {
  metric_scale_kernel<<<n, 256, 0, s>>>(d_g, d_v, n, threshold, d_scaled, d_lambda, d_n_kept);
}

}  // namespace mqcb200

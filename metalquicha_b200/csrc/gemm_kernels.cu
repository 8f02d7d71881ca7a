// gemm_kernels.cu -- a general strided-batched FP64 GEMM on the DMMA pipe, for the contractions
// AROUND the hot path that have no special structure to exploit (the DF gradient densities,
// SURVEY 8f row 4; metric^(-1/2) = (U s^-1/2) U^T, row a5):
//
//     C[b] = alpha * op(A[b]) * op(B[b]) + beta * C[b],      b = 0 .. batch-1
//
// column-major operands with arbitrary leading dimensions and batch strides, op = identity or
// transpose.  64x64 tile per CTA, four warps (2x2) of 32x32 m8n8k4 accumulators, K in steps of
// 16 through shared memory (k-major rows of 68 doubles: 68 = 4 mod 16 makes the m8n8k4 operand
// fragments bank-conflict-free), next step prefetched into registers while this one is
// multiplied.  No TMA here: the strides are the caller's, not fragment order -- the hot
// kernels (k_kernels.cu) get their speed from owning the layout, this one is for everything else.
#include "common.cuh"
#include "kernels.cuh"

namespace mqcb200 {

constexpr int GB_TILE = 64;
constexpr int GB_K = 16;
constexpr int GB_LD = 68;
constexpr int GB_THREADS = 128;

// One thread's share of a 16 x 64 operand tile: 8 elements.  `contig_m` = the operand is contiguous
// along the tile's 64-long (m or n) axis; otherwise along k.
struct GbFetch {
  double v[8];
};

// op(X)(i, k) for the tile rows i0.., k0..: X is (rows x K) stored col-major when !trans (contiguous in
// i), or (K x rows) when trans (contiguous in k).
__device__ __forceinline__ void gb_fetch(const double *__restrict__ x, long long ld, bool trans, int rows, int K, int i0,
                                         int k0, GbFetch &f) {
  const int tid = threadIdx.x;
  if (!trans) {
    // lanes along i: thread handles i = tid % 64, k = tid / 64 + 2*u
    const int i = i0 + (tid & 63);
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int k = k0 + (tid >> 6) + 2 * u;
      f.v[u] = (i < rows && k < K) ? x[(size_t)i + (size_t)ld * k] : 0.0;
    }
  } else {
    // lanes along k: thread handles k = tid % 16, i = tid / 16 + 8*u
    const int k = k0 + (tid & 15);
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int i = i0 + (tid >> 4) + 8 * u;
      f.v[u] = (i < rows && k < K) ? x[(size_t)k + (size_t)ld * i] : 0.0;
    }
  }
}

__device__ __forceinline__ void gb_store(double *s /*[GB_K][GB_LD]*/, bool trans, const GbFetch &f) {
  const int tid = threadIdx.x;
  if (!trans) {
#pragma unroll
    for (int u = 0; u < 8; ++u) s[((tid >> 6) + 2 * u) * GB_LD + (tid & 63)] = f.v[u];
  } else {
#pragma unroll
    for (int u = 0; u < 8; ++u) s[(tid & 15) * GB_LD + (tid >> 4) + 8 * u] = f.v[u];
  }
}

__global__ void __launch_bounds__(GB_THREADS) dgemm_batched_kernel(int M, int N, int K, double alpha,
                                                                  const double *__restrict__ a, long long lda, long long sa,
                                                                  int ta, const double *__restrict__ b, long long ldb,
                                                                  long long sb, int tb, double beta, double *__restrict__ c,
                                                                  long long ldc, long long sc) {
  __shared__ double as[2][GB_K * GB_LD];
  __shared__ double bs[2][GB_K * GB_LD];
  const int batch = blockIdx.z;
  a += (size_t)batch * sa;
  b += (size_t)batch * sb;
  c += (size_t)batch * sc;
  const int m0 = blockIdx.x * GB_TILE, n0 = blockIdx.y * GB_TILE;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wm = warp & 1, wn = warp >> 1;
  const int g = lane >> 2, t = lane & 3;

  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }

  // op(A) is M x K: stored (M x K) when !ta -> contiguous along m; op(B)^T is N x K: B stored (K x N)
  // when !tb -> contiguous along k from the tile's point of view, i.e. "trans" for the fetcher
  GbFetch fa, fb;
  const int n_steps = (K + GB_K - 1) / GB_K;
  gb_fetch(a, lda, ta != 0, M, K, m0, 0, fa);
  gb_fetch(b, ldb, tb == 0, N, K, n0, 0, fb);
  gb_store(as[0], ta != 0, fa);
  gb_store(bs[0], tb == 0, fb);
  __syncthreads();
  for (int step = 0; step < n_steps; ++step) {
    const int cur = step & 1;
    if (step + 1 < n_steps) {
      gb_fetch(a, lda, ta != 0, M, K, m0, (step + 1) * GB_K, fa);
      gb_fetch(b, ldb, tb == 0, N, K, n0, (step + 1) * GB_K, fb);
    }
    const double *a_s = as[cur] + 32 * wm, *b_s = bs[cur] + 32 * wn;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      double af[4], bf[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) af[i] = a_s[(4 * ks + t) * GB_LD + 8 * i + g];
#pragma unroll
      for (int j = 0; j < 4; ++j) bf[j] = b_s[(4 * ks + t) * GB_LD + 8 * j + g];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
    }
    if (step + 1 < n_steps) {
      gb_store(as[cur ^ 1], ta != 0, fa);
      gb_store(bs[cur ^ 1], tb == 0, fb);
    }
    __syncthreads();
  }
  // accumulator (i, j): rows m0 + 32 wm + 8 i + g, columns n0 + 32 wn + 8 j + 2 t (+1)
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = m0 + 32 * wm + 8 * i + g;
    if (row >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int col = n0 + 32 * wn + 8 * j + 2 * t + e;
        if (col >= N) continue;
        double *p = c + (size_t)row + (size_t)ldc * col;
        const double v = alpha * acc[i][j][e];
        *p = beta == 0.0 ? v : fma(beta, *p, v);
      }
    }
  }
}

void launch_dgemm_batched(int M, int N, int K, double alpha, const double *a, long long lda, long long stride_a, bool ta,
                          const double *b, long long ldb, long long stride_b, bool tb, double beta, double *c,
                          long long ldc, long long stride_c, int batch, cudaStream_t s) {
  if (M <= 0 || N <= 0 || batch <= 0) return;
  for (int b0 = 0; b0 < batch; b0 += 65535) {
    const int nb = batch - b0 < 65535 ? batch - b0 : 65535;
    dim3 grid((unsigned)((M + GB_TILE - 1) / GB_TILE), (unsigned)((N + GB_TILE - 1) / GB_TILE), (unsigned)nb);
    dgemm_batched_kernel<<<grid, GB_THREADS, 0, s>>>(M, N, K, alpha, a + (size_t)b0 * stride_a, lda, stride_a, ta ? 1 : 0,
                                                     b + (size_t)b0 * stride_b, ldb, stride_b, tb ? 1 : 0, beta,
                                                     c + (size_t)b0 * stride_c, ldc, stride_c);
  }
}

// ---- small helpers of the gradient-density build -----------------------------------------------
// Packed slabs (q_count of them) -> full symmetric n x n slabs, column-major.
__global__ void __launch_bounds__(256) unpack_tensor_kernel(const double *__restrict__ packed, long long L, int n, int nt,
                                                            double *__restrict__ full) {
  const int q = blockIdx.y;
  const int tile = blockIdx.x;
  // tile id -> (tr, tc), column-major lower triangle
  int tc = 0, off = 0;
  while (tc + 1 < nt && off + (nt - tc) <= tile) { off += nt - tc; ++tc; }
  const int tr = tc + (tile - off);
  const int r = threadIdx.x & 15, cidx = threadIdx.x >> 4;
  const int mu = tr * TILE + r, nu = tc * TILE + cidx;
  if (mu >= n || nu >= n) return;
  const double v = packed[(size_t)q * L + (size_t)tile * TILE_ELEMS + in_tile_offset(r, cidx)];
  double *f = full + (size_t)q * n * n;
  f[(size_t)mu + (size_t)n * nu] = v;
  if (tr != tc) f[(size_t)nu + (size_t)n * mu] = v;
}

void launch_unpack_tensor(const double *d_packed, int n, int q_count, double *d_full, cudaStream_t s) {
  const int nt = num_tiles(n);
  const long long L = packed_row_len(n);
  for (int q0 = 0; q0 < q_count; q0 += 65535) {
    const int qc = q_count - q0 < 65535 ? q_count - q0 : 65535;
    dim3 grid((unsigned)num_lower_tiles(nt), (unsigned)qc);
    unpack_tensor_kernel<<<grid, 256, 0, s>>>(d_packed + (size_t)q0 * L, L, n, nt, d_full + (size_t)q0 * n * n);
  }
}

// gamma(:, :, p) = rho[p] * D - w * z(:, :, p)   for a chunk of auxiliary functions (z may be null == 0;
// with accumulate the result is added to what gamma holds: second spin channel)
__global__ void __launch_bounds__(256) gradient_gamma_kernel(const double *__restrict__ rho, const double *__restrict__ d,
                                                             const double *__restrict__ z, double w, size_t nn,
                                                             int accumulate, double *__restrict__ gamma) {
  const int p = blockIdx.y;
  const double r = rho ? rho[p] : 0.0;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < nn; e += (size_t)gridDim.x * blockDim.x) {
    double v = (d ? r * d[e] : 0.0) - (z ? w * z[(size_t)p * nn + e] : 0.0);
    if (accumulate) v += gamma[(size_t)p * nn + e];
    gamma[(size_t)p * nn + e] = v;
  }
}

void launch_gradient_gamma(const double *d_rho, const double *d_density, const double *d_z, double w, int n, int p_count,
                           bool accumulate, double *d_gamma, cudaStream_t s) {
  const size_t nn = (size_t)n * n;
  unsigned bx = (unsigned)((nn + 255) / 256);
  if (bx > 64) bx = 64;
  for (int p0 = 0; p0 < p_count; p0 += 65535) {
    const int pc = p_count - p0 < 65535 ? p_count - p0 : 65535;
    dim3 grid(bx, (unsigned)pc);
    gradient_gamma_kernel<<<grid, 256, 0, s>>>(d_rho ? d_rho + p0 : nullptr, d_density, d_z ? d_z + (size_t)p0 * nn : nullptr, w,
                                               nn, accumulate ? 1 : 0, d_gamma + (size_t)p0 * nn);
  }
}

}  // namespace mqcb200

// =================================================================================================
// Elementwise / reduction helpers of the general-size device-resident SCF step (engine.cu: scf_general)
// =================================================================================================
namespace mqcb200 {

// Generalised Wolfsberg-Helmholz starting Fock (rhf.f90:1354-1380) or the core guess (F = H).
__global__ void __launch_bounds__(256) scf_guess_kernel(const double *__restrict__ h, const double *__restrict__ s, int n,
                                                        int gwh, double *__restrict__ f) {
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < (size_t)n * n; e += (size_t)gridDim.x * blockDim.x) {
    const int i = (int)(e % n), j = (int)(e / n);
    double v = h[e];
    if (gwh && i != j) v = 0.5 * 1.75 * s[e] * (h[(size_t)i * n + i] + h[(size_t)j * n + j]);
    f[e] = v;
  }
}

// out = a - a^T  (n x n)
__global__ void __launch_bounds__(256) antisym_kernel(const double *__restrict__ a, int n, double *__restrict__ out) {
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < (size_t)n * n; e += (size_t)gridDim.x * blockDim.x) {
    const size_t i = e % n, j = e / n;
    out[e] = a[e] - a[j + (size_t)n * i];
  }
}

// g = (g + g^T)/2 + shift * 1, lower and upper written by the thread that owns (i >= j); shift read from the device
__global__ void __launch_bounds__(256) symmetrize_shift_kernel(double *__restrict__ g, int m, const double *__restrict__ shift) {
  const double sg = shift ? *shift : 0.0;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < (size_t)m * m; e += (size_t)gridDim.x * blockDim.x) {
    const size_t i = e % m, j = e / m;
    if (i < j) continue;
    const double v = 0.5 * (g[i + (size_t)m * j] + g[j + (size_t)m * i]) + (i == j ? sg : 0.0);
    g[i + (size_t)m * j] = v;
    g[j + (size_t)m * i] = v;
  }
}

// Fixed-order two-level reduction, one of: 0: sum a*b   1: sum (a-b)^2   2: sqrt(sum a*a).  Each block reduces a fixed slice;
// the block that finishes last adds the partials in index order.  scratch: 130 doubles, zero before first use (left zero).
__global__ void __launch_bounds__(256) reduce_kernel(const double *__restrict__ a, const double *__restrict__ b, size_t count,
                                                     int mode, double *__restrict__ scratch, double *__restrict__ out) {
  __shared__ double red[8];
  __shared__ bool last;
  double s = 0.0;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < count; e += (size_t)gridDim.x * blockDim.x) {
    if (mode == 0) s = fma(a[e], b[e], s);
    else if (mode == 1) { const double d = a[e] - b[e]; s = fma(d, d, s); }
    else s = fma(a[e], a[e], s);
  }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w];
    scratch[blockIdx.x] = t;
    __threadfence();
    last = atomicAdd(reinterpret_cast<unsigned int *>(scratch + 128), 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    __threadfence();
    double t = 0.0;
    for (unsigned int k = 0; k < gridDim.x; ++k) t += reinterpret_cast<volatile double *>(scratch)[k];
    out[0] = mode == 2 ? sqrt(t) : t;
    *reinterpret_cast<unsigned int *>(scratch + 128) = 0u;
  }
}

// lambda_i = v_i . g_i - shift  (one CTA per column; fixed-order)
__global__ void __launch_bounds__(256) eig_lambda_kernel(const double *__restrict__ g, const double *__restrict__ v, int m,
                                                         const double *__restrict__ shift, double *__restrict__ lambda) {
  const int i = blockIdx.x;
  double s = 0.0;
  for (int r = threadIdx.x; r < m; r += 256) s = fma(v[(size_t)i * m + r], g[(size_t)i * m + r], s);
  __shared__ double red[8];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += red[w];
    lambda[i] = t - (shift ? *shift : 0.0);
  }
}

// order[rank] = i for ascending lambda (ties by index); *n_dropped = number of lambda <= threshold
__global__ void __launch_bounds__(256) rank_sort_kernel(const double *__restrict__ lambda, int m, double threshold,
                                                        int *__restrict__ order, int *__restrict__ n_dropped) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const double wi = lambda[i];
  int rank = 0;
  for (int j = 0; j < m; ++j) {
    const double wj = lambda[j];
    rank += (wj < wi || (wj == wi && j < i)) ? 1 : 0;
  }
  order[rank] = i;
  if (n_dropped && wi <= threshold) atomicAdd(n_dropped, 1);
}

// dst(:, k) = src(:, order[k0 + k]) * (inv_sqrt ? 1/sqrt(lambda[order[k0+k]]) : 1),  k = 0..cols-1  (rows x cols, ld = rows)
__global__ void __launch_bounds__(256) gather_columns_kernel(const double *__restrict__ src, int rows, const int *__restrict__ order,
                                                             int k0, int cols, const double *__restrict__ lambda, int inv_sqrt,
                                                             double *__restrict__ dst) {
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < (size_t)rows * cols; e += (size_t)gridDim.x * blockDim.x) {
    const size_t r = e % rows, k = e / rows;
    const int c = order[k0 + (int)k];
    const double f = inv_sqrt ? 1.0 / sqrt(lambda[c]) : 1.0;
    dst[e] = f * src[r + (size_t)rows * c];
  }
}

// out = sum_i coef[i] * vec[slot[i]]  (count elements each; coef and slot on the device, oldest first)
__global__ void __launch_bounds__(256) lincomb_kernel(const double *__restrict__ vecs, size_t count, const double *__restrict__ coef,
                                                      const int *__restrict__ slots, int n_terms, double *__restrict__ out) {
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < count; e += (size_t)gridDim.x * blockDim.x) {
    double acc = 0.0;
    for (int i = 0; i < n_terms; ++i) acc = acc + coef[i] * vecs[(size_t)slots[i] * count + e];
    out[e] = acc;
  }
}

static unsigned grid_for(size_t count) {
  size_t b = (count + 255) / 256;
  return (unsigned)(b > 148 * 8 ? 148 * 8 : (b < 1 ? 1 : b));
}

void launch_scf_guess(const double *d_h, const double *d_s, int n, bool gwh, double *d_f, cudaStream_t s) {
  scf_guess_kernel<<<grid_for((size_t)n * n), 256, 0, s>>>(d_h, d_s, n, gwh ? 1 : 0, d_f);
}
void launch_antisym(const double *d_a, int n, double *d_out, cudaStream_t s) {
  antisym_kernel<<<grid_for((size_t)n * n), 256, 0, s>>>(d_a, n, d_out);
}
void launch_symmetrize_shift(double *d_g, int m, const double *d_shift, cudaStream_t s) {
  symmetrize_shift_kernel<<<grid_for((size_t)m * m), 256, 0, s>>>(d_g, m, d_shift);
}
void launch_reduce(const double *d_a, const double *d_b, size_t count, int mode, double *d_scratch, double *d_out, cudaStream_t s) {
  unsigned blocks = (unsigned)((count + 4095) / 4096);
  if (blocks > 128) blocks = 128;
  if (blocks < 1) blocks = 1;
  reduce_kernel<<<blocks, 256, 0, s>>>(d_a, d_b, count, mode, d_scratch, d_out);
}
void launch_eig_lambda(const double *d_g, const double *d_v, int m, const double *d_shift, double *d_lambda, cudaStream_t s) {
  eig_lambda_kernel<<<m, 256, 0, s>>>(d_g, d_v, m, d_shift, d_lambda);
}
void launch_rank_sort(const double *d_lambda, int m, double threshold, int *d_order, int *d_n_dropped, cudaStream_t s) {
  rank_sort_kernel<<<(m + 255) / 256, 256, 0, s>>>(d_lambda, m, threshold, d_order, d_n_dropped);
}
void launch_gather_columns(const double *d_src, int rows, const int *d_order, int k0, int cols, const double *d_lambda,
                           bool inv_sqrt, double *d_dst, cudaStream_t s) {
  if (cols <= 0) return;
  gather_columns_kernel<<<grid_for((size_t)rows * cols), 256, 0, s>>>(d_src, rows, d_order, k0, cols, d_lambda, inv_sqrt ? 1 : 0, d_dst);
}
void launch_lincomb(const double *d_vecs, size_t count, const double *d_coef, const int *d_slots, int n_terms, double *d_out,
                    cudaStream_t s) {
  lincomb_kernel<<<grid_for(count), 256, 0, s>>>(d_vecs, count, d_coef, d_slots, n_terms, d_out);
}

}  // namespace mqcb200

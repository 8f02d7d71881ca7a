// j_kernels.cu -- the Coulomb half of build_fock_df on the packed tensor.
//
//   pass 1   gamma_Q = sum_p Bp[Q][p] * w[p]        (reference rhf.f90:1620-1622)
//   pass 2   Jp[p]   = sum_Q gamma_Q * Bp[Q][p]     (reference rhf.f90:1624-1627)
//
// Both passes stream the whole packed tensor once: algorithmic traffic is
// 8*L*Q bytes per pass (L = packed row length), i.e. 16*npair*Q per build up to the
// tile padding.  They are HBM-bound; the design goals are 128-bit coalesced
// streaming loads that bypass L1, enough loads in flight per SM to cover the HBM
// latency, and fixed-order reductions (no floating-point atomics) so that repeated
// builds are bit-identical.
#include "common.cuh"
#include "kernels.cuh"

namespace mqcb200 {

constexpr int J_THREADS = 256;
constexpr int J_ROWS = 8;               // auxiliary rows per CTA in pass 1
constexpr int J_STEP = J_THREADS * 2;   // doubles covered by one CTA-wide 128-bit load

__global__ void __launch_bounds__(J_THREADS) j_gamma_kernel(const double *__restrict__ packed, long long L,
                                                            int q_count, const double *__restrict__ w,
                                                            long long seg_len, int qblk0,
                                                            double *__restrict__ gamma_partial,
                                                            const int *__restrict__ skip_flag) {
  if (skip_flag != nullptr && *skip_flag != 0) return;   // gamma comes from the half-transform
  const int seg = blockIdx.x;
  const int q0 = (qblk0 + blockIdx.y) * J_ROWS;
  const long long start = (long long)seg * seg_len;
  long long end = start + seg_len;
  if (end > L) end = L;

  const double *row[J_ROWS];
#pragma unroll
  for (int r = 0; r < J_ROWS; ++r) {
    int q = q0 + r;
    if (q > q_count - 1) q = q_count - 1;  // tail rows re-read the last slab; their sums are dropped
    row[r] = packed + (size_t)q * L;
  }
  double acc[J_ROWS];
#pragma unroll
  for (int r = 0; r < J_ROWS; ++r) acc[r] = 0.0;
  const uint64_t pol = policy_evict_first();

  long long i = start + 2 * threadIdx.x;
  // two steps per trip: 16 independent 128-bit tensor loads in flight per thread
  for (; i + J_STEP < end; i += 2 * J_STEP) {
    const double2 w0 = *reinterpret_cast<const double2 *>(w + i);
    const double2 w1 = *reinterpret_cast<const double2 *>(w + i + J_STEP);
    double2 v0[J_ROWS], v1[J_ROWS];
#pragma unroll
    for (int r = 0; r < J_ROWS; ++r) v0[r] = ld_stream_f64x2_hint(row[r] + i, pol);
#pragma unroll
    for (int r = 0; r < J_ROWS; ++r) v1[r] = ld_stream_f64x2_hint(row[r] + i + J_STEP, pol);
#pragma unroll
    for (int r = 0; r < J_ROWS; ++r) {
      acc[r] = fma(v0[r].x, w0.x, acc[r]);
      acc[r] = fma(v0[r].y, w0.y, acc[r]);
      acc[r] = fma(v1[r].x, w1.x, acc[r]);
      acc[r] = fma(v1[r].y, w1.y, acc[r]);
    }
  }
  if (i < end) {
    const double2 w0 = *reinterpret_cast<const double2 *>(w + i);
    double2 v0[J_ROWS];
#pragma unroll
    for (int r = 0; r < J_ROWS; ++r) v0[r] = ld_stream_f64x2_hint(row[r] + i, pol);
#pragma unroll
    for (int r = 0; r < J_ROWS; ++r) {
      acc[r] = fma(v0[r].x, w0.x, acc[r]);
      acc[r] = fma(v0[r].y, w0.y, acc[r]);
    }
  }

  __shared__ double red[J_THREADS / 32][J_ROWS];
#pragma unroll
  for (int r = 0; r < J_ROWS; ++r) {
    double s = acc[r];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][r] = s;
  }
  __syncthreads();
  if (threadIdx.x < J_ROWS && q0 + threadIdx.x < q_count) {
    double s = 0.0;
#pragma unroll
    for (int wp = 0; wp < J_THREADS / 32; ++wp) s += red[wp][threadIdx.x];
    gamma_partial[(size_t)seg * q_count + q0 + threadIdx.x] = s;
  }
}

__global__ void __launch_bounds__(256) j_gamma_reduce_kernel(const double *__restrict__ gamma_partial,
                                                             int n_seg, int q_count,
                                                             double *__restrict__ gamma,
                                                             const int *__restrict__ skip_flag) {
  if (skip_flag != nullptr && *skip_flag != 0) return;
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= q_count) return;
  double s = 0.0;
  for (int seg = 0; seg < n_seg; ++seg) s += gamma_partial[(size_t)seg * q_count + q];
  gamma[q] = s;
}

__global__ void __launch_bounds__(256) gamma_from_x_kernel(const double *__restrict__ part_a, int stride_a,
                                                           const double *__restrict__ part_b, int stride_b,
                                                           double factor, int q_count,
                                                           const int *__restrict__ flag,
                                                           double *__restrict__ gamma) {
  if (*flag == 0) return;
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= q_count) return;
  double s = 0.0;
  for (int k = 0; k < stride_a; ++k) s += part_a[(size_t)q * stride_a + k];
  if (part_b != nullptr)
    for (int k = 0; k < stride_b; ++k) s += part_b[(size_t)q * stride_b + k];
  gamma[q] = factor * s;
}

constexpr int J2_QCHUNK = 1024;

__global__ void __launch_bounds__(J_THREADS) j_accumulate_kernel(const double *__restrict__ packed, long long L,
                                                                 int q_count, const double *__restrict__ gamma,
                                                                 int n_slices, double *__restrict__ jpart) {
  __shared__ double g_s[J2_QCHUNK];
  const int slice = blockIdx.y;
  const int qa = (int)((long long)slice * q_count / n_slices);
  const int qb = (int)((long long)(slice + 1) * q_count / n_slices);
  const long long col = (long long)blockIdx.x * J_STEP + 2 * threadIdx.x;
  const bool live = col < L;
  const double *base = packed + (live ? col : 0);
  double ax = 0.0, ay = 0.0;
  const uint64_t pol = policy_evict_first();

  for (int qc = qa; qc < qb; qc += J2_QCHUNK) {
    const int cnt = qb - qc < J2_QCHUNK ? qb - qc : J2_QCHUNK;
    __syncthreads();
    for (int t = threadIdx.x; t < cnt; t += J_THREADS) g_s[t] = gamma[qc + t];
    __syncthreads();
    if (live) {
      int t = 0;
      for (; t + 8 <= cnt; t += 8) {
        double2 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = ld_stream_f64x2_hint(base + (size_t)(qc + t + u) * L, pol);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const double g = g_s[t + u];
          ax = fma(g, v[u].x, ax);
          ay = fma(g, v[u].y, ay);
        }
      }
      for (; t < cnt; ++t) {
        const double2 v = ld_stream_f64x2_hint(base + (size_t)(qc + t) * L, pol);
        const double g = g_s[t];
        ax = fma(g, v.x, ax);
        ay = fma(g, v.y, ay);
      }
    }
  }
  if (live) *reinterpret_cast<double2 *>(jpart + (size_t)slice * L + col) = make_double2(ax, ay);
}

// ------------------------------------------------------------------------------------------------
// Pass 2 as a CO-RESIDENT streaming kernel: J accumulated beside the exchange accumulation.
//
// k_accumulate_kernel<64> holds three CTAs per SM and leaves exactly 4096 registers and ~27 KiB of
// shared memory unused.  This kernel is shaped to fit in that gap -- 64 threads x 64 registers, a
// 24 KiB ring -- so that one of its CTAs sits on every SM next to the three DMMA-bound ones and pulls
// the packed tensor through the otherwise idle HBM interface.  With so few registers the loads
// cannot be kept in flight by the threads; they are kept in flight by the TMA engine instead: thread 0
// issues 2 KiB bulk copies (one row segment of 256 packed columns each, four rows per stage, three
// stages, L2 evict-first) onto mbarriers, and the two warps fold each landed row into their four accumulators per
// thread (Jp[col] += gamma_Q * Bp[Q][col]).  Persistent: the ring runs on across work items
// (column block x auxiliary slice), partials go to jpart[slice][L] like the stand-alone kernel's.
constexpr int JT_THREADS = 64;
constexpr int JT_COLS = 256;                 // doubles per row segment: one 2 KiB bulk copy
constexpr int JT_ROWS = 4;                   // rows per stage
constexpr int JT_STAGES = 3;                 // (12 single-row stages measured worse: per-row barrier traffic, see r02_notes)

struct JtCursor {                            // position in the flattened (work item, row block) sequence
  int item, r0, qa, qb;
  long long col0;
};

__device__ __forceinline__ void jt_open(JtCursor &c, int item, int n_colblk, int n_slices, int q_count) {
  c.item = item;
  const int slice = item / n_colblk;
  c.col0 = (long long)(item - slice * n_colblk) * JT_COLS;
  c.qa = (int)((long long)slice * q_count / n_slices);
  c.qb = (int)((long long)(slice + 1) * q_count / n_slices);
  c.r0 = c.qa;
}

__global__ void __launch_bounds__(JT_THREADS, 16)
j_accumulate_tma_kernel(const double *__restrict__ packed, long long L, int q_count, const double *__restrict__ gamma,
                        int n_slices, int n_colblk, double *__restrict__ jpart) {
  __shared__ __align__(128) double tile[JT_STAGES][JT_ROWS][JT_COLS];
  __shared__ uint64_t full_bar[JT_STAGES], empty_bar[JT_STAGES];
  const int tid = threadIdx.x, lane = tid & 31;
  const int n_items = n_colblk * n_slices;
  if (tid == 0) {
    for (int s = 0; s < JT_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], JT_THREADS / 32); }
    fence_mbar_init();
  }
  __syncthreads();
  if ((int)blockIdx.x >= n_items) return;

  const uint64_t pol = policy_evict_first();
  JtCursor prod, cons;
  jt_open(prod, blockIdx.x, n_colblk, n_slices, q_count);
  cons = prod;
  bool prod_live = true;
  uint32_t it_prod = 0, it_cons = 0;
  auto produce = [&]() {                      // thread 0 only: one stage = up to JT_ROWS row segments
    while (prod_live && prod.r0 >= prod.qb) { // (an empty slice has nothing to load)
      const int next = prod.item + gridDim.x;
      if (next >= n_items) { prod_live = false; break; }
      jt_open(prod, next, n_colblk, n_slices, q_count);
    }
    if (!prod_live) return;
    const int st = it_prod % JT_STAGES;
    mbar_wait(&empty_bar[st], ((it_prod / JT_STAGES) & 1) ^ 1);
    const int rows = prod.qb - prod.r0 < JT_ROWS ? prod.qb - prod.r0 : JT_ROWS;
    mbar_arrive_expect_tx(&full_bar[st], (uint32_t)rows * JT_COLS * sizeof(double));
    for (int r = 0; r < rows; ++r)
      tma_load_1d_hint(&tile[st][r][0], packed + (size_t)(prod.r0 + r) * L + prod.col0, JT_COLS * sizeof(double), &full_bar[st], pol);
    prod.r0 += rows;
    ++it_prod;
  };
  if (tid == 0)
    for (int s = 0; s < JT_STAGES - 1; ++s) produce();

  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  while (true) {
    // skip empty slices on the consumer side too (their partial is zero)
    while (cons.r0 >= cons.qb) {
      *reinterpret_cast<double2 *>(jpart + (size_t)(cons.item / n_colblk) * L + cons.col0 + 4 * tid) = make_double2(acc[0], acc[1]);
      *reinterpret_cast<double2 *>(jpart + (size_t)(cons.item / n_colblk) * L + cons.col0 + 4 * tid + 2) = make_double2(acc[2], acc[3]);
      acc[0] = acc[1] = acc[2] = acc[3] = 0.0;
      const int next = cons.item + gridDim.x;
      if (next >= n_items) return;
      jt_open(cons, next, n_colblk, n_slices, q_count);
    }
    if (tid == 0) produce();                  // keep JT_STAGES - 1 stages in flight ahead of this one
    const int st = it_cons % JT_STAGES;
    const int rows = cons.qb - cons.r0 < JT_ROWS ? cons.qb - cons.r0 : JT_ROWS;
    double g[JT_ROWS];
#pragma unroll
    for (int r = 0; r < JT_ROWS; ++r) g[r] = r < rows ? gamma[cons.r0 + r] : 0.0;
    mbar_wait(&full_bar[st], (it_cons / JT_STAGES) & 1);
#pragma unroll
    for (int r = 0; r < JT_ROWS; ++r) {
      if (r < rows) {
        const double2 v0 = *reinterpret_cast<const double2 *>(&tile[st][r][4 * tid]);
        const double2 v1 = *reinterpret_cast<const double2 *>(&tile[st][r][4 * tid + 2]);
        acc[0] = fma(g[r], v0.x, acc[0]);
        acc[1] = fma(g[r], v0.y, acc[1]);
        acc[2] = fma(g[r], v1.x, acc[2]);
        acc[3] = fma(g[r], v1.y, acc[3]);
      }
    }
    release_stage(&empty_bar[st], lane);
    cons.r0 += rows;
    ++it_cons;
  }
}

void launch_j_accumulate_tma(const double *d_packed, long long L, int q_count, const double *d_gamma, int n_slices,
                             int sm_count, double *d_jpart, cudaStream_t s) {
  const int n_colblk = (int)(L / JT_COLS);
  const long long n_items = (long long)n_colblk * n_slices;
  const unsigned grid = (unsigned)(n_items < sm_count ? n_items : sm_count);
  j_accumulate_tma_kernel<<<grid, JT_THREADS, 0, s>>>(d_packed, L, q_count, d_gamma, n_slices, n_colblk, d_jpart);
}

JPlan plan_j(int n, int q_count) {
  JPlan p;
  const long long L = packed_row_len(n);
  const int target = 148 * 8;
  const int nqblk = (q_count + J_ROWS - 1) / J_ROWS;
  long long max_seg = L / 4096;
  if (max_seg < 1) max_seg = 1;
  long long ns = (target + nqblk - 1) / (nqblk > 0 ? nqblk : 1);
  if (ns > max_seg) ns = max_seg;
  if (ns < 1) ns = 1;
  p.n_seg = (int)ns;
  const long long ncolblk = (L + J_STEP - 1) / J_STEP;
  long long sl = (target + ncolblk - 1) / ncolblk;
  long long max_sl = q_count / 64;
  if (max_sl < 1) max_sl = 1;
  if (sl > max_sl) sl = max_sl;
  if (sl < 1) sl = 1;
  p.n_slices = (int)sl;
  p.gamma_partial_elems = (size_t)p.n_seg * (size_t)(q_count > 0 ? q_count : 1);
  p.j_partial_elems = (size_t)p.n_slices * (size_t)L;
  return p;
}

void launch_gamma_from_x(const double *d_part_a, int stride_a, const double *d_part_b, int stride_b, double factor,
                         int q_count, const int *d_flag, double *d_gamma, cudaStream_t s) {
  gamma_from_x_kernel<<<(q_count + 255) / 256, 256, 0, s>>>(d_part_a, stride_a, d_part_b, stride_b, factor, q_count,
                                                            d_flag, d_gamma);
}

void launch_j_gamma(const double *d_packed, long long L, int q_count, const double *d_w, const JPlan &plan,
                    double *d_gamma_partial, double *d_gamma, const int *d_skip_flag, cudaStream_t s) {
  long long seg_len = (L + plan.n_seg - 1) / plan.n_seg;
  seg_len = (seg_len + J_STEP - 1) / J_STEP * J_STEP;
  const int nqblk = (q_count + J_ROWS - 1) / J_ROWS;
  for (int b0 = 0; b0 < nqblk; b0 += 65535) {
    const int nb = nqblk - b0 < 65535 ? nqblk - b0 : 65535;
    dim3 grid((unsigned)plan.n_seg, (unsigned)nb);
    j_gamma_kernel<<<grid, J_THREADS, 0, s>>>(d_packed, L, q_count, d_w, seg_len, b0, d_gamma_partial, d_skip_flag);
  }
  j_gamma_reduce_kernel<<<(q_count + 255) / 256, 256, 0, s>>>(d_gamma_partial, plan.n_seg, q_count, d_gamma,
                                                              d_skip_flag);
}

void launch_j_accumulate(const double *d_packed, long long L, int q_count, const double *d_gamma,
                         const JPlan &plan, double *d_jpart, cudaStream_t s) {
  const long long ncolblk = (L + J_STEP - 1) / J_STEP;
  dim3 grid((unsigned)ncolblk, (unsigned)plan.n_slices);
  j_accumulate_kernel<<<grid, J_THREADS, 0, s>>>(d_packed, L, q_count, d_gamma, plan.n_slices, d_jpart);
}

}  // namespace mqcb200

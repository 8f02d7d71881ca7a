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
#include "common.cuh"
#include "kernels.cuh"

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
__global__ void __launch_bounds__(256) pack_half_kernel(const double *__restrict__ half, int naux, int nmb, int nkc,
                                                        double *__restrict__ af) {
  const size_t total = (size_t)nkc * nmb * 128;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const int t = e & 3, g = (e >> 2) & 7, ks = (e >> 5) & 3;
    const size_t blk = e >> 7;
    const int mb = (int)(blk % nmb), kc = (int)(blk / nmb);
    const int q = mb * 8 + g, p = kc * 16 + ks * 4 + t;
    af[e] = (q < naux && p < naux) ? half[(size_t)p + (size_t)naux * q] : 0.0;   // contiguous in p
  }
}

__global__ void __launch_bounds__(K_THREADS, 1)
whiten_gemm_kernel(const double *__restrict__ af, int nmb, int nkc, const double *__restrict__ tp, long long L,
                   int naux, double *__restrict__ bp) {
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
        tma_load_1d(a_s + Cfg::kAElems + lane * Cfg::kBStride, tp + (size_t)(16 * kc + lane) * L + p0,
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

#pragma unroll
  for (int m = 0; m < 4; ++m) {
    const int q = 128 * mt + 32 * wm + 8 * m + g;
    if (q >= naux) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const long long c = p0 + 64 * wn + 8 * j + 2 * t;
      *reinterpret_cast<double2 *>(bp + (size_t)q * L + c) = make_double2(acc[m][j][0], acc[m][j][1]);
    }
  }
}

void configure_whiten_kernels() {
  cudaFuncSetAttribute(whiten_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WhitenCfg::kSmemBytes);
}

size_t whiten_half_elems(int naux) {
  const int nmb = (naux + 7) / 8, nkc = (naux + 15) / 16;
  return (size_t)nkc * nmb * 128;
}

void launch_whiten(const double *d_half, int naux, const double *d_tp, int n, double *d_af, double *d_bp,
                   cudaStream_t s) {
  const int nmb = (naux + 7) / 8, nkc = (naux + 15) / 16;
  const long long L = packed_row_len(n);
  const size_t total = (size_t)nkc * nmb * 128;
  unsigned blocks = (unsigned)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  pack_half_kernel<<<blocks, 256, 0, s>>>(d_half, naux, nmb, nkc, d_af);
  const unsigned mtiles = (unsigned)((naux + 127) / 128);
  const long long ntiles = L / 128;
  for (long long y0 = 0; y0 < ntiles; y0 += 65535) {
    const unsigned ny = (unsigned)(ntiles - y0 < 65535 ? ntiles - y0 : 65535);
    dim3 grid(mtiles, ny);
    whiten_gemm_kernel<<<grid, K_THREADS, WhitenCfg::kSmemBytes, s>>>(d_af, nmb, nkc, d_tp + y0 * 128, L, naux,
                                                                      d_bp + y0 * 128);
  }
}

}  // namespace mqcb200

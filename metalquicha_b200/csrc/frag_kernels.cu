// frag_kernels.cu -- one-pass J and K for fragment-sized problems (n <= 80, n_occ <= 64).
//
// The (H2O)n fragments of an MBE run are tiny: a def2-SVP trimer is n = 72, n_occ = 15,
// naux ~ 340, i.e. 79 MFLOP and 7 MB of packed tensor per build.  The general path's five
// kernels are then pure launch latency.  Here ONE kernel does the whole of
//     gamma_Q = B_Q . w,   J += gamma_Q B_Q,   X_Q = B_Q C,   K += X_Q X_Q^T
// with each packed slab B_Q read once: a CTA owns the auxiliary functions q = c, c+G, ...
// and for each of them
//   1. the slab (<= 30 KiB) arrives in shared memory by one bulk-TMA copy (double-buffered,
//      the next slab is in flight while this one is used);
//   2. gamma_Q is a block-wide dot with the density weights (fixed-order reduction);
//   3. X_Q = B_Q C on DMMA, fragments straight out of the packed tiles (direct or
//      transposed, as in the big half-transform), X_Q kept in shared memory in fragment order;
//   4. K += X_Q X_Q^T on DMMA (lower-triangular 8x8 blocks, accumulators stay in registers
//      across the CTA's auxiliary functions) and J += gamma_Q B_Q (registers).
// Per-CTA partials of J (packed) and K (64x64 tiles) go to the same buffers the general path
// uses and are summed in fixed order by finalize_jk_kernel: deterministic, no atomics.
#include "common.cuh"
#include "kernels.cuh"

#include <cstdlib>

namespace mqcb200 {

constexpr int FR_THREADS = 256;
constexpr int FR_WARPS = 8;
constexpr int FR_MAX_NT = 5;      // n <= 80
constexpr int FR_MAX_NIB = 8;     // padded n_occ <= 64
constexpr int FR_MAX_ABLK = (2 * FR_MAX_NT * FR_MAX_NIB + FR_WARPS - 1) / FR_WARPS;              // X blocks per warp
constexpr int FR_MAX_JREG = (FR_MAX_NT * (FR_MAX_NT + 1) / 2 * TILE_ELEMS) / FR_THREADS;   // 15 doubles

template <int NT>
__global__ void __launch_bounds__(FR_THREADS, 1)
fragment_jk_kernel(const double *__restrict__ packed, int L, int q_count,
                   const double *__restrict__ w, const double *__restrict__ ctf, int nib, int want_j, int want_k,
                   double *__restrict__ jpart, double *__restrict__ kpart) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double *b_s = reinterpret_cast<double *>(smem_raw);            // [2][L]
  double *w_s = b_s + 2 * (size_t)L;                             // [L]
  double *c_s = w_s + L;                                         // [nt][nib][4][32]
  constexpr int nt = NT;
  constexpr int nmb = 2 * NT;
  constexpr int KB = (nmb * (nmb + 1) / 2 + FR_WARPS - 1) / FR_WARPS;   // K blocks per warp (last one may be absent)
  const int nkc = (nib + 1) / 2;                                 // 16-wide chunks of the padded occupied range
  double *x_s = c_s + (size_t)nt * nib * 128;                    // [nkc][nmb][4][32]
  double *red_s = x_s + (size_t)nkc * nmb * 128;                 // [FR_WARPS]
  uint64_t *bar = reinterpret_cast<uint64_t *>(red_s + FR_WARPS);   // [2]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int G = gridDim.x;
  if (blockIdx.y != 0) {                     // fragment of a batch: its own slabs, weights, coefficients and partials
    const int np_b = (16 * NT + 63) / 64;
    packed += (size_t)blockIdx.y * (size_t)q_count * L;
    w += (size_t)blockIdx.y * L;
    ctf += (size_t)blockIdx.y * NT * nib * 128;
    jpart += (size_t)blockIdx.y * (size_t)G * L;
    kpart += (size_t)blockIdx.y * (size_t)G * (np_b * (np_b + 1) / 2) * 4096;
  }
  const int my_count = (q_count - (int)blockIdx.x + G - 1) / G;  // auxiliary functions of this CTA

  if (tid == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    fence_mbar_init();
  }
  for (int e = tid; e < L; e += FR_THREADS) w_s[e] = want_j ? w[e] : 0.0;
  for (int e = tid; e < nt * nib * 128; e += FR_THREADS) c_s[e] = want_k ? ctf[e] : 0.0;
  if ((nib & 1) && want_k)                                       // odd nib: the upper half of the last chunk of X stays zero
    for (int e = tid; e < nkc * nmb * 128; e += FR_THREADS) x_s[e] = 0.0;
  __syncthreads();
  if (tid == 0 && my_count > 0) {
    mbar_arrive_expect_tx(&bar[0], (uint32_t)L * sizeof(double));
    tma_load_1d(b_s, packed + (size_t)blockIdx.x * L, (uint32_t)L * sizeof(double), &bar[0]);
  }

  // accumulators that live across the CTA's auxiliary functions
  double jacc[FR_MAX_JREG];
#pragma unroll
  for (int i = 0; i < FR_MAX_JREG; ++i) jacc[i] = 0.0;
  // this warp's lower-triangular K blocks: ids warp, warp + 8, ... of the row-major list (mb, nb <= mb)
  double kacc[KB][2];
  int kb_mb[KB], kb_nb[KB];
  constexpr int n_kblocks = nmb * (nmb + 1) / 2;
  const int my_kblocks = (n_kblocks - warp + FR_WARPS - 1) / FR_WARPS;
#pragma unroll
  for (int k = 0; k < KB; ++k) {
    kacc[k][0] = 0.0;
    kacc[k][1] = 0.0;
    int id = warp + FR_WARPS * k, mb = 0;
    if (id >= n_kblocks) id = 0;
    while ((mb + 1) * (mb + 2) / 2 <= id) ++mb;
    kb_mb[k] = mb;
    kb_nb[k] = id - mb * (mb + 1) / 2;
  }

  const int lane_transp = ((g >> 2) << 5) + (t << 2) + (g & 3);

  for (int it = 0; it < my_count; ++it) {
    const int buf = it & 1;
    double *bq = b_s + (size_t)buf * L;
    // prefetch the next slab into the other buffer (its previous reader finished at the
    // __syncthreads that closed the last iteration; order that read against the async write)
    if (tid == 0 && it + 1 < my_count) {
      mbar_arrive_expect_tx(&bar[buf ^ 1], (uint32_t)L * sizeof(double));
      tma_load_1d(b_s + (size_t)(buf ^ 1) * L, packed + ((size_t)blockIdx.x + (size_t)(it + 1) * G) * L,
                  (uint32_t)L * sizeof(double), &bar[buf ^ 1]);
    }
    mbar_wait(&bar[buf], (it >> 1) & 1);

    // ---- gamma_Q: block-wide dot in fixed order
    double gamma = 0.0;
    if (want_j) {
      double s = 0.0;
      for (int e = tid; e < L; e += FR_THREADS) s = fma(bq[e], w_s[e], s);
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) red_s[warp] = s;
    }

    // ---- X_Q = B_Q C : the nmb x nib output blocks (8x8 each) are dealt round-robin to the
    // warps, so the work is balanced whatever the fragment's shape and no DMMA is issued
    // for a block that does not exist (a predicated-off DMMA still occupies the pipe)
    if (want_k) {
      const int n_blocks = nmb * nib;
      for (int k0 = 0; k0 < FR_MAX_ABLK; k0 += 2) {          // two independent accumulation chains at a time
        const int id0 = warp + FR_WARPS * k0, id1 = id0 + FR_WARPS;
        if (id0 >= n_blocks) break;
        const bool two = id1 < n_blocks;
        const int mb0 = id0 / nib, j0 = id0 - mb0 * nib;
        const int mb1 = two ? id1 / nib : mb0, j1 = two ? id1 - mb1 * nib : j0;
        const int tr0 = mb0 >> 1, rh0 = mb0 & 1, tr1 = mb1 >> 1, rh1 = mb1 & 1;
        // two blocks x two halves of the contraction range = four independent DMMA chains
        // (a single chain is bound by the DMMA latency, not by the pipe)
        double a0 = 0.0, a1 = 0.0, c0 = 0.0, c1 = 0.0, a0h = 0.0, a1h = 0.0, c0h = 0.0, c1h = 0.0;
        constexpr int half = (nt + 1) / 2;
        for (int kc = 0; kc < half; ++kc) {
          const int kh = kc + half;                            // second half (may run past the end: guarded)
          const bool hv = kh < nt;
          const bool tp0 = tr0 < kc, tp1 = tr1 < kc, tp0h = tr0 < kh, tp1h = tr1 < kh;
          const double *tile0 = bq + (size_t)tile_index(tr0 > kc ? tr0 : kc, tr0 > kc ? kc : tr0, nt) * TILE_ELEMS;
          const double *tile1 = bq + (size_t)tile_index(tr1 > kc ? tr1 : kc, tr1 > kc ? kc : tr1, nt) * TILE_ELEMS;
          const int khc = hv ? kh : kc;
          const double *tile0h = bq + (size_t)tile_index(tr0 > khc ? tr0 : khc, tr0 > khc ? khc : tr0, nt) * TILE_ELEMS;
          const double *tile1h = bq + (size_t)tile_index(tr1 > khc ? tr1 : khc, tr1 > khc ? khc : tr1, nt) * TILE_ELEMS;
          const double *cb = c_s + (size_t)kc * nib * 128;
          const double *cbh = c_s + (size_t)khc * nib * 128;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const int od0 = (rh0 * 4 + ks) * 32 + lane, ot0 = (((ks >> 1) * 4 + 2 * rh0) * 32) + (ks & 1) * 16 + lane_transp;
            const int od1 = (rh1 * 4 + ks) * 32 + lane, ot1 = (((ks >> 1) * 4 + 2 * rh1) * 32) + (ks & 1) * 16 + lane_transp;
            const double fb0 = cb[(j0 * 4 + ks) * 32 + lane], fb1 = cb[(j1 * 4 + ks) * 32 + lane];
            const double fb0h = cbh[(j0 * 4 + ks) * 32 + lane], fb1h = cbh[(j1 * 4 + ks) * 32 + lane];
            dmma884(a0, a1, tile0[tp0 ? ot0 : od0], fb0);
            if (two) dmma884(c0, c1, tile1[tp1 ? ot1 : od1], fb1);
            if (hv) {
              dmma884(a0h, a1h, tile0h[tp0h ? ot0 : od0], fb0h);
              if (two) dmma884(c0h, c1h, tile1h[tp1h ? ot1 : od1], fb1h);
            }
          }
        }
        a0 += a0h; a1 += a1h; c0 += c0h; c1 += c1h;
        // X in fragment order: x_s[((kc2*nmb + mb)*4 + ks2)*32 + g*4 + t2], i = 8j + 2t (+1)
        {
          const int i0 = 8 * j0, kc2 = i0 >> 4, ks2 = ((i0 & 15) >> 2) + (t >> 1);
          *reinterpret_cast<double2 *>(x_s + ((size_t)(kc2 * nmb + mb0) * 4 + ks2) * 32 + g * 4 + 2 * (t & 1)) =
              make_double2(a0, a1);
        }
        if (two) {
          const int i0 = 8 * j1, kc2 = i0 >> 4, ks2 = ((i0 & 15) >> 2) + (t >> 1);
          *reinterpret_cast<double2 *>(x_s + ((size_t)(kc2 * nmb + mb1) * 4 + ks2) * 32 + g * 4 + 2 * (t & 1)) =
              make_double2(c0, c1);
        }
      }
    }
    __syncthreads();                                   // X and the gamma partials are complete

    if (want_j) {
#pragma unroll
      for (int wp = 0; wp < FR_WARPS; ++wp) gamma += red_s[wp];
#pragma unroll
      for (int i = 0; i < FR_MAX_JREG; ++i) {
        const int e = tid + i * FR_THREADS;
        if (e < L) jacc[i] = fma(gamma, bq[e], jacc[i]);
      }
    }

    // ---- K += X_Q X_Q^T : the lower-triangular 8x8 blocks (mb, nb <= mb), dealt round-robin;
    // the blocks are the innermost loop so consecutive DMMAs are independent
    if (want_k) {
      for (int kc2 = 0; kc2 < nkc; ++kc2) {
        const double *xk = x_s + (size_t)kc2 * nmb * 128 + lane;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
          for (int k = 0; k < KB; ++k)
            if (k < my_kblocks)                        // warp-uniform; only the last block can be absent
              dmma884(kacc[k][0], kacc[k][1], xk[kb_mb[k] * 128 + ks * 32], xk[kb_nb[k] * 128 + ks * 32]);
        }
      }
    }
    fence_proxy_async();                               // generic reads of this slab before its bulk-TMA refill
    __syncthreads();                                   // everyone is done with this slab and with X
  }

  // ---- per-CTA partials, in the layouts finalize_jk_kernel sums
  if (want_j) {
#pragma unroll
    for (int i = 0; i < FR_MAX_JREG; ++i) {
      const int e = tid + i * FR_THREADS;
      if (e < L) jpart[(size_t)blockIdx.x * L + e] = jacc[i];
    }
  }
  if (want_k) {
    const int np = (16 * nt + 63) / 64;                // 64-row panels
    const int n_ktiles = np * (np + 1) / 2;
    double *out = kpart + (size_t)blockIdx.x * n_ktiles * 4096;
#pragma unroll
    for (int k = 0; k < KB; ++k) {
      if (k >= my_kblocks) continue;
      const int row = 8 * kb_mb[k] + g, col = 8 * kb_nb[k] + 2 * t;
      const int mp = row >> 6, npn = col >> 6;
      double *tp = out + (size_t)(mp * (mp + 1) / 2 + npn) * 4096 + ((row & 63) << 6) + (col & 63);
      *reinterpret_cast<double2 *>(tp) = make_double2(kacc[k][0], kacc[k][1]);
    }
  }
}

void configure_fragment_kernels() {
  cudaFuncSetAttribute(fragment_jk_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(fragment_jk_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(fragment_jk_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(fragment_jk_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(fragment_jk_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
}

bool fragment_path_applies(int n, int n_occ) {
  const int nt = num_tiles(n);
  const int nib = ((n_occ > 0 ? n_occ : 1) + 7) / 8;
  return nt <= FR_MAX_NT && nib <= FR_MAX_NIB;
}

FragPlan plan_fragment(int n, int n_occ, int q_count, int sm_count) {
  FragPlan p;
  p.nt = num_tiles(n);
  p.nib = ((n_occ > 0 ? n_occ : 1) + 7) / 8;
  p.L = (int)packed_row_len(n);
  const int nkc = (p.nib + 1) / 2;
  // a few auxiliary functions per CTA amortise its set-up; the fixed-order sum over the
  // CTAs' partials afterwards costs about as much per CTA, so do not spread too thin
  int g = (q_count + 3) / 4;
  if (g > sm_count) g = sm_count;
  if (g > 64) g = 64;
  if (g < 1) g = 1;
  if (const char *env = getenv("MQCB200_FRAG_GRID")) {   // development override
    const int v = atoi(env);
    if (v >= 1) g = v < q_count ? v : q_count;
  }
  p.grid = g;
  const int np = (16 * p.nt + 63) / 64;
  p.n_ktiles = np * (np + 1) / 2;
  p.smem_bytes = sizeof(double) * ((size_t)3 * p.L + (size_t)p.nt * p.nib * 128 + (size_t)nkc * 2 * p.nt * 128 + FR_WARPS) +
                 2 * sizeof(uint64_t);
  p.jpart_elems = (size_t)p.grid * p.L;
  p.kpart_elems = (size_t)p.grid * p.n_ktiles * 4096;
  return p;
}

void launch_fragment_jk(const double *d_packed, int q_count, const double *d_w, const double *d_ctf,
                        const FragPlan &p, bool want_j, bool want_k, double *d_jpart, double *d_kpart,
                        cudaStream_t s, int batch) {
  // only the lower-triangular 8x8 blocks of the K partial tiles are written -- exactly the
  // elements finalize_jk_kernel reads (row >= col), so no clearing is needed
#define MQCB200_LAUNCH_FRAG(NT_)                                                                              \
  fragment_jk_kernel<NT_><<<dim3((unsigned)p.grid, (unsigned)batch), FR_THREADS, p.smem_bytes, s>>>(d_packed, p.L, q_count, d_w, d_ctf, p.nib, \
                                                                   want_j ? 1 : 0, want_k ? 1 : 0, d_jpart, d_kpart)
  switch (p.nt) {
    case 1: MQCB200_LAUNCH_FRAG(1); break;
    case 2: MQCB200_LAUNCH_FRAG(2); break;
    case 3: MQCB200_LAUNCH_FRAG(3); break;
    case 4: MQCB200_LAUNCH_FRAG(4); break;
    default: MQCB200_LAUNCH_FRAG(5); break;
  }
#undef MQCB200_LAUNCH_FRAG
}

}  // namespace mqcb200

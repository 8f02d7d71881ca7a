// k_kernels.cu -- the exchange half of build_fock_df on the FP64 tensor pipe.
//
//   half-transform   X_Q = B_Q . C_occ                 (reference rhf.f90:1636)
//   accumulation     K  += X_Q . X_Q^T                 (reference rhf.f90:1637)
//
// tcgen05/TMEM have no FP64 kind, so the sm_100a tensor path for this contraction
// is DMMA.8x8x4 (mma.sync.m8n8k4.f64); cuBLAS's own DGEMM on B200 is the same
// instruction (profiles/r01_ubench.md).  What is Blackwell-specific here is the
// feeding: operands live in HBM already in DMMA fragment order, so every pipeline
// stage is a handful of 1-D bulk-TMA copies (cp.async.bulk -> SASS UBLKCP) landing
// on an mbarrier, and every fragment a warp needs is 256 contiguous, bank-conflict-
// free bytes of shared memory.  One producer warp drives the TMA ring; eight
// consumer warps (4 along M x 2 along N) issue DMMA.
//
// Half-transform: the M dimension is the flattened (Q, tile-row) index, so a CTA's
// eight 16-row slots may straddle two auxiliary functions and there is no padding
// waste along M.  Slot s at k-chunk kc needs the 16x16 tile B_Q[tr_s, kc]; it is the
// stored tile (max, min) and is read transposed out of shared memory when
// tr_s < kc (the packed tensor holds each off-diagonal tile once).
//
// Accumulation: SYRK over 64x64 lower-triangular tiles of K (three CTAs per SM) with the
// (Q, i) contraction index split across CTAs; partial tiles are written (or added to, on
// later Q-chunks) in a fixed order -- no floating-point atomics.  The same kernel runs the
// rank-2 (SYR2K) form  sum_P (B_P X)(B_P C)^T + (B_P C)(B_P X)^T  of the response operator
// (reference mqc_libcint_cphf.F90:554-563): the half-transform is fed the stacked [X | C], and a
// pipeline stage pairs the X-chunk of one panel with the C-chunk of the other, then the reverse.
#include "common.cuh"
#include "kernels.cuh"

#include <cstdio>
#include <cstdlib>
#include <type_traits>

namespace mqcb200 {

// ------------------------------------------------------------------------------------
// Half-transform
// ------------------------------------------------------------------------------------
template <int NB>
struct HalfCfg {
  static constexpr int kStages = 4;
  static constexpr int kAElems = K_SLOTS * TILE_ELEMS;   // 2048 doubles = 16 KiB
  static constexpr int kBElems = NB * 256;               // 2*NB i-blocks * 4 k-subs * 32
  static constexpr int kStageElems = kAElems + kBElems;
  static constexpr size_t kSmemBytes = (size_t)kStages * kStageElems * sizeof(double) + 2 * kStages * sizeof(uint64_t);
};

// Persistent: one CTA per SM walks the (row-group, N-tile) list round-robin.  The TMA
// ring runs continuously across tile boundaries -- while the consumer warps store one
// tile's X the producer is already filling stages for the next -- and the consumer
// warps never synchronise with each other (each owns its accumulators and its rows).
template <int NB>
__global__ void __launch_bounds__(K_THREADS, 1)
k_half_transform_kernel(const double *__restrict__ packed, long long L, int nt, int q_count,
                        const double *__restrict__ ctf, int nib, int n_ntiles, double *__restrict__ x, int nkc,
                        int nmb, const double *__restrict__ cep, double *__restrict__ gamma_part, int n_full,
                        int tail_split, int trim_last) {
  using Cfg = HalfCfg<NB>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double *stage_base = reinterpret_cast<double *>(smem_raw);
  uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem_raw + (size_t)Cfg::kStages * Cfg::kStageElems * sizeof(double));
  uint64_t *empty_bar = full_bar + Cfg::kStages;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // 32-bit throughout: q_count * nt fits easily (the launcher checks) and every 64-bit loop variable
  // costs two of the 168 registers this kernel is launched with
  const int total_rows = q_count * nt;                          // flattened (Q, tile-row)
  const int n_groups = (total_rows + K_SLOTS - 1) / K_SLOTS;
  const int n_units = n_groups * n_ntiles;
  // The last, partly filled round of the round-robin walk is cut finer: units from n_full on are
  // dealt as tail_split (2 or 4) pieces -- by warp column, then by the warp's slot -- to CTAs that
  // would otherwise idle.  A piece runs the same DMMA sequence on its share of the accumulators, so
  // X and the gamma partials keep their bits whatever the split.
  const int n_work = n_full + (n_units - n_full) * tail_split;
  // unit -> (row group, N tile).  The N tile is rotated by the round the group falls in, so that a CTA
  // meets every N tile in turn: the last one is cheaper when its padding block is trimmed (trim_last),
  // and a fixed tile per CTA would leave that saving as idle time at the end.
  const int groups_per_round = (int)gridDim.x / n_ntiles > 0 ? (int)gridDim.x / n_ntiles : 1;
  auto coords = [&](int unit, int &group, int &ntile) {
    group = unit / n_ntiles;
    ntile = (unit % n_ntiles + group / groups_per_round) % n_ntiles;
  };
  auto decode = [&](int work, int &unit, int &part_wn, int &part_sl) {
    unit = work; part_wn = -1; part_sl = -1;
    if (work >= n_full) {
      const int r = work - n_full;
      unit = n_full + r / tail_split;
      const int p = r % tail_split;
      part_wn = p & 1;
      part_sl = tail_split == 4 ? (p >> 1) : -1;
    }
  };

  if (threadIdx.x == 0) {
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], K_CONSUMER_WARPS);
    }
    fence_mbar_init();
  }
  __syncthreads();

  if (warp >= K_CONSUMER_WARPS) {
    reg_dealloc_producer();
    if (warp != K_CONSUMER_WARPS) return;
    // ===== TMA producer warp: lanes 0..7 own one slot each, lane 8 owns the C operand =====
    uint32_t it = 0;                                            // ring position, continuous over tiles
    for (int work = blockIdx.x; work < n_work; work += gridDim.x) {
      int unit, part_wn, part_sl;
      decode(work, unit, part_wn, part_sl);
      int group, ntile;
      coords(unit, group, ntile);
      const int f0 = group * K_SLOTS;
      const int ib0 = ntile * (2 * NB);
      // a piece that owns one slot per warp needs only those four rows of the tensor
      const bool load_slot = part_sl < 0 || (lane & 1) == part_sl;
      const uint32_t stage_bytes = (uint32_t)((part_sl < 0 ? Cfg::kAElems : Cfg::kAElems / 2) + Cfg::kBElems) * sizeof(double);
      // slots past the end of the tensor re-load the last row; their results are not stored
      int f = f0 + lane;
      if (f > total_rows - 1) f = total_rows - 1;
      const int q = f / nt, tr = f % nt;
      const double *row = packed + (size_t)q * L;
      for (int kc = 0; kc < nt; ++kc, ++it) {
        const int st = it % Cfg::kStages;
        const uint32_t ph = (it / Cfg::kStages) & 1;
        mbar_wait(&empty_bar[st], ph ^ 1);
        double *a_s = stage_base + (size_t)st * Cfg::kStageElems;
        if (lane == 0) mbar_arrive_expect_tx(&full_bar[st], stage_bytes);
        __syncwarp();
        if (lane < K_SLOTS) {
          const int a = tr > kc ? tr : kc, b = tr > kc ? kc : tr;
          if (load_slot)
            tma_load_1d(a_s + lane * TILE_ELEMS, row + (size_t)tile_index(a, b, nt) * TILE_ELEMS,
                      TILE_ELEMS * sizeof(double), &full_bar[st]);
        } else if (lane == K_SLOTS) {
          tma_load_1d(a_s + Cfg::kAElems, ctf + ((size_t)kc * nib + ib0) * 128, Cfg::kBElems * sizeof(double),
                      &full_bar[st]);
        }
      }
    }
    return;
  }

  // ===== consumer warps =====
  reg_alloc_consumer();
  const int wm = warp & 3, wn = warp >> 2;
  const int g = lane >> 2, t = lane & 3;
  const int lane_direct = lane;                                // (g*4 + t)
  const int lane_transp = ((g >> 2) << 5) + (t << 2) + (g & 3);
  uint32_t it = 0;

  // Fragments are double-buffered in registers and loaded one k-sub ahead of the DMMAs
  // that consume them -- across the stage boundary too: the wait on the next stage's
  // barrier and its first LDS are issued before the last DMMA block of the current stage,
  // so neither the mbarrier round trip nor the LDS latency is exposed.
  double fa[2][4], fb[2][NB];
  auto load_frags = [&](double (&a)[4], double (&b)[NB], int st, int ks, bool tp0, bool tp1) {
    const double *a_s = stage_base + (size_t)st * Cfg::kStageElems + (2 * wm) * TILE_ELEMS;
    const double *b_s = stage_base + (size_t)st * Cfg::kStageElems + Cfg::kAElems + (wn * NB) * 128;
    // direct:     ((rh*4 + ks)*32 + lane)
    // transposed: (((ks>>1)*4 + 2*rh)*32 + (ks&1)*16 + lane_transp)
#pragma unroll
    for (int sl = 0; sl < 2; ++sl) {
      const bool tp = sl ? tp1 : tp0;
#pragma unroll
      for (int rh = 0; rh < 2; ++rh) {
        const int off_d = (rh * 4 + ks) * 32 + lane_direct;
        const int off_t = (((ks >> 1) * 4 + 2 * rh) * 32) + (ks & 1) * 16 + lane_transp;
        a[sl * 2 + rh] = a_s[sl * TILE_ELEMS + (tp ? off_t : off_d)];
      }
    }
#pragma unroll
    for (int j = 0; j < NB; ++j) b[j] = b_s[(j * 4 + ks) * 32 + lane];
  };

  if (blockIdx.x < n_work) {                                   // prologue: first stage, k-sub 0 (kc = 0 is never transposed)
    mbar_wait(&full_bar[0], 0);
    load_frags(fa[0], fb[0], 0, 0, false, false);
  }

  // One work item.  MODE 0: the warp's whole share (both 16-row slots); 1 / 2: slot 0 / slot 1 only;
  // 3: nothing -- the warp only walks the ring, so the stage protocol is the same for whole units and
  // pieces.  The modes are compile-time: a predicated-off DMMA still takes its turn in the pipe.
  // NBW: n8-blocks of this warp -- NB, or NB - 1 for the right-hand warp column of the last N tile when the
  // occupied range ends at least eight columns short of the padded width (trim_last): that block would
  // multiply zeros, and the accumulation never reads its columns (ks_last).
  auto run_item = [&](auto mode_tag, auto nbw_tag, int work, int unit) {
    constexpr int MODE = decltype(mode_tag)::value;
    constexpr int NBW = decltype(nbw_tag)::value;
    constexpr int SL0 = MODE == 2 ? 1 : 0, SL1 = MODE == 0 ? 2 : (MODE == 3 ? 0 : SL0 + 1);   // slots [SL0, SL1)
    int group, ntile;
    coords(unit, group, ntile);
    const int f0 = group * K_SLOTS;
    const int ib0 = ntile * (2 * NB);
    const bool more_work = work + (int)gridDim.x < n_work;
    // the two 16-row slots of this warp
    int tr_s[2], q_s[2];
    bool ok_s[2];
#pragma unroll
    for (int sl = 0; sl < 2; ++sl) {
      int f = f0 + 2 * wm + sl;
      ok_s[sl] = f < total_rows;
      if (!ok_s[sl]) f = total_rows - 1;
      q_s[sl] = f / nt;
      tr_s[sl] = f % nt;
    }

    double acc[4][NBW][2];
#pragma unroll
    for (int m = 0; m < 4; ++m)
#pragma unroll
      for (int j = 0; j < NBW; ++j) { acc[m][j][0] = 0.0; acc[m][j][1] = 0.0; }

    for (int kc = 0; kc < nt; ++kc, ++it) {
      const int st = it % Cfg::kStages;
      const bool tp0 = tr_s[0] < kc, tp1 = tr_s[1] < kc;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        if (ks < 3) {
          if (MODE != 3) load_frags(fa[(ks + 1) & 1], fb[(ks + 1) & 1], st, ks + 1, tp0, tp1);
        } else if (kc + 1 < nt || more_work) {
          const uint32_t it1 = it + 1;
          const int st1 = it1 % Cfg::kStages;
          mbar_wait(&full_bar[st1], (it1 / Cfg::kStages) & 1);
          const bool in_tile = kc + 1 < nt;                    // the next tile starts at kc = 0: direct
          if (MODE != 3 || !in_tile)
            load_frags(fa[0], fb[0], st1, 0, in_tile && tr_s[0] < kc + 1, in_tile && tr_s[1] < kc + 1);
        }
#pragma unroll
        for (int m = 2 * SL0; m < 2 * SL1; ++m)
#pragma unroll
          for (int j = 0; j < NBW; ++j) dmma884(acc[m][j][0], acc[m][j][1], fa[ks & 1][m], fb[ks & 1][j]);
      }
      release_stage(&empty_bar[st], lane);
    }

    // epilogue: X[q][kc_out][mb][ks_out][g*4 + t_out], two adjacent i per thread (16-byte stores)
#pragma unroll
    for (int sl = SL0; sl < SL1; ++sl) {
      if (!ok_s[sl]) continue;
#pragma unroll
      for (int rh = 0; rh < 2; ++rh) {
        const int mb = 2 * tr_s[sl] + rh;
#pragma unroll
        for (int j = 0; j < NBW; ++j) {
          const int i0 = (ib0 + wn * NB + j) * 8;                 // first i of this n8 block
          const int kc_out = i0 >> 4;
          const int ks_out = ((i0 & 15) >> 2) + (t >> 1);
          const size_t off = (((size_t)q_s[sl] * nkc + kc_out) * nmb + mb) * 128 + ks_out * 32 + g * 4 + 2 * (t & 1);
          *reinterpret_cast<double2 *>(x + off) = make_double2(acc[sl * 2 + rh][j][0], acc[sl * 2 + rh][j][1]);
        }
      }
    }

    // Coulomb vector for free when the density is the one these orbitals build:
    // gamma_Q = sum B_Q.D = f * sum_{mu,i} X_Q[mu,i] C[mu,i]  (D = f C C^T).  Each warp reduces
    // its 16-row slot against the C fragments and writes one partial per (Q, tile-row,
    // N tile, N half); they are summed in fixed order by gamma_from_x_kernel.
    if (gamma_part != nullptr) {
#pragma unroll
      for (int sl = SL0; sl < SL1; ++sl) {
        double sum = 0.0;
        if (ok_s[sl]) {
#pragma unroll
          for (int rh = 0; rh < 2; ++rh) {
#pragma unroll
            for (int j = 0; j < NBW; ++j) {
              const int ib = ib0 + wn * NB + j;
              const double2 c2 = *reinterpret_cast<const double2 *>(
                  cep + ((((size_t)tr_s[sl] * nib + ib) * 2 + rh) * 32 + lane) * 2);
              sum = fma(acc[sl * 2 + rh][j][0], c2.x, sum);
              sum = fma(acc[sl * 2 + rh][j][1], c2.y, sum);
            }
          }
        }
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        if (lane == 0 && ok_s[sl])
          gamma_part[(((size_t)q_s[sl] * nt + tr_s[sl]) * n_ntiles + ntile) * 2 + wn] = sum;
      }
    }
  };

  using Whole = std::integral_constant<int, 0>;
  using AllBlocks = std::integral_constant<int, NB>;
  for (int work = blockIdx.x; work < n_work; work += gridDim.x) {
    if (work < n_full) {
      if constexpr (NB > 1) {
        int group, ntile;
        coords(work, group, ntile);
        if (trim_last && wn == 1 && ntile == n_ntiles - 1) {
          run_item(Whole{}, std::integral_constant<int, NB - 1>{}, work, work);
          continue;
        }
      }
      run_item(Whole{}, AllBlocks{}, work, work);
    } else {
      int unit, part_wn, part_sl;
      decode(work, unit, part_wn, part_sl);
      if (wn != part_wn) run_item(std::integral_constant<int, 3>{}, AllBlocks{}, work, unit);
      else if (part_sl < 0) run_item(Whole{}, AllBlocks{}, work, unit);
      else if (part_sl == 0) run_item(std::integral_constant<int, 1>{}, AllBlocks{}, work, unit);
      else run_item(std::integral_constant<int, 2>{}, AllBlocks{}, work, unit);
    }
  }
}

// ------------------------------------------------------------------------------------
// Accumulation (SYRK)
// ------------------------------------------------------------------------------------
// TILE = 128: 8 consumer warps (4x2, 32x64 warp tiles), one CTA per SM, setmaxnreg.
// TILE =  64: 4 consumer warps (2x2, 32x32 warp tiles) + 1 producer warp, three CTAs per
//             SM.  Less diagonal/padding waste (c2: 1.14x the triangle instead of 1.45x) at
//             twice the L2->smem traffic per flop.
template <int TILE>
struct SyrkCfg {
  static constexpr int kWarpsM = TILE == 128 ? 4 : 2;
  static constexpr int kWarpsN = 2;
  static constexpr int kConsumerWarps = kWarpsM * kWarpsN;
  static constexpr int kTN = TILE / (8 * kWarpsN);          // n8-blocks per warp
  static constexpr int kThreads = TILE == 128 ? K_THREADS : (kConsumerWarps + 1) * 32;
  static constexpr int kMinBlocks = TILE == 128 ? 1 : 3;
#ifdef MQCB200_DBG_STAGES
  static constexpr int kStages = MQCB200_DBG_STAGES;
#else
  static constexpr int kStages = 4;
#endif
  static constexpr int kPanelBlocks = TILE / 8;              // 8-row blocks per panel
  static constexpr int kPanelElems = kPanelBlocks * 128;     // * (4 k-subs * 32)
  static constexpr int kStageElems = 2 * kPanelElems;
  static constexpr size_t kSmemBytes = (size_t)kStages * kStageElems * sizeof(double) + 2 * kStages * sizeof(uint64_t);
};

// Consumer side of a DIAGONAL 64x64 tile: only the lower triangle is ever read back, so the
// 36 lower-triangular 8x8 blocks are dealt evenly -- warp W takes block-rows W and 7-W, nine
// blocks -- instead of 16 per warp on a full tile.  Same pipeline protocol as the full path.
template <int W, class Cfg>
__device__ __forceinline__ void syrk_diag_warp(const double *stage_base, uint64_t *full_bar, uint64_t *empty_bar,
                                               int n_steps, int nkc, int ks_last, int lane, double *out,
                                               int accumulate) {
  constexpr int R0 = W, R1 = 7 - W;               // the two 8-row blocks of this warp
  constexpr int NB0 = R0 + 1, NB1 = R1 + 1;       // their column blocks 0..R (NB1 > NB0)
  double acc0[NB0][2], acc1[NB1][2];
#pragma unroll
  for (int j = 0; j < NB0; ++j) { acc0[j][0] = 0.0; acc0[j][1] = 0.0; }
#pragma unroll
  for (int j = 0; j < NB1; ++j) { acc1[j][0] = 0.0; acc1[j][1] = 0.0; }
  double fa[2][2], fb[2][NB1];
  auto load_frags = [&](double (&a)[2], double (&b)[NB1], int st, int ks) {
    const double *p = stage_base + (size_t)st * Cfg::kStageElems + lane;
    a[0] = p[(R0 * 4 + ks) * 32];
    a[1] = p[(R1 * 4 + ks) * 32];
#pragma unroll
    for (int j = 0; j < NB1; ++j) b[j] = p[(j * 4 + ks) * 32];
  };
  if (n_steps > 0) {
    mbar_wait(&full_bar[0], 0);
    load_frags(fa[0], fb[0], 0, 0);
  }
  int kc_idx = 0;
  for (int step = 0; step < n_steps; ++step) {
    const int st = step % Cfg::kStages;
    const int ks_lim = (kc_idx == nkc - 1) ? ks_last : 4;
    kc_idx = (kc_idx + 1 == nkc) ? 0 : kc_idx + 1;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      if (ks < 3) {
        load_frags(fa[(ks + 1) & 1], fb[(ks + 1) & 1], st, ks + 1);
      } else if (step + 1 < n_steps) {
        const int st1 = (step + 1) % Cfg::kStages;
        mbar_wait(&full_bar[st1], ((step + 1) / Cfg::kStages) & 1);
        load_frags(fa[0], fb[0], st1, 0);
      }
      if (ks < ks_lim) {
#pragma unroll
        for (int j = 0; j < NB0; ++j) dmma884(acc0[j][0], acc0[j][1], fa[ks & 1][0], fb[ks & 1][j]);
#pragma unroll
        for (int j = 0; j < NB1; ++j) dmma884(acc1[j][0], acc1[j][1], fa[ks & 1][1], fb[ks & 1][j]);
      }
    }
    release_stage(&empty_bar[st], lane);
  }
  const int g = lane >> 2, t = lane & 3;
  auto store = [&](int rb, int cb, double v0, double v1) {
    double2 *p = reinterpret_cast<double2 *>(out + (8 * rb + g) * 64 + 8 * cb + 2 * t);
    double2 v = make_double2(v0, v1);
    if (accumulate) { const double2 o = *p; v.x += o.x; v.y += o.y; }
    *p = v;
  };
#pragma unroll
  for (int j = 0; j < NB0; ++j) store(R0, j, acc0[j][0], acc0[j][1]);
#pragma unroll
  for (int j = 0; j < NB1; ++j) store(R1, j, acc1[j][0], acc1[j][1]);
}

template <int TILE>
__global__ void __launch_bounds__(SyrkCfg<TILE>::kThreads, SyrkCfg<TILE>::kMinBlocks)
k_accumulate_kernel(const double *__restrict__ x, int nkc, int nmb, int q_count, int n_ktiles, int n_splits,
                    int n_splits_diag, int n_splits_edge, int n_panels, double *__restrict__ kpart, int accumulate,
                    int ks_last, int pair_off, int npairs, int edge_tiles) {
  using Cfg = SyrkCfg<TILE>;
  constexpr int PB = Cfg::kPanelBlocks;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double *stage_base = reinterpret_cast<double *>(smem_raw);
  uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem_raw + (size_t)Cfg::kStages * Cfg::kStageElems * sizeof(double));
  uint64_t *empty_bar = full_bar + Cfg::kStages;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // Units: (tile, split) pairs, class by class (k_unit_decode).  A diagonal tile does 9/16 of the DMMAs
  // (lower triangle only, evenly dealt to the four warps) and a tile of the last panel row LB/8 of them
  // (consume_edge), so they get proportionally longer auxiliary ranges: all units take the same time.
  int tile, split, mp = 0, np = 0;
  k_unit_decode((int)blockIdx.x, n_ktiles, n_panels, n_splits, n_splits_diag, n_splits_edge, mp, np, split);
  tile = mp * (mp + 1) / 2 + np;
  // rank-2 mode (pair_off > 0): chunk k of the X half is paired with chunk k + pair_off of the C half,
  // once as (rows of panel mp: X, rows of panel np: C) and once reversed; every tile -- the
  // diagonal ones too -- then takes the two-panel path, and a step list per auxiliary function is
  // 2*npairs long instead of nkc.
  const bool rank2 = pair_off > 0;
  const bool diag = mp == np && !rank2;
  const int my_splits = (mp == np) ? n_splits_diag : (mp == n_panels - 1 ? n_splits_edge : n_splits);
  const int qa = (int)((long long)split * q_count / my_splits);
  const int qb = (int)((long long)(split + 1) * q_count / my_splits);
  const int per_q = rank2 ? 2 * npairs : nkc;
  const int n_steps = (qb - qa) * per_q;
  const int valid_a = nmb - PB * mp < PB ? nmb - PB * mp : PB;   // 8-row blocks present in each panel
  const int valid_b = nmb - PB * np < PB ? nmb - PB * np : PB;

  // panels past the end of the matrix are never filled by TMA: keep them zero
#ifndef MQCB200_DBG_NOZERO
  for (int e = threadIdx.x; e < Cfg::kStages * Cfg::kStageElems; e += Cfg::kThreads) stage_base[e] = 0.0;
#endif
  if (threadIdx.x == 0) {
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], Cfg::kConsumerWarps);
    }
    fence_mbar_init();
  }
  fence_proxy_async();
  __syncthreads();

  if (warp >= Cfg::kConsumerWarps) {
    if (TILE == 128) reg_dealloc_producer();
    if (warp == Cfg::kConsumerWarps && lane == 0) {
      const uint32_t bytes_a = (uint32_t)valid_a * 128 * sizeof(double);
      const uint32_t bytes_b = diag ? 0u : (uint32_t)valid_b * 128 * sizeof(double);
      for (int step = 0; step < n_steps; ++step) {
        const int st = step % Cfg::kStages;
        const uint32_t ph = (step / Cfg::kStages) & 1;
        mbar_wait(&empty_bar[st], ph ^ 1);
        double *a_s = stage_base + (size_t)st * Cfg::kStageElems;
        const int r = step % per_q;
        const size_t qrow = (size_t)(qa + step / per_q) * nkc;
        const int ka = rank2 ? ((r & 1) ? (r >> 1) + pair_off : (r >> 1)) : r;
        const int kb = rank2 ? ((r & 1) ? (r >> 1) : (r >> 1) + pair_off) : r;
        mbar_arrive_expect_tx(&full_bar[st], bytes_a + bytes_b);
        tma_load_1d(a_s, x + ((qrow + ka) * nmb + PB * mp) * 128, bytes_a, &full_bar[st]);
        if (!diag) tma_load_1d(a_s + Cfg::kPanelElems, x + ((qrow + kb) * nmb + PB * np) * 128, bytes_b, &full_bar[st]);
      }
    }
    return;
  }

  if (TILE == 128) reg_alloc_consumer();
  if (TILE == 64 && diag) {
    double *out_d = kpart + ((size_t)split * n_ktiles + tile) * (TILE * TILE);
    switch (warp) {
      case 0: syrk_diag_warp<0, Cfg>(stage_base, full_bar, empty_bar, n_steps, nkc, ks_last, lane, out_d, accumulate); break;
      case 1: syrk_diag_warp<1, Cfg>(stage_base, full_bar, empty_bar, n_steps, nkc, ks_last, lane, out_d, accumulate); break;
      case 2: syrk_diag_warp<2, Cfg>(stage_base, full_bar, empty_bar, n_steps, nkc, ks_last, lane, out_d, accumulate); break;
      default: syrk_diag_warp<3, Cfg>(stage_base, full_bar, empty_bar, n_steps, nkc, ks_last, lane, out_d, accumulate); break;
    }
    return;
  }
  const int wm = warp % Cfg::kWarpsM, wn = warp / Cfg::kWarpsM;
  const int g = lane >> 2, t = lane & 3;
  constexpr int TN = Cfg::kTN;
  double *out = kpart + ((size_t)split * n_ktiles + tile) * (TILE * TILE);

  // MC: 8-row blocks of this warp that lie inside the matrix.  Four, except in the last panel row,
  // where the rows past n are padding nobody reads back (finalize_jk_kernel stops at n): a warp whose
  // 32 rows are half or wholly padding runs half or none of the DMMAs.  Compile-time, because a
  // predicated-off DMMA still takes its turn in the pipe.
  auto consume = [&](auto mc_tag) {
    constexpr int MC = decltype(mc_tag)::value;
    if constexpr (MC == 0) {
      for (int step = 0; step < n_steps; ++step) {          // nothing to compute: just walk the ring
        const int st = step % Cfg::kStages;
        mbar_wait(&full_bar[st], (step / Cfg::kStages) & 1);
        release_stage(&empty_bar[st], lane);
      }
    } else {
      double acc[MC][TN][2];
#pragma unroll
      for (int m = 0; m < MC; ++m)
#pragma unroll
        for (int j = 0; j < TN; ++j) { acc[m][j][0] = 0.0; acc[m][j][1] = 0.0; }

      // register double-buffered fragments, loaded one k-sub ahead (see the half-transform)
      double fa[2][MC], fb[2][TN];
      auto load_frags = [&](double (&a)[MC], double (&b)[TN], int st, int ks) {
        const double *a_s = stage_base + (size_t)st * Cfg::kStageElems + (4 * wm) * 128;
        const double *b_s = stage_base + (size_t)st * Cfg::kStageElems + (diag ? 0 : Cfg::kPanelElems) + (TN * wn) * 128;
#pragma unroll
        for (int m = 0; m < MC; ++m) a[m] = a_s[(m * 4 + ks) * 32 + lane];
#pragma unroll
        for (int j = 0; j < TN; ++j) b[j] = b_s[(j * 4 + ks) * 32 + lane];
      };
      if (n_steps > 0) {
        mbar_wait(&full_bar[0], 0);
        load_frags(fa[0], fb[0], 0, 0);
      }
      int kc_idx = 0;                                   // position of this step inside its auxiliary function
      const int last_from = rank2 ? per_q - 2 : per_q - 1;    // rank-2: both orders of the last pair
      for (int step = 0; step < n_steps; ++step) {
        const int st = step % Cfg::kStages;
        // the last 16-wide chunk of the padded occupied range may hold fewer than four valid
        // 4-wide k-subs (n_occ = 241: one): the rest is zero padding, skip its DMMAs
        const int ks_lim = (kc_idx >= last_from) ? ks_last : 4;
        kc_idx = (kc_idx + 1 == per_q) ? 0 : kc_idx + 1;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          if (ks < 3) {
            load_frags(fa[(ks + 1) & 1], fb[(ks + 1) & 1], st, ks + 1);
          } else if (step + 1 < n_steps) {
            const int st1 = (step + 1) % Cfg::kStages;
            mbar_wait(&full_bar[st1], ((step + 1) / Cfg::kStages) & 1);
            load_frags(fa[0], fb[0], st1, 0);
          }
          if (ks < ks_lim) {
#pragma unroll
            for (int m = 0; m < MC; ++m)
#pragma unroll
              for (int j = 0; j < TN; ++j) dmma884(acc[m][j][0], acc[m][j][1], fa[ks & 1][m], fb[ks & 1][j]);
          }
        }
        release_stage(&empty_bar[st], lane);
      }

#pragma unroll
      for (int m = 0; m < MC; ++m) {
        const int r = 32 * wm + 8 * m + g;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
          const int c = 8 * TN * wn + 8 * j + 2 * t;
          double2 *p = reinterpret_cast<double2 *>(out + r * TILE + c);
          double2 v = make_double2(acc[m][j][0], acc[m][j][1]);
          if (accumulate) { const double2 o = *p; v.x += o.x; v.y += o.y; }
          *p = v;
        }
      }
    }
  };
  // A tile of the last panel row holds LB < 8 live row blocks.  Skipping the dead blocks warp by warp
  // (consume<2>, consume<0>) idles two of the four sub-partitions while the CTA still waits for the other
  // two; here the LB x 8 live blocks are dealt evenly instead -- every warp takes all LB row blocks of two
  // column blocks -- so the whole CTA finishes in LB/8 of the time and its slot goes to the next unit.
  // Each element is still one accumulator summed over the same steps in the same order: same bits.
  auto consume_edge = [&](auto lb_tag) {
    constexpr int LB = decltype(lb_tag)::value;
    double acc[LB][2][2];
#pragma unroll
    for (int m = 0; m < LB; ++m)
#pragma unroll
      for (int j = 0; j < 2; ++j) { acc[m][j][0] = 0.0; acc[m][j][1] = 0.0; }
    double fa[2][LB], fb[2][2];
    auto load_frags = [&](double (&a)[LB], double (&b)[2], int st, int ks) {
      const double *a_s = stage_base + (size_t)st * Cfg::kStageElems;
      const double *b_s = stage_base + (size_t)st * Cfg::kStageElems + Cfg::kPanelElems + (2 * warp) * 128;
#pragma unroll
      for (int m = 0; m < LB; ++m) a[m] = a_s[(m * 4 + ks) * 32 + lane];
#pragma unroll
      for (int j = 0; j < 2; ++j) b[j] = b_s[(j * 4 + ks) * 32 + lane];
    };
    if (n_steps > 0) {
      mbar_wait(&full_bar[0], 0);
      load_frags(fa[0], fb[0], 0, 0);
    }
    int kc_idx = 0;
    const int last_from = rank2 ? per_q - 2 : per_q - 1;
    for (int step = 0; step < n_steps; ++step) {
      const int st = step % Cfg::kStages;
      const int ks_lim = (kc_idx >= last_from) ? ks_last : 4;
      kc_idx = (kc_idx + 1 == per_q) ? 0 : kc_idx + 1;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        if (ks < 3) {
          load_frags(fa[(ks + 1) & 1], fb[(ks + 1) & 1], st, ks + 1);
        } else if (step + 1 < n_steps) {
          const int st1 = (step + 1) % Cfg::kStages;
          mbar_wait(&full_bar[st1], ((step + 1) / Cfg::kStages) & 1);
          load_frags(fa[0], fb[0], st1, 0);
        }
        if (ks < ks_lim) {
#pragma unroll
          for (int m = 0; m < LB; ++m)
#pragma unroll
            for (int j = 0; j < 2; ++j) dmma884(acc[m][j][0], acc[m][j][1], fa[ks & 1][m], fb[ks & 1][j]);
        }
      }
      release_stage(&empty_bar[st], lane);
    }
#pragma unroll
    for (int m = 0; m < LB; ++m) {
      const int r = 8 * m + g;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int c = 16 * warp + 8 * j + 2 * t;
        double2 *p = reinterpret_cast<double2 *>(out + r * TILE + c);
        double2 v = make_double2(acc[m][j][0], acc[m][j][1]);
        if (accumulate) { const double2 o = *p; v.x += o.x; v.y += o.y; }
        *p = v;
      }
    }
  };
  if (TILE == 64 && edge_tiles && !diag && valid_a < PB) {
    // 8-row blocks are present in pairs (the matrix is padded to 16 rows): 2, 4 or 6 live ones
    if (valid_a <= 2) consume_edge(std::integral_constant<int, 2>{});
    else if (valid_a <= 4) consume_edge(std::integral_constant<int, 4>{});
    else consume_edge(std::integral_constant<int, 6>{});
    return;
  }
  const int live = valid_a - 4 * wm;
  if (TILE == 64 && live <= 0) consume(std::integral_constant<int, 0>{});
  else if (TILE == 64 && live == 2) consume(std::integral_constant<int, 2>{});
  else consume(std::integral_constant<int, 4>{});
}

// ------------------------------------------------------------------------------------
// planning + launch
// ------------------------------------------------------------------------------------
static int ktile_override() {
  static int v = -1;
  if (v < 0) {
    const char *e = getenv("MQCB200_KTILE");
    v = e ? atoi(e) : 0;
    if (v != 64 && v != 128) v = 0;
  }
  return v;
}

KPlan plan_k(int n, int n_occ_in, int q_count, size_t workspace_limit_bytes, int sm_count, int rank2_occ) {
  KPlan p;
  const int nt = num_tiles(n);
  // rank-2 form: the half-transform sees the stacked [X | C], each half padded to whole 16-wide chunks
  const int o16_r2 = rank2_occ > 0 ? (rank2_occ + 15) / 16 * 16 : 0;
  const int n_occ = rank2_occ > 0 ? 2 * o16_r2 : n_occ_in;
  p.pair_off = o16_r2 / 16;
  p.npairs = o16_r2 / 16;
  const int op16 = ((n_occ > 0 ? n_occ : 1) + 15) / 16 * 16;
  p.n_ntiles = (op16 + 127) / 128;
  int bn = (op16 + p.n_ntiles - 1) / p.n_ntiles;
  bn = (bn + 15) / 16 * 16;
  p.nb = bn / 16;
  p.nib = p.n_ntiles * 2 * p.nb;
  p.nkc = p.nib / 2;
  p.nmb = 2 * nt;
  {
    const int valid_last = (n_occ > 0 ? n_occ : 1) - 16 * (p.nkc - 1);   // may be <= 0 when N tiles pad a whole chunk
    p.ks_last = valid_last >= 16 ? 4 : (valid_last <= 0 ? 0 : (valid_last + 3) / 4);
    if (rank2_occ > 0) p.ks_last = (rank2_occ - 16 * (p.npairs - 1) + 3) / 4;   // of the last pair of chunks
  }
  // the last n8-block of the padded width: all padding, and past everything the accumulation reads?
  p.trim_last = (p.nb > 1 && p.nib * 8 - (n_occ > 0 ? n_occ : 1) >= 8 && !getenv("MQCB200_NO_TRIM")) ? 1 : 0;
  const int steps_per_q = rank2_occ > 0 ? 2 * p.npairs : p.nkc;
  // K tile edge: 64 trims the diagonal/padding waste of the SYRK; 128 halves the operand
  // traffic per flop and wins once n is large enough that the waste is small.
  p.ktile = ktile_override() ? ktile_override() : 64;
  p.n_panels = (n + p.ktile - 1) / p.ktile;
  p.n_ktiles = p.n_panels * (p.n_panels + 1) / 2;
  const int slots = sm_count * (p.ktile == 128 ? 1 : 3);
  p.sm_count = sm_count;
  p.gamma_stride = num_tiles(n) * p.n_ntiles * 2;
  p.x_elems_per_q = (size_t)p.nkc * p.nmb * 128;
  size_t qc = workspace_limit_bytes / (p.x_elems_per_q * sizeof(double));
  if (qc < 1) qc = 1;
  if (qc > (size_t)(q_count > 0 ? q_count : 1)) qc = (size_t)(q_count > 0 ? q_count : 1);
  p.q_chunk = (int)qc;
  // Split of the contraction range over CTAs.  Cost model (in pipeline steps of ~0.25 us):
  // waves * (steps per unit + ~12 of fixed per-CTA overhead) + what the later fixed-order sum
  // over the splits costs; small problems get a few dozen short units, big ones whole waves.
  // The partial buffer is capped at 256 MiB.
  const size_t tile_bytes = (size_t)p.ktile * p.ktile * sizeof(double);
  const int n_diag = p.n_panels, n_off = p.n_ktiles - p.n_panels;
  auto diag_splits = [&](int s) {           // 64-wide diagonal tiles cost 9/16 of a full tile
    if (p.ktile != 64 || rank2_occ > 0) return s;
    int sd = (9 * s + 8) / 16;
    return sd < 1 ? 1 : sd;
  };
  // Off-diagonal tiles of a partly filled last panel row: LB of their 8 row blocks are live (consume_edge),
  // so they may get LB/8 of the splits -- a third class of units.  Not in the rank-2 form.
  const int lb_edge = p.nmb - (p.ktile / 8) * (p.n_panels - 1);
  const bool edge_possible = p.ktile == 64 && rank2_occ == 0 && p.n_panels >= 2 && lb_edge < 8;
  auto edge_splits = [&](int s) {
    int se = (lb_edge * s + 4) / 8;
    return se < 1 ? 1 : se;
  };
  auto search = [&](bool use_edge, int &best) {
    best = 1;
    double best_cost = 1e300;
    const int n_edge = use_edge ? p.n_panels - 1 : 0;
    for (int s = 1; s <= 512 && s <= p.q_chunk; ++s) {
      if ((size_t)s * p.n_ktiles * tile_bytes > ((size_t)256 << 20) && s > 1) break;
      const int sd = diag_splits(s), se = use_edge ? edge_splits(s) : s;
      const long long units = (long long)s * (n_off - n_edge) + (long long)se * n_edge + (long long)sd * n_diag;
      const long long waves = (units + slots - 1) / slots;
      const double steps_off = (double)((p.q_chunk + s - 1) / s) * steps_per_q;
      const double steps_edge = n_edge > 0 ? (double)((p.q_chunk + se - 1) / se) * steps_per_q * lb_edge / 8.0 : 0.0;
      const double steps_diag = (double)((p.q_chunk + sd - 1) / sd) * steps_per_q * ((p.ktile == 64 && rank2_occ == 0) ? 9.0 / 16.0 : 1.0);
      double steps = n_off - n_edge > 0 ? (steps_off > steps_diag ? steps_off : steps_diag) : steps_diag;
      if (steps_edge > steps) steps = steps_edge;
      // + the fixed-order sum over the splits in finalize_jk_kernel (~3/4 of a step per split)
      double cost = (double)waves * (steps + 12.0) + 0.75 * (double)s;
      // a single wave of long units measures ~5 % slower than the same work in two waves of
      // shorter ones (c2: 2.28 vs 2.18 ms; co-resident CTAs in lockstep, no second wave to
      // absorb the stragglers)
      if (waves == 1 && units > slots / 2) cost *= 1.05;
      if (cost < best_cost * 0.995) { best_cost = cost; best = s; }
    }
    return best_cost;
  };
  int best = 1;
  const double base_cost = search(false, best);
  // The third class is taken only for launches of a few waves and only where the model sees more than 3 % in
  // it.  With many waves the hardware's own dispatch already hands the slot of an early-finishing edge unit to
  // the next unit, and re-cutting the splits only moves the wave boundary.  Measured in one call on one box
  // (profiles/r02d_split_classes_ab.log): c2 (two waves -> four shorter ones) 6.72 -> 6.52 ms per build;
  // c4 (15 waves) 313.6 ms on two classes against 315.4 (three classes, same 25 splits) and 317.4 (three
  // classes, 5 splits, which the model rated 1.8 % better).
  bool use_edge = false;
  if (edge_possible && !getenv("MQCB200_NO_EDGE_SPLITS")) {
    int best_e = 1;
    const double edge_cost = search(true, best_e);
    const long long base_units = (long long)best * n_off + (long long)diag_splits(best) * n_diag;
    const bool few_waves = (base_units + slots - 1) / slots < 8;
    if ((few_waves && edge_cost < 0.97 * base_cost) || getenv("MQCB200_EDGE_SPLITS")) { use_edge = true; best = best_e; }
  }
  if (const char *e = getenv("MQCB200_KSPLITS")) {   // development override
    const int v = atoi(e);
    if (v >= 1) best = v < p.q_chunk ? v : p.q_chunk;
  }
  p.n_splits = best;
  p.n_splits_diag = diag_splits(best);
  p.n_splits_edge = use_edge ? edge_splits(best) : best;
  p.kpart_elems = (size_t)p.n_splits * p.n_ktiles * p.ktile * p.ktile;
  return p;
}

// How the last round of the half-transform's round-robin walk is cut (see the kernel): the units
// that do not fill a whole round of `ctas` are dealt as 2 or 4 pieces each when that many pieces
// still fit one round.
HalfTail plan_half_tail(long long n_units, int ctas) {
  HalfTail t;
  t.n_full = n_units;
  t.split = 1;
  if (getenv("MQCB200_NO_TAIL_SPLIT") || ctas <= 0) return t;
  const long long rest = n_units % ctas;
  if (rest == 0) return t;
  int split = rest * 4 <= ctas ? 4 : (rest * 2 <= ctas ? 2 : 1);
  if (const char *e = getenv("MQCB200_TAIL_SPLIT")) split = atoi(e) == 4 ? 4 : (atoi(e) == 2 ? 2 : 1);   // development override
  if (split > 1) { t.n_full = n_units - rest; t.split = split; }
  return t;
}

template <int NB>
static void launch_half_nb(const double *d_packed, long long L, int n, int q_count, const double *d_ctf,
                           const double *d_cep, const KPlan &plan, double *d_x, double *d_gamma_part,
                           cudaStream_t s) {
  const int nt = num_tiles(n);
  const long long rows = (long long)q_count * nt;
  const long long n_units = (rows + K_SLOTS - 1) / K_SLOTS * plan.n_ntiles;
  if (rows > 0x3fffffffLL || n_units > 0x1fffffffLL) {     // the kernel indexes rows and work items with 32-bit ints
    std::fprintf(stderr, "mqcb200: half-transform chunk too large (%lld rows); lower the workspace limit\n", rows);
    std::abort();
  }
  const HalfTail tail = plan_half_tail(n_units, plan.sm_count);
  const long long n_work = tail.n_full + (n_units - tail.n_full) * tail.split;
  const unsigned grid = (unsigned)(n_work < plan.sm_count ? n_work : plan.sm_count);
  k_half_transform_kernel<NB><<<grid, K_THREADS, HalfCfg<NB>::kSmemBytes, s>>>(
      d_packed, L, nt, q_count, d_ctf, plan.nib, plan.n_ntiles, d_x, plan.nkc, plan.nmb, d_cep,
      d_cep ? d_gamma_part : nullptr, (int)tail.n_full, tail.split, plan.trim_last);
}

void launch_k_half_transform(const double *d_packed, long long L, int n, int q_count, const double *d_ctf,
                             const double *d_cep, const KPlan &plan, double *d_x, double *d_gamma_part,
                             cudaStream_t s) {
  switch (plan.nb) {
    case 1: launch_half_nb<1>(d_packed, L, n, q_count, d_ctf, d_cep, plan, d_x, d_gamma_part, s); break;
    case 2: launch_half_nb<2>(d_packed, L, n, q_count, d_ctf, d_cep, plan, d_x, d_gamma_part, s); break;
    case 3: launch_half_nb<3>(d_packed, L, n, q_count, d_ctf, d_cep, plan, d_x, d_gamma_part, s); break;
    case 4: launch_half_nb<4>(d_packed, L, n, q_count, d_ctf, d_cep, plan, d_x, d_gamma_part, s); break;
    case 5: launch_half_nb<5>(d_packed, L, n, q_count, d_ctf, d_cep, plan, d_x, d_gamma_part, s); break;
    case 6: launch_half_nb<6>(d_packed, L, n, q_count, d_ctf, d_cep, plan, d_x, d_gamma_part, s); break;
    case 7: launch_half_nb<7>(d_packed, L, n, q_count, d_ctf, d_cep, plan, d_x, d_gamma_part, s); break;
    default: launch_half_nb<8>(d_packed, L, n, q_count, d_ctf, d_cep, plan, d_x, d_gamma_part, s); break;
  }
}

void launch_k_accumulate(const double *d_x, int q_count, const KPlan &plan, double *d_kpart, int accumulate,
                         cudaStream_t s) {
  // A short last chunk still touches every (split, tile) partial -- splits with an empty
  // auxiliary range write (or add) zeros -- so the fixed-order sum in finalize is defined.
  const int splits = plan.n_splits;
  const int edge = getenv("MQCB200_NO_EDGE_TILES") ? 0 : 1;       // development switch (same bits either way)
  const unsigned units = (unsigned)k_unit_count(plan.n_ktiles, plan.n_panels, plan.n_splits, plan.n_splits_diag, plan.n_splits_edge);
  if (plan.ktile == 128)
    k_accumulate_kernel<128><<<units, SyrkCfg<128>::kThreads, SyrkCfg<128>::kSmemBytes, s>>>(
        d_x, plan.nkc, plan.nmb, q_count, plan.n_ktiles, splits, plan.n_splits_diag, plan.n_splits_edge, plan.n_panels,
        d_kpart, accumulate, plan.ks_last, plan.pair_off, plan.npairs, edge);
  else
    k_accumulate_kernel<64><<<units, SyrkCfg<64>::kThreads, SyrkCfg<64>::kSmemBytes, s>>>(
        d_x, plan.nkc, plan.nmb, q_count, plan.n_ktiles, splits, plan.n_splits_diag, plan.n_splits_edge, plan.n_panels,
        d_kpart, accumulate, plan.ks_last, plan.pair_off, plan.npairs, edge);
}

template <int NB>
static void configure_half() {
  cudaFuncSetAttribute(k_half_transform_kernel<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                       (int)HalfCfg<NB>::kSmemBytes);
}

void configure_kernels() {
  configure_fragment_kernels();
  configure_whiten_kernels();
  configure_scf_kernels();
  configure_half<1>(); configure_half<2>(); configure_half<3>(); configure_half<4>();
  configure_half<5>(); configure_half<6>(); configure_half<7>(); configure_half<8>();
  cudaFuncSetAttribute(k_accumulate_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SyrkCfg<128>::kSmemBytes);
  cudaFuncSetAttribute(k_accumulate_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SyrkCfg<64>::kSmemBytes);
}

}  // namespace mqcb200

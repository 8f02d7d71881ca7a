// common.cuh -- packed-tensor layout arithmetic and sm_100a PTX wrappers shared
// by the kernels of the DF-J/K engine.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace mqcb200 {

// ---------------------------------------------------------------------------
// Packed layout of one auxiliary slab B_Q (symmetric n x n), see DESIGN.md:
//   * the matrix is cut into 16x16 tiles; only tiles (tr, tc) with tr >= tc are
//     stored, ordered column by column:  idx(tr,tc) = tc*nt - tc*(tc-1)/2 + (tr-tc);
//   * diagonal tiles are stored as full symmetric 16x16 blocks, rows/cols >= n are 0;
//   * inside a tile the 256 doubles are in "DMMA fragment order": eight 8x4
//     blocks, element (r,c) at ((r>>3)*4 + (c>>2))*32 + (r&7)*4 + (c&3), so that a
//     warp's m8n8k4 A-fragment (and its transpose) is 256 contiguous bytes.
// One slab is L = ntl*256 doubles; slabs are contiguous (row stride L).
// ---------------------------------------------------------------------------
constexpr int TILE = 16;
constexpr int TILE_ELEMS = TILE * TILE;  // 256 doubles = 2 KiB

__host__ __device__ inline int num_tiles(int n) { return (n + TILE - 1) / TILE; }
__host__ __device__ inline long long num_lower_tiles(int nt) { return (long long)nt * (nt + 1) / 2; }
__host__ __device__ inline long long packed_row_len(int n) {
  return num_lower_tiles(num_tiles(n)) * TILE_ELEMS;
}
__host__ __device__ inline int tile_index(int tr, int tc, int nt) {  // tr >= tc
  return tc * nt - (tc * (tc - 1)) / 2 + (tr - tc);
}
__host__ __device__ inline int in_tile_offset(int r, int c) {
  return (((r >> 3) << 2) + (c >> 2)) * 32 + ((r & 7) << 2) + (c & 3);
}

// ---------------------------------------------------------------------------
// Work units of the exchange accumulation (k_accumulate_kernel): unit u -> K tile (mp, np), np <= mp,
// and split `split` of that tile's auxiliary range.  Tiles come in three classes with their own split
// counts, chosen so that every unit takes about the same time: off-diagonal tiles (n_splits),
// DIAGONAL tiles (n_splits_diag: 9/16 of the DMMAs) and, when n_splits_edge != n_splits, the
// off-diagonal tiles of the LAST panel row (n_splits_edge: only LB of their 8 row blocks are inside
// the matrix).  Used by the kernel and by the host-side planner check (tests/native/plan_check.cu).
// ---------------------------------------------------------------------------
__host__ __device__ inline void k_unit_decode(int u, int n_ktiles, int n_panels, int n_splits, int n_splits_diag,
                                              int n_splits_edge, int &mp, int &np, int &split) {
  const int n_off = n_ktiles - n_panels;
  if (n_splits_edge == n_splits) {
    // every tile for the first n_splits_diag splits, then only the off-diagonal tiles
    const int u_full = n_splits_diag * n_ktiles;
    if (u < u_full) {
      split = u / n_ktiles;
      const int tile = u % n_ktiles;
      mp = 0;
      while ((mp + 1) * (mp + 2) / 2 <= tile) ++mp;
      np = tile - mp * (mp + 1) / 2;
    } else {
      const int v = u - u_full;
      split = n_splits_diag + v / n_off;
      const int oi = v % n_off;                       // off-diagonal tiles: oi = mp*(mp-1)/2 + np, np < mp
      mp = 1;
      while ((mp + 1) * mp / 2 <= oi) ++mp;
      np = oi - mp * (mp - 1) / 2;
    }
    return;
  }
  // class by class: the full off-diagonal tiles (panel rows 1 .. n_panels-2), the last panel row, the diagonal
  const int n_edge = n_panels - 1, n_full = n_off - n_edge;
  int v = u;
  if (v < n_splits * n_full) {
    split = v / n_full;
    const int oi = v % n_full;
    mp = 1;
    while ((mp + 1) * mp / 2 <= oi) ++mp;
    np = oi - mp * (mp - 1) / 2;
    return;
  }
  v -= n_splits * n_full;
  if (v < n_splits_edge * n_edge) {
    split = v / n_edge;
    mp = n_panels - 1;
    np = v % n_edge;
    return;
  }
  v -= n_splits_edge * n_edge;
  split = v / n_panels;
  mp = np = v % n_panels;
}
__host__ __device__ inline long long k_unit_count(int n_ktiles, int n_panels, int n_splits, int n_splits_diag,
                                                  int n_splits_edge) {
  const long long n_off = n_ktiles - n_panels, n_edge = n_splits_edge == n_splits ? 0 : n_panels - 1;
  return (long long)n_splits * (n_off - n_edge) + (long long)n_splits_edge * n_edge + (long long)n_splits_diag * n_panels;
}

#ifdef __CUDACC__
// --------------------------- PTX wrappers ----------------------------------
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__device__ __forceinline__ uint64_t policy_evict_first();
__device__ __forceinline__ double2 ld_stream_f64x2(const double *p) {
  double2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}

// Same with an L2 eviction-priority hint: the Coulomb kernels stream the tensor once and must not
// push the exchange kernels' half-transformed panels out of L2 when the two run side by side.
__device__ __forceinline__ double2 ld_stream_f64x2_hint(const double *p, uint64_t policy) {
  double2 v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.f64 {%0,%1}, [%2], %3;"
               : "=d"(v.x), "=d"(v.y) : "l"(p), "l"(policy));
  return v;
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// 1-D bulk async copy global -> shared (TMA engine, SASS UBLKCP); bytes % 16 == 0,
// both addresses 16-byte aligned; completion is signalled on `bar` (complete_tx).
__device__ __forceinline__ void tma_load_1d(void *smem_dst, const void *gmem_src, uint32_t bytes,
                                            uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
// Same with an L2 eviction-priority hint (createpolicy result in `policy`).
__device__ __forceinline__ void tma_load_1d_hint(void *smem_dst, const void *gmem_src, uint32_t bytes,
                                                 uint64_t *bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, "
      "[%3], %4;" ::"r"(smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
// ---- warp-specialised DMMA kernels: shared constants and helpers -------------------
constexpr int K_CONSUMER_WARPS = 8;
// Eight consumer warps (two warpgroups) + one producer warpgroup of which only warp 8
// works.  Three warps share each SM sub-partition's 16K registers, so the launch
// allocation is capped at 168/thread; the producer warpgroup hands its share back
// (setmaxnreg.dec) and the consumers grow to 232 (setmaxnreg.inc): 2*232 + 40 <= 512.
constexpr int K_THREADS = (K_CONSUMER_WARPS + 4) * 32;
constexpr int K_PRODUCER_REGS = 40;
constexpr int K_CONSUMER_REGS = 232;

__device__ __forceinline__ void reg_dealloc_producer() {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(K_PRODUCER_REGS));
}
// Consumer-side release of a pipeline stage.  The stage was read through the generic
// proxy (LDS) and will be overwritten through the async proxy (bulk TMA): without a
// cross-proxy fence ptxas is free to hoist the mbarrier arrive above the DMMAs that
// wait for the LDS results, and the TMA refill then races the reads (seen on B200 as
// run-to-run differences with short pipelines, see profiles/r01_notes.md).
__device__ __forceinline__ void release_stage(uint64_t *empty_bar, int lane) {
  fence_proxy_async();
  __syncwarp();
  if (lane == 0) mbar_arrive(empty_bar);
}
__device__ __forceinline__ void reg_alloc_consumer() {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(K_CONSUMER_REGS));
}
constexpr int K_SLOTS = 8;                               // 16-row slots per CTA (BM = 128)

#endif  // __CUDACC__

}  // namespace mqcb200

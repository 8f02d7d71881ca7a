// xgpu_kernels.cu -- the one exchange step of a sharded build, [J|K] summed over the GPUs,
// as ONE kernel over NVLink peer memory instead of a library all-reduce.
//
// Every rank has its partial [J|K(|K_b)] in a buffer the other ranks can address (CUDA IPC).
// The kernel on rank r
//   1. tells every peer "my partial is complete" (a flag store into the peer's memory),
//      and waits until every peer has said so;
//   2. owns 1/N of the elements: reads that slice from ALL ranks' partials over NVLink, adds
//      them in rank order (fixed order: every rank ends up with bit-identical sums) and
//      PUSHES the result into every rank's output buffer;
//   3. tells every peer "my slice has landed" and waits for theirs, so that when the kernel
//      exits the whole reduced [J|K] is in local memory for the assembly kernel that follows.
// Per rank: (N-1)/N of the span read and written over NVLink, two flag round trips.  The grid
// is small enough to be co-resident (the CTAs wait on one another's GPUs, never on their own
// later CTAs).  Spin loops give up after `timeout_ns` of wall time (%globaltimer; 30 s by
// default, MQCB200_XGPU_TIMEOUT_S) and raise an error flag in HOST-mapped memory rather than
// hang: the host sees it without a synchronisation, also after an asynchronous build.
#include "common.cuh"
#include "kernels.cuh"

namespace mqcb200 {

constexpr int XG_THREADS = 256;

__device__ __forceinline__ unsigned long long ld_flag(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_flag(unsigned long long *p, unsigned long long v) {
  asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ bool wait_flag(const unsigned long long *p, unsigned long long epoch,
                                          unsigned long long timeout_ns) {
  if (ld_flag(p) >= epoch) return true;
  const unsigned long long t0 = global_ns();
  while (ld_flag(p) < epoch) {
    __nanosleep(64);
    if (global_ns() - t0 > timeout_ns) return false;
  }
  return true;
}

__global__ void __launch_bounds__(XG_THREADS) xgpu_allreduce_kernel(XgpuPeers peers, int n_ranks, int rank,
                                                                    unsigned long long epoch, size_t first,
                                                                    size_t count, unsigned int *counter,
                                                                    volatile int *error_flag,
                                                                    unsigned long long timeout_ns) {
  unsigned long long *my_flags = peers.flags[rank];
  __shared__ int ok_s;
  if (threadIdx.x == 0) ok_s = 1;
  __syncthreads();

  // ---- 1. my partial is complete (stream order) -> publish; wait for every peer's
  if (blockIdx.x == 0 && threadIdx.x < n_ranks) {
    __threadfence_system();
    st_flag(peers.flags[threadIdx.x] + rank, epoch);
  }
  if (threadIdx.x < n_ranks && !wait_flag(my_flags + threadIdx.x, epoch, timeout_ns)) ok_s = 0;
  __syncthreads();
  if (!ok_s) {
    if (threadIdx.x == 0) { *error_flag = 1; __threadfence_system(); }
    return;
  }

  // ---- 2. reduce my slice in rank order, push it to everybody
  size_t per = (count + n_ranks - 1) / n_ranks;
  per = (per + 31) / 32 * 32;
  const size_t lo = (size_t)rank * per < count ? (size_t)rank * per : count;
  const size_t hi = lo + per < count ? lo + per : count;
  // 16-byte accesses, and the loads of two vectors from every rank in flight before the first add: the
  // slice costs about one NVLink round trip instead of one per element.  The sums keep the rank order.
  const bool vec_ok = ((first + lo) & 1) == 0;                  // buffers are 256-byte aligned; lo is a multiple of 32
  const size_t n_vec = vec_ok ? (hi - lo) / 2 : 0;
  const size_t stride = (size_t)gridDim.x * XG_THREADS;
  for (size_t v = (size_t)blockIdx.x * XG_THREADS + threadIdx.x; v < n_vec; v += 2 * stride) {
    const size_t e0 = first + lo + 2 * v, e1 = e0 + 2 * stride;
    const bool two = v + stride < n_vec;
    double2 a[XGPU_MAX_RANKS], b[XGPU_MAX_RANKS];
#pragma unroll
    for (int k = 0; k < XGPU_MAX_RANKS; ++k) {
      if (k < n_ranks) {
        a[k] = *reinterpret_cast<const double2 *>(peers.in[k] + e0);
        if (two) b[k] = *reinterpret_cast<const double2 *>(peers.in[k] + e1);
      }
    }
    double2 sa = make_double2(0.0, 0.0), sb = make_double2(0.0, 0.0);
#pragma unroll
    for (int k = 0; k < XGPU_MAX_RANKS; ++k) {
      if (k < n_ranks) {
        sa.x += a[k].x; sa.y += a[k].y;
        if (two) { sb.x += b[k].x; sb.y += b[k].y; }
      }
    }
#pragma unroll
    for (int k = 0; k < XGPU_MAX_RANKS; ++k) {
      if (k < n_ranks) {
        *reinterpret_cast<double2 *>(peers.out[k] + e0) = sa;
        if (two) *reinterpret_cast<double2 *>(peers.out[k] + e1) = sb;
      }
    }
  }
  for (size_t e = lo + 2 * n_vec + (size_t)blockIdx.x * XG_THREADS + threadIdx.x; e < hi; e += stride) {
    double s = 0.0;
    for (int k = 0; k < n_ranks; ++k) s += peers.in[k][first + e];
    for (int k = 0; k < n_ranks; ++k) peers.out[k][first + e] = s;
  }

  // ---- 3. all of my CTAs done -> publish; wait until every peer's slice has landed here
  __threadfence_system();
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) last = atomicAdd(counter, 1u) == gridDim.x - 1;
  __syncthreads();
  if (last) {
    if (threadIdx.x == 0) *counter = 0u;
    if (threadIdx.x < n_ranks) {
      __threadfence_system();
      st_flag(peers.flags[threadIdx.x] + XGPU_MAX_RANKS + rank, epoch);
    }
  }
  if (threadIdx.x < n_ranks && !wait_flag(my_flags + XGPU_MAX_RANKS + threadIdx.x, epoch, timeout_ns)) ok_s = 0;
  __syncthreads();
  if (!ok_s && threadIdx.x == 0) { *error_flag = 1; __threadfence_system(); }
}

void launch_xgpu_allreduce(const XgpuPeers &peers, int n_ranks, int rank, unsigned long long epoch, size_t first,
                           size_t count, unsigned int *d_counter, int *d_error, unsigned long long timeout_ns,
                           cudaStream_t s) {
  size_t per = (count + n_ranks - 1) / n_ranks;
  unsigned blocks = (unsigned)((per + XG_THREADS * 4 - 1) / (XG_THREADS * 4));
  if (blocks < 1) blocks = 1;
  if (blocks > 128) blocks = 128;          // co-resident by construction (148 SMs, 256 threads, no smem)
  xgpu_allreduce_kernel<<<blocks, XG_THREADS, 0, s>>>(peers, n_ranks, rank, epoch, first, count, d_counter, d_error,
                                                      timeout_ns);
}

}  // namespace mqcb200

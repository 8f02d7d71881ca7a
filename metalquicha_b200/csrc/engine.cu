// engine.cu -- the context behind the C ABI of include/mqcb200.h: device binding,
// grow-only pools, tensor slots, the build sequence, NCCL (loaded lazily), the
// fragment FIFO.  No CPU fallback exists anywhere in this file: every entry point
// either runs the CUDA path or fails with a message.
#include "../../include/mqcb200.h"
#include "common.cuh"
#include "kernels.cuh"

#include <dlfcn.h>
#include <unistd.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

namespace mqcb200 {

static thread_local std::string g_last_error;

struct Failure : std::runtime_error {
  using std::runtime_error::runtime_error;
};

#define CUDA_CHECK(expr)                                                                          \
  do {                                                                                            \
    cudaError_t err__ = (expr);                                                                   \
    if (err__ != cudaSuccess)                                                                     \
      throw Failure(std::string("mqcb200: CUDA error '") + cudaGetErrorString(err__) + "' in " + \
                    #expr + " (" __FILE__ ":" + std::to_string(__LINE__) + ")");                  \
  } while (0)

// ---- grow-only device buffer (cf. device_pool_t of the reference's GPU backend,
// backends/cuest/backend/mqc_cuest_context.f90:40-53) -------------------------------
struct DevBuf {
  void *ptr = nullptr;
  size_t cap = 0;
  void ensure(size_t bytes) {
    if (bytes <= cap) return;
    if (ptr) CUDA_CHECK(cudaFree(ptr));
    ptr = nullptr;
    cap = 0;
    CUDA_CHECK(cudaMalloc(&ptr, bytes));
    cap = bytes;
  }
  void release() {
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    cap = 0;
  }
  double *d() const { return static_cast<double *>(ptr); }
};

// grow-only pinned host buffer (staging of small operand sets: one DMA instead of several)
struct PinnedBuf {
  void *ptr = nullptr;
  size_t cap = 0;
  void ensure(size_t bytes) {
    if (bytes <= cap) return;
    if (ptr) CUDA_CHECK(cudaFreeHost(ptr));
    ptr = nullptr;
    cap = 0;
    CUDA_CHECK(cudaHostAlloc(&ptr, bytes, cudaHostAllocDefault));
    cap = bytes;
  }
  void release() {
    if (ptr) cudaFreeHost(ptr);
    ptr = nullptr;
    cap = 0;
  }
  double *d() const { return static_cast<double *>(ptr); }
};

struct TensorSlot {
  DevBuf packed;
  int n = 0;
  int naux_total = 0;
  int q_begin = 0;
  int q_count = 0;
  long long L = 0;
  bool set = false;
};

// ---- NCCL through dlopen: the library must load on a box without NCCL ----------------
struct Id128 {
  char bytes[128];
};
struct NcclApi {
  void *lib = nullptr;
  int (*GetUniqueId)(void *) = nullptr;
  int (*CommInitRank)(void **, int, Id128 /* ncclUniqueId, passed by value */, int) = nullptr;
  int (*AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
  int (*AllGather)(const void *, void *, size_t, int, void *, cudaStream_t) = nullptr;
  int (*CommDestroy)(void *) = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;
static std::mutex g_nccl_mutex;

static void load_nccl() {
  std::lock_guard<std::mutex> lock(g_nccl_mutex);
  if (g_nccl.lib) return;
  const char *names[] = {"libnccl.so.2", "libnccl.so"};
  void *lib = nullptr;
  for (const char *nm : names) {
    lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (lib) break;
  }
  if (!lib) throw Failure("mqcb200: NCCL is required for a multi-GPU build but libnccl.so.2 could not be loaded");
  auto sym = [&](const char *nm) {
    void *p = dlsym(lib, nm);
    if (!p) throw Failure(std::string("mqcb200: NCCL symbol missing: ") + nm);
    return p;
  };
  g_nccl.GetUniqueId = reinterpret_cast<decltype(g_nccl.GetUniqueId)>(sym("ncclGetUniqueId"));
  g_nccl.CommInitRank = reinterpret_cast<decltype(g_nccl.CommInitRank)>(sym("ncclCommInitRank"));
  g_nccl.AllReduce = reinterpret_cast<decltype(g_nccl.AllReduce)>(sym("ncclAllReduce"));
  g_nccl.AllGather = reinterpret_cast<decltype(g_nccl.AllGather)>(sym("ncclAllGather"));
  g_nccl.CommDestroy = reinterpret_cast<decltype(g_nccl.CommDestroy)>(sym("ncclCommDestroy"));
  g_nccl.GetErrorString = reinterpret_cast<decltype(g_nccl.GetErrorString)>(sym("ncclGetErrorString"));
  g_nccl.lib = lib;
}
#define NCCL_CHECK(expr)                                                                   \
  do {                                                                                     \
    int rc__ = (expr);                                                                     \
    if (rc__ != 0)                                                                         \
      throw Failure(std::string("mqcb200: NCCL error '") + g_nccl.GetErrorString(rc__) + \
                    "' in " + #expr);                                                      \
  } while (0)
constexpr int kNcclFloat64 = 8;  // ncclDouble
constexpr int kNcclSum = 0;
constexpr int kNcclChar = 0;     // ncclInt8 / ncclChar

// ---- the engine ------------------------------------------------------------------------
constexpr uint32_t kMagic = 0x4D51B200u;

struct Engine {
  uint32_t magic = kMagic;
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;   // H and D ride up here while the K kernels (which need only C) run
  cudaStream_t j_stream = nullptr;      // the HBM-bound Coulomb kernels, concurrent with the tensor-bound exchange kernels
  cudaEvent_t ev_late_upload = nullptr, ev_fork = nullptr, ev_k1_done = nullptr, ev_j_done = nullptr;
  bool overlap_j = true;
  // ... optionally with pass 2 as the 64-thread TMA-fed kernel that fits beside k_accumulate's three CTAs per SM.
  // Measured (profiles/r02_notes.md): the 16 KiB it can keep in flight per SM stream only ~1.1 TB/s, so it
  // outlasts the accumulation it hides behind (c2 7.95 ms vs 6.72 ms) -- off by default, MQCB200_CORESIDENT_J=1 to try.
  bool coresident_j = false;
  TensorSlot slots[MQCB200_NUM_SLOTS];
  size_t workspace_limit = (size_t)4 << 30;
  size_t fuse_threshold = (size_t)32 << 20;  // below this a pass over B is cheaper than the extra launch of the fused route

  // per-build device operands and scratch (grow-only)
  DevBuf d_w, d_ctf, d_gamma_partial, d_gamma, d_jpart, d_x, d_kpart;
  DevBuf d_gpart_a, d_gpart_b;   // Coulomb-vector partials written by the half-transform
  DevBuf d_cep;                  // coefficients in accumulator order for that epilogue
  DevBuf d_stack;                // [X | C] of a rank-2 exchange, each half on a 16-column boundary
  DevBuf d_scf;                  // device-resident SCF of a fragment: H, S, X, F, D, C, eps, work, DIIS history, scalars
  PinnedBuf h_scf;               // its per-iteration scalars on the host side
  int scf_check_every = 1;       // iterations queued between two looks at the convergence flag
  int scf_last_sweeps = 0;       // Jacobi sweeps of the last diagonalisation of the general-size SCF
  DevBuf d_jk;       // [J | K | K_beta] contiguous so one all-reduce covers them
  DevBuf d_scalar, d_stage, d_escratch;
  DevBuf d_in;                   // [H | D | C_a | C_b] of a host-operand build, contiguous
  DevBuf d_out;                  // [8 scalars | F_a | F_b] of a host-operand build, contiguous
  PinnedBuf h_in, h_out;         // pinned staging for operand sets up to kSmallIoBytes
  PinnedBuf h_stage[2];          // double-buffered pinned staging of lower-triangle slabs (set_tensor)
  DevBuf d_stage_lower[2];
  cudaEvent_t ev_stage[2] = {nullptr, nullptr};
  double whiten_ms = 0.0, whiten_flops = 0.0;
  // slab-streamed whitening in progress (mqcb200_whiten_begin .. _end)
  struct Whiten {
    bool active = false;
    int slot = 0, n = 0, naux = 0, m_rows = 0;
    DevBuf d_af, d_slab, d_tp;
    WhitenDst dst{};
    void *mapped[WHITEN_MAX_RANKS] = {};
    bool multi = false;
    bool failed = false;           // a push failed on this rank of a sharded whitening: whiten_end tells the others
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    double gemm_ms = 0.0, flops = 0.0;
  } wh;
  double metric_ms = 0.0;
  int metric_sweeps = 0;
  double set_tensor_ms = 0.0, set_tensor_h2d_bytes = 0.0;   // last host -> packed upload (stream time incl. host gathers)
  bool last_fuse_attempted = false;
  int last_n = 0;    // shape of the operands of the last build_fock (for last_energy)
  bool have_last_fock = false;
  double last_energy_host = 0.0;   // E = 1/2 sum D (H + F) of the last closed-shell build_fock (host-operand path)

  // multi-GPU
  void *comm = nullptr;
  int n_ranks = 1, rank = 0;
  // NVLink peer-memory exchange (xgpu_kernels.cu): every rank's partial/reduced [J|K] buffers and
  // arrival flags are mapped into every other rank through CUDA IPC
  struct P2P {
    bool ready = false;
    bool failed = false;                          // an exchange timed out: the next sharded build resynchronises the ranks
    XgpuPeers peers{};
    void *flags = nullptr, *misc = nullptr;       // own flags [2*XGPU_MAX_RANKS] u64; misc: CTA counter
    int *host_error = nullptr, *dev_error = nullptr;   // pinned host int the exchange kernel raises, and its device alias
    void *in = nullptr, *out = nullptr;           // own partial / reduced buffers
    size_t cap = 0;                               // bytes of each of in/out
    unsigned long long epoch = 0;
    unsigned long long timeout_ns = 30000000000ull;
    bool same_process[XGPU_MAX_RANKS] = {};       // peer k lives in this process: raw pointer + peer access, no IPC handle
  } p2p;

  // instrumentation: every occurrence of a phase gets its own event pair; a build's
  // per-phase time is the sum over its occurrences (K runs once per Q-chunk and spin)
  bool profiling = false;
  struct Span { int idx; cudaEvent_t a, b; };
  std::vector<Span> spans;          // spans[0..n_spans) are live for the last build
  size_t n_spans = 0;
  double last_ms[MQCB200_NUM_TIMERS] = {};
  int launches = 0;

  void bind() { CUDA_CHECK(cudaSetDevice(device)); }

  void phase_begin(int idx, cudaStream_t on = nullptr) {
    if (!profiling) return;
    if (!on) on = stream;
    if (n_spans == spans.size()) {
      Span sp{idx, nullptr, nullptr};
      CUDA_CHECK(cudaEventCreate(&sp.a));
      CUDA_CHECK(cudaEventCreate(&sp.b));
      spans.push_back(sp);
    }
    spans[n_spans].idx = idx;
    CUDA_CHECK(cudaEventRecord(spans[n_spans].a, on));
    ++n_spans;
  }
  void phase_end(int idx, cudaStream_t on = nullptr) {
    if (!profiling) return;
    if (!on) on = stream;
    for (size_t i = n_spans; i-- > 0;)
      if (spans[i].idx == idx) {
        CUDA_CHECK(cudaEventRecord(spans[i].b, on));
        return;
      }
  }
  // Spans accumulate over builds until mqcb200_last_timings collects and clears them, so
  // a caller can time a whole loop of asynchronous builds without a host sync per build.
  void collect_timers() {
    for (int i = 0; i < MQCB200_NUM_TIMERS; ++i) last_ms[i] = 0.0;
    for (size_t i = 0; i < n_spans; ++i) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, spans[i].a, spans[i].b) == cudaSuccess) last_ms[spans[i].idx] += ms;
    }
    n_spans = 0;
  }
};

static Engine *as_engine(void *h) {
  Engine *e = static_cast<Engine *>(h);
  if (!e || e->magic != kMagic) return nullptr;
  return e;
}

static bool fragment_path_enabled() {
  static int v = -1;
  if (v < 0) {
    const char *env = getenv("MQCB200_NO_FRAGMENT_PATH");   // development switch: always take the general kernels
    v = (env && env[0] == '1') ? 0 : 1;
  }
  return v == 1;
}

static bool use_cooperative_jacobi() {
  static int v = -1;
  if (v < 0) {
    const char *env = getenv("MQCB200_JACOBI_PER_ROUND");   // development switch: one kernel per round instead
    v = (env && env[0] == '1') ? 0 : 1;
  }
  return v == 1;
}

static bool fuse_gamma_enabled() {
  static int v = -1;
  if (v < 0) {
    const char *env = getenv("MQCB200_NO_FUSE_GAMMA");   // development switch: always take the pass over B
    v = (env && env[0] == '1') ? 0 : 1;
  }
  return v == 1;
}

constexpr size_t kSmallIoBytes = (size_t)1 << 20;

enum { T_UPLOAD = 0, T_J1, T_J2, T_K1, T_K2, T_FINAL, T_ALLREDUCE, T_DOWNLOAD };

// ------------------------------- NVLink peer-memory exchange -------------------------------
static bool p2p_enabled_by_env() {
  const char *env = getenv("MQCB200_P2P_ALLREDUCE");    // "0" = use the NCCL all-reduce instead
  return !(env && env[0] == '0');
}

static unsigned long long p2p_timeout_ns() {
  const char *env = getenv("MQCB200_XGPU_TIMEOUT_S");   // how long a rank waits for a peer before it gives up
  double sec = env ? atof(env) : 30.0;
  if (!(sec > 0.0)) sec = 30.0;
  return (unsigned long long)(sec * 1.0e9);
}

// Collective: true iff EVERY rank passed true.  Also a barrier (4-byte NCCL all-gather).
static bool ranks_agree(Engine *e, bool mine) {
  DevBuf sb, rb;
  std::vector<int> all(e->n_ranks, 0);
  int v = mine ? 1 : 0;
  try {
    sb.ensure(sizeof(int));
    rb.ensure(sizeof(int) * e->n_ranks);
    CUDA_CHECK(cudaMemcpyAsync(sb.ptr, &v, sizeof(int), cudaMemcpyHostToDevice, e->stream));
    NCCL_CHECK(g_nccl.AllGather(sb.ptr, rb.ptr, sizeof(int), kNcclChar, e->comm, e->stream));
    CUDA_CHECK(cudaMemcpyAsync(all.data(), rb.ptr, sizeof(int) * e->n_ranks, cudaMemcpyDeviceToHost, e->stream));
    CUDA_CHECK(cudaStreamSynchronize(e->stream));
  } catch (...) {
    sb.release();
    rb.release();
    throw;
  }
  sb.release();
  rb.release();
  bool everyone = true;
  for (int x : all) everyone = everyone && x == 1;
  return everyone;
}

// What a rank tells the others about one of its buffers.  Ranks in OTHER processes map it through
// the CUDA IPC handle; ranks that live in the SAME process (one host thread per GPU, the
// dispatcher shape of SURVEY 3.2) cannot open their own process's handles and use the raw
// pointer with peer access enabled instead.
struct PeerInfo {
  int pid;
  int device;
  unsigned long long ptr;
  cudaIpcMemHandle_t handle;
};

// All ranks call this together (the all-gather inside is unconditional, so a local failure can
// never leave the ranks in different collectives): every rank learns a device-addressable
// pointer to the buffer `local` of every other rank.  Returns false -- with everything it had
// opened closed again -- when this rank could not map all of its peers.
static bool p2p_exchange(Engine *e, void *local, void **mapped /*[n_ranks]*/) {
  PeerInfo mine{};
  bool ok = true;
  mine.pid = (int)getpid();
  mine.device = e->device;
  mine.ptr = (unsigned long long)(uintptr_t)local;
  if (!local || cudaIpcGetMemHandle(&mine.handle, local) != cudaSuccess) {
    ok = false;
    cudaGetLastError();
  }
  const size_t hs = sizeof(PeerInfo);
  DevBuf sendb, recvb;
  std::vector<PeerInfo> all(e->n_ranks);
  try {
    sendb.ensure(hs);
    recvb.ensure(hs * e->n_ranks);
    CUDA_CHECK(cudaMemcpyAsync(sendb.ptr, &mine, hs, cudaMemcpyHostToDevice, e->stream));
    NCCL_CHECK(g_nccl.AllGather(sendb.ptr, recvb.ptr, hs, kNcclChar, e->comm, e->stream));
    CUDA_CHECK(cudaMemcpyAsync(all.data(), recvb.ptr, hs * e->n_ranks, cudaMemcpyDeviceToHost, e->stream));
    CUDA_CHECK(cudaStreamSynchronize(e->stream));
  } catch (...) {
    sendb.release();
    recvb.release();
    throw;
  }
  sendb.release();
  recvb.release();
  for (int k = 0; k < e->n_ranks; ++k) mapped[k] = nullptr;
  for (int k = 0; k < e->n_ranks && ok; ++k) {
    if (k == e->rank) { mapped[k] = local; continue; }
    if (all[k].pid == mine.pid) {
      int can = 0;
      if (all[k].device != e->device) {
        if (cudaDeviceCanAccessPeer(&can, e->device, all[k].device) != cudaSuccess || !can) { ok = false; cudaGetLastError(); break; }
        cudaError_t pe = cudaDeviceEnablePeerAccess(all[k].device, 0);
        if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) { ok = false; cudaGetLastError(); break; }
        cudaGetLastError();
      }
      mapped[k] = (void *)(uintptr_t)all[k].ptr;
      e->p2p.same_process[k] = true;
    } else {
      void *p = nullptr;
      if (cudaIpcOpenMemHandle(&p, all[k].handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = false; cudaGetLastError(); break; }
      mapped[k] = p;
      e->p2p.same_process[k] = false;
    }
  }
  if (!ok) {
    for (int k = 0; k < e->n_ranks; ++k) {
      if (k != e->rank && mapped[k] && !e->p2p.same_process[k]) cudaIpcCloseMemHandle(mapped[k]);
      mapped[k] = nullptr;
    }
  }
  return ok;
}

static void p2p_unmap(Engine *e, void **mapped) {
  for (int k = 0; k < e->n_ranks; ++k) {
    if (k != e->rank && mapped[k] && !e->p2p.same_process[k]) cudaIpcCloseMemHandle(mapped[k]);
    mapped[k] = nullptr;
  }
}

static void p2p_free_local(Engine *e) {
  Engine::P2P &x = e->p2p;
  if (x.flags) cudaFree(x.flags);
  if (x.misc) cudaFree(x.misc);
  if (x.host_error) cudaFreeHost(x.host_error);
  if (x.in) cudaFree(x.in);
  if (x.out) cudaFree(x.out);
  x = Engine::P2P{};
}

// Collective: called by comm_init on every rank (n_ranks >= 2).  Every fallible local step comes
// first, then the ranks agree, and only then are handles exchanged; unless EVERY rank mapped all
// of its peers the build uses the NCCL all-reduce (a mixed choice would deadlock).
static void p2p_setup(Engine *e) {
  Engine::P2P &x = e->p2p;
  x = Engine::P2P{};
  bool ok = e->n_ranks <= XGPU_MAX_RANKS && p2p_enabled_by_env();
  const size_t fbytes = 2 * XGPU_MAX_RANKS * sizeof(unsigned long long);
  if (ok) {
    ok = cudaMalloc(&x.flags, fbytes) == cudaSuccess && cudaMalloc(&x.misc, 64) == cudaSuccess &&
         cudaHostAlloc(reinterpret_cast<void **>(&x.host_error), sizeof(int), cudaHostAllocMapped) == cudaSuccess;
    if (ok) {
      *x.host_error = 0;
      ok = cudaHostGetDevicePointer(reinterpret_cast<void **>(&x.dev_error), x.host_error, 0) == cudaSuccess &&
           cudaMemsetAsync(x.flags, 0, fbytes, e->stream) == cudaSuccess &&
           cudaMemsetAsync(x.misc, 0, 64, e->stream) == cudaSuccess && cudaStreamSynchronize(e->stream) == cudaSuccess;
    }
    cudaGetLastError();
  }
  if (!ranks_agree(e, ok)) { p2p_free_local(e); return; }
  void *mapped[XGPU_MAX_RANKS] = {};
  const bool mapped_ok = p2p_exchange(e, x.flags, mapped);
  if (!ranks_agree(e, mapped_ok)) {
    if (mapped_ok) p2p_unmap(e, mapped);
    p2p_free_local(e);
    return;
  }
  for (int k = 0; k < e->n_ranks; ++k) x.peers.flags[k] = static_cast<unsigned long long *>(mapped[k]);
  x.cap = 0;
  x.epoch = 0;
  x.timeout_ns = p2p_timeout_ns();
  x.ready = true;
}

// Collective: make sure every rank's partial/reduced buffers hold `bytes` and are mapped
// everywhere.  Growth is rare (grow-only, with headroom) and is itself a collective:
// unmap everywhere -> allocate -> agree -> exchange (the all-gathers are the barriers) -> agree.
static void p2p_ensure_buffers(Engine *e, size_t bytes) {
  Engine::P2P &x = e->p2p;
  if (bytes <= x.cap) return;
  CUDA_CHECK(cudaStreamSynchronize(e->stream));
  void *old_in[XGPU_MAX_RANKS], *old_out[XGPU_MAX_RANKS];
  for (int k = 0; k < e->n_ranks; ++k) {
    old_in[k] = const_cast<double *>(x.peers.in[k]);
    old_out[k] = x.peers.out[k];
    x.peers.in[k] = nullptr;
    x.peers.out[k] = nullptr;
  }
  if (x.cap > 0) { p2p_unmap(e, old_in); p2p_unmap(e, old_out); }
  const size_t new_cap = std::max(bytes, 2 * x.cap);
  x.cap = 0;
  void *new_in = nullptr, *new_out = nullptr;
  bool ok = cudaMalloc(&new_in, new_cap) == cudaSuccess && cudaMalloc(&new_out, new_cap) == cudaSuccess;
  cudaGetLastError();
  // the agreement is also the barrier after which nobody still has the old buffers mapped
  const bool all_ok = ranks_agree(e, ok);
  if (x.in) cudaFree(x.in);
  if (x.out) cudaFree(x.out);
  x.in = x.out = nullptr;
  if (!all_ok) {
    if (new_in) cudaFree(new_in);
    if (new_out) cudaFree(new_out);
    throw Failure("mqcb200: a rank could not allocate the peer-mapped exchange buffers");
  }
  x.in = new_in;
  x.out = new_out;
  void *m_in[XGPU_MAX_RANKS] = {}, *m_out[XGPU_MAX_RANKS] = {};
  const bool ok_in = p2p_exchange(e, new_in, m_in);
  const bool ok_out = p2p_exchange(e, new_out, m_out);
  if (!ranks_agree(e, ok_in && ok_out)) {
    if (ok_in) p2p_unmap(e, m_in);
    if (ok_out) p2p_unmap(e, m_out);
    throw Failure("mqcb200: a rank could not map its peers' exchange buffers");
  }
  for (int k = 0; k < e->n_ranks; ++k) {
    x.peers.in[k] = static_cast<const double *>(m_in[k]);
    x.peers.out[k] = static_cast<double *>(m_out[k]);
  }
  x.cap = new_cap;
}

static void p2p_teardown(Engine *e) {
  Engine::P2P &x = e->p2p;
  if (x.flags) {
    void *m[XGPU_MAX_RANKS];
    for (int k = 0; k < e->n_ranks; ++k) m[k] = x.peers.flags[k];
    p2p_unmap(e, m);
    if (x.cap > 0) {
      for (int k = 0; k < e->n_ranks; ++k) m[k] = const_cast<double *>(x.peers.in[k]);
      p2p_unmap(e, m);
      for (int k = 0; k < e->n_ranks; ++k) m[k] = x.peers.out[k];
      p2p_unmap(e, m);
    }
  }
  p2p_free_local(e);
}

// A peer that did not arrive in time leaves garbage in this build's [J|K]; the exchange kernel
// says so in host-mapped memory, so no synchronisation is needed to see it -- also after an
// asynchronous build, whose failure is reported by the NEXT call on the handle.
static void p2p_raise_if_failed(Engine *e) {
  Engine::P2P &x = e->p2p;
  if (!x.ready || !x.host_error || *x.host_error == 0) return;
  *x.host_error = 0;
  x.failed = true;
  throw Failure("mqcb200: the cross-GPU exchange timed out waiting for a peer rank (a rank failed or fell behind by "
                "more than MQCB200_XGPU_TIMEOUT_S); the result of that build is invalid");
}

// Collective, entered by the first sharded build after a failed exchange (every rank of a failed
// exchange sees the failure: whoever is waited for in vain is itself left waiting for the rank that
// gave up): zero the arrival flags and the CTA counter, restart the epochs, carry on.
static void p2p_recover(Engine *e) {
  Engine::P2P &x = e->p2p;
  CUDA_CHECK(cudaStreamSynchronize(e->stream));
  ranks_agree(e, true);                           // nobody is inside an exchange kernel any more
  const size_t fbytes = 2 * XGPU_MAX_RANKS * sizeof(unsigned long long);
  CUDA_CHECK(cudaMemsetAsync(x.flags, 0, fbytes, e->stream));
  CUDA_CHECK(cudaMemsetAsync(x.misc, 0, 64, e->stream));
  CUDA_CHECK(cudaStreamSynchronize(e->stream));
  *x.host_error = 0;
  ranks_agree(e, true);                           // nobody publishes before everyone has zeroed
  x.epoch = 0;
  x.failed = false;
}

// ------------------------------- tensor set-up ------------------------------------------
static void slot_prepare(Engine *e, TensorSlot &sl, int n, int naux_total, int q_begin, int q_count) {
  if (n <= 0) throw Failure("mqcb200: the orbital dimension n must be positive");
  if (naux_total <= 0 || q_count < 0 || q_begin < 0 || q_begin + q_count > naux_total)
    throw Failure("mqcb200: auxiliary range [q_begin, q_begin+q_count) does not fit naux_total");
  sl.set = false;
  sl.n = n;
  sl.naux_total = naux_total;
  sl.q_begin = q_begin;
  sl.q_count = q_count;
  sl.L = packed_row_len(n);
  sl.packed.ensure(std::max<size_t>(16, (size_t)sl.L * (size_t)q_count * sizeof(double)));
  (void)e;
}

// Host full-square slabs -> device, lower triangles only: the packed layout never reads
// mu < nu, so only n(n+1)/2 of the n*n doubles of a slab cross PCIe.  The caller's array is
// pageable (a Fortran allocatable), which the driver would stage through its own pinned
// buffer anyway; here the staging copy gathers just the lower-triangle column segments
// (column nu: rows nu..n-1, contiguous) into one of TWO pinned buffers, so that gathering
// chunk k+1 on the host overlaps the DMA and the packing kernel of chunk k.
static size_t lower_len(int n) { return (size_t)n * (n + 1) / 2; }

// One slab range.  A fragment's columns are short (n = 72: 36 doubles on average), and a memcpy call per
// column costs more than the bytes it moves; here a column goes in fixed 64-byte pieces, the last piece running
// over into the next column's place -- which the next copy then overwrites -- so there is no tail handling at
// all (2.0 -> 0.83 ms per trimer tensor on the build host).  The last columns of the range are copied exactly:
// nothing is read past the caller's array or written past this range's share of dst.  (A piece runs over by
// at most 7 doubles, which must stay inside the NEXT slab of the range: true once a slab's triangle holds 7
// doubles, so matrices smaller than 8 are copied exactly throughout.)
#if defined(__x86_64__)
__attribute__((target("avx2")))
#endif
static void gather_lower_range(const double *src, int n, size_t q_lo, size_t q_hi, size_t src_slab_stride, double *dst) {
  const size_t tri = lower_len(n);
  for (size_t q = q_lo; q < q_hi; ++q) {
    const double *s_q = src + q * src_slab_stride;
    double *d = dst + q * tri;
    const int exact_from = (q + 1 == q_hi || n < 8) ? (n > 8 ? n - 8 : 0) : n;
    for (int nu = 0; nu < n; ++nu) {
      const double *sc = s_q + (size_t)nu * n + nu;
      const int len = n - nu;
      if (nu >= exact_from) {
        std::memcpy(d, sc, (size_t)len * sizeof(double));
      } else {
        for (int i = 0; i < len; i += 8) __builtin_memcpy(d + i, sc + i, 64);
      }
      d += len;
    }
  }
}

void gather_lower(const double *src, int n, size_t q_count, size_t src_slab_stride, double *dst) {
  const size_t tri = lower_len(n);
  auto work = [&](size_t q_lo, size_t q_hi) {
#if defined(__x86_64__)
    static const bool avx2 = __builtin_cpu_supports("avx2");
    if (avx2) { gather_lower_range(src, n, q_lo, q_hi, src_slab_stride, dst); return; }
#endif
    for (size_t q = q_lo; q < q_hi; ++q) {
      const double *s_q = src + q * src_slab_stride;
      double *d_q = dst + q * tri;
      size_t off = 0;
      for (int nu = 0; nu < n; ++nu) {
        std::memcpy(d_q + off, s_q + (size_t)nu * n + nu, (size_t)(n - nu) * sizeof(double));
        off += (size_t)(n - nu);
      }
    }
  };
  const size_t bytes = q_count * tri * sizeof(double);
  unsigned hw = std::thread::hardware_concurrency();
  size_t n_threads = bytes >= ((size_t)16 << 20) ? std::min<size_t>({(size_t)8, (size_t)(hw ? hw : 1), q_count}) : 1;
  if (n_threads <= 1) { work(0, q_count); return; }
  std::vector<std::thread> pool;
  for (size_t t = 0; t < n_threads; ++t)
    pool.emplace_back(work, q_count * t / n_threads, q_count * (t + 1) / n_threads);
  for (auto &th : pool) th.join();
}

// `consume(d_lower, q0, qc)` is called (stream-ordered) for every staged chunk of lower-packed slabs.
template <class Consume>
static void stream_lower_slabs(Engine *e, const double *b, int n, size_t q_count, size_t slab_stride, Consume consume) {
  if (q_count == 0) return;
  const size_t tri = lower_len(n);
  size_t chunk = std::max<size_t>(1, ((size_t)64 << 20) / (tri * sizeof(double)));
  chunk = std::min(chunk, q_count);
  const size_t n_chunks = (q_count + chunk - 1) / chunk;
  const int n_buf = n_chunks > 1 ? 2 : 1;
  for (int i = 0; i < n_buf; ++i) {
    e->h_stage[i].ensure(chunk * tri * sizeof(double));
    e->d_stage_lower[i].ensure(chunk * tri * sizeof(double));
    if (!e->ev_stage[i]) CUDA_CHECK(cudaEventCreateWithFlags(&e->ev_stage[i], cudaEventDisableTiming));
  }
  for (size_t k = 0, q0 = 0; q0 < q_count; ++k, q0 += chunk) {
    const size_t qc = std::min(chunk, q_count - q0);
    const int buf = (int)(k % n_buf);
    if (k >= (size_t)n_buf) CUDA_CHECK(cudaEventSynchronize(e->ev_stage[buf]));   // its previous DMA + packing are done
    gather_lower(b + q0 * slab_stride, n, qc, slab_stride, e->h_stage[buf].d());
    CUDA_CHECK(cudaMemcpyAsync(e->d_stage_lower[buf].ptr, e->h_stage[buf].ptr, qc * tri * sizeof(double),
                               cudaMemcpyHostToDevice, e->stream));
    consume(e->d_stage_lower[buf].d(), q0, qc);
    CUDA_CHECK(cudaGetLastError());
    CUDA_CHECK(cudaEventRecord(e->ev_stage[buf], e->stream));
  }
  CUDA_CHECK(cudaStreamSynchronize(e->stream));
}

static void set_tensor_host(Engine *e, int slot, int n, int naux_total, int q_begin, int q_count,
                            const double *b) {
  if (slot < 0 || slot >= MQCB200_NUM_SLOTS) throw Failure("mqcb200: tensor slot out of range");
  if (!b && q_count > 0) throw Failure("mqcb200: null tensor pointer");
  e->bind();
  TensorSlot &sl = e->slots[slot];
  slot_prepare(e, sl, n, naux_total, q_begin, q_count);
  cudaEvent_t t0 = nullptr, t1 = nullptr;
  CUDA_CHECK(cudaEventCreate(&t0));
  CUDA_CHECK(cudaEventCreate(&t1));
  CUDA_CHECK(cudaEventRecord(t0, e->stream));
  stream_lower_slabs(e, b, n, (size_t)q_count, (size_t)n * n, [&](const double *d_lower, size_t q0, size_t qc) {
    launch_pack_tensor_lower(d_lower, n, (int)qc, sl.packed.d() + q0 * (size_t)sl.L, e->stream);
  });
  CUDA_CHECK(cudaEventRecord(t1, e->stream));
  CUDA_CHECK(cudaEventSynchronize(t1));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, t0, t1);
  cudaEventDestroy(t0);
  cudaEventDestroy(t1);
  e->set_tensor_ms = ms;
  e->set_tensor_h2d_bytes = (double)q_count * (double)lower_len(n) * sizeof(double);
  sl.set = true;
}

// ------------------------------- the build ------------------------------------------------
struct BuildArgs {
  int slot = 0;
  // exactly one of host/device operand sets is used
  bool device_operands = false;
  const double *h = nullptr, *density = nullptr;
  const double *coeff_a = nullptr, *coeff_b = nullptr;
  int lda = 0, ldb = 0, n_a = 0, n_b = 0;
  bool two_spin = false;          // K per spin with factor 1 instead of one K with factor 2
  bool rank2 = false;             // K = sum_P (B_P Ca)(B_P Cb)^T + (B_P Cb)(B_P Ca)^T  (n_a == n_b columns), factor 1
  bool want_j = true, want_k = true;
  // outputs (host pointers unless device_operands)
  double *j = nullptr, *k_a = nullptr, *k_b = nullptr, *fock_a = nullptr, *fock_b = nullptr;
  double k_scale = 1.0, j_scale = 1.0;
  bool assemble = false;
  bool combine = false;           // G = J + ka_coef*Ka + kb_coef*Kb into fock_a (host pointer)
  double ka_coef = 0.0, kb_coef = 0.0;
  bool sync = true;
};

static void run_k(Engine *e, const TensorSlot &sl, const double *d_coeff, int ldc, int n_occ, double *d_kpart_out,
                  double *d_gamma_part, KPlan &plan, cudaEvent_t ev_last_half = nullptr, int rank2_occ = 0) {
  const int n = sl.n;
  plan = plan_k(n, n_occ, sl.q_count, e->workspace_limit, e->sm_count, rank2_occ);
  e->d_ctf.ensure((size_t)num_tiles(n) * plan.nib * 128 * sizeof(double));
  e->d_x.ensure(plan.x_elems_per_q * (size_t)plan.q_chunk * sizeof(double));
  double *d_cep = nullptr;
  if (d_gamma_part) {
    e->d_cep.ensure((size_t)num_tiles(n) * plan.nib * 128 * sizeof(double));
    d_cep = e->d_cep.d();
  }
  launch_pack_coeff(d_coeff, ldc, n, n_occ, plan.nib, e->d_ctf.d(), d_cep, e->stream);
  e->launches += 1;
  for (int q0 = 0, chunk = 0; q0 < sl.q_count; q0 += plan.q_chunk, ++chunk) {
    const int qc = std::min(plan.q_chunk, sl.q_count - q0);
    e->phase_begin(T_K1);
    launch_k_half_transform(sl.packed.d() + (size_t)q0 * sl.L, sl.L, n, qc, e->d_ctf.d(), d_cep, plan, e->d_x.d(),
                            d_gamma_part ? d_gamma_part + (size_t)q0 * plan.gamma_stride : nullptr, e->stream);
    e->phase_end(T_K1);
    if (ev_last_half && q0 + plan.q_chunk >= sl.q_count) CUDA_CHECK(cudaEventRecord(ev_last_half, e->stream));
    e->phase_begin(T_K2);
    launch_k_accumulate(e->d_x.d(), qc, plan, d_kpart_out, chunk > 0, e->stream);
    e->phase_end(T_K2);
    e->launches += 2;
  }
  CUDA_CHECK(cudaGetLastError());
}

static void build(Engine *e, const BuildArgs &a) {
  if (a.slot < 0 || a.slot >= MQCB200_NUM_SLOTS) throw Failure("mqcb200: tensor slot out of range");
  TensorSlot &sl = e->slots[a.slot];
  if (!sl.set) throw Failure("mqcb200: no fitted tensor has been set on this slot (call mqcb200_set_tensor first)");
  if (!a.density && a.want_j) throw Failure("mqcb200: null density");
  if (a.n_a < 0 || a.n_b < 0) throw Failure("mqcb200: negative occupied count");
  if (a.want_k && a.n_a > 0 && (!a.coeff_a || a.lda < sl.n)) throw Failure("mqcb200: bad coefficient matrix / leading dimension");
  if (a.want_k && a.two_spin && a.n_b > 0 && (!a.coeff_b || a.ldb < sl.n)) throw Failure("mqcb200: bad beta coefficient matrix / leading dimension");
  if (a.rank2 && a.want_k && (a.n_a != a.n_b || a.n_a <= 0 || !a.coeff_b || a.ldb < sl.n || a.two_spin))
    throw Failure("mqcb200: the rank-2 exchange needs two factors with the same, positive number of columns");
  e->bind();
  e->launches = 0;
  const int n = sl.n;
  const size_t nn = (size_t)n * n;
  const bool sharded = e->comm != nullptr;
  if (sharded && sl.naux_total == sl.q_count && e->n_ranks > 1)
    throw Failure("mqcb200: a communicator is active but this slot holds the whole tensor; use mqcb200_set_tensor_shard");
  if (sharded && e->n_ranks > 1 && e->p2p.ready) {
    p2p_raise_if_failed(e);              // an asynchronous build's exchange gave up since the last call
    if (e->p2p.failed) p2p_recover(e);   // collective: every rank of a failed exchange comes through here
  }

  // ---- operands on the device
  const double *d_h = nullptr, *d_density = nullptr, *d_ca = nullptr, *d_cb = nullptr;
  int lda = n, ldb = n;
  bool small_io = false;
  const double *late_h = nullptr, *late_d = nullptr;   // host H / D still to be uploaded (large operand sets)
  size_t late_oh = 0, late_od = 0;
  bool late_issued = false;
  auto late_upload = [&]() {
    if (!late_h && !late_d) return;
    late_issued = true;
    double *din = e->d_in.d();
    if (late_d) CUDA_CHECK(cudaMemcpyAsync(din + late_od, late_d, nn * sizeof(double), cudaMemcpyHostToDevice, e->copy_stream));
    if (late_h) CUDA_CHECK(cudaMemcpyAsync(din + late_oh, late_h, nn * sizeof(double), cudaMemcpyHostToDevice, e->copy_stream));
    CUDA_CHECK(cudaEventRecord(e->ev_late_upload, e->copy_stream));
    CUDA_CHECK(cudaStreamWaitEvent(e->stream, e->ev_late_upload, 0));
    late_h = late_d = nullptr;
  };
  {
    e->phase_begin(T_UPLOAD);
    if (a.device_operands) {
      d_h = a.h; d_density = a.density; d_ca = a.coeff_a; d_cb = a.coeff_b;
      lda = a.lda; ldb = a.ldb;
    } else {
      // all operands go into one device block; small sets are first gathered in pinned
      // memory so that the whole upload is ONE DMA (a fragment's 72x72 matrices would
      // otherwise pay the driver's per-copy overhead three times)
      const bool up_ca = a.want_k && a.n_a > 0, up_cb = a.want_k && (a.two_spin || a.rank2) && a.n_b > 0;
      const size_t o_h = 0, o_d = o_h + (a.h ? nn : 0), o_ca = o_d + (a.density ? nn : 0);
      const size_t o_cb = o_ca + (up_ca ? (size_t)n * a.n_a : 0), in_elems = o_cb + (up_cb ? (size_t)n * a.n_b : 0);
      e->d_in.ensure(std::max<size_t>(16, in_elems * sizeof(double)));
      double *din = e->d_in.d();
      small_io = in_elems * sizeof(double) <= kSmallIoBytes;
      auto put = [&](size_t off, const double *src, int rows, int cols, int ld, cudaStream_t st) {
        if (small_io) {
          double *dst = e->h_in.d() + off;
          if (ld == rows) std::memcpy(dst, src, (size_t)rows * cols * sizeof(double));
          else for (int c = 0; c < cols; ++c) std::memcpy(dst + (size_t)rows * c, src + (size_t)ld * c, (size_t)rows * sizeof(double));
        } else if (ld == rows) {
          CUDA_CHECK(cudaMemcpyAsync(din + off, src, (size_t)rows * cols * sizeof(double), cudaMemcpyHostToDevice, st));
        } else {
          CUDA_CHECK(cudaMemcpy2DAsync(din + off, (size_t)rows * sizeof(double), src, (size_t)ld * sizeof(double),
                                       (size_t)rows * sizeof(double), cols, cudaMemcpyHostToDevice, st));
        }
      };
      if (small_io) e->h_in.ensure(std::max<size_t>(16, in_elems * sizeof(double)));
      // The coefficients go first: the K kernels need nothing else.  For operand sets too
      // large for the single staged DMA, H and D are issued later (late_upload below), after
      // the K kernels have been launched, on a second stream -- their transfer (and, for
      // pageable host memory, the driver's staging) then overlaps the exchange build.
      if (up_ca) { put(o_ca, a.coeff_a, n, a.n_a, a.lda, e->stream); d_ca = din + o_ca; }
      if (up_cb) { put(o_cb, a.coeff_b, n, a.n_b, a.ldb, e->stream); d_cb = din + o_cb; }
      if (a.h) d_h = din + o_h;
      if (a.density) d_density = din + o_d;
      late_h = a.h; late_d = a.density; late_oh = o_h; late_od = o_d;
      if (small_io) {
        if (a.h) put(o_h, a.h, n, n, n, e->stream);
        if (a.density) put(o_d, a.density, n, n, n, e->stream);
        late_h = late_d = nullptr;
      }
      if (small_io && in_elems > 0)
        CUDA_CHECK(cudaMemcpyAsync(din, e->h_in.ptr, in_elems * sizeof(double), cudaMemcpyHostToDevice, e->stream));
    }
    e->phase_end(T_UPLOAD);
  }

  const bool do_j = a.want_j;
  const bool do_ka = a.want_k && a.n_a > 0;
  const bool do_kb = a.want_k && a.two_spin && a.n_b > 0;
  // [J | K_a | K_b] in one buffer: a single exchange covers whatever was built.  With the
  // NVLink peer-memory exchange the partials live in the IPC-mapped buffer instead.
  const bool use_p2p = sharded && e->n_ranks > 1 && e->p2p.ready;
  if (use_p2p) p2p_ensure_buffers(e, 3 * nn * sizeof(double));
  else e->d_jk.ensure(3 * nn * sizeof(double));
  double *d_j = use_p2p ? static_cast<double *>(e->p2p.in) : e->d_jk.d(), *d_ka = d_j + nn, *d_kb = d_j + 2 * nn;

  const bool have = sl.q_count > 0;
  const double kfac = (a.two_spin || a.rank2) ? 1.0 : 2.0;
  JPlan jp = plan_j(n, sl.q_count);
  int n_jslices = jp.n_slices;          // partial-buffer shapes handed to finalize below
  int n_ksplits = 0, n_ksplits_diag = -1, n_ksplits_edge = -1, ktile = 64;
  bool ka_finalized = false;
  const int max_occ = std::max(do_ka ? a.n_a : 0, do_kb ? a.n_b : 0);
  const bool frag = have && (do_j || do_ka || do_kb) && fragment_path_enabled() && !a.rank2 &&
                    fragment_path_applies(n, std::max(max_occ, 1));

  if (frag) {
    // ---- fragment-sized problem: J and K in ONE kernel, each slab read once
    e->last_fuse_attempted = false;
    late_upload();
    auto run_frag = [&](const double *d_c, int ldc, int n_occ, bool wj, bool wk) {
      FragPlan fp = plan_fragment(n, std::max(n_occ, 1), sl.q_count, e->sm_count);
      if (wj) {
        e->d_w.ensure((size_t)sl.L * sizeof(double));
        e->d_jpart.ensure(fp.jpart_elems * sizeof(double));
        launch_pack_density(d_density, n, e->d_w.d(), nullptr, e->stream);
        e->launches += 1;
      }
      if (wk) {
        e->d_ctf.ensure((size_t)fp.nt * fp.nib * 128 * sizeof(double));
        e->d_kpart.ensure(fp.kpart_elems * sizeof(double));
        launch_pack_coeff(d_c, ldc, n, n_occ, fp.nib, e->d_ctf.d(), nullptr, e->stream);
        e->launches += 1;
      }
      e->phase_begin(wk ? T_K1 : T_J1);
      launch_fragment_jk(sl.packed.d(), sl.q_count, e->d_w.d(), e->d_ctf.d(), fp, wj, wk, e->d_jpart.d(), e->d_kpart.d(),
                         e->stream);
      e->phase_end(wk ? T_K1 : T_J1);
      e->launches += 1;
      return fp;
    };
    FragPlan fa = run_frag(d_ca, lda, a.n_a, do_j, do_ka);
    n_jslices = fa.grid;
    n_ksplits = fa.grid;
    if (do_kb) {
      if (do_ka) {
        launch_finalize_jk(nullptr, 0, e->d_kpart.d(), fa.grid, 64, n, kfac, nullptr, d_ka, e->stream);
        e->launches += 1;
        ka_finalized = true;
      }
      FragPlan fb = run_frag(d_cb, ldb, a.n_b, false, true);   // stream-ordered after the finalize above
      launch_finalize_jk(nullptr, 0, e->d_kpart.d(), fb.grid, 64, n, kfac, nullptr, d_kb, e->stream);
      e->launches += 1;
    }
    CUDA_CHECK(cudaGetLastError());
  } else {

  // ---- Coulomb vector from the half-transform when the density is the orbitals' own
  // (D = f C C^T: every SCF iteration).  Decided on the device by a consistency check fused
  // into the density packing, so an arbitrary density (guess, response, user-supplied)
  // silently takes the general pass over B instead.
  const bool fuse = do_j && do_ka && have && fuse_gamma_enabled() && !a.rank2 &&
                    (size_t)sl.L * (size_t)sl.q_count * sizeof(double) >= e->fuse_threshold &&
                    (!a.two_spin || do_kb || a.n_b == 0);
  int *d_flag = nullptr;
  KPlan kp{}, kpb{};
  e->last_fuse_attempted = fuse;
  // The Coulomb kernels are HBM-bound, the exchange kernels tensor-bound: with both to do they
  // run on two streams.  ev_fork orders the Coulomb stream after everything the main stream
  // held when this build began (operands of a device build, the previous build's consumers).
  const bool overlap = e->overlap_j && do_j && have && (do_ka || do_kb);
  cudaStream_t js = overlap ? e->j_stream : e->stream;
  if (overlap) CUDA_CHECK(cudaEventRecord(e->ev_fork, e->stream));
  // ---- K (alpha / closed shell); K (beta) reuses the scratch after alpha has been finalized
  if (do_ka && have) {
    KPlan probe = plan_k(n, a.n_a, sl.q_count, e->workspace_limit, e->sm_count, a.rank2 ? a.n_a : 0);
    e->d_kpart.ensure(probe.kpart_elems * sizeof(double));
    if (fuse) e->d_gpart_a.ensure((size_t)probe.gamma_stride * sl.q_count * sizeof(double));
    if (a.rank2) {
      // one half-transform of the stacked [X | C], then the SYR2K-form accumulation
      const int o16 = (a.n_a + 15) / 16 * 16;
      e->d_stack.ensure((size_t)n * 2 * o16 * sizeof(double));
      launch_stack_factors(d_ca, lda, d_cb, ldb, n, a.n_a, e->d_stack.d(), e->stream);
      e->launches += 1;
      run_k(e, sl, e->d_stack.d(), n, 2 * o16, e->d_kpart.d(), nullptr, kp, nullptr, a.n_a);
    } else {
      run_k(e, sl, d_ca, lda, a.n_a, e->d_kpart.d(), fuse ? e->d_gpart_a.d() : nullptr, kp,
            (overlap && fuse && !do_kb) ? e->ev_k1_done : nullptr);
    }
  }
  if (do_kb) {
    if (have) {
      if (do_ka) {
        e->phase_begin(T_FINAL);
        launch_finalize_jk(nullptr, 0, e->d_kpart.d(), kp.n_splits, kp.ktile, n, kfac, nullptr, d_ka, e->stream,
                           nullptr, 0.0, 0.0, nullptr, kp.n_splits_diag, 1, 0, 0, 0, kp.n_splits_edge);
        e->phase_end(T_FINAL);
        e->launches += 1;
        ka_finalized = true;
      }
      KPlan probe = plan_k(n, a.n_b, sl.q_count, e->workspace_limit, e->sm_count);
      e->d_kpart.ensure(probe.kpart_elems * sizeof(double));  // stream-ordered: the finalize above has consumed it
      if (fuse) e->d_gpart_b.ensure((size_t)probe.gamma_stride * sl.q_count * sizeof(double));
      run_k(e, sl, d_cb, ldb, a.n_b, e->d_kpart.d(), fuse ? e->d_gpart_b.d() : nullptr, kpb,
            (overlap && fuse) ? e->ev_k1_done : nullptr);
      e->phase_begin(T_FINAL);
      launch_finalize_jk(nullptr, 0, e->d_kpart.d(), kpb.n_splits, kpb.ktile, n, kfac, nullptr, d_kb, e->stream,
                         nullptr, 0.0, 0.0, nullptr, kpb.n_splits_diag, 1, 0, 0, 0, kpb.n_splits_edge);
      e->phase_end(T_FINAL);
      e->launches += 1;
    } else {
      CUDA_CHECK(cudaMemsetAsync(d_kb, 0, nn * sizeof(double), e->stream));
    }
  }

  // ---- H and D arrive now (they were not needed by the exchange kernels launched above)
  late_upload();

  // ---- J
  if (do_j && have) {
    if (overlap) {
      CUDA_CHECK(cudaStreamWaitEvent(js, e->ev_fork, 0));
      if (late_issued) CUDA_CHECK(cudaStreamWaitEvent(js, e->ev_late_upload, 0));
    }
    // beside the exchange kernels pass 2 runs as the co-resident TMA-fed kernel; alone, as the wide one
    // (same auxiliary slices, same order of the sums inside a slice: the two kernels give the same bits)
    const bool j_beside_k = overlap && e->coresident_j;
    e->d_w.ensure((size_t)sl.L * sizeof(double));
    e->d_gamma_partial.ensure(jp.gamma_partial_elems * sizeof(double));
    e->d_gamma.ensure((size_t)sl.q_count * sizeof(double));
    e->d_jpart.ensure(jp.j_partial_elems * sizeof(double));
    e->phase_begin(T_J1, js);
    if (fuse) {
      unsigned long long *scratch = reinterpret_cast<unsigned long long *>(static_cast<char *>(e->d_scalar.ptr) + 64);
      d_flag = reinterpret_cast<int *>(static_cast<char *>(e->d_scalar.ptr) + 32);
      launch_density_prep(d_density, n, e->d_w.d(), d_ca, lda, a.n_a, do_kb ? d_cb : nullptr, ldb, do_kb ? a.n_b : 0,
                          kfac, scratch, d_flag, js);
    } else {
      launch_pack_density(d_density, n, e->d_w.d(), nullptr, js);
    }
    launch_j_gamma(sl.packed.d(), sl.L, sl.q_count, e->d_w.d(), jp, e->d_gamma_partial.d(), e->d_gamma.d(), d_flag, js);
    e->launches += 3;
    if (fuse) {
      if (overlap) CUDA_CHECK(cudaStreamWaitEvent(js, e->ev_k1_done, 0));
      launch_gamma_from_x(e->d_gpart_a.d(), kp.gamma_stride, do_kb ? e->d_gpart_b.d() : nullptr, kpb.gamma_stride, kfac,
                          sl.q_count, d_flag, e->d_gamma.d(), js);
      e->launches += 1;
    }
    e->phase_end(T_J1, js);
    e->phase_begin(T_J2, js);
    if (j_beside_k)
      launch_j_accumulate_tma(sl.packed.d(), sl.L, sl.q_count, e->d_gamma.d(), jp.n_slices, e->sm_count, e->d_jpart.d(), js);
    else
      launch_j_accumulate(sl.packed.d(), sl.L, sl.q_count, e->d_gamma.d(), jp, e->d_jpart.d(), js);
    e->phase_end(T_J2, js);
    e->launches += 1;
    if (overlap) {
      CUDA_CHECK(cudaEventRecord(e->ev_j_done, js));
      CUDA_CHECK(cudaStreamWaitEvent(e->stream, e->ev_j_done, 0));
    }
  }

    n_ksplits = kp.n_splits;
    n_ksplits_diag = kp.n_splits_diag;
    n_ksplits_edge = kp.n_splits_edge;
    ktile = kp.ktile ? kp.ktile : 64;
  }

  // ---- fixed-order sums of the partial buffers -> full matrices (and, for a closed-shell
  // single-GPU build_fock, the Fock matrix itself in the same pass)
  const bool will_reduce = sharded && e->n_ranks > 1;
  const bool fused_assemble = a.assemble && !a.two_spin && !will_reduce && have;
  // host-operand builds put their results in one block: [8 scalars | F_a | F_b]
  double *d_fa_host = nullptr, *d_fb_host = nullptr;
  if (!a.device_operands && (a.assemble || a.combine)) {
    e->d_out.ensure((8 + 2 * nn) * sizeof(double));
    d_fa_host = e->d_out.d() + 8;
    d_fb_host = d_fa_host + nn;
  }
  double *d_fa_early = nullptr;
  if (fused_assemble) d_fa_early = a.device_operands ? a.fock_a : d_fa_host;
  e->phase_begin(T_FINAL);
  if (have) {
    const bool fin_k = do_ka && !ka_finalized;
    if (do_j || fin_k || fused_assemble) {
      launch_finalize_jk(do_j ? e->d_jpart.d() : nullptr, n_jslices, fin_k ? e->d_kpart.d() : nullptr, n_ksplits,
                         ktile, n, kfac, do_j ? d_j : nullptr, fin_k ? d_ka : nullptr, e->stream,
                         fused_assemble ? d_h : nullptr, a.j_scale, 0.5 * a.k_scale, d_fa_early, n_ksplits_diag, 1, 0, 0, 0,
                         n_ksplits_edge);
      e->launches += 1;
    }
  } else {
    if (do_j) CUDA_CHECK(cudaMemsetAsync(d_j, 0, nn * sizeof(double), e->stream));
    if (do_ka) CUDA_CHECK(cudaMemsetAsync(d_ka, 0, nn * sizeof(double), e->stream));
  }
  e->phase_end(T_FINAL);
  CUDA_CHECK(cudaGetLastError());

  // ---- multi-GPU: one sum exchange over whatever was built
  bool p2p_used = false;
  if (sharded && e->n_ranks > 1) {
    e->phase_begin(T_ALLREDUCE);
    // J, K_a, K_b are adjacent; reduce the smallest contiguous span that covers the built ones
    size_t first = do_j ? 0 : (do_ka ? 1 : 2);
    size_t last = do_kb ? 3 : (do_ka ? 2 : 1);
    if (!do_j && !do_ka && !do_kb) { first = 0; last = 0; }
    if (last > first) {
      if (!do_ka && do_kb && do_j) CUDA_CHECK(cudaMemsetAsync(d_ka, 0, nn * sizeof(double), e->stream));
      if (use_p2p) {
        // one kernel: wait for the peers, sum my slice over all ranks in rank order, push it everywhere
        Engine::P2P &x = e->p2p;
        ++x.epoch;
        launch_xgpu_allreduce(x.peers, e->n_ranks, e->rank, x.epoch, first * nn, (last - first) * nn,
                              static_cast<unsigned int *>(x.misc), x.dev_error, x.timeout_ns, e->stream);
        e->launches += 1;
        d_j = static_cast<double *>(x.out);
        d_ka = d_j + nn;
        d_kb = d_j + 2 * nn;
        p2p_used = true;
      } else {
        NCCL_CHECK(g_nccl.AllReduce(d_j + first * nn, d_j + first * nn, (last - first) * nn, kNcclFloat64, kNcclSum,
                                    e->comm, e->stream));
      }
    }
    e->phase_end(T_ALLREDUCE);
  }

  // ---- assemble and return
  e->have_last_fock = false;
  struct Pending { double *dst; const double *src; size_t bytes; };
  Pending pending[4];
  int n_pending = 0;
  bool energy_pending = false;
  if (a.assemble) {
    e->phase_begin(T_FINAL);
    const double kf_a = a.two_spin ? a.k_scale : 0.5 * a.k_scale;
    double *d_fa = a.device_operands ? a.fock_a : d_fa_host;
    if (!fused_assemble) {
      launch_assemble_fock(d_h, do_j ? d_j : nullptr, do_ka ? d_ka : nullptr, a.j_scale, kf_a, n, d_fa, e->stream);
      e->launches += 1;
    }
    double *d_fb = nullptr;
    if (a.two_spin && a.fock_b) {
      d_fb = a.device_operands ? a.fock_b : d_fb_host;
      launch_assemble_fock(d_h, do_j ? d_j : nullptr, do_kb ? d_kb : nullptr, a.j_scale, a.k_scale, n, d_fb, e->stream);
      e->launches += 1;
    }
    e->phase_end(T_FINAL);
    if (!a.device_operands) {
      const bool with_energy = !a.two_spin && d_h && d_density;
      if (with_energy) {
        // the energy the caller takes next (assemble_fock, rhf.f90:1207) rides along: one tiny
        // kernel and 8 bytes in the same stream instead of a second round trip
        launch_energy(d_density, d_h, d_fa, n, e->d_escratch.d(), e->d_out.d(), e->stream);
        e->launches += 1;
        e->last_n = n;
        e->have_last_fock = true;
      }
      e->phase_begin(T_DOWNLOAD);
      if (small_io) {
        // [scalars | F_a | F_b] in one DMA to pinned memory, handed to the caller after the sync
        const size_t out_elems = 8 + nn * (d_fb ? 2 : 1);
        e->h_out.ensure(out_elems * sizeof(double));
        CUDA_CHECK(cudaMemcpyAsync(e->h_out.ptr, e->d_out.ptr, out_elems * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
        pending[n_pending++] = {a.fock_a, e->h_out.d() + 8, nn * sizeof(double)};
        if (d_fb) pending[n_pending++] = {a.fock_b, e->h_out.d() + 8 + nn, nn * sizeof(double)};
        energy_pending = with_energy;
      } else {
        CUDA_CHECK(cudaMemcpyAsync(a.fock_a, d_fa, nn * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
        if (d_fb) CUDA_CHECK(cudaMemcpyAsync(a.fock_b, d_fb, nn * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
        if (with_energy)
          CUDA_CHECK(cudaMemcpyAsync(&e->last_energy_host, e->d_out.ptr, sizeof(double), cudaMemcpyDeviceToHost, e->stream));
      }
      e->phase_end(T_DOWNLOAD);
    }
  } else if (a.combine) {
    e->phase_begin(T_FINAL);
    launch_combine_g(do_j ? d_j : nullptr, do_ka ? d_ka : nullptr, do_kb ? d_kb : nullptr, a.ka_coef, a.kb_coef, n,
                     d_fa_host, e->stream);
    e->launches += 1;
    e->phase_end(T_FINAL);
    e->phase_begin(T_DOWNLOAD);
    CUDA_CHECK(cudaMemcpyAsync(a.fock_a, d_fa_host, nn * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    e->phase_end(T_DOWNLOAD);
  } else if (!a.device_operands) {
    e->phase_begin(T_DOWNLOAD);
    const bool get_j = a.j && do_j, get_ka = a.k_a && do_ka, get_kb = a.k_b && do_kb;
    if (small_io && (get_j || get_ka || get_kb)) {
      // J, K_a, K_b are adjacent on the device: one DMA over the span that is wanted
      const size_t first = get_j ? 0 : (get_ka ? 1 : 2), last = get_kb ? 3 : (get_ka ? 2 : 1);
      e->h_out.ensure((last - first) * nn * sizeof(double));
      CUDA_CHECK(cudaMemcpyAsync(e->h_out.ptr, d_j + first * nn, (last - first) * nn * sizeof(double),
                                 cudaMemcpyDeviceToHost, e->stream));
      if (get_j) pending[n_pending++] = {a.j, e->h_out.d(), nn * sizeof(double)};
      if (get_ka) pending[n_pending++] = {a.k_a, e->h_out.d() + (1 - first) * nn, nn * sizeof(double)};
      if (get_kb) pending[n_pending++] = {a.k_b, e->h_out.d() + (2 - first) * nn, nn * sizeof(double)};
    } else {
      if (get_j) CUDA_CHECK(cudaMemcpyAsync(a.j, d_j, nn * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
      if (get_ka) CUDA_CHECK(cudaMemcpyAsync(a.k_a, d_ka, nn * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
      if (get_kb) CUDA_CHECK(cudaMemcpyAsync(a.k_b, d_kb, nn * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    }
    // K of an empty closed-shell / alpha channel is zero, not "untouched": only the two-spin
    // beta channel has skip semantics (mqc_cuest_integrals.f90:1694-1701)
    e->phase_end(T_DOWNLOAD);
  }
  CUDA_CHECK(cudaGetLastError());
  if (a.sync) {
    CUDA_CHECK(cudaStreamSynchronize(e->stream));
    if (p2p_used) p2p_raise_if_failed(e);
    for (int i = 0; i < n_pending; ++i) std::memcpy(pending[i].dst, pending[i].src, pending[i].bytes);
    if (energy_pending) e->last_energy_host = e->h_out.d()[0];
    if (!a.device_operands && !a.assemble && a.k_a && a.want_k && !do_ka) std::memset(a.k_a, 0, nn * sizeof(double));
  }
}

// ------------------------------- device-resident SCF of a fragment -------------------------
// run_libcint_rhf (mqc_libcint_rhf.f90:321-680) minus the integrals: H, S and the resident fitted
// tensor in, the converged closed-shell SCF out, with every matrix staying on the GPU between
// iterations.  Per iteration the host queues six kernels (pack D, pack C, one-pass J/K,
// finalize+assemble F, energy, SCF step) and reads back a handful of scalars.
struct ScfArgs {
  int slot = 0;
  const double *h = nullptr, *s = nullptr;
  int n_electrons = 0, guess = 1, max_iter = 100, diis_vectors = 8;
  double energy_tol = 1e-10, density_tol = 1e-8, k_scale = 1.0;
  double *e_electronic = nullptr;
  int *iterations = nullptr, *converged = nullptr, *n_mo = nullptr;
  double *coeff = nullptr, *eps = nullptr, *density = nullptr, *e_history = nullptr;
};

static void fragment_fock_device(Engine *e, const TensorSlot &sl, const double *d_h, const double *d_density,
                                 const double *d_coeff, int n_occ, double k_scale, double *d_j, double *d_k,
                                 double *d_fock, double *d_energy) {
  const int n = sl.n;
  const bool wk = k_scale != 0.0 && n_occ > 0;
  FragPlan fp = plan_fragment(n, std::max(n_occ, 1), sl.q_count, e->sm_count);
  e->d_w.ensure((size_t)sl.L * sizeof(double));
  e->d_jpart.ensure(fp.jpart_elems * sizeof(double));
  launch_pack_density(d_density, n, e->d_w.d(), nullptr, e->stream);
  if (wk) {
    e->d_ctf.ensure((size_t)fp.nt * fp.nib * 128 * sizeof(double));
    e->d_kpart.ensure(fp.kpart_elems * sizeof(double));
    launch_pack_coeff(d_coeff, n, n, n_occ, fp.nib, e->d_ctf.d(), nullptr, e->stream);
  }
  launch_fragment_jk(sl.packed.d(), sl.q_count, e->d_w.d(), e->d_ctf.d(), fp, true, wk, e->d_jpart.d(), e->d_kpart.d(),
                     e->stream);
  launch_finalize_jk(e->d_jpart.d(), fp.grid, wk ? e->d_kpart.d() : nullptr, fp.grid, 64, n, 2.0, d_j, wk ? d_k : nullptr,
                     e->stream, d_h, 1.0, 0.5 * k_scale, d_fock, -1);
  launch_energy(d_density, d_h, d_fock, n, e->d_escratch.d(), d_energy, e->stream);
  e->launches += wk ? 5 : 4;
}

static void scf_fragment(Engine *e, const ScfArgs &a) {
  if (a.slot < 0 || a.slot >= MQCB200_NUM_SLOTS) throw Failure("mqcb200: tensor slot out of range");
  TensorSlot &sl = e->slots[a.slot];
  if (!sl.set) throw Failure("mqcb200: no fitted tensor has been set on this slot (call mqcb200_set_tensor first)");
  if (!a.h || !a.s || !a.e_electronic || !a.iterations || !a.converged)
    throw Failure("mqcb200: null argument to scf_fragment");
  if (e->comm && e->n_ranks > 1) throw Failure("mqcb200: the device-resident SCF runs whole fragments on one GPU, not a sharded tensor");
  if (sl.q_count != sl.naux_total) throw Failure("mqcb200: the device-resident SCF needs the whole tensor on this GPU");
  const int n = sl.n;
  if (a.n_electrons < 0 || (a.n_electrons & 1)) throw Failure("mqcb200: closed-shell SCF needs an even, non-negative electron count");
  const int n_occ = a.n_electrons / 2;
  if (!scf_path_applies(n) || !fragment_path_applies(n, std::max(n_occ, 1)))
    throw Failure("mqcb200: the device-resident SCF step covers fragment-sized problems (n <= 80, n_occ <= 64)");
  if (a.diis_vectors < 0 || a.diis_vectors > 8) throw Failure("mqcb200: diis_vectors must lie in 0..8");
  if (a.max_iter < 1) throw Failure("mqcb200: max_iter must be positive");
  e->bind();
  e->launches = 0;
  const size_t nn = (size_t)n * n;
  const int dmax = a.diis_vectors;
  // layout of the device block (doubles)
  size_t off = 0;
  auto take = [&](size_t count) { const size_t o = off; off += (count + 1) & ~(size_t)1; return o; };
  const size_t o_h = take(nn), o_s = take(nn), o_x = take(nn), o_f = take(nn), o_d = take(nn), o_c = take(nn);
  const size_t o_j = take(nn), o_k = take(nn), o_eps = take(n), o_work = take(4 * nn);
  const size_t o_df = take((size_t)std::max(dmax, 1) * nn), o_de = take((size_t)std::max(dmax, 1) * nn), o_db = take(64);
  const size_t o_scal = take(24), o_state = take(4 /* ints, in 4 doubles */), o_nmo = take(2);
  e->d_scf.ensure(off * sizeof(double));
  e->h_scf.ensure(16 * sizeof(double));
  double *base = e->d_scf.d();
  int *d_state = reinterpret_cast<int *>(base + o_state), *d_nmo = reinterpret_cast<int *>(base + o_nmo);
  // H and S up (one staged DMA when they fit the small-operand buffer), everything else cleared
  const size_t hs_span = (o_s - o_h) + nn;               // S starts on an even offset: one pad double for odd n*n
  e->h_in.ensure((hs_span + 1) * sizeof(double));
  std::memset(e->h_in.ptr, 0, hs_span * sizeof(double));
  std::memcpy(e->h_in.d(), a.h, nn * sizeof(double));
  std::memcpy(e->h_in.d() + (o_s - o_h), a.s, nn * sizeof(double));
  CUDA_CHECK(cudaMemcpyAsync(base + o_h, e->h_in.ptr, hs_span * sizeof(double), cudaMemcpyHostToDevice, e->stream));
  CUDA_CHECK(cudaMemsetAsync(base + o_x, 0, (off - o_x) * sizeof(double), e->stream));
  launch_scf_orthogonalizer(base + o_s, n, base + o_x, d_nmo, e->stream);
  e->launches += 1;
  int n_mo = 0;
  CUDA_CHECK(cudaMemcpyAsync(&n_mo, d_nmo, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
  CUDA_CHECK(cudaStreamSynchronize(e->stream));
  if (n_mo <= 0) throw Failure("SCF: overlap matrix is singular");                           // mqc_scf_common.f90:69
  if (n_occ > n_mo)
    throw Failure("RHF: more occupied orbitals than the basis supports after near-null modes were dropped (" +
                  std::to_string(n_occ) + " occupied, " + std::to_string(n_mo) + " of " + std::to_string(n) +
                  " orbitals kept)");                                                          // rhf.f90:510-513

  ScfStepLaunch st{};
  st.n = n; st.n_mo = n_mo; st.n_occ = n_occ; st.diis_max = dmax; st.guess = a.guess;
  st.h = base + o_h; st.s = base + o_s; st.x = base + o_x; st.fock = base + o_f; st.density = base + o_d;
  st.coeff = base + o_c; st.eps = base + o_eps; st.work = base + o_work; st.diis_f = base + o_df; st.diis_e = base + o_de;
  st.diis_b = base + o_db; st.state = d_state; st.scalars = base + o_scal;
  st.energy_tol = a.energy_tol; st.density_tol = a.density_tol;
  st.mode = 0;
  launch_scf_step(st, e->stream);                      // guess Fock -> C, D
  e->launches += 1;
  st.mode = 1;

  double *hs = e->h_scf.d();                            // [0..7] scalars, [8..9] state as 4 ints
  int iterations = 0, converged = 0;
  const int every = std::max(1, e->scf_check_every);
  while (iterations < a.max_iter && !converged) {
    const int batch = std::min(every, a.max_iter - iterations);
    for (int b = 0; b < batch; ++b) {
      fragment_fock_device(e, sl, st.h, st.density, st.coeff, n_occ, a.k_scale, base + o_j, base + o_k, st.fock, st.scalars);
      launch_scf_step(st, e->stream);
      e->launches += 1;
      if (a.e_history && batch > 1) {
        // queued iterations: their energies are fetched one by one, in stream order
        CUDA_CHECK(cudaMemcpyAsync(a.e_history + iterations + b, st.scalars + 1, sizeof(double), cudaMemcpyDeviceToHost, e->stream));
      }
    }
    CUDA_CHECK(cudaMemcpyAsync(hs, st.scalars, 8 * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    CUDA_CHECK(cudaMemcpyAsync(hs + 8, d_state, 4 * sizeof(int), cudaMemcpyDeviceToHost, e->stream));
    CUDA_CHECK(cudaStreamSynchronize(e->stream));
    const int *hstate = reinterpret_cast<const int *>(hs + 8);
    if (a.e_history && batch == 1) a.e_history[iterations] = hs[1];
    if (getenv("MQCB200_SCF_TRACE")) {
      double tr[24];
      CUDA_CHECK(cudaMemcpy(tr, st.scalars, sizeof(tr), cudaMemcpyDeviceToHost));
      std::fprintf(stderr, "scf step: sweeps %.0f  clocks: commutator+diis %.0f  F' gemms %.0f  jacobi %.0f  back+density %.0f\n",
                   tr[5], tr[8], tr[9], tr[10], tr[11]);
    }
    iterations = hstate[2];
    converged = hstate[3];
  }
  // the energy that goes out belongs to the density that satisfied the test: one more full build (rhf.f90:646-649)
  fragment_fock_device(e, sl, st.h, st.density, st.coeff, n_occ, a.k_scale, base + o_j, base + o_k, st.fock, st.scalars);
  CUDA_CHECK(cudaMemcpyAsync(hs, st.scalars, sizeof(double), cudaMemcpyDeviceToHost, e->stream));
  if (a.coeff) CUDA_CHECK(cudaMemcpyAsync(a.coeff, st.coeff, (size_t)n * n_mo * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
  if (a.eps) CUDA_CHECK(cudaMemcpyAsync(a.eps, st.eps, (size_t)n_mo * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
  if (a.density) CUDA_CHECK(cudaMemcpyAsync(a.density, st.density, nn * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
  CUDA_CHECK(cudaGetLastError());
  CUDA_CHECK(cudaStreamSynchronize(e->stream));
  *a.e_electronic = hs[0];
  *a.iterations = iterations;
  *a.converged = converged;
  if (a.n_mo) *a.n_mo = n_mo;
}

// ------------------------------- device-resident SCF, any size ------------------------------
// The same loop (run_libcint_rhf, rhf.f90:566-649) for problems beyond the one-CTA step: every
// matrix stays on the GPU, the Fock build is the general path above (two streams, fused Coulomb
// vector, the cross-GPU exchange when the tensor is sharded -- every rank then runs the same,
// bit-identical step), the dense algebra is the batched DMMA GEMM, and LAPACK's dsyev is replaced by
// a one-sided Jacobi iteration on the (shifted, hence positive definite) F' = Y^T F Y with Y the
// previous orbitals: nearly diagonal, three or four sweeps.  Per iteration the host sees the DIIS
// overlaps of the newest error vector, the Jacobi "rotated anything?" flags and three scalars.
bool diis_solve_host(const double *overlap /*[8][8] slot coords*/, int newest, int n_stored, int dmax, double *coef,
                            int *slots) {
  // diis_coefficients + solve_diis (src/methods/mqc_diis.f90:146-273): ages oldest -> newest
  if (n_stored < 2) return false;
  const int nb = n_stored + 1;
  double aug[9][10];
  double scale = 0.0;
  for (int i = 0; i < n_stored; ++i) slots[i] = ((newest - n_stored + i) % dmax + dmax) % dmax;
  for (int i = 0; i < n_stored; ++i)
    for (int j = 0; j < n_stored; ++j) {
      aug[i][j] = overlap[slots[i] * 8 + slots[j]];
      scale = std::max(scale, std::fabs(aug[i][j]));
    }
  for (int i = 0; i < n_stored; ++i)
    for (int j = 0; j < n_stored; ++j)
      if (scale > 0.0) aug[i][j] /= scale;
  for (int i = 0; i < n_stored; ++i) { aug[i][n_stored] = -1.0; aug[n_stored][i] = -1.0; aug[i][nb] = 0.0; }
  aug[n_stored][n_stored] = 0.0;
  aug[n_stored][nb] = -1.0;
  for (int i = 0; i < nb; ++i) {
    int piv = i;
    for (int j = i + 1; j < nb; ++j)
      if (std::fabs(aug[j][i]) > std::fabs(aug[piv][i])) piv = j;
    if (piv != i)
      for (int c = 0; c <= nb; ++c) std::swap(aug[i][c], aug[piv][c]);
    const double pivot = aug[i][i];
    if (std::fabs(pivot) < 1.0e-14) return false;                  // PIVOT_FLOOR
    for (int j = i + 1; j < nb; ++j) {
      const double factor = aug[j][i] / pivot;
      for (int c = i; c <= nb; ++c) aug[j][c] -= factor * aug[i][c];
    }
  }
  double sol[9];
  for (int i = nb - 1; i >= 0; --i) {
    double sum = 0.0;
    for (int c = i + 1; c < nb; ++c) sum += aug[i][c] * sol[c];
    sol[i] = (aug[i][nb] - sum) / aug[i][i];
  }
  for (int i = 0; i < n_stored; ++i) coef[i] = sol[i];
  return true;
}

static void scf_general(Engine *e, const ScfArgs &a) {
  if (a.slot < 0 || a.slot >= MQCB200_NUM_SLOTS) throw Failure("mqcb200: tensor slot out of range");
  TensorSlot &sl = e->slots[a.slot];
  if (!sl.set) throw Failure("mqcb200: no fitted tensor has been set on this slot (call mqcb200_set_tensor first)");
  if (!a.h || !a.s || !a.e_electronic || !a.iterations || !a.converged) throw Failure("mqcb200: null argument to scf");
  if (a.n_electrons < 0 || (a.n_electrons & 1)) throw Failure("mqcb200: closed-shell SCF needs an even, non-negative electron count");
  if (a.diis_vectors < 0 || a.diis_vectors > 8) throw Failure("mqcb200: diis_vectors must lie in 0..8");
  if (a.max_iter < 1) throw Failure("mqcb200: max_iter must be positive");
  const int n = sl.n, n_occ = a.n_electrons / 2, dmax = a.diis_vectors;
  const size_t nn = (size_t)n * n;
  e->bind();
  cudaStream_t st = e->stream;
  int total_launches = 0;
  DevBuf blk;
  size_t off = 0;
  auto take = [&](size_t count) { const size_t o = off; off += (count + 1) & ~(size_t)1; return o; };
  const size_t o_h = take(nn), o_s = take(nn), o_x = take(nn), o_f = take(nn), o_d = take(nn), o_c = take(nn), o_cn = take(nn);
  const size_t o_g = take(nn), o_v = take(nn), o_w0 = take(nn), o_w1 = take(nn), o_lam = take(n), o_eps = take(n);
  const size_t o_df = take((size_t)std::max(dmax, 1) * nn), o_de = take((size_t)std::max(dmax, 1) * nn);
  const size_t o_bm = take(16), o_coef = take(16), o_scal = take(16), o_scr = take(130);
  const size_t o_order = take((size_t)n / 2 + 2), o_slots = take(8), o_flags = take(4);
  try {
    blk.ensure(off * sizeof(double));
    double *b = blk.d();
    double *d_h = b + o_h, *d_s = b + o_s, *d_x = b + o_x, *d_f = b + o_f, *d_d = b + o_d, *d_c = b + o_c, *d_cn = b + o_cn;
    double *d_g = b + o_g, *d_v = b + o_v, *d_w0 = b + o_w0, *d_w1 = b + o_w1, *d_lam = b + o_lam, *d_eps = b + o_eps;
    double *d_df = b + o_df, *d_de = b + o_de, *d_bm = b + o_bm, *d_coef = b + o_coef, *d_scal = b + o_scal, *d_scr = b + o_scr;
    int *d_order = reinterpret_cast<int *>(b + o_order), *d_slots = reinterpret_cast<int *>(b + o_slots);
    int *d_flag = reinterpret_cast<int *>(b + o_flags), *d_ndrop = d_flag + 2;   // d_flag[0..1]: the eigensolver's state
    CUDA_CHECK(cudaMemsetAsync(b, 0, off * sizeof(double), st));
    CUDA_CHECK(cudaMemcpyAsync(d_h, a.h, nn * sizeof(double), cudaMemcpyHostToDevice, st));
    CUDA_CHECK(cudaMemcpyAsync(d_s, a.s, nn * sizeof(double), cudaMemcpyHostToDevice, st));
    auto gemm = [&](int M, int N, int K, double alpha, const double *A, int lda, bool ta, const double *B, int ldb, bool tb, double *C,
                    int ldc) {
      launch_dgemm_batched(M, N, K, alpha, A, lda, 0, ta, B, ldb, 0, tb, 0.0, C, ldc, 0, 1, st);
      ++total_launches;
    };
    // one-sided Jacobi on the columns of the m x m positive definite matrix in d_g; eigenvectors accumulate in d_v
    int last_sweeps = 0;
    auto jacobi = [&](int m) {
      launch_set_identity(d_v, m, st);
      const int m_e = (m + 1) & ~1;
      last_sweeps = 0;
      if (use_cooperative_jacobi() && launch_hestenes_solve(d_g, d_v, m, 60, d_flag, st)) {
        total_launches += 1;                              // no host synchronisation: the flag and the sweeps stay on the device
        return;
      }
      cudaGetLastError();
      for (int sweep = 0; sweep < 60; ++sweep) {
        ++last_sweeps;
        CUDA_CHECK(cudaMemsetAsync(d_flag, 0, sizeof(int), st));
        for (int r = 0; r < m_e - 1; ++r) launch_hestenes_round(d_g, d_v, m, r, d_flag, st);
        total_launches += m_e - 1;
        int rotated = 0;
        CUDA_CHECK(cudaMemcpyAsync(&rotated, d_flag, sizeof(int), cudaMemcpyDeviceToHost, st));
        CUDA_CHECK(cudaStreamSynchronize(st));
        if (!rotated) break;
      }
    };
    // ---- orthogonaliser X = U s^-1/2 over eigenvalues > 1e-7, ascending (mqc_scf_common.f90:42-82)
    CUDA_CHECK(cudaMemcpyAsync(d_g, d_s, nn * sizeof(double), cudaMemcpyDeviceToDevice, st));
    jacobi(n);
    launch_eig_lambda(d_g, d_v, n, nullptr, d_lam, st);
    CUDA_CHECK(cudaMemsetAsync(d_ndrop, 0, sizeof(int), st));
    launch_rank_sort(d_lam, n, 1.0e-7, d_order, d_ndrop, st);
    int dropped = 0;
    CUDA_CHECK(cudaMemcpyAsync(&dropped, d_ndrop, sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
    const int m = n - dropped;
    if (m <= 0) throw Failure("SCF: overlap matrix is singular");
    if (n_occ > m)
      throw Failure("RHF: more occupied orbitals than the basis supports after near-null modes were dropped (" +
                    std::to_string(n_occ) + " occupied, " + std::to_string(m) + " of " + std::to_string(n) + " orbitals kept)");
    launch_gather_columns(d_v, n, d_order, dropped, m, d_lam, true, d_x, st);
    const size_t mm = (size_t)m * m;
    // ---- diagonalize (rhf.f90:1464-1489) in the basis y (n x m, S-orthonormal): C <- y V, eps
    auto diagonalize = [&](const double *f_in, const double *y) {
      gemm(n, m, n, 1.0, f_in, n, false, y, n, false, d_w1, n);             // F y
      gemm(m, m, n, 1.0, y, n, true, d_w1, n, false, d_g, m);               // y^T F y
      launch_reduce(d_g, nullptr, mm, 2, d_scr, d_scal + 8, st);            // shift = |F'|_F  >= every |eigenvalue|
      launch_symmetrize_shift(d_g, m, d_scal + 8, st);
      jacobi(m);
      launch_eig_lambda(d_g, d_v, m, d_scal + 8, d_lam, st);
      launch_rank_sort(d_lam, m, -1.0e300, d_order, nullptr, st);
      launch_gather_columns(d_v, m, d_order, 0, m, nullptr, false, d_w0, st);   // V, columns ascending
      launch_gather_columns(d_lam, 1, d_order, 0, m, nullptr, false, d_eps, st);
      gemm(n, m, m, 1.0, y, n, false, d_w0, m, false, d_cn, n);             // C = y V
      CUDA_CHECK(cudaMemcpyAsync(d_c, d_cn, (size_t)n * m * sizeof(double), cudaMemcpyDeviceToDevice, st));
      total_launches += 6;
    };
    auto new_density = [&]() {                                              // D = 2 C_occ C_occ^T, sum (dD)^2 -> scal[3]
      if (n_occ > 0) gemm(n, n, n_occ, 2.0, d_c, n, false, d_c, n, true, d_w1, n);
      else CUDA_CHECK(cudaMemsetAsync(d_w1, 0, nn * sizeof(double), st));
      launch_reduce(d_w1, d_d, nn, 1, d_scr, d_scal + 3, st);
      CUDA_CHECK(cudaMemcpyAsync(d_d, d_w1, nn * sizeof(double), cudaMemcpyDeviceToDevice, st));
      ++total_launches;
    };
    auto fock_build = [&]() {
      BuildArgs ba;
      ba.slot = a.slot; ba.device_operands = true; ba.h = d_h; ba.density = d_d; ba.coeff_a = d_c; ba.lda = n; ba.n_a = n_occ;
      ba.k_scale = a.k_scale; ba.j_scale = 1.0; ba.fock_a = d_f; ba.assemble = true; ba.sync = false;
      ba.want_j = true; ba.want_k = a.k_scale != 0.0 && n_occ > 0;
      build(e, ba);
      total_launches += e->launches;
      launch_energy(d_d, d_h, d_f, n, e->d_escratch.d(), d_scal, st);
      ++total_launches;
    };
    // ---- guess
    launch_scf_guess(d_h, d_s, n, a.guess == 1, d_f, st);
    diagonalize(d_f, d_x);
    new_density();
    // ---- iterations
    double overlap[64] = {0.0};
    int n_stored = 0, newest = 0;                                           // newest: 1-based slot, as in diis_state_t
    double e_old = 0.0;
    int iterations = 0, converged = 0;
    double hs[4];
    for (int it = 1; it <= a.max_iter && !converged; ++it) {
      fock_build();
      const double *f_use = d_f;
      if (dmax > 0) {
        gemm(n, n, n, 1.0, d_f, n, false, d_d, n, false, d_w0, n);          // F D
        gemm(n, n, n, 1.0, d_w0, n, false, d_s, n, false, d_w1, n);         // F D S
        launch_antisym(d_w1, n, d_w0, st);                                  // F D S - S D F
        gemm(n, m, n, 1.0, d_w0, n, false, d_x, n, false, d_w1, n);         // (..) X
        newest = newest % dmax + 1;                                         // diis_push, mqc_diis.f90:94-119
        if (n_stored < dmax) ++n_stored;
        const int slot = newest - 1;
        gemm(m, m, n, 1.0, d_x, n, true, d_w1, n, false, d_de + (size_t)slot * nn, m);   // e = X^T (..) X
        CUDA_CHECK(cudaMemcpyAsync(d_df + (size_t)slot * nn, d_f, nn * sizeof(double), cudaMemcpyDeviceToDevice, st));
        int others[8];
        for (int age = 0; age < n_stored; ++age) {
          others[age] = ((newest - n_stored + age) % dmax + dmax) % dmax;
          launch_reduce(d_de + (size_t)slot * nn, d_de + (size_t)others[age] * nn, mm, 0, d_scr, d_bm + age, st);
        }
        total_launches += 1 + n_stored;
        double row[8];
        CUDA_CHECK(cudaMemcpyAsync(row, d_bm, n_stored * sizeof(double), cudaMemcpyDeviceToHost, st));
        CUDA_CHECK(cudaStreamSynchronize(st));
        for (int age = 0; age < n_stored; ++age) { overlap[slot * 8 + others[age]] = row[age]; overlap[others[age] * 8 + slot] = row[age]; }
        double coef[8];
        int slots[8];
        if (diis_solve_host(overlap, newest, n_stored, dmax, coef, slots)) {
          CUDA_CHECK(cudaMemcpyAsync(d_coef, coef, n_stored * sizeof(double), cudaMemcpyHostToDevice, st));
          CUDA_CHECK(cudaMemcpyAsync(d_slots, slots, n_stored * sizeof(int), cudaMemcpyHostToDevice, st));
          launch_lincomb(d_df, nn, d_coef, d_slots, n_stored, d_w0, st);    // extrapolated Fock, oldest first
          f_use = d_w0;                                                     // (read once, by the first GEMM of diagonalize)
          ++total_launches;
        }
      }
      diagonalize(f_use, d_c);                                              // F' in the basis of the previous orbitals
      new_density();
      CUDA_CHECK(cudaMemcpyAsync(hs, d_scal, 4 * sizeof(double), cudaMemcpyDeviceToHost, st));
      CUDA_CHECK(cudaStreamSynchronize(st));
      const double e_elec = hs[0], de = std::fabs(e_elec - e_old), drms = std::sqrt(hs[3] / (double)nn);
      if (a.e_history) a.e_history[it - 1] = e_elec;
      e_old = e_elec;
      iterations = it;
      if (it > 1 && de < a.energy_tol && drms < a.density_tol) converged = 1;
    }
    fock_build();                                                           // final rebuild, rhf.f90:646-649
    CUDA_CHECK(cudaMemcpyAsync(hs, d_scal, sizeof(double), cudaMemcpyDeviceToHost, st));
    if (a.coeff) CUDA_CHECK(cudaMemcpyAsync(a.coeff, d_c, (size_t)n * m * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (a.eps) CUDA_CHECK(cudaMemcpyAsync(a.eps, d_eps, (size_t)m * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (a.density) CUDA_CHECK(cudaMemcpyAsync(a.density, d_d, nn * sizeof(double), cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaGetLastError());
    CUDA_CHECK(cudaStreamSynchronize(st));
    if (e->comm && e->n_ranks > 1) p2p_raise_if_failed(e);
    *a.e_electronic = hs[0];
    *a.iterations = iterations;
    *a.converged = converged;
    if (a.n_mo) *a.n_mo = m;
    e->launches = total_launches;
    e->scf_last_sweeps = last_sweeps;
  } catch (...) {
    blk.release();
    throw;
  }
  blk.release();
}

// ------------------------------- a batch of fragment SCFs, in lock-step ----------------------
// MBE fragments of one kind (all trimers of a water cluster, say) have the same (n, naux, n_occ).
// `n_frag` of them are driven together: their tensors sit back to back on the slot (fragment f owns
// auxiliary slabs [f*naux, (f+1)*naux) of a tensor set with naux_total = n_frag*naux), and every
// kernel of an iteration -- pack D, pack C, one-pass J/K, finalize+assemble, energy, SCF step -- is
// launched ONCE with the fragment index on a grid axis.  Six launches per iteration whatever the
// batch size; a fragment that has converged idles through the remaining iterations of its batch.
struct ScfBatchArgs {
  int slot = 0, n_frag = 0;
  const double *h = nullptr, *s = nullptr;            // (n, n, n_frag) each
  int n_electrons = 0, guess = 1, max_iter = 100, diis_vectors = 8;
  double energy_tol = 1e-10, density_tol = 1e-8, k_scale = 1.0;
  double *e_electronic = nullptr;                       // [n_frag]
  int *iterations = nullptr, *converged = nullptr, *n_mo = nullptr;   // [n_frag]
  double *coeff = nullptr, *eps = nullptr, *density = nullptr;        // (n, n, n_frag), (n, n_frag), (n, n, n_frag); nullable
};

static void scf_fragment_batch(Engine *e, const ScfBatchArgs &a) {
  if (a.slot < 0 || a.slot >= MQCB200_NUM_SLOTS) throw Failure("mqcb200: tensor slot out of range");
  TensorSlot &sl = e->slots[a.slot];
  if (!sl.set) throw Failure("mqcb200: no fitted tensor has been set on this slot (call mqcb200_set_tensor first)");
  if (!a.h || !a.s || !a.e_electronic || !a.iterations || !a.converged) throw Failure("mqcb200: null argument to scf_fragment_batch");
  if (e->comm && e->n_ranks > 1) throw Failure("mqcb200: the device-resident SCF runs whole fragments on one GPU, not a sharded tensor");
  if (a.n_frag < 1 || sl.q_count != sl.naux_total || sl.q_count % a.n_frag != 0)
    throw Failure("mqcb200: the slot must hold n_fragments tensors of equal size back to back (naux_total = n_fragments * naux)");
  const int n = sl.n, nf = a.n_frag, naux = sl.q_count / nf;
  if (a.n_electrons < 0 || (a.n_electrons & 1)) throw Failure("mqcb200: closed-shell SCF needs an even, non-negative electron count");
  const int n_occ = a.n_electrons / 2;
  if (!scf_path_applies(n) || !fragment_path_applies(n, std::max(n_occ, 1)))
    throw Failure("mqcb200: the device-resident SCF step covers fragment-sized problems (n <= 80, n_occ <= 64)");
  if (a.diis_vectors < 0 || a.diis_vectors > 8) throw Failure("mqcb200: diis_vectors must lie in 0..8");
  if (a.max_iter < 1) throw Failure("mqcb200: max_iter must be positive");
  e->bind();
  e->launches = 0;
  const size_t nn = (size_t)n * n;
  const int dmax = a.diis_vectors;
  size_t off = 0;
  auto take = [&](size_t count) { const size_t o = off; off += (count + 1) & ~(size_t)1; return o; };
  const size_t o_h = take(nn), o_s = take(nn), o_x = take(nn), o_f = take(nn), o_d = take(nn), o_c = take(nn);
  const size_t o_j = take(nn), o_k = take(nn), o_eps = take(n), o_work = take(4 * nn);
  const size_t o_df = take((size_t)std::max(dmax, 1) * nn), o_de = take((size_t)std::max(dmax, 1) * nn), o_db = take(64);
  const size_t o_scal = take(24), o_state = take(4), o_nmo = take(2);
  const size_t blk = off;                                  // doubles per fragment
  const bool wk = a.k_scale != 0.0 && n_occ > 0;
  // a few CTAs per fragment are enough once there are many fragments
  FragPlan fp = plan_fragment(n, std::max(n_occ, 1), naux, std::max(4, 3 * e->sm_count / nf));
  const size_t ctf_stride = (size_t)fp.nt * fp.nib * 128;
  DevBuf d_esc;
  cudaStream_t st = e->stream;
  std::vector<int> states((size_t)4 * nf, 0), nmos(nf, 0);
  try {
    e->d_scf.ensure(blk * nf * sizeof(double));
    d_esc.ensure((size_t)130 * nf * sizeof(double));
    e->d_w.ensure((size_t)sl.L * nf * sizeof(double));
    e->d_jpart.ensure(fp.jpart_elems * nf * sizeof(double));
    if (wk) {
      e->d_ctf.ensure(ctf_stride * nf * sizeof(double));
      e->d_kpart.ensure(fp.kpart_elems * nf * sizeof(double));
    }
    double *base = e->d_scf.d();
    CUDA_CHECK(cudaMemsetAsync(base, 0, blk * nf * sizeof(double), st));
    CUDA_CHECK(cudaMemsetAsync(d_esc.ptr, 0, (size_t)130 * nf * sizeof(double), st));
    CUDA_CHECK(cudaMemcpy2DAsync(base + o_h, blk * sizeof(double), a.h, nn * sizeof(double), nn * sizeof(double), nf, cudaMemcpyHostToDevice, st));
    CUDA_CHECK(cudaMemcpy2DAsync(base + o_s, blk * sizeof(double), a.s, nn * sizeof(double), nn * sizeof(double), nf, cudaMemcpyHostToDevice, st));
    int *d_state = reinterpret_cast<int *>(base + o_state), *d_nmo = reinterpret_cast<int *>(base + o_nmo);
    launch_scf_orthogonalizer(base + o_s, n, base + o_x, d_nmo, st, nf, blk);
    ScfStepLaunch sp{};
    sp.batch = nf; sp.block_stride = blk; sp.n_mo_dev = d_nmo;
    sp.n = n; sp.n_mo = n; sp.n_occ = n_occ; sp.diis_max = dmax; sp.guess = a.guess;
    sp.h = base + o_h; sp.s = base + o_s; sp.x = base + o_x; sp.fock = base + o_f; sp.density = base + o_d;
    sp.coeff = base + o_c; sp.eps = base + o_eps; sp.work = base + o_work; sp.diis_f = base + o_df; sp.diis_e = base + o_de;
    sp.diis_b = base + o_db; sp.state = d_state; sp.scalars = base + o_scal;
    sp.energy_tol = a.energy_tol; sp.density_tol = a.density_tol;
    sp.mode = 0;
    launch_scf_step(sp, st);
    sp.mode = 1;
    e->launches += 2;
    auto fock_build = [&]() {
      launch_pack_density(base + o_d, n, e->d_w.d(), nullptr, st, nf, blk, (size_t)sl.L);
      if (wk) launch_pack_coeff(base + o_c, n, n, n_occ, fp.nib, e->d_ctf.d(), nullptr, st, nf, blk, ctf_stride);
      launch_fragment_jk(sl.packed.d(), naux, e->d_w.d(), e->d_ctf.d(), fp, true, wk, e->d_jpart.d(), e->d_kpart.d(), st, nf);
      launch_finalize_jk(e->d_jpart.d(), fp.grid, wk ? e->d_kpart.d() : nullptr, fp.grid, 64, n, 2.0, base + o_j,
                         wk ? base + o_k : nullptr, st, base + o_h, 1.0, 0.5 * a.k_scale, base + o_f, -1, nf, fp.jpart_elems,
                         fp.kpart_elems, blk);
      launch_energy(base + o_d, base + o_h, base + o_f, n, d_esc.d(), base + o_scal, st, nf, blk, blk);
      e->launches += wk ? 5 : 4;
    };
    const int every = std::max(1, e->scf_check_every);
    int done_iters = 0;
    bool all_done = false;
    while (done_iters < a.max_iter && !all_done) {
      const int batch_it = std::min(every, a.max_iter - done_iters);
      for (int b = 0; b < batch_it; ++b) {
        fock_build();
        launch_scf_step(sp, st);
        e->launches += 1;
      }
      done_iters += batch_it;
      CUDA_CHECK(cudaMemcpy2DAsync(states.data(), 4 * sizeof(int), d_state, blk * sizeof(double), 4 * sizeof(int), nf,
                                   cudaMemcpyDeviceToHost, st));
      CUDA_CHECK(cudaStreamSynchronize(st));
      all_done = true;
      for (int f = 0; f < nf; ++f) all_done = all_done && states[4 * f + 3] != 0;
    }
    fock_build();                                          // the final rebuild (rhf.f90:646-649), every fragment
    CUDA_CHECK(cudaMemcpy2DAsync(a.e_electronic, sizeof(double), base + o_scal, blk * sizeof(double), sizeof(double), nf,
                                 cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaMemcpy2DAsync(nmos.data(), sizeof(int), d_nmo, blk * sizeof(double), sizeof(int), nf, cudaMemcpyDeviceToHost, st));
    if (a.coeff) CUDA_CHECK(cudaMemcpy2DAsync(a.coeff, nn * sizeof(double), base + o_c, blk * sizeof(double), nn * sizeof(double), nf, cudaMemcpyDeviceToHost, st));
    if (a.eps) CUDA_CHECK(cudaMemcpy2DAsync(a.eps, (size_t)n * sizeof(double), base + o_eps, blk * sizeof(double), (size_t)n * sizeof(double), nf, cudaMemcpyDeviceToHost, st));
    if (a.density) CUDA_CHECK(cudaMemcpy2DAsync(a.density, nn * sizeof(double), base + o_d, blk * sizeof(double), nn * sizeof(double), nf, cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaGetLastError());
    CUDA_CHECK(cudaStreamSynchronize(st));
  } catch (...) {
    d_esc.release();
    throw;
  }
  d_esc.release();
  for (int f = 0; f < nf; ++f) {
    a.iterations[f] = states[4 * f + 2];
    a.converged[f] = states[4 * f + 3];                    // 1 converged, 0 not within max_iter, -1 basis collapsed below n_occ
    if (nmos[f] <= 0 || n_occ > nmos[f]) a.converged[f] = -1;
    if (a.n_mo) a.n_mo[f] = nmos[f];
  }
}

// ------------------------------- DF gradient densities -------------------------------------
// The two densities every term of the density-fitted two-electron gradient is contracted with
// (df_two_electron_gradient + add_exchange_channel, backends/libcint/mqc_libcint_gradient.f90:
// 1545-1812), formed on the device from the RESIDENT whitened tensor B = (mu nu|P) . J^(-1/2) and
// the caller's J^(-1/2) (`half`) -- the reference re-generates (mu nu|P) and the metric for this:
//
//   rho    = J^-1 g = half . gamma,          gamma_R = sum B_R . D              (:1672-1676)
//   e~^R   = C^T B_R C ,   f^P = sum_R half(P,R) e~^R   (= J^-1 e, :1781-1795)
//   Gamma^P = rho_P D - sum_spin w_s C f^{s,P} C^T                               (:1678-1681, :1797-1804)
//   Omega   = -1/2 rho rho^T + sum_spin (w_s / 2) sum_ij f^{s,P}_ij f^{s,Q}_ij   (:1696-1700, :1806-1812)
//
// w = 2 kf for the single closed-shell channel, kf for each unrestricted one (:1706-1716); kf = 0
// skips the whole exchange assembly (:1702-1705); with_coulomb = false drops BOTH Coulomb shapes
// (:1677, :1696-1700).  The derivative integrals stay on the host (libcint).
struct GradArgs {
  int slot = 0;
  const double *half = nullptr, *density = nullptr, *ca = nullptr, *cb = nullptr;
  int lda = 0, ldb = 0, n_a = 0, n_b = 0;
  bool unrestricted = false, with_coulomb = true;
  double kf = 1.0;
  double *gamma = nullptr, *omega = nullptr;
};

static void df_gradient_densities(Engine *e, const GradArgs &a) {
  if (a.slot < 0 || a.slot >= MQCB200_NUM_SLOTS) throw Failure("mqcb200: tensor slot out of range");
  TensorSlot &sl = e->slots[a.slot];
  if (!sl.set) throw Failure("mqcb200: no fitted tensor has been set on this slot (call mqcb200_set_tensor first)");
  if (sl.q_count != sl.naux_total || (e->comm && e->n_ranks > 1))
    throw Failure("mqcb200: the gradient densities need the whole tensor on this GPU");
  if (!a.half || !a.density || !a.gamma || !a.omega) throw Failure("mqcb200: null argument to df_gradient_densities");
  if (a.n_a < 0 || a.n_b < 0 || (a.n_a > 0 && (!a.ca || a.lda < sl.n)) || (a.unrestricted && a.n_b > 0 && (!a.cb || a.ldb < sl.n)))
    throw Failure("mqcb200: bad orbital matrix / leading dimension");
  e->bind();
  e->launches = 0;
  const int n = sl.n, Q = sl.naux_total;
  const size_t nn = (size_t)n * n;
  cudaStream_t st = e->stream;
  struct Channel { const double *c_host; int ld, o; double w; DevBuf c, et, f; };
  Channel ch[2];
  int n_ch = 0;
  if (a.kf != 0.0) {
    if (a.unrestricted) {
      if (a.n_a > 0) { ch[n_ch].c_host = a.ca; ch[n_ch].ld = a.lda; ch[n_ch].o = a.n_a; ch[n_ch].w = a.kf; ++n_ch; }
      if (a.n_b > 0) { ch[n_ch].c_host = a.cb; ch[n_ch].ld = a.ldb; ch[n_ch].o = a.n_b; ch[n_ch].w = a.kf; ++n_ch; }
    } else if (a.n_a > 0) {
      ch[n_ch].c_host = a.ca; ch[n_ch].ld = a.lda; ch[n_ch].o = a.n_a; ch[n_ch].w = 2.0 * a.kf; ++n_ch;
    }
  }
  int o_max = 1;
  for (int k = 0; k < n_ch; ++k) o_max = std::max(o_max, ch[k].o);
  // auxiliary functions per pass: unpacked slabs, half-transformed slabs and the Gamma block share ~1 GiB
  size_t chunk = std::max<size_t>(1, ((size_t)384 << 20) / (nn * sizeof(double)));
  chunk = std::min<size_t>(chunk, (size_t)Q);
  DevBuf d_half, d_dens, d_rho, d_omega, d_u, d_x, d_z;
  auto release_all = [&]() {
    for (DevBuf *b : {&d_half, &d_dens, &d_rho, &d_omega, &d_u, &d_x, &d_z}) b->release();
    for (int k = 0; k < 2; ++k) { ch[k].c.release(); ch[k].et.release(); ch[k].f.release(); }
  };
  try {
    d_half.ensure((size_t)Q * Q * sizeof(double));
    d_dens.ensure(nn * sizeof(double));
    d_rho.ensure((size_t)Q * sizeof(double));
    d_omega.ensure((size_t)Q * Q * sizeof(double));
    d_u.ensure(chunk * nn * sizeof(double));
    d_x.ensure(chunk * (size_t)n * o_max * sizeof(double));
    d_z.ensure(chunk * nn * sizeof(double));
    CUDA_CHECK(cudaMemcpyAsync(d_half.ptr, a.half, (size_t)Q * Q * sizeof(double), cudaMemcpyHostToDevice, st));
    CUDA_CHECK(cudaMemcpyAsync(d_dens.ptr, a.density, nn * sizeof(double), cudaMemcpyHostToDevice, st));
    for (int k = 0; k < n_ch; ++k) {
      const int o = ch[k].o;
      ch[k].c.ensure((size_t)n * o * sizeof(double));
      ch[k].et.ensure((size_t)o * o * Q * sizeof(double));
      ch[k].f.ensure((size_t)o * o * Q * sizeof(double));
      CUDA_CHECK(cudaMemcpy2DAsync(ch[k].c.ptr, (size_t)n * sizeof(double), ch[k].c_host, (size_t)ch[k].ld * sizeof(double),
                                   (size_t)n * sizeof(double), o, cudaMemcpyHostToDevice, st));
    }
    // ---- rho = half . gamma, gamma_R = sum B_R . D  (the Coulomb kernels' first pass)
    if (a.with_coulomb) {
      JPlan jp = plan_j(n, Q);
      e->d_w.ensure((size_t)sl.L * sizeof(double));
      e->d_gamma_partial.ensure(jp.gamma_partial_elems * sizeof(double));
      e->d_gamma.ensure((size_t)Q * sizeof(double));
      launch_pack_density(d_dens.d(), n, e->d_w.d(), nullptr, st);
      launch_j_gamma(sl.packed.d(), sl.L, Q, e->d_w.d(), jp, e->d_gamma_partial.d(), e->d_gamma.d(), nullptr, st);
      launch_dgemm_batched(Q, 1, Q, 1.0, d_half.d(), Q, 0, false, e->d_gamma.d(), Q, 0, false, 0.0, d_rho.d(), Q, 0, 1, st);
      launch_dgemm_batched(Q, Q, 1, -0.5, d_rho.d(), Q, 0, false, d_rho.d(), Q, 0, true, 0.0, d_omega.d(), Q, 0, 1, st);
      e->launches += 5;
    } else {
      CUDA_CHECK(cudaMemsetAsync(d_rho.ptr, 0, (size_t)Q * sizeof(double), st));
      CUDA_CHECK(cudaMemsetAsync(d_omega.ptr, 0, (size_t)Q * Q * sizeof(double), st));
    }
    // ---- per channel: e~^R = C^T B_R C for every R, f = e~ . half, Omega += (w/2) f^T f
    for (int k = 0; k < n_ch; ++k) {
      const int o = ch[k].o;
      for (size_t r0 = 0; r0 < (size_t)Q; r0 += chunk) {
        const int rc = (int)std::min(chunk, (size_t)Q - r0);
        launch_unpack_tensor(sl.packed.d() + r0 * (size_t)sl.L, n, rc, d_u.d(), st);
        launch_dgemm_batched(n, o, n, 1.0, d_u.d(), n, (long long)nn, false, ch[k].c.d(), n, 0, false, 0.0, d_x.d(), n,
                             (long long)n * o, rc, st);                                     // X_R = B_R C
        launch_dgemm_batched(o, o, n, 1.0, ch[k].c.d(), n, 0, true, d_x.d(), n, (long long)n * o, false, 0.0,
                             ch[k].et.d() + r0 * (size_t)o * o, o, (long long)o * o, rc, st);   // e~^R = C^T X_R
        e->launches += 3;
      }
      launch_dgemm_batched(o * o, Q, Q, 1.0, ch[k].et.d(), (long long)o * o, 0, false, d_half.d(), Q, 0, false, 0.0,
                           ch[k].f.d(), (long long)o * o, 0, 1, st);                          // f = e~ . half
      launch_dgemm_batched(Q, Q, o * o, 0.5 * ch[k].w, ch[k].f.d(), (long long)o * o, 0, true, ch[k].f.d(),
                           (long long)o * o, 0, false, 1.0, d_omega.d(), Q, 0, 1, st);        // Omega += (w/2) f^T f
      e->launches += 2;
    }
    // ---- Gamma, a block of auxiliary functions at a time, straight to the caller's array
    for (size_t p0 = 0; p0 < (size_t)Q; p0 += chunk) {
      const int pc = (int)std::min(chunk, (size_t)Q - p0);
      launch_gradient_gamma(d_rho.d() + p0, d_dens.d(), nullptr, 0.0, n, pc, false, d_u.d(), st);   // rho_P D
      e->launches += 1;
      for (int k = 0; k < n_ch; ++k) {
        const int o = ch[k].o;
        launch_dgemm_batched(n, o, o, 1.0, ch[k].c.d(), n, 0, false, ch[k].f.d() + p0 * (size_t)o * o, o, (long long)o * o,
                             false, 0.0, d_x.d(), n, (long long)n * o, pc, st);               // C f^P
        launch_dgemm_batched(n, n, o, 1.0, d_x.d(), n, (long long)n * o, false, ch[k].c.d(), n, 0, true, 0.0, d_z.d(), n,
                             (long long)nn, pc, st);                                          // Z^P = (C f^P) C^T
        launch_gradient_gamma(nullptr, nullptr, d_z.d(), ch[k].w, n, pc, true, d_u.d(), st);  // Gamma^P -= w Z^P
        e->launches += 3;
      }
      CUDA_CHECK(cudaMemcpyAsync(a.gamma + p0 * nn, d_u.ptr, (size_t)pc * nn * sizeof(double), cudaMemcpyDeviceToHost, st));
      CUDA_CHECK(cudaStreamSynchronize(st));
    }
    CUDA_CHECK(cudaMemcpyAsync(a.omega, d_omega.ptr, (size_t)Q * Q * sizeof(double), cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaGetLastError());
    CUDA_CHECK(cudaStreamSynchronize(st));
  } catch (...) {
    release_all();
    throw;
  }
  release_all();
}

// ------------------------------- building the tensor on the device -------------------------
// metric^(-1/2) = U s^(-1/2) U^T over the modes above `threshold` (metric_inverse_sqrt,
// mqc_libcint_integrals.F90:992-1038) with the eigendecomposition on the GPU: one-sided Jacobi on the
// columns of the symmetric metric (launch_hestenes_round), a sweep that rotates nothing ends it.
// d_half_out (naux x naux, device) receives the result.
static int metric_inverse_sqrt_device(Engine *e, int naux, const double *metric_host, double threshold, double *d_half_out) {
  if (naux <= 0 || !metric_host) throw Failure("mqcb200: bad metric");
  const size_t qq = (size_t)naux * naux;
  DevBuf d_g, d_v, d_s, d_misc;
  cudaEvent_t t0 = nullptr, t1 = nullptr;
  int kept = 0;
  try {
    d_g.ensure(qq * sizeof(double));
    d_v.ensure(qq * sizeof(double));
    d_s.ensure(qq * sizeof(double));
    d_misc.ensure((size_t)naux * sizeof(double) + 64);
    int *d_flag = reinterpret_cast<int *>(static_cast<char *>(d_misc.ptr) + (size_t)naux * sizeof(double));
    int *d_kept = d_flag + 2;                     // d_flag[0..1]: rotation flag and sweep count of the eigensolver
    CUDA_CHECK(cudaEventCreate(&t0));
    CUDA_CHECK(cudaEventCreate(&t1));
    CUDA_CHECK(cudaMemcpyAsync(d_g.ptr, metric_host, qq * sizeof(double), cudaMemcpyHostToDevice, e->stream));
    CUDA_CHECK(cudaEventRecord(t0, e->stream));
    launch_set_identity(d_v.d(), naux, e->stream);
    const int n_e = (naux + 1) & ~1;
    int sweeps = 0;
    if (use_cooperative_jacobi() && launch_hestenes_solve(d_g.d(), d_v.d(), naux, 60, d_flag, e->stream)) {
      int st2[2] = {0, 0};
      CUDA_CHECK(cudaMemcpyAsync(st2, d_flag, 2 * sizeof(int), cudaMemcpyDeviceToHost, e->stream));
      CUDA_CHECK(cudaStreamSynchronize(e->stream));
      sweeps = st2[1];
    } else {
      cudaGetLastError();
      for (int sweep = 0; sweep < 60; ++sweep) {
        ++sweeps;
        CUDA_CHECK(cudaMemsetAsync(d_flag, 0, sizeof(int), e->stream));
        for (int r = 0; r < n_e - 1; ++r) launch_hestenes_round(d_g.d(), d_v.d(), naux, r, d_flag, e->stream);
        int rotated = 0;
        CUDA_CHECK(cudaMemcpyAsync(&rotated, d_flag, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
        CUDA_CHECK(cudaStreamSynchronize(e->stream));
        CUDA_CHECK(cudaGetLastError());
        if (!rotated) break;
      }
    }
    CUDA_CHECK(cudaMemsetAsync(d_kept, 0, sizeof(int), e->stream));
    launch_metric_scale(d_g.d(), d_v.d(), naux, threshold, d_s.d(), static_cast<double *>(d_misc.ptr), d_kept, e->stream);
    // half = scaled . U^T   (integrals.F90:1036)
    launch_dgemm_batched(naux, naux, naux, 1.0, d_s.d(), naux, 0, false, d_v.d(), naux, 0, true, 0.0, d_half_out, naux, 0, 1,
                         e->stream);
    CUDA_CHECK(cudaEventRecord(t1, e->stream));
    CUDA_CHECK(cudaMemcpyAsync(&kept, d_kept, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
    CUDA_CHECK(cudaStreamSynchronize(e->stream));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, t0, t1);
    e->metric_ms = ms;
    e->metric_sweeps = sweeps;
  } catch (...) {
    if (t0) cudaEventDestroy(t0);
    if (t1) cudaEventDestroy(t1);
    d_g.release(); d_v.release(); d_s.release(); d_misc.release();
    throw;
  }
  cudaEventDestroy(t0);
  cudaEventDestroy(t1);
  d_g.release(); d_v.release(); d_s.release(); d_misc.release();
  if (kept == 0) throw Failure("density fitting: the auxiliary metric is singular");      // integrals.F90:1015-1019
  return kept;
}

static void whiten_abort(Engine *e) {
  Engine::Whiten &w = e->wh;
  if (w.multi) {
    for (int k = 0; k < e->n_ranks; ++k)
      if (k != e->rank && w.mapped[k] && !e->p2p.same_process[k]) cudaIpcCloseMemHandle(w.mapped[k]);
  }
  for (auto &m : w.mapped) m = nullptr;
  w.d_af.release(); w.d_slab.release(); w.d_tp.release();
  if (w.t0) cudaEventDestroy(w.t0);
  if (w.t1) cudaEventDestroy(w.t1);
  w.t0 = w.t1 = nullptr;
  w.active = false;
  w.multi = false;
  w.failed = false;
}

// `half` is on the HOST unless half_on_device.  Collective when a communicator is active: every rank
// calls it with its own [q_begin, q_begin + q_count).
static void whiten_begin(Engine *e, int slot, int n, int naux_total, int q_begin, int q_count, const double *half,
                         bool half_on_device) {
  if (slot < 0 || slot >= MQCB200_NUM_SLOTS) throw Failure("mqcb200: tensor slot out of range");
  if (!half) throw Failure("mqcb200: null metric factor");
  if (e->wh.active) throw Failure("mqcb200: a whitening is already in progress on this handle");
  e->bind();
  Engine::Whiten &w = e->wh;
  TensorSlot &sl = e->slots[slot];
  slot_prepare(e, sl, n, naux_total, q_begin, q_count);
  const bool multi = e->comm && e->n_ranks > 1;
  w.slot = slot; w.n = n; w.naux = naux_total; w.multi = multi;
  w.m_rows = multi ? naux_total : q_count;              // a lone rank forms only the rows it keeps
  w.gemm_ms = 0.0; w.flops = 0.0;
  try {
    CUDA_CHECK(cudaEventCreate(&w.t0));
    CUDA_CHECK(cudaEventCreate(&w.t1));
    w.d_af.ensure(std::max<size_t>(16, whiten_half_elems(w.m_rows, naux_total) * sizeof(double)));
    DevBuf d_half;
    const double *dh = half;
    if (!half_on_device) {
      d_half.ensure((size_t)naux_total * naux_total * sizeof(double));
      CUDA_CHECK(cudaMemcpyAsync(d_half.ptr, half, (size_t)naux_total * naux_total * sizeof(double), cudaMemcpyHostToDevice, e->stream));
      dh = d_half.d();
    }
    if (w.m_rows > 0) launch_pack_half(dh, naux_total, multi ? 0 : q_begin, w.m_rows, w.d_af.d(), e->stream);
    CUDA_CHECK(cudaStreamSynchronize(e->stream));
    d_half.release();
    w.dst = WhitenDst{};
    w.dst.ld = sl.L;
    w.dst.col_off = 0;
    if (!multi) {
      w.dst.n_ranks = 1;
      w.dst.base[0] = sl.packed.d();
      w.dst.q_begin[0] = 0;
      w.dst.q_begin[1] = q_count;
    } else {
      if (e->n_ranks > WHITEN_MAX_RANKS) throw Failure("mqcb200: too many ranks for the peer-memory whitening");
      // everyone learns everyone's slab boundaries and packed-tensor address
      DevBuf sb, rb;
      std::vector<int> all(2 * e->n_ranks, 0);
      int mine[2] = {q_begin, q_count};
      sb.ensure(sizeof(mine));
      rb.ensure(sizeof(mine) * e->n_ranks);
      CUDA_CHECK(cudaMemcpyAsync(sb.ptr, mine, sizeof(mine), cudaMemcpyHostToDevice, e->stream));
      NCCL_CHECK(g_nccl.AllGather(sb.ptr, rb.ptr, sizeof(mine), kNcclChar, e->comm, e->stream));
      CUDA_CHECK(cudaMemcpyAsync(all.data(), rb.ptr, sizeof(mine) * e->n_ranks, cudaMemcpyDeviceToHost, e->stream));
      CUDA_CHECK(cudaStreamSynchronize(e->stream));
      sb.release();
      rb.release();
      int expect = 0;
      for (int k = 0; k < e->n_ranks; ++k) {
        if (all[2 * k] != expect) throw Failure("mqcb200: the ranks' auxiliary slabs must be contiguous and in rank order");
        w.dst.q_begin[k] = all[2 * k];
        expect += all[2 * k + 1];
      }
      if (expect != naux_total) throw Failure("mqcb200: the ranks' auxiliary slabs do not cover the auxiliary range");
      w.dst.q_begin[e->n_ranks] = naux_total;
      w.dst.n_ranks = e->n_ranks;
      const bool ok = p2p_exchange(e, sl.packed.ptr, w.mapped);
      if (!ranks_agree(e, ok)) {
        if (!ok) for (auto &m : w.mapped) m = nullptr;
        throw Failure("mqcb200: a rank could not map its peers' packed tensors (peer access is required for the sharded whitening)");
      }
      for (int k = 0; k < e->n_ranks; ++k) w.dst.base[k] = static_cast<double *>(w.mapped[k]);
    }
    w.active = true;
  } catch (...) {
    whiten_abort(e);
    throw;
  }
}

// One nu-slab of (mu nu|P): `three_cols` points at (mu = 0, nu = nu_begin, P = 0) of the host tensor, the
// auxiliary functions ld_aux doubles apart (n*n inside the full three(n*n, naux); n*nu_count for a compact slab).
static void whiten_push(Engine *e, int slot, int nu_begin, int nu_count, const double *three_cols, long long ld_aux) {
  Engine::Whiten &w = e->wh;
  if (!w.active || w.slot != slot) throw Failure("mqcb200: no whitening in progress on this slot (call mqcb200_whiten_begin)");
  const int n = w.n;
  if (!three_cols || nu_begin < 0 || nu_count <= 0 || nu_begin + nu_count > n || (nu_begin % TILE) != 0 ||
      ((nu_begin + nu_count) % TILE != 0 && nu_begin + nu_count != n) || ld_aux < (long long)n * nu_count)
    throw Failure("mqcb200: a slab must cover whole 16-wide column tiles [nu_begin, nu_begin + nu_count) of the orbital index");
  e->bind();
  TensorSlot &sl = e->slots[slot];
  const int nt = num_tiles(n);
  const int tc0 = nu_begin / TILE, tc1 = (nu_begin + nu_count + TILE - 1) / TILE;
  const long long tile0 = tile_index(tc0, tc0, nt);
  const long long tile1 = tc1 >= nt ? num_lower_tiles(nt) : tile_index(tc1, tc1, nt);
  const long long l_slab = (tile1 - tile0) * TILE_ELEMS;
  // Only the rows the packed layout reads cross PCIe (mu >= nu_begin: on average half of the slab), gathered by a
  // few host threads into two pinned buffers so that the gather of chunk k+1 overlaps the DMA of chunk k.
  const size_t rows = (size_t)(n - nu_begin);
  const size_t slab_elems = rows * nu_count;
  w.d_slab.ensure(slab_elems * w.naux * sizeof(double));
  w.d_tp.ensure((size_t)l_slab * w.naux * sizeof(double));
  {
    size_t chunk = std::max<size_t>(1, ((size_t)64 << 20) / (slab_elems * sizeof(double)));
    chunk = std::min(chunk, (size_t)w.naux);
    const int n_buf = chunk < (size_t)w.naux ? 2 : 1;
    for (int i = 0; i < n_buf; ++i) {
      e->h_stage[i].ensure(chunk * slab_elems * sizeof(double));
      if (!e->ev_stage[i]) CUDA_CHECK(cudaEventCreateWithFlags(&e->ev_stage[i], cudaEventDisableTiming));
    }
    size_t k = 0;
    for (size_t p0 = 0; p0 < (size_t)w.naux; p0 += chunk, ++k) {
      const size_t pc = std::min(chunk, (size_t)w.naux - p0);
      const int buf = (int)(k % n_buf);
      if (k >= (size_t)n_buf) CUDA_CHECK(cudaEventSynchronize(e->ev_stage[buf]));
      double *dst = e->h_stage[buf].d();
      auto work = [&](size_t lo, size_t hi) {
        for (size_t p = lo; p < hi; ++p) {
          const double *src = three_cols + (p0 + p) * (size_t)ld_aux + (size_t)nu_begin;   // (mu = nu_begin, nu = nu_begin)
          double *d_p = dst + p * slab_elems;
          for (int c = 0; c < nu_count; ++c) std::memcpy(d_p + (size_t)c * rows, src + (size_t)c * n, rows * sizeof(double));
        }
      };
      const unsigned hw = std::thread::hardware_concurrency();
      const size_t n_threads = pc * slab_elems * sizeof(double) >= ((size_t)16 << 20) ? std::min<size_t>({(size_t)8, (size_t)(hw ? hw : 1), pc}) : 1;
      if (n_threads <= 1) {
        work(0, pc);
      } else {
        std::vector<std::thread> pool;
        for (size_t t = 0; t < n_threads; ++t) pool.emplace_back(work, pc * t / n_threads, pc * (t + 1) / n_threads);
        for (auto &th : pool) th.join();
      }
      CUDA_CHECK(cudaMemcpyAsync(w.d_slab.d() + p0 * slab_elems, dst, pc * slab_elems * sizeof(double), cudaMemcpyHostToDevice, e->stream));
      CUDA_CHECK(cudaEventRecord(e->ev_stage[buf], e->stream));
    }
  }
  launch_pack_slab(w.d_slab.d(), n, w.naux, nu_begin, nu_count, w.d_tp.d(), e->stream);
  if (w.m_rows > 0) {
    WhitenDst d = w.dst;
    d.col_off = tile0 * TILE_ELEMS;
    CUDA_CHECK(cudaEventRecord(w.t0, e->stream));
    launch_whiten_slab(w.d_af.d(), w.m_rows, w.naux, w.d_tp.d(), l_slab, l_slab, d, e->stream);
    CUDA_CHECK(cudaEventRecord(w.t1, e->stream));
  }
  CUDA_CHECK(cudaGetLastError());
  CUDA_CHECK(cudaStreamSynchronize(e->stream));          // the caller may reuse its slab buffer; ours is reused too
  if (w.m_rows > 0) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, w.t0, w.t1);
    w.gemm_ms += ms;
    w.flops += 2.0 * (double)w.m_rows * (double)w.naux * (double)l_slab;
  }
  (void)sl;
}

static void whiten_end(Engine *e, int slot) {
  Engine::Whiten &w = e->wh;
  if (!w.active || w.slot != slot) throw Failure("mqcb200: no whitening in progress on this slot");
  e->bind();
  CUDA_CHECK(cudaStreamSynchronize(e->stream));
  bool ok = true;
  if (w.multi) ok = ranks_agree(e, !w.failed);           // every rank's stores into my rows have landed, and none gave up
  e->whiten_ms = w.gemm_ms;
  e->whiten_flops = w.flops;
  whiten_abort(e);                                       // releases the scratch and unmaps the peers
  if (!ok) throw Failure("mqcb200: a rank failed during the sharded whitening");
  e->slots[slot].set = true;
}

// Whole host tensor three(n*n, naux) through the slab path (single rank or sharded alike).
static void whiten_all_from_host(Engine *e, int slot, int n, int naux, const double *three, int nu_lo, int nu_hi) {
  // slabs of whole tile columns, at most ~512 MiB of staging each
  const size_t per_col = (size_t)n * naux * sizeof(double) * TILE;
  int cols_per = (int)std::max<size_t>(1, ((size_t)512 << 20) / std::max<size_t>(per_col, 1));
  for (int nu0 = nu_lo; nu0 < nu_hi; nu0 += cols_per * TILE) {
    const int cnt = std::min(cols_per * TILE, nu_hi - nu0);
    whiten_push(e, slot, nu0, cnt, three + (size_t)nu0 * n, (long long)n * n);
  }
}

// ------------------------------- fragment FIFO -------------------------------------------
struct Fifo {
  uint32_t magic = 0x4D51F1F0u;
  std::vector<int64_t> ids;
  size_t head = 0;
  std::mutex m;
};

}  // namespace mqcb200

// =============================================================================================
// C ABI
// =============================================================================================
using namespace mqcb200;

#define API_BEGIN try {
#define API_END                                   \
  }                                               \
  catch (const std::exception &ex) {              \
    g_last_error = ex.what();                     \
    return MQCB200_FAIL;                          \
  }                                               \
  catch (...) {                                   \
    g_last_error = "mqcb200: unknown failure";    \
    return MQCB200_FAIL;                          \
  }                                               \
  return MQCB200_OK;

#define GET_ENGINE(h)                                                   \
  Engine *e = as_engine(h);                                             \
  if (!e) {                                                             \
    g_last_error = "mqcb200: the handle names no engine";               \
    return MQCB200_BAD_HANDLE;                                          \
  }

extern "C" {

int mqcb200_version(void) { return 100; }

void mqcb200_last_error(int buffer_len, char *buffer) {
  if (!buffer || buffer_len <= 0) return;
  std::snprintf(buffer, (size_t)buffer_len, "%s", g_last_error.c_str());
}

int mqcb200_create(int device_rank, void **handle) {
  if (!handle) {
    g_last_error = "mqcb200: null handle pointer";
    return MQCB200_FAIL;
  }
  *handle = nullptr;
  API_BEGIN
  int count = 0;
  cudaError_t err = cudaGetDeviceCount(&count);
  if (err != cudaSuccess || count <= 0)
    throw Failure(std::string("mqcb200: no CUDA device is visible (") + cudaGetErrorString(err) +
                  "); this engine has no CPU fallback");
  if (device_rank < 0) throw Failure("mqcb200: negative device rank");
  Engine *e = new Engine();
  try {
    e->device = device_rank % count;  // cf. mqc_cuest_context.f90:188
    e->bind();
    cudaDeviceProp prop;
    CUDA_CHECK(cudaGetDeviceProperties(&prop, e->device));
    if (prop.major < 10)
      throw Failure(std::string("mqcb200: device '") + prop.name + "' is not sm_100-class; the kernels are built for sm_100a only");
    e->sm_count = prop.multiProcessorCount;
    CUDA_CHECK(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
    CUDA_CHECK(cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking));
    {
      int least = 0, greatest = 0;
      CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&least, &greatest));
      CUDA_CHECK(cudaStreamCreateWithPriority(&e->j_stream, cudaStreamNonBlocking, greatest));
    }
    CUDA_CHECK(cudaEventCreateWithFlags(&e->ev_late_upload, cudaEventDisableTiming));
    CUDA_CHECK(cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming));
    CUDA_CHECK(cudaEventCreateWithFlags(&e->ev_k1_done, cudaEventDisableTiming));
    CUDA_CHECK(cudaEventCreateWithFlags(&e->ev_j_done, cudaEventDisableTiming));
    if (const char *env = getenv("MQCB200_OVERLAP_J")) e->overlap_j = !(env[0] == '0');
    if (const char *env = getenv("MQCB200_CORESIDENT_J")) e->coresident_j = !(env[0] == '0');
    e->d_scalar.ensure(128);
    CUDA_CHECK(cudaMemsetAsync(e->d_scalar.ptr, 0, 128, e->stream));
    e->d_escratch.ensure(160 * sizeof(double));
    CUDA_CHECK(cudaMemsetAsync(e->d_escratch.ptr, 0, 160 * sizeof(double), e->stream));
    configure_kernels();
    CUDA_CHECK(cudaGetLastError());
  } catch (...) {
    delete e;
    throw;
  }
  *handle = e;
  API_END
}

int mqcb200_destroy(void *handle) {
  GET_ENGINE(handle)
  API_BEGIN
  cudaSetDevice(e->device);
  if (e->stream) cudaStreamSynchronize(e->stream);
  if (e->wh.active) whiten_abort(e);
  if (e->comm) p2p_teardown(e);
  if (e->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(e->comm);
  for (auto &sl : e->slots) sl.packed.release();
  DevBuf *bufs[] = {&e->d_w, &e->d_ctf, &e->d_gamma_partial,
                    &e->d_gamma, &e->d_jpart, &e->d_x, &e->d_kpart, &e->d_gpart_a, &e->d_gpart_b, &e->d_cep, &e->d_stack, &e->d_scf, &e->d_jk, &e->d_in, &e->d_out,
                    &e->d_scalar, &e->d_stage, &e->d_escratch};
  for (DevBuf *b : bufs) b->release();
  e->h_in.release();
  e->h_out.release();
  e->h_scf.release();
  for (int i = 0; i < 2; ++i) {
    e->h_stage[i].release();
    e->d_stage_lower[i].release();
    if (e->ev_stage[i]) cudaEventDestroy(e->ev_stage[i]);
  }
  for (auto &sp : e->spans) {
    cudaEventDestroy(sp.a);
    cudaEventDestroy(sp.b);
  }
  if (e->copy_stream) cudaStreamDestroy(e->copy_stream);
  if (e->j_stream) cudaStreamDestroy(e->j_stream);
  for (cudaEvent_t ev : {e->ev_late_upload, e->ev_fork, e->ev_k1_done, e->ev_j_done})
    if (ev) cudaEventDestroy(ev);
  if (e->stream) cudaStreamDestroy(e->stream);
  e->magic = 0;
  delete e;
  API_END
}

int mqcb200_get_stream(void *handle, void **stream) {
  GET_ENGINE(handle)
  if (!stream) { g_last_error = "mqcb200: null stream pointer"; return MQCB200_FAIL; }
  *stream = e->stream;
  return MQCB200_OK;
}

int mqcb200_set_workspace_limit(void *handle, size_t bytes) {
  GET_ENGINE(handle)
  if (bytes < ((size_t)1 << 20)) { g_last_error = "mqcb200: workspace limit below 1 MiB"; return MQCB200_FAIL; }
  e->workspace_limit = bytes;
  return MQCB200_OK;
}

int mqcb200_set_fuse_threshold(void *handle, size_t bytes) {
  GET_ENGINE(handle)
  e->fuse_threshold = bytes;
  return MQCB200_OK;
}

int mqcb200_set_overlap(void *handle, int on) {
  GET_ENGINE(handle)
  e->overlap_j = on != 0;
  return MQCB200_OK;
}

int mqcb200_tensor_shape(void *handle, int slot, int *n, int *naux_total, int *q_begin, int *q_count) {
  GET_ENGINE(handle)
  if (slot < 0 || slot >= MQCB200_NUM_SLOTS) { g_last_error = "mqcb200: tensor slot out of range"; return MQCB200_FAIL; }
  const TensorSlot &sl = e->slots[slot];
  if (n) *n = sl.set ? sl.n : 0;
  if (naux_total) *naux_total = sl.set ? sl.naux_total : 0;
  if (q_begin) *q_begin = sl.set ? sl.q_begin : 0;
  if (q_count) *q_count = sl.set ? sl.q_count : 0;
  return MQCB200_OK;
}

int mqcb200_set_tensor(void *handle, int slot, int n, int naux, const double *b) {
  GET_ENGINE(handle)
  API_BEGIN
  set_tensor_host(e, slot, n, naux, 0, naux, b);
  API_END
}

int mqcb200_set_tensor_shard(void *handle, int slot, int n, int naux_total, int q_begin, int q_count,
                             const double *b_shard) {
  GET_ENGINE(handle)
  API_BEGIN
  set_tensor_host(e, slot, n, naux_total, q_begin, q_count, b_shard);
  API_END
}

int mqcb200_set_tensor_from_3c(void *handle, int slot, int n, int naux, const double *three, const double *half) {
  GET_ENGINE(handle)
  API_BEGIN
  if (!three || !half) throw Failure("mqcb200: null three-centre tensor or metric factor");
  if (e->comm && e->n_ranks > 1) throw Failure("mqcb200: with a communicator active use mqcb200_whiten_begin/push/end (every rank streams its slabs)");
  whiten_begin(e, slot, n, naux, 0, naux, half, false);
  try {
    whiten_all_from_host(e, slot, n, naux, three, 0, n);
  } catch (...) {
    whiten_abort(e);
    throw;
  }
  whiten_end(e, slot);
  API_END
}

int mqcb200_whiten_begin(void *handle, int slot, int n, int naux_total, int q_begin, int q_count, const double *half) {
  GET_ENGINE(handle)
  API_BEGIN
  whiten_begin(e, slot, n, naux_total, q_begin, q_count, half, false);
  API_END
}

int mqcb200_whiten_push(void *handle, int slot, int nu_begin, int nu_count, const double *three_cols, long long ld_aux) {
  GET_ENGINE(handle)
  API_BEGIN
  try {
    whiten_push(e, slot, nu_begin, nu_count, three_cols, ld_aux);
  } catch (...) {
    // a lone rank gives the whitening up at once; a rank of a sharded one must still meet the others in
    // whiten_end (collective), where its failure becomes everyone's
    if (e->wh.active && !e->wh.multi) whiten_abort(e);
    else if (e->wh.active) e->wh.failed = true;
    throw;
  }
  API_END
}

int mqcb200_whiten_end(void *handle, int slot) {
  GET_ENGINE(handle)
  API_BEGIN
  whiten_end(e, slot);
  API_END
}

int mqcb200_metric_inverse_sqrt(void *handle, int naux, const double *metric, double null_threshold, double *half,
                                int *n_kept) {
  GET_ENGINE(handle)
  API_BEGIN
  if (!half) throw Failure("mqcb200: null output for the metric factor");
  e->bind();
  DevBuf d_half;
  try {
    d_half.ensure((size_t)naux * naux * sizeof(double));
    const int kept = metric_inverse_sqrt_device(e, naux, metric, null_threshold, d_half.d());
    CUDA_CHECK(cudaMemcpyAsync(half, d_half.ptr, (size_t)naux * naux * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    CUDA_CHECK(cudaStreamSynchronize(e->stream));
    if (n_kept) *n_kept = kept;
  } catch (...) {
    d_half.release();
    throw;
  }
  d_half.release();
  API_END
}

int mqcb200_build_df_tensor(void *handle, int slot, int n, int naux, const double *three, const double *metric,
                            double null_threshold, double *half_out) {
  GET_ENGINE(handle)
  API_BEGIN
  if (!three || !metric) throw Failure("mqcb200: null three-centre tensor or metric");
  if (e->comm && e->n_ranks > 1) throw Failure("mqcb200: with a communicator active use mqcb200_metric_inverse_sqrt + mqcb200_whiten_begin/push/end");
  e->bind();
  DevBuf d_half;
  try {
    d_half.ensure((size_t)naux * naux * sizeof(double));
    metric_inverse_sqrt_device(e, naux, metric, null_threshold, d_half.d());
    if (half_out) CUDA_CHECK(cudaMemcpyAsync(half_out, d_half.ptr, (size_t)naux * naux * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    whiten_begin(e, slot, n, naux, 0, naux, d_half.d(), true);
    try {
      whiten_all_from_host(e, slot, n, naux, three, 0, n);
    } catch (...) {
      whiten_abort(e);
      throw;
    }
    whiten_end(e, slot);
  } catch (...) {
    d_half.release();
    throw;
  }
  d_half.release();
  API_END
}

int mqcb200_last_metric(void *handle, double *ms, int *sweeps) {
  GET_ENGINE(handle)
  if (ms) *ms = e->metric_ms;
  if (sweeps) *sweeps = e->metric_sweeps;
  return MQCB200_OK;
}

int mqcb200_synth_tensor(void *handle, int slot, int n, int naux_total, int q_begin, int q_count, uint64_t seed,
                         double scale) {
  GET_ENGINE(handle)
  API_BEGIN
  if (slot < 0 || slot >= MQCB200_NUM_SLOTS) throw Failure("mqcb200: tensor slot out of range");
  e->bind();
  TensorSlot &sl = e->slots[slot];
  slot_prepare(e, sl, n, naux_total, q_begin, q_count);
  if (q_count > 0) launch_synth_tensor(sl.packed.d(), n, q_begin, q_count, seed, scale, e->stream);
  CUDA_CHECK(cudaGetLastError());
  CUDA_CHECK(cudaStreamSynchronize(e->stream));
  sl.set = true;
  API_END
}

int mqcb200_clear_tensor(void *handle, int slot) {
  GET_ENGINE(handle)
  API_BEGIN
  if (slot < 0 || slot >= MQCB200_NUM_SLOTS) throw Failure("mqcb200: tensor slot out of range");
  e->bind();
  CUDA_CHECK(cudaStreamSynchronize(e->stream));
  e->slots[slot].packed.release();
  e->slots[slot].set = false;
  API_END
}

int mqcb200_tensor_bytes(void *handle, int slot, size_t *bytes) {
  GET_ENGINE(handle)
  if (slot < 0 || slot >= MQCB200_NUM_SLOTS || !bytes) { g_last_error = "mqcb200: bad slot or null output"; return MQCB200_FAIL; }
  const TensorSlot &sl = e->slots[slot];
  *bytes = sl.set ? (size_t)sl.L * (size_t)sl.q_count * sizeof(double) : 0;
  return MQCB200_OK;
}

int mqcb200_build_fock(void *handle, int slot, const double *h, const double *density, const double *coeff, int ldc,
                       int n_occ, double k_scale, double j_scale, double *fock) {
  GET_ENGINE(handle)
  API_BEGIN
  if (!h || !density || !fock) throw Failure("mqcb200: null matrix argument to build_fock");
  BuildArgs a;
  a.slot = slot; a.h = h; a.density = density; a.coeff_a = coeff; a.lda = ldc; a.n_a = n_occ;
  a.k_scale = k_scale; a.j_scale = j_scale; a.fock_a = fock; a.assemble = true;
  // j_scale == 0 is the attenuated second pass (rhf.f90:1100): J is not needed at all
  a.want_j = j_scale != 0.0;
  a.want_k = k_scale != 0.0;
  build(e, a);
  API_END
}

int mqcb200_build_jk(void *handle, int slot, const double *density, const double *coeff, int ldc, int n_occ,
                     double *j, double *k) {
  GET_ENGINE(handle)
  API_BEGIN
  BuildArgs a;
  a.slot = slot; a.density = density; a.coeff_a = coeff; a.lda = ldc; a.n_a = n_occ;
  a.j = j; a.k_a = k; a.want_j = j != nullptr; a.want_k = k != nullptr;
  if (a.want_j && !density) throw Failure("mqcb200: J was requested without a density");
  build(e, a);
  API_END
}

int mqcb200_build_jk_uhf(void *handle, int slot, const double *density_total, const double *coeff_a, int lda,
                         int n_alpha, const double *coeff_b, int ldb, int n_beta, double *j, double *k_alpha,
                         double *k_beta) {
  GET_ENGINE(handle)
  API_BEGIN
  BuildArgs a;
  a.slot = slot; a.density = density_total; a.two_spin = true;
  a.coeff_a = coeff_a; a.lda = lda; a.n_a = n_alpha; a.coeff_b = coeff_b; a.ldb = ldb; a.n_b = n_beta;
  a.j = j; a.k_a = k_alpha; a.k_b = k_beta; a.want_j = j != nullptr; a.want_k = (k_alpha != nullptr) || (k_beta != nullptr);
  if (a.want_j && !density_total) throw Failure("mqcb200: J was requested without a density");
  build(e, a);
  API_END
}

int mqcb200_build_fock_uhf(void *handle, int slot, const double *h, const double *density_total,
                           const double *coeff_a, int lda, int n_alpha, const double *coeff_b, int ldb, int n_beta,
                           double k_scale, double *fock_a, double *fock_b) {
  GET_ENGINE(handle)
  API_BEGIN
  if (!h || !density_total || !fock_a || !fock_b) throw Failure("mqcb200: null matrix argument to build_fock_uhf");
  BuildArgs a;
  a.slot = slot; a.h = h; a.density = density_total; a.two_spin = true;
  a.coeff_a = coeff_a; a.lda = lda; a.n_a = n_alpha; a.coeff_b = coeff_b; a.ldb = ldb; a.n_b = n_beta;
  a.k_scale = k_scale; a.j_scale = 1.0; a.fock_a = fock_a; a.fock_b = fock_b; a.assemble = true;
  a.want_k = k_scale != 0.0;
  build(e, a);
  API_END
}

int mqcb200_build_fock_device(void *handle, int slot, const double *d_h, const double *d_density,
                              const double *d_coeff, int n_occ, double k_scale, double j_scale, double *d_fock,
                              int sync) {
  GET_ENGINE(handle)
  API_BEGIN
  if (!d_h || !d_density || !d_fock) throw Failure("mqcb200: null device pointer argument to build_fock_device");
  if (slot < 0 || slot >= MQCB200_NUM_SLOTS) throw Failure("mqcb200: tensor slot out of range");
  BuildArgs a;
  a.slot = slot; a.device_operands = true; a.h = d_h; a.density = d_density; a.coeff_a = d_coeff;
  a.lda = e->slots[slot].n; a.n_a = n_occ; a.k_scale = k_scale; a.j_scale = j_scale; a.fock_a = d_fock;
  a.assemble = true; a.sync = sync != 0;
  a.want_j = j_scale != 0.0;
  a.want_k = k_scale != 0.0;
  build(e, a);
  API_END
}

int mqcb200_build_g_two_factor(void *handle, int slot, const double *density, const double *coeff_a, int lda,
                               int n_a, const double *coeff_b, int ldb, int n_b, double ka, double kb, double *g) {
  GET_ENGINE(handle)
  API_BEGIN
  if (!density || !g) throw Failure("mqcb200: null matrix argument to build_g_two_factor");
  BuildArgs a;
  a.slot = slot; a.density = density; a.two_spin = true;
  a.coeff_a = coeff_a; a.lda = lda; a.n_a = n_a; a.coeff_b = coeff_b; a.ldb = ldb; a.n_b = n_b;
  a.combine = true; a.ka_coef = ka; a.kb_coef = kb; a.fock_a = g;
  a.want_k = (ka != 0.0 && n_a > 0) || (kb != 0.0 && n_b > 0);
  build(e, a);
  API_END
}

int mqcb200_response_operator(void *handle, int slot, const double *x, int ldx, const double *c_occ, int ldc,
                              int n_occ, const double *dtilde, double k_scale, double *g) {
  GET_ENGINE(handle)
  API_BEGIN
  if (!dtilde || !g) throw Failure("mqcb200: null matrix argument to response_operator");
  BuildArgs a;
  a.slot = slot; a.density = dtilde; a.rank2 = true;
  a.coeff_a = x; a.lda = ldx; a.n_a = n_occ; a.coeff_b = c_occ; a.ldb = ldc; a.n_b = n_occ;
  a.combine = true; a.ka_coef = -0.5 * k_scale; a.fock_a = g;        // g = coul - 0.5*kf*exch  (cphf.F90:565)
  a.want_k = k_scale != 0.0 && n_occ > 0;
  if (!a.want_k) a.rank2 = false;
  build(e, a);
  API_END
}

int mqcb200_fitted_potential_general(void *handle, int slot, const double *dens, double k_scale, double *g) {
  GET_ENGINE(handle)
  API_BEGIN
  if (!dens || !g) throw Failure("mqcb200: null matrix argument to fitted_potential_general");
  if (slot < 0 || slot >= MQCB200_NUM_SLOTS || !e->slots[slot].set) throw Failure("mqcb200: no fitted tensor has been set on this slot (call mqcb200_set_tensor first)");
  const int n = e->slots[slot].n;
  // sum_P B_P D B_P = 1/2 [(B_P D)(B_P 1)^T + (B_P 1)(B_P D)^T] for the symmetric D the reference
  // requires (cphf.F90:570): the rank-2 kernels with X = D, C = 1 -- no eigendecomposition, no cancellation
  std::vector<double> eye((size_t)n * n, 0.0);
  for (int i = 0; i < n; ++i) eye[(size_t)i * n + i] = 1.0;
  BuildArgs a;
  a.slot = slot; a.density = dens; a.rank2 = true;
  a.coeff_a = dens; a.lda = n; a.n_a = n; a.coeff_b = eye.data(); a.ldb = n; a.n_b = n;
  a.combine = true; a.ka_coef = -0.25 * k_scale; a.fock_a = g;       // g = coul - 0.5*kf*exch  (cphf.F90:615)
  a.want_k = k_scale != 0.0;
  if (!a.want_k) a.rank2 = false;
  build(e, a);
  API_END
}

int mqcb200_build_fock_uhf_device(void *handle, int slot, const double *d_h, const double *d_density_total,
                                  const double *d_coeff_a, int n_alpha, const double *d_coeff_b, int n_beta,
                                  double k_scale, double *d_fock_a, double *d_fock_b, int sync) {
  GET_ENGINE(handle)
  API_BEGIN
  if (!d_h || !d_density_total || !d_fock_a || !d_fock_b)
    throw Failure("mqcb200: null device pointer argument to build_fock_uhf_device");
  if (slot < 0 || slot >= MQCB200_NUM_SLOTS) throw Failure("mqcb200: tensor slot out of range");
  BuildArgs a;
  a.slot = slot; a.device_operands = true; a.two_spin = true; a.h = d_h; a.density = d_density_total;
  a.coeff_a = d_coeff_a; a.lda = e->slots[slot].n; a.n_a = n_alpha;
  a.coeff_b = d_coeff_b; a.ldb = e->slots[slot].n; a.n_b = n_beta;
  a.k_scale = k_scale; a.j_scale = 1.0; a.fock_a = d_fock_a; a.fock_b = d_fock_b;
  a.assemble = true; a.sync = sync != 0;
  a.want_k = k_scale != 0.0;
  build(e, a);
  API_END
}

int mqcb200_scf_fragment(void *handle, int slot, const double *hcore, const double *overlap, int n_electrons,
                         int guess, int max_iter, double energy_tol, double density_tol, int diis_vectors,
                         double k_scale, double *e_electronic, int *iterations, int *converged, int *n_mo,
                         double *coeff, double *orbital_energies, double *density, double *e_history) {
  GET_ENGINE(handle)
  API_BEGIN
  ScfArgs a;
  a.slot = slot; a.h = hcore; a.s = overlap; a.n_electrons = n_electrons; a.guess = guess; a.max_iter = max_iter;
  a.energy_tol = energy_tol; a.density_tol = density_tol; a.diis_vectors = diis_vectors; a.k_scale = k_scale;
  a.e_electronic = e_electronic; a.iterations = iterations; a.converged = converged; a.n_mo = n_mo;
  a.coeff = coeff; a.eps = orbital_energies; a.density = density; a.e_history = e_history;
  scf_fragment(e, a);
  API_END
}

int mqcb200_df_gradient_densities(void *handle, int slot, const double *half, const double *total_density,
                                  const double *orbitals, int lda, int n_occupied, const double *orbitals_beta,
                                  int ldb, int n_occupied_beta, int unrestricted, double exx_fraction,
                                  int with_coulomb, double *gamma, double *omega) {
  GET_ENGINE(handle)
  API_BEGIN
  GradArgs a;
  a.slot = slot; a.half = half; a.density = total_density; a.ca = orbitals; a.lda = lda; a.n_a = n_occupied;
  a.cb = orbitals_beta; a.ldb = ldb; a.n_b = n_occupied_beta; a.unrestricted = unrestricted != 0;
  a.kf = exx_fraction; a.with_coulomb = with_coulomb != 0; a.gamma = gamma; a.omega = omega;
  df_gradient_densities(e, a);
  API_END
}

int mqcb200_scf(void *handle, int slot, const double *hcore, const double *overlap, int n_electrons, int guess, int max_iter,
                double energy_tol, double density_tol, int diis_vectors, double k_scale, double *e_electronic,
                int *iterations, int *converged, int *n_mo, double *coeff, double *orbital_energies, double *density,
                double *e_history) {
  GET_ENGINE(handle)
  API_BEGIN
  ScfArgs a;
  a.slot = slot; a.h = hcore; a.s = overlap; a.n_electrons = n_electrons; a.guess = guess; a.max_iter = max_iter;
  a.energy_tol = energy_tol; a.density_tol = density_tol; a.diis_vectors = diis_vectors; a.k_scale = k_scale;
  a.e_electronic = e_electronic; a.iterations = iterations; a.converged = converged; a.n_mo = n_mo;
  a.coeff = coeff; a.eps = orbital_energies; a.density = density; a.e_history = e_history;
  if (slot < 0 || slot >= MQCB200_NUM_SLOTS) throw Failure("mqcb200: tensor slot out of range");
  const TensorSlot &sl = e->slots[slot];
  const bool fragment_sized = sl.set && scf_path_applies(sl.n) && fragment_path_applies(sl.n, std::max(n_electrons / 2, 1)) &&
                              !(e->comm && e->n_ranks > 1) && sl.q_count == sl.naux_total && fragment_path_enabled();
  if (fragment_sized) scf_fragment(e, a);
  else scf_general(e, a);
  API_END
}

int mqcb200_scf_fragment_batch(void *handle, int slot, int n_fragments, const double *hcore_all, const double *overlap_all,
                               int n_electrons, int guess, int max_iter, double energy_tol, double density_tol,
                               int diis_vectors, double k_scale, double *e_electronic, int *iterations, int *converged,
                               int *n_mo, double *coeff_all, double *orbital_energies_all, double *density_all) {
  GET_ENGINE(handle)
  API_BEGIN
  ScfBatchArgs a;
  a.slot = slot; a.n_frag = n_fragments; a.h = hcore_all; a.s = overlap_all; a.n_electrons = n_electrons; a.guess = guess;
  a.max_iter = max_iter; a.energy_tol = energy_tol; a.density_tol = density_tol; a.diis_vectors = diis_vectors; a.k_scale = k_scale;
  a.e_electronic = e_electronic; a.iterations = iterations; a.converged = converged; a.n_mo = n_mo;
  a.coeff = coeff_all; a.eps = orbital_energies_all; a.density = density_all;
  scf_fragment_batch(e, a);
  API_END
}

int mqcb200_set_scf_check_every(void *handle, int iterations) {
  GET_ENGINE(handle)
  if (iterations < 1) { g_last_error = "mqcb200: check_every must be positive"; return MQCB200_FAIL; }
  e->scf_check_every = iterations;
  return MQCB200_OK;
}

int mqcb200_last_energy(void *handle, double *e_elec) {
  GET_ENGINE(handle)
  API_BEGIN
  if (!e_elec) throw Failure("mqcb200: null energy pointer");
  if (!e->have_last_fock) throw Failure("mqcb200: no closed-shell build_fock has run on this handle yet");
  *e_elec = e->last_energy_host;   // computed on the device by that build, already on the host
  API_END
}

// ---- multi-GPU -------------------------------------------------------------------------
int mqcb200_comm_unique_id(char id[128]) {
  API_BEGIN
  if (!id) throw Failure("mqcb200: null id buffer");
  load_nccl();
  NCCL_CHECK(g_nccl.GetUniqueId(id));
  API_END
}

int mqcb200_comm_init(void *handle, int n_ranks, int rank, const char id[128]) {
  GET_ENGINE(handle)
  API_BEGIN
  if (!id || n_ranks < 1 || rank < 0 || rank >= n_ranks) throw Failure("mqcb200: bad communicator arguments");
  if (e->comm) throw Failure("mqcb200: a communicator is already active on this handle");
  load_nccl();
  e->bind();
  Id128 uid;
  std::memcpy(uid.bytes, id, 128);
  void *comm = nullptr;
  NCCL_CHECK(g_nccl.CommInitRank(&comm, n_ranks, uid, rank));
  e->comm = comm;
  e->n_ranks = n_ranks;
  e->rank = rank;
  // NVLink peer-memory exchange: set up collectively; unless EVERY rank mapped its peers the
  // builds use the NCCL all-reduce (p2p_setup agrees on that itself; a mixed choice would deadlock)
  if (n_ranks > 1) p2p_setup(e);
  API_END
}

int mqcb200_comm_destroy(void *handle) {
  GET_ENGINE(handle)
  API_BEGIN
  if (e->comm) {
    e->bind();
    CUDA_CHECK(cudaStreamSynchronize(e->stream));
    p2p_teardown(e);
    NCCL_CHECK(g_nccl.CommDestroy(e->comm));
    e->comm = nullptr;
    e->n_ranks = 1;
    e->rank = 0;
  }
  API_END
}

// ---- fragment FIFO ------------------------------------------------------------------------
int mqcb200_queue_create(const int64_t *ids, int64_t count, void **queue) {
  API_BEGIN
  if (!queue || count < 0 || (count > 0 && !ids)) throw Failure("mqcb200: bad queue arguments");
  Fifo *f = new Fifo();
  f->ids.assign(ids, ids + count);
  *queue = f;
  API_END
}

static Fifo *as_fifo(void *q) {
  Fifo *f = static_cast<Fifo *>(q);
  return (f && f->magic == 0x4D51F1F0u) ? f : nullptr;
}

int mqcb200_queue_pop(void *queue, int64_t *id, int *has_item) {
  Fifo *f = as_fifo(queue);
  if (!f) { g_last_error = "mqcb200: the handle names no queue"; return MQCB200_BAD_HANDLE; }
  if (!id || !has_item) { g_last_error = "mqcb200: null output pointer"; return MQCB200_FAIL; }
  std::lock_guard<std::mutex> lock(f->m);
  if (f->head >= f->ids.size()) {  // queue_pop, mqc_work_queue.f90:36-40
    *id = -1;
    *has_item = 0;
  } else {
    *id = f->ids[f->head++];
    *has_item = 1;
  }
  return MQCB200_OK;
}

int mqcb200_queue_is_empty(void *queue, int *is_empty) {
  Fifo *f = as_fifo(queue);
  if (!f) { g_last_error = "mqcb200: the handle names no queue"; return MQCB200_BAD_HANDLE; }
  if (!is_empty) { g_last_error = "mqcb200: null output pointer"; return MQCB200_FAIL; }
  std::lock_guard<std::mutex> lock(f->m);
  *is_empty = f->head >= f->ids.size() ? 1 : 0;
  return MQCB200_OK;
}

int mqcb200_queue_destroy(void *queue) {
  Fifo *f = as_fifo(queue);
  if (!f) { g_last_error = "mqcb200: the handle names no queue"; return MQCB200_BAD_HANDLE; }
  f->magic = 0;
  delete f;
  return MQCB200_OK;
}

// ---- instrumentation ----------------------------------------------------------------------
int mqcb200_set_profiling(void *handle, int on) {
  GET_ENGINE(handle)
  e->profiling = on != 0;
  if (!e->profiling) e->n_spans = 0;
  return MQCB200_OK;
}

int mqcb200_last_timings(void *handle, double ms[MQCB200_NUM_TIMERS]) {
  GET_ENGINE(handle)
  if (!ms) { g_last_error = "mqcb200: null output pointer"; return MQCB200_FAIL; }
  API_BEGIN
  e->bind();
  CUDA_CHECK(cudaStreamSynchronize(e->stream));
  e->collect_timers();
  for (int i = 0; i < MQCB200_NUM_TIMERS; ++i) ms[i] = e->last_ms[i];
  if (e->comm && e->n_ranks > 1) p2p_raise_if_failed(e);   // an asynchronous build's exchange that gave up
  API_END
}

int mqcb200_last_set_tensor(void *handle, double *ms, double *h2d_bytes) {
  GET_ENGINE(handle)
  if (!ms || !h2d_bytes) { g_last_error = "mqcb200: null output pointer"; return MQCB200_FAIL; }
  *ms = e->set_tensor_ms;
  *h2d_bytes = e->set_tensor_h2d_bytes;
  return MQCB200_OK;
}

int mqcb200_last_whiten(void *handle, double *ms, double *flops) {
  GET_ENGINE(handle)
  if (!ms || !flops) { g_last_error = "mqcb200: null output pointer"; return MQCB200_FAIL; }
  *ms = e->whiten_ms;
  *flops = e->whiten_flops;
  return MQCB200_OK;
}

int mqcb200_last_gamma_fused(void *handle, int *fused) {
  GET_ENGINE(handle)
  API_BEGIN
  if (!fused) throw Failure("mqcb200: null output pointer");
  *fused = 0;
  if (e->last_fuse_attempted) {
    e->bind();
    int flag = 0;
    CUDA_CHECK(cudaMemcpyAsync(&flag, static_cast<char *>(e->d_scalar.ptr) + 32, sizeof(int), cudaMemcpyDeviceToHost,
                               e->stream));
    CUDA_CHECK(cudaStreamSynchronize(e->stream));
    *fused = flag != 0 ? 1 : 0;
  }
  API_END
}

int mqcb200_last_launches(void *handle, int *n_kernels) {
  GET_ENGINE(handle)
  if (!n_kernels) { g_last_error = "mqcb200: null output pointer"; return MQCB200_FAIL; }
  *n_kernels = e->launches;
  return MQCB200_OK;
}

}  // extern "C"

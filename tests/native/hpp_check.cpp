// The C++ host-side mirror (include/mqcb200.hpp) compiled by a plain host compiler with warnings as errors and
// linked against libmqcb200.so.  Run on the CPU by tests/test_cpp_mirror.py: the work queue (pure host code) is
// exercised in full; the engine must refuse to exist without a CUDA device -- loudly, with the library's message
// and status code -- because there is no CPU fallback behind this interface.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <type_traits>
#include <vector>

#include "../../include/mqcb200.hpp"

static int failures = 0;
#define CHECK(cond, what)                              \
  do {                                                 \
    if (!(cond)) {                                     \
      ++failures;                                      \
      std::printf("FAIL %s (%s)\n", what, #cond);      \
    }                                                  \
  } while (0)

// the mirror keeps the reference's argument order: build_fock_df(h, [b resident], density, coeff, n_occ, fock, k_scale, j_scale)
static_assert(std::is_same<decltype(&mqcb200::FockEngine::build_fock_df),
                           void (mqcb200::FockEngine::*)(const double *, const double *, const double *, int, int, double *,
                                                         double, double, mqcb200::Slot)>::value,
              "build_fock_df signature");
static_assert(!std::is_copy_constructible<mqcb200::FockEngine>::value && std::is_move_constructible<mqcb200::FockEngine>::value,
              "an engine handle is owned once");

int main() {
  // ---- queue_t semantics (mqc_work_queue.f90:10-57): FIFO, has_item = false and id = -1 once drained
  {
    mqcb200::WorkQueue q({7, 3, 11});
    std::int64_t id = 0;
    CHECK(!q.is_empty(), "fresh queue is not empty");
    CHECK(q.pop(id) && id == 7, "first pop");
    CHECK(q.pop(id) && id == 3, "second pop");
    CHECK(q.pop(id) && id == 11, "third pop");
    CHECK(q.is_empty(), "drained");
    CHECK(!q.pop(id) && id == -1, "pop on a drained queue");
    mqcb200::WorkQueue none(std::vector<std::int64_t>{});
    CHECK(none.is_empty() && !none.pop(id), "empty list");
  }
  {  // several host threads (one per GPU in the reference's dispatcher) drain one queue: every id exactly once
    std::vector<std::int64_t> ids(20000);
    for (std::size_t i = 0; i < ids.size(); ++i) ids[i] = static_cast<std::int64_t>(i);
    mqcb200::WorkQueue q(ids);
    std::vector<std::vector<std::int64_t>> got(8);
    std::vector<std::thread> pool;
    for (int t = 0; t < 8; ++t)
      pool.emplace_back([&q, &got, t] {
        std::int64_t id;
        while (q.pop(id)) got[static_cast<std::size_t>(t)].push_back(id);
      });
    for (auto &th : pool) th.join();
    std::vector<int> seen(ids.size(), 0);
    std::size_t total = 0;
    bool ordered = true;
    for (const auto &g : got) {
      total += g.size();
      for (std::size_t i = 0; i < g.size(); ++i) {
        ++seen[static_cast<std::size_t>(g[i])];
        if (i > 0 && g[i] <= g[i - 1]) ordered = false;      // each worker sees the FIFO order
      }
    }
    bool once = total == ids.size();
    for (int s : seen) once = once && s == 1;
    CHECK(once, "every id popped exactly once across 8 threads");
    CHECK(ordered, "FIFO order per worker");
  }
  // ---- the engine: no device, no engine
  bool have_device = true;
  try {
    mqcb200::FockEngine engine(0);
    int n = -1, naux = -1, qb = -1, qc = -1;
    engine.tensor_shape(n, naux, qb, qc);
    CHECK(n == 0 && naux == 0 && qb == 0 && qc == 0, "an empty slot reports zeros");
    bool refused = false;
    try {
      double one = 1.0, f = 0.0;
      engine.build_fock_df(&one, &one, &one, 1, 1, &f);       // no tensor has been set
    } catch (const mqcb200::Error &e) {
      refused = e.code() == MQCB200_FAIL && std::string(e.what()).find("no fitted tensor") != std::string::npos;
    }
    CHECK(refused, "a build before set_tensor is refused with the library's message");
    std::printf("device present: engine created and destroyed\n");
  } catch (const mqcb200::Error &e) {
    have_device = false;
    CHECK(e.code() == MQCB200_FAIL, "status code of the refusal");
    CHECK(std::string(e.what()).find("no CUDA device") != std::string::npos, "the refusal names the missing device");
    CHECK(std::string(e.what()).find("no CPU fallback") != std::string::npos, "the refusal says there is no fallback");
    std::printf("no device: %s\n", e.what());
  }
  (void)have_device;
  if (failures) { std::printf("%d check(s) FAILED\n", failures); return 1; }
  std::printf("C++ mirror checks hold\n");
  return 0;
}

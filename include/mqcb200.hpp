// mqcb200.hpp -- C++17 host-side mirror of the reference's density-fitted Fock-build interface on top
// of the C ABI (include/mqcb200.h).  Header-only; link with -lmqcb200.
//
// The reference is compiled code (Fortran).  Its toolchain is not part of this image, so next to the
// Fortran bindings shipped as source (metalquicha_b200/fortran/) this is the compiled-language face of
// the boundary: the routine names, argument order, argument meaning and error behaviour are those of
// the reference routines each method stands in for (paths relative to the reference tree):
//
//   FockEngine::build_fock_df            build_fock_df            backends/libcint/mqc_libcint_rhf.f90:1576-1646
//   FockEngine::electronic_energy        electronic_energy        mqc_libcint_rhf.f90:1691-1697 (of the last build)
//   FockEngine::set_tensor               the bmat of build_df_tensor, mqc_libcint_integrals.F90:913-990,
//                                        handed over once per geometry (run_libcint_rhf :486-498)
//   FockEngine::metric_inverse_sqrt      metric_inverse_sqrt      mqc_libcint_integrals.F90:992-1038
//   FockEngine::build_df_tensor          build_df_tensor's last two stages, mqc_libcint_integrals.F90:981-987
//   FockEngine::response_operator_df     response_operator_df     backends/libcint/mqc_libcint_cphf.F90:499-566
//   FockEngine::fitted_potential_general fitted_potential_general mqc_libcint_cphf.F90:568-616
//   FockEngine::df_gradient_densities    df_two_electron_gradient's contractions, mqc_libcint_gradient.f90:1545-1812
//   FockEngine::run_rhf                  run_libcint_rhf's loop   mqc_libcint_rhf.f90:321-680
//   WorkQueue                            queue_t                  src/fragmentation/common/mqc_work_queue.f90:10-57
//
// All matrices are float64, column-major, contiguous (a Fortran array as it lies in memory); the
// optional Fortran arguments k_scale / j_scale become defaulted parameters with the values the
// reference uses when they are absent (rhf.f90:1641-1644).  Where the reference sets error_t and
// returns, these methods throw mqcb200::Error carrying the engine's message and status code.
// There is no CPU fallback: constructing a FockEngine without a CUDA device throws.
#ifndef MQCB200_HPP
#define MQCB200_HPP

#include <cstddef>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "mqcb200.h"

namespace mqcb200 {

class Error : public std::runtime_error {
 public:
  Error(int code, const std::string &message) : std::runtime_error(message), code_(code) {}
  int code() const noexcept { return code_; }   // MQCB200_FAIL or MQCB200_BAD_HANDLE

 private:
  int code_;
};

inline std::string last_error() {
  char buffer[1024];
  buffer[0] = '\0';
  mqcb200_last_error(static_cast<int>(sizeof buffer), buffer);
  return std::string(buffer);
}

inline void check(int status) {
  if (status != MQCB200_OK) throw Error(status, last_error());
}

enum class Slot : int { full_range = MQCB200_SLOT_FULL_RANGE, attenuated = MQCB200_SLOT_ATTENUATED };

// rhf_result_t (mqc_libcint_rhf.f90:117-141), the fields the loop itself produces
struct RhfResult {
  double electronic = 0.0;
  int iterations = 0;
  bool converged = false;
  int n_mo = 0;
  std::vector<double> orbitals;           // n x n_mo, column-major
  std::vector<double> orbital_energies;   // n_mo
  std::vector<double> density;            // n x n
  std::vector<double> e_history;          // electronic energy of every iteration's build
};

// One engine == one GPU, mod(device_rank, device_count): the per-rank device context of the reference's
// GPU backend (backends/cuest/backend/mqc_cuest_context.f90:166-233).
class FockEngine {
 public:
  explicit FockEngine(int device_rank = 0) { check(mqcb200_create(device_rank, &handle_)); }
  ~FockEngine() {
    if (handle_) mqcb200_destroy(handle_);
  }
  FockEngine(const FockEngine &) = delete;
  FockEngine &operator=(const FockEngine &) = delete;
  FockEngine(FockEngine &&other) noexcept : handle_(std::exchange(other.handle_, nullptr)) {}
  FockEngine &operator=(FockEngine &&other) noexcept {
    if (this != &other) {
      if (handle_) mqcb200_destroy(handle_);
      handle_ = std::exchange(other.handle_, nullptr);
    }
    return *this;
  }
  void *handle() const noexcept { return handle_; }

  // ---- the fitted tensor: bmat(nao*nao, naux), once per geometry
  void set_tensor(const double *bmat, int n_ao, int naux, Slot slot = Slot::full_range) {
    check(mqcb200_set_tensor(handle_, static_cast<int>(slot), n_ao, naux, bmat));
  }
  void set_tensor_shard(const double *bmat_shard, int n_ao, int naux_total, int q_begin, int q_count,
                        Slot slot = Slot::full_range) {
    check(mqcb200_set_tensor_shard(handle_, static_cast<int>(slot), n_ao, naux_total, q_begin, q_count, bmat_shard));
  }
  void clear_tensor(Slot slot = Slot::full_range) { check(mqcb200_clear_tensor(handle_, static_cast<int>(slot))); }
  // (n, naux_total, q_begin, q_count) of the resident tensor; zeros when the slot is empty
  void tensor_shape(int &n, int &naux_total, int &q_begin, int &q_count, Slot slot = Slot::full_range) const {
    check(mqcb200_tensor_shape(handle_, static_cast<int>(slot), &n, &naux_total, &q_begin, &q_count));
  }

  // metric_inverse_sqrt(metric, half, error): half = U s^-1/2 U^T over the modes above 1e-10; returns the modes kept
  int metric_inverse_sqrt(const double *metric, int naux, double *half, double null_threshold = 1.0e-10) {
    int kept = 0;
    check(mqcb200_metric_inverse_sqrt(handle_, naux, metric, null_threshold, half, &kept));
    return kept;
  }
  // build_df_tensor's last two stages: three(nao*nao, naux) and the metric in, the whitened tensor resident on
  // `slot` out; half_out (naux x naux, may be null) receives metric^-1/2
  void build_df_tensor(const double *three, const double *metric, int n_ao, int naux, double *half_out = nullptr,
                       Slot slot = Slot::full_range, double null_threshold = 1.0e-10) {
    check(mqcb200_build_df_tensor(handle_, static_cast<int>(slot), n_ao, naux, three, metric, null_threshold, half_out));
  }
  // the same GEMM fed nu-block by nu-block (three(nao*nao, naux) never has to exist)
  void whiten_begin(int n_ao, int naux_total, const double *half, int q_begin, int q_count, Slot slot = Slot::full_range) {
    check(mqcb200_whiten_begin(handle_, static_cast<int>(slot), n_ao, naux_total, q_begin, q_count, half));
  }
  void whiten_push(int nu_begin, int nu_count, const double *three_cols, long long ld_aux, Slot slot = Slot::full_range) {
    check(mqcb200_whiten_push(handle_, static_cast<int>(slot), nu_begin, nu_count, three_cols, ld_aux));
  }
  void whiten_end(Slot slot = Slot::full_range) { check(mqcb200_whiten_end(handle_, static_cast<int>(slot))); }

  // ---- build_fock_df(h, b, density, coeff, n_occ, fock, k_scale, j_scale) with b resident:
  // F = H + j_scale*J - (k_scale/2)*K; only coeff(:, 1:n_occ) is read, coeff has leading dimension ldc >= n
  void build_fock_df(const double *h, const double *density, const double *coeff, int ldc, int n_occ, double *fock,
                     double k_scale = 1.0, double j_scale = 1.0, Slot slot = Slot::full_range) {
    check(mqcb200_build_fock(handle_, static_cast<int>(slot), h, density, coeff, ldc, n_occ, k_scale, j_scale, fock));
  }
  // electronic_energy(h, fock, density) of the last build_fock_df, evaluated on the device
  double electronic_energy() const {
    double e = 0.0;
    check(mqcb200_last_energy(handle_, &e));
    return e;
  }
  // DF branch of assemble_fock (rhf.f90:1089-1106, energy :1207) without the range-separated second pass
  double assemble_fock(const double *h, const double *density, const double *coeff, int ldc, int n_occ, double *fock,
                       double k_scale = 1.0) {
    build_fock_df(h, density, coeff, ldc, n_occ, fock, k_scale, 1.0, Slot::full_range);
    return electronic_energy();
  }
  // J and K alone (either may be null); K carries the restricted factor 2 (rhf.f90:1637)
  void build_jk(const double *density, const double *coeff, int ldc, int n_occ, double *j, double *k,
                Slot slot = Slot::full_range) {
    check(mqcb200_build_jk(handle_, static_cast<int>(slot), density, coeff, ldc, n_occ, j, k));
  }
  // two-spin build: F_sigma = H + J[Da+Db] - k_scale*K[C_sigma]
  void build_fock_df_uhf(const double *h, const double *density_total, const double *coeff_a, int lda, int n_alpha,
                         const double *coeff_b, int ldb, int n_beta, double *fock_a, double *fock_b,
                         double k_scale = 1.0, Slot slot = Slot::full_range) {
    check(mqcb200_build_fock_uhf(handle_, static_cast<int>(slot), h, density_total, coeff_a, lda, n_alpha, coeff_b, ldb,
                                 n_beta, k_scale, fock_a, fock_b));
  }

  // ---- response_operator_df(b, x, c_occ, dtilde, g, k_scale) and fitted_potential_general(b, dens, g, k_scale)
  void response_operator_df(const double *x, int ldx, const double *c_occ, int ldc, int n_occ, const double *dtilde,
                            double *g, double k_scale = 1.0, Slot slot = Slot::full_range) {
    check(mqcb200_response_operator(handle_, static_cast<int>(slot), x, ldx, c_occ, ldc, n_occ, dtilde, k_scale, g));
  }
  void fitted_potential_general(const double *dens, double *g, double k_scale = 1.0, Slot slot = Slot::full_range) {
    check(mqcb200_fitted_potential_general(handle_, static_cast<int>(slot), dens, k_scale, g));
  }

  // ---- the contraction half of df_two_electron_gradient: gamma(nao, nao, naux), omega(naux, naux)
  void df_gradient_densities(const double *half, const double *total_density, const double *orbitals, int lda,
                             int n_occupied, double *gamma, double *omega, double exx_fraction = 1.0,
                             bool with_coulomb = true, const double *orbitals_beta = nullptr, int ldb = 0,
                             int n_occupied_beta = 0, bool unrestricted = false, Slot slot = Slot::full_range) {
    check(mqcb200_df_gradient_densities(handle_, static_cast<int>(slot), half, total_density, orbitals, lda, n_occupied,
                                        orbitals_beta, ldb, n_occupied_beta, unrestricted ? 1 : 0, exx_fraction,
                                        with_coulomb ? 1 : 0, gamma, omega));
  }

  // ---- run_libcint_rhf's loop on the GPU (closed shell; any size; also on a sharded tensor)
  // guess: 0 core Hamiltonian, 1 generalised Wolfsberg-Helmholz (the reference's default)
  RhfResult run_rhf(const double *hcore, const double *overlap, int n_ao, int n_electrons, int max_iter = 100,
                    double energy_tol = 1.0e-10, double density_tol = 1.0e-8, int diis_vectors = 8, int guess = 1,
                    double k_scale = 1.0, Slot slot = Slot::full_range) {
    RhfResult r;
    const std::size_t nn = static_cast<std::size_t>(n_ao) * static_cast<std::size_t>(n_ao);
    r.orbitals.assign(nn, 0.0);
    r.orbital_energies.assign(static_cast<std::size_t>(n_ao), 0.0);
    r.density.assign(nn, 0.0);
    r.e_history.assign(static_cast<std::size_t>(max_iter > 0 ? max_iter : 0), 0.0);
    int converged = 0;
    check(mqcb200_scf(handle_, static_cast<int>(slot), hcore, overlap, n_electrons, guess, max_iter, energy_tol,
                      density_tol, diis_vectors, k_scale, &r.electronic, &r.iterations, &converged, &r.n_mo,
                      r.orbitals.data(), r.orbital_energies.data(), r.density.data(), r.e_history.data()));
    r.converged = converged != 0;
    r.orbitals.resize(static_cast<std::size_t>(n_ao) * static_cast<std::size_t>(r.n_mo));
    r.orbital_energies.resize(static_cast<std::size_t>(r.n_mo));
    r.e_history.resize(static_cast<std::size_t>(r.iterations < max_iter ? r.iterations : max_iter));
    return r;
  }

  // ---- whole-molecule builds sharded over GPUs by auxiliary index
  static std::vector<char> comm_unique_id() {
    std::vector<char> id(128, 0);
    check(mqcb200_comm_unique_id(id.data()));
    return id;
  }
  void comm_init(int n_ranks, int rank, const std::vector<char> &id) {
    if (id.size() != 128) throw Error(MQCB200_FAIL, "mqcb200: the communicator id is 128 bytes");
    check(mqcb200_comm_init(handle_, n_ranks, rank, id.data()));
  }
  void comm_destroy() { check(mqcb200_comm_destroy(handle_)); }

 private:
  void *handle_ = nullptr;
};

// queue_t: FIFO of int64 fragment ids (queue_init_from_list / queue_pop / queue_is_empty / queue_destroy)
class WorkQueue {
 public:
  explicit WorkQueue(const std::vector<std::int64_t> &ids) {
    check(mqcb200_queue_create(ids.data(), static_cast<std::int64_t>(ids.size()), &queue_));
  }
  ~WorkQueue() {
    if (queue_) mqcb200_queue_destroy(queue_);
  }
  WorkQueue(const WorkQueue &) = delete;
  WorkQueue &operator=(const WorkQueue &) = delete;
  // queue_pop(queue, item_idx, has_item): false and id = -1 once drained
  bool pop(std::int64_t &id) {
    int has = 0;
    check(mqcb200_queue_pop(queue_, &id, &has));
    return has != 0;
  }
  bool is_empty() const {
    int empty = 0;
    check(mqcb200_queue_is_empty(queue_, &empty));
    return empty != 0;
  }

 private:
  void *queue_ = nullptr;
};

}  // namespace mqcb200

#endif  // MQCB200_HPP

!! Stand-in for the B200 Fock-build engine when the build has no libmqcb200
module mqc_b200_fock
   !! Same name and same entry points as the real module, declining -- the
   !! reference's own stub arrangement (src/methods/mqc_libcint_bridge_stub.f90,
   !! src/methods/mqc_cuest_bridge_stub.f90; selected in src/methods/CMakeLists.txt:21-48),
   !! so call sites need no preprocessor guard and a deck that names the `b200`
   !! backend on a build without it is refused with the build option, never silently
   !! run on the CPU (the refuse-don't-substitute rule, src/methods/mqc_method_hf.F90:190-196).
   use pic_types, only: dp
   use mqc_error, only: error_t, ERROR_VALIDATION
   implicit none
   private

   public :: b200_backend_available
   public :: b200_set_tensor, b200_clear_tensors
   public :: b200_build_fock_df, b200_electronic_energy
   public :: b200_build_fock_df_uhf
   public :: b200_build_df_tensor
   public :: b200_metric_inverse_sqrt, b200_whiten_begin, b200_whiten_push, b200_whiten_end
   public :: b200_set_tensor_shard, b200_comm_init, b200_comm_unique_id
   public :: b200_response_operator_df, b200_fitted_potential_general
   public :: b200_df_gradient_densities
   public :: b200_run_rhf_fragment
   public :: b200_finalize

   character(len=*), parameter :: REFUSAL = &
      "backend 'b200' was asked for, but this build has no B200 engine; configure with -DMQC_ENABLE_B200=ON"

contains

   pure function b200_backend_available() result(available)
      logical :: available
      available = .false.
   end function b200_backend_available

   subroutine b200_set_tensor(device_rank, bmat, n_ao, error, attenuated)
      integer, intent(in) :: device_rank
      real(dp), intent(in) :: bmat(:, :)
      integer, intent(in) :: n_ao
      type(error_t), intent(inout) :: error
      logical, intent(in), optional :: attenuated
      call error%set(ERROR_VALIDATION, REFUSAL)
      if (device_rank < 0 .or. n_ao < 0 .or. size(bmat) < 0 .or. present(attenuated)) return
   end subroutine b200_set_tensor

   subroutine b200_metric_inverse_sqrt(device_rank, metric, half, n_kept, error)
      integer, intent(in) :: device_rank
      real(dp), intent(in) :: metric(:, :)
      real(dp), intent(out) :: half(:, :)
      integer, intent(out) :: n_kept
      type(error_t), intent(inout) :: error
      half = 0.0_dp
      n_kept = 0
      call error%set(ERROR_VALIDATION, REFUSAL)
      if (device_rank < 0 .or. size(metric) < 0) return
   end subroutine b200_metric_inverse_sqrt

   subroutine b200_whiten_begin(device_rank, n_ao, naux_total, q_begin, q_count, half, error, attenuated)
      integer, intent(in) :: device_rank, n_ao, naux_total, q_begin, q_count
      real(dp), intent(in) :: half(:, :)
      type(error_t), intent(inout) :: error
      logical, intent(in), optional :: attenuated
      call error%set(ERROR_VALIDATION, REFUSAL)
      if (device_rank + n_ao + naux_total + q_begin + q_count + size(half) < 0 .or. present(attenuated)) return
   end subroutine b200_whiten_begin

   subroutine b200_whiten_push(nu_begin, three_block, n_ao, error, attenuated)
      integer, intent(in) :: nu_begin, n_ao
      real(dp), intent(in) :: three_block(:, :)
      type(error_t), intent(inout) :: error
      logical, intent(in), optional :: attenuated
      call error%set(ERROR_VALIDATION, REFUSAL)
      if (nu_begin + n_ao + size(three_block) < 0 .or. present(attenuated)) return
   end subroutine b200_whiten_push

   subroutine b200_whiten_end(error, attenuated)
      type(error_t), intent(inout) :: error
      logical, intent(in), optional :: attenuated
      call error%set(ERROR_VALIDATION, REFUSAL)
      if (present(attenuated)) return
   end subroutine b200_whiten_end

   subroutine b200_clear_tensors(error)
      type(error_t), intent(inout) :: error
      if (error%has_error()) return
   end subroutine b200_clear_tensors

   subroutine b200_build_fock_df(h, density, coeff, n_occ, fock, error, k_scale, j_scale, attenuated)
      real(dp), intent(in) :: h(:, :), density(:, :), coeff(:, :)
      integer, intent(in) :: n_occ
      real(dp), intent(out) :: fock(:, :)
      type(error_t), intent(inout) :: error
      real(dp), intent(in), optional :: k_scale, j_scale
      logical, intent(in), optional :: attenuated
      fock = 0.0_dp
      call error%set(ERROR_VALIDATION, REFUSAL)
      if (size(h) + size(density) + size(coeff) + n_occ < 0) return
      if (present(k_scale) .or. present(j_scale) .or. present(attenuated)) return
   end subroutine b200_build_fock_df

   ! Every further entry point of the real module declines the same way; the dummy arguments keep
   ! the real module's names, kinds and ranks so that call sites compile against either.
   subroutine b200_electronic_energy(e_elec, error)
      real(dp), intent(out) :: e_elec
      type(error_t), intent(inout) :: error
      e_elec = 0.0_dp
      call error%set(ERROR_VALIDATION, REFUSAL)
   end subroutine b200_electronic_energy

   subroutine b200_build_fock_df_uhf(h, density_total, coeff_a, n_alpha, coeff_b, n_beta, fock_a, fock_b, error, k_scale)
      real(dp), intent(in) :: h(:, :), density_total(:, :), coeff_a(:, :), coeff_b(:, :)
      integer, intent(in) :: n_alpha, n_beta
      real(dp), intent(out) :: fock_a(:, :), fock_b(:, :)
      type(error_t), intent(inout) :: error
      real(dp), intent(in), optional :: k_scale
      fock_a = 0.0_dp
      fock_b = 0.0_dp
      call error%set(ERROR_VALIDATION, REFUSAL)
      if (size(h) + size(density_total) + size(coeff_a) + size(coeff_b) + n_alpha + n_beta < 0 .or. present(k_scale)) return
   end subroutine b200_build_fock_df_uhf

   subroutine b200_build_df_tensor(device_rank, three, metric, n_ao, half, error, attenuated)
      integer, intent(in) :: device_rank, n_ao
      real(dp), intent(in) :: three(:, :), metric(:, :)
      real(dp), intent(out) :: half(:, :)
      type(error_t), intent(inout) :: error
      logical, intent(in), optional :: attenuated
      half = 0.0_dp
      call error%set(ERROR_VALIDATION, REFUSAL)
      if (device_rank + n_ao + size(three) + size(metric) < 0 .or. present(attenuated)) return
   end subroutine b200_build_df_tensor

   subroutine b200_set_tensor_shard(device_rank, bmat_shard, n_ao, naux_total, q_begin, error)
      integer, intent(in) :: device_rank, n_ao, naux_total, q_begin
      real(dp), intent(in) :: bmat_shard(:, :)
      type(error_t), intent(inout) :: error
      call error%set(ERROR_VALIDATION, REFUSAL)
      if (device_rank + n_ao + naux_total + q_begin + size(bmat_shard) < 0) return
   end subroutine b200_set_tensor_shard

   subroutine b200_comm_unique_id(id, error)
      character(len=1), intent(out) :: id(128)
      type(error_t), intent(inout) :: error
      id = " "
      call error%set(ERROR_VALIDATION, REFUSAL)
   end subroutine b200_comm_unique_id

   subroutine b200_comm_init(device_rank, n_ranks, rank, id, error)
      integer, intent(in) :: device_rank, n_ranks, rank
      character(len=1), intent(in) :: id(128)
      type(error_t), intent(inout) :: error
      call error%set(ERROR_VALIDATION, REFUSAL)
      if (device_rank + n_ranks + rank + size(id) < 0) return
   end subroutine b200_comm_init

   subroutine b200_response_operator_df(x, c_occ, dtilde, g, error, k_scale)
      real(dp), intent(in) :: x(:, :), c_occ(:, :), dtilde(:, :)
      real(dp), intent(out) :: g(:, :)
      type(error_t), intent(inout) :: error
      real(dp), intent(in), optional :: k_scale
      g = 0.0_dp
      call error%set(ERROR_VALIDATION, REFUSAL)
      if (size(x) + size(c_occ) + size(dtilde) < 0 .or. present(k_scale)) return
   end subroutine b200_response_operator_df

   subroutine b200_fitted_potential_general(dens, g, error, k_scale)
      real(dp), intent(in) :: dens(:, :)
      real(dp), intent(out) :: g(:, :)
      type(error_t), intent(inout) :: error
      real(dp), intent(in), optional :: k_scale
      g = 0.0_dp
      call error%set(ERROR_VALIDATION, REFUSAL)
      if (size(dens) < 0 .or. present(k_scale)) return
   end subroutine b200_fitted_potential_general

   subroutine b200_df_gradient_densities(half, total_density, orbitals, n_occupied, gamma, omega, error, &
                                         orbitals_beta, n_occupied_beta, exx_fraction, with_coulomb)
      real(dp), intent(in) :: half(:, :), total_density(:, :), orbitals(:, :)
      integer, intent(in) :: n_occupied
      real(dp), intent(out) :: gamma(:, :, :), omega(:, :)
      type(error_t), intent(inout) :: error
      real(dp), intent(in), optional :: orbitals_beta(:, :)
      integer, intent(in), optional :: n_occupied_beta
      real(dp), intent(in), optional :: exx_fraction
      logical, intent(in), optional :: with_coulomb
      gamma = 0.0_dp
      omega = 0.0_dp
      call error%set(ERROR_VALIDATION, REFUSAL)
      if (size(half) + size(total_density) + size(orbitals) + n_occupied < 0) return
      if (present(orbitals_beta) .or. present(n_occupied_beta) .or. present(exx_fraction) .or. present(with_coulomb)) return
   end subroutine b200_df_gradient_densities

   subroutine b200_run_rhf_fragment(h, overlap, nelec, max_iter, energy_tol, density_tol, electronic, iterations, &
                                    converged, orbitals, orbital_energies, density, error, diis_vectors, k_scale)
      real(dp), intent(in) :: h(:, :), overlap(:, :)
      integer, intent(in) :: nelec, max_iter
      real(dp), intent(in) :: energy_tol, density_tol
      real(dp), intent(out) :: electronic
      integer, intent(out) :: iterations
      logical, intent(out) :: converged
      real(dp), intent(out) :: orbitals(:, :), orbital_energies(:), density(:, :)
      type(error_t), intent(inout) :: error
      integer, intent(in), optional :: diis_vectors
      real(dp), intent(in), optional :: k_scale
      electronic = 0.0_dp
      iterations = 0
      converged = .false.
      orbitals = 0.0_dp
      orbital_energies = 0.0_dp
      density = 0.0_dp
      call error%set(ERROR_VALIDATION, REFUSAL)
      if (size(h) + size(overlap) + nelec + max_iter < 0 .or. energy_tol + density_tol < 0.0_dp) return
      if (present(diis_vectors) .or. present(k_scale)) return
   end subroutine b200_run_rhf_fragment

   subroutine b200_finalize()
   end subroutine b200_finalize

end module mqc_b200_fock

! ============================================================================
!  mqc_b200_iface.f90 -- Fortran 2008 iso_c_binding interface to libmqcb200.so,
!  the B200-native density-fitted J/K Fock-build engine (include/mqcb200.h).
!
!  Written in the style of the reference's own hand-written bindings
!  (backends/cuest/bindings/cublas.f90:20-44): every function returns
!  INTEGER(c_int); handles are TYPE(c_ptr) (VALUE when passed in, INTENT(OUT) when
!  created); sizes are INTEGER(c_int), VALUE; matrices are REAL(c_double)
!  assumed-size arrays, which is Fortran's column-major layout already -- no
!  transposes anywhere.
!
!  NOT COMPILED IN THIS REPOSITORY: the build image has no Fortran compiler.  The
!  C header it mirrors IS compiled and tested (tests/test_abi.py checks every symbol),
!  and tests/test_fortran_binding.py reads this file against the header prototype by
!  prototype: every symbol present, argument count, order and names, VALUE on every C
!  scalar and on no array, kinds (c_int / c_double / c_size_t / c_int64_t), constants.
!
!  Build:  gfortran -c mqc_b200_iface.f90        Link:  ... -lmqcb200
! ============================================================================
module mqc_b200_iface
   use, intrinsic :: iso_c_binding, only: c_ptr, c_int, c_double, c_char, c_size_t, c_int64_t, c_null_ptr
   implicit none
   private

   public :: mqcb200_create, mqcb200_destroy, mqcb200_last_error, mqcb200_version
   public :: mqcb200_set_workspace_limit
   public :: mqcb200_set_tensor, mqcb200_set_tensor_shard, mqcb200_set_tensor_from_3c, mqcb200_clear_tensor
   public :: mqcb200_build_g_two_factor
   public :: mqcb200_build_fock, mqcb200_build_jk, mqcb200_build_jk_uhf, mqcb200_build_fock_uhf
   public :: mqcb200_last_energy
   public :: mqcb200_build_fock_device, mqcb200_build_fock_uhf_device, mqcb200_get_stream, mqcb200_tensor_bytes
   public :: mqcb200_comm_unique_id, mqcb200_comm_init, mqcb200_comm_destroy
   public :: mqcb200_queue_create, mqcb200_queue_pop, mqcb200_queue_is_empty, mqcb200_queue_destroy
   public :: mqcb200_tensor_shape, mqcb200_response_operator, mqcb200_fitted_potential_general
   public :: mqcb200_scf, mqcb200_scf_fragment, mqcb200_scf_fragment_batch, mqcb200_df_gradient_densities
   public :: mqcb200_metric_inverse_sqrt, mqcb200_build_df_tensor
   public :: mqcb200_whiten_begin, mqcb200_whiten_push, mqcb200_whiten_end
   public :: mqcb200_set_fuse_threshold, mqcb200_set_overlap, mqcb200_set_profiling, mqcb200_set_scf_check_every
   public :: mqcb200_last_timings, mqcb200_last_launches, mqcb200_last_gamma_fused, mqcb200_last_set_tensor
   public :: mqcb200_last_whiten, mqcb200_last_metric, mqcb200_synth_tensor
   public :: MQCB200_OK, MQCB200_FAIL, MQCB200_BAD_HANDLE, MQCB200_NUM_TIMERS
   public :: MQCB200_SLOT_FULL_RANGE, MQCB200_SLOT_ATTENUATED

   integer(c_int), parameter :: MQCB200_OK = 0          !! same values as MQC_OK/MQC_FAIL/MQC_BAD_HANDLE,
   integer(c_int), parameter :: MQCB200_FAIL = 1        !! src/interface/mqc_capi_status.f90:21-23
   integer(c_int), parameter :: MQCB200_BAD_HANDLE = 2
   integer(c_int), parameter :: MQCB200_SLOT_FULL_RANGE = 0   !! bmat
   integer(c_int), parameter :: MQCB200_SLOT_ATTENUATED = 1   !! bmat_lr
   integer(c_int), parameter :: MQCB200_NUM_TIMERS = 8        !! length of mqcb200_last_timings' array

   interface
      function mqcb200_create(device_rank, handle) bind(C, name="mqcb200_create") result(status)
         import :: c_int, c_ptr
         integer(c_int), value :: device_rank
         type(c_ptr), intent(out) :: handle
         integer(c_int) :: status
      end function
      function mqcb200_destroy(handle) bind(C, name="mqcb200_destroy") result(status)
         import :: c_int, c_ptr
         type(c_ptr), value :: handle
         integer(c_int) :: status
      end function
      subroutine mqcb200_last_error(buffer_len, buffer) bind(C, name="mqcb200_last_error")
         import :: c_int, c_char
         integer(c_int), value :: buffer_len
         character(kind=c_char) :: buffer(*)
      end subroutine
      function mqcb200_version() bind(C, name="mqcb200_version") result(version)
         import :: c_int
         integer(c_int) :: version
      end function
      function mqcb200_set_workspace_limit(handle, bytes) bind(C, name="mqcb200_set_workspace_limit") result(status)
         import :: c_int, c_ptr, c_size_t
         type(c_ptr), value :: handle
         integer(c_size_t), value :: bytes
         integer(c_int) :: status
      end function
      function mqcb200_set_tensor(handle, slot, n, naux, b) bind(C, name="mqcb200_set_tensor") result(status)
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: handle
         integer(c_int), value :: slot, n, naux
         real(c_double), intent(in) :: b(*)          !! bmat(nao*nao, naux)
         integer(c_int) :: status
      end function
      function mqcb200_set_tensor_shard(handle, slot, n, naux_total, q_begin, q_count, b_shard) &
         bind(C, name="mqcb200_set_tensor_shard") result(status)
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: handle
         integer(c_int), value :: slot, n, naux_total, q_begin, q_count
         real(c_double), intent(in) :: b_shard(*)
         integer(c_int) :: status
      end function
      function mqcb200_set_tensor_from_3c(handle, slot, n, naux, three, half) &
         bind(C, name="mqcb200_set_tensor_from_3c") result(status)
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: handle
         integer(c_int), value :: slot, n, naux
         real(c_double), intent(in) :: three(*)      !! (nao*nao, naux), build_df_tensor's `three`
         real(c_double), intent(in) :: half(*)       !! (naux, naux), metric_inverse_sqrt's result
         integer(c_int) :: status
      end function
      function mqcb200_build_g_two_factor(handle, slot, density, coeff_a, lda, n_a, coeff_b, ldb, n_b, ka, kb, g) &
         bind(C, name="mqcb200_build_g_two_factor") result(status)
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: handle
         integer(c_int), value :: slot, lda, n_a, ldb, n_b
         real(c_double), intent(in) :: density(*), coeff_a(*), coeff_b(*)
         real(c_double), value :: ka, kb
         real(c_double), intent(out) :: g(*)
         integer(c_int) :: status
      end function
      function mqcb200_clear_tensor(handle, slot) bind(C, name="mqcb200_clear_tensor") result(status)
         import :: c_int, c_ptr
         type(c_ptr), value :: handle
         integer(c_int), value :: slot
         integer(c_int) :: status
      end function
      function mqcb200_build_fock(handle, slot, h, density, coeff, ldc, n_occ, k_scale, j_scale, fock) &
         bind(C, name="mqcb200_build_fock") result(status)
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: handle
         integer(c_int), value :: slot, ldc, n_occ
         real(c_double), intent(in) :: h(*), density(*), coeff(*)
         real(c_double), value :: k_scale, j_scale
         real(c_double), intent(out) :: fock(*)
         integer(c_int) :: status
      end function
      !! Device-operand builds: h, density, coeff and fock are DEVICE addresses (type(c_ptr), value -- the style of
      !! the reference's own GPU bindings, backends/cuest/bindings/cublas.f90:20-44), contiguous column-major,
      !! queued on the engine's stream; sync = 0 returns at once.  For a GPU-resident SCF driver of the kind
      !! backends/cuest/backend/mqc_cuest_scf.f90:444-553 runs.
      function mqcb200_build_fock_device(handle, slot, d_h, d_density, d_coeff, n_occ, k_scale, j_scale, d_fock, sync) &
         bind(C, name="mqcb200_build_fock_device") result(status)
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: handle, d_h, d_density, d_coeff, d_fock
         integer(c_int), value :: slot, n_occ, sync
         real(c_double), value :: k_scale, j_scale
         integer(c_int) :: status
      end function
      function mqcb200_build_fock_uhf_device(handle, slot, d_h, d_density_total, d_coeff_a, n_alpha, d_coeff_b, n_beta, &
                                             k_scale, d_fock_a, d_fock_b, sync) &
         bind(C, name="mqcb200_build_fock_uhf_device") result(status)
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: handle, d_h, d_density_total, d_coeff_a, d_coeff_b, d_fock_a, d_fock_b
         integer(c_int), value :: slot, n_alpha, n_beta, sync
         real(c_double), value :: k_scale
         integer(c_int) :: status
      end function
      !! the CUDA stream the engine queues on (a cudaStream_t), so that a caller's own kernels can be ordered with it
      function mqcb200_get_stream(handle, stream) bind(C, name="mqcb200_get_stream") result(status)
         import :: c_int, c_ptr
         type(c_ptr), value :: handle
         type(c_ptr), intent(out) :: stream
         integer(c_int) :: status
      end function
      function mqcb200_tensor_bytes(handle, slot, bytes) bind(C, name="mqcb200_tensor_bytes") result(status)
         import :: c_int, c_ptr, c_size_t
         type(c_ptr), value :: handle
         integer(c_int), value :: slot
         integer(c_size_t), intent(out) :: bytes
         integer(c_int) :: status
      end function
      function mqcb200_build_jk(handle, slot, density, coeff, ldc, n_occ, j, k) &
         bind(C, name="mqcb200_build_jk") result(status)
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: handle
         integer(c_int), value :: slot, ldc, n_occ
         real(c_double), intent(in) :: density(*), coeff(*)
         real(c_double), intent(out) :: j(*), k(*)
         integer(c_int) :: status
      end function
      function mqcb200_build_jk_uhf(handle, slot, density_total, coeff_a, lda, n_alpha, coeff_b, ldb, n_beta, &
                                    j, k_alpha, k_beta) bind(C, name="mqcb200_build_jk_uhf") result(status)
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: handle
         integer(c_int), value :: slot, lda, n_alpha, ldb, n_beta
         real(c_double), intent(in) :: density_total(*), coeff_a(*), coeff_b(*)
         real(c_double), intent(out) :: j(*), k_alpha(*), k_beta(*)
         integer(c_int) :: status
      end function
      function mqcb200_build_fock_uhf(handle, slot, h, density_total, coeff_a, lda, n_alpha, coeff_b, ldb, &
                                      n_beta, k_scale, fock_a, fock_b) &
         bind(C, name="mqcb200_build_fock_uhf") result(status)
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: handle
         integer(c_int), value :: slot, lda, n_alpha, ldb, n_beta
         real(c_double), intent(in) :: h(*), density_total(*), coeff_a(*), coeff_b(*)
         real(c_double), value :: k_scale
         real(c_double), intent(out) :: fock_a(*), fock_b(*)
         integer(c_int) :: status
      end function
      function mqcb200_last_energy(handle, e_elec) bind(C, name="mqcb200_last_energy") result(status)
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: handle
         real(c_double), intent(out) :: e_elec
         integer(c_int) :: status
      end function
      function mqcb200_comm_unique_id(id) bind(C, name="mqcb200_comm_unique_id") result(status)
         import :: c_int, c_char
         character(kind=c_char), intent(out) :: id(128)
         integer(c_int) :: status
      end function
      function mqcb200_comm_init(handle, n_ranks, rank, id) bind(C, name="mqcb200_comm_init") result(status)
         import :: c_int, c_ptr, c_char
         type(c_ptr), value :: handle
         integer(c_int), value :: n_ranks, rank
         character(kind=c_char), intent(in) :: id(128)
         integer(c_int) :: status
      end function
      function mqcb200_comm_destroy(handle) bind(C, name="mqcb200_comm_destroy") result(status)
         import :: c_int, c_ptr
         type(c_ptr), value :: handle
         integer(c_int) :: status
      end function
      function mqcb200_queue_create(ids, count, queue) bind(C, name="mqcb200_queue_create") result(status)
         import :: c_int, c_ptr, c_int64_t
         integer(c_int64_t), intent(in) :: ids(*)
         integer(c_int64_t), value :: count
         type(c_ptr), intent(out) :: queue
         integer(c_int) :: status
      end function
      function mqcb200_queue_pop(queue, id, has_item) bind(C, name="mqcb200_queue_pop") result(status)
         import :: c_int, c_ptr, c_int64_t
         type(c_ptr), value :: queue
         integer(c_int64_t), intent(out) :: id
         integer(c_int), intent(out) :: has_item
         integer(c_int) :: status
      end function
      function mqcb200_queue_is_empty(queue, is_empty) bind(C, name="mqcb200_queue_is_empty") result(status)
         import :: c_int, c_ptr
         type(c_ptr), value :: queue
         integer(c_int), intent(out) :: is_empty
         integer(c_int) :: status
      end function
      function mqcb200_queue_destroy(queue) bind(C, name="mqcb200_queue_destroy") result(status)
         import :: c_int, c_ptr
         type(c_ptr), value :: queue
         integer(c_int) :: status
      end function
      function mqcb200_tensor_shape(handle, slot, n, naux_total, q_begin, q_count) &
         bind(C, name="mqcb200_tensor_shape") result(status)
         import :: c_int, c_ptr
         type(c_ptr), value :: handle
         integer(c_int), value :: slot
         integer(c_int), intent(out) :: n, naux_total, q_begin, q_count
         integer(c_int) :: status
      end function
      function mqcb200_response_operator(handle, slot, x, ldx, c_occ, ldc, n_occ, dtilde, k_scale, g) &
         bind(C, name="mqcb200_response_operator") result(status)
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: handle
         integer(c_int), value :: slot, ldx, ldc, n_occ
         real(c_double), intent(in) :: x(*), c_occ(*), dtilde(*)
         real(c_double), value :: k_scale
         real(c_double), intent(out) :: g(*)
         integer(c_int) :: status
      end function
      function mqcb200_fitted_potential_general(handle, slot, dens, k_scale, g) &
         bind(C, name="mqcb200_fitted_potential_general") result(status)
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: handle
         integer(c_int), value :: slot
         real(c_double), intent(in) :: dens(*)
         real(c_double), value :: k_scale
         real(c_double), intent(out) :: g(*)
         integer(c_int) :: status
      end function
      function mqcb200_scf_fragment(handle, slot, hcore, overlap, n_electrons, guess, max_iter, energy_tol, &
                                    density_tol, diis_vectors, k_scale, e_electronic, iterations, converged, &
                                    n_mo, coeff, orbital_energies, density, e_history) &
         bind(C, name="mqcb200_scf_fragment") result(status)
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: handle
         integer(c_int), value :: slot, n_electrons, guess, max_iter, diis_vectors
         real(c_double), intent(in) :: hcore(*), overlap(*)
         real(c_double), value :: energy_tol, density_tol, k_scale
         real(c_double), intent(out) :: e_electronic
         integer(c_int), intent(out) :: iterations, converged, n_mo
         real(c_double), intent(out) :: coeff(*), orbital_energies(*), density(*), e_history(*)
         integer(c_int) :: status
      end function
      function mqcb200_df_gradient_densities(handle, slot, half, total_density, orbitals, lda, n_occupied, &
                                             orbitals_beta, ldb, n_occupied_beta, unrestricted, exx_fraction, &
                                             with_coulomb, gamma, omega) &
         bind(C, name="mqcb200_df_gradient_densities") result(status)
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: handle
         integer(c_int), value :: slot, lda, n_occupied, ldb, n_occupied_beta, unrestricted, with_coulomb
         real(c_double), intent(in) :: half(*), total_density(*), orbitals(*), orbitals_beta(*)
         real(c_double), value :: exx_fraction
         real(c_double), intent(out) :: gamma(*), omega(*)
         integer(c_int) :: status
      end function
      function mqcb200_metric_inverse_sqrt(handle, naux, metric, null_threshold, half, n_kept) &
         bind(C, name="mqcb200_metric_inverse_sqrt") result(status)
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: handle
         integer(c_int), value :: naux
         real(c_double), intent(in) :: metric(*)
         real(c_double), value :: null_threshold
         real(c_double), intent(out) :: half(*)
         integer(c_int), intent(out) :: n_kept
         integer(c_int) :: status
      end function
      function mqcb200_build_df_tensor(handle, slot, n, naux, three, metric, null_threshold, half_out) &
         bind(C, name="mqcb200_build_df_tensor") result(status)
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: handle
         integer(c_int), value :: slot, n, naux
         real(c_double), intent(in) :: three(*), metric(*)
         real(c_double), value :: null_threshold
         real(c_double), intent(out) :: half_out(*)
         integer(c_int) :: status
      end function
      function mqcb200_whiten_begin(handle, slot, n, naux_total, q_begin, q_count, half) &
         bind(C, name="mqcb200_whiten_begin") result(status)
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: handle
         integer(c_int), value :: slot, n, naux_total, q_begin, q_count
         real(c_double), intent(in) :: half(*)
         integer(c_int) :: status
      end function
      function mqcb200_whiten_push(handle, slot, nu_begin, nu_count, three_cols, ld_aux) &
         bind(C, name="mqcb200_whiten_push") result(status)
         import :: c_int, c_ptr, c_double, c_int64_t
         type(c_ptr), value :: handle
         integer(c_int), value :: slot, nu_begin, nu_count
         real(c_double), intent(in) :: three_cols(*)
         integer(c_int64_t), value :: ld_aux
         integer(c_int) :: status
      end function
      function mqcb200_whiten_end(handle, slot) bind(C, name="mqcb200_whiten_end") result(status)
         import :: c_int, c_ptr
         type(c_ptr), value :: handle
         integer(c_int), value :: slot
         integer(c_int) :: status
      end function
      function mqcb200_scf_fragment_batch(handle, slot, n_fragments, hcore_all, overlap_all, n_electrons, guess, &
                                          max_iter, energy_tol, density_tol, diis_vectors, k_scale, e_electronic, &
                                          iterations, converged, n_mo, coeff_all, orbital_energies_all, density_all) &
         bind(C, name="mqcb200_scf_fragment_batch") result(status)
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: handle
         integer(c_int), value :: slot, n_fragments, n_electrons, guess, max_iter, diis_vectors
         real(c_double), intent(in) :: hcore_all(*), overlap_all(*)
         real(c_double), value :: energy_tol, density_tol, k_scale
         real(c_double), intent(out) :: e_electronic(*)
         integer(c_int), intent(out) :: iterations(*), converged(*), n_mo(*)
         real(c_double), intent(out) :: coeff_all(*), orbital_energies_all(*), density_all(*)
         integer(c_int) :: status
      end function
      function mqcb200_scf(handle, slot, hcore, overlap, n_electrons, guess, max_iter, energy_tol, &
                           density_tol, diis_vectors, k_scale, e_electronic, iterations, converged, &
                           n_mo, coeff, orbital_energies, density, e_history) &
         bind(C, name="mqcb200_scf") result(status)
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: handle
         integer(c_int), value :: slot, n_electrons, guess, max_iter, diis_vectors
         real(c_double), intent(in) :: hcore(*), overlap(*)
         real(c_double), value :: energy_tol, density_tol, k_scale
         real(c_double), intent(out) :: e_electronic
         integer(c_int), intent(out) :: iterations, converged, n_mo
         real(c_double), intent(out) :: coeff(*), orbital_energies(*), density(*), e_history(*)
         integer(c_int) :: status
      end function
      ! ---- tuning and diagnostics (mqcb200.h: the "instrumentation" block and the setters) ----
      function mqcb200_set_fuse_threshold(handle, bytes) bind(C, name="mqcb200_set_fuse_threshold") result(status)
         import :: c_int, c_ptr, c_size_t
         type(c_ptr), value :: handle
         integer(c_size_t), value :: bytes
         integer(c_int) :: status
      end function
      function mqcb200_set_overlap(handle, on) bind(C, name="mqcb200_set_overlap") result(status)
         import :: c_int, c_ptr
         type(c_ptr), value :: handle
         integer(c_int), value :: on
         integer(c_int) :: status
      end function
      function mqcb200_set_profiling(handle, on) bind(C, name="mqcb200_set_profiling") result(status)
         import :: c_int, c_ptr
         type(c_ptr), value :: handle
         integer(c_int), value :: on
         integer(c_int) :: status
      end function
      function mqcb200_set_scf_check_every(handle, iterations) bind(C, name="mqcb200_set_scf_check_every") result(status)
         import :: c_int, c_ptr
         type(c_ptr), value :: handle
         integer(c_int), value :: iterations
         integer(c_int) :: status
      end function
      function mqcb200_last_timings(handle, ms) bind(C, name="mqcb200_last_timings") result(status)
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: handle
         real(c_double), intent(out) :: ms(*)        !! MQCB200_NUM_TIMERS phase sums in milliseconds
         integer(c_int) :: status
      end function
      function mqcb200_last_launches(handle, n_kernels) bind(C, name="mqcb200_last_launches") result(status)
         import :: c_int, c_ptr
         type(c_ptr), value :: handle
         integer(c_int), intent(out) :: n_kernels
         integer(c_int) :: status
      end function
      function mqcb200_last_gamma_fused(handle, fused) bind(C, name="mqcb200_last_gamma_fused") result(status)
         import :: c_int, c_ptr
         type(c_ptr), value :: handle
         integer(c_int), intent(out) :: fused
         integer(c_int) :: status
      end function
      function mqcb200_last_set_tensor(handle, ms, h2d_bytes) bind(C, name="mqcb200_last_set_tensor") result(status)
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: handle
         real(c_double), intent(out) :: ms, h2d_bytes
         integer(c_int) :: status
      end function
      function mqcb200_last_whiten(handle, ms, flops) bind(C, name="mqcb200_last_whiten") result(status)
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: handle
         real(c_double), intent(out) :: ms, flops
         integer(c_int) :: status
      end function
      function mqcb200_last_metric(handle, ms, sweeps) bind(C, name="mqcb200_last_metric") result(status)
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: handle
         real(c_double), intent(out) :: ms
         integer(c_int), intent(out) :: sweeps
         integer(c_int) :: status
      end function
      function mqcb200_synth_tensor(handle, slot, n, naux_total, q_begin, q_count, seed, scale) &
         bind(C, name="mqcb200_synth_tensor") result(status)
         import :: c_int, c_ptr, c_double, c_int64_t
         type(c_ptr), value :: handle
         integer(c_int), value :: slot, n, naux_total, q_begin, q_count
         integer(c_int64_t), value :: seed           !! uint64_t on the C side: the bit pattern is what counts
         real(c_double), value :: scale
         integer(c_int) :: status
      end function
   end interface

end module mqc_b200_iface

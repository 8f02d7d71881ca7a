!! B200 engine for the density-fitted Fock build: the drop-in for build_fock_df
module mqc_b200_fock
   !! Same arguments, same meaning, same error behaviour as the CPU routine it
   !! replaces -- `build_fock_df(h, b, density, coeff, n_occ, fock, k_scale, j_scale)`
   !! in backends/libcint/mqc_libcint_rhf.f90:1576-1646 -- except that the tensor
   !! is handed over once per geometry (`b200_set_tensor`, where run_libcint_rhf
   !! has just called build_df_tensor, :486-498) instead of on every call.
   !!
   !! One engine handle per process, created lazily and kept for the life of the
   !! process, bound to `mod(device_rank, device_count)`: the same arrangement as
   !! the reference's GPU backend (shared_context, get_cuest_context,
   !! backends/cuest/backend/mqc_cuest_context.f90:134-138, :285-299), because method
   !! objects are created and destroyed per fragment and cannot hold device state.
   !!
   !! Failures never stop: they set `error` (ERROR_VALIDATION) with the engine's
   !! message and return, as every routine of the CPU backend does.
   !!
   !! NOT COMPILED IN THIS REPOSITORY (no Fortran compiler in the build image); the
   !! twin `mqc_b200_fock_stub.f90` is what a build without the engine compiles.
   use, intrinsic :: iso_c_binding, only: c_ptr, c_null_ptr, c_associated, c_int, c_double, c_char, c_null_char, &
                                          c_int64_t
   use pic_types, only: dp
   use mqc_error, only: error_t, ERROR_VALIDATION
   use mqc_b200_iface
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

   type(c_ptr), save :: shared_handle = c_null_ptr
   integer, save :: shared_device_rank = -1

contains

   pure function b200_backend_available() result(available)
      logical :: available
      available = .true.
   end function b200_backend_available

   subroutine get_engine(device_rank, handle, error)
      integer, intent(in) :: device_rank
      type(c_ptr), intent(out) :: handle
      type(error_t), intent(inout) :: error

      handle = c_null_ptr
      if (c_associated(shared_handle)) then
         if (device_rank /= shared_device_rank) then
            call error%set(ERROR_VALIDATION, "b200: this process is already bound to another device")
            return
         end if
         handle = shared_handle
         return
      end if
      if (mqcb200_create(int(device_rank, c_int), shared_handle) /= MQCB200_OK) then
         call engine_failure("b200: engine creation failed", error)
         shared_handle = c_null_ptr
         return
      end if
      shared_device_rank = device_rank
      handle = shared_handle
   end subroutine get_engine

   subroutine b200_set_tensor(device_rank, bmat, n_ao, error, attenuated)
      !! Hand the fitted tensor `bmat(nao*nao, naux)` to the device, once per geometry
      integer, intent(in) :: device_rank
      real(dp), intent(in), contiguous :: bmat(:, :)
      integer, intent(in) :: n_ao
      type(error_t), intent(inout) :: error
      logical, intent(in), optional :: attenuated   !! .true. for bmat_lr

      type(c_ptr) :: handle
      integer(c_int) :: slot

      call get_engine(device_rank, handle, error)
      if (error%has_error()) return
      slot = MQCB200_SLOT_FULL_RANGE
      if (present(attenuated)) then
         if (attenuated) slot = MQCB200_SLOT_ATTENUATED
      end if
      if (size(bmat, 1) /= n_ao*n_ao) then
         call error%set(ERROR_VALIDATION, "b200: the fitted tensor is not (nao*nao, naux)")
         return
      end if
      if (mqcb200_set_tensor(handle, slot, int(n_ao, c_int), int(size(bmat, 2), c_int), bmat) /= MQCB200_OK) then
         call engine_failure("b200: set_tensor", error)
      end if
   end subroutine b200_set_tensor

   subroutine b200_clear_tensors(error)
      type(error_t), intent(inout) :: error
      if (.not. c_associated(shared_handle)) return
      if (mqcb200_clear_tensor(shared_handle, MQCB200_SLOT_FULL_RANGE) /= MQCB200_OK) &
         call engine_failure("b200: clear_tensor", error)
      if (mqcb200_clear_tensor(shared_handle, MQCB200_SLOT_ATTENUATED) /= MQCB200_OK) &
         call engine_failure("b200: clear_tensor", error)
   end subroutine b200_clear_tensors

   subroutine b200_build_fock_df(h, density, coeff, n_occ, fock, error, k_scale, j_scale, attenuated)
      !! F = H + j_scale*J - (k_scale/2)*K from the resident tensor
      real(dp), intent(in), contiguous :: h(:, :), density(:, :)
      real(dp), intent(in), contiguous :: coeff(:, :)   !! only (:, 1:n_occ) is read
      integer, intent(in) :: n_occ
      real(dp), intent(out), contiguous :: fock(:, :)
      type(error_t), intent(inout) :: error
      real(dp), intent(in), optional :: k_scale, j_scale
      logical, intent(in), optional :: attenuated

      real(c_double) :: kf, jf
      integer(c_int) :: slot

      if (.not. c_associated(shared_handle)) then
         call error%set(ERROR_VALIDATION, "b200: build_fock_df called before b200_set_tensor")
         return
      end if
      kf = 1.0_c_double
      if (present(k_scale)) kf = k_scale
      jf = 1.0_c_double
      if (present(j_scale)) jf = j_scale
      slot = MQCB200_SLOT_FULL_RANGE
      if (present(attenuated)) then
         if (attenuated) slot = MQCB200_SLOT_ATTENUATED
      end if
      if (.not. operands_match_tensor(slot, size(h, 1), error)) return
      if (mqcb200_build_fock(shared_handle, slot, h, density, coeff, int(size(coeff, 1), c_int), &
                             int(n_occ, c_int), kf, jf, fock) /= MQCB200_OK) then
         call engine_failure("b200: build_fock", error)
      end if
   end subroutine b200_build_fock_df

   function operands_match_tensor(slot, n_operand, error) result(ok)
      !! The build calls carry no size: refuse operands that do not belong to the resident
      !! tensor (a stale tensor of the previous fragment would be read out of bounds)
      integer(c_int), intent(in) :: slot
      integer, intent(in) :: n_operand
      type(error_t), intent(inout) :: error
      logical :: ok
      integer(c_int) :: n, naux_total, q_begin, q_count

      ok = .false.
      if (mqcb200_tensor_shape(shared_handle, slot, n, naux_total, q_begin, q_count) /= MQCB200_OK) then
         call engine_failure("b200: tensor_shape", error)
         return
      end if
      if (n == 0) then
         call error%set(ERROR_VALIDATION, "b200: no fitted tensor is resident on this slot")
         return
      end if
      if (int(n) /= n_operand) then
         call error%set(ERROR_VALIDATION, "b200: the operands do not have the size of the resident tensor")
         return
      end if
      ok = .true.
   end function operands_match_tensor

   subroutine b200_electronic_energy(e_elec, error)
      !! E = 1/2 sum D (H + F) of the last b200_build_fock_df (electronic_energy,
      !! mqc_libcint_rhf.f90:1691-1697), already on the host: no second pass over the matrices
      real(dp), intent(out) :: e_elec
      type(error_t), intent(inout) :: error
      real(c_double) :: e
      e_elec = 0.0_dp
      if (mqcb200_last_energy(shared_handle, e) /= MQCB200_OK) then
         call engine_failure("b200: last_energy", error)
         return
      end if
      e_elec = e
   end subroutine b200_electronic_energy

   subroutine b200_build_fock_df_uhf(h, density_total, coeff_a, n_alpha, coeff_b, n_beta, fock_a, fock_b, error, k_scale)
      !! F_sigma = H + J[Da+Db] - k_scale*K[C_sigma]: the fitted twin of build_fock_uhf
      !! (mqc_libcint_rhf.f90:1648-1681), which the CPU backend refuses (mqc_libcint_bridge.f90:605-612)
      real(dp), intent(in), contiguous :: h(:, :), density_total(:, :), coeff_a(:, :), coeff_b(:, :)
      integer, intent(in) :: n_alpha, n_beta
      real(dp), intent(out), contiguous :: fock_a(:, :), fock_b(:, :)
      type(error_t), intent(inout) :: error
      real(dp), intent(in), optional :: k_scale
      real(c_double) :: kf

      if (.not. c_associated(shared_handle)) then
         call error%set(ERROR_VALIDATION, "b200: build_fock_df_uhf called before b200_set_tensor")
         return
      end if
      kf = 1.0_c_double
      if (present(k_scale)) kf = k_scale
      if (.not. operands_match_tensor(MQCB200_SLOT_FULL_RANGE, size(h, 1), error)) return
      if (mqcb200_build_fock_uhf(shared_handle, MQCB200_SLOT_FULL_RANGE, h, density_total, &
                                 coeff_a, int(size(coeff_a, 1), c_int), int(n_alpha, c_int), &
                                 coeff_b, int(size(coeff_b, 1), c_int), int(n_beta, c_int), kf, fock_a, fock_b) &
          /= MQCB200_OK) call engine_failure("b200: build_fock_uhf", error)
   end subroutine b200_build_fock_df_uhf

   subroutine b200_build_df_tensor(device_rank, three, metric, n_ao, half, error, attenuated)
      !! The last two stages of build_df_tensor (mqc_libcint_integrals.F90:981-987) on the device:
      !! half = metric^(-1/2), b = three . half into the resident packed layout.  `half` comes back
      !! because df_two_electron_gradient needs it again (b200_df_gradient_densities).
      integer, intent(in) :: device_rank
      real(dp), intent(in), contiguous :: three(:, :), metric(:, :)
      integer, intent(in) :: n_ao
      real(dp), intent(out), contiguous :: half(:, :)
      type(error_t), intent(inout) :: error
      logical, intent(in), optional :: attenuated
      type(c_ptr) :: handle
      integer(c_int) :: slot

      call get_engine(device_rank, handle, error)
      if (error%has_error()) return
      slot = MQCB200_SLOT_FULL_RANGE
      if (present(attenuated)) then
         if (attenuated) slot = MQCB200_SLOT_ATTENUATED
      end if
      if (mqcb200_build_df_tensor(handle, slot, int(n_ao, c_int), int(size(three, 2), c_int), three, metric, &
                                  1.0e-10_c_double, half) /= MQCB200_OK) call engine_failure("b200: build_df_tensor", error)
   end subroutine b200_build_df_tensor

   subroutine b200_metric_inverse_sqrt(device_rank, metric, half, n_kept, error)
      !! metric_inverse_sqrt(metric, half, error) (mqc_libcint_integrals.F90:992-1038) on the device: same
      !! 1e-10 threshold, same zeroed modes, same refusal of a singular metric
      integer, intent(in) :: device_rank
      real(dp), intent(in), contiguous :: metric(:, :)
      real(dp), intent(out), contiguous :: half(:, :)
      integer, intent(out) :: n_kept
      type(error_t), intent(inout) :: error
      type(c_ptr) :: handle
      integer(c_int) :: kept

      n_kept = 0
      call get_engine(device_rank, handle, error)
      if (error%has_error()) return
      if (size(metric, 1) /= size(metric, 2) .or. size(half, 1) /= size(metric, 1) .or. &
          size(half, 2) /= size(metric, 2)) then
         call error%set(ERROR_VALIDATION, "b200: the metric and its inverse square root are (naux, naux)")
         return
      end if
      if (mqcb200_metric_inverse_sqrt(handle, int(size(metric, 1), c_int), metric, 1.0e-10_c_double, half, kept) &
          /= MQCB200_OK) then
         call engine_failure("b200: metric_inverse_sqrt", error)
         return
      end if
      n_kept = int(kept)
   end subroutine b200_metric_inverse_sqrt

   subroutine b200_whiten_begin(device_rank, n_ao, naux_total, q_begin, q_count, half, error, attenuated)
      !! Start of the slab-streamed b = three . half (mqc_libcint_integrals.F90:981-987) for this rank's
      !! auxiliary slab [q_begin, q_begin + q_count) (0-based): three(nao*nao, naux) -- 115 GB for the 200-atom
      !! case -- never has to exist.  Collective over the ranks when b200_comm_init has been called.
      integer, intent(in) :: device_rank, n_ao, naux_total, q_begin, q_count
      real(dp), intent(in), contiguous :: half(:, :)
      type(error_t), intent(inout) :: error
      logical, intent(in), optional :: attenuated
      type(c_ptr) :: handle
      integer(c_int) :: slot

      call get_engine(device_rank, handle, error)
      if (error%has_error()) return
      slot = MQCB200_SLOT_FULL_RANGE
      if (present(attenuated)) then
         if (attenuated) slot = MQCB200_SLOT_ATTENUATED
      end if
      if (size(half, 1) /= naux_total .or. size(half, 2) /= naux_total) then
         call error%set(ERROR_VALIDATION, "b200: half is (naux, naux)")
         return
      end if
      if (mqcb200_whiten_begin(handle, slot, int(n_ao, c_int), int(naux_total, c_int), int(q_begin, c_int), &
                               int(q_count, c_int), half) /= MQCB200_OK) call engine_failure("b200: whiten_begin", error)
   end subroutine b200_whiten_begin

   subroutine b200_whiten_push(nu_begin, three_block, n_ao, error, attenuated)
      !! One block of (mu nu|P): three_block(nao*nu_count, naux) holds all mu and all P for the orbital
      !! columns nu in [nu_begin, nu_begin + nu_count) (0-based; whole 16-wide column tiles unless the
      !! block ends at nao) -- what three_centre produces for a range of shells
      integer, intent(in) :: nu_begin, n_ao
      real(dp), intent(in), contiguous :: three_block(:, :)
      type(error_t), intent(inout) :: error
      logical, intent(in), optional :: attenuated
      integer(c_int) :: slot

      if (.not. c_associated(shared_handle)) then
         call error%set(ERROR_VALIDATION, "b200: whiten_push called before b200_whiten_begin")
         return
      end if
      slot = MQCB200_SLOT_FULL_RANGE
      if (present(attenuated)) then
         if (attenuated) slot = MQCB200_SLOT_ATTENUATED
      end if
      if (n_ao <= 0 .or. mod(size(three_block, 1), max(n_ao, 1)) /= 0) then
         call error%set(ERROR_VALIDATION, "b200: a block of the three-centre tensor is (nao*nu_count, naux)")
         return
      end if
      if (mqcb200_whiten_push(shared_handle, slot, int(nu_begin, c_int), int(size(three_block, 1)/n_ao, c_int), &
                              three_block, int(size(three_block, 1), c_int64_t)) /= MQCB200_OK) &
         call engine_failure("b200: whiten_push", error)
   end subroutine b200_whiten_push

   subroutine b200_whiten_end(error, attenuated)
      !! The tensor becomes usable (collective when sharded: a rank whose push failed fails everyone here)
      type(error_t), intent(inout) :: error
      logical, intent(in), optional :: attenuated
      integer(c_int) :: slot

      if (.not. c_associated(shared_handle)) then
         call error%set(ERROR_VALIDATION, "b200: whiten_end called before b200_whiten_begin")
         return
      end if
      slot = MQCB200_SLOT_FULL_RANGE
      if (present(attenuated)) then
         if (attenuated) slot = MQCB200_SLOT_ATTENUATED
      end if
      if (mqcb200_whiten_end(shared_handle, slot) /= MQCB200_OK) call engine_failure("b200: whiten_end", error)
   end subroutine b200_whiten_end

   subroutine b200_set_tensor_shard(device_rank, bmat_shard, n_ao, naux_total, q_begin, error)
      !! This rank's auxiliary slab [q_begin, q_begin + size(bmat_shard, 2)) of a whole-molecule tensor
      !! (0-based q_begin); builds then end in one exchange of [J|K] over NVLink (b200_comm_init first)
      integer, intent(in) :: device_rank, n_ao, naux_total, q_begin
      real(dp), intent(in), contiguous :: bmat_shard(:, :)
      type(error_t), intent(inout) :: error
      type(c_ptr) :: handle

      call get_engine(device_rank, handle, error)
      if (error%has_error()) return
      if (mqcb200_set_tensor_shard(handle, MQCB200_SLOT_FULL_RANGE, int(n_ao, c_int), int(naux_total, c_int), &
                                   int(q_begin, c_int), int(size(bmat_shard, 2), c_int), bmat_shard) /= MQCB200_OK) &
         call engine_failure("b200: set_tensor_shard", error)
   end subroutine b200_set_tensor_shard

   subroutine b200_comm_unique_id(id, error)
      !! Rank 0 creates the 128-byte id and broadcasts it with the host program's own transport
      !! (the reference: bcast of src/parallel/mqc_bcast.f90)
      character(kind=c_char), intent(out) :: id(128)
      type(error_t), intent(inout) :: error
      if (mqcb200_comm_unique_id(id) /= MQCB200_OK) call engine_failure("b200: comm_unique_id", error)
   end subroutine b200_comm_unique_id

   subroutine b200_comm_init(device_rank, n_ranks, rank, id, error)
      integer, intent(in) :: device_rank, n_ranks, rank
      character(kind=c_char), intent(in) :: id(128)
      type(error_t), intent(inout) :: error
      type(c_ptr) :: handle

      call get_engine(device_rank, handle, error)
      if (error%has_error()) return
      if (mqcb200_comm_init(handle, int(n_ranks, c_int), int(rank, c_int), id) /= MQCB200_OK) &
         call engine_failure("b200: comm_init", error)
   end subroutine b200_comm_init

   subroutine b200_response_operator_df(x, c_occ, dtilde, g, error, k_scale)
      !! response_operator_df(b, x, c_occ, dtilde, g, k_scale) with b resident
      !! (backends/libcint/mqc_libcint_cphf.F90:499-566), direct rank-2 form on the device
      real(dp), intent(in), contiguous :: x(:, :), c_occ(:, :), dtilde(:, :)
      real(dp), intent(out), contiguous :: g(:, :)
      type(error_t), intent(inout) :: error
      real(dp), intent(in), optional :: k_scale
      real(c_double) :: kf

      kf = 1.0_c_double
      if (present(k_scale)) kf = k_scale
      if (.not. operands_match_tensor(MQCB200_SLOT_FULL_RANGE, size(dtilde, 1), error)) return
      if (mqcb200_response_operator(shared_handle, MQCB200_SLOT_FULL_RANGE, x, int(size(x, 1), c_int), c_occ, &
                                    int(size(c_occ, 1), c_int), int(size(c_occ, 2), c_int), dtilde, kf, g) /= MQCB200_OK) &
         call engine_failure("b200: response_operator", error)
   end subroutine b200_response_operator_df

   subroutine b200_fitted_potential_general(dens, g, error, k_scale)
      !! fitted_potential_general(b, dens, g, k_scale) with b resident (mqc_libcint_cphf.F90:568-616)
      real(dp), intent(in), contiguous :: dens(:, :)
      real(dp), intent(out), contiguous :: g(:, :)
      type(error_t), intent(inout) :: error
      real(dp), intent(in), optional :: k_scale
      real(c_double) :: kf

      kf = 1.0_c_double
      if (present(k_scale)) kf = k_scale
      if (.not. operands_match_tensor(MQCB200_SLOT_FULL_RANGE, size(dens, 1), error)) return
      if (mqcb200_fitted_potential_general(shared_handle, MQCB200_SLOT_FULL_RANGE, dens, kf, g) /= MQCB200_OK) &
         call engine_failure("b200: fitted_potential_general", error)
   end subroutine b200_fitted_potential_general

   subroutine b200_df_gradient_densities(half, total_density, orbitals, n_occupied, gamma, omega, error, &
                                         orbitals_beta, n_occupied_beta, exx_fraction, with_coulomb)
      !! gamma(nao, nao, naux) and omega(naux, naux) of df_two_electron_gradient
      !! (backends/libcint/mqc_libcint_gradient.f90:1654-1719, add_exchange_channel :1748-1812) from the
      !! resident tensor; the derivative-integral contractions (:1738-1770) stay where they are
      real(dp), intent(in), contiguous :: half(:, :), total_density(:, :), orbitals(:, :)
      integer, intent(in) :: n_occupied
      real(dp), intent(out), contiguous :: gamma(:, :, :), omega(:, :)
      type(error_t), intent(inout) :: error
      real(dp), intent(in), contiguous, optional :: orbitals_beta(:, :)
      integer, intent(in), optional :: n_occupied_beta
      real(dp), intent(in), optional :: exx_fraction
      logical, intent(in), optional :: with_coulomb
      real(c_double) :: kf
      integer(c_int) :: coulomb, status

      kf = 1.0_c_double
      if (present(exx_fraction)) kf = exx_fraction
      coulomb = 1_c_int
      if (present(with_coulomb)) then
         if (.not. with_coulomb) coulomb = 0_c_int
      end if
      if (.not. operands_match_tensor(MQCB200_SLOT_FULL_RANGE, size(total_density, 1), error)) return
      if (present(orbitals_beta)) then
         status = mqcb200_df_gradient_densities(shared_handle, MQCB200_SLOT_FULL_RANGE, half, total_density, orbitals, &
                                                int(size(orbitals, 1), c_int), int(n_occupied, c_int), orbitals_beta, &
                                                int(size(orbitals_beta, 1), c_int), int(n_occupied_beta, c_int), 1_c_int, &
                                                kf, coulomb, gamma, omega)
      else
         status = mqcb200_df_gradient_densities(shared_handle, MQCB200_SLOT_FULL_RANGE, half, total_density, orbitals, &
                                                int(size(orbitals, 1), c_int), int(n_occupied, c_int), orbitals, &
                                                int(size(orbitals, 1), c_int), 0_c_int, 0_c_int, kf, coulomb, gamma, omega)
      end if
      if (status /= MQCB200_OK) call engine_failure("b200: df_gradient_densities", error)
   end subroutine b200_df_gradient_densities

   subroutine b200_run_rhf_fragment(h, overlap, nelec, max_iter, energy_tol, density_tol, electronic, iterations, &
                                    converged, orbitals, orbital_energies, density, error, diis_vectors, k_scale)
      !! run_libcint_rhf's loop (mqc_libcint_rhf.f90:566-649) with every
      !! matrix staying on the GPU: H, S and the resident tensor in, the converged SCF out
      real(dp), intent(in), contiguous :: h(:, :), overlap(:, :)
      integer, intent(in) :: nelec, max_iter
      real(dp), intent(in) :: energy_tol, density_tol
      real(dp), intent(out) :: electronic
      integer, intent(out) :: iterations
      logical, intent(out) :: converged
      real(dp), intent(out), contiguous :: orbitals(:, :), orbital_energies(:), density(:, :)
      type(error_t), intent(inout) :: error
      integer, intent(in), optional :: diis_vectors
      real(dp), intent(in), optional :: k_scale
      real(dp), allocatable :: history(:)
      integer(c_int) :: it, conv, n_mo, diis
      real(c_double) :: e, kf

      diis = 8_c_int                                         ! the reference's default (:443)
      if (present(diis_vectors)) diis = int(diis_vectors, c_int)
      kf = 1.0_c_double
      if (present(k_scale)) kf = k_scale
      converged = .false.
      if (.not. operands_match_tensor(MQCB200_SLOT_FULL_RANGE, size(h, 1), error)) return
      allocate (history(max_iter))
      ! mqcb200_scf: the one-CTA route for fragment-sized problems, the general kernels beyond (any nao, sharded tensors too)
      if (mqcb200_scf(shared_handle, MQCB200_SLOT_FULL_RANGE, h, overlap, int(nelec, c_int), 1_c_int, &
                               int(max_iter, c_int), energy_tol, density_tol, diis, kf, e, it, conv, n_mo, &
                               orbitals, orbital_energies, density, history) /= MQCB200_OK) then
         call engine_failure("b200: scf_fragment", error)
         return
      end if
      electronic = e
      iterations = it
      converged = conv /= 0
   end subroutine b200_run_rhf_fragment

   subroutine b200_finalize()
      integer(c_int) :: status
      if (c_associated(shared_handle)) status = mqcb200_destroy(shared_handle)
      shared_handle = c_null_ptr
      shared_device_rank = -1
   end subroutine b200_finalize

   subroutine engine_failure(context, error)
      character(len=*), intent(in) :: context
      type(error_t), intent(inout) :: error
      character(kind=c_char) :: buffer(512)
      character(len=512) :: message
      integer :: i

      call mqcb200_last_error(int(size(buffer), c_int), buffer)
      message = ""
      do i = 1, size(buffer)
         if (buffer(i) == c_null_char) exit
         message(i:i) = buffer(i)
      end do
      call error%set(ERROR_VALIDATION, context//": "//trim(message))
   end subroutine engine_failure

end module mqc_b200_fock

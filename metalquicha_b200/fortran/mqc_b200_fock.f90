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
   use, intrinsic :: iso_c_binding, only: c_ptr, c_null_ptr, c_associated, c_int, c_double, c_char, c_null_char
   use pic_types, only: dp
   use mqc_error, only: error_t, ERROR_VALIDATION
   use mqc_b200_iface
   implicit none
   private

   public :: b200_backend_available
   public :: b200_set_tensor, b200_clear_tensors
   public :: b200_build_fock_df
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
      if (mqcb200_build_fock(shared_handle, slot, h, density, coeff, int(size(coeff, 1), c_int), &
                             int(n_occ, c_int), kf, jf, fock) /= MQCB200_OK) then
         call engine_failure("b200: build_fock", error)
      end if
   end subroutine b200_build_fock_df

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

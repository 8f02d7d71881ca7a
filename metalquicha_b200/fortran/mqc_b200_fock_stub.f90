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
   public :: b200_build_fock_df
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

   subroutine b200_finalize()
   end subroutine b200_finalize

end module mqc_b200_fock

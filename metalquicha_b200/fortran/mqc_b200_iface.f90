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
!  C header it mirrors IS compiled and tested (tests/test_abi.py checks every symbol).
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
   public :: mqcb200_comm_unique_id, mqcb200_comm_init, mqcb200_comm_destroy
   public :: mqcb200_queue_create, mqcb200_queue_pop, mqcb200_queue_is_empty, mqcb200_queue_destroy
   public :: MQCB200_OK, MQCB200_FAIL, MQCB200_BAD_HANDLE
   public :: MQCB200_SLOT_FULL_RANGE, MQCB200_SLOT_ATTENUATED

   integer(c_int), parameter :: MQCB200_OK = 0          !! same values as MQC_OK/MQC_FAIL/MQC_BAD_HANDLE,
   integer(c_int), parameter :: MQCB200_FAIL = 1        !! src/interface/mqc_capi_status.f90:21-23
   integer(c_int), parameter :: MQCB200_BAD_HANDLE = 2
   integer(c_int), parameter :: MQCB200_SLOT_FULL_RANGE = 0   !! bmat
   integer(c_int), parameter :: MQCB200_SLOT_ATTENUATED = 1   !! bmat_lr

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
   end interface

end module mqc_b200_iface

!------------------------------------------------------------------------------
! recursion_gpu_shim -- how `type recursion` (source/recursion.f90) calls the engine.
!
! The bodies below replace the bodies of the type-bound procedures of the same name in
! recursion.f90 (1807-1866 recur_b, 1655-1737 recur_b_ij, 3485-3532 recur, 1980-2023 zsqr,
! 3057-3130 chebyshev_recur, 2376-2487 chebyshev_recur_ij, 979-1234 compute_moments_stochastic);
! every name, argument list and result member (a_b, b2_b, a, b2, mu_n, mu_nm_stochastic) is
! unchanged, so self.f90 / green.f90 / density_of_states.f90 / conductivity.f90 compile and run
! untouched.  `this%gpu` is one new component of the type: `type(c_ptr) :: gpu = c_null_ptr`.
! Work vectors (psi_b, pmn_b, hpsi, ...) are no longer allocated on the host.
!------------------------------------------------------------------------------
module recursion_gpu_shim
   use, intrinsic :: iso_c_binding
   use rsrec_c_mod
   use mpi_mod, only: rank, start_atom, end_atom, g2l_map, get_mpi_variables
   implicit none
contains

   !> called at the top of every driver: (re)exports the lattice tables and the block sets that
   !> self%run_recursion has just rebuilt (self.f90:777-797)
   subroutine gpu_export(this)
      use recursion_mod, only: recursion
      class(recursion), intent(inout), target :: this
      integer :: nslot
      nslot = size(this%hamiltonian%ee, 3)
      if (.not. c_associated(this%gpu)) then
         call rsrec_check(rsrec_create(this%gpu, int(mod(rank, 8), c_int), int(this%lattice%kk, c_int), &
                                       int(size(this%lattice%nn, 2), c_int), int(nslot, c_int), &
                                       int(this%lattice%ntype, c_int), int(this%lattice%nmax, c_int)), __FILE__, __LINE__)
      end if
      call rsrec_check(rsrec_set_lattice(this%gpu, this%lattice%nn, this%lattice%iz), __FILE__, __LINE__)
      call rsrec_check(rsrec_set_hamiltonian(this%gpu, c_loc(this%hamiltonian%ee), c_loc(this%hamiltonian%eeo), &
                                             c_loc(this%hamiltonian%hall), c_loc(this%hamiltonian%hallo), &
                                             c_loc(this%hamiltonian%lsham), c_loc(this%hamiltonian%enim), &
                                             merge(1_c_int, 0_c_int, this%hamiltonian%hoh)), __FILE__, __LINE__)
   end subroutine

   subroutine recur_b(this)
      use recursion_mod, only: recursion
      class(recursion), intent(inout), target :: this
      integer :: nloc, i, l, ll
      integer(c_int32_t), allocatable :: sites(:), none(:)
      complex(c_double_complex), allocatable :: ones(:)
      call get_mpi_variables(rank, this%lattice%nrec)
      call gpu_export(this)
      nloc = end_atom - start_atom + 1
      allocate (sites(nloc), none(nloc), ones(nloc))
      sites = this%lattice%irec(start_atom:end_atom); none = 0; ones = (1.0d0, 0.0d0)
      call rsrec_check(rsrec_lanczos_block(this%gpu, int(nloc, c_int), sites, none, ones, ones, &
                                           int(this%lattice%control%lld, c_int), this%a_b, this%b2_b), __FILE__, __LINE__)
      do i = 1, nloc
         do ll = 1, this%lattice%control%lld
            do l = 1, 18
               this%a(ll, l, i, 1) = real(this%a_b(l, l, ll, i))
               this%b2(ll, l, i, 1) = real(this%b2_b(l, l, ll, i))
            end do
         end do
      end do
   end subroutine

   subroutine chebyshev_recur(this)
      use recursion_mod, only: recursion
      class(recursion), intent(inout), target :: this
      integer :: nloc
      real(c_double) :: a, b
      integer(c_int32_t), allocatable :: sites(:), none(:)
      complex(c_double_complex), allocatable :: ones(:)
      a = (this%en%energy_max - this%en%energy_min)/(2 - 0.3)
      b = (this%en%energy_max + this%en%energy_min)/2
      call gpu_export(this)
      nloc = end_atom - start_atom + 1
      allocate (sites(nloc), none(nloc), ones(nloc))
      sites = this%lattice%irec(start_atom:end_atom); none = 0; ones = (1.0d0, 0.0d0)
      ! RSREC_EDIVERGED is the reference's own fatal "Chebyshev moments did not converge" (recursion.f90:2594)
      call rsrec_check(rsrec_cheb_moments(this%gpu, int(nloc, c_int), sites, none, ones, ones, &
                                          int(this%lattice%control%lld, c_int), a, b, this%mu_n), __FILE__, __LINE__)
   end subroutine

   subroutine zsqr(this)
      use recursion_mod, only: recursion
      use mpi_mod, only: atoms_per_process
      class(recursion), intent(inout), target :: this
      integer :: na
      na = atoms_per_process
      if (this%lattice%njij /= 0) na = atoms_per_process*4
      call rsrec_check(rsrec_zsqr(this%gpu, this%b2_b, int(this%lattice%control%lld, c_int), int(na, c_int)), &
                       __FILE__, __LINE__)
   end subroutine

   ! recur_b_ij / chebyshev_recur_ij: build the unit list exactly as the reference's loops do --
   ! slot ij_loc*4-4+reci, signs (1,1),(1,-1),(1,i),(1,-i)/sqrt(2), a single (1,1) unit when i == j --
   ! and call rsrec_lanczos_block / rsrec_cheb_moments once; recur -> rsrec_lanczos_scalar;
   ! compute_moments_stochastic -> rsrec_set_operator('a'/'b') + rsrec_kubo_moments with the random
   ! numbers drawn on the host (random_number) and passed as `phases`.  See rslmtoasa_b200/recursion.py
   ! (`_pair_units`, `recur`, `compute_moments_stochastic`) for the executable statement of the same logic.
end module recursion_gpu_shim

!------------------------------------------------------------------------------
! bands_gpu_shim -- how `type bands` (bands.f90) consumes the on-site Green function through the engine.
!
! Bodies for the type-bound procedures of the same name: calculate_fermi (227-347), calculate_band_energy (354-359),
! calculate_magnetic_moments (791-855), calculate_moments (409-524, with calculate_orbital_moments 1075-1156).
! Every Green-function call of green_gpu_shim leaves g0 on the device; with the fused drivers called with
! g0 = c_null_ptr nothing of size 18*18*nv leaves the GPU -- only dtot(nv) and a few numbers per atom come back.
! The file output of calculate_fermi (totaldos.out, *_dos.out) keeps using dtot/dosia/dosial as before.
! NOT compiled in the build container (no Fortran compiler there).
!------------------------------------------------------------------------------
module bands_gpu_shim
   use, intrinsic :: iso_c_binding
   use rsrec_c_mod
   use mpi_mod
   implicit none
contains

   !> bands%calculate_fermi: DOS loops on the device, MPI_ALLREDUCE of dtot as before, Fermi scan on the device
   subroutine calculate_fermi(this)
      use bands_mod, only: bands
      class(bands), intent(inout), target :: this
      integer(c_int) :: nv, nv1, ifail
      real(c_double) :: fermi, e1
      nv = int(this%en%channels_ldos + 10, c_int)
      this%qqv = real(sum(this%symbolic_atom(1:this%lattice%nbulk_bulk)%element%valence))
      call rsrec_check(rsrec_bands_dos(this%recursion%gpu, this%dtot, c_null_ptr, c_null_ptr), __FILE__, __LINE__)
#ifdef USE_MPI
      call MPI_ALLREDUCE(MPI_IN_PLACE, this%dtot, nv, MPI_DOUBLE_PRECISION, MPI_SUM, MPI_COMM_WORLD, ierr)
#endif
      fermi = this%en%fermi
      this%en%chebfermi = this%en%fermi
      nv1 = int(this%en%ik1, c_int)
      if (this%en%fix_fermi .or. this%control%calctype == 'B') then
         call rsrec_check(rsrec_bands_fermi(this%recursion%gpu, this%dtot, nv, this%en%edel, this%en%energy_min, this%qqv, &
                                            merge(1_c_int, 0_c_int, this%en%fix_fermi), fermi, nv1, e1, ifail), &
                          __FILE__, __LINE__)
         if (ifail == 0) then
            this%en%fermi = fermi
            this%nv1 = nv1
            this%e1 = e1
         end if
      end if
   end subroutine

   !> bands%calculate_band_energy
   subroutine calculate_band_energy(this)
      use bands_mod, only: bands
      class(bands), intent(inout), target :: this
      call rsrec_check(rsrec_bands_band_energy(this%recursion%gpu, this%dtot, int(this%en%channels_ldos + 10, c_int), &
                                               this%en%ene, this%en%edel, this%en%fermi, int(this%nv1, c_int), this%e1, &
                                               this%eband), __FILE__, __LINE__)
   end subroutine

   !> bands%calculate_magnetic_moments: the six Simpson integrals per atom come from the device, the scalar
   !> normalisation (mtot, mom, the nsp < 3 rule) stays as in bands.f90:833-846
   subroutine calculate_magnetic_moments(this)
      use bands_mod, only: bands
      class(bands), intent(inout), target :: this
      real(c_double) :: m0(3, atoms_per_process), m1(3, atoms_per_process)
      integer :: na, na_loc
      call rsrec_check(rsrec_bands_magnetic_moments(this%recursion%gpu, this%en%ene, this%en%edel, this%en%fermi, &
                                                    int(this%nv1, c_int), this%e1, m0, m1), __FILE__, __LINE__)
      do na = start_atom, end_atom
         na_loc = g2l_map(na)
         associate (pot => this%symbolic_atom(this%lattice%nbulk + na)%potential)
            pot%mx = m0(1, na_loc); pot%my = m0(2, na_loc); pot%mz = m0(3, na_loc)
            pot%mom0 = m0(:, na_loc)
            pot%mom1 = m1(:, na_loc)
            pot%mtot = sqrt(pot%mx**2 + pot%my**2 + pot%mz**2) + 1.0d-15
            pot%mom = m0(:, na_loc)/pot%mtot
            if (this%control%nsp < 3) pot%mom(:) = [0.0d0, 0.0d0, 1.0d0]
         end associate
      end do
   end subroutine

   !> bands%calculate_moments: occupations / first and second moments per (l, spin) and the orbital moments
   subroutine calculate_moments(this)
      use bands_mod, only: bands
      class(bands), intent(inout), target :: this
      real(c_double) :: mom(3, atoms_per_process), occ(3, 6, atoms_per_process), lmom(3, atoms_per_process)
      real(c_double) :: sgef, pmef, smef
      integer :: na, na_loc, i, nsp, soff
      do na = start_atom, end_atom
         mom(:, g2l_map(na)) = this%symbolic_atom(this%lattice%nbulk + na)%potential%mom
      end do
      call rsrec_check(rsrec_bands_moments(this%recursion%gpu, int(this%en%channels_ldos, c_int), this%en%ene, this%en%edel, &
                                           this%en%fermi, int(this%nv1, c_int), this%e1, mom, occ, lmom), __FILE__, __LINE__)
      do na = start_atom, end_atom
         na_loc = g2l_map(na)
         associate (pot => this%symbolic_atom(this%lattice%nbulk + na)%potential)
            pot%lmom = lmom(:, na_loc)
            do i = 1, 6
               nsp = merge(2, 1, i > 3)
               soff = 3*(nsp - 1)
               sgef = occ(1, i, na_loc); pmef = occ(2, i, na_loc); smef = occ(3, i, na_loc)
               pot%gravity_center(i - soff, nsp) = (pmef/sgef) - pot%vmad
               pot%ql(1, i - soff - 1, nsp) = sgef
               pot%ql(2, i - soff - 1, nsp) = 0.0d0
               pot%ql(3, i - soff - 1, nsp) = smef - 2.0d0*(pmef/sgef)*pmef + ((pmef/sgef)**2)*sgef
            end do
         end associate
      end do
      call this%calculate_pl()
      ! the MPI_ALLREDUCE of the flattened potentials follows unchanged (bands.f90:502-512)
   end subroutine

end module bands_gpu_shim

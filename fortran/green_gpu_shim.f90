!------------------------------------------------------------------------------
! green_gpu_shim -- how the consumers of the recursion results call the engine (SURVEY.md 8f rows 1-3).
!
! Bodies for the type-bound procedures of the same name in green.f90 (588-621 block_green, 628-705 sgreen,
! 1030-1108 chebyshev_green, 1191-1339 bgreen), density_of_states.f90 (248-372 density) and the energy integrand
! of conductivity.f90 (158-306).  Names, arguments and result members (g0, mu_ng, gamma-contracted integrand)
! are unchanged; `this%recursion%gpu` is the handle created by recursion_gpu_shim's gpu_export.
! NOT compiled in the build container (no Fortran compiler there).
!------------------------------------------------------------------------------
module green_gpu_shim
   use, intrinsic :: iso_c_binding
   use rsrec_c_mod
   use mpi_mod, only: atoms_per_process, start_atom, end_atom, g2l_map
   implicit none
contains

   !> green%block_green (green.f90:588-621): get_terminf + bgreen(eta = 0) for this rank's units in one call
   subroutine block_green(this)
      use green_mod, only: green
      class(green), intent(inout), target :: this
      integer :: nv
      nv = this%en%channels_ldos + 10
      call rsrec_check(rsrec_block_green(this%recursion%gpu, this%recursion%a_b, this%recursion%b2_b, &
                                         int(atoms_per_process, c_int), int(this%control%lld, c_int), this%en%ene, &
                                         int(nv, c_int), merge(1_c_int, 0_c_int, this%control%sym_term), this%g0), &
                       __FILE__, __LINE__)
   end subroutine

   !> green%bgreen (green.f90:1191-1339): one unit, a window of energy channels
   subroutine bgreen(this, g_out, ia, ie_start, ie_len, a_inf, b_inf, eta)
      use green_mod, only: green
      class(green), intent(inout), target :: this
      complex(c_double_complex), intent(out) :: g_out(:, :, :)
      integer, intent(in) :: ia, ie_start, ie_len
      real(c_double), intent(in) :: a_inf(18, 18), b_inf(18, 18)
      complex(c_double_complex), intent(in) :: eta
      call rsrec_check(rsrec_bgreen(this%recursion%gpu, this%recursion%a_b(:, :, :, ia), this%recursion%b2_b(:, :, :, ia), &
                                    int(this%control%lld, c_int), this%en%ene, int(size(g_out, 3), c_int), &
                                    int(ie_start, c_int), int(ie_len, c_int), a_inf, b_inf, real(eta, c_double), &
                                    aimag(eta), merge(1_c_int, 0_c_int, this%control%sym_term), g_out), __FILE__, __LINE__)
   end subroutine

   !> green%chebyshev_green (green.f90:1030-1108)
   subroutine chebyshev_green(this)
      use green_mod, only: green
      class(green), intent(inout), target :: this
      call rsrec_check(rsrec_chebyshev_green(this%recursion%gpu, this%recursion%mu_n, int(atoms_per_process, c_int), &
                                             int(this%control%lld, c_int), this%en%ene, &
                                             int(this%en%channels_ldos + 10, c_int), this%en%energy_min, &
                                             this%en%energy_max, this%recursion%mu_ng, this%g0), __FILE__, __LINE__)
   end subroutine

   !> green%sgreen (green.f90:628-705); dw_l = sqrt of the potential's Delta, cshi its band-centre shift per atom
   !> (the two per-atom arrays dos%density reads from symbolic_atoms(..)%potential, density_of_states.f90:300-304)
   subroutine sgreen(this, dw_l, cshi, nmdir)
      use green_mod, only: green
      class(green), intent(inout), target :: this
      real(c_double), intent(in) :: dw_l(18, *), cshi(18, *)
      integer, intent(in) :: nmdir
      call rsrec_check(rsrec_sgreen(this%recursion%gpu, this%recursion%a, this%recursion%b2, int(size(this%recursion%a, 1), c_int), &
                                    int(atoms_per_process, c_int), int(nmdir, c_int), this%en%ene, &
                                    int(this%en%channels_ldos + 10, c_int), dw_l, cshi, this%g0), __FILE__, __LINE__)
   end subroutine

   !> the energy integrand of conductivity%calculate_conductivity_tensor (conductivity.f90:267-290); gamma_nm is
   !> contracted on the fly and never stored, so calculate_gamma_nm's (nv, M, M) array is no longer allocated
   subroutine conductivity_integrand(this, integrand, integrand_at, loop_over)
      use conductivity_mod, only: conductivity
      class(conductivity), intent(inout), target :: this
      complex(c_double_complex), intent(out) :: integrand(18, *), integrand_at(18, this%en%channels_ldos + 10, *)
      integer, intent(in) :: loop_over
      call rsrec_check(rsrec_conductivity_integrand(this%recursion%gpu, this%recursion%mu_nm_stochastic, &
                                                    int(this%control%cond_ll, c_int), int(loop_over, c_int), this%en%ene, &
                                                    int(this%en%channels_ldos + 10, c_int), this%en%energy_min, &
                                                    this%en%energy_max, &
                                                    merge(1_c_int, 0_c_int, this%control%cond_calctype == 'per_type'), &
                                                    integrand, integrand_at), __FILE__, __LINE__)
   end subroutine

   !> green%calculate_intersite_gf (green.f90:425-469): the four Green functions of every pair come from the
   !> device-resident g0 (staged flow: recur_b_ij / chebyshev_recur_ij filled a_b / mu_n with four slots per pair, then
   !> block_green / chebyshev_green over the 4*njij units), and are combined to gij, gji and their spin components there
   subroutine calculate_intersite_gf(this)
      use green_mod, only: green
      class(green), intent(inout), target :: this
      integer(c_int32_t) :: pi(atoms_per_process), pj(atoms_per_process)
      complex(c_double_complex), allocatable, target :: gspin(:, :, :, :, :)
      integer :: ia, ia_glob, nv
      nv = this%en%channels_ldos + 10
      do ia_glob = start_atom, end_atom
         ia = g2l_map(ia_glob)
         pi(ia) = int(this%lattice%ijpair(ia_glob, 1), c_int32_t)
         pj(ia) = int(this%lattice%ijpair(ia_glob, 2), c_int32_t)
      end do
      if (this%control%recur == 'block') then
         call this%recursion%zsqr()
         call rsrec_check(rsrec_block_green(this%recursion%gpu, this%recursion%a_b, this%recursion%b2_b, &
                                            int(4*atoms_per_process, c_int), int(this%control%lld, c_int), this%en%ene, &
                                            int(nv, c_int), merge(1_c_int, 0_c_int, this%control%sym_term), this%g0), &
                          __FILE__, __LINE__)   ! g0 sized (18,18,nv,4*njij_loc) here; only its device copy is consumed
      else
         call rsrec_check(rsrec_chebyshev_green(this%recursion%gpu, this%recursion%mu_n, int(4*atoms_per_process, c_int), &
                                                int(this%control%lld, c_int), this%en%ene, int(nv, c_int), &
                                                this%en%energy_min, this%en%energy_max, this%recursion%mu_ng, this%g0), &
                          __FILE__, __LINE__)
      end if
      allocate (gspin(9, 9, nv, atoms_per_process, 8))
      call rsrec_check(rsrec_intersite_gf(this%recursion%gpu, int(atoms_per_process, c_int), pi, pj, 0_c_int, this%gij, &
                                          this%gji, c_loc(gspin)), __FILE__, __LINE__)
      this%ginmag = gspin(:, :, :, :, 1); this%gix = gspin(:, :, :, :, 2); this%giy = gspin(:, :, :, :, 3)
      this%giz = gspin(:, :, :, :, 4); this%gjnmag = gspin(:, :, :, :, 5); this%gjx = gspin(:, :, :, :, 6)
      this%gjy = gspin(:, :, :, :, 7); this%gjz = gspin(:, :, :, :, 8)
   end subroutine

   ! Fused drivers (no host round trip of a_b/b2_b, mu_n or mu_nm_stochastic): self%run_recursion + self%run_dos of the
   ! block path -> rsrec_recur_b_green; chebyshev_recur + chebyshev_green -> rsrec_cheb_recur_green;
   ! compute_moments_stochastic + calculate_conductivity_tensor -> rsrec_kubo_conductivity; recur_b_ij / chebyshev_recur_ij
   ! + calculate_intersite_gf -> rsrec_recur_b_ij_green / rsrec_cheb_recur_ij_green (g0 = c_null_ptr) + rsrec_intersite_gf.  See
   ! rslmtoasa_b200/green.py (recur_b_green, chebyshev_recur_green, compute_conductivity) for the executable statement.
end module green_gpu_shim

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
! Procedures here: gpu_export, gpu_report_phases, recur_b, recur_b_ij, recur, zsqr, chebyshev_recur, chebyshev_recur_ij,
! compute_moments_stochastic (the consumers' bodies are in green_gpu_shim.f90).  NOT compiled in the build container.
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
      integer :: nslot, ndev
      type(c_ptr) :: p_eeo, p_hall, p_hallo, p_enim
      nslot = size(this%hamiltonian%ee, 3)
      if (.not. c_associated(this%gpu)) then
         ! one process per GPU: the device is the rank modulo the devices this process can see (with one rank per node
         ! slot the launcher's local rank; CUDA_VISIBLE_DEVICES narrows the list further)
         ndev = max(1, int(rsrec_device_count()))
         call rsrec_check(rsrec_create(this%gpu, int(mod(rank, ndev), c_int), int(this%lattice%kk, c_int), &
                                       int(size(this%lattice%nn, 2), c_int), int(nslot, c_int), &
                                       int(this%lattice%ntype, c_int), int(this%lattice%nmax, c_int)), __FILE__, __LINE__)
      end if
      call rsrec_check(rsrec_set_lattice(this%gpu, this%lattice%nn, this%lattice%iz), __FILE__, __LINE__)
      call rsrec_check(rsrec_set_positions(this%gpu, this%lattice%cr), __FILE__, __LINE__)   ! work ordering only
      ! arrays the reference leaves unallocated (hall/hallo without a local region, eeo/hallo/enim without hoh) are passed
      ! as NULL: c_loc of an unallocated array is not defined, and the C side expects NULL for "not used"
      p_eeo = c_null_ptr; p_hall = c_null_ptr; p_hallo = c_null_ptr; p_enim = c_null_ptr
      if (this%lattice%nmax > 0 .and. allocated(this%hamiltonian%hall)) p_hall = c_loc(this%hamiltonian%hall)
      if (this%hamiltonian%hoh) then
         if (allocated(this%hamiltonian%eeo)) p_eeo = c_loc(this%hamiltonian%eeo)
         if (allocated(this%hamiltonian%enim)) p_enim = c_loc(this%hamiltonian%enim)
         if (this%lattice%nmax > 0 .and. allocated(this%hamiltonian%hallo)) p_hallo = c_loc(this%hamiltonian%hallo)
      end if
      call rsrec_check(rsrec_set_hamiltonian(this%gpu, c_loc(this%hamiltonian%ee), p_eeo, p_hall, p_hallo, &
                                             c_loc(this%hamiltonian%lsham), p_enim, &
                                             merge(1_c_int, 0_c_int, this%hamiltonian%hoh)), __FILE__, __LINE__)
   end subroutine

   !> copies the device phase times of the last call into the host's profile tree under the reference's own labels
   !> (recursion.f90:1902-1970, 3104-3127).  g_timer measures wall time between start/stop, so the shim brackets an empty
   !> region and adds the device milliseconds with g_timer%add (a three-line addition to timer.f90, see INTEGRATION.md).
   subroutine gpu_report_phases(this)
      use recursion_mod, only: recursion
      use timer_mod, only: g_timer
      class(recursion), intent(inout) :: this
      real(c_double) :: ms(16)
      integer(c_long_long) :: calls(16)
      integer :: k
      call rsrec_check(rsrec_phase_read(this%gpu, ms, calls), __FILE__, __LINE__)
      do k = 1, int(rsrec_phase_count())
         if (calls(k) > 0) call g_timer%add(rsrec_phase_label_f(k - 1), ms(k)*1.0d-3, int(calls(k)))
      end do
   end subroutine

   subroutine recur_b(this)
      use recursion_mod, only: recursion
      class(recursion), intent(inout), target :: this
      integer :: nloc, i, l, ll
      integer(c_int32_t), allocatable :: sites(:), none(:)
      complex(c_double_complex), allocatable :: ones(:)
      call get_mpi_variables(rank, this%lattice%nrec)
      call gpu_export(this)
      call rsrec_check(rsrec_phase_timing(this%gpu, 1_c_int), __FILE__, __LINE__)   ! 'H|PSI_n>', 'B_n+1', ... of crecal_b
      nloc = end_atom - start_atom + 1
      allocate (sites(nloc), none(nloc), ones(nloc))
      sites = this%lattice%irec(start_atom:end_atom); none = 0; ones = (1.0d0, 0.0d0)
      call rsrec_check(rsrec_lanczos_block(this%gpu, int(nloc, c_int), sites, none, ones, ones, &
                                           int(this%lattice%control%lld, c_int), this%a_b, this%b2_b), __FILE__, __LINE__)
      call gpu_report_phases(this)
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
      call rsrec_check(rsrec_phase_timing(this%gpu, 1_c_int), __FILE__, __LINE__)   ! '<PSI_0|PSI_0>', '<PSI_0|PSI_1>', '<PSI_0|PSI_n>'
      call rsrec_check(rsrec_cheb_moments(this%gpu, int(nloc, c_int), sites, none, ones, ones, &
                                          int(this%lattice%control%lld, c_int), a, b, this%mu_n), __FILE__, __LINE__)
      call gpu_report_phases(this)
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

   !> unit list of recur_b_ij / chebyshev_recur_ij (recursion.f90:1666-1707, 2390-2440): result slot ij_loc*4-4+reci,
   !> signs (1,1),(1,-1),(1,i),(1,-i)/sqrt(2); a pair with i == j keeps only reci = 1 with signs (1,1)
   subroutine pair_units(this, n, si, sj, asg, bsg, slots)
      use recursion_mod, only: recursion
      use math_mod, only: one_over_sqrt_two
      class(recursion), intent(in) :: this
      integer, intent(out) :: n
      integer(c_int32_t), allocatable, intent(out) :: si(:), sj(:)
      complex(c_double_complex), allocatable, intent(out) :: asg(:), bsg(:)
      integer, allocatable, intent(out) :: slots(:)
      integer :: ij, ij_loc, i, j, reci, nmax_units
      complex(c_double_complex), parameter :: bs(4) = [(1.0d0, 0.0d0), (-1.0d0, 0.0d0), (0.0d0, 1.0d0), (0.0d0, -1.0d0)]
      nmax_units = 4*max(end_atom - start_atom + 1, 0)
      allocate (si(nmax_units), sj(nmax_units), asg(nmax_units), bsg(nmax_units), slots(nmax_units))
      n = 0
      do ij = start_atom, end_atom
         ij_loc = g2l_map(ij)
         i = this%lattice%ijpair(ij, 1)
         j = this%lattice%ijpair(ij, 2)
         do reci = 1, 4
            if (i == j .and. reci > 1) cycle
            n = n + 1
            si(n) = i; sj(n) = j; slots(n) = ij_loc*4 - 4 + reci
            if (i == j) then
               asg(n) = (1.0d0, 0.0d0); bsg(n) = (1.0d0, 0.0d0)
            else
               asg(n) = (1.0d0, 0.0d0)*one_over_sqrt_two; bsg(n) = bs(reci)*one_over_sqrt_two
            end if
         end do
      end do
   end subroutine

   subroutine recur_b_ij(this)
      use recursion_mod, only: recursion
      class(recursion), intent(inout), target :: this
      integer :: n, u, lld
      integer(c_int32_t), allocatable :: si(:), sj(:)
      complex(c_double_complex), allocatable :: asg(:), bsg(:), a_u(:, :, :, :), b_u(:, :, :, :)
      integer, allocatable :: slots(:)
      call get_mpi_variables(rank, this%lattice%njij)
      call gpu_export(this)
      call pair_units(this, n, si, sj, asg, bsg, slots)
      lld = this%lattice%control%lld
      allocate (a_u(18, 18, lld, max(n, 1)), b_u(18, 18, lld, max(n, 1)))
      call rsrec_check(rsrec_lanczos_block(this%gpu, int(n, c_int), si, sj, asg, bsg, int(lld, c_int), a_u, b_u), &
                       __FILE__, __LINE__)
      do u = 1, n
         this%a_b(:, :, :, slots(u)) = a_u(:, :, :, u)
         this%b2_b(:, :, :, slots(u)) = b_u(:, :, :, u)
      end do
   end subroutine

   subroutine chebyshev_recur_ij(this)
      use recursion_mod, only: recursion
      class(recursion), intent(inout), target :: this
      integer :: n, u, lld
      real(c_double) :: a, b
      integer(c_int32_t), allocatable :: si(:), sj(:)
      complex(c_double_complex), allocatable :: asg(:), bsg(:), mu_u(:, :, :, :)
      integer, allocatable :: slots(:)
      a = (this%en%energy_max - this%en%energy_min)/(2 - 0.3)
      b = (this%en%energy_max + this%en%energy_min)/2
      call get_mpi_variables(rank, this%lattice%njij)
      call gpu_export(this)
      call pair_units(this, n, si, sj, asg, bsg, slots)
      lld = this%lattice%control%lld
      allocate (mu_u(18, 18, 2*lld + 2, max(n, 1)))
      call rsrec_check(rsrec_cheb_moments(this%gpu, int(n, c_int), si, sj, asg, bsg, int(lld, c_int), a, b, mu_u), &
                       __FILE__, __LINE__)
      do u = 1, n
         this%mu_n(:, :, :, slots(u)) = mu_u(:, :, :, u)
      end do
   end subroutine

   !> scalar recursion (nsp = 1), recursion.f90:3485-3532: the 18 start orbitals of a site run as one block vector
   subroutine recur(this)
      use recursion_mod, only: recursion
      class(recursion), intent(inout), target :: this
      integer :: nloc, lld
      integer(c_int32_t), allocatable :: sites(:)
      real(c_double), allocatable :: a_u(:, :, :), b_u(:, :, :)
      call get_mpi_variables(rank, this%lattice%nrec)
      call gpu_export(this)
      nloc = end_atom - start_atom + 1
      lld = this%lattice%control%lld
      allocate (sites(nloc), a_u(lld, 18, nloc), b_u(lld, 18, nloc))
      sites = this%lattice%irec(start_atom:end_atom)
      call rsrec_check(rsrec_lanczos_scalar(this%gpu, int(nloc, c_int), sites, int(lld, c_int), a_u, b_u), __FILE__, __LINE__)
      this%a(1:lld, :, 1:nloc, 1) = a_u
      this%b2(1:lld, :, 1:nloc, 1) = b_u
   end subroutine

   !> compute_moments_stochastic (recursion.f90:979-1234).  The uniform numbers of the random-phase start vectors are
   !> drawn here with random_number (the reference calls random_seed() without arguments, 1106) and handed to the engine.
   subroutine compute_moments_stochastic(this)
      use recursion_mod, only: recursion
      class(recursion), intent(inout), target :: this
      integer :: loop_over, m
      real(c_double) :: a, b
      integer(c_int32_t), allocatable, target :: sites(:)
      real(c_double), allocatable, target :: phases(:, :)
      a = (this%en%energy_max - this%en%energy_min)/(2 - 0.3)
      b = (this%en%energy_max + this%en%energy_min)/2
      m = this%control%cond_ll
      call gpu_export(this)
      call rsrec_check(rsrec_set_operator(this%gpu, int(iachar('a'), c_int), c_loc(this%hamiltonian%v_a), &
                                          c_loc(this%hamiltonian%vo_a)), __FILE__, __LINE__)
      call rsrec_check(rsrec_set_operator(this%gpu, int(iachar('b'), c_int), c_loc(this%hamiltonian%v_b), &
                                          c_loc(this%hamiltonian%vo_b)), __FILE__, __LINE__)
      select case (this%control%cond_calctype)
      case ('per_type')
         loop_over = this%lattice%ntype
         allocate (sites(loop_over))
         sites = this%lattice%atlist(1:loop_over)
         call rsrec_check(rsrec_kubo_moments(this%gpu, int(loop_over, c_int), 0_c_int, c_loc(sites), c_null_ptr, &
                                             int(m, c_int), a, b, this%mu_nm_stochastic), __FILE__, __LINE__)
      case ('random_vec')
         loop_over = this%control%random_vec_num
         allocate (phases(this%lattice%kk, loop_over))
         call random_seed()
         call random_number(phases)
         call rsrec_check(rsrec_kubo_moments(this%gpu, int(loop_over, c_int), 1_c_int, c_null_ptr, c_loc(phases), &
                                             int(m, c_int), a, b, this%mu_nm_stochastic), __FILE__, __LINE__)
      end select
   end subroutine
end module recursion_gpu_shim

!------------------------------------------------------------------------------
! rsrec_c_mod -- ISO_C_BINDING interface of librsrec.so (include/rsrec.h).
!
! This is the thin layer BASELINE.json's north star asks for: the Fortran host of
! rslmtoasa keeps its `recursion` / `hamiltonian` / `lattice` derived types and calls
! the CUDA engine through these bindings.  One interface per C entry point, argument
! for argument; arrays are passed as they are held by the reference (column-major,
! complex(rp) == complex(c_double_complex), default integers == integer(c_int32_t)).
!
! NOT compiled in the build container (no Fortran compiler there); compile it next to
! source/*.f90 and link with -lrsrec (see INTEGRATION.md).
!------------------------------------------------------------------------------
module rsrec_c_mod
   use, intrinsic :: iso_c_binding
   implicit none
   private

   integer(c_int), parameter, public :: RSREC_OK = 0, RSREC_EINVAL = -1, RSREC_EDIVERGED = -2, &
                                        RSREC_ECUDA = -3, RSREC_ENOMEM = -4

   public :: rsrec_device_count
   public :: rsrec_create, rsrec_destroy, rsrec_set_lattice, rsrec_set_hamiltonian, rsrec_set_operator
   public :: rsrec_lanczos_block, rsrec_lanczos_scalar, rsrec_zsqr, rsrec_cheb_moments, rsrec_cheb_moments_random
   public :: rsrec_kubo_moments, rsrec_ham_vec_matmul, rsrec_velo_vec_matmul, rsrec_last_error_f, rsrec_check
   public :: rsrec_create_ll_map, rsrec_orbital_moments, rsrec_build_nn, rsrec_build_hamiltonian
   public :: rsrec_set_positions
   public :: rsrec_rotate_to_local_axis, rsrec_rotate_from_local_axis, rsrec_lanczos_block_local_axis
   ! consumers of the recursion results (green.f90, density_of_states.f90, conductivity.f90) and fused drivers
   public :: rsrec_bpopt, rsrec_get_terminf, rsrec_bgreen, rsrec_block_green, rsrec_chebyshev_green, rsrec_density
   public :: rsrec_sgreen, rsrec_conductivity_integrand, rsrec_recur_b_green, rsrec_cheb_recur_green
   public :: rsrec_kubo_conductivity
   ! consumers of g0 in the SCF loop (bands.f90)
   public :: rsrec_bands_set_g0, rsrec_bands_get_g0, rsrec_bands_g0_shape, rsrec_bands_dos, rsrec_bands_fermi
   public :: rsrec_bands_magnetic_moments, rsrec_bands_moments, rsrec_bands_band_energy
   ! exchange path: pair-unit fused drivers and calculate_intersite_gf
   public :: rsrec_recur_b_ij_green, rsrec_cheb_recur_ij_green, rsrec_intersite_gf, rsrec_conductivity_cumulative
   public :: rsrec_spin_diag_launch_count
   ! the exchange step of the unit-sharded path (NCCL inside the library) and the per-phase timers (g_timer labels)
   public :: rsrec_comm_unique_id, rsrec_comm_init, rsrec_comm_destroy, rsrec_comm_info, rsrec_shard_range
   public :: rsrec_allreduce, rsrec_allgather_units, rsrec_lanczos_block_sharded, rsrec_cheb_moments_random_sum
   public :: rsrec_phase_timing, rsrec_phase_count, rsrec_phase_label, rsrec_phase_read, rsrec_host_phase_read
   public :: rsrec_phase_label_f, RSREC_COMM_ID_BYTES
   integer(c_int), parameter :: RSREC_COMM_ID_BYTES = 128

   interface
      function rsrec_last_error() bind(C, name='rsrec_last_error') result(msg)
         import :: c_ptr
         type(c_ptr) :: msg
      end function

      ! CUDA devices visible to this process: device_ordinal = mod(local_rank, rsrec_device_count())
      function rsrec_device_count() bind(C, name='rsrec_device_count') result(n)
         import :: c_int
         integer(c_int) :: n
      end function

      ! recursion constructor: recursion.f90:132-143 / allocation 3713-3826
      function rsrec_create(h, device_ordinal, kk, ncols, nslot, ntype, nmax) bind(C, name='rsrec_create') result(rc)
         import :: c_ptr, c_int
         type(c_ptr), intent(out) :: h
         integer(c_int), value :: device_ordinal, kk, ncols, nslot, ntype, nmax
         integer(c_int) :: rc
      end function

      function rsrec_destroy(h) bind(C, name='rsrec_destroy') result(rc)
         import :: c_ptr, c_int
         type(c_ptr), value :: h
         integer(c_int) :: rc
      end function

      ! lattice%nn(kk, ncols), lattice%iz(kk)
      function rsrec_set_lattice(h, nn, iz) bind(C, name='rsrec_set_lattice') result(rc)
         import :: c_ptr, c_int, c_int32_t
         type(c_ptr), value :: h
         integer(c_int32_t), intent(in) :: nn(*), iz(*)
         integer(c_int) :: rc
      end function

      ! optional lattice%cr(3, kk): orders the work for L2 locality only (results unchanged)
      function rsrec_set_positions(h, cr) bind(C, name='rsrec_set_positions') result(rc)
         import :: c_ptr, c_int, c_double
         type(c_ptr), value :: h
         real(c_double), intent(in) :: cr(3, *)
         integer(c_int) :: rc
      end function

      ! hamiltonian%ee, eeo, hall, hallo, lsham, enim, hoh   (pass c_null_ptr for unused arrays)
      function rsrec_set_hamiltonian(h, ee, eeo, hall, hallo, lsham, enim, hoh) &
         bind(C, name='rsrec_set_hamiltonian') result(rc)
         import :: c_ptr, c_int
         type(c_ptr), value :: h, ee, eeo, hall, hallo, lsham, enim
         integer(c_int), value :: hoh
         integer(c_int) :: rc
      end function

      ! hamiltonian%v_a/vo_a (slot = iachar('a')) or v_b/vo_b (slot = iachar('b'))
      function rsrec_set_operator(h, slot, v_op, vo_op) bind(C, name='rsrec_set_operator') result(rc)
         import :: c_ptr, c_int
         type(c_ptr), value :: h, v_op, vo_op
         integer(c_int), value :: slot
         integer(c_int) :: rc
      end function

      ! recur_b / recur_b_ij (+ crecal_b): a_b, b2_b (18,18,lld,nunits)
      function rsrec_lanczos_block(h, nunits, site_i, site_j, asign, bsign, lld, a_b, b2_b) &
         bind(C, name='rsrec_lanczos_block') result(rc)
         import :: c_ptr, c_int, c_int32_t, c_double_complex
         type(c_ptr), value :: h
         integer(c_int), value :: nunits, lld
         integer(c_int32_t), intent(in) :: site_i(*), site_j(*)
         complex(c_double_complex), intent(in) :: asign(*), bsign(*)
         complex(c_double_complex), intent(out) :: a_b(18, 18, lld, *), b2_b(18, 18, lld, *)
         integer(c_int) :: rc
      end function

      ! recur (nsp = 1): a, b2 (lld,18,nunits)
      function rsrec_lanczos_scalar(h, nunits, sites, lld, a, b2) bind(C, name='rsrec_lanczos_scalar') result(rc)
         import :: c_ptr, c_int, c_int32_t, c_double
         type(c_ptr), value :: h
         integer(c_int), value :: nunits, lld
         integer(c_int32_t), intent(in) :: sites(*)
         real(c_double), intent(out) :: a(lld, 18, *), b2(lld, 18, *)
         integer(c_int) :: rc
      end function

      function rsrec_zsqr(h, b2_b, lld, na) bind(C, name='rsrec_zsqr') result(rc)
         import :: c_ptr, c_int, c_double_complex
         type(c_ptr), value :: h
         integer(c_int), value :: lld, na
         complex(c_double_complex), intent(inout) :: b2_b(18, 18, lld, *)
         integer(c_int) :: rc
      end function

      ! chebyshev_recur / chebyshev_recur_ij: mu_n (18,18,2*lld+2,nunits)
      function rsrec_cheb_moments(h, nunits, site_i, site_j, asign, bsign, lld, a_scale, b_shift, mu_n) &
         bind(C, name='rsrec_cheb_moments') result(rc)
         import :: c_ptr, c_int, c_int32_t, c_double, c_double_complex
         type(c_ptr), value :: h
         integer(c_int), value :: nunits, lld
         integer(c_int32_t), intent(in) :: site_i(*), site_j(*)
         complex(c_double_complex), intent(in) :: asign(*), bsign(*)
         real(c_double), value :: a_scale, b_shift
         complex(c_double_complex), intent(out) :: mu_n(18, 18, 2*lld + 2, *)
         integer(c_int) :: rc
      end function

      function rsrec_cheb_moments_random(h, nvec, phases, lld, a_scale, b_shift, mu_n) &
         bind(C, name='rsrec_cheb_moments_random') result(rc)
         import :: c_ptr, c_int, c_double, c_double_complex
         type(c_ptr), value :: h
         integer(c_int), value :: nvec, lld
         real(c_double), intent(in) :: phases(*)
         real(c_double), value :: a_scale, b_shift
         complex(c_double_complex), intent(out) :: mu_n(18, 18, 2*lld + 2, *)
         integer(c_int) :: rc
      end function

      ! compute_moments_stochastic: mu_nm (18,18,cond_ll,cond_ll,nstart)
      function rsrec_kubo_moments(h, nstart, start_kind, start_sites, phases, cond_ll, a_scale, b_shift, mu_nm) &
         bind(C, name='rsrec_kubo_moments') result(rc)
         import :: c_ptr, c_int, c_double, c_double_complex
         type(c_ptr), value :: h, start_sites, phases
         integer(c_int), value :: nstart, start_kind, cond_ll
         real(c_double), value :: a_scale, b_shift
         complex(c_double_complex), intent(out) :: mu_nm(18, 18, cond_ll, cond_ll, *)
         integer(c_int) :: rc
      end function

      function rsrec_ham_vec_matmul(h, psi_in, psi_out, a_scale, b_shift) bind(C, name='rsrec_ham_vec_matmul') result(rc)
         import :: c_ptr, c_int, c_double, c_double_complex
         type(c_ptr), value :: h
         complex(c_double_complex), intent(in) :: psi_in(18, 18, *)
         complex(c_double_complex), intent(out) :: psi_out(18, 18, *)
         real(c_double), value :: a_scale, b_shift
         integer(c_int) :: rc
      end function

      function rsrec_velo_vec_matmul(h, slot, psi_in, psi_out) bind(C, name='rsrec_velo_vec_matmul') result(rc)
         import :: c_ptr, c_int, c_double_complex
         type(c_ptr), value :: h
         integer(c_int), value :: slot
         complex(c_double_complex), intent(in) :: psi_in(18, 18, *)
         complex(c_double_complex), intent(out) :: psi_out(18, 18, *)
         integer(c_int) :: rc
      end function
      ! lattice%nncal + lattice%remd (lattice.f90:3035-3123, 2823-2907) on the device; call once with ncols = 0 and
      ! nn = c_null_ptr to query nm, allocate lattice%nn(kk, nm+1), call again (replaces lattice.f90:1851-1868)
      function rsrec_build_nn(device_ordinal, kk, crd, no, ntot, iu, ct, pbc, nrep, a, alat, ncols, nn, nm) &
         bind(C, name='rsrec_build_nn') result(rc)
         import :: c_ptr, c_int, c_int32_t, c_double
         integer(c_int), value :: device_ordinal, kk, ntot, ncols
         real(c_double), intent(in) :: crd(3, *), a(3, 3)
         integer(c_int32_t), intent(in) :: no(*), iu(*), pbc(3), nrep(3)
         real(c_double), value :: ct, alat
         type(c_ptr), value :: nn
         integer(c_int), intent(out) :: nm
         integer(c_int) :: rc
      end function

      ! device-side build_bulkham / build_locham (hamiltonian.f90:1553-1667 with ham0m_nc 2225-2303, hcpx); optional
      ! downloads are passed as c_loc(array) or c_null_ptr
      function rsrec_build_hamiltonian(h, hhh, jt, it, pot, mom, lsham, hoh, ee, eeo, hall, hallo, enim, obarm) &
         bind(C, name='rsrec_build_hamiltonian') result(rc)
         import :: c_ptr, c_int, c_int32_t, c_double, c_double_complex
         type(c_ptr), value :: h, ee, eeo, hall, hallo, enim, obarm
         real(c_double), intent(in) :: hhh(9, 9, *), mom(3, *)
         integer(c_int32_t), intent(in) :: jt(*), it(*)
         complex(c_double_complex), intent(in) :: pot(9, 12, *), lsham(18, 18, *)
         integer(c_int), value :: hoh
         integer(c_int) :: rc
      end function

      ! hamiltonian%rotate_to_local_axis / rotate_from_local_axis: hamiltonian.f90:2442-2484
      function rsrec_rotate_to_local_axis(h, m_loc) bind(C, name='rsrec_rotate_to_local_axis') result(rc)
         import :: c_ptr, c_int, c_double
         type(c_ptr), value :: h
         real(c_double), intent(in) :: m_loc(3)
         integer(c_int) :: rc
      end function

      function rsrec_rotate_from_local_axis(h) bind(C, name='rsrec_rotate_from_local_axis') result(rc)
         import :: c_ptr, c_int
         type(c_ptr), value :: h
         integer(c_int) :: rc
      end function

      ! recur_b with hamiltonian%local_axis (recursion.f90:1826-1832); mom(3, nunits)
      function rsrec_lanczos_block_local_axis(h, nunits, site_i, mom, lld, a_b, b2_b) &
         bind(C, name='rsrec_lanczos_block_local_axis') result(rc)
         import :: c_ptr, c_int, c_int32_t, c_double, c_double_complex
         type(c_ptr), value :: h
         integer(c_int), value :: nunits, lld
         integer(c_int32_t), intent(in) :: site_i(*)
         real(c_double), intent(in) :: mom(3, *)
         complex(c_double_complex), intent(out) :: a_b(18, 18, lld, *), b2_b(18, 18, lld, *)
         integer(c_int) :: rc
      end function

      ! create_ll_map: recursion.f90:3277-3303 (start mask izeroll(site,1) = 1); izeroll(0:kk, lld+1)
      function rsrec_create_ll_map(h, site, lld, izeroll) bind(C, name='rsrec_create_ll_map') result(rc)
         import :: c_ptr, c_int, c_int32_t
         type(c_ptr), value :: h
         integer(c_int), value :: site, lld
         integer(c_int32_t), intent(out) :: izeroll(*)
         integer(c_int) :: rc
      end function

      ! chebyshev_orbital_mod (moment part): recursion.f90:2901-3008
      function rsrec_orbital_moments(h, nstart, start_sites, cr, alat, lld, a_scale, b_shift, mu_n_orb) &
         bind(C, name='rsrec_orbital_moments') result(rc)
         import :: c_ptr, c_int, c_int32_t, c_double, c_double_complex
         type(c_ptr), value :: h
         integer(c_int), value :: nstart, lld
         integer(c_int32_t), intent(in) :: start_sites(*)
         real(c_double), intent(in) :: cr(3, *)
         real(c_double), value :: alat, a_scale, b_shift
         complex(c_double_complex), intent(out) :: mu_n_orb(18, 18, *)
         integer(c_int) :: rc
      end function

      ! bpopt (+ emami): recursion.f90:3540-3706, batched over chains; a, rb (ll, nchains)
      function rsrec_bpopt(h, nchains, ll, a, rb, ainf, rbinf, ifail) bind(C, name='rsrec_bpopt') result(rc)
         import :: c_ptr, c_int, c_double
         type(c_ptr), value :: h
         integer(c_int), value :: nchains, ll
         real(c_double), intent(in) :: a(ll, *), rb(ll, *)
         real(c_double), intent(out) :: ainf(*), rbinf(*)
         integer(c_int), intent(out) :: ifail(*)
         integer(c_int) :: rc
      end function

      ! get_terminf: recursion.f90:2092-2138
      function rsrec_get_terminf(h, a_b, b_b, na, ll, a_inf, b_inf, a_inf0, b_inf0) &
         bind(C, name='rsrec_get_terminf') result(rc)
         import :: c_ptr, c_int, c_double, c_double_complex
         type(c_ptr), value :: h
         integer(c_int), value :: na, ll
         complex(c_double_complex), intent(in) :: a_b(18, 18, ll, *), b_b(18, 18, ll, *)
         real(c_double), intent(out) :: a_inf(18, 18, *), b_inf(18, 18, *), a_inf0(*), b_inf0(*)
         integer(c_int) :: rc
      end function

      ! bgreen: green.f90:1191-1339 (one unit; eta passed as two reals)
      function rsrec_bgreen(h, a_b, b_b, ll, e, nv, ie_start, ie_len, a_inf, b_inf, eta_re, eta_im, sym_term, g_out) &
         bind(C, name='rsrec_bgreen') result(rc)
         import :: c_ptr, c_int, c_double, c_double_complex
         type(c_ptr), value :: h
         integer(c_int), value :: ll, nv, ie_start, ie_len, sym_term
         complex(c_double_complex), intent(in) :: a_b(18, 18, *), b_b(18, 18, *)
         real(c_double), intent(in) :: e(*), a_inf(18, 18), b_inf(18, 18)
         real(c_double), value :: eta_re, eta_im
         complex(c_double_complex), intent(out) :: g_out(18, 18, *)
         integer(c_int) :: rc
      end function

      ! block_green: green.f90:588-621
      function rsrec_block_green(h, a_b, b_b, na, ll, e, nv, sym_term, g0) bind(C, name='rsrec_block_green') result(rc)
         import :: c_ptr, c_int, c_double, c_double_complex
         type(c_ptr), value :: h
         integer(c_int), value :: na, ll, nv, sym_term
         complex(c_double_complex), intent(in) :: a_b(18, 18, ll, *), b_b(18, 18, ll, *)
         real(c_double), intent(in) :: e(*)
         complex(c_double_complex), intent(out) :: g0(18, 18, nv, *)
         integer(c_int) :: rc
      end function

      ! chebyshev_green: green.f90:1030-1108
      function rsrec_chebyshev_green(h, mu_n, na, lld, ene, nv, energy_min, energy_max, mu_ng, g0) &
         bind(C, name='rsrec_chebyshev_green') result(rc)
         import :: c_ptr, c_int, c_double, c_double_complex
         type(c_ptr), value :: h
         integer(c_int), value :: na, lld, nv
         complex(c_double_complex), intent(in) :: mu_n(18, 18, 2*lld + 2, *)
         real(c_double), intent(in) :: ene(*)
         real(c_double), value :: energy_min, energy_max
         complex(c_double_complex), intent(out) :: mu_ng(18, 18, 2*lld + 2, *), g0(18, 18, nv, *)
         integer(c_int) :: rc
      end function

      ! dos%density + bprldos: density_of_states.f90:248-407, all (atom, direction) pairs at once
      function rsrec_density(h, a, b2, lld, na, nmdir, ene, nv, dw_l, cshi, tdens) bind(C, name='rsrec_density') result(rc)
         import :: c_ptr, c_int, c_double
         type(c_ptr), value :: h
         integer(c_int), value :: lld, na, nmdir, nv
         real(c_double), intent(in) :: a(lld, 18, na, *), b2(lld, 18, na, *), ene(*), dw_l(18, *), cshi(18, *)
         real(c_double), intent(out) :: tdens(18, nv, na, *)
         integer(c_int) :: rc
      end function

      ! sgreen: green.f90:628-705
      function rsrec_sgreen(h, a, b2, lld, na, nmdir, ene, nv, dw_l, cshi, g0) bind(C, name='rsrec_sgreen') result(rc)
         import :: c_ptr, c_int, c_double, c_double_complex
         type(c_ptr), value :: h
         integer(c_int), value :: lld, na, nmdir, nv
         real(c_double), intent(in) :: a(lld, 18, na, *), b2(lld, 18, na, *), ene(*), dw_l(18, *), cshi(18, *)
         complex(c_double_complex), intent(out) :: g0(18, 18, nv, *)
         integer(c_int) :: rc
      end function

      ! calculate_gamma_nm + integrand of calculate_conductivity_tensor: conductivity.f90:158-306
      function rsrec_conductivity_integrand(h, mu_nm, M, nloop, ene, nv, energy_min, energy_max, per_type, &
                                            integrand, integrand_at) bind(C, name='rsrec_conductivity_integrand') result(rc)
         import :: c_ptr, c_int, c_double, c_double_complex
         type(c_ptr), value :: h
         integer(c_int), value :: M, nloop, nv, per_type
         complex(c_double_complex), intent(in) :: mu_nm(18, 18, M, M, *)
         real(c_double), intent(in) :: ene(*)
         real(c_double), value :: energy_min, energy_max
         complex(c_double_complex), intent(out) :: integrand(18, *), integrand_at(18, nv, *)
         integer(c_int) :: rc
      end function

      ! fused: recur_b -> zsqr -> get_terminf -> bgreen (self%run_recursion + self%run_dos, self.f90:799-856)
      function rsrec_recur_b_green(h, nunits, site_i, lld, ene, nv, sym_term, a_b, b2_b, g0) &
         bind(C, name='rsrec_recur_b_green') result(rc)
         import :: c_ptr, c_int, c_int32_t, c_double, c_double_complex
         type(c_ptr), value :: h
         integer(c_int), value :: nunits, lld, nv, sym_term
         integer(c_int32_t), intent(in) :: site_i(*)
         real(c_double), intent(in) :: ene(*)
         complex(c_double_complex), intent(out) :: a_b(18, 18, lld, *), b2_b(18, 18, lld, *), g0(18, 18, nv, *)
         integer(c_int) :: rc
      end function

      ! fused: chebyshev_recur -> chebyshev_green
      function rsrec_cheb_recur_green(h, nunits, site_i, lld, energy_min, energy_max, ene, nv, mu_n, mu_ng, g0) &
         bind(C, name='rsrec_cheb_recur_green') result(rc)
         import :: c_ptr, c_int, c_int32_t, c_double, c_double_complex
         type(c_ptr), value :: h
         integer(c_int), value :: nunits, lld, nv
         integer(c_int32_t), intent(in) :: site_i(*)
         real(c_double), value :: energy_min, energy_max
         real(c_double), intent(in) :: ene(*)
         complex(c_double_complex), intent(out) :: mu_n(18, 18, 2*lld + 2, *), mu_ng(18, 18, 2*lld + 2, *), g0(18, 18, nv, *)
         integer(c_int) :: rc
      end function

      ! fused: compute_moments_stochastic -> calculate_gamma_nm -> integrand (mu_nm may be c_null_ptr)
      function rsrec_kubo_conductivity(h, nstart, start_kind, start_sites, phases, M, energy_min, energy_max, ene, nv, &
                                       mu_nm, integrand, integrand_at) bind(C, name='rsrec_kubo_conductivity') result(rc)
         import :: c_ptr, c_int, c_double, c_double_complex
         type(c_ptr), value :: h, start_sites, phases, mu_nm
         integer(c_int), value :: nstart, start_kind, M, nv
         real(c_double), value :: energy_min, energy_max
         real(c_double), intent(in) :: ene(*)
         complex(c_double_complex), intent(out) :: integrand(18, *), integrand_at(18, nv, *)
         integer(c_int) :: rc
      end function
      ! ---- type bands (bands.f90): g0 stays on the device after every Green-function call ----
      function rsrec_bands_set_g0(h, g0, nunits, nv) bind(C, name='rsrec_bands_set_g0') result(rc)
         import :: c_ptr, c_int, c_double_complex
         type(c_ptr), value :: h
         integer(c_int), value :: nunits, nv
         complex(c_double_complex), intent(in) :: g0(18, 18, nv, *)
         integer(c_int) :: rc
      end function

      function rsrec_bands_get_g0(h, g0) bind(C, name='rsrec_bands_get_g0') result(rc)
         import :: c_ptr, c_int, c_double_complex
         type(c_ptr), value :: h
         complex(c_double_complex), intent(out) :: g0(18, 18, *)
         integer(c_int) :: rc
      end function

      function rsrec_bands_g0_shape(h, nunits, nv) bind(C, name='rsrec_bands_g0_shape') result(rc)
         import :: c_ptr, c_int
         type(c_ptr), value :: h
         integer(c_int), intent(out) :: nunits, nv
         integer(c_int) :: rc
      end function

      ! DOS loops of calculate_fermi (bands.f90:260-273); dosia, dosial may be c_null_ptr
      function rsrec_bands_dos(h, dtot, dosia, dosial) bind(C, name='rsrec_bands_dos') result(rc)
         import :: c_ptr, c_int, c_double
         type(c_ptr), value :: h, dosia, dosial
         real(c_double), intent(out) :: dtot(*)
         integer(c_int) :: rc
      end function

      ! Fermi level of calculate_fermi (bands.f90:322-342 with `fermi`, 366-402)
      function rsrec_bands_fermi(h, dtot, nv, edel, energy_min, qqv, fix_fermi, fermi, nv1, e1, ifail) &
         bind(C, name='rsrec_bands_fermi') result(rc)
         import :: c_ptr, c_int, c_double
         type(c_ptr), value :: h
         real(c_double), intent(in) :: dtot(*)
         integer(c_int), value :: nv, fix_fermi
         real(c_double), value :: edel, energy_min, qqv
         real(c_double), intent(inout) :: fermi
         integer(c_int), intent(inout) :: nv1
         real(c_double), intent(out) :: e1
         integer(c_int), intent(out) :: ifail
         integer(c_int) :: rc
      end function

      ! calculate_magnetic_moments (bands.f90:791-855): mom0 = mx,my,mz; mom1 = potential%mom1, (3,nunits)
      function rsrec_bands_magnetic_moments(h, ene, edel, fermi, nv1, e1, mom0, mom1) &
         bind(C, name='rsrec_bands_magnetic_moments') result(rc)
         import :: c_ptr, c_int, c_double
         type(c_ptr), value :: h
         real(c_double), intent(in) :: ene(*)
         real(c_double), value :: edel, fermi, e1
         integer(c_int), value :: nv1
         real(c_double), intent(out) :: mom0(3, *), mom1(3, *)
         integer(c_int) :: rc
      end function

      ! calculate_moments + calculate_orbital_moments (bands.f90:409-524, 1075-1156): occ(3,6,nunits) = sgef,pmef,smef
      function rsrec_bands_moments(h, channels_ldos, ene, edel, fermi, nv1, e1, mom, occ, lmom) &
         bind(C, name='rsrec_bands_moments') result(rc)
         import :: c_ptr, c_int, c_double
         type(c_ptr), value :: h
         integer(c_int), value :: channels_ldos, nv1
         real(c_double), intent(in) :: ene(*), mom(3, *)
         real(c_double), value :: edel, fermi, e1
         real(c_double), intent(out) :: occ(3, 6, *), lmom(3, *)
         integer(c_int) :: rc
      end function

      ! calculate_band_energy (bands.f90:354-359)
      function rsrec_bands_band_energy(h, dtot, nv, ene, edel, fermi, nv1, e1, eband) &
         bind(C, name='rsrec_bands_band_energy') result(rc)
         import :: c_ptr, c_int, c_double
         type(c_ptr), value :: h
         real(c_double), intent(in) :: dtot(*), ene(*)
         integer(c_int), value :: nv, nv1
         real(c_double), value :: edel, fermi, e1
         real(c_double), intent(out) :: eband
         integer(c_int) :: rc
      end function

      ! fused: recur_b_ij -> zsqr -> block_green_ij (a_b, b2_b, g0 may be c_null_ptr)
      function rsrec_recur_b_ij_green(h, nunits, site_i, site_j, asign, bsign, lld, ene, nv, sym_term, a_b, b2_b, g0) &
         bind(C, name='rsrec_recur_b_ij_green') result(rc)
         import :: c_ptr, c_int, c_int32_t, c_double, c_double_complex
         type(c_ptr), value :: h, a_b, b2_b, g0
         integer(c_int), value :: nunits, lld, nv, sym_term
         integer(c_int32_t), intent(in) :: site_i(*), site_j(*)
         complex(c_double_complex), intent(in) :: asign(*), bsign(*)
         real(c_double), intent(in) :: ene(*)
         integer(c_int) :: rc
      end function

      ! fused: chebyshev_recur_ij -> chebyshev_green_ij (mu_n, mu_ng, g0 may be c_null_ptr)
      function rsrec_cheb_recur_ij_green(h, nunits, site_i, site_j, asign, bsign, lld, energy_min, energy_max, ene, nv, &
                                         mu_n, mu_ng, g0) bind(C, name='rsrec_cheb_recur_ij_green') result(rc)
         import :: c_ptr, c_int, c_int32_t, c_double, c_double_complex
         type(c_ptr), value :: h, mu_n, mu_ng, g0
         integer(c_int), value :: nunits, lld, nv
         integer(c_int32_t), intent(in) :: site_i(*), site_j(*)
         complex(c_double_complex), intent(in) :: asign(*), bsign(*)
         real(c_double), value :: energy_min, energy_max
         real(c_double), intent(in) :: ene(*)
         integer(c_int) :: rc
      end function

      ! calculate_intersite_gf (green.f90:425-469) on the device-resident g0 of the pair units; gspin may be c_null_ptr
      function rsrec_intersite_gf(h, njij, pair_i, pair_j, compact, gij, gji, gspin) &
         bind(C, name='rsrec_intersite_gf') result(rc)
         import :: c_ptr, c_int, c_int32_t, c_double_complex
         type(c_ptr), value :: h, gspin
         integer(c_int), value :: njij, compact
         integer(c_int32_t), intent(in) :: pair_i(*), pair_j(*)
         complex(c_double_complex), intent(out) :: gij(18, 18, *), gji(18, 18, *)
         integer(c_int) :: rc
      end function

      ! tail of calculate_conductivity_tensor (conductivity.f90:300-372): sigma(2,19,nv,1+nat); integrand_at may be c_null_ptr
      function rsrec_conductivity_cumulative(h, integrand, integrand_at, nv, nv1, nat, wstep, loop_over, sigma) &
         bind(C, name='rsrec_conductivity_cumulative') result(rc)
         import :: c_ptr, c_int, c_double, c_double_complex
         type(c_ptr), value :: h, integrand_at
         complex(c_double_complex), intent(in) :: integrand(18, *)
         integer(c_int), value :: nv, nv1, nat, loop_over
         real(c_double), value :: wstep
         real(c_double), intent(out) :: sigma(2, 19, nv, *)
         integer(c_int) :: rc
      end function

      function rsrec_spin_diag_launch_count(h) bind(C, name='rsrec_spin_diag_launch_count') result(n)
         import :: c_ptr, c_long_long
         type(c_ptr), value :: h
         integer(c_long_long) :: n
      end function


      ! ---- exchange step: one process per GPU = one MPI rank of the reference (mpi.f90:32-58) ----
      ! rank 0 creates the id, the host broadcasts its 128 bytes (call MPI_Bcast(id, 128, MPI_BYTE, 0, comm, ierr))
      function rsrec_comm_unique_id(id128) bind(C, name='rsrec_comm_unique_id') result(rc)
         import :: c_int, c_signed_char
         integer(c_signed_char), intent(out) :: id128(*)
         integer(c_int) :: rc
      end function

      function rsrec_comm_init(h, nranks, rank, id128) bind(C, name='rsrec_comm_init') result(rc)
         import :: c_ptr, c_int, c_signed_char
         type(c_ptr), value :: h
         integer(c_int), value :: nranks, rank
         integer(c_signed_char), intent(in) :: id128(*)
         integer(c_int) :: rc
      end function

      function rsrec_comm_destroy(h) bind(C, name='rsrec_comm_destroy') result(rc)
         import :: c_ptr, c_int
         type(c_ptr), value :: h
         integer(c_int) :: rc
      end function

      function rsrec_comm_info(h, nranks, rank, nccl_version) bind(C, name='rsrec_comm_info') result(rc)
         import :: c_ptr, c_int
         type(c_ptr), value :: h
         integer(c_int), intent(out) :: nranks, rank, nccl_version
         integer(c_int) :: rc
      end function

      ! get_mpi_variables (mpi.f90:32-58): first..last = start_atom..end_atom of `rank`
      function rsrec_shard_range(rank, nranks, nunits, first, last) bind(C, name='rsrec_shard_range') result(rc)
         import :: c_int
         integer(c_int), value :: rank, nranks, nunits
         integer(c_int), intent(out) :: first, last
         integer(c_int) :: rc
      end function

      ! MPI_ALLREDUCE(MPI_IN_PLACE, buf, count, .., MPI_SUM) (bands.f90:270-275): dtype 0 real(rp), 1 complex(rp), 2 integer;
      ! buf = c_loc(array)
      function rsrec_allreduce(h, buf, count, dtype) bind(C, name='rsrec_allreduce') result(rc)
         import :: c_ptr, c_int, c_long_long
         type(c_ptr), value :: h, buf
         integer(c_long_long), value :: count
         integer(c_int), value :: dtype
         integer(c_int) :: rc
      end function

      ! the MPI_Allgather of per-unit results recursion.f90:1788-1799 leaves commented out; local, full = c_loc(array)
      function rsrec_allgather_units(h, local, full, doubles_per_unit, nunits_total) &
         bind(C, name='rsrec_allgather_units') result(rc)
         import :: c_ptr, c_int, c_long_long
         type(c_ptr), value :: h, local, full
         integer(c_long_long), value :: doubles_per_unit
         integer(c_int), value :: nunits_total
         integer(c_int) :: rc
      end function

      ! recur_b / recur_b_ij over all units of the job, a_b/b2_b(18,18,lld,nunits_total) gathered on the device
      function rsrec_lanczos_block_sharded(h, nunits_total, site_i, site_j, asign, bsign, lld, a_b, b2_b) &
         bind(C, name='rsrec_lanczos_block_sharded') result(rc)
         import :: c_ptr, c_int, c_int32_t, c_double_complex
         type(c_ptr), value :: h
         integer(c_int), value :: nunits_total, lld
         integer(c_int32_t), intent(in) :: site_i(*), site_j(*)
         complex(c_double_complex), intent(in) :: asign(*), bsign(*)
         complex(c_double_complex), intent(out) :: a_b(18, 18, lld, *), b2_b(18, 18, lld, *)
         integer(c_int) :: rc
      end function

      ! stochastic-trace KPM moments summed over the job's random vectors: mu_sum(18,18,2*lld+2)
      function rsrec_cheb_moments_random_sum(h, nvec_local, phases, lld, a_scale, b_shift, mu_sum) &
         bind(C, name='rsrec_cheb_moments_random_sum') result(rc)
         import :: c_ptr, c_int, c_double, c_double_complex
         type(c_ptr), value :: h
         integer(c_int), value :: nvec_local, lld
         real(c_double), intent(in) :: phases(*)
         real(c_double), value :: a_scale, b_shift
         complex(c_double_complex), intent(out) :: mu_sum(18, 18, *)
         integer(c_int) :: rc
      end function

      ! ---- per-phase device timing under the reference's g_timer labels (recursion.f90:1902-1970, 3104-3127) ----
      function rsrec_phase_timing(h, enable) bind(C, name='rsrec_phase_timing') result(rc)
         import :: c_ptr, c_int
         type(c_ptr), value :: h
         integer(c_int), value :: enable
         integer(c_int) :: rc
      end function

      function rsrec_phase_count() bind(C, name='rsrec_phase_count') result(n)
         import :: c_int
         integer(c_int) :: n
      end function

      function rsrec_phase_label(idx) bind(C, name='rsrec_phase_label') result(label)
         import :: c_ptr, c_int
         integer(c_int), value :: idx
         type(c_ptr) :: label
      end function

      function rsrec_phase_read(h, ms, calls) bind(C, name='rsrec_phase_read') result(rc)
         import :: c_ptr, c_int, c_double, c_long_long
         type(c_ptr), value :: h
         real(c_double), intent(out) :: ms(*)
         integer(c_long_long), intent(out) :: calls(*)
         integer(c_int) :: rc
      end function

      function rsrec_host_phase_read(h, seconds) bind(C, name='rsrec_host_phase_read') result(rc)
         import :: c_ptr, c_int, c_double
         type(c_ptr), value :: h
         real(c_double), intent(out) :: seconds(*)
         integer(c_int) :: rc
      end function

   end interface

contains

   !> C string of rsrec_last_error() as a Fortran string
   function rsrec_last_error_f() result(msg)
      character(len=:), allocatable :: msg
      character(kind=c_char), pointer :: p(:)
      type(c_ptr) :: cp
      integer :: n
      cp = rsrec_last_error()
      msg = ''
      if (.not. c_associated(cp)) return
      call c_f_pointer(cp, p, [1024])
      n = 0
      do while (n < 1024)
         if (p(n + 1) == c_null_char) exit
         n = n + 1
      end do
      allocate (character(len=n) :: msg)
      if (n > 0) msg = transfer(p(1:n), msg)
   end function

   !> g_timer label of phase idx (0-based) as a Fortran string
   function rsrec_phase_label_f(idx) result(label)
      integer, intent(in) :: idx
      character(len=:), allocatable :: label
      character(kind=c_char), pointer :: p(:)
      type(c_ptr) :: cp
      integer :: n
      cp = rsrec_phase_label(int(idx, c_int))
      label = ''
      if (.not. c_associated(cp)) return
      call c_f_pointer(cp, p, [64])
      n = 0
      do while (n < 64)
         if (p(n + 1) == c_null_char) exit
         n = n + 1
      end do
      allocate (character(len=n) :: label)
      if (n > 0) label = transfer(p(1:n), label)
   end function

   !> maps a non-zero status to the reference's fatal convention (logger.f90:186-193)
   subroutine rsrec_check(rc, file, line)
      use logger_mod, only: g_logger
      integer(c_int), intent(in) :: rc
      character(len=*), intent(in) :: file
      integer, intent(in) :: line
      if (rc /= RSREC_OK) call g_logger%fatal('rsrec: '//rsrec_last_error_f(), file, line)
   end subroutine

end module rsrec_c_mod

/*
 * rsrec.h -- C ABI of the B200-native recursion engine (librsrec.so).
 *
 * Drop-in boundary for the hot path of rslmtoasa/rslmtoasa: the type-bound procedures of
 * `type recursion` (reference source/recursion.f90:41-116).  The reference has no FFI; a Fortran host binds these
 * entry points with ISO_C_BINDING (see INTEGRATION.md / fortran/rsrec_c_mod.f90) and keeps its derived-type API.
 *
 * Conventions (identical to the reference's arrays, so the host passes its members unchanged):
 *   - every array is Fortran column-major; `rsrec_cplx` == complex(c_double_complex) == complex(rp);
 *   - site and type indices are 1-based, `nn(i,1)` is the slot count, `nn(i,j)=0` means "no neighbour";
 *   - every function returns 0 on success, <0 on error; rsrec_last_error() gives the message the host passes to
 *     g_logger%fatal (reference logger.f90:186-193).  RSREC_EDIVERGED mirrors the Chebyshev guard
 *     `sum(real(mu)) > 1000 -> fatal` (recursion.f90:2594, 2760);
 *   - calls are synchronous on return (results are in the host arrays); a handle is not thread-safe; one handle
 *     per process/GPU, like one `recursion` object per MPI rank in the reference (mpi.f90:32-58).
 *   - there is NO CPU fallback: if no CUDA device is usable rsrec_create fails with RSREC_ECUDA.
 */
#ifndef RSREC_H
#define RSREC_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
typedef struct { double re, im; } rsrec_cplx;
#else
#include <complex.h>
typedef double _Complex rsrec_cplx;
#endif

typedef struct rsrec_handle_s *rsrec_handle;

#define RSREC_OK 0
#define RSREC_EINVAL (-1)    /* bad argument / call order */
#define RSREC_EDIVERGED (-2) /* Chebyshev moments did not converge (energy window too small) */
#define RSREC_ECUDA (-3)     /* CUDA runtime error, or no device */
#define RSREC_ENOMEM (-4)    /* device memory exhausted */

const char *rsrec_last_error(void);
/* library/ABI version, and the SM architecture the kernels were compiled for (100 => sm_100a) */
int rsrec_version(void);
int rsrec_compiled_arch(void);
/* CUDA devices visible to this process (0 when none): an MPI host picks device_ordinal = mod(local_rank, count) */
int rsrec_device_count(void);

/* recursion constructor (recursion.f90:132-143, allocation 3713-3826): sizes come from
 * lattice%kk, size(lattice%nn,2), size(hamiltonian%ee,3) = maxval(nn(:,1))+1 (hamiltonian.f90:294),
 * lattice%ntype, lattice%nmax. */
int rsrec_create(rsrec_handle *h, int device_ordinal, int kk, int ncols, int nslot, int ntype, int nmax);
int rsrec_destroy(rsrec_handle h);

/* lattice%nn(kk,ncols), lattice%iz(kk) (lattice.f90:1856-1860, 204).  Upload after the Hamiltonian build because
 * chbar_nc may zero entries of nn (hamiltonian.f90:2350-2352). */
int rsrec_set_lattice(rsrec_handle h, const int32_t *nn, const int32_t *iz);

/* Optional: lattice%cr (3,kk) (lattice.f90:239).  Coordinates never enter the arithmetic; they only order the work
 * (8-site tiles along a Morton curve) so that gathers of concurrently running CTAs hit L2.  Results are bit-identical
 * with or without this call; NULL clears.  Call before the first recursion (it re-sorts the tile tables). */
int rsrec_set_positions(rsrec_handle h, const double *cr);

/* hamiltonian%{ee,eeo}(18,18,nslot,ntype), {hall,hallo}(18,18,nslot,nmax), {lsham,enim}(18,18,ntype), hoh
 * (hamiltonian.f90:52-66,294-301).  Called once per SCF iteration after build_bulkham/build_locham
 * (self.f90:777-797).  eeo/hallo/enim may be NULL when hoh==0; hall/hallo may be NULL when nmax==0. */
int rsrec_set_hamiltonian(rsrec_handle h, const rsrec_cplx *ee, const rsrec_cplx *eeo, const rsrec_cplx *hall,
                          const rsrec_cplx *hallo, const rsrec_cplx *lsham, const rsrec_cplx *enim, int hoh);

/* Kubo operators hamiltonian%{v_a,vo_a} (slot 'a') / {v_b,vo_b} (slot 'b'), (18,18,nslot,ntype)
 * (recursion.f90:242-262).  vo_op may be NULL when hoh==0. */
int rsrec_set_operator(rsrec_handle h, int slot, const rsrec_cplx *v_op, const rsrec_cplx *vo_op);

/* recur_b (recursion.f90:1807-1866) when site_j==NULL or site_j[u]==0, recur_b_ij (1655-1737) otherwise:
 * unit u starts from asign[u]*I on site_i[u] and bsign[u]*I on site_j[u] (NULL signs => 1).  crecal_b
 * (1873-1973) runs lld-1 steps.  a_b, b2_b: (18,18,lld,nunits); b2_b holds B^2 (not B) like the reference. */
int rsrec_lanczos_block(rsrec_handle h, int nunits, const int32_t *site_i, const int32_t *site_j,
                        const rsrec_cplx *asign, const rsrec_cplx *bsign, int lld, rsrec_cplx *a_b, rsrec_cplx *b2_b);

/* recur/crecal/hop (recursion.f90:3485-3532, 3423-3478, 3310-3416; nsp=1 only, like the reference).
 * a, b2: real (lld,18,nunits) = this%a(:, :, unit, 1), this%b2(:, :, unit, 1). */
int rsrec_lanczos_scalar(rsrec_handle h, int nunits, const int32_t *sites, int lld, double *a, double *b2);

/* zsqr (recursion.f90:1980-2023): b2_b(18,18,lld,na) <- (b2_b)^(1/2), in place. */
int rsrec_zsqr(rsrec_handle h, rsrec_cplx *b2_b, int lld, int na);

/* chebyshev_recur (recursion.f90:3057-3130) / chebyshev_recur_ij (2376-2487): a_scale, b_shift are the caller's
 * a=(emax-emin)/(2-0.3), b=(emax+emin)/2 (3078-3079).  mu_n: (18,18,2*lld+2,nunits). */
int rsrec_cheb_moments(rsrec_handle h, int nunits, const int32_t *site_i, const int32_t *site_j,
                       const rsrec_cplx *asign, const rsrec_cplx *bsign, int lld, double a_scale, double b_shift,
                       rsrec_cplx *mu_n);

/* The same moment recursion (cheb_0th_mom, cheb_1st_mom, chebyshev_recur_ll) started from the KPM random-phase
 * block exp(2 pi i u_k) I / sqrt(kk) of recursion.f90:1131-1143; phases u: (kk,nvec) from the host because the
 * reference's random_seed() is not repeatable.  mu_n: (18,18,2*lld+2,nvec). */
int rsrec_cheb_moments_random(rsrec_handle h, int nvec, const double *phases, int lld, double a_scale,
                              double b_shift, rsrec_cplx *mu_n);

/* compute_moments_stochastic (recursion.f90:979-1234).  start_kind 0: per_type, start_sites[i] = atlist(i);
 * 1: random_vec, phases (kk,nstart).  mu_nm: (18,18,cond_ll,cond_ll,nstart) = mu_nm_stochastic. */
int rsrec_kubo_moments(rsrec_handle h, int nstart, int start_kind, const int32_t *start_sites,
                       const double *phases, int cond_ll, double a_scale, double b_shift, rsrec_cplx *mu_nm);

/* ham_vec_matmul / ham_hoh_vec_matmul (recursion.f90:913-977 / 785-911): psi_out = (H psi_in - b psi_in)/a,
 * psi_*: (18,18,kk) host arrays.  velo_vec_matmul / velo_hoh_vec_matmul (587-783): psi_out = v_op psi_in. */
int rsrec_ham_vec_matmul(rsrec_handle h, const rsrec_cplx *psi_in, rsrec_cplx *psi_out, double a_scale, double b_shift);
int rsrec_velo_vec_matmul(rsrec_handle h, int slot, const rsrec_cplx *psi_in, rsrec_cplx *psi_out);

/* create_ll_map (recursion.f90:3277-3303) with the start mask of chebyshev_recur (izeroll(site,1) = 1, 3086-3087):
 * izeroll (0:kk, lld+1) int32, column ll+1 = sites reachable within ll applications of H (what izero holds after
 * ll hops; the library itself never materialises the masks -- inactive sites hold exact zeros). */
int rsrec_create_ll_map(rsrec_handle h, int site, int lld, int32_t *izeroll);

/* chebyshev_orbital_mod, moment part (recursion.f90:2901-3008): mu_n_orb (18,18,lld) = sum over start_sites of
 * L_r^H T_{n-1}(H~)|r> with |L_r> = i (Y H~ X - X H~ Y)|r>; cr (3,kk) = lattice%cr, alat = lattice%alat.  The
 * reference loops over all kk sites and divides by kk; the Jackson weighting / trace integration (3012-3048) stay
 * with the caller. */
int rsrec_orbital_moments(rsrec_handle h, int nstart, const int32_t *start_sites, const double *cr, double alat,
                          int lld, double a_scale, double b_shift, rsrec_cplx *mu_n_orb);

/* ---- device-side assembly of the block sets: build_bulkham / build_locham with chbar_nc's
 * orbital part, ham0m_nc, hcpx, build_obarm, build_enim (hamiltonian.f90:1481-1667, 2225-2369; math.f90:1508-1577).
 * Replaces rsrec_set_hamiltonian: the sets the recursion kernels read are produced on the device from
 *   hhh  (9,9,nslot,ncls) real : hhh(ilm,jlm) = real(sbar(jlm,ilm,m,num(ia))) as hmfind returns it, for slot m of class c
 *                                (classes: atom types 1..ntype through atlist, then local sites 1..nmax);
 *   jt   (nslot,ncls) int32    : type iz of the atom in slot m (slot 1 = the atom itself), 0 = no neighbour -> zero block;
 *   it   (ncls) int32          : type of the class's own atom;
 *   pot  (9,12,ntype) complex  : wx0, wx1, cx0, cx1, cex0, cex1, obx0, obx1, cx(:,1), cx(:,2), cex(:,1), cex(:,2)
 *                                of symbolic_atoms(type)%potential (potential.f90:56-65);
 *   mom  (3,ntype), lsham (18,18,ntype) (build_lsham stays on the host: constants).
 * Optional downloads (NULL to skip) in the reference's shapes: ee, eeo (18,18,nslot,ntype), hall, hallo
 * (18,18,nslot,nmax), enim, obarm (18,18,ntype) for the host modules that still read them. */
int rsrec_build_hamiltonian(rsrec_handle h, const double *hhh, const int32_t *jt, const int32_t *it,
                            const rsrec_cplx *pot, const double *mom, const rsrec_cplx *lsham, int hoh, rsrec_cplx *ee,
                            rsrec_cplx *eeo, rsrec_cplx *hall, rsrec_cplx *hallo, rsrec_cplx *enim, rsrec_cplx *obarm);

/* hamiltonian%rotate_to_local_axis / rotate_from_local_axis (hamiltonian.f90:2442-2484; rotmag_loc, ROTMAT, DSs, car2sph
 * of math.f90:1981-2192): the device-resident sets ee, hall (and eeo, hallo, enim when hoh) become R^H X R with R the
 * rotation that takes m_loc to the z axis, always starting from the sets as built (the reference's *_glob copies); lsham
 * is not rotated, like the reference.  No host round trip: the rotated sets exist only on the device. */
int rsrec_rotate_to_local_axis(rsrec_handle h, const double *m_loc /* (3) */);
int rsrec_rotate_from_local_axis(rsrec_handle h);
/* recur_b with hamiltonian%local_axis set (recursion.f90:1826-1832): unit u runs on the sets rotated to mom(:,u)
 * (mom (3,nunits) = symbolic_atoms(i)%potential%mom); the sets stay rotated to the last unit's axis, like the reference. */
int rsrec_lanczos_block_local_axis(rsrec_handle h, int nunits, const int32_t *site_i, const double *mom, int lld,
                                   rsrec_cplx *a_b, rsrec_cplx *b2_b);

/* ---- neighbour table on the device: lattice%nncal + lattice%remd (lattice.f90:3035-3123,
 * 2823-2907) with a cell grid instead of the O(kk^2) pair loop; identical table (integers, bit-exact).
 * crd (3,kk) = cr*alat; no (kk) = lattice%num (bravais type of each site); iu (ntot) = representative site of each
 * bravais type; ct = lattice%ct(1); pbc[3] = b1,b2,b3 flags (NULL = open cluster), nrep[3] = n1,n2,n3, a (3,3) =
 * lattice%a, alat.  nn (kk, ncols) out, ncols >= *nm_out + 1 like lattice%nn (call with nn = NULL to query nm). */
int rsrec_build_nn(int device_ordinal, int kk, const double *crd, const int32_t *no, int ntot, const int32_t *iu,
                   double ct, const int32_t *pbc, const int32_t *nrep, const double *a, double alat, int ncols,
                   int32_t *nn, int *nm_out);

/* ---- consumers either side of the recursion: the reference's green / density_of_states /
 * conductivity back ends that read a_b, b2_b, a, b2, mu_n, mu_nm_stochastic.  Same array shapes as the reference. ---- */

/* bpopt (recursion.f90:3540-3581, with emami 3589-3706) for `nchains` independent chains: a, rb (ll,nchains);
 * ainf, rbinf, ifail (nchains; ifail may be NULL).  Bit-identical to the reference's non-contracted arithmetic. */
int rsrec_bpopt(rsrec_handle h, int nchains, int ll, const double *a, const double *rb, double *ainf, double *rbinf,
                int *ifail);

/* get_terminf / get_cinf (recursion.f90:2092-2138 / 2030-2086): a_b, b_b (18,18,ll,na) with b_b = B (after zsqr) as
 * block_green passes them; a_inf, b_inf (18,18,na) real; a_inf0, b_inf0 (na), may be NULL. */
int rsrec_get_terminf(rsrec_handle h, const rsrec_cplx *a_b, const rsrec_cplx *b_b, int na, int ll, double *a_inf,
                      double *b_inf, double *a_inf0, double *b_inf0);

/* bgreen (green.f90:1191-1339) for one unit: a_b, b_b (18,18,ll); e (nv); channels ie_start..ie_start+ie_len-1
 * (1-based) of g_out (18,18,nv) are computed, the others are zero; a_inf, b_inf (18,18) real; eta complex. */
int rsrec_bgreen(rsrec_handle h, const rsrec_cplx *a_b, const rsrec_cplx *b_b, int ll, const double *e, int nv,
                 int ie_start, int ie_len, const double *a_inf, const double *b_inf, double eta_re, double eta_im,
                 int sym_term, rsrec_cplx *g_out);

/* block_green (green.f90:588-621): get_terminf + bgreen(eta = 0, all channels) for na units; g0 (18,18,nv,na). */
int rsrec_block_green(rsrec_handle h, const rsrec_cplx *a_b, const rsrec_cplx *b_b, int na, int ll, const double *e,
                      int nv, int sym_term, rsrec_cplx *g0);

/* chebyshev_green (green.f90:1030-1108): mu_n (18,18,2*lld+2,na) -> mu_ng (same shape, Jackson-weighted; may be
 * NULL) and g0 (18,18,nv,na); energy_min/max are energy%energy_min/max. */
int rsrec_chebyshev_green(rsrec_handle h, const rsrec_cplx *mu_n, int na, int lld, const double *ene, int nv,
                          double energy_min, double energy_max, rsrec_cplx *mu_ng, rsrec_cplx *g0);

/* dos%density with bprldos (density_of_states.f90:248-407) for every (atom, direction): a, b2 (lld,18,na,nmdir) =
 * recursion%a/b2(:, :, :, 1:nmdir); dw_l, cshi (18,na); tdens (18,nv,na,nmdir). */
int rsrec_density(rsrec_handle h, const double *a, const double *b2, int lld, int na, int nmdir, const double *ene,
                  int nv, const double *dw_l, const double *cshi, double *tdens);

/* sgreen (green.f90:628-705): g0 (18,18,nv,na) from the scalar-recursion coefficients; nmdir 1 or 3. */
int rsrec_sgreen(rsrec_handle h, const double *a, const double *b2, int lld, int na, int nmdir, const double *ene,
                 int nv, const double *dw_l, const double *cshi, rsrec_cplx *g0);

/* calculate_gamma_nm + the energy integrand of calculate_conductivity_tensor (conductivity.f90:158-306):
 * mu_nm (18,18,M,M,nloop) = mu_nm_stochastic; integrand (18,nv) = integrand(l2,l2,:) summed over the loop index;
 * integrand_at (18,nv,nloop) per type when per_type != 0 (zeros otherwise; may be NULL). */
int rsrec_conductivity_integrand(rsrec_handle h, const rsrec_cplx *mu_nm, int M, int nloop, const double *ene, int nv,
                                 double energy_min, double energy_max, int per_type, rsrec_cplx *integrand,
                                 rsrec_cplx *integrand_at);

/* Tail of calculate_conductivity_tensor (conductivity.f90:300-372): for every mesh energy i the T = 0 Fermi-weighted
 * Simpson integral (simpson_f, math.f90:1600-1632, kBT = 1e-15, E_F = wscale(i)) of the integrand -- the sigma(E_F)
 * curves written to cond_total*.out / <symbol>_cond*.out.  The reference's O(nv^2) loop (one exp per term) is a running
 * sum: one pass per series, added in the reference's order (bit-identical).  integrand (18,nv), integrand_at
 * (18,nv,nat) from rsrec_conductivity_integrand / rsrec_kubo_conductivity (nat = 0: no per-type output); nv1 = en%nv1;
 * wstep = wscale(2) - wscale(1); sigma (2,19,nv,1+nat): (re|im, total then orbital l2 = 1..18, energy, summed | type);
 * the summed block is divided by real(loop_over) as in the file output, the per-type blocks are not. */
int rsrec_conductivity_cumulative(rsrec_handle h, const rsrec_cplx *integrand, const rsrec_cplx *integrand_at, int nv,
                                  int nv1, int nat, double wstep, int loop_over, double *sigma);

/* ---- fused entry points: a recursion and its consumer with the coefficients staying on the device ---- */

/* run_recursion + run_dos of the block path (self.f90:799-856): recur_b -> zsqr -> get_terminf -> bgreen(eta=0).
 * a_b, b2_b (18,18,lld,nunits), b2_b = B^2 as recur_b leaves it (either may be NULL); g0 (18,18,nv,nunits), may be
 * NULL: g0 then only stays on the device for the rsrec_bands_* consumers below. */
int rsrec_recur_b_green(rsrec_handle h, int nunits, const int32_t *site_i, int lld, const double *ene, int nv,
                        int sym_term, rsrec_cplx *a_b, rsrec_cplx *b2_b, rsrec_cplx *g0);

/* chebyshev_recur (recursion.f90:3057-3130) + chebyshev_green (green.f90:1030-1108); mu_n, mu_ng, g0 may be NULL. */
int rsrec_cheb_recur_green(rsrec_handle h, int nunits, const int32_t *site_i, int lld, double energy_min,
                           double energy_max, const double *ene, int nv, rsrec_cplx *mu_n, rsrec_cplx *mu_ng,
                           rsrec_cplx *g0);

/* The same two fused drivers for pair start vectors (the exchange path): recur_b_ij (recursion.f90:1655-1737) + zsqr +
 * block_green_ij (green.f90:356-384), and chebyshev_recur_ij (2376-2487) + chebyshev_green_ij (892-952).  Units as in
 * rsrec_lanczos_block / rsrec_cheb_moments (site_i, site_j, asign, bsign). */
int rsrec_recur_b_ij_green(rsrec_handle h, int nunits, const int32_t *site_i, const int32_t *site_j,
                           const rsrec_cplx *asign, const rsrec_cplx *bsign, int lld, const double *ene, int nv,
                           int sym_term, rsrec_cplx *a_b, rsrec_cplx *b2_b, rsrec_cplx *g0);
int rsrec_cheb_recur_ij_green(rsrec_handle h, int nunits, const int32_t *site_i, const int32_t *site_j,
                              const rsrec_cplx *asign, const rsrec_cplx *bsign, int lld, double energy_min,
                              double energy_max, const double *ene, int nv, rsrec_cplx *mu_n, rsrec_cplx *mu_ng,
                              rsrec_cplx *g0);

/* calculate_intersite_gf (green.f90:425-469) on the g0 the last Green-function call left on the device for the pair
 * units: gij, gji (18,18,nv,njij) = ((g1 - g2) +- (g3 - g4)/i)/2 (g1 alone when i == j), and their spin
 * decomposition gspin (9,9,nv,njij,8) = Ginmag, Gix, Giy, Giz, Gjnmag, Gjx, Gjy, Gjz (may be NULL).
 * compact != 0: the units are packed (one unit for an i == j pair, four otherwise -- what the drivers above are given);
 * compact == 0: four slots per pair as in recursion%a_b(:,:,:,4*njij) (slots 2-4 of an i == j pair are ignored). */
int rsrec_intersite_gf(rsrec_handle h, int njij, const int32_t *pair_i, const int32_t *pair_j, int compact,
                       rsrec_cplx *gij, rsrec_cplx *gji, rsrec_cplx *gspin);

/* compute_moments_stochastic + calculate_gamma_nm + integrand of calculate_conductivity_tensor: only the diagonals
 * mu_nm(l,l,n,m,i) the integrand consumes are kept, on the device; mu_nm (18,18,M,M,nstart) is downloaded only when
 * non-NULL.  start_kind 0 = per_type (integrand_at filled), 1 = random_vec. */
int rsrec_kubo_conductivity(rsrec_handle h, int nstart, int start_kind, const int32_t *start_sites,
                            const double *phases, int M, double energy_min, double energy_max, const double *ene,
                            int nv, rsrec_cplx *mu_nm, rsrec_cplx *integrand, rsrec_cplx *integrand_at);

/* ---- `type bands` (bands.f90): what the SCF loop takes from g0 -- total DOS, Fermi level, band moments, charges ----
 * Every Green-function entry point above (block_green, chebyshev_green, sgreen and the fused ones) leaves its
 * g0 (18,18,nv,nunits) on the device; these calls consume that copy, so a fused call with g0 = NULL followed by them
 * moves only O(nv) + O(nunits) numbers to the host.  rsrec_bands_set_g0 uploads a g0 computed elsewhere.
 * g0 is kept only while it leaves half of the free device memory to the recursion (or up to RSREC_G0_RESIDENT_MAX_MB
 * when that environment variable is set); beyond that the Green-function calls stream it to the host per batch of
 * units as before, refuse g0 = NULL with RSREC_ENOMEM, and the calls below report that no g0 is resident. */
int rsrec_bands_set_g0(rsrec_handle h, const rsrec_cplx *g0, int nunits, int nv);
int rsrec_bands_get_g0(rsrec_handle h, rsrec_cplx *g0);
int rsrec_bands_g0_shape(rsrec_handle h, int *nunits, int *nv);

/* DOS part of calculate_fermi (bands.f90:260-273): dtot (nv) summed over this handle's units in the reference's
 * order (bit-exact; the caller all-reduces it across ranks like bands.f90:276), dosia (nv,nunits) and dosial
 * (18,nv,nunits) (either may be NULL). */
int rsrec_bands_dos(rsrec_handle h, double *dtot, double *dosia, double *dosial);

/* Fermi level of calculate_fermi (bands.f90:322-342) from the all-reduced dtot (nv = channels_ldos + 10):
 * fix_fermi == 0: the two `fermi` scans (366-402), fermi in/out = en%fermi, nv1 in/out = en%ik1 / bands%nv1, e1 out =
 * bands%e1, ifail (may be NULL) = 1 when the valence charge qqv is not reached; fix_fermi != 0: nv1 and e1 of the
 * fixed level (338-341), dtot unused. */
int rsrec_bands_fermi(rsrec_handle h, const double *dtot, int nv, double edel, double energy_min, double qqv,
                      int fix_fermi, double *fermi, int *nv1, double *e1, int *ifail);

/* calculate_magnetic_moments (bands.f90:791-855) with calculate_projected_dos (1158-1181): mom0 (3,nunits) = mx,my,mz
 * and mom1 (3,nunits) = potential%mom1, the Simpson integrals (simpson_m, math.f90:1579-1598) of order 0 and 1 of
 * dx,dy,dz up to the Fermi level; ene (nv) = en%ene.  The normalisation to mtot/mom is the caller's (scalar). */
int rsrec_bands_magnetic_moments(rsrec_handle h, const double *ene, double edel, double fermi, int nv1, double e1,
                                 double *mom0, double *mom1);

/* calculate_moments (bands.f90:409-524) with calculate_orbital_moments (1075-1156): mom (3,nunits) = potential%mom;
 * occ (3,6,nunits) = sgef, pmef, smef of channel i = l + 3(isp-1) (from which the caller forms ql and gravity_center,
 * 492-495); lmom (3,nunits) = potential%lmom. */
int rsrec_bands_moments(rsrec_handle h, int channels_ldos, const double *ene, double edel, double fermi, int nv1,
                        double e1, const double *mom, double *occ, double *lmom);

/* calculate_band_energy (bands.f90:354-359): eband = simpson_m(order 1) of the all-reduced dtot. */
int rsrec_bands_band_energy(rsrec_handle h, const double *dtot, int nv, const double *ene, double edel, double fermi,
                            int nv1, double e1, double *eband);

/* ---- device-resident stepping (what bench.py times as `value`; the calls above are the `e2e` path) ----
 * begin: upload start vectors, compute mu(1), mu(2) on the device.  run_steps: enqueue n chebyshev_recur_ll steps on
 * the handle's stream without host synchronisation.  end: wait, download mu_n(18,18,2*lld+2,nvec). */
int rsrec_cheb_begin_random(rsrec_handle h, int nvec, const double *phases, int lld, double a_scale, double b_shift);
int rsrec_cheb_begin_sites(rsrec_handle h, int nunits, const int32_t *site_i, const int32_t *site_j,
                           const rsrec_cplx *asign, const rsrec_cplx *bsign, int lld, double a_scale, double b_shift);
int rsrec_cheb_run_steps(rsrec_handle h, int nsteps);
int rsrec_cheb_end(rsrec_handle h, rsrec_cplx *mu_n);
int rsrec_synchronize(rsrec_handle h);
/* the cudaStream_t all kernels of this handle are launched on (so callers can record events on it) */
void *rsrec_stream(rsrec_handle h);
/* kernels launched by this handle since creation (bench.py's gpu_launches) */
long long rsrec_launch_count(rsrec_handle h);
/* how many of the SpMV launches ran the spin-diagonal two-block kernel (collinear hopping blocks; RSREC_NO_SPIN_DIAG=1
 * in the environment disables it) */
long long rsrec_spin_diag_launch_count(rsrec_handle h);
/* host<->device bytes this handle has moved so far (uploads of lattice/Hamiltonian/start data, downloads of results) */
long long rsrec_h2d_bytes(rsrec_handle h);
long long rsrec_d2h_bytes(rsrec_handle h);
/* per-kernel timing of the gather-SpMV launches with CUDA events on the handle's stream: enable, run, read
 * (read synchronises, returns the summed duration and the number of launches, and resets the counters) */
int rsrec_profile(rsrec_handle h, int enable);
int rsrec_profile_read(rsrec_handle h, double *total_ms, int *nlaunches);
/* Per-phase DEVICE timing under the reference's own g_timer labels (recursion.f90:1902-1970 crecal_b: 'H|PSI_n>',
 * 'H|Psi_n-A_n|Psi_n-B_n|Psi_n-1', 'B_n+1', '<PSI|B_n+1|PSI>'; 3104-3127 / 2454-2478 chebyshev_recur(_ij):
 * '<PSI_0|PSI_0>', '<PSI_0|PSI_1>', '<PSI_0|PSI_n>'), so the host's profile tree keeps its inner entries:
 * rsrec_phase_timing(h,1) starts recording CUDA-event pairs around every phase on the handle's stream (0 stops and
 * resets); rsrec_phase_read synchronises, fills ms[rsrec_phase_count()] / calls[..] (summed since the last read) and
 * resets.  rsrec_phase_label(i) is the g_timer label of entry i. */
int rsrec_phase_timing(rsrec_handle h, int enable);
int rsrec_phase_count(void);
const char *rsrec_phase_label(int idx);
int rsrec_phase_read(rsrec_handle h, double *ms, long long *calls);
/* host wall-clock seconds of the stages of the unit-sharded calls since the last read: seconds[5] = tables (lattice /
 * Hamiltonian export), plan (active-region levels + unit upload), recursion, exchange (NCCL), download.  Stage boundaries
 * synchronise the stream only while rsrec_phase_timing is on. */
int rsrec_host_phase_read(rsrec_handle h, double *seconds);

/* ---- the exchange step of the unit-sharded path ------------------------------------------------
 * The reference shards independent units (recursion sites, pair vectors, random vectors) over MPI ranks with
 * get_mpi_variables (mpi.f90:32-58) and sums afterwards with MPI_ALLREDUCE (bands.f90:270-275, self.f90:887).  One process
 * per GPU: every rank creates its handle, rank 0 calls rsrec_comm_unique_id and broadcasts the 128 bytes with whatever the
 * host already has (MPI_Bcast in the Fortran host), then every rank calls rsrec_comm_init.  NCCL (dlopen'ed libnccl.so.2)
 * then runs on the handle's stream directly on the device-resident results.  With a communicator attached:
 *   - rsrec_bands_dos all-reduces dtot on the device before returning it (bands.f90:276);
 *   - rsrec_kubo_conductivity with random vectors (start_kind 1) all-reduces the integrand over the ranks' vector shards
 *     (nstart = 0 is then legal on a rank that owns no vector);
 *   - rsrec_cheb_moments_random_sum / rsrec_lanczos_block_sharded below exchange their results on the device.
 * Without a communicator every call below degenerates to the single-rank result. */
#define RSREC_COMM_ID_BYTES 128
int rsrec_comm_unique_id(unsigned char *id128);
int rsrec_comm_init(rsrec_handle h, int nranks, int rank, const unsigned char *id128);
int rsrec_comm_destroy(rsrec_handle h);
int rsrec_comm_info(rsrec_handle h, int *nranks, int *rank, int *nccl_version);
/* get_mpi_variables (mpi.f90:32-58): 1-based inclusive start_atom..end_atom of `rank` for nunits units */
int rsrec_shard_range(int rank, int nranks, int nunits, int *first, int *last);
/* MPI_ALLREDUCE(MPI_IN_PLACE, buf, count, <type>, MPI_SUM) of a host array: dtype 0 real(rp), 1 complex(rp), 2 integer */
int rsrec_allreduce(rsrec_handle h, void *buf, long long count, int dtype);
/* all-gather of per-unit host results in global unit order (the MPI_Allgather recursion.f90:1788-1799 leaves commented
 * out): local = this rank's block-rule shard, full = all nunits_total units, doubles_per_unit reals each */
int rsrec_allgather_units(rsrec_handle h, const void *local, void *full, long long doubles_per_unit, int nunits_total);
/* recur_b / recur_b_ij over ALL units of the job: each rank runs its shard, a_b/b2_b (18,18,lld,nunits_total) are
 * gathered on the device and returned on every rank */
int rsrec_lanczos_block_sharded(rsrec_handle h, int nunits_total, const int32_t *site_i, const int32_t *site_j,
                                const rsrec_cplx *asign, const rsrec_cplx *bsign, int lld, rsrec_cplx *a_b, rsrec_cplx *b2_b);
/* stochastic-trace KPM: moments of this rank's random vectors (phases (kk,nvec_local)) summed on the device, all-reduced,
 * mu_sum (18,18,2*lld+2) = sum over all vectors of the job */
int rsrec_cheb_moments_random_sum(rsrec_handle h, int nvec_local, const double *phases, int lld, double a_scale,
                                  double b_shift, rsrec_cplx *mu_sum);

/* the 18x18 reductions of a step inside the SpMV kernel (tensor pipeline): lanczos (default 1: A = sum psi^H H psi of hop_b),
 * cheb (default 0: D1, D2 of chebyshev_recur_ll; measured slower than the separate Gram kernel); -1 leaves a
 * setting unchanged.  Results agree to rounding either way. */
int rsrec_set_fusion(rsrec_handle h, int lanczos, int cheb);
/* select kernel family: 0 = SIMT reference kernels, 1 = DMMA (FP64 tensor core) pipeline (default) */
int rsrec_set_kernel_family(rsrec_handle h, int family);

#ifdef __cplusplus
}
#endif
#endif /* RSREC_H */

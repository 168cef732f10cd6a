/*
 * rsrec_oracle.h -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * A plain-C restatement, loop for loop, of the recursion hot path of rslmtoasa/rslmtoasa
 * (`source/recursion.f90`).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library; the product (librsrec.so) never does.
 *
 * PARITY PINNED by the reference's own golden fixtures for the SCF chain: oracle/ref_bccfe.py rebuilds the reference's
 * bccFe regression case (tests/scf/references/Example_bulk_bccFe_<case>/ref.json, 12 cases: nsp 1/2/4, hoh on/off, block
 * and Chebyshev recursion, lld 16/21, two energy windows) from its input files and reproduces every stored
 * totaldos.out value to all printed digits through this oracle (tests/test_reference_golden.py) and through the
 * CUDA library (tests/test_gpu_reference_golden.py); oracle/ref_fccpt.py does the same for the Kubo-Bastin path with the
 * stored Pt_cond.out curves of tests/postproc/references/Example_exchange_conductivity_fccPt{,_hoh} (kubo moments, velocity
 * products, Gamma contraction, Fermi-weighted tail), and oracle/ref_exchange.py for the pair path (recur_b_ij,
 * calculate_intersite_gf) with the stored J_ij of tests/postproc/references/Example_exchange_bccFe{,_hoh}.  Routines no
 * reference fixture reaches (scalar recursion, site-indexed `hall` region, orbital moments, random-vector starts) are pinned by the independent dense
 * numpy restatement (oracle/dense_check*.py) and by invariants (tests/test_oracle*.py).  The reference itself cannot
 * be compiled here (no Fortran compiler), so there is no oracle/_ref.
 *
 * All arrays are Fortran column-major exactly as the reference holds them; site / type indices are 1-based.
 */
#ifndef RSREC_ORACLE_H
#define RSREC_ORACLE_H
#include <stdint.h>
#include <complex.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef double _Complex orc_cplx;

typedef struct orc_ctx orc_ctx;

/* lattice%{kk,nn,iz,nmax}, hamiltonian%{ee,eeo,hall,hallo,lsham,enim,hoh}; pointers are borrowed. */
orc_ctx *orc_create(int kk, int ncols, int nslot, int ntype, int nmax,
                    const int32_t *nn, const int32_t *iz);
void orc_destroy(orc_ctx *c);
void orc_set_hamiltonian(orc_ctx *c, const orc_cplx *ee, const orc_cplx *eeo, const orc_cplx *hall,
                         const orc_cplx *hallo, const orc_cplx *lsham, const orc_cplx *enim, int hoh);
void orc_set_operator(orc_ctx *c, int slot /* 'a' | 'b' */, const orc_cplx *v_op, const orc_cplx *vo_op);
/* use_mask=1 (default): the reference's izero/idum/irlist logic; 0: every site always active. */
void orc_set_use_mask(orc_ctx *c, int use_mask);
void orc_set_threads(int nthreads);
int orc_get_max_threads(void);

/* recur_b (site_j[u]==0) / recur_b_ij (start = asign*I on site_i, bsign*I on site_j): recursion.f90:1807-1866,
 * 1655-1737 + crecal_b 1873-1973.  a_b, b2_b: 18x18xlldxnunits. */
int orc_lanczos_block(orc_ctx *c, int nunits, const int32_t *site_i, const int32_t *site_j,
                      const orc_cplx *asign, const orc_cplx *bsign, int lld, orc_cplx *a_b, orc_cplx *b2_b);
/* recur (scalar, nsp=1): recursion.f90:3485-3532.  a, b2: lld x 18 x nunits (real). */
int orc_lanczos_scalar(orc_ctx *c, int nunits, const int32_t *sites, int lld, double *a, double *b2);
/* chebyshev_recur / chebyshev_recur_ij: recursion.f90:3057-3130, 2376-2487.  mu_n: 18x18x(2lld+2)xnunits.
 * Returns -2 if the reference's divergence guard (sum Re mu > 1000) would have called fatal. */
int orc_cheb_moments(orc_ctx *c, int nunits, const int32_t *site_i, const int32_t *site_j,
                     const orc_cplx *asign, const orc_cplx *bsign, int lld, double a, double b, orc_cplx *mu_n);
/* same recursion, KPM random-phase start block exp(2 pi i u_k) I / sqrt(kk) on every site
 * (start vector of recursion.f90:1131-1143 fed to the chebyshev_recur_ll loop); phases: kk x nvec. */
int orc_cheb_moments_random(orc_ctx *c, int nvec, const double *phases, int lld, double a, double b,
                            orc_cplx *mu_n);
/* timing hook for the CPU baseline: seconds spent in `nsteps` chebyshev_recur_ll steps of one random vector */
int orc_cheb_time_steps(orc_ctx *c, const double *phases, int nsteps, double a, double b, double *seconds);
/* compute_moments_stochastic: recursion.f90:979-1234.  start_kind 0 = per_type (start_sites[i] = atlist(i)),
 * 1 = random_vec (phases kk x nstart).  mu_nm: 18x18xMxMxnstart. */
int orc_kubo_moments(orc_ctx *c, int nstart, int start_kind, const int32_t *start_sites, const double *phases,
                     int cond_ll, double a, double b, orc_cplx *mu_nm);
/* the same chains, contracted only against the left indices msel[0..nsel-1] (1-based): mu_sel(18,18,M,nsel,nstart) */
int orc_kubo_moments_cols(orc_ctx *c, int nstart, int start_kind, const int32_t *start_sites, const double *phases,
                          int M, double a, double b, const int32_t *msel, int nsel, orc_cplx *mu_sel);
/* zsqr: recursion.f90:1980-2023, in place on b2_b(18,18,lld,na). */
int orc_zsqr(orc_cplx *b2_b, int lld, int na);
/* single operator applications (ham_vec_matmul 913-977, ham_hoh_vec_matmul 785-911, velo_* 587-783);
 * izero (kk+1 ints, index 0 unused/0) is updated to idum on return like the callers do. */
void orc_ham_vec_matmul(orc_ctx *c, const orc_cplx *psi_in, orc_cplx *psi_out, double a, double b, int32_t *izero);
void orc_velo_vec_matmul(orc_ctx *c, int slot, const orc_cplx *psi_in, orc_cplx *psi_out, int32_t *izero);
/* Hermitian 18x18 eigen-decomposition used in place of LAPACK zheev (cyclic Jacobi). u: in = matrix, out = vectors */
int orc_heev18(orc_cplx *u, double *ev);
/* create_ll_map (recursion.f90:3277-3303): izeroll (0:kk, lld+1) int32 column-major, column 1 = start mask on entry */
void orc_create_ll_map(const orc_ctx *c, int lld, int32_t *izeroll);
/* chebyshev_orbital_mod, moment part (recursion.f90:2901-3008): mu_n_orb(18,18,lld) summed over the start sites;
 * cr (3,kk), alat as lattice%cr, lattice%alat */
int orc_orbital_moments(orc_ctx *c, int nstart, const int32_t *start_sites, const double *cr, double alat, int lld,
                        double a, double b, orc_cplx *mu_n_orb);
/* number of active sites after the last hop (this%irnum) */
int orc_last_irnum(const orc_ctx *c);


/* ---- consumers either side of the hot path, rsrec_oracle_post.c -------------------------------- */
/* emami / bpopt (recursion.f90:3589-3706 / 3540-3581) with 0-based arrays: as, bs (n); a, rb (ll), n = ll-1 */
void orc_emami(int n, const double *as, const double *bs, double *emax, double *emin);
void orc_bpopt(int ll, const double *a, const double *rb, double *ainf, double *rbinf, int *ifail);
/* get_terminf (recursion.f90:2092-2138): a_b, b_b (18,18,ll,na) -> a_inf, b_inf (18,18,na), a_inf0, b_inf0 (na) */
void orc_get_terminf(const orc_cplx *a_b, const orc_cplx *b_b, int na, int ll, double *a_inf, double *b_inf,
                     double *a_inf0, double *b_inf0);
/* bgreen (green.f90:1191-1339) for one unit; block_green (588-621) for na units, eta = 0, all channels */
void orc_bgreen(const orc_cplx *a_b, const orc_cplx *b_b, int ll, const double *e, int nv, int ie_start, int ie_len,
                const double *a_inf, const double *b_inf, double eta_re, double eta_im, int sym_term, orc_cplx *g_out);
void orc_block_green(const orc_cplx *a_b, const orc_cplx *b_b, int na, int ll, const double *e, int nv, int sym_term,
                     orc_cplx *g0);
void orc_jackson_kernel(int n, double *k);
void orc_lorentz_kernel(int n, double lambda, double *k);
/* chebyshev_green (green.f90:1030-1108) */
void orc_chebyshev_green(const orc_cplx *mu_n, int na, int lld, const double *ene, int nv, double energy_min,
                         double energy_max, orc_cplx *mu_ng, orc_cplx *g0);
/* bprldos / density (density_of_states.f90:378-407 / 248-372), sgreen (green.f90:628-705) */
double orc_bprldos(double e, const double *a, const double *b2, int ll, const double *ei);
void orc_density(const double *a, const double *b2, int lld, const double *ene, int nv, const double *dw_l,
                 const double *cshi, double *tdens);
void orc_sgreen(const double *a, const double *b2, int lld, int na, int nmdir, const double *ene, int nv,
                const double *dw_l, const double *cshi, orc_cplx *g0);
/* calculate_gamma_nm / integrand of calculate_conductivity_tensor (conductivity.f90:158-306) */
void orc_gamma_nm(const double *ene, int nv, int M, double energy_min, double energy_max, orc_cplx *gamma);
void orc_conductivity_integrand(const orc_cplx *mu_nm, int M, int nloop, const double *ene, int nv, double energy_min,
                                double energy_max, int per_type, orc_cplx *integrand, orc_cplx *integrand_at);

/* tail of calculate_conductivity_tensor (conductivity.f90:300-372), literal O(nv^2) simpson_f loop; sigma (2,19,nv,1+nat) */
void orc_conductivity_cumulative(const orc_cplx *integrand, const orc_cplx *integrand_at, int nv, int nv1, int nat,
                                 const double *wscale, int loop_over, double *sigma);
/* calculate_intersite_gf (green.f90:425-469): g0 (18,18,nv,4*njij) -> gij, gji (18,18,nv,njij), gspin (9,9,nv,njij,8) */
void orc_intersite_gf(const orc_cplx *g0, int nv, int njij, const int32_t *pair_i, const int32_t *pair_j, orc_cplx *gij,
                      orc_cplx *gji, orc_cplx *gspin);

/* ---- neighbour table, rsrec_oracle_lattice.c: nncal + remd (lattice.f90:3035-3123, 2823-2907) ---- */
int orc_build_nn(int kk, const double *crd, const int32_t *no, int ntot, const int32_t *iu, double ct, int use_pbc,
                 const int *b, const int *nrep, const double *a, double alat, int ncols, int32_t *nn, int *nm_out);

/* ---- `type bands` (bands.f90), rsrec_oracle_bands.c: g0 (18,18,nv,nunits) -> DOS, Fermi level, band moments ---- */
void orc_bands_dos(const orc_cplx *g0, int nv, int nunits, double *dtot, double *dosia, double *dosial);
void orc_fermi(double *ef, double h, int *ik1, double ainf, int npts, const double *y, int *ifail, double qqv, double *e1);
void orc_bands_fermi(const double *dtot, int nv, double edel, double energy_min, double qqv, int fix_fermi, double *fermi,
                     int *nv1, double *e1, int *ifail);
double orc_simpson_m(double h, double ef, int npts, const double *y, double ea, int nexp, const double *ene);
void orc_bands_magnetic_moments(const orc_cplx *g0, int nv, int nunits, const double *ene, double edel, double fermi, int nv1,
                                double e1, double *mom0, double *mom1);
void orc_bands_moments(const orc_cplx *g0, int nv, int channels_ldos, int nunits, const double *mom, const orc_cplx *lsph,
                       const double *ene, double edel, double fermi, int nv1, double e1, double *occ, double *lmom);

#ifdef __cplusplus
}
#endif
#endif

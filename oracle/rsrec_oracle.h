/*
 * rsrec_oracle.h -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * A plain-C restatement, loop for loop, of the recursion hot path of rslmtoasa/rslmtoasa
 * (`source/recursion.f90`).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library; the product (librsrec.so) never does.
 *
 * PARITY UNPINNED at the a_n/b_n/mu_n boundary: the reference stores no golden coefficient or moment
 * vectors (only end-to-end etot/DOS values that need the whole Fortran program) and cannot be compiled in
 * this image (no Fortran compiler).  The oracle is pinned instead by an independent dense numpy restatement
 * (oracle/dense_check.py) and by mathematical invariants (tests/test_oracle.py).
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
/* zsqr: recursion.f90:1980-2023, in place on b2_b(18,18,lld,na). */
int orc_zsqr(orc_cplx *b2_b, int lld, int na);
/* single operator applications (ham_vec_matmul 913-977, ham_hoh_vec_matmul 785-911, velo_* 587-783);
 * izero (kk+1 ints, index 0 unused/0) is updated to idum on return like the callers do. */
void orc_ham_vec_matmul(orc_ctx *c, const orc_cplx *psi_in, orc_cplx *psi_out, double a, double b, int32_t *izero);
void orc_velo_vec_matmul(orc_ctx *c, int slot, const orc_cplx *psi_in, orc_cplx *psi_out, int32_t *izero);
/* Hermitian 18x18 eigen-decomposition used in place of LAPACK zheev (cyclic Jacobi). u: in = matrix, out = vectors */
int orc_heev18(orc_cplx *u, double *ev);
/* number of active sites after the last hop (this%irnum) */
int orc_last_irnum(const orc_ctx *c);

#ifdef __cplusplus
}
#endif
#endif

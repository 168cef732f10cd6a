/*
 * rsrec_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).  See rsrec_oracle.h.
 *
 * PARITY PINNED by the reference's bccFe golden fixtures (recur_b, hop_b(_hoh), chebyshev_recur, zsqr; see
 * rsrec_oracle.h) and, for the routines those do not reach, by oracle/dense_check.py.
 *
 * Restates, with the reference's own pass structure, masks and per-site 18x18x18 complex products:
 *   hop_b            recursion.f90:1560-1648        hop_b_hoh          recursion.f90:1411-1552
 *   crecal_b         recursion.f90:1873-1973        recur_b / _ij      recursion.f90:1807-1866 / 1655-1737
 *   zsqr             recursion.f90:1980-2023        hop/crecal/recur   recursion.f90:3310-3532
 *   cheb_0th_mom     recursion.f90:2145-2162        cheb_1st_mom(_hoh) recursion.f90:2169-2369
 *   chebyshev_recur_ll(_hoh) recursion.f90:2495-2763   chebyshev_recur(_ij) recursion.f90:3057-3130 / 2376-2487
 *   ham(_hoh)_vec_matmul recursion.f90:913-977 / 785-911   velo(_hoh)_vec_matmul recursion.f90:587-783
 *   compute_moments_stochastic recursion.f90:979-1234
 * BLAS/LAPACK calls of the reference (zgemm 18^3, zaxpy, zheev) are restated with plain loops / cyclic Jacobi.
 *
 * Build: see oracle/Makefile (gcc -O3 -march=native -fopenmp -shared -fPIC).
 */
#include "rsrec_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define NB 18
#define BLK (NB * NB)
typedef orc_cplx cplx;

struct orc_ctx {
  int kk, ncols, nslot, ntype, nmax, hoh, use_mask;
  const int32_t *nn, *iz;
  const cplx *ee, *eeo, *hall, *hallo, *lsham, *enim;
  const cplx *v_a, *v_b, *vo_a, *vo_b;
  /* recursion members (recursion.f90:52-76) */
  cplx *psi_b, *pmn_b, *hpsi, *hohpsi, *enupsi, *socpsi, *psi0, *psi1, *psi2;
  cplx *psi, *pmn, *v;
  int32_t *izero, *idum, *irlist; /* (0:kk) */
  int irnum;
  cplx cheb_mom_temp[BLK];
};

/* ---- index helpers (Fortran 1-based, column-major) ---- */
#define NN(c, i, j) ((c)->nn[((size_t)(i) - 1) + (size_t)(c)->kk * ((j) - 1)])
#define HBLK(arr, c, j, t) ((arr) + (size_t)BLK * (((size_t)(j) - 1) + (size_t)(c)->nslot * ((size_t)(t) - 1)))
#define TBLK(arr, t) ((arr) + (size_t)BLK * ((size_t)(t) - 1))
#define SBLK(arr, i) ((arr) + (size_t)BLK * ((size_t)(i) - 1))

/* ---- 18x18x18 complex products: the reference's zgemm calls ---- */
/* C += alpha * A * B   (zgemm 'n','n') */
static inline void gemm_nn(cplx *restrict C, const cplx *restrict A, const cplx *restrict B, double alpha) {
  for (int j = 0; j < NB; j++) {
    double cr[NB], ci[NB];
    for (int i = 0; i < NB; i++) { cr[i] = 0.0; ci[i] = 0.0; }
    for (int k = 0; k < NB; k++) {
      const double br = creal(B[k + NB * j]), bi = cimag(B[k + NB * j]);
      const cplx *a = A + NB * k;
      for (int i = 0; i < NB; i++) {
        const double ar = creal(a[i]), ai = cimag(a[i]);
        cr[i] += ar * br - ai * bi;
        ci[i] += ar * bi + ai * br;
      }
    }
    for (int i = 0; i < NB; i++) C[i + NB * j] += alpha * (cr[i] + I * ci[i]);
  }
}
/* C = A * B   (zgemm 'n','n', beta = 0) */
static inline void gemm_nn_set(cplx *restrict C, const cplx *restrict A, const cplx *restrict B) {
  for (int i = 0; i < BLK; i++) C[i] = 0.0;
  gemm_nn(C, A, B, 1.0);
}
/* C += A^H * B   (zgemm 'c','n') */
static inline void gemm_cn(cplx *restrict C, const cplx *restrict A, const cplx *restrict B) {
  for (int j = 0; j < NB; j++)
    for (int i = 0; i < NB; i++) {
      double sr = 0.0, si = 0.0;
      const cplx *a = A + NB * i, *b = B + NB * j;
      for (int k = 0; k < NB; k++) {
        const double ar = creal(a[k]), ai = -cimag(a[k]);
        const double br = creal(b[k]), bi = cimag(b[k]);
        sr += ar * br - ai * bi;
        si += ar * bi + ai * br;
      }
      C[i + NB * j] += sr + I * si;
    }
}
/* C = A * B^H  (zgemm 'n','c', beta = 0) */
static inline void gemm_nc_set(cplx *restrict C, const cplx *restrict A, const cplx *restrict B) {
  for (int j = 0; j < NB; j++)
    for (int i = 0; i < NB; i++) {
      cplx s = 0.0;
      for (int k = 0; k < NB; k++) s += A[i + NB * k] * conj(B[j + NB * k]);
      C[i + NB * j] = s;
    }
}

/* ---- Hermitian eigen-solver (stands in for LAPACK zheev('v','u')) ---- */
int orc_heev18(cplx *a, double *ev) {
  cplx v[BLK];
  for (int i = 0; i < BLK; i++) v[i] = 0.0;
  for (int i = 0; i < NB; i++) v[i + NB * i] = 1.0;
  /* zheev('u') reads the upper triangle only: symmetrise from it */
  for (int j = 0; j < NB; j++) {
    a[j + NB * j] = creal(a[j + NB * j]);
    for (int i = j + 1; i < NB; i++) a[i + NB * j] = conj(a[j + NB * i]);
  }
  for (int sweep = 0; sweep < 100; sweep++) {
    double off = 0.0, tot = 0.0;
    for (int j = 0; j < NB; j++)
      for (int i = 0; i < NB; i++) {
        double m = creal(a[i + NB * j]) * creal(a[i + NB * j]) + cimag(a[i + NB * j]) * cimag(a[i + NB * j]);
        tot += m;
        if (i != j) off += m;
      }
    if (off <= 1e-62 * tot || off == 0.0) break;
    for (int p = 0; p < NB - 1; p++)
      for (int q = p + 1; q < NB; q++) {
        cplx apq = a[p + NB * q];
        double mag = cabs(apq);
        if (mag == 0.0) continue;
        double app = creal(a[p + NB * p]), aqq = creal(a[q + NB * q]);
        if (mag < 1e-300) continue;
        cplx ph = apq / mag; /* e^{i phi} */
        double tau = (aqq - app) / (2.0 * mag);
        double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
        double cs = 1.0 / sqrt(1.0 + t * t), sn = t * cs;
        /* R = [[cs, sn*ph],[ -sn*conj(ph), cs ]] acting on columns (p,q):  A <- R^H A R, V <- V R */
        cplx rpq = sn * ph, rqp = -sn * conj(ph);
        for (int k = 0; k < NB; k++) { /* columns */
          cplx akp = a[k + NB * p], akq = a[k + NB * q];
          a[k + NB * p] = akp * cs + akq * rqp;
          a[k + NB * q] = akp * rpq + akq * cs;
          cplx vkp = v[k + NB * p], vkq = v[k + NB * q];
          v[k + NB * p] = vkp * cs + vkq * rqp;
          v[k + NB * q] = vkp * rpq + vkq * cs;
        }
        for (int k = 0; k < NB; k++) { /* rows: R^H */
          cplx apk = a[p + NB * k], aqk = a[q + NB * k];
          a[p + NB * k] = cs * apk + conj(rqp) * aqk;
          a[q + NB * k] = conj(rpq) * apk + cs * aqk;
        }
        a[p + NB * q] = 0.0;
        a[q + NB * p] = 0.0;
        a[p + NB * p] = creal(a[p + NB * p]);
        a[q + NB * q] = creal(a[q + NB * q]);
      }
  }
  for (int i = 0; i < NB; i++) ev[i] = creal(a[i + NB * i]);
  memcpy(a, v, sizeof(v));
  return 0;
}

/* B = U sqrt(ev) U^H and (optionally) B^-1 = U ev^-1/2 U^H  (recursion.f90:1947-1959, 2015-2020) */
static void sqrt_from_eig(const cplx *u, const double *ev, cplx *b, cplx *b_i) {
  cplx lam[BLK], lam_i[BLK], dum[BLK];
  for (int i = 0; i < BLK; i++) { lam[i] = 0.0; lam_i[i] = 0.0; }
  for (int i = 0; i < NB; i++) {
    lam[i + NB * i] = sqrt(ev[i]); /* NaN for ev<0, like the reference */
    lam_i[i + NB * i] = 1.0 / lam[i + NB * i];
  }
  gemm_nn_set(dum, u, lam);
  gemm_nc_set(b, dum, u);
  if (b_i) {
    gemm_nn_set(dum, u, lam_i);
    gemm_nc_set(b_i, dum, u);
  }
}

/* ---- context ---- */
orc_ctx *orc_create(int kk, int ncols, int nslot, int ntype, int nmax, const int32_t *nn, const int32_t *iz) {
  orc_ctx *c = (orc_ctx *)calloc(1, sizeof(orc_ctx));
  c->kk = kk; c->ncols = ncols; c->nslot = nslot; c->ntype = ntype; c->nmax = nmax;
  c->nn = nn; c->iz = iz; c->use_mask = 1;
  size_t n = (size_t)BLK * kk;
  cplx **blocks[] = {&c->psi_b, &c->pmn_b, &c->hpsi, &c->hohpsi, &c->enupsi, &c->socpsi, &c->psi0, &c->psi1, &c->psi2};
  for (unsigned i = 0; i < sizeof(blocks) / sizeof(blocks[0]); i++) *blocks[i] = (cplx *)calloc(n, sizeof(cplx));
  c->psi = (cplx *)calloc((size_t)NB * kk, sizeof(cplx));
  c->pmn = (cplx *)calloc((size_t)NB * kk, sizeof(cplx));
  c->v = (cplx *)calloc((size_t)NB * kk, sizeof(cplx));
  c->izero = (int32_t *)calloc(kk + 1, sizeof(int32_t));
  c->idum = (int32_t *)calloc(kk + 1, sizeof(int32_t));
  c->irlist = (int32_t *)calloc(kk + 1, sizeof(int32_t));
  return c;
}
void orc_destroy(orc_ctx *c) {
  if (!c) return;
  free(c->psi_b); free(c->pmn_b); free(c->hpsi); free(c->hohpsi); free(c->enupsi); free(c->socpsi);
  free(c->psi0); free(c->psi1); free(c->psi2); free(c->psi); free(c->pmn); free(c->v);
  free(c->izero); free(c->idum); free(c->irlist); free(c);
}
void orc_set_hamiltonian(orc_ctx *c, const cplx *ee, const cplx *eeo, const cplx *hall, const cplx *hallo,
                         const cplx *lsham, const cplx *enim, int hoh) {
  c->ee = ee; c->eeo = eeo; c->hall = hall; c->hallo = hallo; c->lsham = lsham; c->enim = enim; c->hoh = hoh;
}
void orc_set_operator(orc_ctx *c, int slot, const cplx *v_op, const cplx *vo_op) {
  if (slot == 'a') { c->v_a = v_op; c->vo_a = vo_op; } else { c->v_b = v_op; c->vo_b = vo_op; }
}
void orc_set_use_mask(orc_ctx *c, int use_mask) { c->use_mask = use_mask; }
void orc_set_threads(int n) {
#ifdef _OPENMP
  omp_set_num_threads(n);
#else
  (void)n;
#endif
}
int orc_get_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
int orc_last_irnum(const orc_ctx *c) { return c->irnum; }

static void clear_mask(orc_ctx *c) {
  for (int i = 0; i <= c->kk; i++) c->izero[i] = c->use_mask ? 0 : (i > 0);
}
static void zero_blocks(cplx *p, int kk) { memset(p, 0, sizeof(cplx) * (size_t)BLK * kk); }

/* The gather-SpMV shared by every routine:  out_i += Hsite(:,:,1,i)*in_i [+ extra on-site terms handled by caller]
 * + sum_{j=2..nr} H(:,:,j,.) * in_{nn(i,j)},  H = hsite(:,:,j,i) for i<=nmax else htype(:,:,j,iz(i));
 * neighbour term only if nn/=0 and izero(nn)/=0; idum(i) = izero(i) or any active neighbour.
 * `onsite`: 0 = skip the on-site product (caller does it), 1 = H(:,:,1,.) , 2 = H(:,:,1,.) + lsham (locham). */
static void spmv_site(const orc_ctx *c, int i, const cplx *htype, const cplx *hsite, const cplx *in, cplx *out,
                      int onsite, const int32_t *izero, int32_t *idum) {
  const int t = c->iz[i - 1];
  const int local = (i <= c->nmax);
  const int nr = NN(c, i, 1);
  idum[i] = izero[i];
  if (izero[i] != 0 && onsite) {
    const cplx *h1 = local ? HBLK(hsite, c, 1, i) : HBLK(htype, c, 1, t);
    if (onsite == 2) {
      cplx locham[BLK];
      const cplx *ls = TBLK(c->lsham, t);
      for (int k = 0; k < BLK; k++) locham[k] = h1[k] + ls[k];
      gemm_nn(SBLK(out, i), locham, SBLK(in, i), 1.0);
    } else {
      gemm_nn(SBLK(out, i), h1, SBLK(in, i), 1.0);
    }
  }
  for (int j = 2; j <= nr; j++) {
    const int nb = NN(c, i, j);
    if (nb != 0 && izero[nb] != 0) {
      const cplx *h = local ? HBLK(hsite, c, j, i) : HBLK(htype, c, j, t);
      gemm_nn(SBLK(out, i), h, SBLK(in, nb), 1.0);
      idum[i] = 1;
    }
  }
}

/* rebuild izero / irlist / irnum from idum (recursion.f90:1628-1636, 2569-2577) */
static void rebuild_active(orc_ctx *c, const int32_t *idum) {
  c->irnum = 0;
  for (int i = 1; i <= c->kk; i++) {
    c->izero[i] = idum[i];
    if (c->izero[i] != 0) c->irlist[++c->irnum] = i;
  }
}

/* ===================== block Lanczos ===================== */
/* hop_b: recursion.f90:1560-1648 */
static void hop_b(orc_ctx *c, cplx *atemp_ll) {
  const int kk = c->kk;
  int32_t *idum = (int32_t *)calloc(kk + 1, sizeof(int32_t)); /* local idum, recursion.f90:1567 */
  zero_blocks(c->hpsi, kk);
#pragma omp parallel for schedule(dynamic, 100)
  for (int i = 1; i <= kk; i++) spmv_site(c, i, c->ee, c->hall, c->psi_b, c->hpsi, 2, c->izero, idum);
  rebuild_active(c, idum);
  double sr[BLK], si[BLK];
  for (int k = 0; k < BLK; k++) { sr[k] = 0.0; si[k] = 0.0; }
#pragma omp parallel for schedule(dynamic, 100) reduction(+ : sr[:BLK], si[:BLK])
  for (int n = 1; n <= c->irnum; n++) {
    const int i = c->irlist[n];
    cplx *pm = SBLK(c->pmn_b, i);
    const cplx *hp = SBLK(c->hpsi, i);
    for (int k = 0; k < BLK; k++) pm[k] = hp[k] - pm[k];
    cplx s[BLK];
    for (int k = 0; k < BLK; k++) s[k] = 0.0;
    gemm_cn(s, SBLK(c->psi_b, i), hp);
    for (int k = 0; k < BLK; k++) { sr[k] += creal(s[k]); si[k] += cimag(s[k]); }
  }
  for (int k = 0; k < BLK; k++) atemp_ll[k] = sr[k] + I * si[k];
  free(idum);
}

/* hop_b_hoh: recursion.f90:1411-1552 */
static void hop_b_hoh(orc_ctx *c, cplx *atemp_ll) {
  const int kk = c->kk;
  int32_t *idum = (int32_t *)calloc(kk + 1, sizeof(int32_t));
  zero_blocks(c->hpsi, kk); zero_blocks(c->hohpsi, kk); zero_blocks(c->enupsi, kk); zero_blocks(c->socpsi, kk);
#pragma omp parallel for schedule(dynamic, 100)
  for (int i = 1; i <= kk; i++) {
    if (c->izero[i] != 0) {
      const int t = c->iz[i - 1];
      gemm_nn(SBLK(c->enupsi, i), TBLK(c->enim, t), SBLK(c->psi_b, i), 1.0);
      gemm_nn(SBLK(c->socpsi, i), TBLK(c->lsham, t), SBLK(c->psi_b, i), 1.0);
    }
    spmv_site(c, i, c->ee, c->hall, c->psi_b, c->hpsi, 1, c->izero, idum);
  }
  memcpy(c->izero, idum, sizeof(int32_t) * (kk + 1)); /* mapping update, 1480 */
#pragma omp parallel for schedule(dynamic, 100)
  for (int i = 1; i <= kk; i++) spmv_site(c, i, c->eeo, c->hallo, c->hpsi, c->hohpsi, 1, c->izero, idum);
  rebuild_active(c, idum);
  double sr[BLK], si[BLK];
  for (int k = 0; k < BLK; k++) { sr[k] = 0.0; si[k] = 0.0; }
#pragma omp parallel for schedule(dynamic, 100) reduction(+ : sr[:BLK], si[:BLK])
  for (int n = 1; n <= c->irnum; n++) {
    const int i = c->irlist[n];
    cplx *hp = SBLK(c->hpsi, i), *pm = SBLK(c->pmn_b, i);
    const cplx *ho = SBLK(c->hohpsi, i), *en = SBLK(c->enupsi, i), *so = SBLK(c->socpsi, i);
    for (int k = 0; k < BLK; k++) hp[k] = hp[k] - ho[k] + en[k] + so[k]; /* H = h - hoh + e_nu + l.s */
    for (int k = 0; k < BLK; k++) pm[k] = hp[k] - pm[k];
    cplx s[BLK];
    for (int k = 0; k < BLK; k++) s[k] = 0.0;
    gemm_cn(s, SBLK(c->psi_b, i), hp);
    for (int k = 0; k < BLK; k++) { sr[k] += creal(s[k]); si[k] += cimag(s[k]); }
  }
  for (int k = 0; k < BLK; k++) atemp_ll[k] = sr[k] + I * si[k];
  free(idum);
}

/* crecal_b: recursion.f90:1873-1973.  atemp_b, b2temp_b: 18x18xlld; b2temp_b(:,:,1) preset by the caller. */
static int crecal_b(orc_ctx *c, int lld, cplx *atemp_b, cplx *b2temp_b) {
  const int kk = c->kk;
  cplx *psi_t = (cplx *)malloc(sizeof(cplx) * (size_t)BLK * kk);
  cplx sum_b[BLK];
  memcpy(sum_b, b2temp_b, sizeof(sum_b));
  for (int ll = 1; ll <= lld - 1; ll++) {
    cplx *a_ll = atemp_b + (size_t)BLK * (ll - 1);
    for (int k = 0; k < BLK; k++) a_ll[k] = 0.0;
    if (c->hoh) hop_b_hoh(c, a_ll); else hop_b(c, a_ll);
    for (int n = 1; n <= c->irnum; n++)
      memcpy(SBLK(psi_t, c->irlist[n]), SBLK(c->psi_b, c->irlist[n]), sizeof(cplx) * BLK);
    memcpy(b2temp_b + (size_t)BLK * (ll - 1), sum_b, sizeof(sum_b));
    double sr[BLK], si[BLK];
    for (int k = 0; k < BLK; k++) { sr[k] = 0.0; si[k] = 0.0; }
#pragma omp parallel for reduction(+ : sr[:BLK], si[:BLK])
    for (int n = 1; n <= c->irnum; n++) {
      const int i = c->irlist[n];
      cplx *pm = SBLK(c->pmn_b, i);
      gemm_nn(pm, SBLK(c->psi_b, i), a_ll, -1.0); /* pmn -= psi * A */
      cplx s[BLK];
      for (int k = 0; k < BLK; k++) s[k] = 0.0;
      gemm_cn(s, pm, pm);                          /* B^2 += pmn^H pmn */
      for (int k = 0; k < BLK; k++) { sr[k] += creal(s[k]); si[k] += cimag(s[k]); }
    }
    for (int k = 0; k < BLK; k++) sum_b[k] = sr[k] + I * si[k];
    cplx u[BLK], b[BLK], b_i[BLK];
    double ev[NB];
    memcpy(u, sum_b, sizeof(u));
    orc_heev18(u, ev);
    sqrt_from_eig(u, ev, b, b_i);
#pragma omp parallel for
    for (int n = 1; n <= c->irnum; n++) {
      const int i = c->irlist[n];
      gemm_nn_set(SBLK(c->psi_b, i), SBLK(c->pmn_b, i), b_i); /* psi = pmn * B^-1 */
      gemm_nn_set(SBLK(c->pmn_b, i), SBLK(psi_t, i), b);      /* pmn = psi_old * B */
    }
  }
  memcpy(b2temp_b + (size_t)BLK * (lld - 1), sum_b, sizeof(sum_b));
  free(psi_t);
  return 0;
}

int orc_lanczos_block(orc_ctx *c, int nunits, const int32_t *site_i, const int32_t *site_j, const cplx *asign,
                      const cplx *bsign, int lld, cplx *a_b, cplx *b2_b) {
  const int kk = c->kk;
  cplx *atemp_b = (cplx *)calloc((size_t)BLK * lld, sizeof(cplx));
  cplx *b2temp_b = (cplx *)calloc((size_t)BLK * lld, sizeof(cplx));
  for (int u = 0; u < nunits; u++) {
    const int i = site_i[u], j = site_j ? site_j[u] : 0;
    clear_mask(c);
    zero_blocks(c->psi_b, kk); zero_blocks(c->pmn_b, kk);
    memset(atemp_b, 0, sizeof(cplx) * (size_t)BLK * lld);
    memset(b2temp_b, 0, sizeof(cplx) * (size_t)BLK * lld);
    const cplx as = asign ? asign[u] : 1.0, bs = bsign ? bsign[u] : 1.0;
    for (int l = 0; l < NB; l++) {
      SBLK(c->psi_b, i)[l + NB * l] = as;
      if (j > 0) SBLK(c->psi_b, j)[l + NB * l] = bs; /* i==j: bsign overwrites (recursion.f90:1710-1711) */
      b2temp_b[l + NB * l] = 1.0;
    }
    c->izero[i] = 1;
    if (j > 0) c->izero[j] = 1;
    crecal_b(c, lld, atemp_b, b2temp_b);
    memcpy(a_b + (size_t)BLK * lld * u, atemp_b, sizeof(cplx) * (size_t)BLK * lld);
    memcpy(b2_b + (size_t)BLK * lld * u, b2temp_b, sizeof(cplx) * (size_t)BLK * lld);
  }
  free(atemp_b); free(b2temp_b);
  return 0;
}

int orc_zsqr(cplx *b2_b, int lld, int na) {
  for (int n = 0; n < na; n++)
    for (int l = 0; l < lld; l++) {
      cplx *blk = b2_b + (size_t)BLK * (l + (size_t)lld * n);
      cplx u[BLK], b[BLK];
      double ev[NB];
      memcpy(u, blk, sizeof(u));
      orc_heev18(u, ev);
      sqrt_from_eig(u, ev, b, NULL);
      memcpy(blk, b, sizeof(b));
    }
  return 0;
}

/* ===================== scalar Lanczos (nsp = 1) ===================== */
/* hop: recursion.f90:3310-3416 -- two independent 9x9 spin blocks per neighbour, no lsham */
static double hop_scalar(orc_ctx *c) {
  const int kk = c->kk;
  int32_t *idum = (int32_t *)calloc(kk + 1, sizeof(int32_t));
#pragma omp parallel for schedule(dynamic, 100)
  for (int i = 1; i <= kk; i++) {
    cplx dum[NB];
    for (int l = 0; l < NB; l++) dum[l] = 0.0;
    const int t = c->iz[i - 1], local = (i <= c->nmax), nr = NN(c, i, 1);
    idum[i] = c->izero[i];
    for (int j = 1; j <= nr; j++) {
      const int nb = (j == 1) ? i : NN(c, i, j);
      if (nb == 0 || c->izero[nb] == 0) continue;
      const cplx *h = local ? HBLK(c->hall, c, j, i) : HBLK(c->ee, c, j, t);
      const cplx *p = c->psi + (size_t)NB * (nb - 1);
      for (int m = 0; m < 9; m++)
        for (int l = 0; l < 9; l++) {
          dum[l] += h[l + NB * m] * p[m];
          dum[l + 9] += h[(l + 9) + NB * (m + 9)] * p[m + 9];
        }
      if (j > 1) idum[i] = 1;
    }
    for (int l = 0; l < NB; l++) c->v[(size_t)NB * (i - 1) + l] = dum[l];
  }
  double summ = 0.0;
  for (int i = 1; i <= kk; i++) {
    c->izero[i] = idum[i];
    for (int l = 0; l < NB; l++) {
      const size_t o = (size_t)NB * (i - 1) + l;
      summ += creal(c->v[o] * conj(c->psi[o]));
      c->pmn[o] = c->v[o] + c->pmn[o];
    }
  }
  free(idum);
  return summ;
}

/* crecal + recur: recursion.f90:3423-3478, 3485-3532 */
int orc_lanczos_scalar(orc_ctx *c, int nunits, const int32_t *sites, int lld, double *a, double *b2) {
  const size_t n = (size_t)NB * c->kk;
  for (int u = 0; u < nunits; u++)
    for (int l = 0; l < NB; l++) {
      double *atemp = a + (size_t)lld * (l + (size_t)NB * u), *b2temp = b2 + (size_t)lld * (l + (size_t)NB * u);
      clear_mask(c);
      memset(c->psi, 0, sizeof(cplx) * n); memset(c->pmn, 0, sizeof(cplx) * n);
      c->psi[(size_t)NB * (sites[u] - 1) + l] = 1.0;
      c->izero[sites[u]] = 1;
      for (int k = 0; k < lld; k++) { atemp[k] = 0.0; b2temp[k] = 0.0; }
      b2temp[0] = 1.0;
      double summ = b2temp[0];
      for (int ll = 1; ll <= lld - 1; ll++) {
        atemp[ll - 1] = hop_scalar(c);
        b2temp[ll - 1] = summ;
        const double ajc = -atemp[ll - 1];
        for (size_t k = 0; k < n; k++) c->pmn[k] += ajc * c->psi[k]; /* zaxpy */
        summ = 0.0;
        for (size_t k = 0; k < n; k++) summ += creal(conj(c->pmn[k]) * c->pmn[k]);
        double s = 1.0 / sqrt(summ);
        const double s2 = sqrt(summ);
        for (size_t k = 0; k < n; k++) {
          const cplx th = c->pmn[k] * s;
          c->pmn[k] = -c->psi[k] * s2;
          c->psi[k] = th;
        }
      }
      b2temp[lld - 1] = summ;
    }
  return 0;
}

/* ===================== Chebyshev ===================== */
/* cheb_0th_mom: recursion.f90:2145-2162 */
static void cheb_0th_mom(orc_ctx *c, const cplx *psiref) {
  for (int k = 0; k < BLK; k++) c->cheb_mom_temp[k] = 0.0;
  for (int k = 1; k <= c->kk; k++)
    if (c->izero[k] != 0) gemm_cn(c->cheb_mom_temp, SBLK(psiref, k), SBLK(c->psi0, k));
}

/* H|in> into out for one site, no-hoh flavour of the Chebyshev routines: on-site = ee|hall (+) lsham.
 * cheb_1st_mom applies ee and lsham as two separate products (2189-2190), chebyshev_recur_ll sums them first
 * (2515, 2545): `two_gemm` selects which. */
static void cheb_spmv_site(const orc_ctx *c, int i, const cplx *in, cplx *out, int two_gemm, const int32_t *izero,
                           int32_t *idum) {
  if (two_gemm) {
    if (izero[i] != 0) { /* reference order: H_onsite, lsham, then the neighbours */
      const int t = c->iz[i - 1];
      const cplx *h1 = (i <= c->nmax) ? HBLK(c->hall, c, 1, i) : HBLK(c->ee, c, 1, t);
      gemm_nn(SBLK(out, i), h1, SBLK(in, i), 1.0);
      gemm_nn(SBLK(out, i), TBLK(c->lsham, t), SBLK(in, i), 1.0);
    }
    spmv_site(c, i, c->ee, c->hall, in, out, 0, izero, idum);
  } else {
    spmv_site(c, i, c->ee, c->hall, in, out, 2, izero, idum);
  }
}

/* hoh flavour, whole cluster: out = h in - (h o)(h in) + e_nu in + l.s in  (recursion.f90:2245-2357, 2605-2723, 785-904).
 * `out` must be zero on entry.  Updates izero after the first sweep (2312 / 2674 / 855) and leaves the final
 * reachability in idum. */
static void hoh_apply(orc_ctx *c, const cplx *in, cplx *out) {
  const int kk = c->kk;
  zero_blocks(c->hohpsi, kk); zero_blocks(c->enupsi, kk); zero_blocks(c->socpsi, kk);
#pragma omp parallel for
  for (int i = 1; i <= kk; i++) {
    if (c->izero[i] != 0) {
      const int t = c->iz[i - 1];
      gemm_nn(SBLK(c->enupsi, i), TBLK(c->enim, t), SBLK(in, i), 1.0);
      gemm_nn(SBLK(c->socpsi, i), TBLK(c->lsham, t), SBLK(in, i), 1.0);
    }
    spmv_site(c, i, c->ee, c->hall, in, out, 1, c->izero, c->idum);
  }
  memcpy(c->izero, c->idum, sizeof(int32_t) * (kk + 1));
#pragma omp parallel for
  for (int i = 1; i <= kk; i++) spmv_site(c, i, c->eeo, c->hallo, out, c->hohpsi, 1, c->izero, c->idum);
  const size_t n = (size_t)BLK * kk;
#pragma omp parallel for
  for (size_t k = 0; k < n; k++) out[k] = out[k] - c->hohpsi[k] + c->enupsi[k] + c->socpsi[k];
}

/* cheb_1st_mom / cheb_1st_mom_hoh: recursion.f90:2169-2238 / 2245-2369 */
static void cheb_1st_mom(orc_ctx *c, const cplx *psiref, double a, double b) {
  const int kk = c->kk;
  for (int k = 0; k < BLK; k++) c->cheb_mom_temp[k] = 0.0;
  if (c->hoh) {
    hoh_apply(c, c->psi0, c->psi1);
    const size_t n = (size_t)BLK * kk;
    for (size_t k = 0; k < n; k++) { c->psi1[k] = c->psi1[k] - b * c->psi0[k]; c->psi1[k] = c->psi1[k] / a; }
  } else {
    for (int i = 1; i <= kk; i++) {
      cheb_spmv_site(c, i, c->psi0, c->psi1, 1, c->izero, c->idum);
      cplx *p1 = SBLK(c->psi1, i);
      const cplx *p0 = SBLK(c->psi0, i);
      for (int k = 0; k < BLK; k++) { p1[k] = p1[k] - b * p0[k]; p1[k] = p1[k] / a; }
    }
  }
  for (int n = 1; n <= kk; n++)
    if (c->izero[n] != 0) gemm_cn(c->cheb_mom_temp, SBLK(psiref, n), SBLK(c->psi1, n));
}

/* chebyshev_recur_ll / _hoh: recursion.f90:2495-2597 / 2605-2763.  mu: 18x18x(2lld+2) of this unit. */
static int chebyshev_recur_ll(orc_ctx *c, int ll, double a, double b, cplx *mu) {
  const int kk = c->kk;
  if (c->hoh) {
    hoh_apply(c, c->psi1, c->psi2);
    const size_t n = (size_t)BLK * kk;
#pragma omp parallel for
    for (size_t k = 0; k < n; k++) {
      c->psi2[k] = c->psi2[k] - b * c->psi1[k];
      c->psi2[k] = c->psi2[k] / a;
      c->psi2[k] = 2 * c->psi2[k];
    }
  } else {
#pragma omp parallel for
    for (int i = 1; i <= kk; i++) {
      cheb_spmv_site(c, i, c->psi1, c->psi2, 0, c->izero, c->idum);
      cplx *p2 = SBLK(c->psi2, i);
      const cplx *p1 = SBLK(c->psi1, i);
      for (int k = 0; k < BLK; k++) { p2[k] = p2[k] - b * p1[k]; p2[k] = p2[k] / a; p2[k] = 2 * p2[k]; }
    }
  }
  rebuild_active(c, c->idum);
  double d1r[BLK], d1i[BLK], d2r[BLK], d2i[BLK];
  for (int k = 0; k < BLK; k++) { d1r[k] = d1i[k] = d2r[k] = d2i[k] = 0.0; }
#pragma omp parallel for reduction(+ : d1r[:BLK], d1i[:BLK], d2r[:BLK], d2i[:BLK])
  for (int n = 1; n <= c->irnum; n++) {
    const int i = c->irlist[n];
    cplx *p0 = SBLK(c->psi0, i), *p1 = SBLK(c->psi1, i), *p2 = SBLK(c->psi2, i);
    for (int k = 0; k < BLK; k++) p2[k] = p2[k] - p0[k];
    cplx s1[BLK], s2[BLK];
    for (int k = 0; k < BLK; k++) { s1[k] = 0.0; s2[k] = 0.0; }
    gemm_cn(s1, p1, p1);
    gemm_cn(s2, p2, p1);
    for (int k = 0; k < BLK; k++) {
      d1r[k] += creal(s1[k]); d1i[k] += cimag(s1[k]); d2r[k] += creal(s2[k]); d2i[k] += cimag(s2[k]);
      p0[k] = p1[k]; p1[k] = p2[k]; p2[k] = 0.0;
    }
  }
  cplx *m1 = mu + (size_t)BLK * (2 * ll), *m2 = mu + (size_t)BLK * (2 * ll + 1); /* mu_n(:,:,2ll+1), (:,:,2ll+2) */
  double guard = 0.0;
  for (int k = 0; k < BLK; k++) {
    m1[k] = 2.0 * (d1r[k] + I * d1i[k]) - mu[k];
    m2[k] = 2.0 * (d2r[k] + I * d2i[k]) - mu[BLK + k];
    guard += creal(m2[k]);
  }
  return guard > 1000.0 ? -2 : 0;
}

static int cheb_unit(orc_ctx *c, const cplx *psiref, int lld, double a, double b, cplx *mu) {
  int rc = 0;
  cheb_0th_mom(c, psiref);
  memcpy(mu, c->cheb_mom_temp, sizeof(cplx) * BLK);
  cheb_1st_mom(c, psiref, a, b);
  memcpy(mu + BLK, c->cheb_mom_temp, sizeof(cplx) * BLK);
  memcpy(c->izero, c->idum, sizeof(int32_t) * (c->kk + 1)); /* 3118 / 2469 */
  for (int ll = 1; ll <= lld; ll++)
    if (chebyshev_recur_ll(c, ll, a, b, mu) != 0) rc = -2;
  return rc;
}

int orc_cheb_moments(orc_ctx *c, int nunits, const int32_t *site_i, const int32_t *site_j, const cplx *asign,
                     const cplx *bsign, int lld, double a, double b, cplx *mu_n) {
  const int kk = c->kk;
  int rc = 0;
  cplx *psiref = (cplx *)malloc(sizeof(cplx) * (size_t)BLK * kk);
  for (int u = 0; u < nunits; u++) {
    const int i = site_i[u], j = site_j ? site_j[u] : 0;
    clear_mask(c);
    zero_blocks(c->psi0, kk); zero_blocks(c->psi1, kk); zero_blocks(c->psi2, kk); zero_blocks(psiref, kk);
    const cplx as = asign ? asign[u] : 1.0, bs = bsign ? bsign[u] : 1.0;
    for (int l = 0; l < NB; l++) {
      SBLK(c->psi0, i)[l + NB * l] = as;
      if (j > 0) SBLK(c->psi0, j)[l + NB * l] = bs;
    }
    memcpy(psiref, c->psi0, sizeof(cplx) * (size_t)BLK * kk);
    c->izero[i] = 1;
    if (j > 0) c->izero[j] = 1;
    if (cheb_unit(c, psiref, lld, a, b, mu_n + (size_t)BLK * (2 * lld + 2) * u) != 0) rc = -2;
  }
  free(psiref);
  return rc;
}

static void random_start(const orc_ctx *c, const double *u, cplx *psiref) {
  const int kk = c->kk;
  /* `sqrt(real(this%lattice%kk))` is evaluated in DEFAULT (single) real kind in the reference (no
   * -fdefault-real-8 in its build flags), then promoted: keep that, it matters at the 1e-8 level. */
  const double nrm = (double)sqrtf((float)kk);
  zero_blocks(psiref, kk);
  for (int k = 1; k <= kk; k++) {
    const cplx ph = cexp(2.0 * M_PI * I * u[k - 1]);
    for (int m = 0; m < NB; m++) SBLK(psiref, k)[m + NB * m] = ph / nrm; /* recursion.f90:1135-1142 */
  }
}

int orc_cheb_moments_random(orc_ctx *c, int nvec, const double *phases, int lld, double a, double b, cplx *mu_n) {
  const int kk = c->kk;
  int rc = 0;
  cplx *psiref = (cplx *)malloc(sizeof(cplx) * (size_t)BLK * kk);
  for (int v = 0; v < nvec; v++) {
    for (int i = 0; i <= kk; i++) c->izero[i] = (i > 0);
    zero_blocks(c->psi1, kk); zero_blocks(c->psi2, kk);
    random_start(c, phases + (size_t)kk * v, psiref);
    memcpy(c->psi0, psiref, sizeof(cplx) * (size_t)BLK * kk);
    if (cheb_unit(c, psiref, lld, a, b, mu_n + (size_t)BLK * (2 * lld + 2) * v) != 0) rc = -2;
  }
  free(psiref);
  return rc;
}

/* CPU-baseline timing hook: the chebyshev_recur_ll loop alone (what the reference's g_timer labels
 * '<PSI_0|PSI_n>', recursion.f90:3121-3127), wall-clock seconds for `nsteps` steps of one random-phase vector. */
int orc_cheb_time_steps(orc_ctx *c, const double *phases, int nsteps, double a, double b, double *seconds) {
  const int kk = c->kk;
  cplx *psiref = (cplx *)malloc(sizeof(cplx) * (size_t)BLK * kk);
  cplx *mu = (cplx *)calloc((size_t)BLK * (2 * nsteps + 2), sizeof(cplx));
  for (int i = 0; i <= kk; i++) c->izero[i] = (i > 0);
  zero_blocks(c->psi1, kk); zero_blocks(c->psi2, kk);
  random_start(c, phases, psiref);
  memcpy(c->psi0, psiref, sizeof(cplx) * (size_t)BLK * kk);
  cheb_0th_mom(c, psiref);
  memcpy(mu, c->cheb_mom_temp, sizeof(cplx) * BLK);
  cheb_1st_mom(c, psiref, a, b);
  memcpy(mu + BLK, c->cheb_mom_temp, sizeof(cplx) * BLK);
  memcpy(c->izero, c->idum, sizeof(int32_t) * (kk + 1));
#ifdef _OPENMP
  double t0 = omp_get_wtime();
#else
  double t0 = 0.0;
#endif
  for (int ll = 1; ll <= nsteps; ll++) chebyshev_recur_ll(c, ll, a, b, mu);
#ifdef _OPENMP
  *seconds = omp_get_wtime() - t0;
#else
  *seconds = -1.0;
#endif
  double chk = creal(mu[(size_t)BLK * (2 * nsteps)]);
  free(psiref); free(mu);
  return chk == chk ? 0 : -1;
}

/* ===================== H~ and velocity applications, Kubo-Bastin moments ===================== */
/* ham_vec_matmul / ham_hoh_vec_matmul: recursion.f90:913-977 / 785-911 (uses the MEMBER idum; caller copies) */
static void ham_apply(orc_ctx *c, const cplx *in, cplx *out, double a, double b) {
  const int kk = c->kk;
  const size_t n = (size_t)BLK * kk;
  zero_blocks(out, kk);
  if (c->hoh) {
    hoh_apply(c, in, out);
  } else {
#pragma omp parallel for
    for (int i = 1; i <= kk; i++) spmv_site(c, i, c->ee, c->hall, in, out, 2, c->izero, c->idum);
  }
#pragma omp parallel for
  for (size_t k = 0; k < n; k++) { out[k] = out[k] - b * in[k]; out[k] = out[k] / a; }
}

/* velo_vec_matmul / velo_hoh_vec_matmul: recursion.f90:587-654 / 656-783.  Local-H region not implemented in
 * the reference: loops start at nmax+1 and idum(1..nmax) keeps stale values. */
static void velo_apply(orc_ctx *c, const cplx *v_op, const cplx *vo_op, const cplx *in, cplx *out) {
  const int kk = c->kk;
  zero_blocks(out, kk);
  if (!c->hoh) {
#pragma omp parallel for
    for (int i = c->nmax + 1; i <= kk; i++) spmv_site(c, i, v_op, NULL, in, out, 1, c->izero, c->idum);
    return;
  }
  /* hoh: psi2 = v*in, psi1 = ee*in (both incl. on-site), izero=idum, hohpsi = sum_{nb>=2} vo(nb)*psi1(nn),
   * out = psi2 - hohpsi (+ enupsi + socpsi which stay zero, 712-713) */
  cplx *t2 = c->psi2, *t1 = c->psi1;
  zero_blocks(t1, kk); zero_blocks(t2, kk);
  zero_blocks(c->hohpsi, kk); zero_blocks(c->enupsi, kk); zero_blocks(c->socpsi, kk);
#pragma omp parallel for
  for (int i = c->nmax + 1; i <= kk; i++) {
    spmv_site(c, i, v_op, NULL, in, t2, 1, c->izero, c->idum);
    spmv_site(c, i, c->ee, NULL, in, t1, 1, c->izero, c->idum);
  }
  memcpy(c->izero, c->idum, sizeof(int32_t) * (kk + 1));
#pragma omp parallel for
  for (int i = c->nmax + 1; i <= kk; i++) spmv_site(c, i, vo_op, NULL, t1, c->hohpsi, 0, c->izero, c->idum);
  const size_t n = (size_t)BLK * kk;
  for (size_t k = 0; k < n; k++) out[k] = t2[k] - c->hohpsi[k] + c->enupsi[k] + c->socpsi[k];
  zero_blocks(t1, kk); zero_blocks(t2, kk);
}

void orc_ham_vec_matmul(orc_ctx *c, const cplx *psi_in, cplx *psi_out, double a, double b, int32_t *izero) {
  memcpy(c->izero, izero, sizeof(int32_t) * (c->kk + 1));
  ham_apply(c, psi_in, psi_out, a, b);
  memcpy(izero, c->idum, sizeof(int32_t) * (c->kk + 1));
}
void orc_velo_vec_matmul(orc_ctx *c, int slot, const cplx *psi_in, cplx *psi_out, int32_t *izero) {
  memcpy(c->izero, izero, sizeof(int32_t) * (c->kk + 1));
  velo_apply(c, slot == 'a' ? c->v_a : c->v_b, slot == 'a' ? c->vo_a : c->vo_b, psi_in, psi_out);
  memcpy(izero, c->idum, sizeof(int32_t) * (c->kk + 1));
}

/* compute_moments_stochastic: recursion.f90:1105-1230.  msel/nsel (test infrastructure for full-size lattices, where the
 * M*M contractions take minutes on a CPU): when msel != NULL only the left indices m = msel[0..nsel-1] (1-based) are
 * contracted and the result is packed as mu(18,18,M,nsel,nstart); the two Chebyshev chains are the same either way. */
static int kubo_moments_impl(orc_ctx *c, int nstart, int start_kind, const int32_t *start_sites, const double *phases,
                             int M, double a, double b, const int32_t *msel, int nsel, cplx *mu_nm) {
  const int kk = c->kk;
  const size_t n = (size_t)BLK * kk;
  cplx *psiref = (cplx *)calloc(n, sizeof(cplx)), *w0 = (cplx *)calloc(n, sizeof(cplx));
  cplx *w1 = (cplx *)calloc(n, sizeof(cplx)), *w2 = (cplx *)calloc(n, sizeof(cplx));
  cplx *v0 = (cplx *)calloc(n, sizeof(cplx)), *v1 = (cplx *)calloc(n, sizeof(cplx)), *v2 = (cplx *)calloc(n, sizeof(cplx));
  cplx *right = (cplx *)calloc(n, sizeof(cplx));
  cplx *left = (cplx *)calloc(n * (size_t)M, sizeof(cplx));
  if (!left) return -1;
  for (int s = 0; s < nstart; s++) {
    memset(w0, 0, n * sizeof(cplx)); memset(w1, 0, n * sizeof(cplx)); memset(w2, 0, n * sizeof(cplx));
    memset(v0, 0, n * sizeof(cplx)); memset(v1, 0, n * sizeof(cplx)); memset(v2, 0, n * sizeof(cplx));
    int j = 0;
    if (start_kind == 0) {
      j = start_sites[s];
      clear_mask(c);
      c->izero[j] = 1;
      zero_blocks(psiref, kk);
      for (int m = 0; m < NB; m++) SBLK(psiref, j)[m + NB * m] = 1.0;
    } else {
      for (int i = 0; i <= kk; i++) c->izero[i] = (i > 0);
      random_start(c, phases + (size_t)kk * s, psiref);
    }
    /* left vectors <r|T_m(H) */
    for (int m = 1; m <= M; m++) {
      if (m == 1) {
        memcpy(w1, psiref, n * sizeof(cplx));
      } else if (m == 2) {
        memcpy(w0, w1, n * sizeof(cplx));
        ham_apply(c, w0, w1, a, b);
        memcpy(c->izero, c->idum, sizeof(int32_t) * (kk + 1));
      } else {
        ham_apply(c, w1, w2, a, b);
        memcpy(c->izero, c->idum, sizeof(int32_t) * (kk + 1));
        for (size_t k = 0; k < n; k++) { w2[k] = 2 * w2[k] - w0[k]; w0[k] = w1[k]; w1[k] = w2[k]; w2[k] = 0.0; }
      }
      memcpy(left + n * (size_t)(m - 1), w1, n * sizeof(cplx));
    }
    if (start_kind == 0) { clear_mask(c); c->izero[j] = 1; }
    velo_apply(c, c->v_b, c->vo_b, psiref, v0);
    memcpy(c->izero, c->idum, sizeof(int32_t) * (kk + 1));
    for (int nn_ = 1; nn_ <= M; nn_++) {
      if (nn_ == 1) {
        memcpy(v1, v0, n * sizeof(cplx));
      } else if (nn_ == 2) {
        memcpy(v0, v1, n * sizeof(cplx));
        ham_apply(c, v0, v1, a, b);
        memcpy(c->izero, c->idum, sizeof(int32_t) * (kk + 1));
      } else {
        ham_apply(c, v1, v2, a, b);
        memcpy(c->izero, c->idum, sizeof(int32_t) * (kk + 1));
        for (size_t k = 0; k < n; k++) { v2[k] = 2 * v2[k] - v0[k]; v0[k] = v1[k]; v1[k] = v2[k]; v2[k] = 0.0; }
      }
      velo_apply(c, c->v_a, c->vo_a, v1, right);
      memcpy(c->izero, c->idum, sizeof(int32_t) * (kk + 1));
      const int mcount = msel ? nsel : M;
#pragma omp parallel for schedule(dynamic)
      for (int mi = 1; mi <= mcount; mi++) {
        const int m = msel ? msel[mi - 1] : mi;
        cplx dum[BLK];
        for (int k = 0; k < BLK; k++) dum[k] = 0.0;
        for (int k = 1; k <= kk; k++) gemm_cn(dum, SBLK(left + n * (size_t)(m - 1), k), SBLK(right, k));
        /* mu_nm_stochastic(:,:,n,m,i) */
        memcpy(mu_nm + (size_t)BLK * ((size_t)(nn_ - 1) + (size_t)M * ((size_t)(mi - 1) + (size_t)mcount * s)), dum, sizeof(dum));
      }
    }
  }
  free(psiref); free(w0); free(w1); free(w2); free(v0); free(v1); free(v2); free(right); free(left);
  return 0;
}
int orc_kubo_moments(orc_ctx *c, int nstart, int start_kind, const int32_t *start_sites, const double *phases,
                     int M, double a, double b, cplx *mu_nm) {
  return kubo_moments_impl(c, nstart, start_kind, start_sites, phases, M, a, b, NULL, 0, mu_nm);
}
int orc_kubo_moments_cols(orc_ctx *c, int nstart, int start_kind, const int32_t *start_sites, const double *phases,
                          int M, double a, double b, const int32_t *msel, int nsel, cplx *mu_sel) {
  for (int i = 0; i < nsel; i++) if (msel[i] < 1 || msel[i] > M) return -2;
  return kubo_moments_impl(c, nstart, start_kind, start_sites, phases, M, a, b, msel, nsel, mu_sel);
}

/* ===================== create_ll_map (recursion.f90:3277-3303) ===================== */
/* izeroll: (0:kk, lld+1) int32, column-major; column 1 must hold the start mask (chebyshev_recur sets
 * izeroll(j,1) = 1, 3086-3087). */
void orc_create_ll_map(const orc_ctx *c, int lld, int32_t *izeroll) {
  const int kk = c->kk;
  const size_t ld = (size_t)kk + 1;
  int32_t *idumll = (int32_t *)calloc(ld, sizeof(int32_t));
  for (int ll = 1; ll <= lld; ll++) {
    memset(idumll, 0, ld * sizeof(int32_t));
    for (int i = 1; i <= kk; i++) {
      idumll[i] = izeroll[i + ld * (ll - 1)];
      const int nr = NN(c, i, 1);
      if (nr >= 2)
        for (int j = 2; j <= nr; j++) {
          const int nnmap = NN(c, i, j);
          if (nnmap != 0 && izeroll[nnmap + ld * (ll - 1)] != 0) idumll[i] = 1;
        }
    }
    memcpy(izeroll + ld * ll, idumll, ld * sizeof(int32_t));
  }
  free(idumll);
}

/* ===================== chebyshev_orbital_mod, moment part (recursion.f90:2901-3008) ===================== */
/* The reference loops `random` over ALL kk start sites and never zeroes mu_n_orb; here the start sites are an argument
 * and mu_n_orb(18,18,lld) is the plain sum over them (the caller divides by kk, 3010).  The left vector
 * i (Y H~ X - X H~ Y)|r> is built with ham_vec_matmul (the non-hoh routine, 2957/2973) even when hoh is set; the
 * Chebyshev chain uses ham_hoh_vec_matmul when hoh (2984-2996).  cr: (3,kk) lattice%cr. */
static void ham_apply_nohoh(orc_ctx *c, const cplx *in, cplx *out, double a, double b) {
  const int hoh = c->hoh;
  c->hoh = 0;
  ham_apply(c, in, out, a, b);
  c->hoh = hoh;
}
int orc_orbital_moments(orc_ctx *c, int nstart, const int32_t *start_sites, const double *cr, double alat, int lld,
                        double a, double b, cplx *mu_n_orb) {
  const int kk = c->kk;
  const size_t n = (size_t)BLK * kk;
  cplx *psiref = (cplx *)calloc(n, sizeof(cplx)), *w0 = (cplx *)calloc(n, sizeof(cplx));
  cplx *l1 = (cplx *)calloc(n, sizeof(cplx)), *l2 = (cplx *)calloc(n, sizeof(cplx)), *left = (cplx *)calloc(n, sizeof(cplx));
  cplx *v0 = (cplx *)calloc(n, sizeof(cplx)), *v1 = (cplx *)calloc(n, sizeof(cplx)), *v2 = (cplx *)calloc(n, sizeof(cplx));
  memset(mu_n_orb, 0, sizeof(cplx) * (size_t)BLK * lld);
  for (int s = 0; s < nstart; s++) {
    const int r = start_sites[s];
    memset(v0, 0, n * sizeof(cplx)); memset(v1, 0, n * sizeof(cplx)); memset(v2, 0, n * sizeof(cplx));
    zero_blocks(psiref, kk);
    for (int i = 0; i <= kk; i++) c->izero[i] = (i > 0); /* this%izero(:) = 1 */
    for (int m = 0; m < NB; m++) SBLK(psiref, r)[m + NB * m] = 1.0;
    for (int k = 1; k <= kk; k++)
      for (int e = 0; e < BLK; e++) SBLK(l1, k)[e] = cr[0 + 3 * (k - 1)] * SBLK(psiref, k)[e] * alat;
    ham_apply_nohoh(c, l1, w0, a, b);
    for (int k = 1; k <= kk; k++)
      for (int e = 0; e < BLK; e++) SBLK(l1, k)[e] = cr[1 + 3 * (k - 1)] * SBLK(w0, k)[e] * alat;
    for (int k = 1; k <= kk; k++)
      for (int e = 0; e < BLK; e++) SBLK(l2, k)[e] = cr[1 + 3 * (k - 1)] * SBLK(psiref, k)[e] * alat;
    ham_apply_nohoh(c, l2, w0, a, b);
    for (int k = 1; k <= kk; k++)
      for (int e = 0; e < BLK; e++) SBLK(l2, k)[e] = cr[0 + 3 * (k - 1)] * SBLK(w0, k)[e] * alat;
    for (size_t e = 0; e < n; e++) left[e] = I * (l1[e] - l2[e]);
    for (int nn_ = 1; nn_ <= lld; nn_++) {
      if (nn_ == 1) {
        memcpy(v1, psiref, n * sizeof(cplx));
      } else if (nn_ == 2) {
        memcpy(v0, v1, n * sizeof(cplx));
        ham_apply(c, v0, v1, a, b);
        memcpy(c->izero, c->idum, sizeof(int32_t) * (kk + 1));
      } else {
        ham_apply(c, v1, v2, a, b);
        memcpy(c->izero, c->idum, sizeof(int32_t) * (kk + 1));
        for (size_t k = 0; k < n; k++) { v2[k] = 2 * v2[k] - v0[k]; v0[k] = v1[k]; v1[k] = v2[k]; v2[k] = 0.0; }
      }
      cplx dum[BLK];
      for (int k = 0; k < BLK; k++) dum[k] = 0.0;
      for (int k = 1; k <= kk; k++) gemm_cn(dum, SBLK(left, k), SBLK(v1, k));
      for (int k = 0; k < BLK; k++) mu_n_orb[k + (size_t)BLK * (nn_ - 1)] += dum[k];
    }
  }
  free(psiref); free(w0); free(l1); free(l2); free(left); free(v0); free(v1); free(v2);
  return 0;
}

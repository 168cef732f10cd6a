/*
 * rsrec_oracle_lattice.c -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE) for the neighbour-table construction
 * that feeds the recursion ("analytic nn builder"): a plain-C restatement of
 *   lattice%nncal              lattice.f90:3035-3123   (O(kk^2) pair loop, cut-off test of `mapa`, 2956-2973)
 *   lattice%f_wrap_coord_diff  lattice.f90:2975-3018   (minimum image over the 27 supercell shifts)
 *   lattice%remd               lattice.f90:2823-2907   (slots reordered to the representative atom's vector set)
 * PARITY: the table it builds for the reference's bccFe cluster feeds the golden-value reproduction of
 * tests/test_reference_golden.py (see rsrec_oracle.h); otherwise pinned by tests/test_oracle_lattice.py (numpy
 * restatement + invariants).
 * Compiled with -ffp-contract=off (distance tests are branch decisions).
 */
#include "rsrec_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  int b[3], n[3];
  const double *a; /* (3,3) column-major lattice%a */
  double alat;
} pbc_t;

/* cdiff = minimum image of coord(:,j) - coord(:,i); 1-based i, j */
static void wrap_diff(const pbc_t *p, const double *crd, int i, int j, double *cdiff) {
  double odiff[3], mdiff[3];
  for (int l = 0; l < 3; l++) { odiff[l] = crd[l + 3 * (j - 1)] - crd[l + 3 * (i - 1)]; mdiff[l] = odiff[l]; }
  const int xmax = p->b[0] ? 1 : 0, ymax = p->b[1] ? 1 : 0, zmax = p->b[2] ? 1 : 0;
  for (int z = -zmax; z <= zmax; z++)
    for (int y = -ymax; y <= ymax; y++)
      for (int x = -xmax; x <= xmax; x++) {
        double os[3];
        for (int l = 0; l < 3; l++)
          os[l] = odiff[l] + (double)(x * p->n[0]) * p->a[l + 3 * 0] * p->alat + (double)(y * p->n[1]) * p->a[l + 3 * 1] * p->alat +
                  (double)(z * p->n[2]) * p->a[l + 3 * 2] * p->alat;
        const double no = sqrt(os[0] * os[0] + os[1] * os[1] + os[2] * os[2]);
        const double nm = sqrt(mdiff[0] * mdiff[0] + mdiff[1] * mdiff[1] + mdiff[2] * mdiff[2]);
        if (no < nm) { mdiff[0] = os[0]; mdiff[1] = os[1]; mdiff[2] = os[2]; }
      }
  cdiff[0] = mdiff[0]; cdiff[1] = mdiff[1]; cdiff[2] = mdiff[2];
}

static int mapa(double r2, double ct) {
  const double ctm = (ct + ct) / 2.;
  const double ctsm = ctm * ctm;
  return r2 >= ctsm ? 0 : 1;
}

/* nncal + remd.  crd (3,kk) = cr*alat; no (kk) bravais type of each site (lattice%num); iu (ntot) representative site of
 * each bravais type; use_pbc: 0 = open cluster.  nn (kk, ncols) column-major int32 out, *nm_out = nnmax (nn needs
 * nm+1 columns like lattice%nn).  Returns 0 ok, -1 ncols too small (nothing written), -2 "VECTOR NOT FOUND",
 * -3 "TYPE NO NOT FOUND". */
int orc_build_nn(int kk, const double *crd, const int32_t *no, int ntot, const int32_t *iu, double ct, int use_pbc,
                 const int *b, const int *nrep, const double *a, double alat, int ncols, int32_t *nn, int *nm_out) {
  pbc_t p;
  for (int l = 0; l < 3; l++) { p.b[l] = use_pbc ? b[l] : 0; p.n[l] = nrep ? nrep[l] : 1; }
  p.a = a; p.alat = alat;
  /* nncal with a growable row store: cnt[i] = nn(i,1), rows[i][..] = nn(i,2..) */
  int *cnt = (int *)malloc(sizeof(int) * (kk + 1)), *cap = (int *)calloc(kk + 1, sizeof(int));
  int32_t **rows = (int32_t **)calloc(kk + 1, sizeof(int32_t *));
  int nnmax = 0;
  for (int i = 1; i <= kk; i++) cnt[i] = 1;
#define PUSH(I, J) do { if (cnt[I] - 1 >= cap[I]) { cap[I] = cap[I] ? 2 * cap[I] : 32; rows[I] = (int32_t *)realloc(rows[I], sizeof(int32_t) * cap[I]); } \
                        rows[I][cnt[I] - 1] = (J); cnt[I]++; if (cnt[I] > nnmax) nnmax = cnt[I]; } while (0)
  for (int i = 2; i <= kk; i++)
    for (int j = 1; j <= i - 1; j++) {
      double d[3], r2 = 0.0;
      if (use_pbc) {
        wrap_diff(&p, crd, i, j, d);
        r2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
      } else {
        for (int l = 0; l < 3; l++) { d[l] = crd[l + 3 * (i - 1)] - crd[l + 3 * (j - 1)]; r2 = r2 + d[l] * d[l]; }
      }
      if (mapa(r2, ct) != 0) { PUSH(i, j); PUSH(j, i); }
    }
  *nm_out = nnmax;
  int rc = 0;
  if (nnmax + 1 > ncols) rc = -1;
  /* remd */
  double *set = NULL;
  if (rc == 0) {
    set = (double *)calloc((size_t)3 * (ntot + 1) * (nnmax + 2), sizeof(double));
#define SET(m, t, j) set[(m) + 3 * ((size_t)(t) + (size_t)(ntot + 1) * (j))]
    for (int t = 1; t <= ntot; t++) {
      const int la = iu[t - 1];
      for (int j = 2; j <= cnt[la]; j++) {
        const int jj = rows[la][j - 2];
        if (use_pbc) { double d[3]; wrap_diff(&p, crd, la, jj, d); for (int m = 0; m < 3; m++) SET(m, t, j) = d[m]; }
        else for (int m = 0; m < 3; m++) SET(m, t, j) = crd[m + 3 * (la - 1)] - crd[m + 3 * (jj - 1)];
      }
    }
    memset(nn, 0, sizeof(int32_t) * (size_t)kk * ncols);
    int32_t *idnn = (int32_t *)malloc(sizeof(int32_t) * (nnmax + 2));
    for (int i = 1; i <= kk && rc == 0; i++) {
      const int n = no[i - 1];
      int ino = 0;
      for (int lk = 1; lk <= ntot; lk++) { ino = iu[lk - 1]; if (no[ino - 1] == n) break; ino = 0; }
      if (!ino || n < 1 || n > ntot) { rc = -3; break; }
      const int imax = cnt[ino];
      for (int k = 1; k <= imax; k++) idnn[k] = 0;
      for (int j = 2; j <= cnt[i]; j++) {
        const int jj = rows[i][j - 2];
        double ret[3];
        if (use_pbc) wrap_diff(&p, crd, i, jj, ret);
        else for (int m = 0; m < 3; m++) ret[m] = crd[m + 3 * (i - 1)] - crd[m + 3 * (jj - 1)];
        int k = 0;
        for (int ii = 2; ii <= imax; ii++) {
          const double a1 = ret[0] - SET(0, n, ii), a2 = ret[1] - SET(1, n, ii), a3 = ret[2] - SET(2, n, ii);
          const double aaa = a1 * a1 + a2 * a2 + a3 * a3;
          if (aaa < (double).0001f) { k = ii; break; } /* eps = .0001 is a default-real (single) literal */
        }
        if (!k) { rc = -2; break; }
        idnn[k] = jj;
      }
      if (rc) break;
      nn[(i - 1)] = imax;
      for (int j = 2; j <= imax; j++) nn[(i - 1) + (size_t)kk * (j - 1)] = idnn[j];
    }
    free(idnn);
  }
  for (int i = 1; i <= kk; i++) free(rows[i]);
  free(rows); free(cnt); free(cap); free(set);
  return rc;
}

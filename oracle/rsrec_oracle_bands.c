/*
 * rsrec_oracle_bands.c -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE) for the consumer of the on-site Green
 * function in the SCF loop, `type bands`.  See rsrec_oracle.h.  Plain-C restatement, statement for statement, of
 *   calculate_fermi (DOS loops + the two `fermi` calls)   bands.f90:227-347      fermi            bands.f90:366-402
 *   calculate_band_energy                                 bands.f90:354-359      simpson_m        math.f90:1579-1598
 *   calculate_projected_dos + calculate_magnetic_moments  bands.f90:1158-1181, 791-855 (integrals only)
 *   calculate_moments (dspd + the three Simpson moments)  bands.f90:409-524
 *   calculate_orbital_moments                             bands.f90:1075-1156 (imtrace(L g0) + integral)
 * PARITY: orc_bands_dos reproduces the reference's stored totaldos.out values of the bccFe regression case
 * (tests/test_reference_golden.py); the integrals are pinned by closed forms and a numpy restatement
 * (tests/test_oracle_bands.py).  Compiled with -ffp-contract=off (the Fermi scan is a branch decision).
 */
#include "rsrec_oracle.h"
#include <math.h>
#include <string.h>
#include <stdlib.h>

#define PI_RP 3.14159265358979323846
#define G0(r, c, ie, u) g0[(size_t)(r) + 18 * ((size_t)(c) + 18 * ((size_t)(ie) + (size_t)nv * (u)))]

/* dtot (nv), dosia (nv,nunits), dosial (18,nv,nunits); either of the last two may be NULL */
void orc_bands_dos(const orc_cplx *g0, int nv, int nunits, double *dtot, double *dosia, double *dosial) {
  for (int i = 0; i < nv; i++) dtot[i] = 0.0;
  for (int u = 0; u < nunits; u++)
    for (int i = 0; i < nv; i++) {
      double d = 0.0;
      for (int j = 0; j < 9; j++) {
        const double up = cimag(G0(j, j, i, u)), dn = cimag(G0(j + 9, j + 9, i, u));
        dtot[i] = dtot[i] - (up + dn) / PI_RP;
        d = d - (up + dn) / PI_RP;
        if (dosial) {
          dosial[j + 18 * ((size_t)i + (size_t)nv * u)] = -up / PI_RP;
          dosial[j + 9 + 18 * ((size_t)i + (size_t)nv * u)] = -dn / PI_RP;
        }
      }
      if (dosia) dosia[(size_t)i + (size_t)nv * u] = d;
    }
}

/* `fermi` (bands.f90:366-402); y is 1-based in the reference */
void orc_fermi(double *ef, double h, int *ik1, double ainf, int npts, const double *y, int *ifail, double qqv, double *e1) {
  double aint = 0.0, aint0 = 0.0;
  int i;
  *ifail = 1;
  for (i = 2; i <= npts - 1; i += 2) {
    aint = aint + h * (y[i - 2] + 4.0 * y[i - 1] + y[i]) / 3.0;
    if (aint >= qqv) goto found;
    aint0 = aint;
  }
  return;
found:
  *ifail = 0;
  if (aint == qqv) {
    *ik1 = i + 1;
    *ef = ainf + h * i;
    *e1 = *ef;
  } else {
    const double alpha = (aint - aint0) / 2.0 / h;
    *ik1 = i - 1;
    *e1 = ainf + h * (i - 2);
    *ef = ((qqv - aint0) / alpha) + *e1;
  }
}

/* Fermi part of calculate_fermi (bands.f90:322-342): fermi in/out, nv1 in (en%ik1) / out, e1 out */
void orc_bands_fermi(const double *dtot, int nv, double edel, double energy_min, double qqv, int fix_fermi, double *fermi,
                     int *nv1, double *e1, int *ifail) {
  *ifail = 0;
  if (!fix_fermi) {
    double ef_mag = *fermi, e1_mag = *fermi;
    int ik1_mag = 0, ik1 = *nv1;
    orc_fermi(&ef_mag, edel, &ik1_mag, energy_min, nv, dtot, ifail, qqv, &e1_mag);
    orc_fermi(fermi, edel, &ik1, energy_min, nv, dtot, ifail, qqv, &e1_mag);
    *nv1 = ik1;
    *e1 = e1_mag;
  } else {
    const int ik1 = (int)lround((*fermi - energy_min) / edel);
    *e1 = energy_min + (ik1 - 1) * edel;
    *nv1 = ik1;
  }
}

static double ipow(double x, int n) { return n == 0 ? 1.0 : n == 1 ? x : x * x; }

/* simpson_m (math.f90:1579-1598) */
double orc_simpson_m(double h, double ef, int npts, const double *y, double ea, int nexp, const double *ene) {
  double aint = 0.0;
  for (int i = 2; i <= npts - 1; i += 2)
    aint = aint + y[i - 2] * ipow(ene[i - 2], nexp) + 4.0 * y[i - 1] * ipow(ene[i - 1], nexp) + y[i] * ipow(ene[i], nexp);
  aint = h * aint / 3.0;
  if (ea != ef)
    aint = aint + (ef - ea) * (y[npts - 1] * ipow(ene[npts - 1], nexp) + 4.0 * y[npts] * ipow(ene[npts], nexp) +
                               y[npts + 1] * ipow(ene[npts + 1], nexp)) / 6.0;
  return aint;
}

/* mom0, mom1 (3,nunits): mx,my,mz and potential%mom1 */
void orc_bands_magnetic_moments(const orc_cplx *g0, int nv, int nunits, const double *ene, double edel, double fermi, int nv1,
                                double e1, double *mom0, double *mom1) {
  double *d = (double *)malloc(sizeof(double) * 3 * (size_t)nv);
  for (int u = 0; u < nunits; u++) {
    double *dx = d, *dy = d + nv, *dz = d + 2 * (size_t)nv;
    for (int ie = 0; ie < nv; ie++) {
      dx[ie] = dy[ie] = dz[ie] = 0.0;
      for (int i = 0; i < 9; i++) {
        dz[ie] = dz[ie] - cimag(G0(i, i, ie, u) - G0(i + 9, i + 9, ie, u)) / PI_RP;
        dy[ie] = dy[ie] - cimag(I * G0(i, i + 9, ie, u) - I * G0(i + 9, i, ie, u)) / PI_RP;
        dx[ie] = dx[ie] - cimag(G0(i, i + 9, ie, u) + G0(i + 9, i, ie, u)) / PI_RP;
      }
    }
    for (int k = 0; k < 3; k++) {
      mom0[k + 3 * u] = orc_simpson_m(edel, fermi, nv1, d + (size_t)k * nv, e1, 0, ene);
      mom1[k + 3 * u] = orc_simpson_m(edel, fermi, nv1, d + (size_t)k * nv, e1, 1, ene);
    }
  }
  free(d);
}

/* mom (3,nunits); lsph (9,9,3) = hcpx(L_x), hcpx(L_y), hcpx(L_z); occ (3,6,nunits) = sgef,pmef,smef; lmom (3,nunits) */
void orc_bands_moments(const orc_cplx *g0, int nv, int channels_ldos, int nunits, const double *mom, const orc_cplx *lsph,
                       const double *ene, double edel, double fermi, int nv1, double e1, double *occ, double *lmom) {
  double *y = (double *)malloc(sizeof(double) * (size_t)nv);
  for (int u = 0; u < nunits; u++) {
    for (int isp = 1; isp <= 2; isp++) {
      const double isgn = isp == 1 ? 1.0 : -1.0;
      for (int l = 1; l <= 3; l++) {
        for (int ie = 0; ie < nv; ie++) y[ie] = 0.0;
        for (int m = 1; m <= 2 * l - 1; m++) {
          const int o = (l - 1) * (l - 1) + m - 1;
          for (int ie = 0; ie < channels_ldos; ie++)
            y[ie] = y[ie] - cimag(G0(o, o, ie, u) + G0(o + 9, o + 9, ie, u)) -
                    isgn * mom[2 + 3 * u] * cimag(G0(o, o, ie, u) - G0(o + 9, o + 9, ie, u)) -
                    isgn * mom[1 + 3 * u] * cimag(I * G0(o, o + 9, ie, u) - I * G0(o + 9, o, ie, u)) -
                    isgn * mom[0 + 3 * u] * cimag(G0(o, o + 9, ie, u) + G0(o + 9, o, ie, u));
        }
        for (int ie = 0; ie < nv; ie++) y[ie] = y[ie] * 0.5 / PI_RP;
        const int q = l - 1 + 3 * (isp - 1);
        for (int k = 0; k < 3; k++) occ[k + 3 * (q + 6 * (size_t)u)] = orc_simpson_m(edel, fermi, nv1, y, e1, k, ene);
      }
    }
    for (int dir = 0; dir < 3; dir++) {
      const orc_cplx *L = lsph + 81 * dir;
      for (int ie = 0; ie < nv; ie++) {  /* imtrace(matmul(mL_ext, g0)) */
        double complex tr = 0.0;
        for (int sp = 0; sp < 2; sp++)
          for (int i = 0; i < 9; i++)
            for (int k = 0; k < 9; k++) tr += L[i + 9 * k] * G0(k + 9 * sp, i + 9 * sp, ie, u);
        y[ie] = cimag(tr);
      }
      lmom[dir + 3 * u] = -(orc_simpson_m(edel, fermi, nv1, y, e1, 0, ene) / PI_RP);
    }
  }
  free(y);
}

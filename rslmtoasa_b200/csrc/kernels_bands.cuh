// kernels_bands.cuh -- the consumer of the on-site Green function g0 in the SCF loop (`type bands`, bands.f90): the
// energy-resolved projections of g0 and their Simpson integrals up to the Fermi level, so that g0 (18*18*nv complex per
// unit) never leaves the device and only the band moments / charges / band energy go back to the host.
//
//   k_bands_dtot       total DOS of calculate_fermi (bands.f90:260-273), one thread per energy, the reference's own
//                      summation order (units outer, orbitals inner) with explicit IEEE operations: bit-exact for
//                      identical g0, so the Fermi search below takes the reference's branches
//   k_bands_ldos       dosia / dosial (per-unit and per-orbital DOS) of the same loop
//   k_bands_fermi      `fermi` (bands.f90:366-402): sequential Simpson scan of dtot, twice, as calculate_fermi calls it
//   k_bands_spin       dx, dy, dz of calculate_projected_dos (1158-1181)
//   k_bands_dspd       dspd(l + 3(isp-1), ie, unit) of calculate_moments (438-452) and Im Tr(L_d g0) of
//                      calculate_orbital_moments (1116-1120)
//   k_bands_simpson    simpson_m (math.f90:1579-1598): one warp per integral, fixed-order tree (deterministic)
#pragma once
#include "kernels_ham.cuh"

__constant__ double2 c_bands_L[3][81];  // hcpx(L_x), hcpx(L_y), hcpx(L_z), col-major 9x9

// L_x, L_y, L_z of math.f90:133-165 (real 9x9 tables times -i) taken to the spherical basis like
// calculate_orbital_moments does (bands.f90:1094-1101)
static int bands_configure() {
  const double s3 = sqrt(3.0);
  double L[3][81];
  memset(L, 0, sizeof(L));
  auto S = [&](int d, int row, int col, double v) { L[d][(row - 1) + 9 * (col - 1)] = v; };
  // column c of each table lists (row, value)
  S(0, 4, 3, -1); S(0, 3, 4, 1); S(0, 7, 5, -1); S(0, 8, 6, -1); S(0, 9, 6, -s3); S(0, 5, 7, 1); S(0, 6, 8, 1); S(0, 6, 9, s3);
  S(1, 4, 2, 1); S(1, 2, 4, -1); S(1, 6, 5, 1); S(1, 5, 6, -1); S(1, 8, 7, -1); S(1, 9, 7, s3); S(1, 7, 8, 1); S(1, 7, 9, -s3);
  S(2, 3, 2, -1); S(2, 2, 3, 1); S(2, 8, 5, 2); S(2, 7, 6, 1); S(2, 6, 7, -1); S(2, 5, 8, -2);
  double2 v[81], vc[81], out[3][81];
  hcpx_host_matrices(v, vc);
  for (int d = 0; d < 3; d++) {
    double2 tmp[81];
    for (int j = 0; j < 9; j++)
      for (int i = 0; i < 9; i++) {
        double2 s = make_double2(0.0, 0.0);
        for (int k = 0; k < 9; k++) {  // (L * -i)(i,k) * v(k,j)
          const double2 l = make_double2(0.0, -L[d][i + 9 * k]);
          const double2 w = v[k + 9 * j];
          s.x += l.x * w.x - l.y * w.y; s.y += l.x * w.y + l.y * w.x;
        }
        tmp[i + 9 * j] = s;
      }
    for (int j = 0; j < 9; j++)
      for (int i = 0; i < 9; i++) {
        double2 s = make_double2(0.0, 0.0);
        for (int k = 0; k < 9; k++) {
          const double2 a = vc[i + 9 * k], b = tmp[k + 9 * j];
          s.x += a.x * b.x - a.y * b.y; s.y += a.x * b.y + a.y * b.x;
        }
        out[d][i + 9 * j] = s;
      }
  }
  return cudaMemcpyToSymbol(c_bands_L, out, sizeof(out)) == cudaSuccess ? 0 : -1;
}

// g0: (18,18,nv,nunits) complex col-major.  element (r,c) 0-based of energy ie, unit u
__device__ __forceinline__ double2 g0_at(const double2 *__restrict__ g0, int r, int c, size_t blk) { return __ldg(g0 + blk * 324 + r + 18 * c); }

__global__ void k_bands_dtot(const double2 *__restrict__ g0, int nv, int nunits, double *__restrict__ dtot) {
  const int ie = blockIdx.x * blockDim.x + threadIdx.x;
  if (ie >= nv) return;
  double d = 0.0;
  for (int u = 0; u < nunits; u++) {
    const size_t blk = (size_t)u * nv + ie;
    for (int j = 0; j < 9; j++)
      d = __dsub_rn(d, __ddiv_rn(__dadd_rn(g0_at(g0, j, j, blk).y, g0_at(g0, j + 9, j + 9, blk).y), PI_RP));
  }
  dtot[ie] = d;
}

// dosial: (18, nv, nunits); dosia: (nv, nunits)
__global__ void k_bands_ldos(const double2 *__restrict__ g0, int nv, int nunits, double *__restrict__ dosia, double *__restrict__ dosial) {
  const size_t blk = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // = u*nv + ie
  if (blk >= (size_t)nv * nunits) return;
  double d = 0.0;
  for (int j = 0; j < 9; j++) {
    const double up = g0_at(g0, j, j, blk).y, dn = g0_at(g0, j + 9, j + 9, blk).y;
    d = __dsub_rn(d, __ddiv_rn(__dadd_rn(up, dn), PI_RP));
    if (dosial) { dosial[blk * 18 + j] = __ddiv_rn(-up, PI_RP); dosial[blk * 18 + j + 9] = __ddiv_rn(-dn, PI_RP); }
  }
  if (dosia) dosia[blk] = d;
}

// `fermi` (bands.f90:366-402).  pan[j] = H*(Y(I-1) + 4 Y(I) + Y(I+1))/3 of panel I = 2 j + 2 (precomputed by all threads
// with the reference's operations); the scan itself is a serial add-and-compare.  y = dtot(1..npts) is only needed for
// nothing else, so the caller passes the panel values.
__device__ void bands_fermi_scan(double &ef, double h, int &ik1, double ainf, int npts, const double *pan, int &ifail, double qqv, double &e1) {
  ifail = 1;
  double aint = 0.0, aint0 = 0.0;
  int i;
  bool hit = false;
  for (i = 2; i <= npts - 1; i += 2) {
    aint = __dadd_rn(aint, pan[(i >> 1) - 1]);
    if (aint >= qqv) { hit = true; break; }
    aint0 = aint;
  }
  if (!hit) return;
  ifail = 0;
  if (aint == qqv) {
    ik1 = i + 1;
    ef = __dadd_rn(ainf, __dmul_rn(h, (double)i));
    e1 = ef;
  } else {
    const double alpha = __ddiv_rn(__ddiv_rn(__dsub_rn(aint, aint0), 2.0), h);
    ik1 = i - 1;
    e1 = __dadd_rn(ainf, __dmul_rn(h, (double)(i - 2)));
    ef = __dadd_rn(__ddiv_rn(__dsub_rn(qqv, aint0), alpha), e1);
  }
}

// the two calls of calculate_fermi (bands.f90:327-334): res = {fermi, e1, nv1, ifail}; pan: npts/2 doubles of scratch
// (shared memory when it fits, else global)
__global__ void k_bands_fermi(const double *__restrict__ dtot, int npts, double h, double ainf, double qqv, double fermi_in, int ik1_in,
                              double *__restrict__ res, double *__restrict__ pan_global) {
  extern __shared__ double fm_smem[];
  double *pan = pan_global ? pan_global : fm_smem;
  for (int i = 2 + 2 * threadIdx.x; i <= npts - 1; i += 2 * blockDim.x) {
    const double s = __dadd_rn(__dadd_rn(dtot[i - 2], __dmul_rn(4.0, dtot[i - 1])), dtot[i]);
    pan[(i >> 1) - 1] = __ddiv_rn(__dmul_rn(h, s), 3.0);
  }
  __syncthreads();
  if (blockIdx.x || threadIdx.x) return;
  double ef_mag = fermi_in, e1_mag = fermi_in, ef = fermi_in;
  int ik1_mag = 0, ik1 = ik1_in, ifail = 1;
  bands_fermi_scan(ef_mag, h, ik1_mag, ainf, npts, pan, ifail, qqv, e1_mag);
  bands_fermi_scan(ef, h, ik1, ainf, npts, pan, ifail, qqv, e1_mag);
  res[0] = ef; res[1] = e1_mag; res[2] = (double)ik1; res[3] = (double)ifail;
}

// y: (nv, 3, nunits): dx, dy, dz
__global__ void k_bands_spin(const double2 *__restrict__ g0, int nv, int nunits, double *__restrict__ y) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)nv * nunits * 3) return;
  const int ie = (int)(t % nv), d = (int)((t / nv) % 3), u = (int)(t / ((size_t)nv * 3));
  const size_t blk = (size_t)u * nv + ie;
  double s = 0.0;
  for (int i = 0; i < 9; i++) {
    double im;
    if (d == 2) im = g0_at(g0, i, i, blk).y - g0_at(g0, i + 9, i + 9, blk).y;
    else {
      const double2 ud = g0_at(g0, i, i + 9, blk), du = g0_at(g0, i + 9, i, blk);
      im = d == 0 ? ud.y + du.y : ud.x - du.x;  // aimag(ud + du) ; aimag(i*ud - i*du) = re(ud) - re(du)
    }
    s -= im / PI_RP;
  }
  y[t] = s;
}

// y: (nv, 9, nunits): q = 0..5 dspd(l + 3(isp-1)), q = 6..8 Im Tr(L_x g0), Im Tr(L_y g0), Im Tr(L_z g0)
__global__ void k_bands_dspd(const double2 *__restrict__ g0, int nv, int channels, int nunits, const double *__restrict__ mom, double *__restrict__ y) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)nv * nunits * 9) return;
  const int ie = (int)(t % nv), q = (int)((t / nv) % 9), u = (int)(t / ((size_t)nv * 9));
  const size_t blk = (size_t)u * nv + ie;
  double s = 0.0;
  if (q < 6) {
    if (ie < channels) {
      const int l = q % 3 + 1;
      const double isgn = q < 3 ? 1.0 : -1.0;
      const double m1 = mom[3 * u], m2 = mom[3 * u + 1], m3 = mom[3 * u + 2];
      for (int m = 1; m <= 2 * l - 1; m++) {
        const int o = (l - 1) * (l - 1) + m - 1;
        const double2 uu = g0_at(g0, o, o, blk), dd = g0_at(g0, o + 9, o + 9, blk), ud = g0_at(g0, o, o + 9, blk), du = g0_at(g0, o + 9, o, blk);
        s = s - (uu.y + dd.y) - isgn * m3 * (uu.y - dd.y) - isgn * m2 * (ud.x - du.x) - isgn * m1 * (ud.y + du.y);
      }
      s = s * 0.5 / PI_RP;
    }
  } else {
    const double2 *L = c_bands_L[q - 6];
    for (int sp = 0; sp < 2; sp++)
      for (int i = 0; i < 9; i++)
        for (int k = 0; k < 9; k++) {
          const double2 l = L[i + 9 * k];
          if (l.x == 0.0 && l.y == 0.0) continue;
          const double2 g = g0_at(g0, k + 9 * sp, i + 9 * sp, blk);
          s += l.x * g.y + l.y * g.x;
        }
  }
  y[t] = s;
}

// integral j (one warp): y_j = Y + j*nv, nexp = nexp0 + (j % nord) if nord > 0 ... here: integrand index = j / nord,
// order = ord0 + j % nord.  out[j] = simpson_m(h, ef, npts, y, ea, nexp, ene)
__global__ void k_bands_simpson(const double *__restrict__ Y, int nv, int nint, int nord, int ord0, const double *__restrict__ ene, double h,
                                double ef, int npts, double ea, double *__restrict__ out) {
  const int j = (int)(((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (j >= nint * nord) return;
  const double *y = Y + (size_t)(j / nord) * nv;
  const int nexp = ord0 + j % nord;
  auto term = [&](int i) {  // 1-based index
    const double e = ene[i - 1];
    return y[i - 1] * (nexp == 0 ? 1.0 : nexp == 1 ? e : e * e);
  };
  double s = 0.0;
  for (int i = 2 + 2 * lane; i <= npts - 1; i += 64) s += term(i - 1) + 4.0 * term(i) + term(i + 1);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) {
    s = h * s / 3.0;
    if (ea != ef) s += (ef - ea) * (term(npts) + 4.0 * term(npts + 1) + term(npts + 2)) / 6.0;
    out[j] = s;
  }
}

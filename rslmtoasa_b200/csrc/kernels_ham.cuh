// kernels_ham.cuh -- device-side assembly of the Hamiltonian block sets that feed the recursion,
// and the packing of complex 18x18 blocks into the HR36 real embedding the tensor-pipe kernels consume.
//
//   k_ham_blocks     ham0m_nc (hamiltonian.f90:2225-2303) + hcpx 'cart2sph' (math.f90:1508-1577) + the spin-block
//                    composition of build_bulkham / build_locham (1553-1616 / 1618-1667) for one (slot, class)
//   k_ham_onsite18   build_obarm / build_enim (1481-1551)
//   k_ham_times_o    eeo = ee * obarm(type of the atom in the slot)  (1597-1606, 1651-1660)
//   k_pack_hr36      complex block (+ optional on-site addend, sign, spin-diagonal mask) -> [[Hr,-Hi],[Hi,Hr]]
// Classes: 0..ntype-1 = atom types (ee), ntype.. = site-indexed local region (hall), as everywhere in the library.
#pragma once
#include "kernels_post.cuh"

#define POT_NPAR 12  // wx0 wx1 cx0 cx1 cex0 cex1 obx0 obx1 cx(:,1) cx(:,2) cex(:,1) cex(:,2), each (9) per type

__constant__ double2 c_hcpx_v[81], c_hcpx_vc[81];  // col-major 9x9: element (i,j) at i + 9 j

// V / V^H of hcpx (math.f90:1525-1558), col-major 9x9, on the host
static void hcpx_host_matrices(double2 *v, double2 *vc) {
  for (int e = 0; e < 81; e++) v[e] = vc[e] = make_double2(0.0, 0.0);
  const double c = 1.0 / sqrt(2.0);
  auto S = [](double2 *m, int i, int j, double re, double im) { m[(i - 1) + 9 * (j - 1)] = make_double2(re, im); };
  // base Y(lm) in the order (00)(1-1)(10)(11)(2-2)(2-1)(20)(21)(22), math.f90:1525-1558
  S(v, 1, 1, 1, 0); S(vc, 1, 1, 1, 0);
  S(v, 2, 4, -c, 0); S(vc, 4, 2, -c, 0); S(v, 2, 2, c, 0); S(vc, 2, 2, c, 0);
  S(v, 3, 4, 0, c); S(vc, 4, 3, 0, -c); S(v, 3, 2, 0, c); S(vc, 2, 3, 0, -c);
  S(v, 4, 3, 1, 0); S(vc, 3, 4, 1, 0);
  S(v, 5, 5, 0, c); S(v, 5, 9, 0, -c); S(v, 6, 6, 0, c); S(v, 6, 8, 0, c);
  S(v, 7, 6, c, 0); S(v, 7, 8, -c, 0); S(v, 8, 5, c, 0); S(v, 8, 9, c, 0); S(v, 9, 7, 1, 0);
  S(vc, 5, 5, 0, -c); S(vc, 9, 5, 0, c); S(vc, 6, 6, 0, -c); S(vc, 8, 6, 0, -c);
  S(vc, 6, 7, c, 0); S(vc, 8, 7, -c, 0); S(vc, 5, 8, c, 0); S(vc, 9, 8, c, 0); S(vc, 7, 9, 1, 0);
}

static int ham_configure() {
  double2 v[81], vc[81];
  hcpx_host_matrices(v, vc);
  if (cudaMemcpyToSymbol(c_hcpx_v, v, sizeof(v)) != cudaSuccess) return -1;
  if (cudaMemcpyToSymbol(c_hcpx_vc, vc, sizeof(vc)) != cudaSuccess) return -1;
  return 0;
}

// in-place hcpx 'cart2sph' of a 9x9 matrix in shared memory (col-major), 81 active threads, tmp = scratch
__device__ __forceinline__ void hcpx81(double2 *h, double2 *tmp, int t) {
  const int i = t % 9, j = t / 9;
  double2 s = make_double2(0.0, 0.0);
  if (t < 81) for (int k = 0; k < 9; k++) s = c_add(s, c_mul(h[i + 9 * k], c_hcpx_v[k + 9 * j]));
  __syncthreads();
  if (t < 81) tmp[t] = s;
  __syncthreads();
  s = make_double2(0.0, 0.0);
  if (t < 81) for (int k = 0; k < 9; k++) s = c_add(s, c_mul(c_hcpx_vc[i + 9 * k], tmp[k + 9 * j]));
  __syncthreads();
  if (t < 81) h[t] = s;
  __syncthreads();
}

// pot: (9, POT_NPAR, ntype) complex; mom: (3, ntype); hhh: (9,9,nslot,ncls) real; jt: (nslot,ncls); it: (ncls)
__global__ void __launch_bounds__(96)
k_ham_blocks(const double *__restrict__ hhh, const int32_t *__restrict__ jt, const int32_t *__restrict__ it,
             const double2 *__restrict__ pot, const double *__restrict__ mom, int hoh, int nslot, double2 *__restrict__ blk) {
  __shared__ double2 hh[4][81], tmp[81];
  const int m = blockIdx.x, c = blockIdx.y, t = threadIdx.x, ilm = t % 9, jlm = t / 9;
  double2 *out = blk + (size_t)BLKC * (m + (size_t)nslot * c);
  const int jtv = jt[m + nslot * c] - 1, itv = it[c] - 1;
  if (jtv < 0) {  // no atom in this slot: hmag stays zero (chbar_nc 2339)
    for (int e = t; e < BLKC; e += 96) out[e] = make_double2(0.0, 0.0);
    return;
  }
  const double *mi = mom + 3 * itv, *mj = mom + 3 * jtv;
  if (t < 81) {
    const double2 *pi = pot + (size_t)9 * POT_NPAR * itv, *pj = pot + (size_t)9 * POT_NPAR * jtv;
    const double2 wx0i = pi[ilm], wx1i = pi[9 + ilm], wx0j = pj[jlm], wx1j = pj[9 + jlm];
    const double2 hc = make_double2(hhh[t + 81 * (m + (size_t)nslot * c)], 0.0);
    const double2 dot = make_double2(mi[0] * mj[0] + mi[1] * mj[1] + mi[2] * mj[2], 0.0);
    const double cr[3] = {mi[1] * mj[2] - mi[2] * mj[1], mi[2] * mj[0] - mi[0] * mj[2], mi[0] * mj[1] - mi[1] * mj[0]};
    const double2 a00 = c_mul(c_mul(wx0i, hc), wx0j), a11 = c_mul(c_mul(wx1i, hc), wx1j);
    const double2 a10 = c_mul(c_mul(wx1i, hc), wx0j), a01 = c_mul(c_mul(wx0i, hc), wx1j);
    double2 h4 = c_add(a00, c_mul(a11, dot));
    const bool on = (m == 0) && ilm == jlm;  // norm2(vet) <= 0.01: the atom itself (slot 1)
    if (on) h4 = c_add(h4, pi[(hoh ? 4 : 2) * 9 + ilm]);  // cex0 | cx0
    hh[3][t] = h4;
    for (int d = 0; d < 3; d++) {
      // (wx1 S wx0) mom_i + (wx0 S wx1) mom_j + i wx1 S wx1 cross
      const double2 ia11 = c_mul(make_double2(-a11.y, a11.x), make_double2(cr[d], 0.0));
      double2 v = c_add(c_add(c_mul(a10, make_double2(mi[d], 0.0)), c_mul(a01, make_double2(mj[d], 0.0))), ia11);
      if (on) v = c_add(v, c_mul(pi[(hoh ? 5 : 3) * 9 + ilm], make_double2(mi[d], 0.0)));  // cex1 | cx1
      hh[d][t] = v;
    }
  }
  __syncthreads();
  for (int d = 0; d < 4; d++) hcpx81(hh[d], tmp, t);
  if (t < 81) {
    const int a = ilm, b = jlm;
    const double2 H1 = hh[0][t], H2 = hh[1][t], H3 = hh[2][t], H4 = hh[3][t];
    out[a + NB * b] = c_add(H4, H3);                                        // H0 + Hz
    out[(a + 9) + NB * (b + 9)] = c_sub(H4, H3);                            // H0 - Hz
    out[a + NB * (b + 9)] = make_double2(H1.x + H2.y, H1.y - H2.x);         // Hx - i Hy
    out[(a + 9) + NB * b] = make_double2(H1.x - H2.y, H1.y + H2.x);         // Hx + i Hy
  }
}

// which = 0: obarm from obx0/obx1; 1: enim from 0.5 (eu +- ed), eu = cx(:,1) - cex(:,1), ed = cx(:,2) - cex(:,2)
__global__ void __launch_bounds__(96)
k_ham_onsite18(const double2 *__restrict__ pot, const double *__restrict__ mom, double2 *__restrict__ obarm,
               double2 *__restrict__ enim) {
  __shared__ double2 q[4][81], tmp[81];
  const int ty = blockIdx.x, which = blockIdx.y, t = threadIdx.x, i = t % 9, j = t / 9;
  const double2 *p = pot + (size_t)9 * POT_NPAR * ty;
  const double *mm = mom + 3 * ty;
  if (t < 81) {
    double2 d0 = make_double2(0.0, 0.0), d1 = d0;
    if (i == j) {
      if (which == 0) { d0 = p[6 * 9 + i]; d1 = p[7 * 9 + i]; }
      else {
        const double2 eu = c_sub(p[8 * 9 + i], p[10 * 9 + i]), ed = c_sub(p[9 * 9 + i], p[11 * 9 + i]);
        d0 = c_scale(c_add(eu, ed), 0.5); d1 = c_scale(c_sub(eu, ed), 0.5);
      }
    }
    const double2 m3 = c_mul(d1, make_double2(mm[2], 0.0)), m1 = c_mul(d1, make_double2(mm[0], 0.0));
    const double2 m2 = c_mul(d1, make_double2(mm[1], 0.0)), im2 = make_double2(-m2.y, m2.x);
    q[0][t] = c_add(d0, m3);   // (m, l)
    q[1][t] = c_sub(d0, m3);   // (m+9, l+9)
    q[2][t] = c_sub(m1, im2);  // (l, m+9)   (diagonal: l == m)
    q[3][t] = c_add(m1, im2);  // (l+9, m)
  }
  __syncthreads();
  for (int d = 0; d < 4; d++) hcpx81(q[d], tmp, t);
  if (t < 81) {
    double2 *o = (which == 0 ? obarm : enim) + (size_t)BLKC * ty;
    o[i + NB * j] = q[0][t];
    o[(i + 9) + NB * (j + 9)] = q[1][t];
    o[i + NB * (j + 9)] = q[2][t];
    o[(i + 9) + NB * j] = q[3][t];
  }
}

// blko(:,:,m,c) = blk(:,:,m,c) * obarm(:,:,jt(m,c)); zero where the slot is empty
__global__ void __launch_bounds__(BLKC)
k_ham_times_o(const double2 *__restrict__ blk, const double2 *__restrict__ obarm, const int32_t *__restrict__ jt, int nslot,
              double2 *__restrict__ blko) {
  __shared__ double2 A[BLKC], O[BLKC];
  const int m = blockIdx.x, c = blockIdx.y, t = threadIdx.x, r = t % NB, k = t / NB;
  const size_t off = (size_t)BLKC * (m + (size_t)nslot * c);
  const int jtv = jt[m + nslot * c] - 1;
  if (jtv < 0) { blko[off + t] = make_double2(0.0, 0.0); return; }
  A[t] = blk[off + t];
  O[t] = obarm[(size_t)BLKC * jtv + t];
  __syncthreads();
  double2 s = make_double2(0.0, 0.0);
  for (int l = 0; l < NB; l++) s = c_add(s, c_mul(A[r + NB * l], O[l + NB * k]));
  blko[off + t] = s;
}

// dst[(c*nslot + m)] = HR36( scale * mask(src(:,:,m,c)) ) + (m == 0 ? HR36(add(:,:,cls_type(c))) : 0)
// src: class-major blocks (18,18,nslot,ncls_src); classes >= ncls_src produce zero blocks (velocity sets);
// skip_onsite: slot 0 left zero (vo operator); per_class: src is (18,18,ntype) indexed by cls_type (Hx = enim).
__global__ void __launch_bounds__(BLKC)
k_pack_hr36(const double2 *__restrict__ src, int ncls_src, const double2 *__restrict__ add, const int32_t *__restrict__ cls_type,
            int nslot, double scale, int spin_diag, int skip_onsite, int per_class, double *__restrict__ dst) {
  const int m = blockIdx.x, c = blockIdx.y, t = threadIdx.x, r = t % NB, k = t / NB;
  double2 v = make_double2(0.0, 0.0);
  if (c < ncls_src && !(skip_onsite && m == 0)) {
    v = per_class ? src[(size_t)BLKC * cls_type[c] + t] : src[(size_t)BLKC * (m + (size_t)nslot * c) + t];
    if (spin_diag && ((r < 9) != (k < 9))) v = make_double2(0.0, 0.0);
    v = make_double2(scale * v.x, scale * v.y);
  }
  if (add && m == 0) { const double2 a = add[(size_t)BLKC * cls_type[c] + t]; v = make_double2(v.x + a.x, v.y + a.y); }
  double *d = dst + ((size_t)c * nslot + m) * HBLK;
  d[r * COLD + k] = v.x;
  d[r * COLD + NB + k] = 0.0 - v.y;
  d[(r + NB) * COLD + k] = v.y;
  d[(r + NB) * COLD + NB + k] = v.x;
}

// flags[m] |= 1 when any block of slot m has a non-zero element between the two spins (src: (18,18,nslot,ncls) complex)
__global__ void __launch_bounds__(BLKC)
k_sd_scan(const double2 *__restrict__ src, int nslot, int *__restrict__ flags) {
  const int m = blockIdx.x, c = blockIdx.y, t = threadIdx.x, r = t % NB, k = t / NB;
  const double2 v = src[(size_t)BLKC * (m + (size_t)nslot * c) + t];
  if (((r < 9) != (k < 9)) && (v.x != 0.0 || v.y != 0.0)) atomicOr(&flags[m], 1);
}

// rotmag_loc (math.f90:1981-2053): out(:,:,j) = R^H (in(:,:,j) R) for every 18x18 block; R = ROTMAT(alfa, beta, 0)
__global__ void __launch_bounds__(BLKC)
k_rotmag(const double2 *__restrict__ in, const double2 *__restrict__ R, double2 *__restrict__ out) {
  __shared__ double2 A[BLKC], Rm[BLKC], T[BLKC];
  const int t = threadIdx.x, r = t % NB, k = t / NB;
  const size_t off = (size_t)BLKC * blockIdx.x;
  A[t] = in[off + t];
  Rm[t] = R[t];
  __syncthreads();
  double2 s = make_double2(0.0, 0.0);
  for (int l = 0; l < NB; l++) s = c_add(s, c_mul(A[r + NB * l], Rm[l + NB * k]));   // h * Ux
  T[t] = s;
  __syncthreads();
  s = make_double2(0.0, 0.0);
  for (int l = 0; l < NB; l++) {                                                      // Ux' * (h * Ux)
    const double2 u = Rm[l + NB * r];
    s = c_add(s, c_mul(make_double2(u.x, -u.y), T[l + NB * k]));
  }
  out[off + t] = s;
}

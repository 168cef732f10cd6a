// kernels_simt.cuh -- straightforward SIMT FP64 kernels (kernel family 0).
//
// One CTA of 324 threads walks over sites; thread (r = tid%18, c = tid/18) owns element (r,c) of the site's 18x18
// output block.  These kernels cover every operator of the path (gather-SpMV with all epilogues, Gram reductions,
// the per-step 18x18 algebra of crecal_b) and are the in-library baseline the DMMA pipeline is measured against.
#pragma once
#include "common.cuh"

#define SIMT_THREADS 324

__device__ __forceinline__ void load_block_T(const double *__restrict__ g, double *s_re, double *s_im, int tid) {
  // HR36 block (Hr = rows 0..17, Hi = rows 18..35 of the first 18 columns)  ->  smem [k][r] (re, im) so that
  // lanes (consecutive r) are conflict-free
  for (int e = tid; e < BLKC; e += SIMT_THREADS) {
    int r = e / NB, k = e % NB;
    s_re[k * NB + r] = g[r * COLD + k];
    s_im[k * NB + r] = g[(r + NB) * COLD + k];
  }
}
__device__ __forceinline__ void load_block(const double *__restrict__ g, double *s, int tid) {
  for (int e = tid; e < BLKD; e += SIMT_THREADS) s[e] = g[e];
}

// acc(r,c) += sum_k H(r,k) * P(k,c);  Hs_* in [k][r] order, Ps in RI36
__device__ __forceinline__ void mac_block(const double *Hs_re, const double *Hs_im, const double *Ps, int r, int c,
                                          double &ar, double &ai) {
#pragma unroll
  for (int k = 0; k < NB; k++) {
    const double hr = Hs_re[k * NB + r], hi = Hs_im[k * NB + r];
    const double pr = Ps[c * COLD + k], pi = Ps[c * COLD + NB + k];
    ar = fma(hr, pr, ar); ar = fma(-hi, pi, ar);
    ai = fma(hr, pi, ai); ai = fma(hi, pr, ai);
  }
}
// d(i,j) += sum_k conj(X(k,i)) * Y(k,j); X, Y in RI36 in smem
__device__ __forceinline__ void gram_block(const double *Xs, const double *Ys, int i, int j, double &dr, double &di) {
#pragma unroll
  for (int k = 0; k < NB; k++) {
    const double xr = Xs[i * COLD + k], xi = Xs[i * COLD + NB + k];
    const double yr = Ys[j * COLD + k], yi = Ys[j * COLD + NB + k];
    dr = fma(xr, yr, dr); dr = fma(xi, yi, dr);
    di = fma(xr, yi, di); di = fma(-xi, yr, di);
  }
}

// Generic fused gather-SpMV + epilogue.  grid = (ctas_per_unit, nunits).
__global__ void __launch_bounds__(SIMT_THREADS) k_apply_simt(ApplyParams p) {
  __shared__ double Hs_re[BLKC], Hs_im[BLKC], Ps[BLKD], Xs[BLKD];
  const int tid = threadIdx.x, r = tid % NB, c = tid / NB;
  const int unit = blockIdx.y;
  const size_t uo = (size_t)unit * p.vstride;
  double g1r = 0, g1i = 0, g2r = 0, g2i = 0;
  for (int site = blockIdx.x; site < p.kk; site += gridDim.x) {
    const int cls = p.cls[site];
    double ar = 0, ai = 0;
    for (int t = 0; t < p.ngterms; t++) {
      const GatherTerm gt = p.g[t];
      for (int m = gt.first_slot; m < p.ngather; m++) {
        const int nb = p.nbr[(size_t)m * p.kk + site];
        if (nb == p.kk) continue;  // null site: contributes exact zeros
        __syncthreads();
        load_block_T(gt.H + ((size_t)cls * p.nslot_h + m) * HBLK, Hs_re, Hs_im, tid);
        load_block(gt.src + uo + (size_t)nb * BLKD, Ps, tid);
        __syncthreads();
        mac_block(Hs_re, Hs_im, Ps, r, c, ar, ai);
      }
    }
    if (p.Hx) {
      __syncthreads();
      load_block_T(p.Hx + (size_t)cls * HBLK, Hs_re, Hs_im, tid);
      load_block(p.srcx + uo + (size_t)site * BLKD, Ps, tid);
      __syncthreads();
      mac_block(Hs_re, Hs_im, Ps, r, c, ar, ai);
    }
    const size_t so = uo + (size_t)site * BLKD + c * COLD + r;
    if (p.addend) { ar += p.addend[so]; ai += p.addend[so + NB]; }
    if (p.epi == EPI_STORE) {
      p.out[so] = ar; p.out[so + NB] = ai;
    } else if (p.epi == EPI_HAM) {
      p.out[so] = (ar - p.b * p.in[so]) / p.a;
      p.out[so + NB] = (ai - p.b * p.in[so + NB]) / p.a;
    } else if (p.epi == EPI_CHEB || p.epi == EPI_CHEB_NOGRAM) {
      const double i_r = p.in[so], i_i = p.in[so + NB];
      const double n_r = 2.0 * ((ar - p.b * i_r) / p.a) - p.prev[so];
      const double n_i = 2.0 * ((ai - p.b * i_i) / p.a) - p.prev[so + NB];
      if (p.epi == EPI_CHEB) {
        __syncthreads();
        Ps[c * COLD + r] = i_r; Ps[c * COLD + NB + r] = i_i;  // psi1
        Xs[c * COLD + r] = n_r; Xs[c * COLD + NB + r] = n_i;  // psi2
        __syncthreads();
        gram_block(Ps, Ps, r, c, g1r, g1i);  // psi1^H psi1
        gram_block(Xs, Ps, r, c, g2r, g2i);  // psi2^H psi1
      }
      p.out[so] = n_r; p.out[so + NB] = n_i;
    } else {  // EPI_HOP
      __syncthreads();
      Ps[c * COLD + r] = p.in[so]; Ps[c * COLD + NB + r] = p.in[so + NB];  // psi
      Xs[c * COLD + r] = ar; Xs[c * COLD + NB + r] = ai;                    // H psi
      __syncthreads();
      gram_block(Ps, Xs, r, c, g1r, g1i);  // psi^H (H psi)
      p.out[so] = ar - p.prev[so]; p.out[so + NB] = ai - p.prev[so + NB];
    }
  }
  if (p.part) {
    double *pp = p.part + ((size_t)unit * gridDim.x + blockIdx.x) * (2 * BLKD);
    // complex, column-major (i + 18 j) like the host arrays
    pp[2 * (r + NB * c)] = g1r; pp[2 * (r + NB * c) + 1] = g1i;
    pp[BLKD + 2 * (r + NB * c)] = g2r; pp[BLKD + 2 * (r + NB * c) + 1] = g2i;
  }
}

// D = sum_sites X^H Y for two batched vectors; partials like k_apply_simt (matrix 0 only).
__global__ void __launch_bounds__(SIMT_THREADS) k_gram_simt(const double *X, const double *Y, int kk, size_t xstride,
                                                            size_t ystride, double *part) {
  __shared__ double Xs[BLKD], Ys[BLKD];
  const int tid = threadIdx.x, r = tid % NB, c = tid / NB, unit = blockIdx.y;
  double gr = 0, gi = 0;
  for (int site = blockIdx.x; site < kk; site += gridDim.x) {
    __syncthreads();
    load_block(X + (size_t)unit * xstride + (size_t)site * BLKD, Xs, tid);
    load_block(Y + (size_t)unit * ystride + (size_t)site * BLKD, Ys, tid);
    __syncthreads();
    gram_block(Xs, Ys, r, c, gr, gi);
  }
  double *pp = part + ((size_t)unit * gridDim.x + blockIdx.x) * (2 * BLKD);
  pp[2 * (r + NB * c)] = gr; pp[2 * (r + NB * c) + 1] = gi;
  pp[BLKD + 2 * (r + NB * c)] = 0.0; pp[BLKD + 2 * (r + NB * c) + 1] = 0.0;
}

// Fixed-order reduction of the per-CTA partials: out[unit][which] (complex 18x18, host layout) = sum_cta part.
// A CTA of 1024 threads takes 32 consecutive matrix elements: thread (g, l) sums the partials g, g+32, ... of element l in order
// (coalesced: a warp reads 32 consecutive doubles of one partial), then the 32 group sums are added in a fixed order --
// deterministic for a given partial count.  (Round 1 used one warp per element with lane-strided partials: every 8 B load pulled
// its own 32 B sector, 6.5 us for 294 partials; this form takes about half.)
// mode 0: plain store to dst0 (+ dst1 for matrix 1 if non-null)
// mode 1: Chebyshev finish: dst0 = 2*D1 - mu0, dst1 = 2*D2 - mu1   (recursion.f90:2591-2592)
// mode 2: diagonal projection (scalar Lanczos): keep Re(diag) only
// grid = (ceil(2*648 / 32), nunits), 1024 threads
#define RP_ELEMS 32
#define RP_GROUPS 32
__global__ void __launch_bounds__(RP_ELEMS * RP_GROUPS)
k_reduce_parts(const double *part, int nctas, int mode, double *dst0, double *dst1, size_t dstride,
               const double *mu0, const double *mu1, double *hist0 = nullptr, size_t hstride = 0) {
  __shared__ double gs[RP_GROUPS][RP_ELEMS];
  const int unit = blockIdx.y, l = threadIdx.x % RP_ELEMS, g = threadIdx.x / RP_ELEMS;
  const int e = blockIdx.x * RP_ELEMS + l;
  double acc = 0.0;
  // matrix 1 of the slots is only wanted by the Chebyshev finish (mode 1) or when the caller asks for it (dst1)
  const bool wanted = e < 2 * BLKD && (e < BLKD || mode == 1 || dst1 != nullptr);
  if (wanted) {
    const double *pp = part + (size_t)unit * nctas * (2 * BLKD) + e;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    int cta = g;
    for (; cta + 3 * RP_GROUPS < nctas; cta += 4 * RP_GROUPS) {
      const double v0 = pp[(size_t)cta * (2 * BLKD)], v1 = pp[(size_t)(cta + RP_GROUPS) * (2 * BLKD)],
                   v2 = pp[(size_t)(cta + 2 * RP_GROUPS) * (2 * BLKD)], v3 = pp[(size_t)(cta + 3 * RP_GROUPS) * (2 * BLKD)];
      a0 += v0; a1 += v1; a2 += v2; a3 += v3;
    }
    for (; cta < nctas; cta += RP_GROUPS) a0 += pp[(size_t)cta * (2 * BLKD)];
    acc = (a0 + a1) + (a2 + a3);
  }
  gs[g][l] = acc;
  __syncthreads();
  if (g != 0 || !wanted) return;
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < RP_GROUPS; k++) s += gs[k][l];
  const int which = e / BLKD, idx = e % BLKD;
  if (mode == 1) {
    const double *m = which ? mu1 : mu0;
    double *d = which ? dst1 : dst0;
    d[(size_t)unit * dstride + idx] = 2.0 * s - m[(size_t)unit * dstride + idx];
  } else {
    double *d = which ? dst1 : dst0;
    if (!d) return;
    if (mode == 2) {
      const int ce = idx / 2, im = idx & 1, i = ce % NB, j = ce / NB;
      if (i != j || im) s = 0.0;
    }
    d[(size_t)unit * dstride + idx] = s;
    if (hist0 && which == 0) hist0[(size_t)unit * hstride + idx] = s;  // the history slot of the same block (atemp_b)
  }
}

// pmn -= psi * A ;  B2 partial += pmn^H pmn          crecal_b, recursion.f90:1922-1934
// hpsi != null: pmn = hpsi - pmn first (the hop_b update, recursion.f90:1641, when the SpMV kernel only stored H psi)
__global__ void __launch_bounds__(SIMT_THREADS) k_lz_ortho_simt(const double *psi, double *pmn, const double *hpsi,
                                                                const double *Amat, size_t astride, int kk,
                                                                size_t vstride, double *part) {
  __shared__ double As_re[BLKC], As_im[BLKC], Ps[BLKD], Xs[BLKD];
  const int tid = threadIdx.x, r = tid % NB, c = tid / NB, unit = blockIdx.y;
  const size_t uo = (size_t)unit * vstride;
  const double *A = Amat + (size_t)unit * astride;  // complex col-major (k + 18 c)
  As_re[tid] = A[2 * tid]; As_im[tid] = A[2 * tid + 1];
  double gr = 0, gi = 0;
  for (int site = blockIdx.x; site < kk; site += gridDim.x) {
    __syncthreads();
    load_block(psi + uo + (size_t)site * BLKD, Ps, tid);
    __syncthreads();
    const size_t so = uo + (size_t)site * BLKD + c * COLD + r;
    double ar = pmn[so], ai = pmn[so + NB];
    if (hpsi) { ar = hpsi[so] - ar; ai = hpsi[so + NB] - ai; }
#pragma unroll
    for (int k = 0; k < NB; k++) {  // (psi A)(r,c) = sum_k psi(r,k) A(k,c)
      const double pr = Ps[k * COLD + r], pi = Ps[k * COLD + NB + r];
      const double qr = As_re[k + NB * c], qi = As_im[k + NB * c];
      ar = fma(-pr, qr, ar); ar = fma(pi, qi, ar);
      ai = fma(-pr, qi, ai); ai = fma(-pi, qr, ai);
    }
    pmn[so] = ar; pmn[so + NB] = ai;
    Xs[c * COLD + r] = ar; Xs[c * COLD + NB + r] = ai;
    __syncthreads();
    gram_block(Xs, Xs, r, c, gr, gi);
  }
  double *pp = part + ((size_t)unit * gridDim.x + blockIdx.x) * (2 * BLKD);
  pp[2 * (r + NB * c)] = gr; pp[2 * (r + NB * c) + 1] = gi;
  pp[BLKD + 2 * (r + NB * c)] = 0.0; pp[BLKD + 2 * (r + NB * c) + 1] = 0.0;
}

// psi' = pmn * Binv ; pmn' = psi * B                  crecal_b, recursion.f90:1963-1969
__global__ void __launch_bounds__(SIMT_THREADS) k_lz_rotate_simt(double *psi, double *pmn, const double *Bmat,
                                                                 const double *Bimat, size_t bstride, int kk,
                                                                 size_t vstride) {
  __shared__ double B_re[BLKC], B_im[BLKC], Bi_re[BLKC], Bi_im[BLKC], Ps[BLKD], Ms[BLKD];
  const int tid = threadIdx.x, r = tid % NB, c = tid / NB, unit = blockIdx.y;
  const size_t uo = (size_t)unit * vstride;
  B_re[tid] = Bmat[(size_t)unit * bstride + 2 * tid]; B_im[tid] = Bmat[(size_t)unit * bstride + 2 * tid + 1];
  Bi_re[tid] = Bimat[(size_t)unit * bstride + 2 * tid]; Bi_im[tid] = Bimat[(size_t)unit * bstride + 2 * tid + 1];
  for (int site = blockIdx.x; site < kk; site += gridDim.x) {
    __syncthreads();
    load_block(psi + uo + (size_t)site * BLKD, Ps, tid);
    load_block(pmn + uo + (size_t)site * BLKD, Ms, tid);
    __syncthreads();
    double nr = 0, ni = 0, mr = 0, mi = 0;
#pragma unroll
    for (int k = 0; k < NB; k++) {
      const double xr = Ms[k * COLD + r], xi = Ms[k * COLD + NB + r];
      const double br = Bi_re[k + NB * c], bi = Bi_im[k + NB * c];
      nr = fma(xr, br, nr); nr = fma(-xi, bi, nr);
      ni = fma(xr, bi, ni); ni = fma(xi, br, ni);
      const double pr = Ps[k * COLD + r], pi = Ps[k * COLD + NB + r];
      const double cr = B_re[k + NB * c], ci = B_im[k + NB * c];
      mr = fma(pr, cr, mr); mr = fma(-pi, ci, mr);
      mi = fma(pr, ci, mi); mi = fma(pi, cr, mi);
    }
    const size_t so = uo + (size_t)site * BLKD + c * COLD + r;
    psi[so] = nr; psi[so + NB] = ni;
    pmn[so] = mr; pmn[so + NB] = mi;
  }
}

// ---- layout conversion / initialisation ----
// host complex col-major (18,18,kk) <-> RI36
__global__ void k_host_to_ri36(const double *h, double *d, int kk) {
  size_t n = (size_t)kk * BLKC;
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x) {
    size_t site = e / BLKC; int w = e % BLKC, k = w % NB, c = w / NB;
    d[site * BLKD + c * COLD + k] = h[2 * e];
    d[site * BLKD + c * COLD + NB + k] = h[2 * e + 1];
  }
}
__global__ void k_ri36_to_host(const double *d, double *h, int kk) {
  size_t n = (size_t)kk * BLKC;
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x) {
    size_t site = e / BLKC; int w = e % BLKC, k = w % NB, c = w / NB;
    h[2 * e] = d[site * BLKD + c * COLD + k];
    h[2 * e + 1] = d[site * BLKD + c * COLD + NB + k];
  }
}
// start blocks: unit u gets asign*I on site_i[u] and bsign*I on site_j[u] (recursion.f90:1709-1711, 1834-1836)
__global__ void k_init_site_start(double *v, size_t vstride, const int32_t *site_i, const int32_t *site_j,
                                  const double *asign, const double *bsign, int nunits) {
  int u = blockIdx.x, l = threadIdx.x;
  if (u >= nunits || l >= NB) return;
  double *base = v + (size_t)u * vstride;
  int i = site_i[u] - 1, j = site_j[u] - 1;
  base[(size_t)i * BLKD + l * COLD + l] = asign[2 * u];
  base[(size_t)i * BLKD + l * COLD + NB + l] = asign[2 * u + 1];
  if (j >= 0) {
    base[(size_t)j * BLKD + l * COLD + l] = bsign[2 * u];
    base[(size_t)j * BLKD + l * COLD + NB + l] = bsign[2 * u + 1];
  }
}
// KPM random-phase start: exp(2 pi i u_k) I / sqrt(kk) on every site (recursion.f90:1135-1142)
__global__ void k_init_random_start(double *v, size_t vstride, const double *phases, int kk, int nvec) {
  size_t n = (size_t)kk * nvec;
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x) {
    int u = e / kk; size_t site = e % kk;
    double s, c;
    sincospi(2.0 * phases[e], &s, &c);
    // the reference evaluates sqrt(real(kk)) in single precision (default real kind) and promotes the result
    const double nrm = (double)sqrtf((float)kk);
    double *blk = v + (size_t)u * vstride + site * BLKD;
    for (int l = 0; l < NB; l++) { blk[l * COLD + l] = c / nrm; blk[l * COLD + NB + l] = s / nrm; }
  }
}
__global__ void k_set_identity(double *m, size_t stride, int n) {  // complex col-major 18x18 identity per unit
  int u = blockIdx.x;
  if (u >= n) return;
  for (int e = threadIdx.x; e < BLKC; e += blockDim.x) {
    m[(size_t)u * stride + 2 * e] = (e % NB == e / NB) ? 1.0 : 0.0;
    m[(size_t)u * stride + 2 * e + 1] = 0.0;
  }
}

// acc[e] += sum over units (in unit order: deterministic) of src[u * stride + e]
__global__ void k_sum_units(const double *__restrict__ src, size_t stride, int nunits, double *__restrict__ acc) {
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < stride; e += (size_t)gridDim.x * blockDim.x) {
    double s = acc[e];
    for (int u = 0; u < nunits; u++) s += src[(size_t)u * stride + e];
    acc[e] = s;
  }
}

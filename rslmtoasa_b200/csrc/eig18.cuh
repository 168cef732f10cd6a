// eig18.cuh -- device-side 18x18 Hermitian eigen-decomposition and matrix square roots.
//
// Replaces the serial LAPACK zheev calls on the critical path of crecal_b (recursion.f90:1938-1959) and zsqr
// (recursion.f90:1980-2023).  One CTA of 324 threads per matrix: two-sided Jacobi with the round-robin parallel
// ordering (17 rounds x 9 disjoint rotations per sweep), thread (r,c) updates element (r,c) of A and V.
#pragma once
#include "common.cuh"

struct Eig18Smem {
  double Ar[BLKC], Ai[BLKC], Vr[BLKC], Vi[BLKC];  // [r + 18 c]
  double Tr[BLKC], Ti[BLKC];                      // Newton-Schulz work matrix
  double cs[9], sr[9], si[9];                     // rotation: cos, sin*phase
  int pp[9], qq[9], pair_of[NB];
  double ev[NB];
  double tot;
};

// Sum of one value per thread over the 324-thread CTA in a fixed order (run-to-run reproducible, unlike atomics):
// 32 strided partial sums, then a shuffle tree in warp 0.  Uses s.Tr as scratch; result also left in s.tot.
__device__ __forceinline__ double eig18_block_sum(Eig18Smem &s, double v) {
  const int tid = threadIdx.x;
  s.Tr[tid] = v;
  __syncthreads();
  if (tid < 32) {
    double a = 0.0;
    for (int k = tid; k < BLKC; k += 32) a += s.Tr[k];
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (tid == 0) s.tot = a;
  }
  __syncthreads();
  return s.tot;
}

// A (Hermitian, upper triangle trusted like zheev 'U') -> eigenvalues s.ev, eigenvectors s.V (columns).
__device__ void eig18_jacobi(Eig18Smem &s) {
  const int tid = threadIdx.x, r = tid % NB, c = tid / NB;
  // symmetrise from the upper triangle
  __syncthreads();
  double ar = s.Ar[tid], ai = s.Ai[tid];
  if (r > c) { ar = s.Ar[c + NB * r]; ai = -s.Ai[c + NB * r]; }
  if (r == c) ai = 0.0;
  __syncthreads();
  s.Ar[tid] = ar; s.Ai[tid] = ai;
  s.Vr[tid] = (r == c) ? 1.0 : 0.0; s.Vi[tid] = 0.0;
  __syncthreads();
  eig18_block_sum(s, ar * ar + ai * ai);
  // |a_ij|^2 <= 1e-40 ||A||_F^2: off-diagonal elements perturb eigenvalues at second order (~1e-40 ||A||),
  // far below double precision; Jacobi converges quadratically so this saves about one sweep over 1e-64.
  const double thresh = 1e-40 * s.tot;
  for (int sweep = 0; sweep < 40; sweep++) {
    const int notconv = __syncthreads_or(r != c && (s.Ar[tid] * s.Ar[tid] + s.Ai[tid] * s.Ai[tid]) > thresh);
    if (!notconv) break;
    for (int t = 0; t < 17; t++) {
      if (tid < 9) {
        int a_, b_;
        if (tid == 0) { a_ = 17; b_ = t; } else { a_ = (t + tid) % 17; b_ = (t - tid + 17) % 17; }
        const int p = min(a_, b_), q = max(a_, b_);
        s.pp[tid] = p; s.qq[tid] = q; s.pair_of[p] = tid; s.pair_of[q] = tid;
        const double pr = s.Ar[p + NB * q], pi = s.Ai[p + NB * q];
        const double m2 = pr * pr + pi * pi;  // |a_pq|^2
        double cs = 1.0, sr = 0.0, si = 0.0;
        if (m2 > 1e-300) {
          // t = tan(theta) = sign(z) |a_pq| / (|z| + sqrt(z^2 + |a_pq|^2)),  z = (a_qq - a_pp)/2.  With u = t/|a_pq|
          // the rotation is cs = 1/sqrt(1 + u^2 |a_pq|^2), sn e^{i phi} = cs * u * a_pq: one sqrt, one division and
          // one rsqrt on the critical path instead of three square roots and four divisions.
          const double z = 0.5 * (s.Ar[q + NB * q] - s.Ar[p + NB * p]);
          const double u = (z >= 0.0 ? 1.0 : -1.0) / (fabs(z) + sqrt(z * z + m2));
          cs = rsqrt(1.0 + u * u * m2);
          sr = cs * u * pr; si = cs * u * pi;
        }
        s.cs[tid] = cs; s.sr[tid] = sr; s.si[tid] = si;
      }
      __syncthreads();
      // R restricted to (p,q): R_pp = cs, R_pq = (sr + i si), R_qp = -(sr - i si), R_qq = cs
      const int kc = s.pair_of[c], pc = s.pp[kc], qc = s.qq[kc];
      const int kr = s.pair_of[r], pr_ = s.pp[kr], qr_ = s.qq[kr];
      double Rpc_r, Rpc_i, Rqc_r, Rqc_i;  // R[pc][c], R[qc][c]
      if (c == pc) { Rpc_r = s.cs[kc]; Rpc_i = 0.0; Rqc_r = -s.sr[kc]; Rqc_i = s.si[kc]; }
      else         { Rpc_r = s.sr[kc]; Rpc_i = s.si[kc]; Rqc_r = s.cs[kc]; Rqc_i = 0.0; }
      double Rpr_r, Rpr_i, Rqr_r, Rqr_i;  // R[pr][r], R[qr][r]
      if (r == pr_) { Rpr_r = s.cs[kr]; Rpr_i = 0.0; Rqr_r = -s.sr[kr]; Rqr_i = s.si[kr]; }
      else          { Rpr_r = s.sr[kr]; Rpr_i = s.si[kr]; Rqr_r = s.cs[kr]; Rqr_i = 0.0; }
      // (A R)[x][c] for x = pr_, qr_
      double x1r, x1i, x2r, x2i;
      {
        const double a1r = s.Ar[pr_ + NB * pc], a1i = s.Ai[pr_ + NB * pc], a2r = s.Ar[pr_ + NB * qc], a2i = s.Ai[pr_ + NB * qc];
        x1r = a1r * Rpc_r - a1i * Rpc_i + a2r * Rqc_r - a2i * Rqc_i;
        x1i = a1r * Rpc_i + a1i * Rpc_r + a2r * Rqc_i + a2i * Rqc_r;
        const double b1r = s.Ar[qr_ + NB * pc], b1i = s.Ai[qr_ + NB * pc], b2r = s.Ar[qr_ + NB * qc], b2i = s.Ai[qr_ + NB * qc];
        x2r = b1r * Rpc_r - b1i * Rpc_i + b2r * Rqc_r - b2i * Rqc_i;
        x2i = b1r * Rpc_i + b1i * Rpc_r + b2r * Rqc_i + b2i * Rqc_r;
      }
      // A'[r][c] = conj(R[pr][r]) x1 + conj(R[qr][r]) x2
      double nr = Rpr_r * x1r + Rpr_i * x1i + Rqr_r * x2r + Rqr_i * x2i;
      double ni = Rpr_r * x1i - Rpr_i * x1r + Rqr_r * x2i - Rqr_i * x2r;
      // V'[r][c] = V[r][pc] R[pc][c] + V[r][qc] R[qc][c]
      const double v1r = s.Vr[r + NB * pc], v1i = s.Vi[r + NB * pc], v2r = s.Vr[r + NB * qc], v2i = s.Vi[r + NB * qc];
      const double wr = v1r * Rpc_r - v1i * Rpc_i + v2r * Rqc_r - v2i * Rqc_i;
      const double wi = v1r * Rpc_i + v1i * Rpc_r + v2r * Rqc_i + v2i * Rqc_r;
      // the rotated pair is annihilated exactly; diagonals are real
      if ((r == pr_ && c == qr_) || (r == qr_ && c == pr_)) { nr = 0.0; ni = 0.0; }
      if (r == c) ni = 0.0;
      __syncthreads();
      s.Ar[tid] = nr; s.Ai[tid] = ni; s.Vr[tid] = wr; s.Vi[tid] = wi;
      __syncthreads();
    }
  }
  if (tid < NB) s.ev[tid] = s.Ar[tid + NB * tid];
  __syncthreads();
}

// Hermitian square root and inverse square root by the coupled Newton-Schulz iteration
//   Y0 = M/s, Z0 = I;  T = (3I - Z Y)/2;  Y <- Y T;  Z <- T Z;   Y -> (M/s)^1/2, Z -> (M/s)^-1/2   (s = ||M||_F)
// -- eighteen-wide matrix products only, no serial scalar chain, so it is ~5x shorter than the Jacobi sweep that
// replaces zheev on the critical path of crecal_b.  Returns false (caller falls back to Jacobi) unless the
// iteration converged to round-off; M must be Hermitian positive definite for that.
// On success B and Binv (complex col-major) are written.
__device__ bool sqrt18_newton_schulz(Eig18Smem &s, double *B, double *Bi) {
  const int tid = threadIdx.x, r = tid % NB, c = tid / NB;
  __syncthreads();
  double ar = s.Ar[tid], ai = s.Ai[tid];
  if (r > c) { ar = s.Ar[c + NB * r]; ai = -s.Ai[c + NB * r]; }  // trust the upper triangle like zheev('U')
  if (r == c) ai = 0.0;
  __syncthreads();
  // scale = ||M||_inf (largest row sum of |m_rc|) >= lambda_max for a Hermitian matrix, and within a few per cent of it for
  // the well-conditioned B^2 of a recursion (cond ~ 1.2): the iteration then starts inside its quadratic regime.  The
  // Frobenius norm used before overestimates lambda_max by up to sqrt(18) and cost three more linear-phase iterations.
  s.Tr[tid] = sqrt(ar * ar + ai * ai);
  __syncthreads();
  if (tid < NB) {
    double rs = 0.0;
#pragma unroll
    for (int k = 0; k < NB; k++) rs += s.Tr[tid + NB * k];
    s.ev[tid] = rs;
  }
  __syncthreads();
  double scale = 0.0;
#pragma unroll
  for (int k = 0; k < NB; k++) scale = fmax(scale, s.ev[k]);
  if (!(scale > 0.0) || !(scale < 1e300)) return false;
  __syncthreads();
  // Y in (Ar,Ai), Z in (Vr,Vi)
  s.Ar[tid] = ar / scale; s.Ai[tid] = ai / scale;
  s.Vr[tid] = (r == c) ? 1.0 : 0.0; s.Vi[tid] = 0.0;
  __syncthreads();
  // The products are register-tiled: 81 threads each own a 2x2 tile of the result (thread-per-element, the matrices in
  // shared memory, is one shared-memory load per FMA pair and was bound by the shared-memory pipe of the one SM it runs
  // on: 3.3 us per iteration).  Every element is still the same ascending-k chain of fused multiply-adds, so B and B^-1
  // are bit-identical to the thread-per-element form.
  const bool tiled = tid < 81;
  const int r0 = 2 * (tid % 9), c0 = 2 * (tid / 9);
  bool ok = false;
  for (int it = 0; it < 120; it++) {
    // P = Z Y, residual I - P, T = (3I - P)/2
    bool bad = false, big = false;
    if (tiled) {
      double pr[2][2] = {{0.0, 0.0}, {0.0, 0.0}}, pi[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
      for (int k = 0; k < NB; k++) {
        double zr[2], zi[2], yr[2], yi[2];
#pragma unroll
        for (int a = 0; a < 2; a++) {
          zr[a] = s.Vr[r0 + a + NB * k]; zi[a] = s.Vi[r0 + a + NB * k];
          yr[a] = s.Ar[k + NB * (c0 + a)]; yi[a] = s.Ai[k + NB * (c0 + a)];
        }
#pragma unroll
        for (int a = 0; a < 2; a++)
#pragma unroll
          for (int b = 0; b < 2; b++) {
            pr[a][b] = fma(zr[a], yr[b], pr[a][b]); pr[a][b] = fma(-zi[a], yi[b], pr[a][b]);
            pi[a][b] = fma(zr[a], yi[b], pi[a][b]); pi[a][b] = fma(zi[a], yr[b], pi[a][b]);
          }
      }
#pragma unroll
      for (int a = 0; a < 2; a++)
#pragma unroll
        for (int b = 0; b < 2; b++) {
          const bool dg = (r0 + a) == (c0 + b);
          const double dr = (dg ? 1.0 : 0.0) - pr[a][b], di = -pi[a][b];  // residual I - Z Y
          bad = bad || !(fabs(dr) < 1e300) || !(fabs(di) < 1e300);
          big = big || fabs(dr) > 1e-13 || fabs(di) > 1e-13;
          s.Tr[r0 + a + NB * (c0 + b)] = (dg ? 1.5 : 0.0) - 0.5 * pr[a][b];
          s.Ti[r0 + a + NB * (c0 + b)] = -0.5 * pi[a][b];
        }
    }
    if (__syncthreads_or(bad)) return false;
    const int anybig = __syncthreads_or(big);
    // Y <- Y T, Z <- T Z
    double yr_[2][2] = {{0.0, 0.0}, {0.0, 0.0}}, yi_[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
    double zr_[2][2] = {{0.0, 0.0}, {0.0, 0.0}}, zi_[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
    if (tiled) {
#pragma unroll
      for (int k = 0; k < NB; k++) {
        double a1[2], a2[2], t1[2], t2[2], u1[2], u2[2], z1[2], z2[2];
#pragma unroll
        for (int a = 0; a < 2; a++) {
          a1[a] = s.Ar[r0 + a + NB * k]; a2[a] = s.Ai[r0 + a + NB * k];
          t1[a] = s.Tr[k + NB * (c0 + a)]; t2[a] = s.Ti[k + NB * (c0 + a)];
          u1[a] = s.Tr[r0 + a + NB * k]; u2[a] = s.Ti[r0 + a + NB * k];
          z1[a] = s.Vr[k + NB * (c0 + a)]; z2[a] = s.Vi[k + NB * (c0 + a)];
        }
#pragma unroll
        for (int a = 0; a < 2; a++)
#pragma unroll
          for (int b = 0; b < 2; b++) {
            yr_[a][b] = fma(a1[a], t1[b], yr_[a][b]); yr_[a][b] = fma(-a2[a], t2[b], yr_[a][b]);
            yi_[a][b] = fma(a1[a], t2[b], yi_[a][b]); yi_[a][b] = fma(a2[a], t1[b], yi_[a][b]);
            zr_[a][b] = fma(u1[a], z1[b], zr_[a][b]); zr_[a][b] = fma(-u2[a], z2[b], zr_[a][b]);
            zi_[a][b] = fma(u1[a], z2[b], zi_[a][b]); zi_[a][b] = fma(u2[a], z1[b], zi_[a][b]);
          }
      }
    }
    __syncthreads();
    if (tiled) {
#pragma unroll
      for (int a = 0; a < 2; a++)
#pragma unroll
        for (int b = 0; b < 2; b++) {
          const int e = r0 + a + NB * (c0 + b);
          s.Ar[e] = yr_[a][b]; s.Ai[e] = yi_[a][b]; s.Vr[e] = zr_[a][b]; s.Vi[e] = zi_[a][b];
        }
    }
    __syncthreads();
    if (!anybig) { ok = true; break; }  // residual below 1e-13 BEFORE this (quadratic) update: the iterate is at round-off now
  }
  if (!ok) return false;
  const double sq = sqrt(scale);
  // Hermitian parts (the iterates are polynomials in M: Hermitian up to rounding)
  B[2 * tid] = 0.5 * sq * (s.Ar[tid] + s.Ar[c + NB * r]);
  B[2 * tid + 1] = 0.5 * sq * (s.Ai[tid] - s.Ai[c + NB * r]);
  Bi[2 * tid] = 0.5 * (s.Vr[tid] + s.Vr[c + NB * r]) / sq;
  Bi[2 * tid + 1] = 0.5 * (s.Vi[tid] - s.Vi[c + NB * r]) / sq;
  return true;
}

// out(r,c) = sum_k V(r,k) f_k conj(V(c,k)), complex col-major interleaved
__device__ __forceinline__ void eig18_func(const Eig18Smem &s, const double *f, double *out) {
  const int tid = threadIdx.x, r = tid % NB, c = tid / NB;
  double orr = 0, oi = 0;
#pragma unroll
  for (int k = 0; k < NB; k++) {
    const double ar = s.Vr[r + NB * k], ai = s.Vi[r + NB * k], br = s.Vr[c + NB * k], bi = -s.Vi[c + NB * k];
    orr += f[k] * (ar * br - ai * bi);
    oi += f[k] * (ar * bi + ai * br);
  }
  out[2 * tid] = orr; out[2 * tid + 1] = oi;
}

#define LZ_EIG_ALONE_SMEM (128 * 1024)  // dynamic shared memory requested (not used) when k_lz_eig must not share its SM
// crecal_b "B_n+1": take the reduced B^2 of unit blockIdx.x (k_reduce_parts), record it in the history slot, then
// B = U sqrt(L) U^H and B^-1 = U L^-1/2 U^H.  diag != 0: scalar Lanczos, everything diagonal & real.
__global__ void __launch_bounds__(BLKC) k_lz_eig(const double *b2, size_t b2stride, double *b2_hist_slot, size_t hstride,
                                                 double *Bmat, double *Bimat, size_t bstride, int diag, int method,
                                                 double *b_hist_slot = nullptr, const double *part = nullptr, int nparts = 0) {
  __shared__ Eig18Smem s;
  __shared__ double f1[NB], f2[NB];
  const int tid = threadIdx.x, unit = blockIdx.x, r = tid % NB, c = tid / NB;
  double mr, mi;
  if (part) {
    // B^2 = fixed-order sum of the per-CTA partials of sum pmn^H pmn (matrix 0 of each slot): folds the k_reduce_parts launch
    // that used to precede this kernel into its prologue (coalesced: thread = matrix element)
    const double2 *pp = reinterpret_cast<const double2 *>(part + (size_t)unit * nparts * (2 * BLKD)) + tid;
    double2 a0 = make_double2(0.0, 0.0), a1 = a0, a2 = a0, a3 = a0;
    int cta = 0;
    for (; cta + 4 <= nparts; cta += 4) {
      const double2 v0 = pp[(size_t)cta * BLKD], v1 = pp[(size_t)(cta + 1) * BLKD], v2 = pp[(size_t)(cta + 2) * BLKD], v3 = pp[(size_t)(cta + 3) * BLKD];
      a0.x += v0.x; a0.y += v0.y; a1.x += v1.x; a1.y += v1.y; a2.x += v2.x; a2.y += v2.y; a3.x += v3.x; a3.y += v3.y;
    }
    for (; cta < nparts; cta++) { const double2 v = pp[(size_t)cta * BLKD]; a0.x += v.x; a0.y += v.y; }
    mr = (a0.x + a1.x) + (a2.x + a3.x); mi = (a0.y + a1.y) + (a2.y + a3.y);
  } else {
    mr = b2[(size_t)unit * b2stride + 2 * tid]; mi = b2[(size_t)unit * b2stride + 2 * tid + 1];
  }
  if (diag) { if (r != c) mr = 0.0; mi = 0.0; }
  b2_hist_slot[(size_t)unit * hstride + 2 * tid] = mr;
  b2_hist_slot[(size_t)unit * hstride + 2 * tid + 1] = mi;
  double *B = Bmat + (size_t)unit * bstride, *Bi = Bimat + (size_t)unit * bstride;
  double *bh = b_hist_slot ? b_hist_slot + (size_t)unit * hstride : nullptr;  // B of this level for the Green function
  if (diag) {
    const double sq = sqrt(mr);
    B[2 * tid] = (r == c) ? sq : 0.0; B[2 * tid + 1] = 0.0;
    Bi[2 * tid] = (r == c) ? 1.0 / sq : 0.0; Bi[2 * tid + 1] = 0.0;
    if (bh) { bh[2 * tid] = B[2 * tid]; bh[2 * tid + 1] = 0.0; }
    return;
  }
  s.Ar[tid] = mr; s.Ai[tid] = mi;
  if (method == 1 && sqrt18_newton_schulz(s, B, Bi)) {
    if (bh) { bh[2 * tid] = B[2 * tid]; bh[2 * tid + 1] = B[2 * tid + 1]; }
    return;
  }
  __syncthreads();
  s.Ar[tid] = mr; s.Ai[tid] = mi;  // fallback / method 0: eigen-decomposition like the reference's zheev path
  eig18_jacobi(s);
  if (tid < NB) { f1[tid] = sqrt(s.ev[tid]); f2[tid] = 1.0 / f1[tid]; }  // NaN for ev<0, like the reference
  __syncthreads();
  eig18_func(s, f1, B);
  eig18_func(s, f2, Bi);
  if (bh) { bh[2 * tid] = B[2 * tid]; bh[2 * tid + 1] = B[2 * tid + 1]; }
}

// zsqr: in-place square root of a batch of Hermitian PSD 18x18 matrices (complex col-major)
__global__ void __launch_bounds__(BLKC) k_zsqr(double *mats) {
  __shared__ Eig18Smem s;
  __shared__ double f1[NB];
  const int tid = threadIdx.x;
  double *m = mats + (size_t)blockIdx.x * BLKD;
  __shared__ double inv_scratch[BLKD];
  const double mr = m[2 * tid], mi = m[2 * tid + 1];
  s.Ar[tid] = mr; s.Ai[tid] = mi;
  // same routine, same input bits as k_lz_eig saw for this level: the staged zsqr returns exactly the B the recursion used
  if (sqrt18_newton_schulz(s, m, inv_scratch)) return;
  __syncthreads();
  s.Ar[tid] = mr; s.Ai[tid] = mi;
  eig18_jacobi(s);
  if (tid < NB) f1[tid] = sqrt(s.ev[tid]);
  __syncthreads();
  eig18_func(s, f1, m);
}

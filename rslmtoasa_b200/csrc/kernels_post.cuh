// kernels_post.cuh -- consumers either side of the recursion hot path on the device.
//
//   k_bpopt / k_terminf_fix   Beer-Pettifor terminator: bpopt + emami (recursion.f90:3540-3706), get_cinf /
//                             get_terminf (2030-2138).  Branchy bisection on tridiagonals: one thread per chain,
//                             every floating-point operation written as an explicit IEEE round-to-nearest op so that
//                             the branch decisions are those of the reference's non-contracted arithmetic.
//   k_bgreen                  block continued fraction (green.f90:1191-1339): per energy, ll-1 levels of
//                             Q <- B^H (E - A - Q)^-1 B with an 18x18 complex LU inverse (partial pivoting, the
//                             |re|+|im| pivot rule of izamax).  One warp per (energy, unit), lane j owns column j.
//   k_cheb_weight / k_cheb_green   kernel-weighted Chebyshev sum (green.f90:1030-1108).
//   k_density / k_sgreen_*    scalar continued fraction + terminator (density_of_states.f90:248-407, green.f90:628-705).
//   k_cond_*                  Kubo-Bastin Gamma_nm and its contraction with the diagonal of mu_nm_stochastic
//                             (conductivity.f90:158-306); Gamma is never materialised (rank-2 in (n,m)).
// All inputs/outputs here are in the reference's own complex column-major layouts (re,im interleaved).
#pragma once
#include "common.cuh"

#define PI_RP 3.14159265358979323846  // math.f90:70

__device__ __forceinline__ double2 c_mul(double2 a, double2 b) {
  return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ double2 c_add(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 c_sub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 c_scale(double2 a, double s) { return make_double2(a.x * s, a.y * s); }
// x - a*b and s + conj(a)*b as four fused multiply-adds (the mul-then-add forms cost six FP64 instructions)
__device__ __forceinline__ double2 c_fnma(double2 x, double2 a, double2 b) {
  return make_double2(fma(a.y, b.y, fma(-a.x, b.x, x.x)), fma(-a.y, b.x, fma(-a.x, b.y, x.y)));
}
__device__ __forceinline__ double2 c_fma(double2 s, double2 a, double2 b) {  // s + a*b
  return make_double2(fma(-a.y, b.y, fma(a.x, b.x, s.x)), fma(a.y, b.x, fma(a.x, b.y, s.y)));
}
__device__ __forceinline__ double2 c_fma_conj(double2 s, double2 a, double2 b) {
  return make_double2(fma(a.y, b.y, fma(a.x, b.x, s.x)), fma(-a.y, b.x, fma(a.x, b.y, s.y)));
}
// a / b with Smith's scaling (what the compilers' complex division does)
__device__ __forceinline__ double2 c_div(double2 a, double2 b) {
  if (fabs(b.x) < fabs(b.y)) {
    const double r = b.x / b.y, den = b.x * r + b.y;
    return make_double2((a.x * r + a.y) / den, (a.y * r - a.x) / den);
  }
  const double r = b.y / b.x, den = b.y * r + b.x;
  return make_double2((a.y * r + a.x) / den, (a.y - a.x * r) / den);
}
// principal square root of the REAL number x seen as x + 0i (csqrt of green.f90:1255, density_of_states.f90:396)
__device__ __forceinline__ double2 c_sqrt_real(double x) {
  return x >= 0.0 ? make_double2(sqrt(x), 0.0) : make_double2(0.0, sqrt(-x));
}

// ---- terminator ------------------------------------------------------------------------------------------------
// Chain c reads a(l) = A[off(c) + (l-1)*lstride], rb(l) = RB[off(c) + (l-1)*lstride], l = 1..ll, with
// off(c) = (c / inner)*outer_stride + (c % inner)*inner_stride  (all in doubles).
struct ChainLayout {
  long long inner, inner_stride, outer_stride, lstride;
};

// The reference's bpopt builds AZ/RBZ and emami copies them; the copies are pure functions of (a, rb, ainf), so they
// are recomputed on the fly (same operations, same roundings) instead of being stored per thread.
struct BpoptChain {
  const double *a, *rb;
  long long ls;
  int n;  // = ll - 1
  double ainf;
  // optional per-thread copies in shared memory (element i of this thread at s?z[i * sstride]): the Sturm counts read
  // every coefficient ~50 times per emami call, and from global memory each read is a dependent L1/L2 access
  double *saz = nullptr, *sbz = nullptr;
  int sstride = 0;
  __device__ __forceinline__ double az_g(int i) const {
    const double d = __dsub_rn(__ldg(a + (long long)(i - 1) * ls), ainf);
    return i == n ? d : __dmul_rn(0.5, d);
  }
  __device__ __forceinline__ double bz_g(int i) const {  // emami's B: B(1) = 0, B(n+1) = 0
    if (i <= 1 || i > n) return 0.0;
    const double r = __ldg(rb + (long long)(i - 1) * ls);
    return i == n ? __dmul_rn(0.70710678118654746 /* 1/sqrt(2) as the reference's double expression */, r) : __dmul_rn(0.5, r);
  }
  __device__ __forceinline__ void stage() {  // after every change of ainf
    if (!saz) return;
    for (int i = 1; i <= n + 1; i++) {
      if (i <= n) saz[i * sstride] = az_g(i);
      sbz[i * sstride] = bz_g(i);
    }
  }
  __device__ __forceinline__ double az(int i) const { return saz ? saz[i * sstride] : az_g(i); }
  __device__ __forceinline__ double bz(int i) const { return sbz ? sbz[i * sstride] : bz_g(i); }
  // shared-memory carve-up for a CTA of nthreads: [2][n + 2][nthreads] doubles
  __device__ __forceinline__ void attach(double *smem, int nthreads, int tid) {
    saz = smem + tid;
    sbz = smem + (size_t)(n + 2) * nthreads + tid;
    sstride = nthreads;
  }
};
static inline size_t bpopt_smem_bytes(int ll, int nthreads) { return (size_t)2 * (ll + 1) * nthreads * sizeof(double); }

__device__ inline int sturm_count(const BpoptChain &c, double e) {
  const double relfeh = 1.8189894035458565e-12;  // 2^-39
  int num = 0;
  double p = __dsub_rn(c.az(1), e);
  if (p < 0.0) num++;
  for (int i = 2; i <= c.n; i++) {
    const double d = __dsub_rn(c.az(i), e), b = c.bz(i);
    if (p == 0.0) p = __dsub_rn(d, __ddiv_rn(fabs(b), relfeh));
    else p = __dsub_rn(d, __ddiv_rn(__dmul_rn(b, b), p));
    if (p < 0.0) num++;
  }
  return num;
}

// emami (recursion.f90:3589-3706), including its early return ("goto 1000") after 50 bisections
__device__ inline void dev_emami(const BpoptChain &c, double &emax_o, double &emin_o) {
  const int n = c.n;
  double emax0 = -1.0e6, emin0 = 1.0e6;
  for (int i = 1; i <= n; i++) {
    const double ai = c.az(i), b0 = fabs(c.bz(i)), b1 = fabs(c.bz(i + 1));
    const double x1 = __dadd_rn(__dadd_rn(ai, b0), b1), x2 = __dsub_rn(__dsub_rn(ai, b0), b1);
    if (emax0 <= x1) emax0 = x1;
    if (emin0 > x2) emin0 = x2;
  }
  const double eps = 1.0e-6;
  double emax = emax0, emin = emin0, e = 0.0;
  for (int istop = 1;; istop++) {
    e = __ddiv_rn(__dadd_rn(emax, emin), 2.0);
    if (istop > 50) { emax_o = emax; emin_o = emin; return; }
    const int num = sturm_count(c, e);
    if (num == n) emax = e;
    if (num < n) emin = e;
    const double dele = fabs(__ddiv_rn(__dsub_rn(emax, emin), __ddiv_rn(__dadd_rn(emax, emin), 2.0)));
    if (dele <= eps) break;
  }
  const double e1 = e;
  emax = e1; emin = emin0;
  for (int istop = 1;; istop++) {
    e = __ddiv_rn(__dadd_rn(emax, emin), 2.0);
    if (istop > 50) { emax_o = emax; emin_o = emin; return; }
    const int num = sturm_count(c, e);
    if (num == 0) emin = e;
    if (num > 0) emax = e;
    const double dele = fabs(__ddiv_rn(__dsub_rn(emax, emin), __ddiv_rn(__dadd_rn(emax, emin), 2.0)));
    if (dele <= eps) break;
  }
  emax_o = e1;
  emin_o = e;
}

// ---- warp-per-chain emami: the bisections are sequential chains of Sturm counts (20 dependent divisions each), so a
// thread-per-chain kernel is pure latency.  A warp instead evaluates the 31 nodes of the next FIVE bisection levels at
// once: node k (heap numbering, lane k-1) derives its trial energy from the bounds its ancestors WOULD have set (only
// halvings, no counts needed), all lanes count in parallel, and the true path is then walked with shuffles using the
// reference's branch and stop rules.  Every number on the true path is computed by the same operations as the serial
// loop, so the result is bit-identical; ~22 serial counts become 5 rounds.
// kind 0: first bisection (num == n -> emax = e ; num < n -> emin = e); kind 1: second (num == 0 -> emin ; else emax).
__device__ inline bool emami_rounds(const BpoptChain &c, int kind, double &emax, double &emin, double &e_last, int lane) {
  const unsigned FULL = 0xffffffffu;
  const int n = c.n;
  const double eps = 1.0e-6;
  int istop = 1;
  for (;;) {
    // this lane's node: k = lane + 1 in 1..31 (lane 31 repeats node 31), depth d = floor(log2 k)
    const int k = min(lane + 1, 31);
    const int d = 31 - __clz(k);
    double hi = emax, lo = emin, e = 0.0;
    for (int lev = 0; lev <= d; lev++) {
      e = __ddiv_rn(__dadd_rn(hi, lo), 2.0);
      if (lev == d) break;
      const int bit = (k >> (d - 1 - lev)) & 1;  // 1: the branch that moves emax down to e
      if (bit) hi = e; else lo = e;
    }
    const int num = sturm_count(c, e);
    // walk the true path
    int node = 1;
    for (int lev = 0; lev < 5; lev++) {
      const double en = __shfl_sync(FULL, e, node - 1);
      const int cnt = __shfl_sync(FULL, num, node - 1);
      e_last = en;
      if (istop > 50) return true;  // the reference's "goto 1000": caller returns (emax, emin) as they are
      int bit;
      if (kind == 0) { bit = cnt == n; if (cnt == n) emax = en; if (cnt < n) emin = en; }
      else { bit = cnt > 0; if (cnt == 0) emin = en; if (cnt > 0) emax = en; }
      const double dele = fabs(__ddiv_rn(__dsub_rn(emax, emin), __ddiv_rn(__dadd_rn(emax, emin), 2.0)));
      if (dele <= eps) return false;
      istop++;
      node = 2 * node + bit;
    }
  }
}

__device__ inline void dev_emami_warp(const BpoptChain &c, double &emax_o, double &emin_o, int lane) {
  const int n = c.n;
  double emax0 = -1.0e6, emin0 = 1.0e6;
  for (int i = 1; i <= n; i++) {
    const double ai = c.az(i), b0 = fabs(c.bz(i)), b1 = fabs(c.bz(i + 1));
    const double x1 = __dadd_rn(__dadd_rn(ai, b0), b1), x2 = __dsub_rn(__dsub_rn(ai, b0), b1);
    if (emax0 <= x1) emax0 = x1;
    if (emin0 > x2) emin0 = x2;
  }
  double emax = emax0, emin = emin0, e = 0.0;
  if (emami_rounds(c, 0, emax, emin, e, lane)) { emax_o = emax; emin_o = emin; return; }
  const double e1 = e;
  emax = e1; emin = emin0;
  if (emami_rounds(c, 1, emax, emin, e, lane)) { emax_o = emax; emin_o = emin; return; }
  emax_o = e1;
  emin_o = e;
}

// bpopt for nchains chains, one WARP per chain; diag != 0: chain = (orbital i, unit), results go to element (i,i) of the
// (18,18,unit) arrays (the 18 diagonal chains block_green consumes); shared memory: 2 (ll + 1) doubles per warp
__global__ void k_bpopt_warp(const double *A, const double *RB, ChainLayout lay, int ll, int nchains, double *ainf_o,
                             double *rbinf_o, int *ifail_o, int diag) {
  extern __shared__ double bp_smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const int ch = blockIdx.x * wpb + wib;
  if (ch >= nchains) return;
  const long long off = (ch / lay.inner) * lay.outer_stride + (ch % lay.inner) * lay.inner_stride;
  BpoptChain c;
  c.a = A + off; c.rb = RB + off; c.ls = lay.lstride; c.n = ll - 1;
  c.ainf = __ldg(c.a + (long long)(c.n - 1) * c.ls);
  c.saz = bp_smem + (size_t)wib * 2 * (ll + 1);
  c.sbz = c.saz + (ll + 1);
  c.sstride = 1;
  double bmax = 0.0, bmin = 0.0;
  int ifail = 0;
  for (int jiter = 1;; jiter++) {
    for (int i = 1 + lane; i <= c.n + 1; i += 32) {
      if (i <= c.n) c.saz[i] = c.az_g(i);
      c.sbz[i] = c.bz_g(i);
    }
    __syncwarp();
    dev_emami_warp(c, bmax, bmin, lane);
    __syncwarp();
    const double s = __dadd_rn(bmax, bmin);
    c.ainf = __dadd_rn(c.ainf, s);
    if (fabs(s) <= 1.0e-05) break;
    if (jiter > 300) { ifail = 1; break; }
  }
  if (lane == 0) {
    const double rbinf = __ddiv_rn(__dsub_rn(bmax, bmin), 2.0);
    if (diag) {
      const int i = ch % NB, unit = ch / NB;
      ainf_o[(size_t)unit * BLKC + i * (NB + 1)] = c.ainf;
      rbinf_o[(size_t)unit * BLKC + i * (NB + 1)] = rbinf;
    } else {
      ainf_o[ch] = c.ainf;
      rbinf_o[ch] = rbinf;
      if (ifail_o) ifail_o[ch] = ifail;
    }
  }
}

// bpopt (recursion.f90:3540-3581) for nchains independent chains
__global__ void k_bpopt(const double *A, const double *RB, ChainLayout lay, int ll, int nchains, double *ainf_o,
                        double *rbinf_o, int *ifail_o, int use_smem) {
  const int ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= nchains) return;
  const long long off = (ch / lay.inner) * lay.outer_stride + (ch % lay.inner) * lay.inner_stride;
  extern __shared__ double bp_smem[];
  BpoptChain c;
  c.a = A + off; c.rb = RB + off; c.ls = lay.lstride; c.n = ll - 1;
  c.ainf = __ldg(c.a + (long long)(c.n - 1) * c.ls);  // AINF = A(N)
  if (use_smem) c.attach(bp_smem, blockDim.x, threadIdx.x);
  double bmax = 0.0, bmin = 0.0;
  int ifail = 0;
  for (int jiter = 1;; jiter++) {
    c.stage();
    dev_emami(c, bmax, bmin);
    const double s = __dadd_rn(bmax, bmin);
    c.ainf = __dadd_rn(c.ainf, s);
    if (fabs(s) <= 1.0e-05) break;
    if (jiter > 300) { ifail = 1; break; }
  }
  ainf_o[ch] = c.ainf;
  rbinf_o[ch] = __ddiv_rn(__dsub_rn(bmax, bmin), 2.0);
  if (ifail_o) ifail_o[ch] = ifail;
}

// the 18 diagonal chains of each unit only: chain c = (i, unit), results written to element (i,i) of a_inf/b_inf
__global__ void k_bpopt_diag(const double *A, const double *RB, ChainLayout lay, int ll, int nchains, double *ainf_o,
                             double *rbinf_o, int use_smem) {
  const int ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= nchains) return;
  const long long off = (ch / lay.inner) * lay.outer_stride + (ch % lay.inner) * lay.inner_stride;
  extern __shared__ double bp_smem[];
  BpoptChain c;
  c.a = A + off; c.rb = RB + off; c.ls = lay.lstride; c.n = ll - 1;
  c.ainf = __ldg(c.a + (long long)(c.n - 1) * c.ls);
  if (use_smem) c.attach(bp_smem, blockDim.x, threadIdx.x);
  double bmax = 0.0, bmin = 0.0;
  for (int jiter = 1;; jiter++) {
    c.stage();
    dev_emami(c, bmax, bmin);
    const double s = __dadd_rn(bmax, bmin);
    c.ainf = __dadd_rn(c.ainf, s);
    if (fabs(s) <= 1.0e-05 || jiter > 300) break;
  }
  const int i = ch % NB, unit = ch / NB;
  ainf_o[(size_t)unit * BLKC + i * (NB + 1)] = c.ainf;
  rbinf_o[(size_t)unit * BLKC + i * (NB + 1)] = __ddiv_rn(__dsub_rn(bmax, bmin), 2.0);
}

// get_terminf's fix-ups (recursion.f90:2110-2136): a_inf, b_inf (18,18,na) in place; a_inf0, b_inf0 (na)
__global__ void k_terminf_fix(double *a_inf, double *b_inf, double *a_inf0, double *b_inf0) {
  double *ai = a_inf + (size_t)blockIdx.x * BLKC, *bi = b_inf + (size_t)blockIdx.x * BLKC;
  const int t = threadIdx.x, i = t % NB, j = t / NB;
  if (isnan(ai[t])) ai[t] = 0.0;
  if (isnan(bi[t])) bi[t] = 0.0;
  if (i == j) {
    if (ai[t] == 0.0) ai[t] = 0.5;
    if (bi[t] == 0.0) bi[t] = 0.5;
  }
  __syncthreads();
  if (t == 0) {
    double s = 0.0;
    for (int k = 0; k < NB; k++) s = __dadd_rn(s, ai[k + NB * k]);
    a_inf0[blockIdx.x] = __ddiv_rn(s, (double)NB);
    bi[0] = __dmul_rn(bi[0], 1.01);
    bi[9 + NB * 9] = __dmul_rn(bi[9 + NB * 9], 1.01);
    s = 0.0;
    for (int k = 0; k < NB; k++) s = __dadd_rn(s, bi[k + NB * k]);
    b_inf0[blockIdx.x] = __ddiv_rn(s, (double)NB);
  }
}

// ---- block continued fraction (bgreen) ------------------------------------------------------------------------
// Two geometries: 8 warps per CTA (128 registers, 16 warps per SM) and 9 warps per CTA (<= 113 registers, a little more
// spilling, 18 warps per SM).  A chain is one warp's sequential work, so what counts is the number of ROUNDS the grid
// needs: the standard mesh of one atom (2510 energies) is 1.06 rounds of 148 x 16 warps but 0.94 of 148 x 18.
#define BG_WARPS 8
#define BG_WARPS_WIDE 9
#ifndef BG_MINB
#define BG_MINB 2
#endif
#define BG_LD 19  // padded column stride (complex) so that the 16-byte column accesses of 8 lanes hit 8 bank groups
#define BG_MAT (NB * BG_LD)
#define BG_WSTRIDE (BG_MAT + NB + 5)  // per-warp shared memory: Q, the 18 pivot reciprocals, the row permutation (18 ints)

// a_b, b_b: (18,18,ll,na) complex, b_b = B (after zsqr); g: (18,18,nv,na).  Channels ie0..ie0+ie_len-1 (0-based) are
// written, the rest of g is left untouched (the caller zeroes it, like bgreen's g_out = 0).
template <int NW>
__global__ void __launch_bounds__(NW * 32, BG_MINB)
k_bgreen(const double2 *__restrict__ a_b, const double2 *__restrict__ b_b, int ll, const double *__restrict__ ene,
         int nv, int ie0, int ie_len, const double *__restrict__ a_inf, const double *__restrict__ b_inf, double eta_re,
         double eta_im, int sym_term, double2 *__restrict__ g) {
  extern __shared__ double2 bg_smem[];
  double2 *sA = bg_smem, *sB = bg_smem + BG_MAT, *Q = bg_smem + 2 * BG_MAT + (threadIdx.x >> 5) * BG_WSTRIDE;
  double2 *dinv = Q + BG_MAT;
  int *perm = reinterpret_cast<int *>(dinv + NB);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, unit = blockIdx.y;
  const int iel = blockIdx.x * NW + warp;
  const bool live = iel < ie_len;
  const int ei = ie0 + (live ? iel : ie_len - 1);
  const bool act = lane < NB;
  const int j = act ? lane : 0;
  const double en = ene[ei];
  const double *ai = a_inf + (size_t)unit * BLKC, *bi = b_inf + (size_t)unit * BLKC;
  const double2 *au = a_b + (size_t)unit * ll * BLKC, *bu = b_b + (size_t)unit * ll * BLKC;
  double2 *q = Q + j * BG_LD;  // own column
  // terminator: Q = diag((E + eta - a_inf - sqrt((E-etop)(E-ebot)))/2)  (green.f90:1240-1262)
  if (act) {
    for (int i = 0; i < NB; i++) q[i] = make_double2(0.0, 0.0);
    double ad, bd, ws = 1.0;
    if (sym_term) {
      ad = __dmul_rn(__dadd_rn(ai[0], ai[9 + NB * 9]), 0.5);
      bd = __dmul_rn(__dadd_rn(bi[0], bi[9 + NB * 9]), 0.5);
    } else {
      ad = ai[j + NB * j];
      bd = bi[j + NB * j];
      if (j == 0 || j == 9) ws = 1.025;
    }
    const double hw = __dmul_rn(__dmul_rn(2.0, bd), ws);  // explicit roundings: E close to a band edge amplifies them
    const double etop = __dadd_rn(ad, hw), ebot = __dsub_rn(ad, hw);
    const double2 zoff = c_sqrt_real(__dmul_rn(__dsub_rn(en, etop), __dsub_rn(en, ebot)));
    q[j] = make_double2(((en + eta_re) - ad - zoff.x) * 0.5, (eta_im - zoff.y) * 0.5);
  }
  const double2 P = en != 0.0 ? make_double2(en + eta_re, eta_im) : make_double2(en, 0.0);
  for (int l = ll - 1; l >= 1; l--) {
    __syncthreads();  // every warp is done with the previous level's B
    for (int t = threadIdx.x; t < BLKC; t += NW * 32) {
      sA[(t / NB) * BG_LD + t % NB] = au[(size_t)(l - 1) * BLKC + t];
      sB[(t / NB) * BG_LD + t % NB] = bu[(size_t)(l - 1) * BLKC + t];
    }
    __syncthreads();
    // Q <- P - A - Q   (own column)
    if (act) {
      const double2 *ac = sA + j * BG_LD;
      const double2 qjj = q[j];
      for (int i = 0; i < NB; i++) q[i] = c_sub(c_sub(make_double2(0.0, 0.0), ac[i]), q[i]);
      q[j] = c_sub(c_sub(P, ac[j]), qjj);
    }
    __syncwarp();
    // LU with partial pivoting (zgetrf semantics); perm = the row permutation (P x)(i) = x(perm(i))
    if (act) perm[j] = j;
    __syncwarp();
#pragma unroll
    for (int k = 0; k < NB; k++) {
      // pivot of column k (izamax rule: first largest |re| + |im| among rows k..17): lane i looks at row i, a shuffle
      // tree keeps the larger value, the smaller row on ties -- the serial scan by lane k alone was 17 dependent steps
      double2 *ck = Q + k * BG_LD;
      const bool below = lane >= k && lane < NB;
      const double2 mine = below ? ck[lane] : make_double2(0.0, 0.0);
      double best = below ? fabs(mine.x) + fabs(mine.y) : -1.0;
      int p = below ? lane : NB;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int op = __shfl_xor_sync(0xffffffffu, p, o);
        if (ov > best || (ov == best && op < p)) { best = ov; p = op; }
      }
      if (lane == 0 && p != k) { const int t = perm[k]; perm[k] = perm[p]; perm[p] = t; }
      // 1/u_kk = conj(u)/|u|^2: one division; kept for the back substitution (x_i /= u_ii becomes a product).
      // Every lane forms it (u = row p of column k before the swap), lane i scales its own multiplier l_ik.
      const double2 u = ck[p];
      const double inv = 1.0 / (u.x * u.x + u.y * u.y);
      const double2 r = make_double2(u.x * inv, -u.y * inv);
      __syncwarp();  // column k has been read by every lane
      if (act && p != k) { const double2 t = q[k]; q[k] = q[p]; q[p] = t; }
      __syncwarp();
      if (lane == k) dinv[k] = r;
      if (lane > k && lane < NB) ck[lane] = c_mul(ck[lane], r);
      __syncwarp();
      if (act && j > k) {
        const double2 uj = q[k];
#pragma unroll
        for (int i = k + 1; i < NB; i++) q[i] = c_fnma(q[i], ck[i], uj);
      }
      __syncwarp();
    }
    // Q_new(:,j) = B^H (M^-1 B(:,j)): the reference inverts M and multiplies twice (zgetrf, zgetri, two zgemm: green.f90:1317-1323); solving
    // M x = B(:,j) with the factors gives the same column without forming M^-1 (one 18^3 product less per level; the
    // results differ by rounding only, parity tolerance 1e-10).  x = P b_j, forward, backward (registers, fully unrolled)
    const double2 *bc = sB + j * BG_LD;
    double2 x[NB];
#pragma unroll
    for (int i = 0; i < NB; i++) x[i] = bc[perm[i]];
#pragma unroll
    for (int i = 1; i < NB; i++) {
#pragma unroll
      for (int k = 0; k < i; k++) x[i] = c_fnma(x[i], Q[k * BG_LD + i], x[k]);
    }
#pragma unroll
    for (int i = NB - 1; i >= 0; i--) {
#pragma unroll
      for (int k = i + 1; k < NB; k++) x[i] = c_fnma(x[i], Q[k * BG_LD + i], x[k]);
      x[i] = c_mul(x[i], dinv[i]);
    }
    __syncwarp();  // all lanes have read the factors
    if (act) {
#pragma unroll
      for (int i = 0; i < NB; i++) {
        const double2 *bi_ = sB + i * BG_LD;
        double2 s = make_double2(0.0, 0.0);
#pragma unroll
        for (int k = 0; k < NB; k++) s = c_fma_conj(s, bi_[k], x[k]);
        q[i] = s;
      }
    }
    __syncwarp();
  }
  if (act && live) {
    double2 *go = g + ((size_t)unit * nv + ei) * BLKC + NB * j;
    for (int i = 0; i < NB; i++) go[i] = q[i];
  }
}

// ---- Chebyshev Green function -------------------------------------------------------------------------------------
// mu_ng = mu_n * kernel(i), and * 2 for i >= 2 (green.f90:1063-1068); blocks of 324 complex, nk per unit
__global__ void k_cheb_weight(const double2 *mu, const double *kernel, int nk, size_t total, double2 *mg) {
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const int i = (int)((t / BLKC) % nk);
    double2 v = c_scale(mu[t], kernel[i]);
    if (i >= 1) v = c_scale(v, 2.0);
    mg[t] = v;
  }
}

#define CG_EB 8    // energies per CTA
#define CG_CH 32   // moments per staged chunk
#define CG_THREADS 352
// g0(:,:,ie,n) = sum_i mu_ng(:,:,i,n) * (-i exp(-i (i-1) acos(w))) / sqrt(a^2 - (e-b)^2)   (green.f90:1077-1092)
__global__ void __launch_bounds__(CG_THREADS)
k_cheb_green(const double2 *__restrict__ mg, int nk, const double *__restrict__ ene, int nv, double a, double b,
             double2 *__restrict__ g0) {
  __shared__ double2 f[CG_CH][CG_EB];
  __shared__ double theta[CG_EB];
  const int tid = threadIdx.x, unit = blockIdx.y, e0 = blockIdx.x * CG_EB;
  if (tid < CG_EB) {
    const int ie = min(e0 + tid, nv - 1);
    theta[tid] = acos(__ddiv_rn(__dsub_rn(ene[ie], b), a));
  }
  double2 acc[CG_EB];
#pragma unroll
  for (int e = 0; e < CG_EB; e++) acc[e] = make_double2(0.0, 0.0);
  const double2 *mu = mg + (size_t)unit * nk * BLKC;
  for (int i0 = 0; i0 < nk; i0 += CG_CH) {
    __syncthreads();
    for (int t = tid; t < CG_CH * CG_EB; t += CG_THREADS) {
      const int ii = t / CG_EB, e = t % CG_EB;
      double s, c;
      sincos(__dmul_rn((double)(i0 + ii), theta[e]), &s, &c);
      f[ii][e] = make_double2(-s, -c);
    }
    __syncthreads();
    if (tid < BLKC) {
      const int nch = min(CG_CH, nk - i0);
      for (int ii = 0; ii < nch; ii++) {
        const double2 m = mu[(size_t)(i0 + ii) * BLKC + tid];
#pragma unroll
        for (int e = 0; e < CG_EB; e++) acc[e] = c_fma(acc[e], m, f[ii][e]);
      }
    }
  }
  if (tid < BLKC) {
#pragma unroll
    for (int e = 0; e < CG_EB; e++) {
      const int ie = e0 + e;
      if (ie >= nv) break;
      const double d = __dsub_rn(ene[ie], b);
      const double den = sqrt(__dsub_rn(__dmul_rn(a, a), __dmul_rn(d, d)));
      g0[((size_t)unit * nv + ie) * BLKC + tid] = make_double2(acc[e].x / den, acc[e].y / den);
    }
  }
}

// ---- scalar continued fraction -----------------------------------------------------------------------------------
__global__ void k_sqrt_array(const double *in, double *out, size_t n) {
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (size_t)gridDim.x * blockDim.x) out[t] = sqrt(in[t]);
}

// density (density_of_states.f90:248-372) for nch = na*nmdir (atom, direction) chains-of-18:
// a, b2: (lld,18,nch); am1, bm1: bpopt results (18,nch); dw_l, cshi: (18,na) indexed by atom = ch % na;
// tdens: (18,nv,nch)
__global__ void k_density(const double *a, const double *b2, int lld, int nch, int na, const double *am1,
                          const double *bm1, const double *ene, int nv, const double *dw_l, const double *cshi,
                          double *tdens) {
  const size_t total = (size_t)NB * nv * nch;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const int nl = (int)(t % NB), ie = (int)((t / NB) % nv), ch = (int)(t / ((size_t)NB * nv)), ia = ch % na;
    double bm = bm1[nl + NB * ch];
    if (nl == 0 || nl == 9) bm = __dmul_rn(1.01, bm);
    const double ebot = __dsub_rn(am1[nl + NB * ch], __dmul_rn(2.0, bm)), etop = __dadd_rn(ebot, __dmul_rn(4.0, bm));
    const double dw = dw_l[nl + NB * ia];
    const double e = __dsub_rn(__ddiv_rn(ene[ie], dw), cshi[nl + NB * ia]);
    // bprldos (378-407)
    const double emid = 0.5 * (etop + ebot);
    const double2 zoff = c_sqrt_real((e - etop) * (e - ebot));
    double2 Qt = make_double2((e - emid - zoff.x) * 0.5, (0.0 - zoff.y) * 0.5);
    if (Qt.y > 0.0) Qt = make_double2((e - emid + zoff.x) * 0.5, zoff.y * 0.5);
    const double *ac = a + (size_t)lld * (nl + NB * ch), *bc = b2 + (size_t)lld * (nl + NB * ch);
    for (int l = lld - 1; l >= 1; l--)
      Qt = c_div(make_double2(bc[l - 1], 0.0), make_double2(e - ac[l - 1] - Qt.x, -Qt.y));
    tdens[t] = 0.0 + 1.0 * (-Qt.y / PI_RP) / dw;
  }
}

// sgreen (green.f90:628-705): g0 (18,18,nv,na) from doso = tdens (18,nv,na,nmdir); g0 must be zeroed by the caller
__global__ void k_sgreen_assemble(const double *doso, int nv, int na, int nmdir, double2 *g0) {
  const size_t total = (size_t)(nmdir == 1 ? NB : 9) * nv * na;
  const int per = nmdir == 1 ? NB : 9;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const int j = (int)(t % per), ie = (int)((t / per) % nv), ia = (int)(t / ((size_t)per * nv));
    double2 *gb = g0 + ((size_t)ia * nv + ie) * BLKC;
    if (nmdir == 1) {
      gb[j + NB * j] = make_double2(0.0, -doso[j + NB * (ie + (size_t)nv * ia)] * PI_RP);
      continue;
    }
    const double2 dfac = make_double2(0.0, PI_RP / 2.0);  // i*pi/2
    const double2 gfac[2][3] = {{{1.0, 0.0}, {0.0, -1.0}, {1.0, 0.0}}, {{1.0, 0.0}, {0.0, 1.0}, {-1.0, 0.0}}};
    const int goff[4][3] = {{0, 0, 0}, {9, 9, 0}, {9, 9, 9}, {0, 0, 9}};
    for (int mdir = 0; mdir < nmdir; mdir++) {
      const double *d = doso + (size_t)NB * (ie + (size_t)nv * (ia + (size_t)na * mdir));
      const double up = d[j], dn = d[j + 9];
      const double2 ch = c_scale(c_scale(dfac, up + dn), 1.0 / 3.0);
      gb[j + NB * j] = c_sub(gb[j + NB * j], ch);
      gb[(j + 9) + NB * (j + 9)] = c_sub(gb[(j + 9) + NB * (j + 9)], ch);
      const int e1 = (j + goff[0][mdir]) + NB * (j + goff[1][mdir]), e2 = (j + goff[2][mdir]) + NB * (j + goff[3][mdir]);
      gb[e1] = c_sub(gb[e1], c_mul(c_scale(gfac[0][mdir], up - dn), dfac));
      gb[e2] = c_sub(gb[e2], c_mul(c_scale(gfac[1][mdir], up - dn), dfac));
    }
  }
}

// ---- Kubo-Bastin back end -----------------------------------------------------------------------------------------
// Tables of calculate_gamma_nm (conductivity.f90:184-213), [n][i] so that energies are coalesced:
//   CN = cn(i,n) s_n, CM = cm(i,n) s_n, TS = T_n(w_i) s_n with s_n = g_kernel(n) weights(n);  inv(i) = 1/(1-w^2)^2
__global__ void k_cond_tables(const double *ene, int nv, int M, double a, double b, const double *sk, double2 *CN,
                              double2 *CM, double *TS, double *inv) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nv) return;
  const double w = __ddiv_rn(__dsub_rn(ene[i], b), a), ac = acos(w), om = __dsub_rn(1.0, __dmul_rn(w, w)), sq = sqrt(om);
  inv[i] = __dmul_rn(om, om);
  double t0 = 1.0, t1 = w;
  for (int n = 0; n < M; n++) {
    double tn;
    if (n == 0) tn = 1.0;
    else if (n == 1) tn = w;
    else { tn = __dsub_rn(__dmul_rn(__dmul_rn(2.0, w), t1), t0); t0 = t1; t1 = tn; }
    double s, c;
    const double fn = (double)n;
    sincos(__dmul_rn(fn, ac), &s, &c);
    const double q = __dmul_rn(fn, sq);
    // cn = (w - i n sq)(c + i s), cm = (w + i n sq)(c - i s)
    const double2 cn = c_mul(make_double2(w, -q), make_double2(c, s)), cm = c_mul(make_double2(w, q), make_double2(c, -s));
    const size_t o = (size_t)n * nv + i;
    CN[o] = c_scale(cn, sk[n]);
    CM[o] = c_scale(cm, sk[n]);
    TS[o] = tn * sk[n];
  }
}

// D[t][n][m][l2] = mu_nm(l2,l2,n,m,t)  (block index n + M m, complex column-major blocks)
__global__ void k_cond_diag(const double2 *mu, int M, size_t nblocks, double2 *D) {
  const size_t total = nblocks * NB;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const int l2 = (int)(t % NB);
    const size_t blk = t / NB, mm = (size_t)M * M, tt = blk / mm, r = blk % mm, n = r % M, m = r / M;
    D[((tt * M + n) * M + m) * NB + l2] = mu[blk * BLKC + l2 + NB * l2];
  }
}

#define CD_THREADS 128
// part[t][y][l2][i] = sum_{n in chunk y} sum_m Gamma'(i,n,m) D[t][n][m][l2],  Gamma' = CN_n TS_m + CM_m TS_n
__global__ void __launch_bounds__(CD_THREADS)
k_cond_contract(const double2 *__restrict__ CN, const double2 *__restrict__ CM, const double *__restrict__ TS,
                const double2 *__restrict__ D, int nv, int M, int nchunk, double2 *__restrict__ part) {
  const int i = blockIdx.x * CD_THREADS + threadIdx.x, y = blockIdx.y, t = blockIdx.z;
  const int ic = min(i, nv - 1);
  const int n0 = (int)((long long)M * y / nchunk), n1 = (int)((long long)M * (y + 1) / nchunk);
  double2 acc[NB];
#pragma unroll
  for (int l = 0; l < NB; l++) acc[l] = make_double2(0.0, 0.0);
  for (int n = n0; n < n1; n++) {
    const double2 cn = CN[(size_t)n * nv + ic];
    const double tn = TS[(size_t)n * nv + ic];
    const double2 *d = D + (((size_t)t * M + n) * M) * NB;
    for (int m = 0; m < M; m++) {
      const double tm = TS[(size_t)m * nv + ic];
      const double2 cm = CM[(size_t)m * nv + ic];
      const double2 G = make_double2(cn.x * tm + cm.x * tn, cn.y * tm + cm.y * tn);
#pragma unroll
      for (int l = 0; l < NB; l++) acc[l] = c_fma(acc[l], G, d[(size_t)m * NB + l]);
    }
  }
  if (i < nv) {
    double2 *p = part + (((size_t)t * nchunk + y) * NB) * nv;
#pragma unroll
    for (int l = 0; l < NB; l++) p[(size_t)l * nv + i] = acc[l];
  }
}

// integrand_at(l2,i,t) = factor/(1-w^2)^2 * sum_y part; integrand(l2,i) = sum_t (conductivity.f90:276-290)
__global__ void k_cond_finish(const double2 *part, const double *inv, int nv, int nchunk, int nloop, double factor,
                              double2 *integrand, double2 *integrand_at) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= NB * nv) return;
  const int l2 = t % NB, i = t / NB;
  double2 tot = make_double2(0.0, 0.0);
  for (int tt = 0; tt < nloop; tt++) {
    double2 s = make_double2(0.0, 0.0);
    for (int y = 0; y < nchunk; y++) s = c_add(s, part[(((size_t)tt * nchunk + y) * NB + l2) * nv + i]);
    s = c_scale(s, factor / inv[i]);
    if (integrand_at) integrand_at[l2 + (size_t)NB * (i + (size_t)nv * tt)] = s;
    tot = c_add(tot, s);
  }
  integrand[l2 + (size_t)NB * i] = tot;
}

// ---- create_ll_map (recursion.f90:3277-3303): one reachability level ------------------------------------------
// prev/next: izeroll(0:kk, ll) / izeroll(0:kk, ll+1); nbr [ng][kk] 0-based with kk = "no neighbour", row 0 = self.
__global__ void k_ll_map_step(const int32_t *__restrict__ nbr, int ng, int kk, const int32_t *__restrict__ prev,
                              int32_t *__restrict__ next) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < kk; i += gridDim.x * blockDim.x) {
    int v = prev[i + 1];
    for (int j = 1; j < ng && !v; j++) {
      const int nb = nbr[(size_t)j * kk + i];
      if (nb < kk && prev[nb + 1] != 0) v = 1;
    }
    next[i + 1] = v;
    if (i == 0) next[0] = 0;
  }
}

// ---- chebyshev_orbital_mod helpers (recursion.f90:2949-2979): RI36 block vectors ---------------------------------
// out_k = (pos[comp + 3 k] * in_k) * alat
__global__ void k_scale_by_pos(const double *__restrict__ in, double *__restrict__ out, const double *__restrict__ pos,
                               int comp, double alat, int kk) {
  const size_t total = (size_t)kk * BLKD;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x)
    out[t] = __dmul_rn(__dmul_rn(pos[comp + 3 * (t / BLKD)], in[t]), alat);
}
// out = i (x - y):  (re, im) -> (-(xi - yi), xr - yr); RI36 column = 18 reals then 18 imaginaries
__global__ void k_i_times_diff(const double *__restrict__ x, const double *__restrict__ y, double *__restrict__ out, int kk) {
  const size_t total = (size_t)kk * BLKC;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const size_t col = t / NB, r = t % NB, o = col * COLD + r;
    const double dr = x[o] - y[o], di = x[o + NB] - y[o + NB];
    out[o] = -di;
    out[o + NB] = dr;
  }
}
__global__ void k_add_block(double *dst, const double *src) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < BLKD) dst[t] += src[t];
}

// ---- calculate_intersite_gf (green.f90:425-469) ---------------------------------------------------------------------
// meta[2p] = first unit of pair p in g0 (18,18,nv,nunits), meta[2p+1] = 1 when i == j (only that unit is used).
// gij = ((g1 - g2) + (1/i g3 - 1/i g4)) / 2, gji = ((g1 - g2) - (1/i g3 - 1/i g4)) / 2
__global__ void k_intersite_combine(const double2 *__restrict__ g0, int nv, int njij, const int32_t *__restrict__ meta,
                                    double2 *__restrict__ gij, double2 *__restrict__ gji) {
  const size_t total = (size_t)njij * nv * BLKC;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const int e = (int)(t % BLKC), ie = (int)((t / BLKC) % nv), p = (int)(t / ((size_t)BLKC * nv));
    const int u = meta[2 * p];
    const double2 g1 = g0[((size_t)u * nv + ie) * BLKC + e];
    if (meta[2 * p + 1]) { gij[t] = g1; gji[t] = g1; continue; }
    const double2 g2 = g0[((size_t)(u + 1) * nv + ie) * BLKC + e], g3 = g0[((size_t)(u + 2) * nv + ie) * BLKC + e],
                  g4 = g0[((size_t)(u + 3) * nv + ie) * BLKC + e];
    const double2 d = c_sub(g1, g2);
    const double2 m3 = make_double2(g3.y, -g3.x), m4 = make_double2(g4.y, -g4.x);  // (1/i) g = -i g
    const double2 tt = c_sub(m3, m4);
    gij[t] = c_scale(c_add(d, tt), 0.5);
    gji[t] = c_scale(c_sub(d, tt), 0.5);
  }
}
// gspin (9,9,nv,njij,8): Ginmag, Gix, Giy, Giz, Gjnmag, Gjx, Gjy, Gjz (green.f90:452-465)
__global__ void k_intersite_pauli(const double2 *__restrict__ gij, const double2 *__restrict__ gji, int nv, int njij,
                                  double2 *__restrict__ gspin) {
  const size_t nblk = (size_t)njij * nv, total = nblk * 81;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const int e = (int)(t % 81), j = e % 9, i = e / 9;
    const size_t blk = t / 81;
#pragma unroll
    for (int w = 0; w < 2; w++) {
      const double2 *g = (w ? gji : gij) + blk * BLKC;
      const double2 uu = g[j + 18 * i], dd = g[j + 9 + 18 * (i + 9)], ud = g[j + 18 * (i + 9)], du = g[j + 9 + 18 * i];
      double2 *o = gspin + (size_t)(4 * w) * total + t;
      o[0] = c_scale(c_add(uu, dd), 0.5);
      o[total] = c_scale(c_add(ud, du), 0.5);                                        // x
      const double2 iud = make_double2(-ud.y, ud.x), idu = make_double2(-du.y, du.x);  // i * g
      o[2 * total] = c_scale(c_sub(iud, idu), 0.5);                                  // y
      o[3 * total] = c_scale(c_sub(uu, dd), 0.5);                                    // z
    }
  }
}

// ---- tail of calculate_conductivity_tensor (conductivity.f90:300-372) ------------------------------------------------
// For every mesh energy i the reference integrates each integrand series with simpson_f (math.f90:1600-1632) using the
// Fermi function at T = 0 (kBT = 1e-15) and E_F = wscale(i): f = 1 below, 1/2 at, 0 above the point, so the O(nv^2)
// loop with an exp per term is a running Simpson sum.  One thread per series walks the panels once, adding them in the
// reference's order (bit-identical to the literal loop); the boundary panel(s) are evaluated with the reference's
// grouping  (Y(I-1) f + 4 Y(I) f) + Y(I+1) f.  Series: s = 0 total, 1..18 orbital l2; component c = re / im;
// block g = 0 the summed integrand, 1..nat the per-type ones.  The term Y(nv+1) the reference reads one past the end
// of the arrays always carries f = 0 and is dropped.
//   integrand (18,nv) complex, integrand_at (18,nv,nat) complex; out: ((19 series, 2 comps), nv, 1+nat), divided by
//   `div` for g = 0 (real(loop_over)) and not for g > 0, as written to the cond_*.out files.
__global__ void k_cond_cumulative(const double2 *__restrict__ integrand, const double2 *__restrict__ integrand_at, int nv, int npts9,
                                  int nat, double h, double div, double *__restrict__ out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int nser = 38;
  if (t >= nser * (1 + nat)) return;
  const int g = t / nser, sc = t % nser, s = sc >> 1, c = sc & 1;
  const double2 *src = g == 0 ? integrand : integrand_at + (size_t)(g - 1) * 18 * nv;
  auto Y = [&](int j) {  // 1-based mesh index
    if (j > nv) return 0.0;
    if (s > 0) { const double2 v = __ldg(src + (s - 1) + 18 * (size_t)(j - 1)); return c ? v.y : v.x; }
    double a = 0.0;
    for (int l = 0; l < 18; l++) { const double2 v = __ldg(src + l + 18 * (size_t)(j - 1)); a = __dadd_rn(a, c ? v.y : v.x); }
    return a;
  };
  double *o = out + ((size_t)g * nv) * nser + sc;  // out[(g*nv + (i-1))*38 + sc]
  const double d = g == 0 ? div : 1.0;
  // the literal loop multiplies EVERY point by its Fermi factor: a non-finite Y(j) above E_F gives NaN * 0 = NaN, so
  // one non-finite sample makes every integral NaN
  bool bad = false;
  for (int j = 1; j <= min(nv, npts9 + 1); j++) bad |= !isfinite(Y(j));
  if (bad) {
    for (int i = 0; i < nv; i++) o[(size_t)i * nser] = nan("");
    return;
  }
  auto fin = [&](double aint) { return __ddiv_rn(__ddiv_rn(__dmul_rn(h, aint), 3.0), d); };
  double P = 0.0;    // the panels I = 2, 4, .. that lie fully below the Fermi point, added term by term like the reference
  double ym = Y(1);  // Y(I-1) of the current panel
  o[0] = fin(__dmul_rn(ym, 0.5));
  for (int I = 2; I <= npts9; I += 2) {
    const double y0 = Y(I), yp = Y(I + 1);
    const double A = __dadd_rn(P, ym);
    if (I <= nv) o[(size_t)(I - 1) * nser] = fin(__dadd_rn(A, __dmul_rn(__dmul_rn(4.0, y0), 0.5)));   // E_F at the even point I
    const double F = __dadd_rn(A, __dmul_rn(4.0, y0));
    if (I + 1 <= nv) {                                                                              // E_F at the odd point I+1
      double B = __dadd_rn(F, __dmul_rn(yp, 0.5));
      if (I + 2 <= npts9) B = __dadd_rn(B, __dmul_rn(yp, 0.5));
      o[(size_t)I * nser] = fin(B);
    }
    P = __dadd_rn(F, yp);
    ym = yp;
  }
}

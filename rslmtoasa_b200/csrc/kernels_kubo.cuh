// kernels_kubo.cuh -- the Kubo-Bastin moment contraction of compute_moments_stochastic (recursion.f90:1220-1228)
// as ONE dense FP64 tensor-core GEMM per batch of right vectors, hand-written for sm_100a.
//
//   mu_nm(:,:,n,m) = sum_k left_vec(:,:,k,m)^H right_vec_n(:,:,k)                      (18x18 complex per (n,m))
//
// In the RI36 layout a column of a site block is 36 consecutive doubles [re(18) | im(18)], and
//   Re D(i,j) = <Lcol_i, Rcol_j>_36,   Im D(i,j) = <Lcol_i, (J R)col_j>_36,   (J x)[k'] = k'<18 ? x[k'+18] : -x[k'-18],
// so for KB_MB left vectors and KB_NR right vectors the contraction is the real GEMM
//   C[(m,i)][(n,j')] = sum_{site,k'} A[(m,i)][(site,k')] * B[(site,k')][(n,j')],   A = left columns, B = [R | J R]
// with M = 18*KB_MB = 72 rows (9 m-tiles), N = 36*KB_NR = 144 columns (18 n-tiles), K = 36*kk: NO padding of the
// m8n8k4 DMMA shape (the per-unit Gram kernel pads 18x36 -> 24x40, 48 % idle tensor work, and reads the right
// vector once per left vector).  Both operands are K-contiguous in shared memory exactly as the TMA bulk copy
// delivers the site blocks; the J swizzle/sign of B is applied on the fragment load.
//
// Work decomposition ("stream-K"): the linear index u = mblock*kk + site over all (left-block, site) pairs is cut
// into gridDim.x equal contiguous ranges, one per persistent CTA (148 on B200), so every SM gets the same number of
// DMMAs whatever cond_ll and kk are.  A CTA writes one partial C tile per left-block segment of its range to slot
// (cta + mblock); k_kubo_reduce sums the slots of each left block in fixed order (deterministic) into mu.
// Producer warp: per pair, 8 TMA bulk copies of 5184 B (4 left + 4 right site blocks) into a 4-stage mbarrier ring.
// 8 consumer warps: 162 (m-tile,n-tile) units per k-step split 20/20/20/21 x2 (40/40/41/41 per SM sub-partition).
#pragma once
#include "kernels_dmma.cuh"

#define KB_MB 4
#define KB_NR 4
#define KB_STAGES 4
#define KB_STAGE_D ((KB_MB + KB_NR) * BLKD)  // 5184 doubles = 41472 B
#define KB_CONSUMERS 8
#define KB_THREADS (32 * (KB_CONSUMERS + 1))
#define KB_SMEM_BYTES (KB_STAGES * KB_STAGE_D * 8 + 64)
#define KB_ROWS (KB_MB * NB)           // 72
#define KB_COLS (KB_NR * COLD)         // 144
#define KB_TILE_D (KB_ROWS * KB_COLS)  // 10368 doubles per partial C tile

__device__ __forceinline__ double flip_sign(double v, int mask) {
  return __hiloint2double(__double2hiint(v) ^ mask, __double2loint(v));
}

template <int XN>
__device__ __forceinline__ void kubo_consumer(int kk, long long lo, long long hi, const double *stages, uint64_t *full,
                                              uint64_t *empty, double *part, int warp, int lane) {
  const int g = lane >> 2, q = lane & 3;
  const int half = warp >> 2, w4 = warp & 3;
  const int mt0 = 2 * w4, mt1 = 2 * w4 + 1;                            // two private m-tiles, m-tile 8 is shared
  const int xs = (half == 0 || w4 < 3) ? 2 * w4 : 7;                    // first n-tile (within the half) of the extras
  int aoff[3], blo[9], bhi[9], bmask[9], xlo[XN], xhi[XN], xmask[XN];
  auto a_off = [&](int mt) { const int r = mt * 8 + g; return (r / NB) * BLKD + (r % NB) * COLD + q; };
  // B column J = (9*half + t)*8 + g: right vector J/36, j' = J%36; j' >= 18 is the J-rotated copy (imaginary part)
  // (lo: offset used while k' < 18, hi: while k' >= 18, mask: sign-bit flip applied while k' >= 18)
  auto b_off = [&](int t, int &lo_, int &hi_, int &mask) {
    const int J = (9 * half + t) * 8 + g, nl = J / COLD, jp = J % COLD;
    const bool sw = jp >= NB;
    const int off = (KB_MB + nl) * BLKD + (jp % NB) * COLD + q;
    mask = sw ? (int)0x80000000 : 0;
    lo_ = off + (sw ? NB : 0);
    hi_ = off - (sw ? NB : 0);
  };
  aoff[0] = a_off(mt0); aoff[1] = a_off(mt1); aoff[2] = a_off(8);
#pragma unroll
  for (int t = 0; t < 9; t++) b_off(t, blo[t], bhi[t], bmask[t]);
#pragma unroll
  for (int x = 0; x < XN; x++) b_off(xs + x, xlo[x], xhi[x], xmask[x]);
  // fragment of B at k-step ks for a (possibly swapped) column: k = 4 ks + q
  //   plain: sm[off + 4ks];  swapped: k<18 -> sm[off + 4ks + 18], k>=18 -> -sm[off + 4ks - 18]
  auto b_frag = [&](const double *sm, int lo_, int hi_, int mask, int ks) -> double {
    if (ks < 4) return sm[lo_ + 4 * ks];
    if (ks > 4) return flip_sign(sm[hi_ + 4 * ks], mask);
    const bool lowk = q < 2;  // ks == 4: k = 16 + q
    return flip_sign(sm[(lowk ? lo_ : hi_) + 16], lowk ? 0 : mask);
  };

  double acc[2][9][2], xacc[XN][2];
  auto zero = [&]() {
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
      for (int t = 0; t < 9; t++) acc[i][t][0] = acc[i][t][1] = 0.0;
#pragma unroll
    for (int x = 0; x < XN; x++) xacc[x][0] = xacc[x][1] = 0.0;
  };
  auto flush = [&](int mb) {
    double *pp = part + (size_t)(blockIdx.x + mb) * KB_TILE_D;
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
      for (int t = 0; t < 9; t++)
        *reinterpret_cast<double2 *>(pp + ((i == 0 ? mt0 : mt1) * 8 + g) * KB_COLS + (9 * half + t) * 8 + 2 * q) =
            make_double2(acc[i][t][0], acc[i][t][1]);
#pragma unroll
    for (int x = 0; x < XN; x++)
      *reinterpret_cast<double2 *>(pp + (8 * 8 + g) * KB_COLS + (9 * half + xs + x) * 8 + 2 * q) =
          make_double2(xacc[x][0], xacc[x][1]);
  };
  zero();
  int mb = (int)(lo / kk), site = (int)(lo - (long long)mb * kk);
  uint32_t it = 0;
  for (long long u = lo; u < hi; u++, it++, site++) {
    if (site == kk) { flush(mb); zero(); mb++; site = 0; }
    const int slot = it % KB_STAGES;
    mbar_wait(&full[slot], (it / KB_STAGES) & 1);
    const double *sm = stages + (size_t)slot * KB_STAGE_D;
#pragma unroll
    for (int ks = 0; ks < 9; ks++) {
      double a[3], b[9], xb[XN];
#pragma unroll
      for (int i = 0; i < 3; i++) a[i] = sm[aoff[i] + 4 * ks];
#pragma unroll
      for (int t = 0; t < 9; t++) b[t] = b_frag(sm, blo[t], bhi[t], bmask[t], ks);
#pragma unroll
      for (int x = 0; x < XN; x++) xb[x] = b_frag(sm, xlo[x], xhi[x], xmask[x], ks);
#pragma unroll
      for (int t = 0; t < 9; t++) {
        dmma(acc[0][t][0], acc[0][t][1], a[0], b[t]);
        dmma(acc[1][t][0], acc[1][t][1], a[1], b[t]);
      }
#pragma unroll
      for (int x = 0; x < XN; x++) dmma(xacc[x][0], xacc[x][1], a[2], xb[x]);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[slot]);
  }
  if (hi > lo) flush(mb);
}

// left: nmb*KB_MB vectors (stride lstride doubles), right: KB_NR vectors (stride rstride).  part: (gridDim.x + nmb) tiles.
__global__ void __launch_bounds__(KB_THREADS, 1)
k_kubo_gemm(const double *__restrict__ left, size_t lstride, int nmb, const double *__restrict__ right, size_t rstride,
            int kk, long long per, double *part) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double *stages = reinterpret_cast<double *>(smem_raw);
  uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + (size_t)KB_STAGES * KB_STAGE_D * 8);
  uint64_t *empty = full + KB_STAGES;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < KB_STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], KB_CONSUMERS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const long long total = (long long)nmb * kk;
  const long long lo = min(total, (long long)blockIdx.x * per), hi = min(total, lo + per);
  if (warp == KB_CONSUMERS) {
    if (lane < KB_MB + KB_NR) {
      int mb = (int)(lo / kk), site = (int)(lo - (long long)mb * kk);
      uint32_t it = 0;
      for (long long u = lo; u < hi; u++, it++, site++) {
        if (site == kk) { mb++; site = 0; }
        const int slot = it % KB_STAGES;
        mbar_wait(&empty[slot], ((it / KB_STAGES) & 1) ^ 1);
        double *sm = stages + (size_t)slot * KB_STAGE_D;
        if (lane == 0) mbar_expect_tx(&full[slot], KB_STAGE_D * 8);
        __syncwarp(0xff);
        const double *src = lane < KB_MB ? left + (size_t)(mb * KB_MB + lane) * lstride : right + (size_t)(lane - KB_MB) * rstride;
        bulk_g2s(sm + lane * BLKD, src + (size_t)site * BLKD, BLKD * 8, &full[slot]);
      }
    }
    return;
  }
  if (warp == 3 || warp == 6) kubo_consumer<3>(kk, lo, hi, stages, full, empty, part, warp, lane);
  else kubo_consumer<2>(kk, lo, hi, stages, full, empty, part, warp, lane);
}

// mu[((m*M + n0+nl)*648 + 2*(i + 18 j) + im] = sum over the slots of left block mb, fixed order.  grid = (nmb, KB_RED_Y)
#define KB_RED_Y 8
__global__ void __launch_bounds__(256)
k_kubo_reduce(const double *__restrict__ part, int nctas, int kk, long long per, int M, int n0, double *mu) {
  const int mb = blockIdx.x;
  const long long ulo = (long long)mb * kk, uhi = ulo + kk - 1;
  const int bfirst = (int)(ulo / per), blast = min((int)(uhi / per), nctas - 1);
  for (int e = blockIdx.y * 256 + threadIdx.x; e < KB_TILE_D; e += KB_RED_Y * 256) {
    const int r = e / KB_COLS, J = e % KB_COLS, ml = r / NB, i = r % NB, nl = J / COLD, jp = J % COLD;
    const int m = mb * KB_MB + ml, n = n0 + nl;
    if (m >= M || n >= M) continue;
    double s = 0.0;
    for (int b = bfirst; b <= blast; b++) s += part[(size_t)(b + mb) * KB_TILE_D + e];
    mu[((size_t)m * M + n) * BLKD + 2 * (i + NB * (jp % NB)) + (jp >= NB ? 1 : 0)] = s;
  }
}

static int kubo_configure() {
  return cudaFuncSetAttribute(k_kubo_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, KB_SMEM_BYTES) == cudaSuccess ? 0 : -3;
}
static int kubo_grid(int nmb, int kk, int sms) { return (int)std::max<long long>(1, std::min<long long>(sms, (long long)nmb * kk)); }
static size_t kubo_part_doubles(int nmb, int kk, int sms) { return (size_t)(kubo_grid(nmb, kk, sms) + nmb) * KB_TILE_D; }

// contracts right vectors n0..n0+3 (ring `right`) against all left vectors; results into mu(:,:,n0+nl,m)
static int kubo_launch(const double *left, size_t lstride, int M, const double *right, size_t rstride, int kk, int n0,
                       double *part, double *mu, int sms, cudaStream_t st, long long *launches) {
  const int nmb = (M + KB_MB - 1) / KB_MB, grid = kubo_grid(nmb, kk, sms);
  const long long total = (long long)nmb * kk, per = (total + grid - 1) / grid;
  k_kubo_gemm<<<grid, KB_THREADS, KB_SMEM_BYTES, st>>>(left, lstride, nmb, right, rstride, kk, per, part);
  k_kubo_reduce<<<dim3(nmb, KB_RED_Y), 256, 0, st>>>(part, grid, kk, per, M, n0, mu);
  (*launches) += 2;
  return cudaGetLastError() == cudaSuccess ? 0 : -3;
}

// ---- diagonal-only contraction ------------------------------------------------------------------------------------------
// calculate_conductivity_tensor consumes only mu_nm(l,l,n,m) (conductivity.f90:276-279): 18 of the 324 entries of every
// moment block.  When the caller does not ask for mu_nm_stochastic itself (rsrec_kubo_conductivity), the contraction is
//   D(l; m, n) = sum_{site,k} conj(L_m(k,l)) R_n(k,l)
// i.e. one real GEMM per orbital column l with K = 36*kk:  C_l[m][(n, re|im)] = A_l[m][(site,k')] B_l[(site,k')][(n,c)],
// A = column l of the left blocks, B = column l of the right blocks and of their J-rotation -- 18x fewer flops than the
// full blocks.  CTA tile: 16 left x 16 right vectors; a pipeline stage = the 9-column half (2592 B, contiguous in RI36)
// of those 32 site blocks (83 kB), two stages; 72 (column, m-tile, n-tile) units per stage = 9 per consumer warp:
// warp w owns all 8 units of column w of the half and one unit of column 8.  Work item = (left block, right block,
// site chunk); partial tiles are summed in fixed order by k_kdiag_reduce.  The kernel is L2-bandwidth bound
// (0.5 flop per byte and pair), not tensor bound.
#define KD_MB 16
#define KD_NR 16
#define KD_HALF_D 324                                   // doubles in 9 columns of a site block
#define KD_STAGE_D ((KD_MB + KD_NR) * KD_HALF_D)        // 10368 doubles = 82944 B
#define KD_STAGES 2
#define KD_CONSUMERS 8
#define KD_THREADS (32 * (KD_CONSUMERS + 1))
#define KD_SMEM_BYTES (KD_STAGES * KD_STAGE_D * 8 + 64)
#define KD_TILE_D (NB * KD_MB * 2 * KD_NR)              // doubles per partial tile: [l][m][ncol] = 9216

__global__ void __launch_bounds__(KD_THREADS, 1)
k_kubo_diag(const double *__restrict__ left, size_t lstride, int nmb, const double *__restrict__ right, size_t rstride,
            int nnb, int kk, int nchunk, double *__restrict__ part) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double *stages = reinterpret_cast<double *>(smem_raw);
  uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + (size_t)KD_STAGES * KD_STAGE_D * 8);
  uint64_t *empty = full + KD_STAGES;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < KD_STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], KD_CONSUMERS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int nitems = nmb * nnb * nchunk;
  if (warp == KD_CONSUMERS) {  // producer: 32 lanes = the 32 vectors of the tile
    uint32_t it = 0;
    for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
      const int c = item % nchunk, nb = (item / nchunk) % nnb, mb = item / (nchunk * nnb);
      const int s0 = (int)((long long)kk * c / nchunk), s1 = (int)((long long)kk * (c + 1) / nchunk);
      const double *src = lane < KD_MB ? left + (size_t)(mb * KD_MB + lane) * lstride
                                       : right + (size_t)(nb * KD_NR + lane - KD_MB) * rstride;
      for (int site = s0; site < s1; site++)
        for (int half = 0; half < 2; half++, it++) {
          const int slot = it % KD_STAGES;
          mbar_wait(&empty[slot], ((it / KD_STAGES) & 1) ^ 1);
          double *sm = stages + (size_t)slot * KD_STAGE_D;
          if (lane == 0) mbar_expect_tx(&full[slot], KD_STAGE_D * 8);
          __syncwarp();
          bulk_g2s(sm + lane * KD_HALF_D, src + (size_t)site * BLKD + half * KD_HALF_D, KD_HALF_D * 8, &full[slot]);
        }
    }
    return;
  }
  const int g = lane >> 2, q = lane & 3;
  // unit lists: column `warp` of the half with (mt, nt) = (0..1, 0..3); plus column 8 with (mt, nt) = (warp/4, warp%4)
  const int xmt = warp >> 2, xnt = warp & 3;
  // A fragment: left vector mt*8+g, column lc of the half, k' = 4ks+q  -> sm[(mt*8+g)*324 + lc*36 + 4ks + q]
  // B fragment: right vector (nt&1)*8+g; nt >= 2 is the J-rotated copy: k'<18 -> x[k'+18], k'>=18 -> -x[k'-18]
  auto bfrag = [&](const double *col, int nt, int ks) -> double {
    const double *x = col + (size_t)(KD_MB + (nt & 1) * 8 + g) * KD_HALF_D;
    const int k = 4 * ks + q;
    if (nt < 2) return x[k];
    return k < NB ? x[k + NB] : -x[k - NB];
  };
  uint32_t it = 0;
  for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
    const int c = item % nchunk;
    const int s0 = (int)((long long)kk * c / nchunk), s1 = (int)((long long)kk * (c + 1) / nchunk);
    double acc[2][8][2], xacc[2][2];  // [half][mt*4+nt][e], [half][e]
#pragma unroll
    for (int h = 0; h < 2; h++) {
#pragma unroll
      for (int u = 0; u < 8; u++) acc[h][u][0] = acc[h][u][1] = 0.0;
      xacc[h][0] = xacc[h][1] = 0.0;
    }
    for (int site = s0; site < s1; site++) {
#pragma unroll
      for (int half = 0; half < 2; half++, it++) {
        const int slot = it % KD_STAGES;
        mbar_wait(&full[slot], (it / KD_STAGES) & 1);
        const double *sm = stages + (size_t)slot * KD_STAGE_D;
        const double *colw = sm + warp * COLD, *col8 = sm + 8 * COLD;
#pragma unroll
        for (int ks = 0; ks < 9; ks++) {
          const double a0 = colw[(size_t)g * KD_HALF_D + 4 * ks + q], a1 = colw[(size_t)(8 + g) * KD_HALF_D + 4 * ks + q];
          double b[4];
#pragma unroll
          for (int nt = 0; nt < 4; nt++) b[nt] = bfrag(colw, nt, ks);
          const double xa = col8[(size_t)(xmt * 8 + g) * KD_HALF_D + 4 * ks + q], xb = bfrag(col8, xnt, ks);
#pragma unroll
          for (int nt = 0; nt < 4; nt++) {
            dmma(acc[half][nt][0], acc[half][nt][1], a0, b[nt]);
            dmma(acc[half][4 + nt][0], acc[half][4 + nt][1], a1, b[nt]);
          }
          dmma(xacc[half][0], xacc[half][1], xa, xb);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[slot]);
      }
    }
    // partial tile: part[item][l][m][ncol], ncol < 16: Re for right vector ncol, ncol >= 16: Im for right vector ncol-16
    double *pp = part + (size_t)item * KD_TILE_D;
#pragma unroll
    for (int h = 0; h < 2; h++) {
#pragma unroll
      for (int u = 0; u < 8; u++) {
        const int l = h * 9 + warp, m = (u >> 2) * 8 + g, ncol = (u & 3) * 8 + 2 * q;
        *reinterpret_cast<double2 *>(pp + ((size_t)l * KD_MB + m) * (2 * KD_NR) + ncol) = make_double2(acc[h][u][0], acc[h][u][1]);
      }
      const int l = h * 9 + 8, m = xmt * 8 + g, ncol = xnt * 8 + 2 * q;
      *reinterpret_cast<double2 *>(pp + ((size_t)l * KD_MB + m) * (2 * KD_NR) + ncol) = make_double2(xacc[h][0], xacc[h][1]);
    }
  }
}

// D[(n0 + n) * M + m][l] (complex, the layout of k_cond_diag for one start vector) = sum over the site chunks, fixed order
__global__ void k_kdiag_reduce(const double *__restrict__ part, int nmb, int nnb, int nchunk, int M, int n0, double2 *__restrict__ D) {
  const size_t total = (size_t)nmb * nnb * NB * KD_MB * KD_NR;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const int nl = (int)(t % KD_NR), ml = (int)((t / KD_NR) % KD_MB), l = (int)((t / (KD_NR * KD_MB)) % NB);
    const int nb = (int)((t / ((size_t)KD_NR * KD_MB * NB)) % nnb), mb = (int)(t / ((size_t)KD_NR * KD_MB * NB * nnb));
    const int m = mb * KD_MB + ml, n = n0 + nb * KD_NR + nl;
    if (m >= M || n >= M) continue;
    double re = 0.0, im = 0.0;
    for (int c = 0; c < nchunk; c++) {
      const double *pp = part + ((size_t)((mb * nnb + nb) * nchunk + c)) * KD_TILE_D + ((size_t)l * KD_MB + ml) * (2 * KD_NR);
      re += pp[nl];
      im += pp[KD_NR + nl];
    }
    D[((size_t)n * M + m) * NB + l] = make_double2(re, im);
  }
}

static int kdiag_configure() {
  return cudaFuncSetAttribute(k_kubo_diag, cudaFuncAttributeMaxDynamicSharedMemorySize, KD_SMEM_BYTES) == cudaSuccess ? 0 : -3;
}
// site chunks per (left block, right block) pair so that the work items fill the SMs in whole waves
static int kdiag_chunks(int nmb, int nnb, int kk, int sms) {
  const int pairs = nmb * nnb;
  int nchunk = std::max(1, (4 * sms + pairs - 1) / pairs);
  return std::min(nchunk, std::max(1, kk / 64));
}
static size_t kdiag_part_doubles(int nmb, int nnb, int nchunk) { return (size_t)nmb * nnb * nchunk * KD_TILE_D; }
// contracts the right vectors n0 .. n0 + 16*nnb - 1 (padding vectors must be zero) against all left vectors
static int kdiag_launch(const double *left, size_t lstride, int M, const double *right, size_t rstride, int nnb, int kk, int n0,
                        double *part, double2 *D, int sms, cudaStream_t st, long long *launches) {
  const int nmb = (M + KD_MB - 1) / KD_MB, nchunk = kdiag_chunks(nmb, nnb, kk, sms);
  const int nitems = nmb * nnb * nchunk;
  k_kubo_diag<<<std::min(nitems, sms), KD_THREADS, KD_SMEM_BYTES, st>>>(left, lstride, nmb, right, rstride, nnb, kk, nchunk, part);
  const size_t total = (size_t)nmb * nnb * NB * KD_MB * KD_NR;
  k_kdiag_reduce<<<(unsigned)std::min<size_t>((total + 255) / 256, 4096), 256, 0, st>>>(part, nmb, nnb, nchunk, M, n0, D);
  (*launches) += 2;
  return cudaGetLastError() == cudaSuccess ? 0 : -3;
}

// kernels_dmma.cuh -- FP64 tensor-core pipeline (kernel family 1), hand-written for sm_100a.
//
// The 18-column block recursion is FP64-COMPUTE bound on B200 (~51 flop/B against a machine balance of
// ~6 flop/B), so the hot loops run on the FP64 tensor pipe: `mma.sync.m8n8k4.f64` (SASS DMMA.8x8x4), measured at
// 37.1 TFLOP/s on this B200 against 33.6 for plain DFMA (profiles/r01_fp64_peak_microbench.txt).  tcgen05/TMEM
// has no FP64 kind, so DMMA is the Blackwell tensor path for this arithmetic.
//
// k_apply_dmma: persistent, warp-specialised gather-SpMV.  A CTA owns a tile of 8 sites (144 vector columns) of one
// Hamiltonian class.  For every neighbour slot the producer warp issues TMA bulk copies (cp.async.bulk, one 5184 B
// psi block per neighbour + one 10368 B HR36 Hamiltonian block) into a 3-stage shared-memory ring guarded by
// mbarriers; 8 consumer warps compute  C^T[n][r] += Psi[k'][n] * Hreal[r][k']  (M = 144 site-columns, N = 36 -> 40
// output rows, K = 36) with DMMA and keep the accumulators in registers across all slots.  The 90 (m-tile, n-tile)
// units of a stage are split 22/22/23/23 over the four SM sub-partitions (11 or 12 per warp).  The epilogue (scale/shift, three-term
// update) is applied straight from the accumulator fragments with 16 B global accesses; the on-site slot is
// scheduled last so that psi_self is still in shared memory when the epilogue needs it.
//
// k_gram_dmma: D = sum_sites Y^H X (and X^H X) as real DMMA products on the RI36 columns:  Re D(i,j) = <Ycol_i, Xcol_j>,
// Im D(i,j) = <Ycol_i, J Xcol_j>.  One site per warp per step, per-warp 2-stage TMA ring, 25 accumulator tiles per warp.
#pragma once
#include "common.cuh"
#include <algorithm>
#include <cstring>
#include <type_traits>
#include <vector>

#define DM_S 8                               // sites per tile
#define DM_STAGES 4
#define DM_STAGE_D (HBLK + DM_S * BLKD)      // doubles per stage: H block + 8 psi blocks = 6480 (51840 B)
#define DM_CONSUMERS 8
#define DM_THREADS (32 * (DM_CONSUMERS + 1))
#define DM_MAXST 48
#ifndef DM_APPLY_S
#define DM_APPLY_S 4  // default geometry of the SpMV kernel (8: one CTA per SM, 4: two half-tile CTAs per SM)
#endif
#define DM_SMEM_BYTES (DM_STAGES * DM_STAGE_D * 8 + 64)
// The SpMV kernel comes in two geometries (template parameter S = sites per CTA pass):
//   S = 8: one CTA per SM, 8 consumer warps on a whole tile, 4-stage ring of 51.8 kB stages;
//   S = 4: two CTAs per SM, each with 4 consumer warps on HALF a tile (one warp per SM sub-partition), 3-stage ring of
//          31.1 kB stages.  The two warps of a sub-partition then belong to different CTAs and drift out of phase, so
//          stage boundaries and epilogues of one overlap the DMMAs of the other.
template <int S> struct ApGeom {
  static constexpr int kConsumers = S;
  static constexpr int kStages = S == 8 ? 4 : 3;
  static constexpr int kStageD = HBLK + S * BLKD;
  static constexpr int kThreads = 32 * (S + 1);
  static constexpr int kSmem = kStages * kStageD * 8 + 64;
  static constexpr int kSmemGram = kSmem + S * BLKD * 8;   // + the output tile of a pass (operand of the fused Gram products)
  static constexpr int kMinBlocks = S == 8 ? 1 : 2;
};
// Gram products fused into the SpMV kernel (S = 4 geometry): 0 none, 1 = A = sum IN^H OUT (EPI_HOP_GRAM), 2 = D1 = sum IN^H IN and
// D2 = sum OUT^H IN (EPI_CHEB).
template <int EPI> struct EpiTraits {
  static constexpr int kGram = EPI == EPI_CHEB ? 2 : EPI == EPI_HOP_GRAM ? 1 : 0;
  static constexpr bool kCheb = EPI == EPI_CHEB || EPI == EPI_CHEB_NOGRAM;
  static constexpr bool kHop = EPI == EPI_HOP || EPI == EPI_HOP_GRAM;
  static constexpr bool kPrev = kCheb || kHop;             // the epilogue reads `prev`
  static constexpr bool kScale = EPI == EPI_HAM || kCheb;  // (acc - b in)/a
};
#define GR_SLOTS 7  // accumulator tiles a consumer warp owns in the fused Gram products

struct DmmaTiles {
  int ntiles = 0, ng = 0, kk = 0;
  int32_t *d_sites = nullptr;  // [ntiles][8]      site ids (kk = null)
  int32_t *d_cls = nullptr;    // [ntiles][2]   Hamiltonian class of each half tile (sites 0-3 / 4-7)
  int32_t *d_nbr = nullptr;    // [ntiles][ng][8]  neighbour site ids per slot
  std::vector<int32_t> h_sites;  // host copy of d_sites (for the active-region planner)
};

struct DmmaStages {  // stage list of one tile, in execution order (self stage last)
  const double *H[DM_MAXST];
  const double *src[DM_MAXST];
  int hstride[DM_MAXST];  // doubles between classes
  int slot[DM_MAXST];     // neighbour slot, 0 = self
  unsigned char sd[DM_MAXST];  // 1: every block of this stage is spin-diagonal (only the two 9x9 spin blocks are non-zero)
  int n;
};

// ---- PTX wrappers -----------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
// non-blocking phase test: issued a few k-steps before the end of a stage for the NEXT stage's barrier, so that the latency of the
// barrier query (~200 cycles per stage on the critical path of every consumer warp when it is issued at the stage boundary:
// 7.5 % of the consumer samples in profiles/r02h_ncu_apply_dmma_1M.txt sat on that branch) overlaps the remaining DMMAs
__device__ __forceinline__ uint32_t mbar_test(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok;
}
// TMA 1-D bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}



// ---- work list of an SpMV launch -------------------------------------------------------------------------------------------
// All passes (S sites of one tile) of all units form ONE list, walked by the persistent CTAs with stride gridDim.x: pass w
// belongs to unit u iff base_u <= w < base_u + n_u, n_u = (tiles of unit u reachable at this step) * (8 / S).  With several
// units in a batch the CTAs therefore share the passes of every unit (a per-unit loop left all but n_u CTAs idle while a unit's
// active region was small, and rounded the rounds up per unit).  Producer and consumers walk the list with the same cursor.
struct PassCursor {
  int w, u, base, n_u;
  template <int S>
  __device__ __forceinline__ static int count(const int32_t *__restrict__ cnt, int ntiles, int u) { return (cnt ? cnt[u] : ntiles) * (DM_S / S); }
};

// ---- Gram products fused into the SpMV kernel --------------------------------------------------------------------------------
// After the epilogue of a pass the CTA holds, in shared memory and in RI36 layout, the S = 4 site blocks of `in` (the last
// pipeline stage) and of the fresh output tile (gbuf).  The 18x18 reductions of the step are real DMMA products on those
// columns, exactly as in k_gram_dmma (Re D(i,j) = <Ycol_i, Xcol_j>, Im D(i,j) = <Ycol_i, J Xcol_j>):
//   GRAM 2 (chebyshev_recur_ll, recursion.f90:2566-2592): rows 0..17 = IN columns (D1 = IN^H IN), rows 18..35 = OUT columns
//           (D2 = OUT^H IN); right operand IN.  5 x 5 accumulator tiles: warp w owns m-tile w (5 tiles) and tile (4, w);
//           warp 0 also (4, 4).
//   GRAM 1 (hop_b, recursion.f90:1643-1647): rows = IN columns, right operand OUT = H psi.  3 x 5 tiles: warp w < 3 owns
//           (w, 0..3), warp 3 owns (0..2, 4).
// Every tile belongs to exactly one warp of the CTA and is accumulated over all passes of the CTA in a fixed order, so the
// per-CTA partials need no cross-warp reduction and the result is bitwise reproducible.
template <int GRAM>
__device__ __forceinline__ void gram_slot_tile(int w, int s, int &mt, int &nt, bool &valid) {
  if (GRAM == 2) {
    if (s < 5) { mt = w; nt = s; valid = true; }
    else if (s == 5) { mt = 4; nt = w; valid = true; }
    else { mt = 4; nt = 4; valid = (w == 0); }
  } else {
    if (w < 3) { mt = w; nt = s; valid = s < 4; }
    else { mt = s; nt = 4; valid = s < 3; }
  }
}
template <int GRAM, int S, int N>
__device__ __forceinline__ void gram_run(const double *in_tile, const double *out_tile, int warp, int lane,
                                         double (&gacc)[GR_SLOTS][2]) {
  const int g = lane >> 2, q = lane & 3;
  int aoff[N], boff[N];
  bool aout[N], bswap[N];
#pragma unroll
  for (int s = 0; s < N; s++) {
    int mt, nt; bool valid;
    gram_slot_tile<GRAM>(warp, s, mt, nt, valid);
    const int i = mt * 8 + g, j = nt * 8 + g;
    // left operand column: GRAM 2: i < 18 -> IN col i, else OUT col i-18;  GRAM 1: IN col i
    aout[s] = GRAM == 2 && i >= NB;
    aoff[s] = min(aout[s] ? i - NB : i, NB - 1) * COLD;
    bswap[s] = j >= NB;
    boff[s] = min(bswap[s] ? j - NB : j, NB - 1) * COLD;
  }
#pragma unroll 1
  for (int site = 0; site < S; site++) {
    const double *xin = in_tile + site * BLKD, *xout = out_tile + site * BLKD;
    const double *rt = GRAM == 2 ? xin : xout;  // right operand
#pragma unroll
    for (int ks = 0; ks < 9; ks++) {
      const int k = 4 * ks + q;
      const int kj = k < NB ? k + NB : k - NB;   // row of (J X): (J X)[k] = k < 18 ? X[k+18] : -X[k-18]
      double a[N], b[N];
#pragma unroll
      for (int s = 0; s < N; s++) {
        a[s] = (aout[s] ? xout : xin)[aoff[s] + k];
        b[s] = rt[boff[s] + (bswap[s] ? kj : k)];
        if (bswap[s] && k >= NB) b[s] = -b[s];
      }
#pragma unroll
      for (int s = 0; s < N; s++) dmma(gacc[s][0], gacc[s][1], a[s], b[s]);
    }
  }
}
// the slot count is warp-uniform: branch once, so that no DMMA is ever issued predicated-off
template <int GRAM, int S>
__device__ __forceinline__ void gram_accumulate(const double *in_tile, const double *out_tile, int warp, int lane,
                                                double (&gacc)[GR_SLOTS][2]) {
  if (GRAM == 2) {
    if (warp == 0) gram_run<GRAM, S, 7>(in_tile, out_tile, warp, lane, gacc);
    else gram_run<GRAM, S, 6>(in_tile, out_tile, warp, lane, gacc);
  } else {
    if (warp < 3) gram_run<GRAM, S, 4>(in_tile, out_tile, warp, lane, gacc);
    else gram_run<GRAM, S, 3>(in_tile, out_tile, warp, lane, gacc);
  }
}
// per-CTA partials in the layout k_reduce_parts sums: [2][18 x 18 complex, column-major]
template <int GRAM>
__device__ __forceinline__ void gram_flush(double *pp, int warp, int lane, double (&gacc)[GR_SLOTS][2]) {
  const int g = lane >> 2, q = lane & 3;
  constexpr int NS = GRAM == 2 ? 7 : 4;
#pragma unroll
  for (int s = 0; s < NS; s++) {
    int mt, nt; bool valid;
    gram_slot_tile<GRAM>(warp, s, mt, nt, valid);
    if (valid) {
      const int R = mt * 8 + g;
#pragma unroll
      for (int e = 0; e < 2; e++) {
        const int C = nt * 8 + 2 * q + e;
        if (R < 2 * NB && C < 2 * NB && (GRAM == 2 || R < NB))
          pp[(R / NB) * BLKD + 2 * ((R % NB) + NB * (C % NB)) + (C / NB)] = gacc[s][e];
      }
    }
    gacc[s][0] = gacc[s][1] = 0.0;
  }
  // GRAM == 1: matrix 1 of the slot is not produced and k_reduce_parts does not read it (modes 0 / 2 without a second destination)
}
__device__ __forceinline__ void consumer_bar(int nthreads) { asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory"); }

// One consumer warp of k_apply_dmma.  XN = number of units this warp owns in its shared m-tile (compile time so that
// no DMMA is ever issued predicated-off: a predicated-off DMMA still occupies the tensor pipe).
template <int EPI, bool ADDEND, int XN, int S>
__device__ __forceinline__ void dmma_consumer(const ApplyParams &p, const DmmaStages &st,
                                              const int32_t *__restrict__ tile_sites, double *stages, uint64_t *full,
                                              uint64_t *empty, int ntiles, int nunits,
                                              const int32_t *__restrict__ order, const int32_t *__restrict__ cnt,
                                              int warp, int lane, double *gbuf) {
  typedef EpiTraits<EPI> ET;
  constexpr int GRAM = S == 4 ? ET::kGram : 0;
  const int g = lane >> 2, q = lane & 3;
  const int nst = st.n;
  double gacc[GR_SLOTS][2];  // fused Gram accumulators of this warp's tiles (GRAM != 0)
#pragma unroll
  for (int s = 0; s < GR_SLOTS; s++) gacc[s][0] = gacc[s][1] = 0.0;
  const double inv_a = 1.0 / p.a;  // the epilogue multiplies by 1/a (<= 1 ulp from the reference's division)
  constexpr int STG = ApGeom<S>::kStages, STGD = ApGeom<S>::kStageD;
  // S = 8: 18 m-tiles, warps share m-tiles 16/17;  S = 4: 9 m-tiles, the four warps share m-tile 8
  const int mt0 = 2 * warp, mt1 = 2 * warp + 1, mt2 = S == 8 ? 16 + (warp >> 2) : 8;
  const int w4 = warp & 3;
  const int xn0 = (S == 8 && warp == 6) ? 2 : (S == 8 && warp == 7) ? 4 : w4;  // first extra n-tile
  const int xn1 = xn0 + 1;                                  // second one (XN == 2 only)
  int aoff[3], boff[5], xoff[2];
  aoff[0] = HBLK + (mt0 * 8 + g) * COLD + q;
  aoff[1] = HBLK + (mt1 * 8 + g) * COLD + q;
  aoff[2] = HBLK + (mt2 * 8 + g) * COLD + q;
#pragma unroll
  for (int nt = 0; nt < 5; nt++) boff[nt] = min(nt * 8 + g, 35) * COLD + q;
  xoff[0] = min(xn0 * 8 + g, 35) * COLD + q;
  xoff[1] = min(xn1 * 8 + g, 35) * COLD + q;

  uint32_t it = 0, ready = 0;  // ready: the barrier of the upcoming stage has already been seen complete (mbar_test)
  int u = 0, base = 0, n_u = nunits > 0 ? PassCursor::count<S>(cnt, ntiles, 0) : 0;
  for (int w = blockIdx.x;; w += gridDim.x) {
    // units left behind are complete for this CTA: flush their Gram partials (zeros where the CTA had no pass)
    while (u < nunits && w >= base + n_u) {
      if (GRAM) gram_flush<GRAM>(p.part + ((size_t)u * gridDim.x + blockIdx.x) * (2 * BLKD), warp, lane, gacc);
      base += n_u; u++;
      if (u < nunits) n_u = PassCursor::count<S>(cnt, ntiles, u);
    }
    if (u >= nunits) break;
    {
      const size_t uo = (size_t)u * p.vstride;
      const int ti = w - base;
      const int tpos = S == 8 ? ti : ti >> 1, half = S == 8 ? 0 : (ti & 1) * S;
      const int tile = order ? order[(size_t)u * ntiles + tpos] : tpos;
      double acc[2][5][2], xacc[XN][2];
#pragma unroll
      for (int i = 0; i < 2; i++)
#pragma unroll
        for (int nt = 0; nt < 5; nt++) acc[i][nt][0] = acc[i][nt][1] = 0.0;
#pragma unroll
      for (int x = 0; x < XN; x++) xacc[x][0] = xacc[x][1] = 0.0;
      // global element offsets of this lane's accumulator rows (one site-column per m-tile)
      size_t goff[3];
      bool gval[3];
#pragma unroll
      for (int i = 0; i < 3; i++) {
        const int n = (i == 0 ? mt0 : i == 1 ? mt1 : mt2) * 8 + g;
        const int site = tile_sites[tile * DM_S + half + n / NB];
        gval[i] = site < p.kk;
        goff[i] = uo + (size_t)site * BLKD + (n % NB) * COLD + 2 * q;
      }
      double2 pv[2][5], xpv[XN];  // prefetched `prev` (psi0) fragments
      for (int j = 0; j < nst; j++, it++) {
        const int slot = it % STG;
        if (ET::kPrev && j == nst - 1) {
          // issue the epilogue's global loads now; they land while the last stage is being computed
#pragma unroll
          for (int i = 0; i < 2; i++)
#pragma unroll
            for (int nt = 0; nt < 5; nt++)
              pv[i][nt] = (gval[i] && p.prev && (nt < 4 || q < 2)) ? __ldg(reinterpret_cast<const double2 *>(p.prev + goff[i] + nt * 8))
                                                                   : make_double2(0.0, 0.0);  // prev == null: plain W = H in
#pragma unroll
          for (int x = 0; x < XN; x++) {
            const int nt = x == 0 ? xn0 : xn1;
            xpv[x] = (gval[2] && p.prev && (nt < 4 || q < 2)) ? __ldg(reinterpret_cast<const double2 *>(p.prev + goff[2] + nt * 8))
                                                              : make_double2(0.0, 0.0);
          }
        }
        if (!ready) mbar_wait(&full[slot], (it / STG) & 1);
        const double *sm = stages + (size_t)slot * STGD;
        // fragments of k-step ks+1 are loaded before the DMMAs of k-step ks are issued (register double buffering)
        double b[2][5], a[2][3], xb[2][XN];
#pragma unroll
        for (int nt = 0; nt < 5; nt++) b[0][nt] = sm[boff[nt]];
#pragma unroll
        for (int i = 0; i < 3; i++) a[0][i] = sm[aoff[i]];
#pragma unroll
        for (int x = 0; x < XN; x++) xb[0][x] = sm[xoff[x]];
#pragma unroll
        for (int ks = 0; ks < 9; ks++) {
          const int c = ks & 1, n = c ^ 1;
          if (ks == 6) ready = mbar_test(&full[(it + 1) % STG], ((it + 1) / STG) & 1);  // next stage's data: usually there already
          if (ks < 8) {
#pragma unroll
            for (int nt = 0; nt < 5; nt++) b[n][nt] = sm[boff[nt] + 4 * (ks + 1)];
#pragma unroll
            for (int i = 0; i < 3; i++) a[n][i] = sm[aoff[i] + 4 * (ks + 1)];
#pragma unroll
            for (int x = 0; x < XN; x++) xb[n][x] = sm[xoff[x] + 4 * (ks + 1)];
          }
#pragma unroll
          for (int nt = 0; nt < 5; nt++) {
            dmma(acc[0][nt][0], acc[0][nt][1], a[c][0], b[c][nt]);
            dmma(acc[1][nt][0], acc[1][nt][1], a[c][1], b[c][nt]);
          }
#pragma unroll
          for (int x = 0; x < XN; x++) dmma(xacc[x][0], xacc[x][1], a[c][2], xb[c][x]);
        }
        if (j < nst - 1) {
          __syncwarp();
          if (lane == 0) mbar_arrive(&empty[slot]);
        }
      }
      // ===== epilogue from the accumulator fragments; the last stage (self blocks of `in`) is still held =====
      const int lslot = (it - 1) % STG;
      const double *sm = stages + (size_t)lslot * STGD;
      if (GRAM) consumer_bar(32 * S);  // every warp is done reading the previous pass's output tile
      auto finish = [&](double v0, double v1, int n, int nt, size_t go, double2 prev, bool valid) {
        if (ADDEND) {
          const double2 ad = valid ? __ldg(reinterpret_cast<const double2 *>(p.addend + go)) : make_double2(0.0, 0.0);
          v0 += ad.x; v1 += ad.y;
        }
        if (ET::kScale) {
          const double2 in = *reinterpret_cast<const double2 *>(sm + HBLK + n * COLD + nt * 8 + 2 * q);
          v0 = (v0 - p.b * in.x) * inv_a; v1 = (v1 - p.b * in.y) * inv_a;
          if (ET::kCheb) { v0 = 2.0 * v0 - prev.x; v1 = 2.0 * v1 - prev.y; }
        }
        if (ET::kHop) {  // hop_b: hpsi = H psi (kept for A = psi^H hpsi), pmn = hpsi - pmn (recursion.f90:1641)
          if (GRAM == 1) *reinterpret_cast<double2 *>(gbuf + n * COLD + nt * 8 + 2 * q) = make_double2(v0, v1);
          else if (valid) *reinterpret_cast<double2 *>(p.out2 + go) = make_double2(v0, v1);
          v0 -= prev.x; v1 -= prev.y;
        }
        if (GRAM == 2) *reinterpret_cast<double2 *>(gbuf + n * COLD + nt * 8 + 2 * q) = make_double2(v0, v1);
        if (valid) *reinterpret_cast<double2 *>(p.out + go) = make_double2(v0, v1);
      };
#pragma unroll
      for (int i = 0; i < 2; i++) {
        const int n = (i == 0 ? mt0 : mt1) * 8 + g;
#pragma unroll
        for (int nt = 0; nt < 5; nt++)
          if ((GRAM || gval[i]) && (nt < 4 || q < 2)) finish(acc[i][nt][0], acc[i][nt][1], n, nt, goff[i] + nt * 8, pv[i][nt], gval[i]);
      }
#pragma unroll
      for (int x = 0; x < XN; x++) {
        const int nt = x == 0 ? xn0 : xn1;
        if ((GRAM || gval[2]) && (nt < 4 || q < 2)) finish(xacc[x][0], xacc[x][1], mt2 * 8 + g, nt, goff[2] + nt * 8, xpv[x], gval[2]);
      }
      if (GRAM) {
        consumer_bar(32 * S);  // the output tile is complete in shared memory
        gram_accumulate<GRAM, S>(sm + HBLK, gbuf, warp, lane, gacc);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[lslot]);
    }
  }
}

// ---- fused gather-SpMV + epilogue ---------------------------------------------------------------------------
template <int EPI, bool ADDEND, int S>
__global__ void __launch_bounds__(ApGeom<S>::kThreads, ApGeom<S>::kMinBlocks)
k_apply_dmma(ApplyParams p, DmmaStages st, const int32_t *__restrict__ tile_sites, const int32_t *__restrict__ tile_cls,
             const int32_t *__restrict__ tile_nbr, int ntiles, int nunits, const int32_t *__restrict__ order,
             const int32_t *__restrict__ cnt) {
  // order/cnt (optional): per unit, the tiles reachable from the unit's start sites sorted by the step at which they
  // are first reached, and how many of them are reachable at this step -- the reference's izero/irlist bookkeeping
  // (recursion.f90:1629-1636): unreached tiles hold exact zeros and are skipped.
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double *stages = reinterpret_cast<double *>(smem_raw);
  constexpr int STG = ApGeom<S>::kStages, STGD = ApGeom<S>::kStageD, NCONS = ApGeom<S>::kConsumers;
  uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + (size_t)STG * STGD * 8);
  uint64_t *empty = full + STG;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < STG; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], NCONS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int nst = st.n, ng = p.ngather;

  if (warp == NCONS) {
    // ===== producer warp: TMA bulk copies =====
    // The (site index, class) of the next stage is fetched from global memory while the warp waits for the ring slot
    // of the current one, so the index-load latency is off the critical path of the pipeline.
    // cursor over the flattened pass list (PassCursor) and the stage within the pass
    auto settle = [&](PassCursor &c) {
      while (c.u < nunits && c.w >= c.base + c.n_u) {
        c.base += c.n_u; c.u++;
        if (c.u < nunits) c.n_u = PassCursor::count<S>(cnt, ntiles, c.u);
      }
    };
    auto fetch = [&](const PassCursor &c, int j, int &site, int &cls) {
      const int i = c.w - c.base;
      const int tpos = S == 8 ? i : i >> 1, half = S == 8 ? 0 : (i & 1) * S;
      const int tile = order ? order[(size_t)c.u * ntiles + tpos] : tpos;
      cls = tile_cls[2 * tile + (half ? 1 : 0)];
      const int m = st.slot[j];
      site = (lane < S) ? ((m == 0) ? tile_sites[tile * DM_S + half + lane] : tile_nbr[((size_t)tile * ng + m) * DM_S + half + lane]) : 0;
    };
    PassCursor cur{(int)blockIdx.x, 0, 0, nunits > 0 ? PassCursor::count<S>(cnt, ntiles, 0) : 0};
    int cj = 0;
    settle(cur);
    int site = 0, cls = 0;
    if (cur.u < nunits) fetch(cur, cj, site, cls);
    for (uint32_t it = 0; cur.u < nunits; it++) {
      PassCursor nxt = cur;
      int nj = cj + 1, nsite = 0, ncls = 0;
      if (nj == nst) { nj = 0; nxt.w += gridDim.x; settle(nxt); }
      if (nxt.u < nunits) fetch(nxt, nj, nsite, ncls);
      const int slot = it % STG;
      mbar_wait(&empty[slot], ((it / STG) & 1) ^ 1);
      double *sm = stages + (size_t)slot * STGD;
      if (lane == 0) mbar_expect_tx(&full[slot], STGD * 8);
      __syncwarp();
      if (lane < S) {
        bulk_g2s(sm + HBLK + lane * BLKD, st.src[cj] + (size_t)cur.u * p.vstride + (size_t)site * BLKD, BLKD * 8, &full[slot]);
      } else if (lane == S) {
        bulk_g2s(sm, st.H[cj] + (size_t)cls * st.hstride[cj], HBLK * 8, &full[slot]);
      }
      cur = nxt; cj = nj; site = nsite; cls = ncls;
    }
    return;
  }

  // ===== consumer warps: DMMA =====
  // 90 (m-tile, n-tile) units per k-step.  Every warp owns two full m-tiles (10 units) plus 1 or 2 units of the two
  // shared m-tiles 16/17, so the sub-partitions carry 22/22/23/23 units and the two warps of a sub-partition stay
  // within one unit of each other (both keep the tensor pipe fed).
  //   m-tile 16: w0:n0  w1:n1  w2:n2  w3:n3,n4        m-tile 17: w4:n0  w5:n1  w6:n2,n3  w7:n4
  double *gbuf = reinterpret_cast<double *>(smem_raw + ApGeom<S>::kSmem);  // output tile of a pass (Gram variants only)
  if (warp == 3 || (S == 8 && warp == 6))
    dmma_consumer<EPI, ADDEND, 2, S>(p, st, tile_sites, stages, full, empty, ntiles, nunits, order, cnt, warp, lane, gbuf);
  else
    dmma_consumer<EPI, ADDEND, 1, S>(p, st, tile_sites, stages, full, empty, ntiles, nunits, order, cnt, warp, lane, gbuf);
}

// ---- spin-resolved variant of the SpMV (S = 4 geometry) -----------------------------------------------------------------------
// In a collinear calculation the hopping blocks ee(:,:,m,type) are spin-diagonal -- only l.s on the on-site block couples the
// spins -- so the 36x36 real embedding of such a block is two independent 18x18 products: 2 x 3 x 5 = 30 DMMAs per m-tile
// and slot instead of 5 x 9 = 45.  With a third less arithmetic per stage the round-1 form of this kernel (ordinary 10 kB
// HR36 blocks, 3-stage ring of 31 kB stages) waited for its data (DMMA pipe 67 % active, 17 % of the samples in barrier
// waits).  Round 2: every stage carries a HALF block -- the "HS18" packing, 2 spins x 18 rows x 20 k (5760 B; k padded with
// explicit zeros) of either the spin-diagonal part of the slot's block (kind 0: output rows of spin s contract psi rows of
// spin s) or its spin-off-diagonal part (kind 1: rows of spin s contract psi rows of the other spin; only emitted for slots
// that couple the spins, i.e. the on-site slot with l.s) -- so stages are 26.5 kB and the ring is FOUR deep (three in the
// variants that also hold the output tile for the fused Gram products).
// Output rows are grouped by spin, 24 per spin (3 n-tiles, 18 used):
//   up:   rho 0..8 = Re rows 0..8, rho 9 pad, rho 10..18 = Im rows 18..26, rho 19..23 pad
//   down: rho 24 pad, 25..33 = Re rows 9..17, rho 34 pad, 35..43 = Im rows 27..35, rho 44..47 pad
// (the pads are placed so that every accumulator pair (rho, rho+1) of two valid rows is an aligned double2 in RI36).
#define SDH 720   // doubles per HS18 half block: [spin][18 rows][20 k]
template <bool GRAMV> struct SdGeom {
  static constexpr int kStages = GRAMV ? 3 : 4;
  static constexpr int kStageD = SDH + 4 * BLKD;                    // 3312 doubles = 26496 B
  static constexpr int kRing = kStages * kStageD * 8 + 64;
  static constexpr int kSmem = kRing + (GRAMV ? 4 * BLKD * 8 : 0);  // + the output tile of a pass
};
__device__ __forceinline__ int sd_row(int rho) {  // accumulator row -> RI36 row, -1 = padding
  if (rho < 24) {
    if (rho < 9) return rho;
    if (rho == 9) return -1;
    return rho < 19 ? 18 + (rho - 10) : -1;
  }
  const int r = rho - 24;
  if (r == 0 || r == 10) return -1;
  if (r < 10) return 9 + (r - 1);
  return r < 20 ? 27 + (r - 11) : -1;
}
__device__ __forceinline__ int sd_prow(int rho) {  // accumulator row -> row (0..17) of its spin's HS18 sub-block, -1 = padding
  const int r = sd_row(rho);
  if (r < 0) return -1;
  return r < NB ? r % 9 : 9 + (r - NB) % 9;
}
__device__ __forceinline__ int sd_kcol(int spin, int kappa) {  // k index (0..19) of a half stage -> RI36 row of psi
  if (kappa < 9) return 9 * spin + kappa;
  if (kappa < 18) return 18 + 9 * spin + (kappa - 9);
  return 0;  // padding: the HS18 block holds explicit zeros there, any finite psi value will do
}
// HR36 set -> HS18 set: dst[(b*2 + kind)][s][r][kappa] = Hreal_b[row(s, r)][kcol(kind ? 1-s : s, kappa)], zero for kappa >= 18
__global__ void k_pack_hs18(const double *__restrict__ hr36, double *__restrict__ hs18, int nblocks) {
  const int b = blockIdx.x;
  if (b >= nblocks) return;
  for (int e = threadIdx.x; e < 2 * SDH; e += blockDim.x) {
    const int kind = e / SDH, rem = e % SDH, sp = rem / 360, r = (rem % 360) / 20, kap = rem % 20;
    const int row = r < 9 ? 9 * sp + r : 18 + 9 * sp + (r - 9);
    const int ks = kind ? 1 - sp : sp;
    double v = 0.0;
    if (kap < 18) v = hr36[(size_t)b * HBLK + row * COLD + (kap < 9 ? 9 * ks + kap : 18 + 9 * ks + (kap - 9))];
    hs18[(size_t)b * 2 * SDH + e] = v;
  }
}

// XN shared-m-tile units of this warp, all of spin XSPIN: warp 0: n-tiles 0,1  warp 1: 2  warp 2: 3,4  warp 3: 5
template <int EPI, bool ADDEND, int XN, int XSPIN>
__device__ __forceinline__ void dmma_consumer_sd(const ApplyParams &p, const DmmaStages &st,
                                                 const int32_t *__restrict__ tile_sites, double *stages, uint64_t *full,
                                                 uint64_t *empty, int ntiles, int nunits,
                                                 const int32_t *__restrict__ order, const int32_t *__restrict__ cnt,
                                                 int warp, int lane, double *gbuf) {
  typedef EpiTraits<EPI> ET;
  constexpr int GRAM = ET::kGram;
  constexpr int S = 4;
  typedef SdGeom<(GRAM != 0)> G;
  constexpr int STG = G::kStages, STGD = G::kStageD;
  double gacc[GR_SLOTS][2];  // fused Gram accumulators of this warp's tiles (GRAM != 0)
#pragma unroll
  for (int s = 0; s < GR_SLOTS; s++) gacc[s][0] = gacc[s][1] = 0.0;
  const int g = lane >> 2, q = lane & 3;
  const int nst = st.n;
  const double inv_a = 1.0 / p.a;
  const int mt0 = 2 * warp, mt1 = 2 * warp + 1, mt2 = 8;
  const int xn0 = XSPIN * 3 + (warp & 1 ? 2 : 0);  // warp 0: 0,1  warp 1: 2  warp 2: 3,4  warp 3: 5
  // psi (A fragment) addresses.  k-step kt of spin s reads psi row sd_kcol(s, 4 kt + q) = 9 s + q + {0, 4, c2, 21, 25}: per
  // m-tile and spin ONE lane-dependent base (a0) and immediates, except k-step 2 (the Re/Im seam: c2 = 8 for q = 0, 17
  // otherwise) and the two padding lanes of k-step 4 (c4 points them at row 0 of the column; the HS18 block is zero there, the
  // value only has to be finite).  The round-1 form looked the row up in a table: one IMAD/SEL per shared-memory load.
  int a0[3][2];
#pragma unroll
  for (int sp = 0; sp < 2; sp++) {
    a0[0][sp] = SDH + (mt0 * 8 + g) * COLD + 9 * sp + q;
    a0[1][sp] = SDH + (mt1 * 8 + g) * COLD + 9 * sp + q;
    a0[2][sp] = SDH + (mt2 * 8 + g) * COLD + 9 * sp + q;
  }
  const int c2 = q == 0 ? 8 : 17;
  int c4[2];
  c4[0] = q < 2 ? 25 : -q;
  c4[1] = q < 2 ? 25 : -(q + 9);
  int brow[6], xrow[XN];                           // HS18 row base offsets of this lane's B fragments (pads read row 0)
#pragma unroll
  for (int nt = 0; nt < 6; nt++) brow[nt] = ((nt / 3) * NB + max(sd_prow(nt * 8 + g), 0)) * 20 + q;
#pragma unroll
  for (int x = 0; x < XN; x++) xrow[x] = (((xn0 + x) / 3) * NB + max(sd_prow((xn0 + x) * 8 + g), 0)) * 20 + q;
  // epilogue rows of this lane's accumulator pairs: RI36 offsets (or -1)
  int er0[6], er1[6];
#pragma unroll
  for (int nt = 0; nt < 6; nt++) { er0[nt] = sd_row(nt * 8 + 2 * q); er1[nt] = sd_row(nt * 8 + 2 * q + 1); }

  uint32_t it = 0, ready = 0;
  int u = 0, base = 0, n_u = nunits > 0 ? PassCursor::count<S>(cnt, ntiles, 0) : 0;
  for (int w = blockIdx.x;; w += gridDim.x) {
    while (u < nunits && w >= base + n_u) {
      if (GRAM) gram_flush<GRAM>(p.part + ((size_t)u * gridDim.x + blockIdx.x) * (2 * BLKD), warp, lane, gacc);
      base += n_u; u++;
      if (u < nunits) n_u = PassCursor::count<S>(cnt, ntiles, u);
    }
    if (u >= nunits) break;
    {
      const size_t uo = (size_t)u * p.vstride;
      const int ti = w - base;
      const int tpos = ti >> 1, half = (ti & 1) * S;
      const int tile = order ? order[(size_t)u * ntiles + tpos] : tpos;
      double acc[2][6][2], xacc[XN][2];
#pragma unroll
      for (int i = 0; i < 2; i++)
#pragma unroll
        for (int nt = 0; nt < 6; nt++) acc[i][nt][0] = acc[i][nt][1] = 0.0;
#pragma unroll
      for (int x = 0; x < XN; x++) xacc[x][0] = xacc[x][1] = 0.0;
      size_t gbase[3];
      bool gval[3];
      int ncol[3];
#pragma unroll
      for (int i = 0; i < 3; i++) {
        const int n = (i == 0 ? mt0 : i == 1 ? mt1 : mt2) * 8 + g;
        const int site = tile_sites[tile * DM_S + half + n / NB];
        gval[i] = site < p.kk;
        gbase[i] = uo + (size_t)site * BLKD + (n % NB) * COLD;
        ncol[i] = n;
      }
      double pv[2][6][2], xpv[XN][2];  // prefetched `prev` values of this lane's accumulator rows
      auto fetch_prev = [&](double *dst, int nt, size_t gb) {
        const int r0 = er0[nt], r1 = er1[nt];
        dst[0] = dst[1] = 0.0;
        if (r0 >= 0 && r1 >= 0) {
          const double2 t2 = __ldg(reinterpret_cast<const double2 *>(p.prev + gb + r0));
          dst[0] = t2.x; dst[1] = t2.y;
        } else if (r0 >= 0) dst[0] = __ldg(p.prev + gb + r0);
        else if (r1 >= 0) dst[1] = __ldg(p.prev + gb + r1);
      };
      for (int j = 0; j < nst; j++, it++) {
        const int slot = it % STG;
        if (ET::kPrev && j == nst - 1) {
          // issue the epilogue's global loads now; they land while the last stage is being computed
#pragma unroll
          for (int i = 0; i < 2; i++)
#pragma unroll
            for (int nt = 0; nt < 6; nt++)
              if (gval[i] && p.prev) fetch_prev(pv[i][nt], nt, gbase[i]); else pv[i][nt][0] = pv[i][nt][1] = 0.0;
#pragma unroll
          for (int x = 0; x < XN; x++)
            if (gval[2] && p.prev) fetch_prev(xpv[x], xn0 + x, gbase[2]); else xpv[x][0] = xpv[x][1] = 0.0;
        }
        if (!ready) mbar_wait(&full[slot], (it / STG) & 1);
        const double *sm = stages + (size_t)slot * STGD;
        // kind 1 (st.sd[j] != 0): rows of spin s contract the psi rows of the OTHER spin.  Two instantiations of the stage body so
        // that every address stays base + immediate.  Fragments of step t+1 are loaded before the DMMAs of step t are issued.
        auto stage = [&](auto cross_tag) {
          constexpr bool CROSS = decltype(cross_tag)::value;
          double fa[2][3], fb[2][3], fx[2][XN];
          auto aoff = [&](int i, int sp, int kt) {
            const int spa = CROSS ? 1 - sp : sp;
            return a0[i][spa] + (kt < 2 ? 4 * kt : kt == 2 ? c2 : kt == 3 ? 21 : c4[spa]);
          };
          auto load = [&](int t, int buf) {  // t = 5 * spin + k-step
            const int sp = t / 5, kt = t % 5;
            fa[buf][0] = sm[aoff(0, sp, kt)]; fa[buf][1] = sm[aoff(1, sp, kt)];
#pragma unroll
            for (int t3 = 0; t3 < 3; t3++) fb[buf][t3] = sm[brow[3 * sp + t3] + 4 * kt];
            if (sp == XSPIN) {
              fa[buf][2] = sm[aoff(2, sp, kt)];
#pragma unroll
              for (int x = 0; x < XN; x++) fx[buf][x] = sm[xrow[x] + 4 * kt];
            }
          };
          load(0, 0);
#pragma unroll
          for (int t = 0; t < 10; t++) {
            const int c = t & 1, sp = t / 5;
            if (t == 7) ready = mbar_test(&full[(it + 1) % STG], ((it + 1) / STG) & 1);
            if (t < 9) load(t + 1, c ^ 1);
#pragma unroll
            for (int t3 = 0; t3 < 3; t3++) {
              dmma(acc[0][3 * sp + t3][0], acc[0][3 * sp + t3][1], fa[c][0], fb[c][t3]);
              dmma(acc[1][3 * sp + t3][0], acc[1][3 * sp + t3][1], fa[c][1], fb[c][t3]);
            }
            if (sp == XSPIN) {
#pragma unroll
              for (int x = 0; x < XN; x++) dmma(xacc[x][0], xacc[x][1], fa[c][2], fx[c][x]);
            }
          }
        };
        if (st.sd[j]) stage(std::true_type()); else stage(std::false_type());
        if (j < nst - 1) {
          __syncwarp();
          if (lane == 0) mbar_arrive(&empty[slot]);
        }
      }
      // ===== epilogue: the last stage (self blocks of `in`) is still held =====
      const int lslot = (it - 1) % STG;
      const double *sm = stages + (size_t)lslot * STGD;
      if (GRAM) consumer_bar(32 * S);  // every warp is done reading the previous pass's output tile
      auto fin1 = [&](double v, int n, int r, size_t gb, double pr, bool valid) {  // one row
        const size_t go = gb + r;
        if (ADDEND && valid) v += __ldg(p.addend + go);
        if (ET::kScale) {
          v = (v - p.b * sm[SDH + n * COLD + r]) * inv_a;
          if (ET::kCheb) v = 2.0 * v - pr;
        }
        if (ET::kHop) {
          if (GRAM == 1) gbuf[n * COLD + r] = v; else if (valid) p.out2[go] = v;
          v -= pr;
        }
        if (GRAM == 2) gbuf[n * COLD + r] = v;
        if (valid) p.out[go] = v;
      };
      auto fin2 = [&](double v0, double v1, int n, int r, size_t gb, double pr0, double pr1, bool valid) {  // two adjacent rows, r even
        const size_t go = gb + r;
        if (ADDEND && valid) { const double2 ad = __ldg(reinterpret_cast<const double2 *>(p.addend + go)); v0 += ad.x; v1 += ad.y; }
        if (ET::kScale) {
          const double2 in = *reinterpret_cast<const double2 *>(sm + SDH + n * COLD + r);
          v0 = (v0 - p.b * in.x) * inv_a; v1 = (v1 - p.b * in.y) * inv_a;
          if (ET::kCheb) { v0 = 2.0 * v0 - pr0; v1 = 2.0 * v1 - pr1; }
        }
        if (ET::kHop) {
          if (GRAM == 1) *reinterpret_cast<double2 *>(gbuf + n * COLD + r) = make_double2(v0, v1);
          else if (valid) *reinterpret_cast<double2 *>(p.out2 + go) = make_double2(v0, v1);
          v0 -= pr0; v1 -= pr1;
        }
        if (GRAM == 2) *reinterpret_cast<double2 *>(gbuf + n * COLD + r) = make_double2(v0, v1);
        if (valid) *reinterpret_cast<double2 *>(p.out + go) = make_double2(v0, v1);
      };
      auto finish = [&](double v0, double v1, int n, int nt, size_t gb, const double *pr, bool valid) {
        const int r0 = er0[nt], r1 = er1[nt];
        if (r0 >= 0 && r1 >= 0) fin2(v0, v1, n, r0, gb, pr[0], pr[1], valid);
        else if (r0 >= 0) fin1(v0, n, r0, gb, pr[0], valid);
        else if (r1 >= 0) fin1(v1, n, r1, gb, pr[1], valid);
      };
#pragma unroll
      for (int i = 0; i < 2; i++)
#pragma unroll
        for (int nt = 0; nt < 6; nt++)
          if (GRAM || gval[i]) finish(acc[i][nt][0], acc[i][nt][1], ncol[i], nt, gbase[i], pv[i][nt], gval[i]);
#pragma unroll
      for (int x = 0; x < XN; x++)
        if (GRAM || gval[2]) finish(xacc[x][0], xacc[x][1], ncol[2], xn0 + x, gbase[2], xpv[x], gval[2]);
      if (GRAM) {
        consumer_bar(32 * S);  // the output tile is complete in shared memory
        gram_accumulate<GRAM, S>(sm + SDH, gbuf, warp, lane, gacc);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[lslot]);
    }
  }
}

template <int EPI, bool ADDEND>
__global__ void __launch_bounds__(ApGeom<4>::kThreads, ApGeom<4>::kMinBlocks)
k_apply_dmma_sd(ApplyParams p, DmmaStages st, const int32_t *__restrict__ tile_sites, const int32_t *__restrict__ tile_cls,
                const int32_t *__restrict__ tile_nbr, int ntiles, int nunits, const int32_t *__restrict__ order,
                const int32_t *__restrict__ cnt) {
  constexpr int S = 4;
  typedef SdGeom<(EpiTraits<EPI>::kGram != 0)> G;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double *stages = reinterpret_cast<double *>(smem_raw);
  constexpr int STG = G::kStages, STGD = G::kStageD, NCONS = ApGeom<S>::kConsumers;
  uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + (size_t)STG * STGD * 8);
  uint64_t *empty = full + STG;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < STG; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], NCONS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int nst = st.n, ng = p.ngather;
  if (warp == NCONS) {
    // producer warp: the neighbour indices are fetched TWO stages ahead of the copy that needs them
    auto settle = [&](PassCursor &c) {
      while (c.u < nunits && c.w >= c.base + c.n_u) {
        c.base += c.n_u; c.u++;
        if (c.u < nunits) c.n_u = PassCursor::count<S>(cnt, ntiles, c.u);
      }
    };
    auto advance = [&](PassCursor &c, int &j) { if (++j == nst) { j = 0; c.w += gridDim.x; settle(c); } };
    auto fetch = [&](const PassCursor &c, int j, int &site, int &cls) {
      const int i = c.w - c.base;
      const int tpos = i >> 1, half = (i & 1) * S;
      const int tile = order ? order[(size_t)c.u * ntiles + tpos] : tpos;
      cls = tile_cls[2 * tile + (half ? 1 : 0)];
      const int m = st.slot[j];
      site = (lane < S) ? ((m == 0) ? tile_sites[tile * DM_S + half + lane] : tile_nbr[((size_t)tile * ng + m) * DM_S + half + lane]) : 0;
    };
    PassCursor c0{(int)blockIdx.x, 0, 0, nunits > 0 ? PassCursor::count<S>(cnt, ntiles, 0) : 0};  // stage being copied
    int j0 = 0;
    settle(c0);
    PassCursor c1 = c0;                             // one stage ahead
    int j1 = j0;
    int site0 = 0, cls0 = 0, site1 = 0, cls1 = 0;
    if (c0.u < nunits) { fetch(c0, j0, site0, cls0); advance(c1, j1); if (c1.u < nunits) fetch(c1, j1, site1, cls1); }
    for (uint32_t it = 0; c0.u < nunits; it++) {
      PassCursor c2 = c1;                           // two stages ahead
      int j2 = j1, site2 = 0, cls2 = 0;
      if (c2.u < nunits) { advance(c2, j2); if (c2.u < nunits) fetch(c2, j2, site2, cls2); }
      const int slot = it % STG;
      mbar_wait(&empty[slot], ((it / STG) & 1) ^ 1);
      double *sm = stages + (size_t)slot * STGD;
      if (lane == 0) mbar_expect_tx(&full[slot], STGD * 8);
      __syncwarp();
      if (lane < S) {
        bulk_g2s(sm + SDH + lane * BLKD, st.src[j0] + (size_t)c0.u * p.vstride + (size_t)site0 * BLKD, BLKD * 8, &full[slot]);
      } else if (lane == S) {
        bulk_g2s(sm, st.H[j0] + (size_t)cls0 * st.hstride[j0], SDH * 8, &full[slot]);   // HS18 half block of this stage
      }
      c0 = c1; j0 = j1; site0 = site1; cls0 = cls1;
      c1 = c2; j1 = j2; site1 = site2; cls1 = cls2;
    }
    return;
  }
  double *gbuf = reinterpret_cast<double *>(smem_raw + G::kRing);  // output tile of a pass (Gram variants only)
  if (warp == 0) dmma_consumer_sd<EPI, ADDEND, 2, 0>(p, st, tile_sites, stages, full, empty, ntiles, nunits, order, cnt, warp, lane, gbuf);
  else if (warp == 1) dmma_consumer_sd<EPI, ADDEND, 1, 0>(p, st, tile_sites, stages, full, empty, ntiles, nunits, order, cnt, warp, lane, gbuf);
  else if (warp == 2) dmma_consumer_sd<EPI, ADDEND, 2, 1>(p, st, tile_sites, stages, full, empty, ntiles, nunits, order, cnt, warp, lane, gbuf);
  else dmma_consumer_sd<EPI, ADDEND, 1, 1>(p, st, tile_sites, stages, full, empty, ntiles, nunits, order, cnt, warp, lane, gbuf);
}

// ---- A = sum IN^H OUT inside an 8-warp SpMV CTA (EPI_HOP_GRAM) ---------------------------------------------------------------
// Same products as gram_run<1, ...> (3 x 5 accumulator tiles on the RI36 columns of the pass's `in` and output tiles), spread over
// eight warps: tile t = 2 w + s (s < 2) -> (m-tile t / 5, n-tile t % 5); warp 7 owns one tile.  Every tile belongs to one warp and
// is accumulated over the CTA's passes in a fixed order (bitwise reproducible, no cross-warp reduction).
__device__ __forceinline__ void gram8_tile(int w, int s, int &mt, int &nt) { const int t = min(2 * w + s, 14); mt = t / 5; nt = t % 5; }
template <int N>
__device__ __forceinline__ void gram8_run(const double *in_tile, const double *out_tile, int warp, int lane, double (&gacc)[2][2]) {
  const int g = lane >> 2, q = lane & 3;
  int aoff[N], boff[N];
  bool bswap[N];
#pragma unroll
  for (int s = 0; s < N; s++) {
    int mt, nt;
    gram8_tile(warp, s, mt, nt);
    const int i = mt * 8 + g, j = nt * 8 + g;
    aoff[s] = min(i, NB - 1) * COLD;
    bswap[s] = j >= NB;
    boff[s] = min(bswap[s] ? j - NB : j, NB - 1) * COLD;
  }
#pragma unroll 1
  for (int site = 0; site < 4; site++) {
    const double *xin = in_tile + site * BLKD, *xout = out_tile + site * BLKD;
#pragma unroll
    for (int ks = 0; ks < 9; ks++) {
      const int k = 4 * ks + q;
      const int kj = k < NB ? k + NB : k - NB;   // row of (J X): (J X)[k] = k < 18 ? X[k+18] : -X[k-18]
      double a[N], b[N];
#pragma unroll
      for (int s = 0; s < N; s++) {
        a[s] = xin[aoff[s] + k];
        b[s] = xout[boff[s] + (bswap[s] ? kj : k)];
        if (bswap[s] && k >= NB) b[s] = -b[s];
      }
#pragma unroll
      for (int s = 0; s < N; s++) dmma(gacc[s][0], gacc[s][1], a[s], b[s]);
    }
  }
}
__device__ __forceinline__ void gram8_flush(double *pp, int warp, int lane, double (&gacc)[2][2]) {
  const int g = lane >> 2, q = lane & 3;
#pragma unroll
  for (int s = 0; s < 2; s++) {
    if (2 * warp + s < 15) {
      int mt, nt;
      gram8_tile(warp, s, mt, nt);
      const int R = mt * 8 + g;
#pragma unroll
      for (int e = 0; e < 2; e++) {
        const int C = nt * 8 + 2 * q + e;
        if (R < NB && C < 2 * NB) pp[2 * (R + NB * (C % NB)) + (C / NB)] = gacc[s][e];
      }
    }
    gacc[s][0] = gacc[s][1] = 0.0;
  }
  // matrix 1 of the slot is not produced and k_reduce_parts does not read it (modes 0 / 2 without a second destination)
}

// ---- the spin-resolved SpMV with EIGHT consumer warps per CTA ------------------------------------------------------------------
// The 4-warp form above issues two thirds of the full-block kernel's DMMAs in the same number of instructions and takes the same
// time: with 2.45 warps per scheduler a warp issues one instruction per ~11 cycles, so the stage is bound by the non-DMMA
// instruction stream (profiles/r02k_ncu_apply_dmma_sd_config1_collinear.txt).  Here a CTA has 8 consumer warps (4 per scheduler
// with the two CTAs of an SM), each owning ONE m-tile (8 columns) and, for warps 0..5, one (m-tile 8, n-tile w) unit: 6 or 7
// independent accumulators per k-step and both spins in the same k-step, no prefetch of `prev`, <= 96 registers.  EPI_HOP_GRAM
// (hop_b of the Lanczos step: A = sum psi^H H psi from the output tile in shared memory, 3-stage ring) is carried too;
// EPI_CHEB (opt-in) stays with the 4-warp kernel.
#define SD8_CONS 8
#define SD8_THREADS (32 * (SD8_CONS + 1))
template <int EPI, bool ADDEND, int XN, int XSPIN>
__device__ __forceinline__ void dmma_consumer_sd8(const ApplyParams &p, const DmmaStages &st,
                                                  const int32_t *__restrict__ tile_sites, double *stages, uint64_t *full,
                                                  uint64_t *empty, int ntiles, int nunits,
                                                  const int32_t *__restrict__ order, const int32_t *__restrict__ cnt,
                                                  int warp, int lane, double *gbuf) {
  typedef EpiTraits<EPI> ET;
  constexpr int S = 4;
  constexpr int GRAM = ET::kGram;
  static_assert(GRAM != 2, "EPI_CHEB runs on the 4-warp kernel");
  typedef SdGeom<(GRAM != 0)> G;
  double gacc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};  // fused Gram accumulators of this warp's tiles (GRAM == 1)
  constexpr int STG = G::kStages, STGD = G::kStageD;
  const int g = lane >> 2, q = lane & 3;
  const int nst = st.n;
  const double inv_a = 1.0 / p.a;
  const int mt0 = warp, mt2 = 8, xn0 = warp;       // own m-tile; the shared m-tile's n-tile of this warp (XN = 1: warps 0..5)
  int a0[2][2];                                    // [own | shared m-tile][psi spin]: base + q + 9 spin
#pragma unroll
  for (int sp = 0; sp < 2; sp++) {
    a0[0][sp] = SDH + (mt0 * 8 + g) * COLD + 9 * sp + q;
    a0[1][sp] = SDH + (mt2 * 8 + g) * COLD + 9 * sp + q;
  }
  const int c2 = q == 0 ? 8 : 17;
  int c4[2];
  c4[0] = q < 2 ? 25 : -q;
  c4[1] = q < 2 ? 25 : -(q + 9);
  int brow[6];
#pragma unroll
  for (int nt = 0; nt < 6; nt++) brow[nt] = ((nt / 3) * NB + max(sd_prow(nt * 8 + g), 0)) * 20 + q;
  const int xrow = ((xn0 / 3) * NB + max(sd_prow(xn0 * 8 + g), 0)) * 20 + q;

  uint32_t it = 0, ready = 0;
  int u = 0, base = 0, n_u = nunits > 0 ? PassCursor::count<S>(cnt, ntiles, 0) : 0;
  for (int w = blockIdx.x;; w += gridDim.x) {
    while (u < nunits && w >= base + n_u) {
      if (GRAM) gram8_flush(p.part + ((size_t)u * gridDim.x + blockIdx.x) * (2 * BLKD), warp, lane, gacc);
      base += n_u; u++;
      if (u < nunits) n_u = PassCursor::count<S>(cnt, ntiles, u);
    }
    if (u >= nunits) break;
    const size_t uo = (size_t)u * p.vstride;
    const int ti = w - base;
    const int tpos = ti >> 1, half = (ti & 1) * S;
    const int tile = order ? order[(size_t)u * ntiles + tpos] : tpos;
    double acc[6][2], xacc[2] = {0.0, 0.0};
#pragma unroll
    for (int nt = 0; nt < 6; nt++) acc[nt][0] = acc[nt][1] = 0.0;
    for (int j = 0; j < nst; j++, it++) {
      const int slot = it % STG;
      if (!ready) mbar_wait(&full[slot], (it / STG) & 1);
      const double *sm = stages + (size_t)slot * STGD;
      auto stage = [&](auto cross_tag) {
        constexpr bool CROSS = decltype(cross_tag)::value;
        constexpr int XS = CROSS ? 1 - XSPIN : XSPIN;
        auto koff = [&](int kt, int sp) { return kt < 2 ? 4 * kt : kt == 2 ? c2 : kt == 3 ? 21 : c4[sp]; };
#pragma unroll
        for (int kt = 0; kt < 5; kt++) {
          double fa[2], fa2 = 0.0, fb[6], fx = 0.0;
          fa[0] = sm[a0[0][0] + koff(kt, 0)];
          fa[1] = sm[a0[0][1] + koff(kt, 1)];
#pragma unroll
          for (int nt = 0; nt < 6; nt++) fb[nt] = sm[brow[nt] + 4 * kt];
          if (XN) { fa2 = sm[a0[1][XS] + koff(kt, XS)]; fx = sm[xrow + 4 * kt]; }
          if (kt == 3) ready = mbar_test(&full[(it + 1) % STG], ((it + 1) / STG) & 1);
#pragma unroll
          for (int nt = 0; nt < 6; nt++) dmma(acc[nt][0], acc[nt][1], fa[CROSS ? 1 - nt / 3 : nt / 3], fb[nt]);
          if (XN) dmma(xacc[0], xacc[1], fa2, fx);
        }
      };
      if (st.sd[j]) stage(std::true_type()); else stage(std::false_type());
      if (j < nst - 1) {
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[slot]);
      }
    }
    // ===== epilogue: the last stage (self blocks of `in`) is still held =====
    const int lslot = (it - 1) % STG;
    const double *sm = stages + (size_t)lslot * STGD;
    if (GRAM) consumer_bar(32 * SD8_CONS);  // every warp is done reading the previous pass's output tile
    // `valid` is false only in the Gram variant (columns of a null site still fill the output tile in shared memory)
    auto out_pair = [&](double v0, double v1, int n, int nt, size_t gb, bool valid) {
      const int r0 = sd_row(nt * 8 + 2 * q), r1 = sd_row(nt * 8 + 2 * q + 1);
      auto one = [&](double v, int r) {
        const size_t go = gb + r;
        if (ADDEND && valid) v += __ldg(p.addend + go);
        if (ET::kScale) {
          v = (v - p.b * sm[SDH + n * COLD + r]) * inv_a;
          if (ET::kCheb) v = 2.0 * v - __ldg(p.prev + go);
        }
        if (ET::kHop) {
          if (GRAM == 1) gbuf[n * COLD + r] = v; else p.out2[go] = v;
          if (valid && p.prev) v -= __ldg(p.prev + go);
        }
        if (valid) p.out[go] = v;
      };
      if (r0 >= 0 && r1 >= 0) {
        const size_t go = gb + r0;
        if (ADDEND && valid) { const double2 ad = __ldg(reinterpret_cast<const double2 *>(p.addend + go)); v0 += ad.x; v1 += ad.y; }
        if (ET::kScale) {
          const double2 in = *reinterpret_cast<const double2 *>(sm + SDH + n * COLD + r0);
          v0 = (v0 - p.b * in.x) * inv_a; v1 = (v1 - p.b * in.y) * inv_a;
          if (ET::kCheb) { const double2 pr = __ldg(reinterpret_cast<const double2 *>(p.prev + go)); v0 = 2.0 * v0 - pr.x; v1 = 2.0 * v1 - pr.y; }
        }
        if (ET::kHop) {
          if (GRAM == 1) *reinterpret_cast<double2 *>(gbuf + n * COLD + r0) = make_double2(v0, v1);
          else *reinterpret_cast<double2 *>(p.out2 + go) = make_double2(v0, v1);
          if (valid && p.prev) { const double2 pr = __ldg(reinterpret_cast<const double2 *>(p.prev + go)); v0 -= pr.x; v1 -= pr.y; }
        }
        if (valid) *reinterpret_cast<double2 *>(p.out + go) = make_double2(v0, v1);
      } else if (r0 >= 0) one(v0, r0);
      else if (r1 >= 0) one(v1, r1);
    };
    {
      const int n = mt0 * 8 + g;
      const int site = tile_sites[tile * DM_S + half + n / NB];
      if (GRAM || site < p.kk) {
        const size_t gb = uo + (size_t)site * BLKD + (n % NB) * COLD;
#pragma unroll
        for (int nt = 0; nt < 6; nt++) out_pair(acc[nt][0], acc[nt][1], n, nt, gb, site < p.kk);
      }
    }
    if (XN) {
      const int n = mt2 * 8 + g;
      const int site = tile_sites[tile * DM_S + half + n / NB];
      if (GRAM || site < p.kk) out_pair(xacc[0], xacc[1], n, xn0, uo + (size_t)site * BLKD + (n % NB) * COLD, site < p.kk);
    }
    if (GRAM) {
      consumer_bar(32 * SD8_CONS);  // the output tile is complete in shared memory
      if (warp < 7) gram8_run<2>(sm + SDH, gbuf, warp, lane, gacc);
      else gram8_run<1>(sm + SDH, gbuf, warp, lane, gacc);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[lslot]);
  }
}

template <int EPI, bool ADDEND>
__global__ void __launch_bounds__(SD8_THREADS, 2)
k_apply_dmma_sd8(ApplyParams p, DmmaStages st, const int32_t *__restrict__ tile_sites, const int32_t *__restrict__ tile_cls,
                 const int32_t *__restrict__ tile_nbr, int ntiles, int nunits, const int32_t *__restrict__ order,
                 const int32_t *__restrict__ cnt) {
  constexpr int S = 4;
  typedef SdGeom<(EpiTraits<EPI>::kGram != 0)> G;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double *stages = reinterpret_cast<double *>(smem_raw);
  constexpr int STG = G::kStages, STGD = G::kStageD;
  uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + (size_t)STG * STGD * 8);
  uint64_t *empty = full + STG;
  double *gbuf = reinterpret_cast<double *>(smem_raw + G::kRing);  // output tile of a pass (Gram variant only)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < STG; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], SD8_CONS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int nst = st.n, ng = p.ngather;
  if (warp == SD8_CONS) {
    // producer warp (as in k_apply_dmma_sd): indices fetched two stages ahead of the copy
    auto settle = [&](PassCursor &c) {
      while (c.u < nunits && c.w >= c.base + c.n_u) {
        c.base += c.n_u; c.u++;
        if (c.u < nunits) c.n_u = PassCursor::count<S>(cnt, ntiles, c.u);
      }
    };
    auto advance = [&](PassCursor &c, int &j) { if (++j == nst) { j = 0; c.w += gridDim.x; settle(c); } };
    auto fetch = [&](const PassCursor &c, int j, int &site, int &cls) {
      const int i = c.w - c.base;
      const int tpos = i >> 1, half = (i & 1) * S;
      const int tile = order ? order[(size_t)c.u * ntiles + tpos] : tpos;
      cls = tile_cls[2 * tile + (half ? 1 : 0)];
      const int m = st.slot[j];
      site = (lane < S) ? ((m == 0) ? tile_sites[tile * DM_S + half + lane] : tile_nbr[((size_t)tile * ng + m) * DM_S + half + lane]) : 0;
    };
    PassCursor c0{(int)blockIdx.x, 0, 0, nunits > 0 ? PassCursor::count<S>(cnt, ntiles, 0) : 0};
    int j0 = 0;
    settle(c0);
    PassCursor c1 = c0;
    int j1 = j0;
    int site0 = 0, cls0 = 0, site1 = 0, cls1 = 0;
    if (c0.u < nunits) { fetch(c0, j0, site0, cls0); advance(c1, j1); if (c1.u < nunits) fetch(c1, j1, site1, cls1); }
    for (uint32_t it = 0; c0.u < nunits; it++) {
      PassCursor c2 = c1;
      int j2 = j1, site2 = 0, cls2 = 0;
      if (c2.u < nunits) { advance(c2, j2); if (c2.u < nunits) fetch(c2, j2, site2, cls2); }
      const int slot = it % STG;
      mbar_wait(&empty[slot], ((it / STG) & 1) ^ 1);
      double *sm = stages + (size_t)slot * STGD;
      if (lane == 0) mbar_expect_tx(&full[slot], STGD * 8);
      __syncwarp();
      if (lane < S) {
        bulk_g2s(sm + SDH + lane * BLKD, st.src[j0] + (size_t)c0.u * p.vstride + (size_t)site0 * BLKD, BLKD * 8, &full[slot]);
      } else if (lane == S) {
        bulk_g2s(sm, st.H[j0] + (size_t)cls0 * st.hstride[j0], SDH * 8, &full[slot]);
      }
      c0 = c1; j0 = j1; site0 = site1; cls0 = cls1;
      c1 = c2; j1 = j2; site1 = site2; cls1 = cls2;
    }
    return;
  }
  if (warp < 3) dmma_consumer_sd8<EPI, ADDEND, 1, 0>(p, st, tile_sites, stages, full, empty, ntiles, nunits, order, cnt, warp, lane, gbuf);
  else if (warp < 6) dmma_consumer_sd8<EPI, ADDEND, 1, 1>(p, st, tile_sites, stages, full, empty, ntiles, nunits, order, cnt, warp, lane, gbuf);
  else dmma_consumer_sd8<EPI, ADDEND, 0, 0>(p, st, tile_sites, stages, full, empty, ntiles, nunits, order, cnt, warp, lane, gbuf);
}

// ---- Gram reductions on the tensor pipe -------------------------------------------------------------------------
#define GR_WARPS 8
#define GR_THREADS (32 * GR_WARPS)
#define GR_STAGE_D (2 * BLKD)   // X block + Y block
#define GR_SMEM_BYTES (GR_WARPS * 2 * GR_STAGE_D * 8 + GR_WARPS * 2 * 8)

// two != 0:  part[.][0] = sum X^H X, part[.][1] = sum Y^H X;   two == 0:  part[.][0] = sum Y^H X, part[.][1] = 0.
// grid = (ctas, nunits); partials in complex column-major (i + 18 j) like the host arrays.
__global__ void __launch_bounds__(GR_THREADS, 1)
k_gram_dmma(const double *__restrict__ X, const double *__restrict__ Y, int two, int kk, size_t xstride, size_t ystride,
            double *part, const int32_t *__restrict__ border, const int32_t *__restrict__ bcnt, int nblocks) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double *sbase = reinterpret_cast<double *>(smem_raw);
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + (size_t)GR_WARPS * 2 * GR_STAGE_D * 8);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, q = lane & 3;
  const int unit = blockIdx.y;
  const double *Xu = X + (size_t)unit * xstride, *Yu = Y + (size_t)unit * ystride;
  double *wsm = sbase + (size_t)warp * 2 * GR_STAGE_D;
  uint64_t *wbar = bars + warp * 2;
  if (lane == 0) { mbar_init(&wbar[0], 1); mbar_init(&wbar[1], 1); }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();

  // fragment source offsets (doubles, relative to the stage base: X block at 0, Y block at BLKD)
  const int nmt = two ? 5 : 3;
  int aoff[5], boff[5];
  bool bswap[5];
#pragma unroll
  for (int mt = 0; mt < 5; mt++) {
    const int i = mt * 8 + g;
    if (two) aoff[mt] = (i < NB) ? i * COLD : BLKD + min(i - NB, NB - 1) * COLD;
    else aoff[mt] = BLKD + min(i, NB - 1) * COLD;
  }
#pragma unroll
  for (int nt = 0; nt < 5; nt++) {
    const int j = nt * 8 + g;
    bswap[nt] = j >= NB;
    boff[nt] = (j < NB ? j : min(j - NB, NB - 1)) * COLD;
  }
  double acc[5][5][2];
#pragma unroll
  for (int mt = 0; mt < 5; mt++)
#pragma unroll
    for (int nt = 0; nt < 5; nt++) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;

  const int stride = gridDim.x * GR_WARPS;
  const int s0 = blockIdx.x * GR_WARPS + warp;
  // border/bcnt (optional): the contiguous 8-site blocks that can be non-zero at this step (active-region plan);
  // positions past kk in the last block map to the null site kk, whose block is always zero
  const int nact = border ? bcnt[unit] * DM_S : kk;
  const int32_t *bo = border ? border + (size_t)unit * nblocks : nullptr;
  auto site_of = [&](int idx) { return bo ? min(bo[idx >> 3] * DM_S + (idx & 7), kk) : idx; };
  // prologue: up to two sites in flight
  for (int pf = 0; pf < 2; pf++) {
    const int idx = s0 + pf * stride;
    const int site = idx < nact ? site_of(idx) : 0;
    if (idx < nact && lane == 0) {
      mbar_expect_tx(&wbar[pf], GR_STAGE_D * 8);
      bulk_g2s(wsm + (size_t)pf * GR_STAGE_D, Xu + (size_t)site * BLKD, BLKD * 8, &wbar[pf]);
      bulk_g2s(wsm + (size_t)pf * GR_STAGE_D + BLKD, Yu + (size_t)site * BLKD, BLKD * 8, &wbar[pf]);
    }
  }
  int itn = 0;
  for (int idx = s0; idx < nact; idx += stride, itn++) {
    const int slot = itn & 1;
    mbar_wait(&wbar[slot], (itn >> 1) & 1);
    const double *sm = wsm + (size_t)slot * GR_STAGE_D;
#pragma unroll
    for (int ks = 0; ks < 9; ks++) {
      const int k = 4 * ks + q;
      double b[5];
#pragma unroll
      for (int nt = 0; nt < 5; nt++) {
        // Re part: Xcol_j[k'];  Im part: (J X)col_j[k'] = k'<18 ? Xcol[k'+18] : -Xcol[k'-18]
        const int kk2 = bswap[nt] ? (k < NB ? k + NB : k - NB) : k;
        double v = sm[boff[nt] + kk2];
        if (bswap[nt] && k >= NB) v = -v;
        b[nt] = v;
      }
#pragma unroll
      for (int mt = 0; mt < 5; mt++) {
        if (mt >= nmt) continue;
        const double a = sm[aoff[mt] + k];
#pragma unroll
        for (int nt = 0; nt < 5; nt++) dmma(acc[mt][nt][0], acc[mt][nt][1], a, b[nt]);
      }
    }
    __syncwarp();
    const int nidx = idx + 2 * stride;
    const int nsite = nidx < nact ? site_of(nidx) : 0;
    if (nidx < nact && lane == 0) {
      mbar_expect_tx(&wbar[slot], GR_STAGE_D * 8);
      bulk_g2s(wsm + (size_t)slot * GR_STAGE_D, Xu + (size_t)nsite * BLKD, BLKD * 8, &wbar[slot]);
      bulk_g2s(wsm + (size_t)slot * GR_STAGE_D + BLKD, Yu + (size_t)nsite * BLKD, BLKD * 8, &wbar[slot]);
    }
  }
  // fixed-order cross-warp reduction through shared memory (reusing the staging area): red[warp][40][40]
  __syncthreads();
  double *red = sbase;  // 8 * 1600 doubles = 102400 B <= staging size
#pragma unroll
  for (int mt = 0; mt < 5; mt++)
#pragma unroll
    for (int nt = 0; nt < 5; nt++) {
      red[(size_t)warp * 1600 + (mt * 8 + g) * 40 + nt * 8 + 2 * q] = acc[mt][nt][0];
      red[(size_t)warp * 1600 + (mt * 8 + g) * 40 + nt * 8 + 2 * q + 1] = acc[mt][nt][1];
    }
  __syncthreads();
  double *pp = part + ((size_t)unit * gridDim.x + blockIdx.x) * (2 * BLKD);
  for (int e = tid; e < 2 * BLKD; e += GR_THREADS) {
    const int which = e / BLKD, idx = e % BLKD, ce = idx >> 1, im = idx & 1, i = ce % NB, j = ce / NB;
    double s = 0.0;
    if (two || which == 0) {
      const int row = (two ? which * NB : 0) + i, col = im * NB + j;
      for (int w = 0; w < GR_WARPS; w++) s += red[(size_t)w * 1600 + row * 40 + col];
    }
    pp[e] = s;
  }
}

// ---- right-multiplications of crecal_b on the tensor pipe ------------------------------------------------------
// X(18,18,site) * M(18x18) for every site is, in RI36/real form,  OUT[(s,k)][c'] = sum_j' Xhat_s[k][j'] Mhat[j'][c']
// with Xhat = [Xre | Xim] (18x36) and Mhat = [[Mre, Mim],[-Mim, Mre]] (36x36): M = 144 rows per 8-site tile (exact),
// N = 36 -> 40, K = 36 -- the same 90-unit shape as one SpMV stage, with the tile read "transposed" from shared
// memory.  These passes are HBM-bound (2.5 flop/B), so tiles are contiguous 41 kB TMA bulk copies, 2-stage ring.
//   RM_ORTHO : pmn <- pmn - psi*A                     (crecal_b 1927; pmn already holds hpsi - pmn_old from EPI_HOP)
//   RM_ROTATE: psi <- pmn*Binv ; pmn <- psi_old*B     (crecal_b 1966-1967)
enum RmulMode { RM_ORTHO = 0, RM_ROTATE = 1 };
#define RM_TILE_D (DM_S * BLKD)                       // 5184 doubles = 41472 B
#define RM_SMEM_BYTES (2 * 2 * RM_TILE_D * 8 + 2 * HBLK * 8 + 64)

template <int XN>
__device__ __forceinline__ void rmul_product(const double *xs, const double *ts, const int (&aoff)[3], const int (&koff)[9],
                                             const int (&boff)[5], const int (&xoff)[2], double (&acc)[2][5][2],
                                             double (&xacc)[2][2]) {
#pragma unroll
  for (int i = 0; i < 2; i++)
#pragma unroll
    for (int nt = 0; nt < 5; nt++) acc[i][nt][0] = acc[i][nt][1] = 0.0;
#pragma unroll
  for (int x = 0; x < 2; x++) xacc[x][0] = xacc[x][1] = 0.0;
#pragma unroll
  for (int ks = 0; ks < 9; ks++) {
    double b[5], xb[2];
#pragma unroll
    for (int nt = 0; nt < 5; nt++) b[nt] = ts[boff[nt] + 4 * ks];
#pragma unroll
    for (int x = 0; x < XN; x++) xb[x] = ts[xoff[x] + 4 * ks];
    const double a0 = xs[aoff[0] + koff[ks]], a1 = xs[aoff[1] + koff[ks]], a2 = xs[aoff[2] + koff[ks]];
#pragma unroll
    for (int nt = 0; nt < 5; nt++) {
      dmma(acc[0][nt][0], acc[0][nt][1], a0, b[nt]);
      dmma(acc[1][nt][0], acc[1][nt][1], a1, b[nt]);
    }
#pragma unroll
    for (int x = 0; x < XN; x++) dmma(xacc[x][0], xacc[x][1], a2, xb[x]);
  }
}


// B^2 = sum pmn^H pmn of crecal_b (recursion.f90:1929-1934) inside the ORTHO pass: the freshly orthogonalised tile is written back
// into its shared-memory slot (RI36, in place of the addend it replaces) and its 3 x 5 accumulator tiles are spread over the 8
// consumer warps (tile t = warp + 8 s).  NT = tiles of this warp (2, or 1 for warp 7); same operand forms as k_gram_dmma.
template <int NT>
__device__ __forceinline__ void rmul_gram(const double *x, int ns, int warp, int lane, double (&gacc)[2][2]) {
  const int g = lane >> 2, q = lane & 3;
  int aoff[NT], boff[NT];
  bool bswap[NT];
#pragma unroll
  for (int s = 0; s < NT; s++) {
    const int t = warp + 8 * s, i = (t / 5) * 8 + g, j = (t % 5) * 8 + g;
    aoff[s] = min(i, NB - 1) * COLD;
    bswap[s] = j >= NB;
    boff[s] = min(bswap[s] ? j - NB : j, NB - 1) * COLD;
  }
#pragma unroll 1
  for (int site = 0; site < ns; site++) {
    const double *xs = x + site * BLKD;
#pragma unroll
    for (int ks = 0; ks < 9; ks++) {
      const int k = 4 * ks + q, kj = k < NB ? k + NB : k - NB;
      double a[NT], b[NT];
#pragma unroll
      for (int s = 0; s < NT; s++) {
        a[s] = xs[aoff[s] + k];
        b[s] = xs[boff[s] + (bswap[s] ? kj : k)];
        if (bswap[s] && k >= NB) b[s] = -b[s];
      }
#pragma unroll
      for (int s = 0; s < NT; s++) dmma(gacc[s][0], gacc[s][1], a[s], b[s]);
    }
  }
}

template <int MODE, int XN>
__device__ __forceinline__ void rmul_consumer(double *psi, double *pmn, const double *hpsi, int kk, double *tiles,
                                              const double *tmat, uint64_t *full, uint64_t *empty, int warp, int lane,
                                              const int32_t *bo, int nact, double *gpart) {
  const int g = lane >> 2, q = lane & 3;
  double gacc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};  // fused B^2 partial tiles of this warp (ORTHO with gpart != null)
  const int mt[3] = {2 * warp, 2 * warp + 1, 16 + (warp >> 2)};
  const int w4 = warp & 3;
  const int xn[2] = {(warp == 6) ? 2 : (warp == 7) ? 4 : w4, ((warp == 6) ? 2 : (warp == 7) ? 4 : w4) + 1};
  int aoff[3], koff[9], boff[5], xoff[2], rows[3], rowk[3];
#pragma unroll
  for (int i = 0; i < 3; i++) {
    const int n = mt[i] * 8 + g;
    rows[i] = n / NB; rowk[i] = n % NB;
    aoff[i] = rows[i] * BLKD + rowk[i];
  }
#pragma unroll
  for (int ks = 0; ks < 9; ks++) { const int j = 4 * ks + q; koff[ks] = (j % NB) * COLD + (j / NB) * NB; }
#pragma unroll
  for (int nt = 0; nt < 5; nt++) boff[nt] = min(nt * 8 + g, 35) * COLD + q;
#pragma unroll
  for (int x = 0; x < 2; x++) xoff[x] = min(xn[x] * 8 + g, 35) * COLD + q;
  uint32_t it = 0;
  for (int ti = blockIdx.x; ti < nact; ti += gridDim.x, it++) {
    const int slot = it & 1;
    const double *sm = tiles + (size_t)slot * 2 * RM_TILE_D;
    const int site0 = (bo ? bo[ti] : ti) * DM_S;
    // output element (row n = (s,k), column c') -> RI36 offset
    auto gofs = [&](int i, int c) { return (size_t)(site0 + rows[i]) * BLKD + (c % NB) * COLD + (c / NB) * NB + rowk[i]; };
    auto valid = [&](int i, int c) { return c < 2 * NB && site0 + rows[i] < kk; };
    double acc[2][5][2], xacc[2][2];
    if (MODE == RM_ORTHO) {
      // tile 0 = psi, tile 1 = pmn (the addend, read from shared memory in the accumulator-fragment pattern)
      mbar_wait(&full[slot], (it >> 1) & 1);
      rmul_product<XN>(sm, tmat, aoff, koff, boff, xoff, acc, xacc);  // psi * (-A)
      const double *ad = sm + RM_TILE_D;
      auto sofs = [&](int i, int c) { return rows[i] * BLKD + (c % NB) * COLD + (c / NB) * NB + rowk[i]; };
#pragma unroll
      for (int i = 0; i < 2; i++)
#pragma unroll
        for (int nt = 0; nt < 5; nt++)
#pragma unroll
          for (int e = 0; e < 2; e++) {
            const int c = nt * 8 + 2 * q + e;
            if (valid(i, c)) acc[i][nt][e] += ad[sofs(i, c)];
          }
#pragma unroll
      for (int x = 0; x < XN; x++)
#pragma unroll
        for (int e = 0; e < 2; e++) {
          const int c = xn[x] * 8 + 2 * q + e;
          if (valid(2, c)) xacc[x][e] += ad[sofs(2, c)];
        }
      if (!gpart) {
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[slot]);
      }
      double *adw = const_cast<double *>(ad);
#pragma unroll
      for (int i = 0; i < 2; i++)
#pragma unroll
        for (int nt = 0; nt < 5; nt++)
#pragma unroll
          for (int e = 0; e < 2; e++) {
            const int c = nt * 8 + 2 * q + e;
            if (valid(i, c)) { pmn[gofs(i, c)] = acc[i][nt][e]; if (gpart) adw[sofs(i, c)] = acc[i][nt][e]; }
          }
#pragma unroll
      for (int x = 0; x < XN; x++)
#pragma unroll
        for (int e = 0; e < 2; e++) {
          const int c = xn[x] * 8 + 2 * q + e;
          if (valid(2, c)) { pmn[gofs(2, c)] = xacc[x][e]; if (gpart) adw[sofs(2, c)] = xacc[x][e]; }
        }
      if (gpart) {
        asm volatile("bar.sync 1, %0;" ::"r"(32 * DM_CONSUMERS) : "memory");  // the orthogonalised tile is complete in shared memory
        const int ns = min(DM_S, kk - site0);
        if (warp + 8 < 15) rmul_gram<2>(ad, ns, warp, lane, gacc); else rmul_gram<1>(ad, ns, warp, lane, gacc);
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[slot]);
      }
    } else {
      mbar_wait(&full[slot], (it >> 1) & 1);
      // tile 0 = pmn, tile 1 = psi;  tmat 0 = Binv, tmat 1 = B
#pragma unroll
      for (int pass = 0; pass < 2; pass++) {
        rmul_product<XN>(sm + (size_t)pass * RM_TILE_D, tmat + (size_t)pass * HBLK, aoff, koff, boff, xoff, acc, xacc);
        if (pass == 1) {
          __syncwarp();
          if (lane == 0) mbar_arrive(&empty[slot]);
        }
        double *out = pass == 0 ? psi : pmn;
#pragma unroll
        for (int i = 0; i < 2; i++)
#pragma unroll
          for (int nt = 0; nt < 5; nt++)
#pragma unroll
            for (int e = 0; e < 2; e++) {
              const int c = nt * 8 + 2 * q + e;
              if (valid(i, c)) out[gofs(i, c)] = acc[i][nt][e];
            }
#pragma unroll
        for (int x = 0; x < XN; x++)
#pragma unroll
          for (int e = 0; e < 2; e++) {
            const int c = xn[x] * 8 + 2 * q + e;
            if (valid(2, c)) out[gofs(2, c)] = xacc[x][e];
          }
      }
    }
  }
  if (MODE == RM_ORTHO && gpart) {  // this warp's tiles of the per-CTA partial of B^2 (matrix 0 of the slot, complex column-major)
#pragma unroll
    for (int s = 0; s < 2; s++) {
      const int t = warp + 8 * s;
      if (t < 15) {
        const int R = (t / 5) * 8 + g;
#pragma unroll
        for (int e = 0; e < 2; e++) {
          const int C = (t % 5) * 8 + 2 * q + e;
          if (R < NB && C < 2 * NB) gpart[2 * (R + NB * (C % NB)) + (C / NB)] = gacc[s][e];
        }
      }
    }
  }
}

// grid = (ctas, nunits).  m0/m1: complex column-major 18x18 per unit (stride mstride doubles): ORTHO m0 = A;
// ROTATE m0 = Binv, m1 = B.
template <int MODE>
__global__ void __launch_bounds__(DM_THREADS, 1)
k_rmul_dmma(double *psi_all, double *pmn_all, const double *hpsi_all, const double *m0, const double *m1, size_t mstride,
            int kk, size_t vstride, const int32_t *__restrict__ border, const int32_t *__restrict__ bcnt, int nblocks,
            double *part) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double *tiles = reinterpret_cast<double *>(smem_raw);                       // [2 slots][2 tiles]
  double *tmat = tiles + 2 * 2 * RM_TILE_D;                                   // [2][36x36]: T[c'][j'] = Mhat[j'][c']
  uint64_t *full = reinterpret_cast<uint64_t *>(tmat + 2 * HBLK);
  uint64_t *empty = full + 2;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, unit = blockIdx.y;
  double *psi = psi_all + (size_t)unit * vstride, *pmn = pmn_all + (size_t)unit * vstride;
  const double *hpsi = hpsi_all ? hpsi_all + (size_t)unit * vstride : nullptr;
  if (tid == 0) {
    for (int s = 0; s < 2; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], DM_CONSUMERS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // Mhat^T for the one or two matrices: Mhat = [[Mre, Mim],[-Mim, Mre]]; ORTHO multiplies by -A
  for (int e = tid; e < 2 * HBLK; e += DM_THREADS) {
    const int w = e / HBLK, r = e % HBLK, c = r / COLD, j = r % COLD;  // T[c'][j']
    double v = 0.0;
    if (w == 0 || MODE == RM_ROTATE) {
      const double *m = (w == 0 ? m0 : m1) + (size_t)unit * mstride;
      const double re = m[2 * ((j % NB) + NB * (c % NB))], im = m[2 * ((j % NB) + NB * (c % NB)) + 1];
      v = (j < NB) == (c < NB) ? re : (j < NB ? im : -im);
      if (MODE == RM_ORTHO) v = -v;
    }
    tmat[e] = v;
  }
  __syncthreads();
  // border/bcnt (optional): contiguous 8-site blocks that can be non-zero at this step (active-region plan)
  const int nact = border ? bcnt[unit] : (kk + DM_S - 1) / DM_S;
  const int32_t *bo = border ? border + (size_t)unit * nblocks : nullptr;
  if (warp == DM_CONSUMERS) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int ti = blockIdx.x; ti < nact; ti += gridDim.x, it++) {
        const int slot = it & 1;
        mbar_wait(&empty[slot], ((it >> 1) & 1) ^ 1);
        const int site0 = (bo ? bo[ti] : ti) * DM_S, ns = min(DM_S, kk - site0);
        const uint32_t bytes = (uint32_t)ns * BLKD * 8;
        double *sm = tiles + (size_t)slot * 2 * RM_TILE_D;
        if (MODE == RM_ORTHO) {
          mbar_expect_tx(&full[slot], 2 * bytes);
          bulk_g2s(sm, psi + (size_t)site0 * BLKD, bytes, &full[slot]);
          bulk_g2s(sm + RM_TILE_D, pmn + (size_t)site0 * BLKD, bytes, &full[slot]);
        } else {
          mbar_expect_tx(&full[slot], 2 * bytes);
          bulk_g2s(sm, pmn + (size_t)site0 * BLKD, bytes, &full[slot]);
          bulk_g2s(sm + RM_TILE_D, psi + (size_t)site0 * BLKD, bytes, &full[slot]);
        }
      }
    }
    return;
  }
  // ORTHO with part != null: B^2 = sum pmn^H pmn partials of this CTA go to slot [unit][blockIdx.x] (k_lz_eig sums them)
  double *gpart = (MODE == RM_ORTHO && part) ? part + ((size_t)unit * gridDim.x + blockIdx.x) * (2 * BLKD) : nullptr;
  if (warp == 3 || warp == 6)
    rmul_consumer<MODE, 2>(psi, pmn, hpsi, kk, tiles, tmat, full, empty, warp, lane, bo, nact, gpart);
  else
    rmul_consumer<MODE, 1>(psi, pmn, hpsi, kk, tiles, tmat, full, empty, warp, lane, bo, nact, gpart);
}

// ---- pipelined block Lanczos: rotate + orthogonalise in ONE pass ---------------------------------------------------------------
// crecal_b (recursion.f90:1873-1973) computes, per step,  hpsi = H psi_n ; R = hpsi - psi_{n-1} B_n - psi_n A_n ; B_{n+1}^2 = R^H R ;
// B_{n+1} = (B^2)^1/2 ; psi_{n+1} = R B^-1 -- a chain in which the 18x18 square root (a single-CTA kernel, ~25 us) sits between two
// passes over the vectors.  Right-multiplication by B^-1 commutes with H, so the driver applies H to the UNNORMALISED residual,
//     W = H R_n ,  G = R_n^H W          (the SpMV kernel with its fused Gram product, running WHILE k_lz_eig forms B, B^-1 on a
//                                         second stream)
// and everything else of the step is right-multiplications by 18x18 matrices known once B^-1 and G are:
//     psi_{n+1} = R_n B_{n+1}^-1 ,   A_{n+1} = psi_{n+1}^H H psi_{n+1} = B^-1 G B^-1 ,
//     R_{n+1}   = W B^-1 - psi_n B_{n+1} - R_n (B^-1 A_{n+1}) ,   B_{n+2}^2 = sum R_{n+1}^H R_{n+1}.
// The normalised vectors are never needed as such: psi_n = R_{n-1} B_n^-1 (R_{-1} = the start vector, B_0 = 1), so the state of
// the recursion is the last TWO residuals and
//     R_{n+1} = W B_{n+1}^-1  -  R_{n-1} (B_n^-1 B_{n+1})  -  R_n (B_{n+1}^-1 A_{n+1})
// is three right-multiplications per site, written over R_{n-1}.  k_rotortho_dmma does that in one pass: three launches on the
// critical path of a step (SpMV, reduce, this) instead of five, one pass over the vectors instead of two, and the square root
// off the critical path.  Same algebra as the reference, different association of the products (results agree to rounding,
// parity tolerance 1e-10).
// Geometry: one CTA per SM, 8 consumer warps + TMA producer, tiles of FOUR sites (three input tiles R_n, W, R_{n-1} per ring
// slot: 62 kB, two slots); warp w owns m-tile w of the 9 (72 rows) and, for w < 5, unit (m-tile 8, n-tile w).
#define RO_S 4
#define RO_TILE_D (RO_S * BLKD)                                       // 2592 doubles = 20736 B
#define RO_SMEM_BYTES (2 * 3 * RO_TILE_D * 8 + 3 * HBLK * 8 + 8 * BLKD * 8 + 64)   // tiles, three matrices, prologue scratch: 197 056 B

// acc (+)= X * M for this warp's units: X tile in shared memory (RI36), ts = transposed real embedding of M
template <int XN, bool ACCUM>
__device__ __forceinline__ void ro_product(const double *xs, const double *ts, int aoff0, int aoff2, const int (&koff)[9],
                                           const int (&boff)[5], int xoff, double (&acc)[5][2], double (&xacc)[2]) {
  if (!ACCUM) {
#pragma unroll
    for (int nt = 0; nt < 5; nt++) acc[nt][0] = acc[nt][1] = 0.0;
    xacc[0] = xacc[1] = 0.0;
  }
#pragma unroll
  for (int ks = 0; ks < 9; ks++) {
    double b[5], xb = 0.0, a2 = 0.0;
#pragma unroll
    for (int nt = 0; nt < 5; nt++) b[nt] = ts[boff[nt] + 4 * ks];
    const double a0 = xs[aoff0 + koff[ks]];
    if (XN) { xb = ts[xoff + 4 * ks]; a2 = xs[aoff2 + koff[ks]]; }
#pragma unroll
    for (int nt = 0; nt < 5; nt++) dmma(acc[nt][0], acc[nt][1], a0, b[nt]);
    if (XN) dmma(xacc[0], xacc[1], a2, xb);
  }
}

template <int XN>
__device__ __forceinline__ void rotortho_consumer(double *rnew, int kk, double *tiles, const double *tmat, double *outbuf,
                                                  uint64_t *full, uint64_t *empty, int warp, int lane, const int32_t *bo,
                                                  int nact, double *gpart) {
  const int g = lane >> 2, q = lane & 3;
  double gacc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
  const int mt[2] = {warp, 8};
  const int xn = warp;  // n-tile of the shared m-tile (XN = 1: warps 0..4)
  int aoff[2], rows[2], rowk[2], koff[9], boff[5];
#pragma unroll
  for (int i = 0; i < 2; i++) {
    const int n = mt[i] * 8 + g;
    rows[i] = n / NB; rowk[i] = n % NB;
    aoff[i] = rows[i] * BLKD + rowk[i];
  }
#pragma unroll
  for (int ks = 0; ks < 9; ks++) { const int j = 4 * ks + q; koff[ks] = (j % NB) * COLD + (j / NB) * NB; }
#pragma unroll
  for (int nt = 0; nt < 5; nt++) boff[nt] = min(nt * 8 + g, 35) * COLD + q;
  const int xoff = min(xn * 8 + g, 35) * COLD + q;
  const double *M1 = tmat, *M2 = tmat + HBLK, *M3 = tmat + 2 * HBLK;  // B^-1, -B_prev^-1 B, -B^-1 A
  uint32_t it = 0;
  for (int ti = blockIdx.x; ti < 2 * nact; ti += gridDim.x) {
    const int blk = ti >> 1;
    const int site0 = (bo ? bo[blk] : blk) * DM_S + (ti & 1) * RO_S;
    if (site0 >= kk) continue;  // second half of a last, partial block (the producer skips it too)
    const int slot = it & 1;
    const double *tC = tiles + (size_t)slot * 3 * RO_TILE_D, *tW = tC + RO_TILE_D, *tP = tC + 2 * RO_TILE_D;   // R_n, W = H R_n, R_{n-1}
    double *tN = outbuf + (size_t)slot * RO_TILE_D;   // R_{n+1} of this tile (the prologue's scratch area: dead by now)
    auto gofs = [&](int i, int c) { return (size_t)(site0 + rows[i]) * BLKD + (c % NB) * COLD + (c / NB) * NB + rowk[i]; };
    auto sofs = [&](int i, int c) { return rows[i] * BLKD + (c % NB) * COLD + (c / NB) * NB + rowk[i]; };
    auto valid = [&](int i, int c) { return c < 2 * NB && site0 + rows[i] < kk; };
    double acc[5][2], xacc[2];
    mbar_wait(&full[slot], (it >> 1) & 1);
    // R_{n+1} = R_n (-B^-1 A) + W B^-1 + R_{n-1} (-B_prev^-1 B)
    ro_product<XN, false>(tC, M3, aoff[0], aoff[1], koff, boff, xoff, acc, xacc);
    ro_product<XN, true>(tW, M1, aoff[0], aoff[1], koff, boff, xoff, acc, xacc);
    ro_product<XN, true>(tP, M2, aoff[0], aoff[1], koff, boff, xoff, acc, xacc);
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[slot]);  // the input tiles are free: the producer refills the slot during the stores and the Gram products
    // R_{n+1} goes to its own tile (two tiles ago every warp passed the barrier below, i.e. finished the Gram products that read it)
#pragma unroll
    for (int nt = 0; nt < 5; nt++)
#pragma unroll
      for (int e = 0; e < 2; e++) {
        const int c = nt * 8 + 2 * q + e;
        if (valid(0, c)) { rnew[gofs(0, c)] = acc[nt][e]; tN[sofs(0, c)] = acc[nt][e]; }
      }
    if (XN) {
#pragma unroll
      for (int e = 0; e < 2; e++) {
        const int c = xn * 8 + 2 * q + e;
        if (valid(1, c)) { rnew[gofs(1, c)] = xacc[e]; tN[sofs(1, c)] = xacc[e]; }
      }
    }
    asm volatile("bar.sync 1, %0;" ::"r"(32 * DM_CONSUMERS) : "memory");  // the new residual tile is complete in shared memory
    const int ns = min(RO_S, kk - site0);
    if (warp + 8 < 15) rmul_gram<2>(tN, ns, warp, lane, gacc); else rmul_gram<1>(tN, ns, warp, lane, gacc);
    it++;
  }
#pragma unroll
  for (int s = 0; s < 2; s++) {  // this warp's tiles of the per-CTA partial of B^2 (matrix 0 of the slot, complex column-major)
    const int t = warp + 8 * s;
    if (t < 15) {
      const int R = (t / 5) * 8 + g;
#pragma unroll
      for (int e = 0; e < 2; e++) {
        const int C = (t % 5) * 8 + 2 * q + e;
        if (R < NB && C < 2 * NB) gpart[2 * (R + NB * (C % NB)) + (C / NB)] = gacc[s][e];
      }
    }
  }
}

// grid = (ctas, nunits).  Bmat, Bimat, Biprev, Gmat: complex column-major 18x18 per unit (stride mstride doubles; Biprev with its own
// stride, 0 = one matrix for all units: the identity of the first step).  cur = R_n, wv = W = H R_n, prev = R_{n-1} -> R_{n+1}.
// a_out (stride mstride) and a_hist (stride hstride): A_{n+1} = B^-1 G B^-1 (written by CTA 0 of each unit).  diag: scalar Lanczos
// (18 independent chains: only the real diagonal of G counts, like k_reduce_parts mode 2).
__global__ void __launch_bounds__(DM_THREADS, 1)
k_rotortho_dmma(double *prev_all, const double *cur_all, const double *w_all, const double *Bmat, const double *Bimat,
                const double *Biprev, size_t pstride, const double *Gmat, size_t mstride, double *a_out, double *a_hist, size_t hstride, int diag, int kk,
                size_t vstride, const int32_t *__restrict__ border, const int32_t *__restrict__ bcnt, int nblocks,
                double *part) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double *tiles = reinterpret_cast<double *>(smem_raw);            // [2 slots][3 tiles]
  double *tmat = tiles + 2 * 3 * RO_TILE_D;                        // [3][36x36]: T[c'][j'] = Mhat[j'][c']
  double *scratch = tmat + 3 * HBLK;                               // eight 18x18 complex matrices of the prologue
  uint64_t *full = reinterpret_cast<uint64_t *>(scratch + 8 * BLKD);
  uint64_t *empty = full + 2;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, unit = blockIdx.y;
  double *prev = prev_all + (size_t)unit * vstride;
  const double *cur = cur_all + (size_t)unit * vstride, *wv = w_all + (size_t)unit * vstride;
  if (tid == 0) {
    for (int s = 0; s < 2; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], DM_CONSUMERS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int nact = border ? bcnt[unit] : (kk + DM_S - 1) / DM_S;
  const int32_t *bo = border ? border + (size_t)unit * nblocks : nullptr;
  if (warp == DM_CONSUMERS) {
    // the tiles do not depend on the matrices: the copies of the first two start while the consumer warps form them
    if (lane == 0) {
      uint32_t it = 0;
      for (int ti = blockIdx.x; ti < 2 * nact; ti += gridDim.x) {
        const int blk = ti >> 1;
        const int site0 = (bo ? bo[blk] : blk) * DM_S + (ti & 1) * RO_S, ns = min(RO_S, kk - site0);
        if (site0 >= kk) continue;
        const int slot = it & 1;
        mbar_wait(&empty[slot], ((it >> 1) & 1) ^ 1);
        const uint32_t bytes = (uint32_t)ns * BLKD * 8;
        double *sm = tiles + (size_t)slot * 3 * RO_TILE_D;
        mbar_expect_tx(&full[slot], 3 * bytes);
        bulk_g2s(sm, cur + (size_t)site0 * BLKD, bytes, &full[slot]);
        bulk_g2s(sm + RO_TILE_D, wv + (size_t)site0 * BLKD, bytes, &full[slot]);
        bulk_g2s(sm + 2 * RO_TILE_D, prev + (size_t)site0 * BLKD, bytes, &full[slot]);
        it++;
      }
    }
    return;
  }
  // ---- the three 18x18 matrices of the step, formed by the 8 consumer warps (complex column-major in `scratch`) ----
  constexpr int NT = 32 * DM_CONSUMERS;
  auto cbar = [&]() { asm volatile("bar.sync 1, %0;" ::"r"(NT) : "memory"); };
  double *cB = scratch, *cBi = cB + BLKD, *cG = cBi + BLKD, *cT = cG + BLKD, *cA = cT + BLKD, *cM3 = cA + BLKD, *cBp = cM3 + BLKD,
         *cM2 = cBp + BLKD;
  for (int e = tid; e < BLKD; e += NT) {
    cB[e] = Bmat[(size_t)unit * mstride + e];
    cBi[e] = Bimat[(size_t)unit * mstride + e];
    cBp[e] = Biprev[(size_t)unit * pstride + e];
    double gv = Gmat[(size_t)unit * mstride + e];
    if (diag) { const int ce = e >> 1, i = ce % NB, j = ce / NB; if (i != j || (e & 1)) gv = 0.0; }
    cG[e] = gv;
  }
  cbar();
  // Z0 = s0 X0 Y0 and (n == 2) Z1 = s1 X1 Y1 in one sweep
  auto cmul18 = [&](int n, const double *X0, const double *Y0, double *Z0, double s0, const double *X1, const double *Y1, double *Z1,
                    double s1) {
    for (int e2 = tid; e2 < n * BLKC; e2 += NT) {
      const bool second = e2 >= BLKC;
      const int e = second ? e2 - BLKC : e2, r = e % NB, c = e / NB;
      const double *X = second ? X1 : X0, *Y = second ? Y1 : Y0;
      double zr = 0.0, zi = 0.0;
#pragma unroll
      for (int k = 0; k < NB; k++) {
        const double xr = X[2 * (r + NB * k)], xi = X[2 * (r + NB * k) + 1], yr = Y[2 * (k + NB * c)], yi = Y[2 * (k + NB * c) + 1];
        zr = fma(xr, yr, zr); zr = fma(-xi, yi, zr);
        zi = fma(xr, yi, zi); zi = fma(xi, yr, zi);
      }
      double *Z = second ? Z1 : Z0;
      const double sg = second ? s1 : s0;
      Z[2 * e] = sg * zr; Z[2 * e + 1] = sg * zi;
    }
    cbar();
  };
  cmul18(2, cG, cBi, cT, 1.0, cBp, cB, cM2, -1.0);            // T = G B^-1 ;  M2 = -B_prev^-1 B
  cmul18(1, cBi, cT, cA, 1.0, nullptr, nullptr, nullptr, 0.0);   // A = B^-1 G B^-1
  cmul18(1, cBi, cA, cM3, -1.0, nullptr, nullptr, nullptr, 0.0); // M3 = -B^-1 A
  if (blockIdx.x == 0)
    for (int e = tid; e < BLKD; e += NT) {
      double v = cA[e];
      if (diag) { const int ce = e >> 1, i = ce % NB, j = ce / NB; if (i != j || (e & 1)) v = 0.0; }
      a_out[(size_t)unit * mstride + e] = v;
      if (a_hist) a_hist[(size_t)unit * hstride + e] = v;
    }
  for (int e = tid; e < 3 * HBLK; e += NT) {
    const int w = e / HBLK, r = e % HBLK, c = r / COLD, j = r % COLD;  // T[c'][j']
    const double *m = w == 0 ? cBi : w == 1 ? cM2 : cM3;
    const double re = m[2 * ((j % NB) + NB * (c % NB))], im = m[2 * ((j % NB) + NB * (c % NB)) + 1];
    tmat[e] = (j < NB) == (c < NB) ? re : (j < NB ? im : -im);
  }
  cbar();
  double *gpart = part + ((size_t)unit * gridDim.x + blockIdx.x) * (2 * BLKD);
  if (warp < 5) rotortho_consumer<1>(prev, kk, tiles, tmat, scratch, full, empty, warp, lane, bo, nact, gpart);
  else rotortho_consumer<0>(prev, kk, tiles, tmat, scratch, full, empty, warp, lane, bo, nact, gpart);
}

// ---- host side ------------------------------------------------------------------------------------------------
static int dmma_configure() {
#define DM_ATTR(E, A) \
  if (cudaFuncSetAttribute(k_apply_dmma<E, A, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, ApGeom<8>::kSmem) != cudaSuccess) return -3; \
  if (cudaFuncSetAttribute(k_apply_dmma<E, A, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, ApGeom<4>::kSmem) != cudaSuccess) return -3; \
  if (cudaFuncSetAttribute(k_apply_dmma_sd<E, A>, cudaFuncAttributeMaxDynamicSharedMemorySize, SdGeom<false>::kSmem) != cudaSuccess) return -3; \
  if (cudaFuncSetAttribute(k_apply_dmma_sd8<E, A>, cudaFuncAttributeMaxDynamicSharedMemorySize, SdGeom<false>::kSmem) != cudaSuccess) return -3;
#define DM_ATTR_GRAM(E, A) \
  if (cudaFuncSetAttribute(k_apply_dmma<E, A, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, ApGeom<4>::kSmemGram) != cudaSuccess) return -3; \
  if (cudaFuncSetAttribute(k_apply_dmma_sd<E, A>, cudaFuncAttributeMaxDynamicSharedMemorySize, SdGeom<true>::kSmem) != cudaSuccess) return -3;
  DM_ATTR(EPI_STORE, false)
  DM_ATTR(EPI_STORE, true)
  DM_ATTR(EPI_HAM, false)
  DM_ATTR(EPI_HAM, true)
  DM_ATTR(EPI_CHEB_NOGRAM, false)
  DM_ATTR(EPI_CHEB_NOGRAM, true)
  DM_ATTR(EPI_HOP, false)
  DM_ATTR(EPI_HOP, true)
  DM_ATTR_GRAM(EPI_CHEB, false)
  DM_ATTR_GRAM(EPI_CHEB, true)
  DM_ATTR_GRAM(EPI_HOP_GRAM, false)
  DM_ATTR_GRAM(EPI_HOP_GRAM, true)
  if (cudaFuncSetAttribute(k_apply_dmma_sd8<EPI_HOP_GRAM, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SdGeom<true>::kSmem) != cudaSuccess) return -3;
  if (cudaFuncSetAttribute(k_apply_dmma_sd8<EPI_HOP_GRAM, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SdGeom<true>::kSmem) != cudaSuccess) return -3;
#undef DM_ATTR
#undef DM_ATTR_GRAM
  if (cudaFuncSetAttribute(k_gram_dmma, cudaFuncAttributeMaxDynamicSharedMemorySize, GR_SMEM_BYTES) != cudaSuccess) return -3;
  if (cudaFuncSetAttribute(k_rmul_dmma<RM_ORTHO>, cudaFuncAttributeMaxDynamicSharedMemorySize, RM_SMEM_BYTES) != cudaSuccess) return -3;
  if (cudaFuncSetAttribute(k_rmul_dmma<RM_ROTATE>, cudaFuncAttributeMaxDynamicSharedMemorySize, RM_SMEM_BYTES) != cudaSuccess) return -3;
  if (cudaFuncSetAttribute(k_rotortho_dmma, cudaFuncAttributeMaxDynamicSharedMemorySize, RO_SMEM_BYTES) != cudaSuccess) return -3;
  return 0;
}

static void dmma_free_tiles(DmmaTiles &t) {
  if (t.d_sites) cudaFree(t.d_sites);
  if (t.d_cls) cudaFree(t.d_cls);
  if (t.d_nbr) cudaFree(t.d_nbr);
  t = DmmaTiles();
}

// Tiles = up to 8 sites of one Hamiltonian class, ordered by their first site so that consecutive CTAs touch
// neighbouring parts of the vector (L2 reuse of the gathered blocks).  nbr: [ng][kk], cls: [kk].
// pos (optional, 3 x kk = lattice%cr): tiles are then ordered along a Morton (Z-order) curve through their first site,
// so that the 148 persistent CTAs, which walk the list side by side, gather from a compact region of the vector: at
// 10^6 sites in storage order the three z-planes a tile row touches (3 x 104 MB) exceed the 126 MB L2 and every psi
// block was fetched from DRAM 2.6 times.
static uint32_t morton_spread10(uint32_t v) {  // 10 bits -> every third bit
  v &= 0x3ffu;
  v = (v | (v << 16)) & 0x030000ffu;
  v = (v | (v << 8)) & 0x0300f00fu;
  v = (v | (v << 4)) & 0x030c30c3u;
  v = (v | (v << 2)) & 0x09249249u;
  return v;
}
static int dmma_apply_geom();
static int dmma_build_tiles(DmmaTiles &t, const std::vector<int32_t> &nbr, const std::vector<int32_t> &cls, int kk, int ng,
                            int ncls, int ncls_type, const double *pos = nullptr) {
  dmma_free_tiles(t);
  std::vector<std::vector<int32_t>> by_cls(ncls);
  for (int i = 0; i < kk; i++) by_cls[cls[i]].push_back(i);
  // Type-indexed classes (c < ncls_type): 8 sites per tile, one H block per stage.  Site-indexed classes (hall: one class per
  // site): a class has a single site, so a tile takes TWO of them, one per half (the S = 4 geometry loads the H block of a
  // pass by the class of its half tile); with the S = 8 geometry a tile has one class and such a site has a tile of its own.
  struct T { int32_t s[DM_S]; int32_t c[2]; };
  std::vector<T> tiles;
  const bool halves = dmma_apply_geom() == 4;
  T open_local; bool have_open = false;
  for (int c = 0; c < ncls; c++) {
    const bool local = c >= ncls_type;
    if (local && halves) {
      for (size_t o = 0; o < by_cls[c].size(); o++) {   // one site per local class (more only if the caller reuses a class)
        if (!have_open) {
          for (int k = 0; k < DM_S; k++) open_local.s[k] = kk;
          open_local.s[0] = by_cls[c][o]; open_local.c[0] = open_local.c[1] = c;
          have_open = true;
        } else {
          open_local.s[DM_S / 2] = by_cls[c][o]; open_local.c[1] = c;
          tiles.push_back(open_local);
          have_open = false;
        }
      }
      continue;
    }
    for (size_t o = 0; o < by_cls[c].size(); o += DM_S) {
      T x;
      x.c[0] = x.c[1] = c;
      for (int k = 0; k < DM_S; k++) x.s[k] = (o + k < by_cls[c].size()) ? by_cls[c][o + k] : kk;
      tiles.push_back(x);
    }
  }
  if (have_open) tiles.push_back(open_local);
  if (pos) {
    double lo[3], hi[3];
    for (int l = 0; l < 3; l++) lo[l] = hi[l] = pos[l];
    for (int i = 1; i < kk; i++)
      for (int l = 0; l < 3; l++) { lo[l] = std::min(lo[l], pos[l + 3 * (size_t)i]); hi[l] = std::max(hi[l], pos[l + 3 * (size_t)i]); }
    const double ext = std::max(std::max(hi[0] - lo[0], hi[1] - lo[1]), std::max(hi[2] - lo[2], 1e-300));
    std::vector<uint32_t> key(tiles.size());
    std::vector<size_t> perm(tiles.size());
    for (size_t i = 0; i < tiles.size(); i++) {
      uint32_t q[3];
      for (int l = 0; l < 3; l++) q[l] = (uint32_t)std::min(1023.0, std::floor((pos[l + 3 * (size_t)tiles[i].s[0]] - lo[l]) / ext * 1024.0));
      key[i] = morton_spread10(q[0]) | (morton_spread10(q[1]) << 1) | (morton_spread10(q[2]) << 2);
      perm[i] = i;
    }
    std::sort(perm.begin(), perm.end(), [&](size_t a, size_t b) { return key[a] != key[b] ? key[a] < key[b] : tiles[a].s[0] < tiles[b].s[0]; });
    std::vector<T> sorted(tiles.size());
    for (size_t i = 0; i < tiles.size(); i++) sorted[i] = tiles[perm[i]];
    tiles.swap(sorted);
  } else {
    std::sort(tiles.begin(), tiles.end(), [](const T &a, const T &b) { return a.s[0] < b.s[0]; });
  }
  const int nt = (int)tiles.size();
  std::vector<int32_t> hs((size_t)nt * DM_S), hc((size_t)2 * nt), hn((size_t)nt * ng * DM_S);
  for (int i = 0; i < nt; i++) {
    hc[2 * i] = tiles[i].c[0]; hc[2 * i + 1] = tiles[i].c[1];
    for (int k = 0; k < DM_S; k++) {
      const int s = tiles[i].s[k];
      hs[(size_t)i * DM_S + k] = s;
      for (int m = 0; m < ng; m++) hn[((size_t)i * ng + m) * DM_S + k] = (s < kk) ? nbr[(size_t)m * kk + s] : kk;
    }
  }
  if (cudaMalloc(&t.d_sites, hs.size() * 4) != cudaSuccess || cudaMalloc(&t.d_cls, hc.size() * 4) != cudaSuccess ||
      cudaMalloc(&t.d_nbr, hn.size() * 4) != cudaSuccess)
    return -4;
  cudaMemcpy(t.d_sites, hs.data(), hs.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(t.d_cls, hc.data(), hc.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(t.d_nbr, hn.data(), hn.size() * 4, cudaMemcpyHostToDevice);
  t.ntiles = nt; t.ng = ng; t.kk = kk;
  t.h_sites = hs;
  return 0;
}

// The epilogues that read `in` need its self blocks in the last pipeline stage.
static bool dmma_supported(const ApplyParams &p) {
  // the epilogues with fused Gram products (S = 4 geometry only) need the partial buffer; EPI_HOP needs out2 (hpsi)
  if ((p.epi == EPI_CHEB || p.epi == EPI_HOP_GRAM) && !p.part) return false;
  if (p.epi == EPI_HOP && !p.out2) return false;
  int nst = 0;
  for (int t = 0; t < p.ngterms; t++) nst += p.ngather - p.g[t].first_slot;
  if (p.Hx) nst++;
  if (nst < 1 || nst > DM_MAXST) return false;
  // the epilogues that read `in` (and the fused Gram products) need its self blocks in the last pipeline stage
  if (p.epi == EPI_HAM || p.epi == EPI_CHEB_NOGRAM || p.epi == EPI_CHEB || p.epi == EPI_HOP_GRAM) {
    if (p.Hx) return p.srcx == p.in;
    return p.ngterms == 1 && p.g[0].first_slot == 0 && p.g[0].src == p.in;
  }
  return true;
}
static int dmma_apply_geom() {
  static const int geom = getenv("RSREC_APPLY_S") ? atoi(getenv("RSREC_APPLY_S")) : DM_APPLY_S;
  return geom;
}
// CTAs of an SpMV launch = per-unit partial-sum slots the fused Gram variants write
static int dmma_apply_grid(const DmmaTiles &t, int sms, int nunits) {
  const long long passes = (long long)t.ntiles * std::max(1, nunits) * (dmma_apply_geom() == 4 ? 2 : 1);
  return (int)std::max<long long>(1, std::min<long long>(passes, dmma_apply_geom() == 4 ? 2 * sms : sms));
}
static int dmma_grid(const DmmaTiles &t, int sms) { return std::max(1, std::min(t.ntiles, sms)); }
static int dmma_gram_ctas(int kk, int sms) { return std::max(1, std::min(sms, (kk + GR_WARPS - 1) / GR_WARPS)); }

// sdflags(H set, slot) -> true when every block of that slot is spin-diagonal (nullptr: never)
// (H set, slot) -> true when every block of that slot is spin-diagonal; *hs18 = the HS18 twin of the set (null: none)
typedef bool (*SdLookup)(const void *ctx, const double *Hset, int slot, const double **hs18);
static int dmma_launch_apply(DmmaTiles &t, ApplyParams &p, int nunits, int sms, cudaStream_t st, long long *launches,
                             const int32_t *order = nullptr, const int32_t *cnt = nullptr, SdLookup sdl = nullptr,
                             const void *sdctx = nullptr, long long *sd_launches = nullptr, int *nparts_out = nullptr) {
  DmmaStages sg, sg2;
  memset(&sg, 0, sizeof(sg));
  memset(&sg2, 0, sizeof(sg2));
  sg.n = 0;
  bool have_hs18 = sdl != nullptr;
  // the spin-resolved kernel's list: one HS18 half stage per slot (kind 0), a second one (kind 1) where the slot couples the spins
  auto add2 = [&](const double *hs18, size_t half_index, int hstride, const double *src, int slot, bool diag) {
    for (int kind = 0; kind < (diag ? 1 : 2); kind++) {
      if (!hs18 || sg2.n >= DM_MAXST) { have_hs18 = false; return; }
      sg2.H[sg2.n] = hs18 + (half_index * 2 + kind) * SDH;
      sg2.src[sg2.n] = src; sg2.hstride[sg2.n] = hstride; sg2.slot[sg2.n] = slot; sg2.sd[sg2.n] = (unsigned char)kind;
      sg2.n++;
    }
  };
  // neighbour slots of every term first, then the on-site slots, then the on-site extra term (self stage last)
  for (int pass = 0; pass < 2; pass++)
    for (int tm = 0; tm < p.ngterms; tm++)
      for (int m = (pass == 0 ? std::max(1, p.g[tm].first_slot) : 0); m < (pass == 0 ? p.ngather : 1); m++) {
        if (pass == 1 && p.g[tm].first_slot > 0) continue;
        sg.H[sg.n] = p.g[tm].H + (size_t)m * HBLK;
        sg.src[sg.n] = p.g[tm].src;
        sg.hstride[sg.n] = p.nslot_h * HBLK;
        sg.slot[sg.n] = m;
        const double *hs = nullptr;
        sg.sd[sg.n] = sdl && sdl(sdctx, p.g[tm].H, m, &hs);
        if (sdl) add2(hs, (size_t)m, p.nslot_h * 2 * SDH, p.g[tm].src, m, sg.sd[sg.n] != 0);
        sg.n++;
      }
  if (p.Hx) {
    sg.H[sg.n] = p.Hx; sg.src[sg.n] = p.srcx; sg.hstride[sg.n] = HBLK; sg.slot[sg.n] = 0;
    const double *hs = nullptr;
    sg.sd[sg.n] = sdl && sdl(sdctx, p.Hx, 0, &hs);
    if (sdl) add2(hs, 0, 2 * SDH, p.srcx, 0, sg.sd[sg.n] != 0);
    sg.n++;
  }
  int nsd = 0;
  for (int j = 0; j < sg.n; j++) nsd += sg.sd[j];
  const int geom = dmma_apply_geom();
  const int grid = dmma_apply_grid(t, sms, nunits);
  // half stages pay off when most slots are spin-diagonal (a coupling slot costs two of them: 60 instead of 45 DMMAs per m-tile)
  const bool use_sd = geom == 4 && have_hs18 && 2 * nsd > sg.n;
  if (use_sd && sd_launches) (*sd_launches)++;
  const char *sdw = getenv("RSREC_SD_WARPS");      // A/B switch, read per launch: 8 (default) or 4 consumer warps
  const bool sd8 = !(sdw && atoi(sdw) == 4);
#define DM_LAUNCH(E, A)                                                                                                   \
  do {                                                                                                                    \
    if (use_sd && sd8)                                                                                                    \
      k_apply_dmma_sd8<E, A><<<grid, SD8_THREADS, SdGeom<false>::kSmem, st>>>(p, sg2, t.d_sites, t.d_cls, t.d_nbr,         \
                                                                           t.ntiles, nunits, order, cnt);                \
    else if (use_sd)                                                                                                      \
      k_apply_dmma_sd<E, A><<<grid, ApGeom<4>::kThreads, SdGeom<false>::kSmem, st>>>(p, sg2, t.d_sites, t.d_cls, t.d_nbr,  \
                                                                                  t.ntiles, nunits, order, cnt);         \
    else if (geom == 4)                                                                                                   \
      k_apply_dmma<E, A, 4><<<grid, ApGeom<4>::kThreads, ApGeom<4>::kSmem, st>>>(p, sg, t.d_sites, t.d_cls, t.d_nbr,       \
                                                                              t.ntiles, nunits, order, cnt);             \
    else                                                                                                                  \
      k_apply_dmma<E, A, 8><<<grid, ApGeom<8>::kThreads, ApGeom<8>::kSmem, st>>>(p, sg, t.d_sites, t.d_cls, t.d_nbr,       \
                                                                              t.ntiles, nunits, order, cnt);             \
  } while (0)
#define DM_LAUNCH_GRAM(E, A)                                                                                              \
  do {                                                                                                                    \
    if (use_sd && sd8 && E == EPI_HOP_GRAM)                                                                               \
      k_apply_dmma_sd8<EPI_HOP_GRAM, A><<<grid, SD8_THREADS, SdGeom<true>::kSmem, st>>>(p, sg2, t.d_sites, t.d_cls,        \
                                                                                     t.d_nbr, t.ntiles, nunits, order, cnt); \
    else if (use_sd)                                                                                                      \
      k_apply_dmma_sd<E, A><<<grid, ApGeom<4>::kThreads, SdGeom<true>::kSmem, st>>>(p, sg2, t.d_sites, t.d_cls, t.d_nbr,   \
                                                                                 t.ntiles, nunits, order, cnt);          \
    else                                                                                                                  \
      k_apply_dmma<E, A, 4><<<grid, ApGeom<4>::kThreads, ApGeom<4>::kSmemGram, st>>>(p, sg, t.d_sites, t.d_cls, t.d_nbr,   \
                                                                                  t.ntiles, nunits, order, cnt);         \
  } while (0)
  const bool ad = p.addend != nullptr;
  if (nparts_out) *nparts_out = grid;
  switch (p.epi) {
    case EPI_STORE: if (ad) DM_LAUNCH(EPI_STORE, true); else DM_LAUNCH(EPI_STORE, false); break;
    case EPI_HAM: if (ad) DM_LAUNCH(EPI_HAM, true); else DM_LAUNCH(EPI_HAM, false); break;
    case EPI_CHEB_NOGRAM: if (ad) DM_LAUNCH(EPI_CHEB_NOGRAM, true); else DM_LAUNCH(EPI_CHEB_NOGRAM, false); break;
    case EPI_HOP: if (ad) DM_LAUNCH(EPI_HOP, true); else DM_LAUNCH(EPI_HOP, false); break;
    case EPI_CHEB: if (geom != 4) return -1; if (ad) DM_LAUNCH_GRAM(EPI_CHEB, true); else DM_LAUNCH_GRAM(EPI_CHEB, false); break;
    case EPI_HOP_GRAM: if (geom != 4) return -1; if (ad) DM_LAUNCH_GRAM(EPI_HOP_GRAM, true); else DM_LAUNCH_GRAM(EPI_HOP_GRAM, false); break;
    default: return -1;
  }
#undef DM_LAUNCH
#undef DM_LAUNCH_GRAM
  (*launches)++;
  return cudaGetLastError() == cudaSuccess ? 0 : -3;
}

static int dmma_launch_rotortho(double *prev, const double *cur, const double *w, const double *B, const double *Bi,
                                const double *Biprev, size_t pstride, const double *G, size_t mstride, double *a_out, double *a_hist, size_t hstride, int diag, int kk, size_t vstride,
                                int nunits, int ctas, cudaStream_t st, long long *launches, const int32_t *border,
                                const int32_t *bcnt, int nblocks, double *part) {
  dim3 grid(ctas, nunits);
  k_rotortho_dmma<<<grid, DM_THREADS, RO_SMEM_BYTES, st>>>(prev, cur, w, B, Bi, Biprev, pstride, G, mstride, a_out, a_hist, hstride, diag, kk,
                                                           vstride, border, bcnt, nblocks, part);
  (*launches)++;
  return cudaGetLastError() == cudaSuccess ? 0 : -3;
}

static int dmma_launch_gram(const double *X, size_t xstride, const double *Y, size_t ystride, int two, int kk,
                            int nunits, int ctas, double *part, cudaStream_t st, long long *launches,
                            const int32_t *border = nullptr, const int32_t *bcnt = nullptr, int nblocks = 0) {
  dim3 grid(ctas, nunits);
  k_gram_dmma<<<grid, GR_THREADS, GR_SMEM_BYTES, st>>>(X, Y, two, kk, xstride, ystride, part, border, bcnt, nblocks);
  (*launches)++;
  return cudaGetLastError() == cudaSuccess ? 0 : -3;
}

static int dmma_rmul_ctas(int kk, int sms, int nunits) { return std::max(1, std::min((kk + DM_S - 1) / DM_S, sms / std::max(1, nunits))); }
static int dmma_launch_rmul(int mode, double *psi, double *pmn, const double *hpsi, const double *m0, const double *m1,
                            size_t mstride, int kk, size_t vstride, int nunits, int sms, cudaStream_t st,
                            long long *launches, const int32_t *border = nullptr, const int32_t *bcnt = nullptr,
                            int nblocks = 0, double *part = nullptr, int *nparts_out = nullptr) {
  dim3 grid(dmma_rmul_ctas(kk, sms, nunits), nunits);
  if (nparts_out) *nparts_out = (int)grid.x;
  if (mode == RM_ORTHO)
    k_rmul_dmma<RM_ORTHO><<<grid, DM_THREADS, RM_SMEM_BYTES, st>>>(psi, pmn, hpsi, m0, m1, mstride, kk, vstride, border, bcnt, nblocks, part);
  else
    k_rmul_dmma<RM_ROTATE><<<grid, DM_THREADS, RM_SMEM_BYTES, st>>>(psi, pmn, hpsi, m0, m1, mstride, kk, vstride, border, bcnt, nblocks, nullptr);
  (*launches)++;
  return cudaGetLastError() == cudaSuccess ? 0 : -3;
}

// rsrec.cu -- C ABI (include/rsrec.h) and host orchestration of the B200-native recursion engine.
//
// Host-side structure mirrors the reference drivers: recur_b / recur_b_ij / crecal_b (recursion.f90:1655-1973),
// recur / crecal (3423-3532), chebyshev_recur(_ij) (2376-2487, 3057-3130), compute_moments_stochastic (979-1234).
// All arithmetic runs in CUDA kernels; the host only sequences launches.  No CPU fallback exists.
#include "../../include/rsrec.h"
#include "common.cuh"
#include "kernels_simt.cuh"
#include "eig18.cuh"
#include "kernels_dmma.cuh"
#include "kernels_kubo.cuh"
#include "kernels_post.cuh"
#include "kernels_lattice.cuh"
#include "kernels_ham.cuh"
#include "kernels_bands.cuh"
#include "comm_nccl.cuh"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

static thread_local std::string g_err;
static int fail(int code, const std::string &msg) { g_err = msg; return code; }
#define CUDA_TRY(x)                                                                                         \
  do {                                                                                                      \
    cudaError_t e_ = (x);                                                                                   \
    if (e_ != cudaSuccess) {                                                                                \
      cudaGetLastError();                                                                                   \
      return fail(e_ == cudaErrorMemoryAllocation ? RSREC_ENOMEM : RSREC_ECUDA,                             \
                  std::string(#x) + ": " + cudaGetErrorString(e_) + " (" + __FILE__ + ":" + std::to_string(__LINE__) + ")"); \
    }                                                                                                       \
  } while (0)
#define TRY(x) do { int rc_ = (x); if (rc_ != RSREC_OK) return rc_; } while (0)

typedef rsrec_cplx cplx;

struct DevBuf {
  double *p = nullptr;
  size_t n = 0;  // doubles
};

struct rsrec_handle_s {
  int dev = 0, kk = 0, ncols = 0, nslot = 0, ntype = 0, nmax = 0, hoh = 0, ncls = 0;
  int family = 1;
  bool fuse_lanczos = true, fuse_cheb = false;  // Gram products inside the SpMV kernel (rsrec_set_fusion)
  bool fam_ok = true;  // every operator application of this handle fits the tensor pipeline's stage list (set by ensure_ready)
  int sms = 148;
  cudaStream_t st = nullptr;
  cudaStream_t st2 = nullptr;               // side stream of the pipelined Lanczos step (k_lz_eig beside the SpMV), created on first use
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  int sm_reserve = 0;                       // SMs the next gather-SpMV launch leaves one CTA slot free on (for the side stream's kernel)
  long long launches = 0;
  long long sd_launches = 0;  // SpMV launches that took the spin-diagonal two-block kernel
  int last_parts = 0;  // partial-sum slots per unit written by the last fused apply
  double *out2 = nullptr;  // second output of the next EPI_HOP application (hpsi)
  int sqrt_method = 1;     // B = (B^2)^1/2: 1 = Newton-Schulz with Jacobi fallback, 0 = Jacobi eigen-decomposition
  long long h2d_bytes = 0, d2h_bytes = 0;  // bytes moved over PCIe/C2C by this handle (bench.py's e2e accounting)
  bool profile = false;                    // record CUDA events around every gather-SpMV launch
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_events;
  size_t prof_used = 0;
  // per-phase device timing under the reference's g_timer labels (recursion.f90:1902-1970, 3104-3127)
  bool phase_on = false;
  struct PhaseEv { int phase; cudaEvent_t a, b; };
  std::vector<PhaseEv> phase_events;
  size_t phase_used = 0;
  // host copies of the reference arrays (small) so that device sets can be (re)built in any call order
  std::vector<int32_t> nn, iz;
  std::vector<double> pos;  // optional lattice%cr (3,kk): only used to order the work (L2 locality), never in arithmetic
  std::vector<cplx> ee, eeo, hall, hallo, lsham, enim, v_a, v_b, vo_a, vo_b;
  bool have_lat = false, have_ham = false, have_op[2] = {false, false}, dirty = true /* lattice tables */, dirty_ham = true /* block sets */;
  // device operator data
  int32_t *d_nbr = nullptr, *d_cls = nullptr;
  DevBuf Hmain, Hh, Hho_neg, Hx, Hscalar, Hva, Hvb, Hvoa_neg, Hvob_neg;
  // complex block sets in the reference's layout, classes = types then local sites: ee|hall, eeo|hallo, lsham, enim,
  // obarm.  Filled from the host arrays (rsrec_set_hamiltonian) or assembled on the device (rsrec_build_hamiltonian).
  DevBuf cBLK, cBLKO, cLS, cENIM, cOBARM, cV[2], cVO[2];
  DevBuf gBLK, gBLKO, gENIM;  // the reference's *_glob copies (hamiltonian.f90:70-80) while rotated to a local axis
  bool have_glob = false;
  int32_t *d_cls_type = nullptr;
  bool ham_on_device = false;  // cBLK.. are current (device-side assembly): do not re-stage from the host copies
  DmmaTiles tiles;
  // work vectors and small matrices
  std::vector<DevBuf> vecs;
  DevBuf part, A, B, Bi, B2, mu, ahist, b2hist, bhist /* B = (B^2)^1/2 per level, what zsqr would return */, scratch;
  // per H set (key = device pointer of the packed set): slot -> 1 when every block of the slot is spin-diagonal
  std::map<const double *, std::vector<unsigned char>> sdmap;
  // HS18 twins of the packed sets (half blocks per spin for the spin-resolved SpMV kernel), keyed like sdmap
  std::map<const double *, DevBuf> hs18;
  DevBuf post[12];  // work arrays of the post-recursion consumers (terminator, Green functions, Kubo back end)
  // on-site Green function of the last Green-function call, kept for the `bands` consumers (bands.f90)
  DevBuf g0all, bands_y, bands_out;
  int g0_units = 0, g0_nv = 0;
  int32_t *d_si = nullptr, *d_sj = nullptr;
  double *d_as = nullptr, *d_bs = nullptr;
  int units_cap = 0;
  // active-region plan (the reference's izero/irlist): tiles reachable per step from the units' start sites
  std::vector<int32_t> radj_off, radj;  // reverse adjacency of the neighbour table (who gathers from site a)
  std::vector<int32_t> nbr_host;        // [ncols][kk] 0-based neighbour table as uploaded
  struct {
    bool on = false;
    int level = 0, maxlevel = 0, nunits = 0;
    std::vector<int32_t> key;  // start sites (i, j per unit) the plan was built for
    int32_t *d_order = nullptr, *d_counts = nullptr;
    size_t order_cap = 0, counts_cap = 0;
    // the same levels for contiguous 8-site blocks (what the site-ordered Gram / right-multiplication kernels walk)
    int32_t *d_border = nullptr, *d_bcounts = nullptr;
    size_t border_cap = 0, bcounts_cap = 0;
    int nblocks = 0;
  } plan;
  // host wall-clock seconds per stage of a sharded call (rsrec_host_phase_read): tables, plan, recursion, exchange, download
  double hphase[5] = {0, 0, 0, 0, 0};
  // NCCL communicator of the unit-sharded job (rsrec_comm_init); null = single rank
  rs_ncclComm_t comm = nullptr;
  int comm_rank = 0, comm_size = 1;
  DevBuf comm_buf;  // staging for collectives on host arrays
  DevBuf comm_res;  // device-resident results that are exchanged (gathered histories, summed moments)
  // Chebyshev stepping session
  struct {
    bool active = false;
    int nunits = 0, lld = 0, done = 0, nctas = 0;
    double a = 1, b = 0;
    int i0 = 0, i1 = 1;  // indices into vecs of psi0, psi1
  } cheb;
};
typedef rsrec_handle_s H;
// Effective kernel family of a call.  The tensor pipeline lists one stage per gathered slot (DM_MAXST of them); when an
// operator of this handle needs more (very long neighbour lists), EVERY kernel of the call runs in family 0 -- mixing the
// families inside one recursion is not supported (the SIMT SpMV has no hpsi output and does not advance the plan).
static bool fam1(const H *h) { return h->family == 1 && h->fam_ok; }
// The phases the reference times with g_timer inside crecal_b and chebyshev_recur, in its own words.
enum Phase { PH_HPSI = 0, PH_ORTHO, PH_BNEXT, PH_ROTATE, PH_MOM0, PH_MOM1, PH_MOMN, PH_COUNT };
static const char *const kPhaseLabel[PH_COUNT] = {"H|PSI_n>", "H|Psi_n-A_n|Psi_n-B_n|Psi_n-1", "B_n+1", "<PSI|B_n+1|PSI>",
                                                   "<PSI_0|PSI_0>", "<PSI_0|PSI_1>", "<PSI_0|PSI_n>"};
// Host wall-clock per stage of a call, for the strong-scaling breakdown: 0 tables (ensure_ready), 1 plan (active-region
// BFS + unit upload), 2 recursion (enqueue + device time), 3 exchange (NCCL), 4 download.  The stage boundaries are stream
// synchronisations only while phase timing is on (otherwise a stage's time is its enqueue time and the device time lands in
// the first stage that synchronises).
enum HostPhase { HP_TABLES = 0, HP_PLAN, HP_RECUR, HP_EXCHANGE, HP_DOWNLOAD, HP_COUNT };
struct HostScope {
  H *h; int k; std::chrono::steady_clock::time_point t0;
  HostScope(H *h_, int k_) : h(h_), k(k_), t0(std::chrono::steady_clock::now()) {}
  ~HostScope() {
    if (h->phase_on) cudaStreamSynchronize(h->st);
    h->hphase[k] += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  }
};
// records a CUDA event pair around a phase on the handle's stream while phase timing is enabled (two event records per
// phase; nothing otherwise)
struct PhaseScope {
  H *h; size_t idx = (size_t)-1; cudaStream_t s;
  PhaseScope(H *h_, int phase, cudaStream_t stream = nullptr) : h(h_), s(stream ? stream : h_->st) {
    if (!h->phase_on) return;
    if (h->phase_used == h->phase_events.size()) {
      rsrec_handle_s::PhaseEv e; e.phase = phase;
      if (cudaEventCreate(&e.a) != cudaSuccess || cudaEventCreate(&e.b) != cudaSuccess) return;
      h->phase_events.push_back(e);
    }
    idx = h->phase_used++;
    h->phase_events[idx].phase = phase;
    cudaEventRecord(h->phase_events[idx].a, s);
  }
  ~PhaseScope() { if (idx != (size_t)-1) cudaEventRecord(h->phase_events[idx].b, s); }
};
static int post_configure();
static int comm_allreduce_dev(rsrec_handle_s *h, double *d, size_t n);
static int grid_for(size_t n, int threads, int cap);

// ------------------------------------------------------------------------------------------------------------
static int dev_alloc(DevBuf &b, size_t n, bool zero) {
  if (b.n < n) {
    if (b.p) cudaFree(b.p);
    b.p = nullptr; b.n = 0;
    CUDA_TRY(cudaMalloc(&b.p, n * sizeof(double)));
    b.n = n;
    if (zero) CUDA_TRY(cudaMemset(b.p, 0, n * sizeof(double)));
  }
  return RSREC_OK;
}
static void dev_free(DevBuf &b) { if (b.p) cudaFree(b.p); b.p = nullptr; b.n = 0; }

static size_t vstride(const H *h) { return (size_t)(h->kk + 1) * BLKD; }

static int get_vec(H *h, int idx, int nunits, double **out) {
  if ((int)h->vecs.size() <= idx) h->vecs.resize(idx + 1);
  size_t need = vstride(h) * nunits;
  if (h->vecs[idx].n < need) {
    dev_free(h->vecs[idx]);
    TRY(dev_alloc(h->vecs[idx], need, true));
  }
  *out = h->vecs[idx].p;
  return RSREC_OK;
}
static int zero_vec(H *h, double *v, int nunits) {
  CUDA_TRY(cudaMemsetAsync(v, 0, vstride(h) * nunits * sizeof(double), h->st));
  return RSREC_OK;
}

// Stream-ordered on the handle's stream: a synchronous cudaMemcpy from pageable memory returns once the data is in the
// driver's staging buffer, NOT when it has reached the device, and the handle's stream is non-blocking -- a kernel
// launched on it right afterwards could read the destination before the DMA lands.
static int upload(rsrec_handle_s *h, DevBuf &b, const std::vector<double> &host);

// Build the device operator sets from the host copies (the "device-resident CSR-of-blocks" export of the
// hamiltonian builder and the device index array of the lattice): classes 0..ntype-1 = bulk types (ee),
// ntype..ntype+nmax-1 = site-indexed local region (hall) -- hamiltonian.f90:1553-1667, lattice.f90:1856-1860.
static int upload(H *h, DevBuf &b, const std::vector<double> &host) {
  TRY(dev_alloc(b, host.size(), false));
  CUDA_TRY(cudaMemcpyAsync(b.p, host.data(), host.size() * sizeof(double), cudaMemcpyHostToDevice, h->st));
  h->h2d_bytes += (long long)(host.size() * sizeof(double));
  return RSREC_OK;
}

static int ensure_ready(H *h) {
  if (!h->have_lat) return fail(RSREC_EINVAL, "rsrec_set_lattice has not been called");
  if (!h->have_ham) return fail(RSREC_EINVAL, "rsrec_set_hamiltonian has not been called");
  {  // longest stage list any operator of this handle builds: H (+ the on-site e_nu + l.s stage with hoh), v - vo.(ee.)
    int maxst = h->ncols + (h->hoh ? 1 : 0);
    if (h->hoh && (h->have_op[0] || h->have_op[1])) maxst = std::max(maxst, 2 * h->ncols - 1);
    h->fam_ok = maxst <= DM_MAXST;
  }
  if (!h->dirty && !h->dirty_ham) return RSREC_OK;
  const int kk = h->kk, nslot = h->nslot, ncls = h->ncls, ng = h->ncols;
  if (h->dirty) {
  // neighbour table [slot][site], slot 0 = self, missing -> kk
  std::vector<int32_t> nbr((size_t)ng * kk), cls(kk);
  for (int i = 0; i < kk; i++) {
    const int nr = h->nn[i];
    if (nr > ng) return fail(RSREC_EINVAL, "nn(i,1) exceeds the second dimension of nn");
    nbr[i] = i;
    for (int j = 1; j < ng; j++) {
      int nb = (j < nr) ? h->nn[(size_t)i + (size_t)kk * j] : 0;
      if (nb < 0 || nb > kk) return fail(RSREC_EINVAL, "nn entry out of range");
      nbr[(size_t)j * kk + i] = nb == 0 ? kk : nb - 1;
    }
    const int t = h->iz[i];
    if (t < 1 || t > h->ntype) return fail(RSREC_EINVAL, "iz entry out of range");
    cls[i] = (i < h->nmax) ? h->ntype + i : t - 1;
  }
  if (!h->d_nbr) CUDA_TRY(cudaMalloc(&h->d_nbr, nbr.size() * sizeof(int32_t)));
  if (!h->d_cls) CUDA_TRY(cudaMalloc(&h->d_cls, cls.size() * sizeof(int32_t)));
  CUDA_TRY(cudaMemcpyAsync(h->d_nbr, nbr.data(), nbr.size() * sizeof(int32_t), cudaMemcpyHostToDevice, h->st));
  CUDA_TRY(cudaMemcpyAsync(h->d_cls, cls.data(), cls.size() * sizeof(int32_t), cudaMemcpyHostToDevice, h->st));
  h->h2d_bytes += (long long)((nbr.size() + cls.size()) * sizeof(int32_t));
  h->nbr_host.swap(nbr);  // kept for the active-region planner (reverse adjacency is built on first use)
  h->radj_off.clear(); h->radj.clear();
  h->plan.on = false; h->plan.key.clear();
  const std::vector<int32_t> &nbr_c = h->nbr_host;
  if (dmma_build_tiles(h->tiles, nbr_c, cls, kk, ng, ncls, h->ntype, h->pos.empty() ? nullptr : h->pos.data()) != 0) return fail(RSREC_ENOMEM, "cannot allocate the tile tables");
  h->h2d_bytes += (long long)h->tiles.ntiles * (DM_S + 2 + (long long)ng * DM_S) * 4;
  CUDA_TRY(cudaDeviceSynchronize());  // the tile tables went through the legacy stream: make sure they have landed
  h->dirty = false;
  h->dirty_ham = true;
  }  // lattice tables

  // class -> atom type (for the per-type on-site terms lsham / enim)
  std::vector<int32_t> cls_type(ncls);
  for (int c = 0; c < ncls; c++) cls_type[c] = c < h->ntype ? c : h->iz[c - h->ntype] - 1;
  if (!h->d_cls_type) CUDA_TRY(cudaMalloc(&h->d_cls_type, ncls * sizeof(int32_t)));
  CUDA_TRY(cudaMemcpyAsync(h->d_cls_type, cls_type.data(), ncls * sizeof(int32_t), cudaMemcpyHostToDevice, h->st));
  const size_t nblk = (size_t)ncls * nslot, ntb = (size_t)h->ntype * nslot;
  const size_t setn = nblk * HBLK;
  if (!h->ham_on_device) {  // stage the reference's arrays: ee followed by hall = blocks of classes 0..ncls-1
    auto stage = [&](DevBuf &dst, const std::vector<cplx> &ty, const std::vector<cplx> &loc) -> int {
      std::vector<double> st(nblk * BLKD, 0.0);
      if (!ty.empty()) memcpy(st.data(), ty.data(), ntb * BLKD * sizeof(double));
      if (!loc.empty()) memcpy(st.data() + ntb * BLKD, loc.data(), (nblk - ntb) * BLKD * sizeof(double));
      return upload(h, dst, st);
    };
    TRY(stage(h->cBLK, h->ee, h->hall));
    std::vector<double> ls((const double *)h->lsham.data(), (const double *)h->lsham.data() + (size_t)h->ntype * BLKD);
    TRY(upload(h, h->cLS, ls));
    if (h->hoh) {
      TRY(stage(h->cBLKO, h->eeo, h->hallo));
      std::vector<double> en((const double *)h->enim.data(), (const double *)h->enim.data() + (size_t)h->ntype * BLKD);
      TRY(upload(h, h->cENIM, en));
    }
  }
  // HR36 packing on the device (hamiltonian.f90:1553-1667 exports): locham = H_on + lsham summed before the product
  // (1582, 1608); scalar recursion: only the two 9x9 spin-diagonal sub-blocks act, no lsham (recursion.f90:3337-3342)
  const dim3 pg(nslot, ncls);
  auto pack = [&](DevBuf &dst, const double *src, int ncls_src, const double *add, double scale, int spin_diag, int skip_on) -> int {
    TRY(dev_alloc(dst, setn, false));
    k_pack_hr36<<<pg, BLKC, 0, h->st>>>((const double2 *)src, ncls_src, (const double2 *)add, h->d_cls_type, nslot, scale, spin_diag, skip_on, 0, dst.p);
    h->launches++;
    return RSREC_OK;
  };
  TRY(pack(h->Hmain, h->cBLK.p, ncls, h->cLS.p, 1.0, 0, 0));
  TRY(pack(h->Hscalar, h->cBLK.p, ncls, nullptr, 1.0, 1, 0));
  if (h->hoh) {
    TRY(pack(h->Hh, h->cBLK.p, ncls, nullptr, 1.0, 0, 0));
    TRY(pack(h->Hho_neg, h->cBLKO.p, ncls, nullptr, -1.0, 0, 0));
    TRY(dev_alloc(h->Hx, (size_t)ncls * HBLK, false));  // enim + lsham of the class's type
    k_pack_hr36<<<dim3(1, ncls), BLKC, 0, h->st>>>((const double2 *)h->cENIM.p, ncls, (const double2 *)h->cLS.p, h->d_cls_type, 1, 1.0, 0, 0, 1, h->Hx.p);
    h->launches++;
  }
  // velocity operators are type-indexed; the reference skips the site-indexed region entirely
  // (loops start at nmax+1, recursion.f90:603-634): local classes keep zero blocks.
  for (int s = 0; s < 2; s++) {
    if (!h->have_op[s]) continue;
    const std::vector<cplx> &v = s == 0 ? h->v_a : h->v_b, &vo = s == 0 ? h->vo_a : h->vo_b;
    std::vector<double> sv((const double *)v.data(), (const double *)v.data() + ntb * BLKD);
    TRY(upload(h, h->cV[s], sv));
    TRY(pack(s == 0 ? h->Hva : h->Hvb, h->cV[s].p, h->ntype, nullptr, 1.0, 0, 0));
    if (h->hoh) {
      if (!vo.empty()) {  // on-site vo term is commented out in the reference (761): slot 0 stays zero
        std::vector<double> svo((const double *)vo.data(), (const double *)vo.data() + ntb * BLKD);
        TRY(upload(h, h->cVO[s], svo));
        TRY(pack(s == 0 ? h->Hvoa_neg : h->Hvob_neg, h->cVO[s].p, h->ntype, nullptr, -1.0, 0, 1));
      } else {
        TRY(dev_alloc(s == 0 ? h->Hvoa_neg : h->Hvob_neg, setn, false));
        CUDA_TRY(cudaMemsetAsync((s == 0 ? h->Hvoa_neg : h->Hvob_neg).p, 0, setn * sizeof(double), h->st));
      }
    }
  }
  CUDA_TRY(cudaGetLastError());
  // spin-diagonal slots (collinear hopping blocks): the tensor SpMV runs them as two 18x18 products (k_apply_dmma_sd)
  h->sdmap.clear();
  if (!getenv("RSREC_NO_SPIN_DIAG")) {
    // scan(src, classes) -> per-slot "mixes the spins" flags
    const int nsl = nslot;
    TRY(dev_alloc(h->scratch, std::max<size_t>(h->scratch.n, (size_t)8 * nsl), false));
    int *d_fl = (int *)h->scratch.p;
    CUDA_TRY(cudaMemsetAsync(d_fl, 0, (size_t)8 * nsl * sizeof(int), h->st));
    auto scan = [&](int row, const double *src, int slots, int classes) {
      if (src && classes > 0) k_sd_scan<<<dim3(slots, classes), BLKC, 0, h->st>>>((const double2 *)src, slots, d_fl + row * nsl);
    };
    scan(0, h->cBLK.p, nsl, ncls);
    if (h->hoh) scan(1, h->cBLKO.p, nsl, ncls);
    scan(2, h->cLS.p, 1, h->ntype);
    if (h->hoh) scan(3, h->cENIM.p, 1, h->ntype);
    for (int sx = 0; sx < 2; sx++) {
      if (!h->have_op[sx]) continue;
      scan(4 + sx, h->cV[sx].p, nsl, h->ntype);
      if (h->hoh && !(sx == 0 ? h->vo_a : h->vo_b).empty()) scan(6 + sx, h->cVO[sx].p, nsl, h->ntype);
    }
    std::vector<int> fl((size_t)8 * nsl);
    CUDA_TRY(cudaMemcpyAsync(fl.data(), d_fl, fl.size() * sizeof(int), cudaMemcpyDeviceToHost, h->st));
    CUDA_TRY(cudaStreamSynchronize(h->st));
    auto diag = [&](int row, bool extra_mix0) {
      std::vector<unsigned char> v(nsl);
      for (int m = 0; m < nsl; m++) v[m] = !(fl[(size_t)row * nsl + m] || (m == 0 && extra_mix0));
      return v;
    };
    const bool ls_mix = fl[(size_t)2 * nsl] != 0;
    h->sdmap[h->Hmain.p] = diag(0, ls_mix);                                   // on-site slot carries lsham
    h->sdmap[h->Hscalar.p] = std::vector<unsigned char>(nsl, 1);              // masked to the spin blocks by construction
    if (h->hoh) {
      h->sdmap[h->Hh.p] = diag(0, false);
      h->sdmap[h->Hho_neg.p] = diag(1, false);
      h->sdmap[h->Hx.p] = std::vector<unsigned char>(1, !(fl[(size_t)3 * nsl] || ls_mix));
    }
    for (int sx = 0; sx < 2; sx++) {
      if (!h->have_op[sx]) continue;
      h->sdmap[(sx == 0 ? h->Hva : h->Hvb).p] = diag(4 + sx, false);
      if (h->hoh) h->sdmap[(sx == 0 ? h->Hvoa_neg : h->Hvob_neg).p] = diag(6 + sx, false);   // all-zero set: diagonal too
    }
  }
  // HS18 twins: [block][kind][spin][18 rows][20 k] from the HR36 blocks (kind 0 = spin-diagonal part, 1 = coupling part)
  for (auto &kv : h->sdmap) {
    const size_t nb = (size_t)ncls * kv.second.size();
    DevBuf &dst = h->hs18[kv.first];
    TRY(dev_alloc(dst, nb * 2 * SDH, false));
    k_pack_hs18<<<(unsigned)nb, 256, 0, h->st>>>(kv.first, dst.p, (int)nb);
    h->launches++;
  }
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaStreamSynchronize(h->st));
  h->dirty_ham = false;
  return RSREC_OK;
}

static int ensure_units(H *h, int nunits) {
  if (h->units_cap >= nunits) return RSREC_OK;
  if (h->d_si) { cudaFree(h->d_si); cudaFree(h->d_sj); cudaFree(h->d_as); cudaFree(h->d_bs); }
  CUDA_TRY(cudaMalloc(&h->d_si, nunits * sizeof(int32_t)));
  CUDA_TRY(cudaMalloc(&h->d_sj, nunits * sizeof(int32_t)));
  CUDA_TRY(cudaMalloc(&h->d_as, nunits * 2 * sizeof(double)));
  CUDA_TRY(cudaMalloc(&h->d_bs, nunits * 2 * sizeof(double)));
  h->units_cap = nunits;
  return RSREC_OK;
}

static int nctas_for(const H *h, int nunits) {
  int n = (h->sms * 6 + nunits - 1) / nunits;
  return std::max(1, std::min(n, h->kk));
}

// Device memory the work vectors may use: what is free plus what the vectors with index < nreuse already hold (get_vec
// re-uses those in place); vectors with a higher index are leftovers of an earlier call with a larger working set (the
// Kubo left-vector store, for instance) and are released first.  10% headroom.
static double vec_memory(H *h, int nreuse) {
  for (size_t i = nreuse; i < h->vecs.size(); i++) dev_free(h->vecs[i]);
  size_t fr = 0, tot = 0;
  cudaMemGetInfo(&fr, &tot);
  size_t held = 0;
  for (size_t i = 0; i < h->vecs.size() && (int)i < nreuse; i++) held += h->vecs[i].n * sizeof(double);
  return 0.9 * (double)(fr + held);
}
// work vectors of one recursion call: psi, pmn / psi0, psi1 (+ hpsi on the tensor pipeline's Lanczos, + the hoh scratch)
// The 18x18 reduction of a Lanczos step (A = sum psi^H H psi) runs inside the SpMV kernel (north star: SpMV + three-term
// update + dot products in one pass): one launch and one block vector (hpsi) less, same speed within noise on configs 1-3
// (DESIGN.md 4).  RSREC_NO_FUSED_GRAM=1 restores the separate k_gram_dmma launch (A/B switch), as does the S = 8 geometry.
static bool fused_gram(const H *h) { return fam1(h) && h->fuse_lanczos && dmma_apply_geom() == 4; }
// The same fusion for chebyshev_recur_ll (D1, D2 inside the SpMV kernel) exists but is OFF by default: both kernels are bound
// by the FP64 tensor pipe, the stand-alone Gram kernel keeps it 96 % busy, and the fused form measured 27.8 ms per step
// against 26.4 ms at 10^6 sites (profiles/r02b_*).  RSREC_FUSED_CHEB=1 turns it on.
static bool fused_cheb(const H *h) { return fam1(h) && h->fuse_cheb && dmma_apply_geom() == 4; }
// Pipelined step (kernels_dmma.cuh, k_rotortho_dmma): H is applied to the unnormalised residual while k_lz_eig runs on a second
// stream, rotate + orthogonalise are one pass.  Needs the fused Gram products.  One more block vector per unit (W = H R).
// It pays where a step is bound by launch and single-CTA latencies, i.e. on the small lattices of the reference's own cases
// (5984 sites: 4.21 -> 3.67 ms, 3838 sites: 3.42 -> 2.82 ms; 4 units of 5984 sites: 12.6 -> 12.1 ms); on batches that fill the
// GPU the square root it hides is a negligible part of a step (8 x 5984 sites: 23.97 vs 23.76 ms; profiles/r02r_*).
// RSREC_LZ_PIPELINE=1 / 0 forces it on / off.
#define LZ_PIPELINE_MAX_SITES 32768
static bool lz_pipelined(const H *h, int nunits) {
  const char *env = getenv("RSREC_LZ_PIPELINE");  // read per call (tests switch it)
  if (!fused_gram(h)) return false;
  if (env && *env) return atoi(env) != 0;
  return (long long)nunits * h->kk <= LZ_PIPELINE_MAX_SITES;
}
// (the pipelined form holds W = H R; sized for it whenever a batch could take it)
static int lanczos_nvec(const H *h, bool diag) { return 2 + ((fam1(h) && (!fused_gram(h) || lz_pipelined(h, 1))) ? 1 : 0) + ((h->hoh && !diag) ? 1 : 0); }
static int cheb_nvec(const H *h) { return 2 + (h->hoh ? 1 : 0); }
// largest unit batch whose `nvec` work vectors (+ extra bytes per unit: histories, a non-resident g0) fit
static int unit_batch(H *h, int nunits, int nvec, size_t extra_bytes_per_unit = 0) {
  const double avail = vec_memory(h, 4);  // vecs[0..3] are the recursion drivers' own
  const size_t per = vstride(h) * sizeof(double) * nvec + extra_bytes_per_unit;
  int nb = (int)std::max(1.0, std::floor(avail / (double)per));
  if (const char *f = getenv("RSREC_UNIT_BATCH")) nb = std::max(1, std::min(nb, atoi(f)));  // tests: force small batches
  return std::min(nunits, nb);
}
// how many single block vectors fit next to everything else that is allocated (no clamping: 0 means none)
static long long vectors_that_fit(H *h) {
  size_t fr = 0, tot = 0;
  cudaMemGetInfo(&fr, &tot);
  size_t held = 0;
  for (auto &v : h->vecs) held += v.n * sizeof(double);
  return (long long)std::floor(0.9 * (double)(fr + held) / (double)(vstride(h) * sizeof(double)));
}

// ---- operator application: out = epilogue( H src ) for a unit batch ---------------------------------------
enum OpKind { OP_HAM = 0, OP_SCALAR = 1, OP_VELO_A = 2, OP_VELO_B = 3, OP_HAM_NOHOH = 4 /* ham_vec_matmul even when hoh */ };

// Active-region plan for site-started recursions.  level(site) = number of operator applications after which the
// site can be non-zero (start sites: 0); a tile is processed by the k-th application iff its smallest level <= k.
// This is exactly the set the reference tracks with izero/idum/irlist (recursion.f90:1604-1636); skipped tiles hold
// exact zeros, so results are unchanged.
static int plan_build(H *h, int nunits, const int32_t *site_i, const int32_t *site_j) {
  // the plan depends on the neighbour table and the start sites only: the SCF loop repeats the same call every iteration
  {
    std::vector<int32_t> key((size_t)2 * nunits);
    for (int u = 0; u < nunits; u++) { key[2 * u] = site_i[u]; key[2 * u + 1] = site_j ? site_j[u] : 0; }
    if (h->plan.on && h->plan.nunits == nunits && h->plan.key == key && fam1(h)) { h->plan.level = 0; return RSREC_OK; }
    h->plan.key.swap(key);
  }
  h->plan.on = false;
  if (!fam1(h) || (size_t)nunits * h->kk > (size_t)64 << 20) return RSREC_OK;
  const int kk = h->kk, nt = h->tiles.ntiles;
  if (h->radj_off.empty()) {  // reverse adjacency (CSR): radj[a] = sites i that gather from a (i != a), built on first use
    const int ng = h->ncols;
    const std::vector<int32_t> &nbr = h->nbr_host;
    h->radj_off.assign(kk + 1, 0);
    for (int j = 1; j < ng; j++)
      for (int i = 0; i < kk; i++) { const int a = nbr[(size_t)j * kk + i]; if (a < kk) h->radj_off[a + 1]++; }
    for (int a = 0; a < kk; a++) h->radj_off[a + 1] += h->radj_off[a];
    h->radj.resize(h->radj_off[kk]);
    std::vector<int32_t> fill(h->radj_off.begin(), h->radj_off.end() - 1);
    for (int j = 1; j < ng; j++)
      for (int i = 0; i < kk; i++) { const int a = nbr[(size_t)j * kk + i]; if (a < kk) h->radj[fill[a]++] = i; }
  }
  std::vector<int32_t> order((size_t)nunits * nt), level(kk), tl(nt), frontier, next;
  std::vector<std::vector<int32_t>> cum(nunits), bcum(nunits);
  const int nb = (kk + DM_S - 1) / DM_S;
  std::vector<int32_t> border((size_t)nunits * nb), bl(nb);
  int maxlevel = 0;
  for (int u = 0; u < nunits; u++) {
    std::fill(level.begin(), level.end(), INT32_MAX);
    frontier.clear();
    for (int s : {site_i[u] - 1, site_j ? site_j[u] - 1 : -1})
      if (s >= 0 && level[s] != 0) { level[s] = 0; frontier.push_back(s); }
    for (int L = 1; !frontier.empty(); L++) {
      next.clear();
      for (int a : frontier)
        for (int e = h->radj_off[a]; e < h->radj_off[a + 1]; e++) {
          const int i = h->radj[e];
          if (level[i] == INT32_MAX) { level[i] = L; next.push_back(i); }
        }
      frontier.swap(next);
    }
    int ml = 0;
    for (int t = 0; t < nt; t++) {
      int m = INT32_MAX;
      for (int k = 0; k < DM_S; k++) { const int s = h->tiles.h_sites[(size_t)t * DM_S + k]; if (s < kk) m = std::min(m, level[s]); }
      tl[t] = m;
      if (m != INT32_MAX) ml = std::max(ml, m);
    }
    maxlevel = std::max(maxlevel, ml);
    // counting sort by level (stable: tile order inside a level is kept for L2 locality)
    std::vector<int32_t> c(ml + 2, 0);
    for (int t = 0; t < nt; t++) if (tl[t] != INT32_MAX) c[tl[t] + 1]++;
    for (int L = 0; L <= ml; L++) c[L + 1] += c[L];
    cum[u].assign(c.begin() + 1, c.end());  // cum[L] = tiles with level <= L
    std::vector<int32_t> pos(c.begin(), c.end() - 1);
    int32_t *o = order.data() + (size_t)u * nt;
    for (int t = 0; t < nt; t++) if (tl[t] != INT32_MAX) o[pos[tl[t]]++] = t;
    // contiguous blocks of 8 sites, same counting sort
    for (int b = 0; b < nb; b++) {
      int m = INT32_MAX;
      for (int k = 0; k < DM_S && b * DM_S + k < kk; k++) m = std::min(m, level[b * DM_S + k]);
      bl[b] = m;
    }
    // a contiguous block can lie deeper than every class tile (tiles collect far-apart sites of one class): its own maximum
    int mlb = 0;
    for (int b = 0; b < nb; b++) if (bl[b] != INT32_MAX) mlb = std::max(mlb, bl[b]);
    maxlevel = std::max(maxlevel, mlb);
    std::vector<int32_t> bc(mlb + 2, 0);
    for (int b = 0; b < nb; b++) if (bl[b] != INT32_MAX) bc[bl[b] + 1]++;
    for (int L = 0; L <= mlb; L++) bc[L + 1] += bc[L];
    bcum[u].assign(bc.begin() + 1, bc.end());
    std::vector<int32_t> bpos(bc.begin(), bc.end() - 1);
    int32_t *bo = border.data() + (size_t)u * nb;
    for (int b = 0; b < nb; b++) if (bl[b] != INT32_MAX) bo[bpos[bl[b]]++] = b;
  }
  std::vector<int32_t> counts((size_t)(maxlevel + 1) * nunits), bcounts((size_t)(maxlevel + 1) * nunits);
  for (int L = 0; L <= maxlevel; L++)
    for (int u = 0; u < nunits; u++) {
      counts[(size_t)L * nunits + u] = cum[u][std::min<size_t>(L, cum[u].size() - 1)];
      bcounts[(size_t)L * nunits + u] = bcum[u][std::min<size_t>(L, bcum[u].size() - 1)];
    }
  auto &pl = h->plan;
  if (pl.border_cap < border.size()) { if (pl.d_border) cudaFree(pl.d_border); CUDA_TRY(cudaMalloc(&pl.d_border, border.size() * 4)); pl.border_cap = border.size(); }
  if (pl.bcounts_cap < bcounts.size()) { if (pl.d_bcounts) cudaFree(pl.d_bcounts); CUDA_TRY(cudaMalloc(&pl.d_bcounts, bcounts.size() * 4)); pl.bcounts_cap = bcounts.size(); }
  CUDA_TRY(cudaMemcpyAsync(pl.d_border, border.data(), border.size() * 4, cudaMemcpyHostToDevice, h->st));
  CUDA_TRY(cudaMemcpyAsync(pl.d_bcounts, bcounts.data(), bcounts.size() * 4, cudaMemcpyHostToDevice, h->st));
  h->h2d_bytes += (long long)((border.size() + bcounts.size()) * 4);
  pl.nblocks = nb;
  if (pl.order_cap < order.size()) { if (pl.d_order) cudaFree(pl.d_order); CUDA_TRY(cudaMalloc(&pl.d_order, order.size() * 4)); pl.order_cap = order.size(); }
  if (pl.counts_cap < counts.size()) { if (pl.d_counts) cudaFree(pl.d_counts); CUDA_TRY(cudaMalloc(&pl.d_counts, counts.size() * 4)); pl.counts_cap = counts.size(); }
  CUDA_TRY(cudaMemcpyAsync(pl.d_order, order.data(), order.size() * 4, cudaMemcpyHostToDevice, h->st));
  CUDA_TRY(cudaMemcpyAsync(pl.d_counts, counts.data(), counts.size() * 4, cudaMemcpyHostToDevice, h->st));
  CUDA_TRY(cudaStreamSynchronize(h->st));
  h->h2d_bytes += (long long)((order.size() + counts.size()) * 4);
  pl.on = true; pl.level = 0; pl.maxlevel = maxlevel; pl.nunits = nunits;
  return RSREC_OK;
}

// the active contiguous-block list at the plan's current level (null when no plan is active)
static void plan_blocks(const H *h, int nunits, const int32_t **border, const int32_t **bcnt, int *nblocks) {
  *border = nullptr; *bcnt = nullptr; *nblocks = 0;
  static const bool off = getenv("RSREC_NO_BLOCK_PLAN") != nullptr;  // A/B switch for measurements
  if (!off && h->plan.on && h->plan.nunits == nunits && fam1(h)) {
    *border = h->plan.d_border;
    *bcnt = h->plan.d_bcounts + (size_t)h->plan.level * nunits;
    *nblocks = h->plan.nblocks;
  }
}
static int launch_apply_inner(H *h, ApplyParams &p, int nunits, int nctas);
static int launch_apply(H *h, ApplyParams &p, int nunits, int nctas) {
  if (!h->profile) return launch_apply_inner(h, p, nunits, nctas);
  if (h->prof_used == h->prof_events.size()) {
    cudaEvent_t a, b;
    CUDA_TRY(cudaEventCreate(&a));
    CUDA_TRY(cudaEventCreate(&b));
    h->prof_events.push_back({a, b});
  }
  auto &ev = h->prof_events[h->prof_used++];
  CUDA_TRY(cudaEventRecord(ev.first, h->st));
  int rc = launch_apply_inner(h, p, nunits, nctas);
  CUDA_TRY(cudaEventRecord(ev.second, h->st));
  return rc;
}
static int launch_apply_inner(H *h, ApplyParams &p, int nunits, int nctas) {
  if (fam1(h)) {
    if (!dmma_supported(p)) return fail(RSREC_EINVAL, "internal: operator application not supported by the tensor pipeline");
    const int32_t *order = nullptr, *cnt = nullptr;
    if (h->plan.on && h->plan.nunits == nunits) {  // this application reaches level+1
      h->plan.level = std::min(h->plan.level + 1, h->plan.maxlevel);
      order = h->plan.d_order;
      cnt = h->plan.d_counts + (size_t)h->plan.level * nunits;
    }
    auto sdl = [](const void *ctx, const double *Hset, int slot, const double **hs18) -> bool {
      const H *hh = (const H *)ctx;
      auto ith = hh->hs18.find(Hset);
      *hs18 = ith != hh->hs18.end() ? ith->second.p : nullptr;
      auto itf = hh->sdmap.find(Hset);
      return itf != hh->sdmap.end() && slot < (int)itf->second.size() && itf->second[slot];
    };
    int nparts = 0;
    if (dmma_launch_apply(h->tiles, p, nunits, h->sms - h->sm_reserve, h->st, &h->launches, order, cnt, sdl, h, &h->sd_launches, &nparts) != 0)
      return fail(RSREC_ECUDA, std::string("k_apply_dmma launch failed: ") + cudaGetErrorString(cudaGetLastError()));
    h->last_parts = (p.epi == EPI_CHEB || p.epi == EPI_HOP_GRAM) ? nparts : 0;  // one partial slot per CTA of the launch
    return RSREC_OK;
  }
  h->last_parts = nctas;
  dim3 grid(nctas, nunits);
  k_apply_simt<<<grid, SIMT_THREADS, 0, h->st>>>(p);
  h->launches++;
  CUDA_TRY(cudaGetLastError());
  return RSREC_OK;
}

// out = epi( Op * in ).  tmp: scratch vector for the two-pass hoh forms.  part: gram partial buffer or null.
static int apply_op(H *h, OpKind op, const double *in, double *out, const double *prev, double *tmp, int epi,
                    double a, double b, int nunits, int nctas, double *part) {
  ApplyParams p;
  memset(&p, 0, sizeof(p));
  p.kk = h->kk; p.nslot_h = h->nslot; p.ngather = h->ncols; p.vstride = vstride(h);
  p.nbr = h->d_nbr; p.cls = h->d_cls; p.in = in; p.prev = prev; p.out = out; p.a = a; p.b = b; p.epi = epi; p.part = part;
  p.out2 = h->out2;
  const bool hoh = h->hoh && op != OP_SCALAR && op != OP_HAM_NOHOH;
  if (!hoh) {
    const double *Hs = (op == OP_HAM || op == OP_HAM_NOHOH) ? h->Hmain.p : op == OP_SCALAR ? h->Hscalar.p : op == OP_VELO_A ? h->Hva.p : h->Hvb.p;
    p.g[0] = GatherTerm{Hs, in, 0};
    p.ngterms = 1;
    return launch_apply(h, p, nunits, nctas);
  }
  // pass A: tmp = h * in  (first sweep of the hoh routines, e.g. recursion.f90:1455-1477)
  ApplyParams pa = p;
  pa.g[0] = GatherTerm{h->Hh.p, in, 0};
  pa.ngterms = 1; pa.epi = EPI_STORE; pa.out = tmp; pa.out2 = nullptr; pa.part = nullptr; pa.prev = nullptr;
  TRY(launch_apply(h, pa, nunits, nctas));
  if (op == OP_HAM) {
    // pass B: acc = -(h o)*tmp + (e_nu + l.s)*in + tmp   (H = h - hoh + e_nu + l.s, recursion.f90:1543)
    p.g[0] = GatherTerm{h->Hho_neg.p, tmp, 0};
    p.ngterms = 1; p.Hx = h->Hx.p; p.srcx = in; p.addend = tmp;
  } else {
    // velo_hoh_vec_matmul: out = v*in - sum_{nb>=2} vo(nb) * (ee*in)(nn)   (recursion.f90:704-778)
    p.g[0] = GatherTerm{op == OP_VELO_A ? h->Hva.p : h->Hvb.p, in, 0};
    p.g[1] = GatherTerm{op == OP_VELO_A ? h->Hvoa_neg.p : h->Hvob_neg.p, tmp, 1};
    p.ngterms = 2;
  }
  return launch_apply(h, p, nunits, nctas);
}

// part[unit][cta][0] = sum_sites X^H Y  (first factor conjugated); xs/ys: doubles between units (0 = shared vector)
static int launch_gram_strided(H *h, const double *X, size_t xs, const double *Y, size_t ys, int nunits, int nctas,
                               double *part) {
  if (fam1(h)) {
    // many units in one launch (Kubo contraction): fewer CTAs per unit, the grid is filled by the unit dimension
    const int ctas = std::max(1, std::min(dmma_gram_ctas(h->kk, h->sms), (2 * h->sms) / nunits));
    const int32_t *bo, *bc; int nbk;
    plan_blocks(h, nunits, &bo, &bc, &nbk);
    if (dmma_launch_gram(Y, ys, X, xs, 0, h->kk, nunits, ctas, part, h->st, &h->launches, bo, bc, nbk) != 0)
      return fail(RSREC_ECUDA, std::string("k_gram_dmma launch failed: ") + cudaGetErrorString(cudaGetLastError()));
    h->last_parts = ctas;
    return RSREC_OK;
  }
  const int ctas = std::max(1, std::min(nctas, (6 * h->sms + nunits - 1) / nunits));
  dim3 grid(ctas, nunits);
  k_gram_simt<<<grid, SIMT_THREADS, 0, h->st>>>(X, Y, h->kk, xs, ys, part);
  h->launches++;
  h->last_parts = ctas;
  CUDA_TRY(cudaGetLastError());
  return RSREC_OK;
}
static int launch_gram(H *h, const double *X, const double *Y, int nunits, int nctas, double *part) {
  return launch_gram_strided(h, X, vstride(h), Y, vstride(h), nunits, nctas, part);
}
static size_t part_doubles(const H *h, int nunits, int nctas) {
  const int simt = std::max(1, std::min(nctas, (6 * h->sms + nunits - 1) / nunits));
  const int dm = std::max(1, std::min(dmma_gram_ctas(h->kk, h->sms), (2 * h->sms) / nunits));
  const int fused = std::max(h->tiles.ntiles > 0 ? dmma_apply_grid(h->tiles, h->sms, nunits) : 2 * h->sms, dmma_rmul_ctas(h->kk, h->sms, nunits));
  return (size_t)nunits * std::max(std::max(std::max(simt, dm), fused), std::min(nctas, nctas_for(h, nunits))) * 2 * BLKD;
}
static int launch_reduce(H *h, int nunits, int /*nctas*/, int mode, double *d0, double *d1, size_t dstride,
                         const double *m0, const double *m1, double *hist0 = nullptr, size_t hstride = 0) {
  k_reduce_parts<<<dim3((2 * BLKD + RP_ELEMS - 1) / RP_ELEMS, nunits), RP_ELEMS * RP_GROUPS, 0, h->st>>>(h->part.p, h->last_parts, mode, d0, d1, dstride, m0, m1, hist0, hstride);
  h->launches++;
  CUDA_TRY(cudaGetLastError());
  return RSREC_OK;
}

static int upload_units(H *h, int nunits, const int32_t *site_i, const int32_t *site_j, const cplx *asign,
                        const cplx *bsign) {
  TRY(ensure_units(h, nunits));
  std::vector<int32_t> sj(nunits, 0);
  std::vector<double> as(2 * nunits), bs(2 * nunits);
  for (int u = 0; u < nunits; u++) {
    if (site_i[u] < 1 || site_i[u] > h->kk) return fail(RSREC_EINVAL, "start site out of range");
    if (site_j) {
      if (site_j[u] < 0 || site_j[u] > h->kk) return fail(RSREC_EINVAL, "start site j out of range");
      sj[u] = site_j[u];
    }
    as[2 * u] = asign ? asign[u].re : 1.0; as[2 * u + 1] = asign ? asign[u].im : 0.0;
    bs[2 * u] = bsign ? bsign[u].re : 1.0; bs[2 * u + 1] = bsign ? bsign[u].im : 0.0;
  }
  CUDA_TRY(cudaMemcpyAsync(h->d_si, site_i, nunits * sizeof(int32_t), cudaMemcpyHostToDevice, h->st)); h->h2d_bytes += (long long)(nunits * sizeof(int32_t));
  CUDA_TRY(cudaMemcpyAsync(h->d_sj, sj.data(), nunits * sizeof(int32_t), cudaMemcpyHostToDevice, h->st)); h->h2d_bytes += (long long)(nunits * sizeof(int32_t));
  CUDA_TRY(cudaMemcpyAsync(h->d_as, as.data(), 2 * nunits * sizeof(double), cudaMemcpyHostToDevice, h->st)); h->h2d_bytes += (long long)(2 * nunits * sizeof(double));
  CUDA_TRY(cudaMemcpyAsync(h->d_bs, bs.data(), 2 * nunits * sizeof(double), cudaMemcpyHostToDevice, h->st)); h->h2d_bytes += (long long)(2 * nunits * sizeof(double));
  CUDA_TRY(cudaStreamSynchronize(h->st));  // host staging vectors go out of scope
  return RSREC_OK;
}

// ------------------------------------------------------------------------------------------------------------
// Block / scalar Lanczos driver: crecal_b (recursion.f90:1873-1973) for a batch of units, entirely on the device.
static int lanczos_batch(H *h, int nunits, int lld, bool diag, double *a_host /*18x18xlld complex per unit*/,
                         double *b2_host) {
  const int nctas = nctas_for(h, nunits);
  double *psi, *pmn, *tmp = nullptr;
  TRY(get_vec(h, 0, nunits, &psi));
  TRY(get_vec(h, 1, nunits, &pmn));
  if (h->hoh && !diag) TRY(get_vec(h, 2, nunits, &tmp));
  double *hpsi = nullptr;
  const bool fused = fused_gram(h);
  const bool pipelined = lz_pipelined(h, nunits) && lld >= 2;
  if (fam1(h) && (!fused || pipelined)) TRY(get_vec(h, 3, nunits, &hpsi));  // hpsi of the staged form / W = H R of the pipelined one
  const size_t pd = part_doubles(h, nunits, nctas);
  TRY(dev_alloc(h->part, pipelined ? 2 * pd : pd, false));  // pipelined: second half = the B^2 partials (read by the side stream)
  double *part2 = h->part.p + pd;
  TRY(dev_alloc(h->A, (size_t)nunits * BLKD, false));
  TRY(dev_alloc(h->B, (size_t)nunits * BLKD, false));
  TRY(dev_alloc(h->Bi, (size_t)2 * nunits * BLKD, false));  // two: the pipelined step needs B^-1 of the previous level too
  TRY(dev_alloc(h->B2, (size_t)nunits * BLKD, false));
  const size_t hs = (size_t)lld * BLKD;  // history stride per unit (doubles)
  TRY(dev_alloc(h->ahist, (size_t)nunits * hs, false));
  TRY(dev_alloc(h->b2hist, (size_t)nunits * hs, false));
  TRY(dev_alloc(h->bhist, (size_t)nunits * hs, false));
  TRY(zero_vec(h, psi, nunits));
  TRY(zero_vec(h, pmn, nunits));
  if (hpsi) TRY(zero_vec(h, hpsi, nunits));  // tiles the active-region plan skips must read as zeros
  if (tmp) TRY(zero_vec(h, tmp, nunits));
  CUDA_TRY(cudaMemsetAsync(h->ahist.p, 0, (size_t)nunits * hs * sizeof(double), h->st));
  CUDA_TRY(cudaMemsetAsync(h->b2hist.p, 0, (size_t)nunits * hs * sizeof(double), h->st));
  CUDA_TRY(cudaMemsetAsync(h->bhist.p, 0, (size_t)nunits * hs * sizeof(double), h->st));
  k_init_site_start<<<nunits, 32, 0, h->st>>>(psi, vstride(h), h->d_si, h->d_sj, h->d_as, h->d_bs, nunits);
  k_set_identity<<<nunits, 64, 0, h->st>>>(h->b2hist.p, hs, nunits);  // b2temp_b(:,:,1) = I
  k_set_identity<<<nunits, 64, 0, h->st>>>(h->bhist.p, hs, nunits);   // and its square root
  h->launches += 3;
  const dim3 grid(nctas, nunits);
  if (pipelined) {
    if (!h->st2) {
      CUDA_TRY(cudaStreamCreateWithFlags(&h->st2, cudaStreamNonBlocking));
      CUDA_TRY(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
      CUDA_TRY(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
    }
    const OpKind op = diag ? OP_SCALAR : OP_HAM;
    cudaStream_t st2 = h->st2;
    double *W = hpsi, *G = h->B2.p;
    double *prev = psi, *cur = pmn;  // R_{n-1} (R_{-1} = the start vector) and R_n: the state of the pipelined recursion
    int np2 = 0;
    const int32_t *bo, *bc; int nbk;
    // step 0 in the reference's order: pmn = H psi_0 (pmn was zero), A_0, R_0 = pmn - psi_0 A_0, partials of B_1^2
    {
      PhaseScope ph_(h, PH_HPSI);
      TRY(apply_op(h, op, psi, pmn, pmn, tmp, EPI_HOP_GRAM, 1.0, 0.0, nunits, nctas, h->part.p));
      TRY(launch_reduce(h, nunits, nctas, diag ? 2 : 0, h->A.p, nullptr, BLKD, nullptr, nullptr, h->ahist.p, hs));
    }
    {
      PhaseScope ph_(h, PH_ORTHO);
      plan_blocks(h, nunits, &bo, &bc, &nbk);
      if (dmma_launch_rmul(RM_ORTHO, psi, pmn, nullptr, h->A.p, nullptr, BLKD, h->kk, vstride(h), nunits, h->sms, h->st, &h->launches, bo, bc, nbk, part2, &np2) != 0)
        return fail(RSREC_ECUDA, std::string("k_rmul_dmma launch failed: ") + cudaGetErrorString(cudaGetLastError()));
    }
    for (int ll = 0; ll < lld - 1; ll++) {
      // B_{ll+1}^2 -> history, B, B^-1 on the side stream ...
      CUDA_TRY(cudaEventRecord(h->ev_fork, h->st));
      CUDA_TRY(cudaStreamWaitEvent(st2, h->ev_fork, 0));
      {
        PhaseScope ph_(h, PH_BNEXT, st2);
        // 128 kB of (unused) dynamic shared memory: no SpMV CTA (>= 93 kB) fits beside this CTA, so the square root has an SM --
        // and its FP64 pipe, which a co-resident DMMA kernel saturates (the iteration took 93 us instead of 25) -- to itself
        k_lz_eig<<<nunits, BLKC, LZ_EIG_ALONE_SMEM, st2>>>(nullptr, BLKD, h->b2hist.p + (size_t)(ll + 1) * BLKD, hs, h->B.p,
                                              h->Bi.p + (size_t)(ll & 1) * nunits * BLKD,
                                              BLKD, diag ? 1 : 0, h->sqrt_method, h->bhist.p + (size_t)(ll + 1) * BLKD,
                                              part2, np2);
        h->launches++;
      }
      CUDA_TRY(cudaEventRecord(h->ev_join, st2));
      if (ll == lld - 2) { CUDA_TRY(cudaStreamWaitEvent(h->st, h->ev_join, 0)); break; }  // the last level needs no further A
      // ... while H is applied to the unnormalised residual: W = H R, G = sum R^H W (one SM per eig CTA left free)
      {
        PhaseScope ph_(h, PH_HPSI);
        h->sm_reserve = std::min(h->sms / 2, nunits);
        const int rc = apply_op(h, op, cur, W, nullptr, tmp, EPI_HOP_GRAM, 1.0, 0.0, nunits, nctas, h->part.p);
        h->sm_reserve = 0;
        TRY(rc);
        TRY(launch_reduce(h, nunits, nctas, diag ? 2 : 0, G, nullptr, BLKD, nullptr, nullptr));
      }
      CUDA_TRY(cudaStreamWaitEvent(h->st, h->ev_join, 0));
      // A_{ll+1} = B^-1 G B^-1 ; R_{ll+1} = W B^-1 - R_{ll-1} (B_prev^-1 B) - R_ll B^-1 A_{ll+1} (over R_{ll-1}) ; partials of B_{ll+2}^2
      {
        PhaseScope ph_(h, PH_ORTHO);
        plan_blocks(h, nunits, &bo, &bc, &nbk);
        np2 = dmma_rmul_ctas(h->kk, h->sms, nunits);
        const double *Bi_cur = h->Bi.p + (size_t)(ll & 1) * nunits * BLKD;
        // B^-1 of the previous level; level 0 is the identity (b2temp_b(:,:,1) = I: slot 0 of the B history)
        const double *Bi_prev = ll == 0 ? h->bhist.p : h->Bi.p + (size_t)((ll - 1) & 1) * nunits * BLKD;
        if (dmma_launch_rotortho(prev, cur, W, h->B.p, Bi_cur, Bi_prev, ll == 0 ? hs : (size_t)BLKD, G, BLKD, h->A.p,
                                 h->ahist.p + (size_t)(ll + 1) * BLKD, hs, diag ? 1 : 0, h->kk, vstride(h), nunits, np2, h->st,
                                 &h->launches, bo, bc, nbk, part2) != 0)
          return fail(RSREC_ECUDA, std::string("k_rotortho_dmma launch failed: ") + cudaGetErrorString(cudaGetLastError()));
        std::swap(prev, cur);
      }
      { PhaseScope ph_(h, PH_ROTATE); }  // '<PSI|B_n+1|PSI>' is part of the merged pass: the label stays in the host's timer tree (calls counted, ~0 ms)
      CUDA_TRY(cudaGetLastError());
    }
  } else
  for (int ll = 0; ll < lld - 1; ll++) {
    // hop_b / hop_b_hoh: pmn = H psi - pmn ; A = sum psi^H H psi
    {
    PhaseScope ph_(h, PH_HPSI);
    if (fused) {    // one kernel: pmn = H psi - pmn and the per-CTA partials of A = sum psi^H (H psi)
      TRY(apply_op(h, diag ? OP_SCALAR : OP_HAM, psi, pmn, pmn, tmp, EPI_HOP_GRAM, 1.0, 0.0, nunits, nctas, h->part.p));
    } else if (fam1(h)) {  // tensor-pipe SpMV: hpsi = H psi, pmn = hpsi - pmn; A = sum psi^H hpsi on the tensor pipe too
      h->out2 = hpsi;
      TRY(apply_op(h, diag ? OP_SCALAR : OP_HAM, psi, pmn, pmn, tmp, EPI_HOP, 1.0, 0.0, nunits, nctas, nullptr));
      h->out2 = nullptr;
      TRY(launch_gram(h, psi, hpsi, nunits, nctas, h->part.p));
    } else {
      TRY(apply_op(h, diag ? OP_SCALAR : OP_HAM, psi, pmn, pmn, tmp, EPI_HOP, 1.0, 0.0, nunits, nctas, h->part.p));
    }
    TRY(launch_reduce(h, nunits, nctas, diag ? 2 : 0, h->A.p, nullptr, BLKD, nullptr, nullptr, h->ahist.p + (size_t)ll * BLKD, hs));
    }
    // pmn -= psi A ; B2 = sum pmn^H pmn
    {
    PhaseScope ph_(h, PH_ORTHO);
    if (fam1(h)) {
      const int32_t *bo, *bc; int nbk;
      plan_blocks(h, nunits, &bo, &bc, &nbk);
      if (fused) {  // pmn -= psi A and the per-CTA partials of B^2 = sum pmn^H pmn in one pass over the tiles
        int np = 0;
        if (dmma_launch_rmul(RM_ORTHO, psi, pmn, nullptr, h->A.p, nullptr, BLKD, h->kk, vstride(h), nunits, h->sms, h->st, &h->launches, bo, bc, nbk, h->part.p, &np) != 0)
          return fail(RSREC_ECUDA, std::string("k_rmul_dmma launch failed: ") + cudaGetErrorString(cudaGetLastError()));
        h->last_parts = np;
      } else {
      if (dmma_launch_rmul(RM_ORTHO, psi, pmn, nullptr, h->A.p, nullptr, BLKD, h->kk, vstride(h), nunits, h->sms, h->st, &h->launches, bo, bc, nbk) != 0)
        return fail(RSREC_ECUDA, std::string("k_rmul_dmma launch failed: ") + cudaGetErrorString(cudaGetLastError()));
      TRY(launch_gram(h, pmn, pmn, nunits, nctas, h->part.p));
      }
    } else {
      k_lz_ortho_simt<<<grid, SIMT_THREADS, 0, h->st>>>(psi, pmn, hpsi, h->A.p, BLKD, h->kk, vstride(h), h->part.p);
      h->last_parts = nctas;
      h->launches++;
    }
    }
    // B2 -> history slot ll+1, B, B^-1
    {
    PhaseScope ph_(h, PH_BNEXT);
    // the fixed-order sum of the B^2 partials is the prologue of k_lz_eig (one launch less per step)
    k_lz_eig<<<nunits, BLKC, 0, h->st>>>(nullptr, BLKD, h->b2hist.p + (size_t)(ll + 1) * BLKD, hs, h->B.p, h->Bi.p,
                                         BLKD, diag ? 1 : 0, h->sqrt_method, h->bhist.p + (size_t)(ll + 1) * BLKD,
                                         h->part.p, h->last_parts);
    h->launches++;
    }
    // psi = pmn B^-1 ; pmn = psi_old B
    {
    PhaseScope ph_(h, PH_ROTATE);
    if (fam1(h)) {
      const int32_t *bo, *bc; int nbk;
      plan_blocks(h, nunits, &bo, &bc, &nbk);
      if (dmma_launch_rmul(RM_ROTATE, psi, pmn, nullptr, h->Bi.p, h->B.p, BLKD, h->kk, vstride(h), nunits, h->sms, h->st, &h->launches, bo, bc, nbk) != 0)
        return fail(RSREC_ECUDA, std::string("k_rmul_dmma launch failed: ") + cudaGetErrorString(cudaGetLastError()));
    } else {
      k_lz_rotate_simt<<<grid, SIMT_THREADS, 0, h->st>>>(psi, pmn, h->B.p, h->Bi.p, BLKD, h->kk, vstride(h));
      h->launches++;
    }
    }
    CUDA_TRY(cudaGetLastError());
  }
  if (h->phase_on && getenv("RSREC_LZ_TIMELINE")) {  // diagnostics: start / end of every timed phase of this call, us from the first
    cudaStreamSynchronize(h->st);
    if (h->st2) cudaStreamSynchronize(h->st2);
    for (size_t i = 0; i < h->phase_used; i++) {
      float ta = 0.f, tb = 0.f;
      cudaEventElapsedTime(&ta, h->phase_events[0].a, h->phase_events[i].a);
      cudaEventElapsedTime(&tb, h->phase_events[0].a, h->phase_events[i].b);
      fprintf(stderr, "timeline phase %d  %9.1f -> %9.1f us\n", h->phase_events[i].phase, 1e3 * ta, 1e3 * tb);
    }
  }
  if (a_host) { CUDA_TRY(cudaMemcpyAsync(a_host, h->ahist.p, (size_t)nunits * hs * sizeof(double), cudaMemcpyDeviceToHost, h->st)); h->d2h_bytes += (long long)((size_t)nunits * hs * sizeof(double)); }
  if (b2_host) { CUDA_TRY(cudaMemcpyAsync(b2_host, h->b2hist.p, (size_t)nunits * hs * sizeof(double), cudaMemcpyDeviceToHost, h->st)); h->d2h_bytes += (long long)((size_t)nunits * hs * sizeof(double)); }
  CUDA_TRY(cudaStreamSynchronize(h->st));
  return RSREC_OK;
}

// ------------------------------------------------------------------------------------------------------------
// Chebyshev session: cheb_0th_mom, cheb_1st_mom(_hoh), chebyshev_recur_ll(_hoh).
static int cheb_begin_common(H *h, int nunits, int lld, double a, double b) {
  if (lld < 0) return fail(RSREC_EINVAL, "lld must be >= 0");
  if (a == 0.0) return fail(RSREC_EINVAL, "a_scale must be non-zero");
  auto &c = h->cheb;
  c.nunits = nunits; c.lld = lld; c.a = a; c.b = b; c.done = 0; c.i0 = 0; c.i1 = 1;
  c.nctas = nctas_for(h, nunits);
  TRY(dev_alloc(h->part, part_doubles(h, nunits, c.nctas), false));
  TRY(dev_alloc(h->mu, (size_t)nunits * (2 * lld + 2) * BLKD, false));
  CUDA_TRY(cudaMemsetAsync(h->mu.p, 0, (size_t)nunits * (2 * lld + 2) * BLKD * sizeof(double), h->st));
  return RSREC_OK;
}
// psi0 (vecs[0]) holds the start vector = psiref.  Computes mu(1), mu(2) and psi1.
static int cheb_first_moments(H *h) {
  auto &c = h->cheb;
  double *p0, *p1, *tmp = nullptr;
  TRY(get_vec(h, 0, c.nunits, &p0));
  TRY(get_vec(h, 1, c.nunits, &p1));
  if (h->hoh) TRY(get_vec(h, 2, c.nunits, &tmp));
  const size_t ms = (size_t)(2 * c.lld + 2) * BLKD;
  {
    PhaseScope ph_(h, PH_MOM0);
    TRY(launch_gram(h, p0, p0, c.nunits, c.nctas, h->part.p));
    TRY(launch_reduce(h, c.nunits, c.nctas, 0, h->mu.p, nullptr, ms, nullptr, nullptr));
  }
  {
    PhaseScope ph_(h, PH_MOM1);
    TRY(apply_op(h, OP_HAM, p0, p1, nullptr, tmp, EPI_HAM, c.a, c.b, c.nunits, c.nctas, nullptr));
    TRY(launch_gram(h, p0, p1, c.nunits, c.nctas, h->part.p));
    TRY(launch_reduce(h, c.nunits, c.nctas, 0, h->mu.p + BLKD, nullptr, ms, nullptr, nullptr));
  }
  c.active = true;
  return RSREC_OK;
}
static int cheb_steps(H *h, int nsteps) {
  auto &c = h->cheb;
  if (!c.active) return fail(RSREC_EINVAL, "no Chebyshev session (call rsrec_cheb_begin_* first)");
  if (c.done + nsteps > c.lld) return fail(RSREC_EINVAL, "more steps than lld requested");
  double *tmp = nullptr;
  if (h->hoh) TRY(get_vec(h, 2, c.nunits, &tmp));
  const size_t ms = (size_t)(2 * c.lld + 2) * BLKD;
  for (int s = 0; s < nsteps; s++) {
    PhaseScope ph_(h, PH_MOMN);
    const int ll = c.done + 1;
    double *p0 = h->vecs[c.i0].p, *p1 = h->vecs[c.i1].p;
    // psi2 = 2 (H psi1 - b psi1)/a - psi0, written over psi0; D1 = sum psi1^H psi1, D2 = sum psi2^H psi1
    if (fused_cheb(h)) {  // SpMV + three-term update + D1, D2 partials in one kernel
      TRY(apply_op(h, OP_HAM, p1, p0, p0, tmp, EPI_CHEB, c.a, c.b, c.nunits, c.nctas, h->part.p));
    } else if (fam1(h)) {
      TRY(apply_op(h, OP_HAM, p1, p0, p0, tmp, EPI_CHEB_NOGRAM, c.a, c.b, c.nunits, c.nctas, nullptr));
      const int gctas = std::max(1, std::min(dmma_gram_ctas(h->kk, h->sms), (2 * h->sms) / c.nunits));
      const int32_t *bo, *bc; int nbk;
      plan_blocks(h, c.nunits, &bo, &bc, &nbk);
      if (dmma_launch_gram(p1, vstride(h), p0, vstride(h), 1, h->kk, c.nunits, gctas, h->part.p, h->st, &h->launches, bo, bc, nbk) != 0)
        return fail(RSREC_ECUDA, std::string("k_gram_dmma launch failed: ") + cudaGetErrorString(cudaGetLastError()));
      h->last_parts = gctas;
    } else {
      TRY(apply_op(h, OP_HAM, p1, p0, p0, tmp, EPI_CHEB, c.a, c.b, c.nunits, c.nctas, h->part.p));
    }
    const int nparts = h->last_parts;
    // mu(2ll+1) = 2 D1 - mu(1), mu(2ll+2) = 2 D2 - mu(2)
    TRY(launch_reduce(h, c.nunits, nparts, 1, h->mu.p + (size_t)(2 * ll) * BLKD,
                      h->mu.p + (size_t)(2 * ll + 1) * BLKD, ms, h->mu.p, h->mu.p + BLKD));
    std::swap(c.i0, c.i1);
    c.done++;
  }
  return RSREC_OK;
}
static int cheb_finish(H *h, cplx *mu_n) {
  auto &c = h->cheb;
  const size_t n = (size_t)c.nunits * (2 * c.lld + 2) * BLKD;
  CUDA_TRY(cudaMemcpyAsync(mu_n, h->mu.p, n * sizeof(double), cudaMemcpyDeviceToHost, h->st)); h->d2h_bytes += (long long)(n * sizeof(double));
  CUDA_TRY(cudaStreamSynchronize(h->st));
  c.active = false;
  // the reference's divergence guard (recursion.f90:2594): sum(real(mu(:,:,2ll+2))) > 1000 -> fatal
  const double *m = (const double *)mu_n;
  for (int u = 0; u < c.nunits; u++)
    for (int k = 3; k < 2 * c.lld + 2; k += 2) {
      double s = 0.0;
      const double *blk = m + ((size_t)u * (2 * c.lld + 2) + k) * BLKD;
      for (int e = 0; e < BLKC; e++) s += blk[2 * e];
      if (!(s <= 1000.0))
        return fail(RSREC_EDIVERGED, "Chebyshev moments did not converge. Check energy limits energy_min and energy_max");
    }
  return RSREC_OK;
}

// ============================================================================================================
extern "C" {

const char *rsrec_last_error(void) { return g_err.c_str(); }
int rsrec_version(void) { return 100; }
int rsrec_compiled_arch(void) { return 100; }
int rsrec_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int rsrec_create(rsrec_handle *out, int device, int kk, int ncols, int nslot, int ntype, int nmax) {
  if (!out) return fail(RSREC_EINVAL, "null handle pointer");
  *out = nullptr;
  if (kk < 1 || ncols < 1 || nslot < ncols || ntype < 1 || nmax < 0 || nmax > kk)
    return fail(RSREC_EINVAL, "rsrec_create: inconsistent sizes (need kk>=1, 1<=ncols<=nslot, ntype>=1, 0<=nmax<=kk)");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(RSREC_ECUDA, std::string("no CUDA device available (this library has no CPU path): ") + cudaGetErrorString(e));
  }
  if (device < 0 || device >= ndev) return fail(RSREC_EINVAL, "device ordinal out of range");
  CUDA_TRY(cudaSetDevice(device));
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) return fail(RSREC_ECUDA, std::string("kernels are built for sm_100a only; device is ") + prop.name);
  H *h = new H();
  h->dev = device; h->kk = kk; h->ncols = ncols; h->nslot = nslot; h->ntype = ntype; h->nmax = nmax;
  h->ncls = ntype + nmax; h->sms = prop.multiProcessorCount;
  if (const char *f = getenv("RSREC_KERNEL_FAMILY")) h->family = atoi(f);
  if (const char *f = getenv("RSREC_SQRT_METHOD")) h->sqrt_method = atoi(f);
  if (getenv("RSREC_NO_FUSED_GRAM")) h->fuse_lanczos = false;
  if (getenv("RSREC_FUSED_CHEB")) h->fuse_cheb = true;
  CUDA_TRY(cudaStreamCreateWithFlags(&h->st, cudaStreamNonBlocking));
  if (dmma_configure() != 0 || kubo_configure() != 0 || kdiag_configure() != 0 || post_configure() != 0 || ham_configure() != 0 || bands_configure() != 0) {
    cudaStreamDestroy(h->st);
    delete h;
    return fail(RSREC_ECUDA, "cannot reserve shared memory for the DMMA kernels");
  }
  *out = h;
  return RSREC_OK;
}

int rsrec_destroy(rsrec_handle h) {
  if (!h) return RSREC_OK;
  cudaSetDevice(h->dev);
  cudaStreamSynchronize(h->st);
  for (auto &v : h->vecs) dev_free(v);
  DevBuf *bufs[] = {&h->Hmain, &h->Hh, &h->Hho_neg, &h->Hx, &h->Hscalar, &h->Hva, &h->Hvb, &h->Hvoa_neg, &h->Hvob_neg,
                    &h->part, &h->A, &h->B, &h->Bi, &h->B2, &h->mu, &h->ahist, &h->b2hist, &h->bhist, &h->scratch};
  for (auto b : bufs) dev_free(*b);
  for (auto &b : h->post) dev_free(b);
  dev_free(h->g0all); dev_free(h->bands_y); dev_free(h->bands_out);
  dmma_free_tiles(h->tiles);
  if (h->d_nbr) cudaFree(h->d_nbr);
  if (h->d_cls) cudaFree(h->d_cls);
  if (h->d_cls_type) cudaFree(h->d_cls_type);
  for (auto &kv : h->hs18) dev_free(kv.second);
  { DevBuf *cb[] = {&h->cBLK, &h->cBLKO, &h->cLS, &h->cENIM, &h->cOBARM, &h->cV[0], &h->cV[1], &h->cVO[0], &h->cVO[1], &h->gBLK, &h->gBLKO, &h->gENIM}; for (auto b : cb) dev_free(*b); }
  if (h->plan.d_order) cudaFree(h->plan.d_order);
  if (h->plan.d_counts) cudaFree(h->plan.d_counts);
  if (h->plan.d_border) cudaFree(h->plan.d_border);
  if (h->plan.d_bcounts) cudaFree(h->plan.d_bcounts);
  if (h->d_si) { cudaFree(h->d_si); cudaFree(h->d_sj); cudaFree(h->d_as); cudaFree(h->d_bs); }
  for (auto &ev : h->prof_events) { cudaEventDestroy(ev.first); cudaEventDestroy(ev.second); }
  for (auto &ev : h->phase_events) { cudaEventDestroy(ev.a); cudaEventDestroy(ev.b); }
  if (h->comm && nccl_api()->dl) nccl_api()->CommDestroy(h->comm);
  dev_free(h->comm_buf); dev_free(h->comm_res);
  if (h->st2) { cudaStreamDestroy(h->st2); cudaEventDestroy(h->ev_fork); cudaEventDestroy(h->ev_join); }
  cudaStreamDestroy(h->st);
  delete h;
  return RSREC_OK;
}

int rsrec_set_kernel_family(rsrec_handle h, int family) {
  if (!h || family < 0 || family > 1) return fail(RSREC_EINVAL, "bad kernel family");
  h->family = family;
  return RSREC_OK;
}

int rsrec_set_fusion(rsrec_handle h, int lanczos, int cheb) {
  if (!h) return fail(RSREC_EINVAL, "null handle");
  if (lanczos >= 0) h->fuse_lanczos = lanczos != 0;
  if (cheb >= 0) h->fuse_cheb = cheb != 0;
  return RSREC_OK;
}

int rsrec_set_lattice(rsrec_handle h, const int32_t *nn, const int32_t *iz) {
  if (!h || !nn || !iz) return fail(RSREC_EINVAL, "rsrec_set_lattice: null argument");
  h->nn.assign(nn, nn + (size_t)h->kk * h->ncols);
  h->iz.assign(iz, iz + h->kk);
  h->have_lat = true; h->dirty = true;
  return RSREC_OK;
}

int rsrec_set_hamiltonian(rsrec_handle h, const cplx *ee, const cplx *eeo, const cplx *hall, const cplx *hallo,
                          const cplx *lsham, const cplx *enim, int hoh) {
  if (!h || !ee || !lsham) return fail(RSREC_EINVAL, "rsrec_set_hamiltonian: ee and lsham are required");
  if (h->nmax > 0 && !hall) return fail(RSREC_EINVAL, "rsrec_set_hamiltonian: hall is required when nmax > 0");
  if (hoh && (!eeo || !enim || (h->nmax > 0 && !hallo)))
    return fail(RSREC_EINVAL, "rsrec_set_hamiltonian: eeo, enim (and hallo when nmax > 0) are required when hoh is set");
  const size_t nt = (size_t)BLKC * h->nslot * h->ntype, nl = (size_t)BLKC * h->nslot * h->nmax, n1 = (size_t)BLKC * h->ntype;
  h->ee.assign(ee, ee + nt);
  h->lsham.assign(lsham, lsham + n1);
  if (h->nmax > 0) h->hall.assign(hall, hall + nl); else h->hall.clear();
  if (hoh) {
    h->eeo.assign(eeo, eeo + nt);
    h->enim.assign(enim, enim + n1);
    if (h->nmax > 0) h->hallo.assign(hallo, hallo + nl); else h->hallo.clear();
  }
  h->hoh = hoh ? 1 : 0;
  h->have_ham = true; h->dirty_ham = true; h->ham_on_device = false; h->have_glob = false;
  return RSREC_OK;
}

int rsrec_set_operator(rsrec_handle h, int slot, const cplx *v_op, const cplx *vo_op) {
  if (!h || !v_op || (slot != 'a' && slot != 'b')) return fail(RSREC_EINVAL, "rsrec_set_operator: slot must be 'a' or 'b'");
  const size_t nt = (size_t)BLKC * h->nslot * h->ntype;
  const int s = slot == 'a' ? 0 : 1;
  (s == 0 ? h->v_a : h->v_b).assign(v_op, v_op + nt);
  if (vo_op) (s == 0 ? h->vo_a : h->vo_b).assign(vo_op, vo_op + nt); else (s == 0 ? h->vo_a : h->vo_b).clear();
  h->have_op[s] = true; h->dirty_ham = true;
  return RSREC_OK;
}

int rsrec_lanczos_block(rsrec_handle h, int nunits, const int32_t *site_i, const int32_t *site_j, const cplx *asign,
                        const cplx *bsign, int lld, cplx *a_b, cplx *b2_b) {
  if (!h || nunits < 0 || !a_b || !b2_b || lld < 1 || (nunits > 0 && !site_i)) return fail(RSREC_EINVAL, "rsrec_lanczos_block: bad argument");
  if (nunits == 0) return RSREC_OK;
  CUDA_TRY(cudaSetDevice(h->dev));
  { HostScope hs_(h, HP_TABLES); TRY(ensure_ready(h)); }
  const int ub = unit_batch(h, nunits, lanczos_nvec(h, false), (size_t)3 * lld * BLKD * sizeof(double));
  for (int u0 = 0; u0 < nunits; u0 += ub) {
    const int n = std::min(ub, nunits - u0);
    {
      HostScope hs_(h, HP_PLAN);
      TRY(upload_units(h, n, site_i + u0, site_j ? site_j + u0 : nullptr, asign ? asign + u0 : nullptr, bsign ? bsign + u0 : nullptr));
      TRY(plan_build(h, n, site_i + u0, site_j ? site_j + u0 : nullptr));
    }
    HostScope hs_(h, HP_RECUR);
    TRY(lanczos_batch(h, n, lld, false, (double *)(a_b + (size_t)u0 * lld * BLKC), (double *)(b2_b + (size_t)u0 * lld * BLKC)));
  }
  return RSREC_OK;
}

int rsrec_lanczos_scalar(rsrec_handle h, int nunits, const int32_t *sites, int lld, double *a, double *b2) {
  if (!h || nunits < 0 || !a || !b2 || lld < 1 || (nunits > 0 && !sites)) return fail(RSREC_EINVAL, "rsrec_lanczos_scalar: bad argument");
  if (nunits == 0) return RSREC_OK;
  CUDA_TRY(cudaSetDevice(h->dev));
  TRY(ensure_ready(h));
  // The 18 independent scalar recursions of a site (one per start orbital, recursion.f90:3499-3520) are the 18
  // columns of one block vector with diagonal A and B: run them together and keep Re(diag).
  const int ub = unit_batch(h, nunits, lanczos_nvec(h, true), (size_t)3 * lld * BLKD * sizeof(double));
  std::vector<double> ah((size_t)ub * lld * BLKD), bh((size_t)ub * lld * BLKD);
  for (int u0 = 0; u0 < nunits; u0 += ub) {
    const int n = std::min(ub, nunits - u0);
    TRY(upload_units(h, n, sites + u0, nullptr, nullptr, nullptr));
    TRY(plan_build(h, n, sites + u0, nullptr));
    TRY(lanczos_batch(h, n, lld, true, ah.data(), bh.data()));
    for (int u = 0; u < n; u++)
      for (int l = 0; l < NB; l++)
        for (int ll = 0; ll < lld; ll++) {
          const size_t src = ((size_t)u * lld + ll) * BLKD + 2 * (l + NB * l);
          a[(size_t)ll + (size_t)lld * (l + (size_t)NB * (u0 + u))] = ah[src];
          b2[(size_t)ll + (size_t)lld * (l + (size_t)NB * (u0 + u))] = bh[src];
        }
  }
  return RSREC_OK;
}

int rsrec_zsqr(rsrec_handle h, cplx *b2_b, int lld, int na) {
  if (!h || !b2_b || lld < 0 || na < 0) return fail(RSREC_EINVAL, "rsrec_zsqr: bad argument");
  const size_t nmat = (size_t)lld * na;
  if (nmat == 0) return RSREC_OK;
  CUDA_TRY(cudaSetDevice(h->dev));
  TRY(dev_alloc(h->scratch, nmat * BLKD, false));
  CUDA_TRY(cudaMemcpyAsync(h->scratch.p, b2_b, nmat * BLKD * sizeof(double), cudaMemcpyHostToDevice, h->st)); h->h2d_bytes += (long long)(nmat * BLKD * sizeof(double));
  k_zsqr<<<(unsigned)nmat, BLKC, 0, h->st>>>(h->scratch.p);
  h->launches++;
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpyAsync(b2_b, h->scratch.p, nmat * BLKD * sizeof(double), cudaMemcpyDeviceToHost, h->st)); h->d2h_bytes += (long long)(nmat * BLKD * sizeof(double));
  CUDA_TRY(cudaStreamSynchronize(h->st));
  return RSREC_OK;
}

int rsrec_cheb_begin_sites(rsrec_handle h, int nunits, const int32_t *site_i, const int32_t *site_j, const cplx *asign,
                           const cplx *bsign, int lld, double a, double b) {
  if (!h || nunits < 1 || !site_i) return fail(RSREC_EINVAL, "rsrec_cheb_begin_sites: bad argument");
  CUDA_TRY(cudaSetDevice(h->dev));
  TRY(ensure_ready(h));
  if (unit_batch(h, nunits, cheb_nvec(h)) < nunits) return fail(RSREC_ENOMEM, "unit batch does not fit in device memory");
  TRY(upload_units(h, nunits, site_i, site_j, asign, bsign));
  TRY(plan_build(h, nunits, site_i, site_j));
  TRY(cheb_begin_common(h, nunits, lld, a, b));
  double *p0, *p1;
  TRY(get_vec(h, 0, nunits, &p0));
  TRY(get_vec(h, 1, nunits, &p1));
  TRY(zero_vec(h, p0, nunits));
  TRY(zero_vec(h, p1, nunits));
  if (h->hoh) { double *t2; TRY(get_vec(h, 2, nunits, &t2)); TRY(zero_vec(h, t2, nunits)); }
  k_init_site_start<<<nunits, 32, 0, h->st>>>(p0, vstride(h), h->d_si, h->d_sj, h->d_as, h->d_bs, nunits);
  h->launches++;
  return cheb_first_moments(h);
}

int rsrec_cheb_begin_random(rsrec_handle h, int nvec, const double *phases, int lld, double a, double b) {
  if (!h || nvec < 1 || !phases) return fail(RSREC_EINVAL, "rsrec_cheb_begin_random: bad argument");
  CUDA_TRY(cudaSetDevice(h->dev));
  TRY(ensure_ready(h));
  if (unit_batch(h, nvec, cheb_nvec(h)) < nvec) return fail(RSREC_ENOMEM, "vector batch does not fit in device memory");
  h->plan.on = false;  // every site is active from the start
  TRY(cheb_begin_common(h, nvec, lld, a, b));
  double *p0, *p1;
  TRY(get_vec(h, 0, nvec, &p0));
  TRY(get_vec(h, 1, nvec, &p1));
  TRY(zero_vec(h, p0, nvec));
  TRY(zero_vec(h, p1, nvec));
  TRY(dev_alloc(h->scratch, (size_t)h->kk * nvec, false));
  CUDA_TRY(cudaMemcpyAsync(h->scratch.p, phases, (size_t)h->kk * nvec * sizeof(double), cudaMemcpyHostToDevice, h->st)); h->h2d_bytes += (long long)((size_t)h->kk * nvec * sizeof(double));
  k_init_random_start<<<h->sms * 4, 256, 0, h->st>>>(p0, vstride(h), h->scratch.p, h->kk, nvec);
  h->launches++;
  return cheb_first_moments(h);
}

int rsrec_cheb_run_steps(rsrec_handle h, int nsteps) {
  if (!h || nsteps < 0) return fail(RSREC_EINVAL, "rsrec_cheb_run_steps: bad argument");
  CUDA_TRY(cudaSetDevice(h->dev));
  return cheb_steps(h, nsteps);
}

int rsrec_cheb_end(rsrec_handle h, cplx *mu_n) {
  if (!h || !mu_n) return fail(RSREC_EINVAL, "rsrec_cheb_end: bad argument");
  if (!h->cheb.active) return fail(RSREC_EINVAL, "no Chebyshev session");
  CUDA_TRY(cudaSetDevice(h->dev));
  return cheb_finish(h, mu_n);
}

int rsrec_cheb_moments(rsrec_handle h, int nunits, const int32_t *site_i, const int32_t *site_j, const cplx *asign,
                       const cplx *bsign, int lld, double a, double b, cplx *mu_n) {
  if (!h || nunits < 0 || !mu_n || (nunits > 0 && !site_i)) return fail(RSREC_EINVAL, "rsrec_cheb_moments: bad argument");
  if (nunits == 0) return RSREC_OK;
  CUDA_TRY(cudaSetDevice(h->dev));
  TRY(ensure_ready(h));
  const int ub = unit_batch(h, nunits, cheb_nvec(h), (size_t)(2 * lld + 2) * BLKD * sizeof(double));
  int rc_all = RSREC_OK;
  for (int u0 = 0; u0 < nunits; u0 += ub) {
    const int n = std::min(ub, nunits - u0);
    TRY(rsrec_cheb_begin_sites(h, n, site_i + u0, site_j ? site_j + u0 : nullptr, asign ? asign + u0 : nullptr,
                               bsign ? bsign + u0 : nullptr, lld, a, b));
    TRY(cheb_steps(h, lld));
    int rc = cheb_finish(h, mu_n + (size_t)u0 * (2 * lld + 2) * BLKC);
    if (rc == RSREC_EDIVERGED) rc_all = rc; else TRY(rc);
  }
  return rc_all;
}

int rsrec_cheb_moments_random(rsrec_handle h, int nvec, const double *phases, int lld, double a, double b, cplx *mu_n) {
  if (!h || nvec < 0 || !mu_n || (nvec > 0 && !phases)) return fail(RSREC_EINVAL, "rsrec_cheb_moments_random: bad argument");
  if (nvec == 0) return RSREC_OK;
  CUDA_TRY(cudaSetDevice(h->dev));
  TRY(ensure_ready(h));
  const int ub = unit_batch(h, nvec, cheb_nvec(h), (size_t)(2 * lld + 2) * BLKD * sizeof(double));
  int rc_all = RSREC_OK;
  for (int u0 = 0; u0 < nvec; u0 += ub) {
    const int n = std::min(ub, nvec - u0);
    TRY(rsrec_cheb_begin_random(h, n, phases + (size_t)u0 * h->kk, lld, a, b));
    TRY(cheb_steps(h, lld));
    int rc = cheb_finish(h, mu_n + (size_t)u0 * (2 * lld + 2) * BLKC);
    if (rc == RSREC_EDIVERGED) rc_all = rc; else TRY(rc);
  }
  return rc_all;
}

// host (18,18,kk) complex <-> device RI36 through the scratch buffer
static int upload_vec(H *h, const cplx *src, double *dst) {
  TRY(dev_alloc(h->scratch, (size_t)h->kk * BLKD, false));
  CUDA_TRY(cudaMemcpyAsync(h->scratch.p, src, (size_t)h->kk * BLKD * sizeof(double), cudaMemcpyHostToDevice, h->st)); h->h2d_bytes += (long long)((size_t)h->kk * BLKD * sizeof(double));
  k_host_to_ri36<<<h->sms * 4, 256, 0, h->st>>>(h->scratch.p, dst, h->kk);
  h->launches++;
  return RSREC_OK;
}
static int download_vec(H *h, const double *src, cplx *dst) {
  TRY(dev_alloc(h->scratch, (size_t)h->kk * BLKD, false));
  k_ri36_to_host<<<h->sms * 4, 256, 0, h->st>>>(src, h->scratch.p, h->kk);
  h->launches++;
  CUDA_TRY(cudaMemcpyAsync(dst, h->scratch.p, (size_t)h->kk * BLKD * sizeof(double), cudaMemcpyDeviceToHost, h->st)); h->d2h_bytes += (long long)((size_t)h->kk * BLKD * sizeof(double));
  CUDA_TRY(cudaStreamSynchronize(h->st));
  return RSREC_OK;
}

int rsrec_ham_vec_matmul(rsrec_handle h, const cplx *psi_in, cplx *psi_out, double a, double b) {
  if (!h || !psi_in || !psi_out || a == 0.0) return fail(RSREC_EINVAL, "rsrec_ham_vec_matmul: bad argument");
  CUDA_TRY(cudaSetDevice(h->dev));
  TRY(ensure_ready(h));
  double *vin, *vout, *tmp = nullptr;
  TRY(get_vec(h, 0, 1, &vin));
  TRY(get_vec(h, 1, 1, &vout));
  if (h->hoh) TRY(get_vec(h, 2, 1, &tmp));
  TRY(upload_vec(h, psi_in, vin));
  h->plan.on = false;
  TRY(apply_op(h, OP_HAM, vin, vout, nullptr, tmp, EPI_HAM, a, b, 1, nctas_for(h, 1), nullptr));
  return download_vec(h, vout, psi_out);
}

int rsrec_velo_vec_matmul(rsrec_handle h, int slot, const cplx *psi_in, cplx *psi_out) {
  if (!h || !psi_in || !psi_out || (slot != 'a' && slot != 'b')) return fail(RSREC_EINVAL, "rsrec_velo_vec_matmul: bad argument");
  if (!h->have_op[slot == 'a' ? 0 : 1]) return fail(RSREC_EINVAL, "rsrec_set_operator has not been called for this slot");
  CUDA_TRY(cudaSetDevice(h->dev));
  TRY(ensure_ready(h));
  double *vin, *vout, *tmp = nullptr;
  TRY(get_vec(h, 0, 1, &vin));
  TRY(get_vec(h, 1, 1, &vout));
  if (h->hoh) TRY(get_vec(h, 2, 1, &tmp));
  TRY(upload_vec(h, psi_in, vin));
  h->plan.on = false;
  TRY(apply_op(h, slot == 'a' ? OP_VELO_A : OP_VELO_B, vin, vout, nullptr, tmp, EPI_STORE, 1.0, 0.0, 1, nctas_for(h, 1), nullptr));
  return download_vec(h, vout, psi_out);
}

// compute_moments_stochastic (recursion.f90:1105-1230): left vectors T_m|r> are stored (like the reference's
// left_vec), the right chain v_a T_n v_b |r> is contracted against all of them after every step.
// mu_nm (host, may be null): every start's moments are downloaded; d_diag (device, may be null): D[t][n][m][l2], the
// diagonals calculate_conductivity_tensor consumes, are kept on the device instead.
static int kubo_moments_impl(rsrec_handle h, int nstart, int start_kind, const int32_t *start_sites, const double *phases,
                             int M, double a, double b, cplx *mu_nm, double *d_diag) {
  if (!h || nstart < 0 || M < 1 || (!mu_nm && !d_diag) || a == 0.0) return fail(RSREC_EINVAL, "rsrec_kubo_moments: bad argument");
  if (start_kind == 0 ? !start_sites : !phases) return fail(RSREC_EINVAL, "rsrec_kubo_moments: missing start data");
  if (!h->have_op[0] || !h->have_op[1]) return fail(RSREC_EINVAL, "rsrec_set_operator must be called for slots 'a' and 'b'");
  if (nstart == 0) return RSREC_OK;
  CUDA_TRY(cudaSetDevice(h->dev));
  TRY(ensure_ready(h));
  const bool gemm = fam1(h);  // tensor pipeline: batched GEMM contraction (kernels_kubo.cuh)
  // only the diagonals mu(l,l,n,m) are wanted: column-wise contraction, 18x fewer flops (k_kubo_diag)
  static const bool no_diag = getenv("RSREC_NO_KUBO_DIAG") != nullptr;
  const bool diag_only = gemm && !mu_nm && d_diag && !no_diag;
  const int M4 = diag_only ? (M + KD_MB - 1) / KD_MB * KD_MB : gemm ? (M + KB_MB - 1) / KB_MB * KB_MB : M;
  // right vectors kept per contraction: 4 for the full-block GEMM; for the diagonal kernel as many 16-blocks as fit
  int nring = gemm ? KB_NR : 0;
  if (diag_only) {
    const long long fit = vectors_that_fit(h) - (M4 + 6);
    nring = (int)std::max<long long>(KD_NR, std::min<long long>((M + KD_NR - 1) / KD_NR * KD_NR, fit / KD_NR * KD_NR));
  }
  if (vectors_that_fit(h) < (long long)M4 + 6 + nring)
    return fail(RSREC_ENOMEM, "rsrec_kubo_moments: the left-vector storage (cond_ll + " + std::to_string(6 + nring) + " block vectors of " +
                                  std::to_string(vstride(h) * sizeof(double) >> 20) + " MiB) does not fit in device memory");
  h->plan.on = false;
  const int nctas = nctas_for(h, 1);
  // vecs: 0 psiref, 1 tmp(hoh), 2 v0, 3 v1, 4 right, 5 spare, 6 left[m] (batched), 7 ring of KB_NR right vectors
  double *psiref, *tmp, *v0, *v1, *right, *spare;
  TRY(get_vec(h, 0, 1, &psiref)); TRY(get_vec(h, 1, 1, &tmp)); TRY(get_vec(h, 2, 1, &v0));
  TRY(get_vec(h, 3, 1, &v1)); TRY(get_vec(h, 4, 1, &right)); TRY(get_vec(h, 5, 1, &spare));
  std::vector<double *> left(M);
  double *leftbuf, *ring = nullptr;  // the reference's left_vec(18,18,kk,cond_ll): one batched allocation, unit m = T_m|r>
  TRY(get_vec(h, 6, M4, &leftbuf));
  for (int m = 0; m < M; m++) left[m] = leftbuf + (size_t)m * vstride(h);
  if (gemm) {
    TRY(get_vec(h, 7, nring, &ring));
    TRY(zero_vec(h, ring, nring));
    if (M4 > M) TRY(zero_vec(h, leftbuf + (size_t)M * vstride(h), M4 - M));  // padding rows of the last left block
    if (diag_only)
      TRY(dev_alloc(h->part, kdiag_part_doubles(M4 / KD_MB, nring / KD_NR, kdiag_chunks(M4 / KD_MB, nring / KD_NR, h->kk, h->sms)), false));
    else
      TRY(dev_alloc(h->part, kubo_part_doubles(M4 / KB_MB, h->kk, h->sms), false));
  } else {
    TRY(dev_alloc(h->part, part_doubles(h, M, nctas), false));
  }
  TRY(dev_alloc(h->mu, (size_t)M * M * BLKD, false));
  for (int s = 0; s < nstart; s++) {
    TRY(zero_vec(h, psiref, 1));
    if (start_kind == 0) {
      TRY(upload_units(h, 1, start_sites + s, nullptr, nullptr, nullptr));
      k_init_site_start<<<1, 32, 0, h->st>>>(psiref, vstride(h), h->d_si, h->d_sj, h->d_as, h->d_bs, 1);
    } else {
      TRY(dev_alloc(h->scratch, (size_t)h->kk, false));
      CUDA_TRY(cudaMemcpyAsync(h->scratch.p, phases + (size_t)s * h->kk, (size_t)h->kk * sizeof(double), cudaMemcpyHostToDevice, h->st)); h->h2d_bytes += (long long)((size_t)h->kk * sizeof(double));
      k_init_random_start<<<h->sms * 4, 256, 0, h->st>>>(psiref, vstride(h), h->scratch.p, h->kk, 1);
    }
    h->launches++;
    // left chain: T_1 = start, T_2 = H~ T_1, T_m = 2 H~ T_{m-1} - T_{m-2}
    CUDA_TRY(cudaMemcpyAsync(left[0], psiref, vstride(h) * sizeof(double), cudaMemcpyDeviceToDevice, h->st));
    if (M > 1) TRY(apply_op(h, OP_HAM, left[0], left[1], nullptr, tmp, EPI_HAM, a, b, 1, nctas, nullptr));
    for (int m = 2; m < M; m++)
      TRY(apply_op(h, OP_HAM, left[m - 1], left[m], left[m - 2], tmp, EPI_CHEB_NOGRAM, a, b, 1, nctas, nullptr));
    // right chain
    TRY(apply_op(h, OP_VELO_B, psiref, v0, nullptr, tmp, EPI_STORE, 1.0, 0.0, 1, nctas, nullptr));
    double *c0 = v0, *c1 = v1, *c2 = spare;  // T_{n-2}, T_{n-1}, scratch
    for (int n = 0; n < M; n++) {
      double *cur;
      if (n == 0) {
        cur = c0;
      } else if (n == 1) {
        TRY(apply_op(h, OP_HAM, c0, c1, nullptr, tmp, EPI_HAM, a, b, 1, nctas, nullptr));
        cur = c1;
      } else {
        TRY(apply_op(h, OP_HAM, c1, c2, c0, tmp, EPI_CHEB_NOGRAM, a, b, 1, nctas, nullptr));
        double *t = c0; c0 = c1; c1 = c2; c2 = t;
        cur = c1;
      }
      // mu_nm_stochastic(:,:,n,m,i) = sum_k left_vec(:,:,k,m)^H right_vec(:,:,k) for all m (1220-1228)
      if (diag_only) {
        // right vectors are collected nring at a time; only the diagonals mu(l,l,n,m) are contracted
        TRY(apply_op(h, OP_VELO_A, cur, ring + (size_t)(n % nring) * vstride(h), nullptr, tmp, EPI_STORE, 1.0, 0.0, 1, nctas, nullptr));
        if (n % nring == nring - 1 || n == M - 1) {
          const int n0 = n - n % nring, filled = n - n0 + 1, nnb = (filled + KD_NR - 1) / KD_NR;
          if (filled < nnb * KD_NR)  // stale vectors of the previous batch in the last 16-block: zero them
            TRY(zero_vec(h, ring + (size_t)filled * vstride(h), nnb * KD_NR - filled));
          if (kdiag_launch(leftbuf, vstride(h), M, ring, vstride(h), nnb, h->kk, n0, h->part.p,
                           (double2 *)d_diag + (size_t)s * M * M * NB, h->sms, h->st, &h->launches) != 0)
            return fail(RSREC_ECUDA, std::string("k_kubo_diag launch failed: ") + cudaGetErrorString(cudaGetLastError()));
        }
      } else if (gemm) {
        // right vectors are collected KB_NR at a time, then contracted against every left vector in one GEMM
        TRY(apply_op(h, OP_VELO_A, cur, ring + (size_t)(n % KB_NR) * vstride(h), nullptr, tmp, EPI_STORE, 1.0, 0.0, 1, nctas, nullptr));
        if (n % KB_NR == KB_NR - 1 || n == M - 1) {
          if (kubo_launch(leftbuf, vstride(h), M, ring, vstride(h), h->kk, n - n % KB_NR, h->part.p, h->mu.p, h->sms, h->st, &h->launches) != 0)
            return fail(RSREC_ECUDA, std::string("k_kubo_gemm launch failed: ") + cudaGetErrorString(cudaGetLastError()));
        }
      } else {
        TRY(apply_op(h, OP_VELO_A, cur, right, nullptr, tmp, EPI_STORE, 1.0, 0.0, 1, nctas, nullptr));
        TRY(launch_gram_strided(h, leftbuf, vstride(h), right, 0, M, nctas, h->part.p));
        TRY(launch_reduce(h, M, nctas, 0, h->mu.p + (size_t)n * BLKD, nullptr, (size_t)M * BLKD, nullptr, nullptr));
      }
    }
    if (d_diag && !diag_only) {
      k_cond_diag<<<grid_for((size_t)M * M * NB, 256, h->sms * 16), 256, 0, h->st>>>((const double2 *)h->mu.p, M, (size_t)M * M,
                                                                                  (double2 *)d_diag + (size_t)s * M * M * NB);
      h->launches++;
    }
    if (mu_nm) { CUDA_TRY(cudaMemcpyAsync(mu_nm + (size_t)s * M * M * BLKC, h->mu.p, (size_t)M * M * BLKD * sizeof(double), cudaMemcpyDeviceToHost, h->st)); h->d2h_bytes += (long long)((size_t)M * M * BLKD * sizeof(double)); }
    CUDA_TRY(cudaStreamSynchronize(h->st));
  }
  return RSREC_OK;
}

int rsrec_kubo_moments(rsrec_handle h, int nstart, int start_kind, const int32_t *start_sites, const double *phases,
                       int M, double a, double b, cplx *mu_nm) {
  if (!mu_nm) return fail(RSREC_EINVAL, "rsrec_kubo_moments: bad argument");
  return kubo_moments_impl(h, nstart, start_kind, start_sites, phases, M, a, b, mu_nm, nullptr);
}

}  // extern "C"

// ============================================================================================================
// Consumers either side of the hot path: terminator, block / Chebyshev Green functions,
// scalar continued fraction, Kubo-Bastin integrand.  d_* functions work on device arrays (so that the fused entry
// points can chain them behind a recursion without a host round trip); the rsrec_* wrappers marshal host arrays.
// Terminator launcher: one warp per chain while the chains fit a few waves of resident warps (the bisections run five
// levels per round there), one thread per chain beyond that.
static void launch_bpopt(H *h, const double *A, const double *RB, ChainLayout lay, int ll, int nchains, double *ainf, double *rbinf,
                         int *ifail, int diag);
#define BPOPT_SMEM_MAX (48 * 1024)  // beyond this (ll > ~45) the chains are read from global memory as before
static void launch_bpopt(H *h, const double *A, const double *RB, ChainLayout lay, int ll, int nchains, double *ainf, double *rbinf,
                         int *ifail, int diag) {
  const size_t wsm = (size_t)4 * 2 * (ll + 1) * sizeof(double);  // 4 warps per CTA
  const char *force = getenv("RSREC_BPOPT_KERNEL");              // "thread" | "warp" (A/B switch)
  const bool warp = force ? force[0] == 'w' : (nchains <= 4 * 32 * h->sms && wsm <= BPOPT_SMEM_MAX);
  if (warp) {
    k_bpopt_warp<<<(nchains + 3) / 4, 128, wsm, h->st>>>(A, RB, lay, ll, nchains, ainf, rbinf, ifail, diag);
  } else if (diag) {
    const size_t sm = bpopt_smem_bytes(ll, 32);
    k_bpopt_diag<<<(nchains + 31) / 32, 32, sm <= BPOPT_SMEM_MAX ? sm : 0, h->st>>>(A, RB, lay, ll, nchains, ainf, rbinf, sm <= BPOPT_SMEM_MAX);
  } else {
    const size_t sm = bpopt_smem_bytes(ll, 64);
    k_bpopt<<<(nchains + 63) / 64, 64, sm <= BPOPT_SMEM_MAX ? sm : 0, h->st>>>(A, RB, lay, ll, nchains, ainf, rbinf, ifail, sm <= BPOPT_SMEM_MAX);
  }
}
static int bgreen_smem(int nw) { return (2 * BG_MAT + nw * BG_WSTRIDE) * (int)sizeof(double2); }
static int post_configure() {
  if (cudaFuncSetAttribute(k_bgreen<BG_WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, bgreen_smem(BG_WARPS)) != cudaSuccess) return -1;
  if (cudaFuncSetAttribute(k_bgreen<BG_WARPS_WIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, bgreen_smem(BG_WARPS_WIDE)) != cudaSuccess) return -1;
  if (cudaFuncSetAttribute(k_lz_eig, cudaFuncAttributeMaxDynamicSharedMemorySize, LZ_EIG_ALONE_SMEM) != cudaSuccess) return -1;
  return 0;
}
static int to_dev(H *h, DevBuf &b, const void *src, size_t ndoubles) {
  TRY(dev_alloc(b, std::max<size_t>(ndoubles, 1), false));
  if (ndoubles) CUDA_TRY(cudaMemcpyAsync(b.p, src, ndoubles * sizeof(double), cudaMemcpyHostToDevice, h->st));
  h->h2d_bytes += (long long)(ndoubles * sizeof(double));
  return RSREC_OK;
}
static int to_host(H *h, void *dst, const double *src, size_t ndoubles) {
  if (ndoubles) CUDA_TRY(cudaMemcpyAsync(dst, src, ndoubles * sizeof(double), cudaMemcpyDeviceToHost, h->st));
  h->d2h_bytes += (long long)(ndoubles * sizeof(double));
  return RSREC_OK;
}
static int grid_for(size_t n, int threads, int cap) { return (int)std::max<size_t>(1, std::min<size_t>((n + threads - 1) / threads, (size_t)cap)); }
// g0 (18,18,nv,na) stays on the device for the bands consumers
static int g0_reserve(H *h, int na, int nv) {
  h->g0_units = 0; h->g0_nv = 0;
  TRY(dev_alloc(h->g0all, std::max<size_t>(1, (size_t)na * nv * BLKD), false));
  h->g0_units = na; h->g0_nv = nv;
  return RSREC_OK;
}
// true when g0 of na units may stay resident: it must leave half of the free device memory to the recursion itself
static bool g0_fits(H *h, int na, int nv) {
  const size_t need = (size_t)na * nv * BLKD * sizeof(double);
  if (const char *cap = getenv("RSREC_G0_RESIDENT_MAX_MB")) return need <= (size_t)atoll(cap) * 1048576;  // test hook / user cap
  if (h->g0all.n * sizeof(double) >= need) return true;
  size_t fr = 0, tot = 0;
  if (cudaMemGetInfo(&fr, &tot) != cudaSuccess) return false;
  return need <= (fr + h->g0all.n * sizeof(double)) / 2;
}
static int g0_keep(H *h, const double *d_src, int na, int nv) {
  if (!g0_fits(h, na, nv)) { h->g0_units = 0; h->g0_nv = 0; return RSREC_OK; }  // too large to keep: the host copy is the only one
  TRY(g0_reserve(h, na, nv));
  CUDA_TRY(cudaMemcpyAsync(h->g0all.p, d_src, (size_t)na * nv * BLKD * sizeof(double), cudaMemcpyDeviceToDevice, h->st));
  return RSREC_OK;
}

// get_terminf (recursion.f90:2092-2138) = get_cinf (324 bpopt chains per unit over the real parts) + fix-ups
static int d_terminf(H *h, const double *d_ab, const double *d_bb, int na, int ll, double *d_ainf, double *d_binf,
                     double *d_a0, double *d_b0, bool diag_only = false) {
  // diag_only: bgreen reads only a_inf(i,i), b_inf(i,i) (green.f90:1243-1262) and a_inf0/b_inf0 average the diagonal, so
  // the callers that do not hand a_inf/b_inf back (block_green, the fused driver) run 18 chains per unit instead of 324
  const ChainLayout lay{BLKC, 2, 2LL * BLKC * ll, 2LL * BLKC};
  const int nch = BLKC * na;
  if (diag_only) {
    CUDA_TRY(cudaMemsetAsync(d_ainf, 0, 2 * (size_t)nch * sizeof(double), h->st));  // a_inf and b_inf are adjacent
    const ChainLayout dl{NB, 2LL * (NB + 1), 2LL * BLKC * ll, 2LL * BLKC};
    launch_bpopt(h, d_ab, d_bb, dl, ll, NB * na, d_ainf, d_binf, nullptr, 1);
  } else {
    launch_bpopt(h, d_ab, d_bb, lay, ll, nch, d_ainf, d_binf, nullptr, 0);
  }
  k_terminf_fix<<<na, BLKC, 0, h->st>>>(d_ainf, d_binf, d_a0, d_b0);
  h->launches += 2;
  CUDA_TRY(cudaGetLastError());
  return RSREC_OK;
}
// bgreen (green.f90:1191-1339) for na units sharing the mesh; a_inf, b_inf (18,18,na); g (18,18,nv,na) zeroed here
static int d_bgreen(H *h, const double *d_ab, const double *d_bb, int ll, int na, const double *d_ene, int nv, int ie0,
                    int ie_len, const double *d_ainf, const double *d_binf, double eta_re, double eta_im, int sym_term,
                    double *d_g) {
  CUDA_TRY(cudaMemsetAsync(d_g, 0, (size_t)na * nv * BLKD * sizeof(double), h->st));
  if (ie_len <= 0) return RSREC_OK;
  // rounds of resident warps each geometry needs (2 CTAs per SM); the wide one pays 10-17 % per chain for its spills (measured)
  auto rounds = [&](int nw) { return ((long long)((ie_len + nw - 1) / nw) * na + 2LL * h->sms - 1) / (2LL * h->sms); };
  const char *force = getenv("RSREC_BGREEN_WARPS");
  const bool wide = force ? atoi(force) == BG_WARPS_WIDE : 117 * rounds(BG_WARPS_WIDE) < 100 * rounds(BG_WARPS);
  if (wide)
    k_bgreen<BG_WARPS_WIDE><<<dim3((ie_len + BG_WARPS_WIDE - 1) / BG_WARPS_WIDE, na), BG_WARPS_WIDE * 32, bgreen_smem(BG_WARPS_WIDE), h->st>>>(
        (const double2 *)d_ab, (const double2 *)d_bb, ll, d_ene, nv, ie0, ie_len, d_ainf, d_binf, eta_re, eta_im, sym_term,
        (double2 *)d_g);
  else
    k_bgreen<BG_WARPS><<<dim3((ie_len + BG_WARPS - 1) / BG_WARPS, na), BG_WARPS * 32, bgreen_smem(BG_WARPS), h->st>>>(
        (const double2 *)d_ab, (const double2 *)d_bb, ll, d_ene, nv, ie0, ie_len, d_ainf, d_binf, eta_re, eta_im, sym_term,
        (double2 *)d_g);
  h->launches++;
  CUDA_TRY(cudaGetLastError());
  return RSREC_OK;
}
// jackson_kernel / lorentz_kernel (math.f90:1641-1677): `real(ll)` is default (single) real in the reference
static void host_jackson_kernel(int n, std::vector<double> &k) {
  k.resize(n);
  const float bign = (float)n;
  for (int ll = 1; ll <= n; ll++) {
    const double theta = PI_RP * ((float)ll - 1.0f) / (bign + 1.0f);
    k[ll - 1] = (double)(bign - ((float)ll - 1.0f) + 1.0f) * std::cos(theta) + std::sin(theta) / std::tan(PI_RP / (bign + 1.0f));
    k[ll - 1] = k[ll - 1] / (bign + 1.0f);
  }
}
static void host_lorentz_kernel(int n, double lambda, std::vector<double> &k) {
  k.resize(n);
  for (int ll = 1; ll <= n; ll++) {
    const float q = ((float)ll - 1.0f) / (float)n;
    const double theta = lambda * (double)(1.0f - q);
    k[ll - 1] = std::sinh(theta) / std::sinh(lambda);
  }
}
// chebyshev_green (green.f90:1030-1108): d_mu (18,18,nk,na) -> d_mug (weighted moments), d_g0 (18,18,nv,na)
static int d_cheb_green(H *h, const double *d_mu, int na, int lld, const double *d_ene, int nv, double emin, double emax,
                        double *d_mug, double *d_g0) {
  const int nk = 2 * lld + 2;
  const double a = (emax - emin) / (2 - 0.3), b = (emax + emin) / 2;
  std::vector<double> kern;
  host_jackson_kernel(nk, kern);
  TRY(to_dev(h, h->post[10], kern.data(), nk));
  const size_t total = (size_t)na * nk * BLKC;
  k_cheb_weight<<<grid_for(total, 256, h->sms * 8), 256, 0, h->st>>>((const double2 *)d_mu, h->post[10].p, nk, total, (double2 *)d_mug);
  k_cheb_green<<<dim3((nv + CG_EB - 1) / CG_EB, na), CG_THREADS, 0, h->st>>>((const double2 *)d_mug, nk, d_ene, nv, a, b, (double2 *)d_g0);
  h->launches += 2;
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaStreamSynchronize(h->st));  // `kern` goes out of scope
  return RSREC_OK;
}
// density for nch = na*nmdir chains-of-18 (density_of_states.f90:248-372): d_a, d_b2 (lld,18,na,nmdir) -> d_td (18,nv,na,nmdir)
static int d_density(H *h, const double *d_a, const double *d_b2, int lld, int na, int nmdir, const double *d_ene, int nv,
                     const double *d_dw, const double *d_cs, double *d_td) {
  const int nch = na * nmdir, nchain = NB * nch;
  const size_t n = (size_t)lld * nchain;
  TRY(dev_alloc(h->post[8], n + 2 * (size_t)nchain, false));
  double *sq = h->post[8].p, *am1 = sq + n, *bm1 = am1 + nchain;
  k_sqrt_array<<<grid_for(n, 256, h->sms * 8), 256, 0, h->st>>>(d_b2, sq, n);
  const ChainLayout lay{(long long)nchain, (long long)lld, 0, 1};
  launch_bpopt(h, d_a, sq, lay, lld, nchain, am1, bm1, nullptr, 0);
  const size_t total = (size_t)NB * nv * nch;
  k_density<<<grid_for(total, 256, h->sms * 16), 256, 0, h->st>>>(d_a, d_b2, lld, nch, na, am1, bm1, d_ene, nv, d_dw, d_cs, d_td);
  h->launches += 3;
  CUDA_TRY(cudaGetLastError());
  return RSREC_OK;
}
// calculate_gamma_nm + the integrand of calculate_conductivity_tensor (conductivity.f90:158-306).
// d_mu: (18,18,M,M,nloop) device moments; outputs on the host.
// d_mu: (18,18,M,M,nloop) device moments, or null when d_diag_in already holds D[t][n][m][l2].
static int d_cond_integrand(H *h, const double *d_mu, const double *d_diag_in, int M, int nloop, const double *ene, int nv,
                            double emin, double emax, int per_type, cplx *integrand, cplx *integrand_at, bool reduce_ranks = false) {
  const double a = (emax - emin) / (2 - 0.3), b = (emax + emin) / 2, de = emax - emin;
  const double factor = 16 / (PI_RP * (de * de));
  std::vector<double> sk;
  host_lorentz_kernel(M, 6.0, sk);
  sk[0] *= 0.5;  // weights(1) = 0.5
  const int nchunk = std::max(1, std::min(M, (2 * h->sms) / std::max(1, ((nv + CD_THREADS - 1) / CD_THREADS) * nloop)));
  const size_t nvM = (size_t)nv * M;
  TRY(to_dev(h, h->post[0], ene, nv));
  TRY(to_dev(h, h->post[1], sk.data(), M));
  TRY(dev_alloc(h->post[2], 2 * nvM, false));  // CN
  TRY(dev_alloc(h->post[3], 2 * nvM, false));  // CM
  TRY(dev_alloc(h->post[4], nvM + nv, false)); // TS, inv
  if (d_mu) TRY(dev_alloc(h->post[5], 2 * (size_t)nloop * M * M * NB, false));   // D
  const double *d_D = d_mu ? h->post[5].p : d_diag_in;
  TRY(dev_alloc(h->post[6], 2 * (size_t)nloop * nchunk * NB * nv, false));      // partials
  TRY(dev_alloc(h->post[7], 2 * (size_t)NB * nv * (1 + nloop), false));          // integrand, integrand_at
  double *TS = h->post[4].p, *inv = TS + nvM;
  double *d_int = h->post[7].p, *d_at = d_int + 2 * (size_t)NB * nv;
  k_cond_tables<<<(nv + 127) / 128, 128, 0, h->st>>>(h->post[0].p, nv, M, a, b, h->post[1].p, (double2 *)h->post[2].p, (double2 *)h->post[3].p, TS, inv);
  const size_t nblocks = (size_t)nloop * M * M;
  if (d_mu) k_cond_diag<<<grid_for(nblocks * NB, 256, h->sms * 16), 256, 0, h->st>>>((const double2 *)d_mu, M, nblocks, (double2 *)h->post[5].p);
  k_cond_contract<<<dim3((nv + CD_THREADS - 1) / CD_THREADS, nchunk, nloop), CD_THREADS, 0, h->st>>>(
      (const double2 *)h->post[2].p, (const double2 *)h->post[3].p, TS, (const double2 *)d_D, nv, M, nchunk, (double2 *)h->post[6].p);
  k_cond_finish<<<(NB * nv + 127) / 128, 128, 0, h->st>>>((const double2 *)h->post[6].p, inv, nv, nchunk, nloop, factor, (double2 *)d_int,
                                                         per_type ? (double2 *)d_at : nullptr);
  h->launches += 4;
  CUDA_TRY(cudaGetLastError());
  // random vectors sharded over ranks: the integrand is linear in the moments, so the sum over all vectors of the job is
  // one all-reduce of 18 x nv complex numbers on the device
  if (reduce_ranks) TRY(comm_allreduce_dev(h, d_int, 2 * (size_t)NB * nv));
  TRY(to_host(h, integrand, d_int, 2 * (size_t)NB * nv));
  if (integrand_at) {
    if (per_type) TRY(to_host(h, integrand_at, d_at, 2 * (size_t)NB * nv * nloop));
    else memset(integrand_at, 0, sizeof(cplx) * (size_t)NB * nv * nloop);
  }
  CUDA_TRY(cudaStreamSynchronize(h->st));
  return RSREC_OK;
}

extern "C" {

int rsrec_bpopt(rsrec_handle h, int nchains, int ll, const double *a, const double *rb, double *ainf, double *rbinf, int *ifail) {
  if (!h || nchains < 0 || ll < 2 || !a || !rb || !ainf || !rbinf) return fail(RSREC_EINVAL, "rsrec_bpopt: bad argument (need ll >= 2)");
  if (nchains == 0) return RSREC_OK;
  CUDA_TRY(cudaSetDevice(h->dev));
  const size_t n = (size_t)ll * nchains;
  TRY(to_dev(h, h->post[0], a, n));
  TRY(to_dev(h, h->post[1], rb, n));
  TRY(dev_alloc(h->post[2], 3 * (size_t)nchains, false));
  double *d_ai = h->post[2].p, *d_bi = d_ai + nchains;
  int *d_if = (int *)(d_bi + nchains);
  const ChainLayout lay{(long long)nchains, (long long)ll, 0, 1};
  launch_bpopt(h, h->post[0].p, h->post[1].p, lay, ll, nchains, d_ai, d_bi, d_if, 0);
  h->launches++;
  CUDA_TRY(cudaGetLastError());
  TRY(to_host(h, ainf, d_ai, nchains));
  TRY(to_host(h, rbinf, d_bi, nchains));
  if (ifail) { CUDA_TRY(cudaMemcpyAsync(ifail, d_if, nchains * sizeof(int), cudaMemcpyDeviceToHost, h->st)); h->d2h_bytes += nchains * 4; }
  CUDA_TRY(cudaStreamSynchronize(h->st));
  return RSREC_OK;
}

int rsrec_get_terminf(rsrec_handle h, const cplx *a_b, const cplx *b_b, int na, int ll, double *a_inf, double *b_inf,
                      double *a_inf0, double *b_inf0) {
  if (!h || na < 0 || ll < 2 || !a_b || !b_b || !a_inf || !b_inf) return fail(RSREC_EINVAL, "rsrec_get_terminf: bad argument (need ll >= 2)");
  if (na == 0) return RSREC_OK;
  CUDA_TRY(cudaSetDevice(h->dev));
  const size_t n = (size_t)na * ll * BLKD;
  TRY(to_dev(h, h->post[0], a_b, n));
  TRY(to_dev(h, h->post[1], b_b, n));
  TRY(dev_alloc(h->post[2], 2 * (size_t)na * (BLKC + 1), false));
  double *d_ai = h->post[2].p, *d_bi = d_ai + (size_t)na * BLKC, *d_a0 = d_bi + (size_t)na * BLKC, *d_b0 = d_a0 + na;
  TRY(d_terminf(h, h->post[0].p, h->post[1].p, na, ll, d_ai, d_bi, d_a0, d_b0));
  TRY(to_host(h, a_inf, d_ai, (size_t)na * BLKC));
  TRY(to_host(h, b_inf, d_bi, (size_t)na * BLKC));
  if (a_inf0) TRY(to_host(h, a_inf0, d_a0, na));
  if (b_inf0) TRY(to_host(h, b_inf0, d_b0, na));
  CUDA_TRY(cudaStreamSynchronize(h->st));
  return RSREC_OK;
}

int rsrec_bgreen(rsrec_handle h, const cplx *a_b, const cplx *b_b, int ll, const double *e, int nv, int ie_start, int ie_len,
                 const double *a_inf, const double *b_inf, double eta_re, double eta_im, int sym_term, cplx *g_out) {
  if (!h || ll < 1 || nv < 0 || !a_b || !b_b || !e || !a_inf || !b_inf || !g_out || ie_start < 1 || ie_len < 0 || ie_start - 1 + ie_len > nv)
    return fail(RSREC_EINVAL, "rsrec_bgreen: bad argument");
  if (nv == 0) return RSREC_OK;
  CUDA_TRY(cudaSetDevice(h->dev));
  TRY(to_dev(h, h->post[0], a_b, (size_t)ll * BLKD));
  TRY(to_dev(h, h->post[1], b_b, (size_t)ll * BLKD));
  TRY(to_dev(h, h->post[2], a_inf, BLKC));
  TRY(to_dev(h, h->post[3], b_inf, BLKC));
  TRY(to_dev(h, h->post[4], e, nv));
  TRY(dev_alloc(h->post[5], (size_t)nv * BLKD, false));
  TRY(d_bgreen(h, h->post[0].p, h->post[1].p, ll, 1, h->post[4].p, nv, ie_start - 1, ie_len, h->post[2].p, h->post[3].p, eta_re, eta_im, sym_term, h->post[5].p));
  TRY(to_host(h, g_out, h->post[5].p, (size_t)nv * BLKD));
  CUDA_TRY(cudaStreamSynchronize(h->st));
  return RSREC_OK;
}

int rsrec_block_green(rsrec_handle h, const cplx *a_b, const cplx *b_b, int na, int ll, const double *e, int nv, int sym_term, cplx *g0) {
  if (!h || na < 0 || ll < 2 || nv < 0 || !a_b || !b_b || !e || !g0) return fail(RSREC_EINVAL, "rsrec_block_green: bad argument (need ll >= 2)");
  if (na == 0 || nv == 0) return RSREC_OK;
  CUDA_TRY(cudaSetDevice(h->dev));
  const size_t n = (size_t)na * ll * BLKD;
  TRY(to_dev(h, h->post[0], a_b, n));
  TRY(to_dev(h, h->post[1], b_b, n));
  TRY(dev_alloc(h->post[2], 2 * (size_t)na * (BLKC + 1), false));
  TRY(to_dev(h, h->post[4], e, nv));
  TRY(dev_alloc(h->post[5], (size_t)na * nv * BLKD, false));
  double *d_ai = h->post[2].p, *d_bi = d_ai + (size_t)na * BLKC, *d_a0 = d_bi + (size_t)na * BLKC, *d_b0 = d_a0 + na;
  TRY(d_terminf(h, h->post[0].p, h->post[1].p, na, ll, d_ai, d_bi, d_a0, d_b0, true));
  TRY(d_bgreen(h, h->post[0].p, h->post[1].p, ll, na, h->post[4].p, nv, 0, nv, d_ai, d_bi, 0.0, 0.0, sym_term, h->post[5].p));
  TRY(g0_keep(h, h->post[5].p, na, nv));
  TRY(to_host(h, g0, h->post[5].p, (size_t)na * nv * BLKD));
  CUDA_TRY(cudaStreamSynchronize(h->st));
  return RSREC_OK;
}

int rsrec_chebyshev_green(rsrec_handle h, const cplx *mu_n, int na, int lld, const double *ene, int nv, double energy_min,
                          double energy_max, cplx *mu_ng, cplx *g0) {
  if (!h || na < 0 || lld < 0 || nv < 0 || !mu_n || !ene || !g0 || energy_max == energy_min) return fail(RSREC_EINVAL, "rsrec_chebyshev_green: bad argument");
  if (na == 0) return RSREC_OK;
  CUDA_TRY(cudaSetDevice(h->dev));
  const size_t n = (size_t)na * (2 * lld + 2) * BLKD;
  TRY(to_dev(h, h->post[0], mu_n, n));
  TRY(dev_alloc(h->post[1], n, false));
  TRY(to_dev(h, h->post[4], ene, nv));
  TRY(dev_alloc(h->post[5], std::max<size_t>(1, (size_t)na * nv * BLKD), false));
  if (nv == 0) {  // only the weighted moments
    std::vector<double> kern;
    host_jackson_kernel(2 * lld + 2, kern);
    TRY(to_dev(h, h->post[10], kern.data(), kern.size()));
    k_cheb_weight<<<grid_for(n / 2, 256, h->sms * 8), 256, 0, h->st>>>((const double2 *)h->post[0].p, h->post[10].p, 2 * lld + 2, n / 2, (double2 *)h->post[1].p);
    h->launches++;
    CUDA_TRY(cudaStreamSynchronize(h->st));
  } else {
    TRY(d_cheb_green(h, h->post[0].p, na, lld, h->post[4].p, nv, energy_min, energy_max, h->post[1].p, h->post[5].p));
  }
  if (mu_ng) TRY(to_host(h, mu_ng, h->post[1].p, n));
  if (nv > 0) TRY(g0_keep(h, h->post[5].p, na, nv));
  TRY(to_host(h, g0, h->post[5].p, (size_t)na * nv * BLKD));
  CUDA_TRY(cudaStreamSynchronize(h->st));
  return RSREC_OK;
}

int rsrec_density(rsrec_handle h, const double *a, const double *b2, int lld, int na, int nmdir, const double *ene, int nv,
                  const double *dw_l, const double *cshi, double *tdens) {
  if (!h || lld < 2 || na < 0 || nmdir < 1 || nv < 0 || !a || !b2 || !ene || !dw_l || !cshi || !tdens) return fail(RSREC_EINVAL, "rsrec_density: bad argument (need lld >= 2)");
  if (na == 0 || nv == 0) return RSREC_OK;
  CUDA_TRY(cudaSetDevice(h->dev));
  const size_t n = (size_t)lld * NB * na * nmdir;
  TRY(to_dev(h, h->post[0], a, n));
  TRY(to_dev(h, h->post[1], b2, n));
  TRY(to_dev(h, h->post[2], dw_l, (size_t)NB * na));
  TRY(to_dev(h, h->post[3], cshi, (size_t)NB * na));
  TRY(to_dev(h, h->post[4], ene, nv));
  TRY(dev_alloc(h->post[5], (size_t)NB * nv * na * nmdir, false));
  TRY(d_density(h, h->post[0].p, h->post[1].p, lld, na, nmdir, h->post[4].p, nv, h->post[2].p, h->post[3].p, h->post[5].p));
  TRY(to_host(h, tdens, h->post[5].p, (size_t)NB * nv * na * nmdir));
  CUDA_TRY(cudaStreamSynchronize(h->st));
  return RSREC_OK;
}

int rsrec_sgreen(rsrec_handle h, const double *a, const double *b2, int lld, int na, int nmdir, const double *ene, int nv,
                 const double *dw_l, const double *cshi, cplx *g0) {
  if (!h || lld < 2 || na < 0 || (nmdir != 1 && nmdir != 3) || nv < 0 || !a || !b2 || !ene || !dw_l || !cshi || !g0)
    return fail(RSREC_EINVAL, "rsrec_sgreen: bad argument (need lld >= 2, nmdir 1 or 3)");
  if (na == 0 || nv == 0) return RSREC_OK;
  CUDA_TRY(cudaSetDevice(h->dev));
  const size_t n = (size_t)lld * NB * na * nmdir;  // the first nmdir direction slabs of recursion%a(lld,18,na,3)
  TRY(to_dev(h, h->post[0], a, n));
  TRY(to_dev(h, h->post[1], b2, n));
  TRY(to_dev(h, h->post[2], dw_l, (size_t)NB * na));
  TRY(to_dev(h, h->post[3], cshi, (size_t)NB * na));
  TRY(to_dev(h, h->post[4], ene, nv));
  TRY(dev_alloc(h->post[5], (size_t)NB * nv * na * nmdir, false));
  TRY(dev_alloc(h->post[6], (size_t)na * nv * BLKD, false));
  TRY(d_density(h, h->post[0].p, h->post[1].p, lld, na, nmdir, h->post[4].p, nv, h->post[2].p, h->post[3].p, h->post[5].p));
  CUDA_TRY(cudaMemsetAsync(h->post[6].p, 0, (size_t)na * nv * BLKD * sizeof(double), h->st));
  const size_t total = (size_t)(nmdir == 1 ? NB : 9) * nv * na;
  k_sgreen_assemble<<<grid_for(total, 256, h->sms * 8), 256, 0, h->st>>>(h->post[5].p, nv, na, nmdir, (double2 *)h->post[6].p);
  h->launches++;
  CUDA_TRY(cudaGetLastError());
  TRY(g0_keep(h, h->post[6].p, na, nv));
  TRY(to_host(h, g0, h->post[6].p, (size_t)na * nv * BLKD));
  CUDA_TRY(cudaStreamSynchronize(h->st));
  return RSREC_OK;
}

int rsrec_conductivity_integrand(rsrec_handle h, const cplx *mu_nm, int M, int nloop, const double *ene, int nv, double energy_min,
                                 double energy_max, int per_type, cplx *integrand, cplx *integrand_at) {
  if (!h || M < 1 || nloop < 1 || nv < 1 || !mu_nm || !ene || !integrand || energy_max == energy_min)
    return fail(RSREC_EINVAL, "rsrec_conductivity_integrand: bad argument");
  CUDA_TRY(cudaSetDevice(h->dev));
  TRY(to_dev(h, h->post[9], mu_nm, (size_t)nloop * M * M * BLKD));
  return d_cond_integrand(h, h->post[9].p, nullptr, M, nloop, ene, nv, energy_min, energy_max, per_type, integrand, integrand_at);
}

// ---- fused entry points: recursion + its consumer without a host round trip of the coefficients -----------------
// run_recursion + run_dos of the block path (self.f90:799-856): recur_b -> zsqr -> get_terminf -> bgreen.
// a_b, b2_b (18,18,lld,nunits; b2_b = B^2 as recur_b leaves it; either may be NULL), g0 (18,18,nv,nunits).
static int recur_b_green_impl(rsrec_handle h, int nunits, const int32_t *site_i, const int32_t *site_j, const cplx *asign,
                             const cplx *bsign, int lld, const double *ene, int nv, int sym_term, cplx *a_b, cplx *b2_b, cplx *g0) {
  if (!h || nunits < 0 || lld < 2 || nv < 1 || !ene || (nunits > 0 && !site_i)) return fail(RSREC_EINVAL, "rsrec_recur_b_green: bad argument (need lld >= 2)");
  if (nunits == 0) return RSREC_OK;
  CUDA_TRY(cudaSetDevice(h->dev));
  TRY(ensure_ready(h));
  const bool keep = g0_fits(h, nunits, nv);
  if (!keep && !g0) return fail(RSREC_ENOMEM, "rsrec_recur_b_green: g0 of all units does not fit on the device; pass a host g0");
  if (keep) TRY(g0_reserve(h, nunits, nv)); else { h->g0_units = 0; h->g0_nv = 0; }
  TRY(to_dev(h, h->post[4], ene, nv));
  const int ub = unit_batch(h, nunits, lanczos_nvec(h, false), (size_t)4 * lld * BLKD * sizeof(double) + (keep ? 0 : (size_t)nv * BLKD * sizeof(double)));
  const size_t hs = (size_t)lld * BLKD;
  for (int u0 = 0; u0 < nunits; u0 += ub) {
    const int n = std::min(ub, nunits - u0);
    TRY(upload_units(h, n, site_i + u0, site_j ? site_j + u0 : nullptr, asign ? asign + u0 : nullptr, bsign ? bsign + u0 : nullptr));
    TRY(plan_build(h, n, site_i + u0, site_j ? site_j + u0 : nullptr));
    TRY(lanczos_batch(h, n, lld, false, a_b ? (double *)(a_b + (size_t)u0 * lld * BLKC) : nullptr,
                      b2_b ? (double *)(b2_b + (size_t)u0 * lld * BLKC) : nullptr));
    // zsqr is free here: crecal_b already formed B = (B^2)^1/2 at every level (k_lz_eig keeps it in bhist); then the
    // terminator and the continued fraction
    TRY(dev_alloc(h->post[1], (size_t)n * hs, false));
    CUDA_TRY(cudaMemcpyAsync(h->post[1].p, h->bhist.p, (size_t)n * hs * sizeof(double), cudaMemcpyDeviceToDevice, h->st));
    TRY(dev_alloc(h->post[2], 2 * (size_t)n * (BLKC + 1), false));
    if (!keep) TRY(dev_alloc(h->post[5], (size_t)n * nv * BLKD, false));
    double *d_g0 = keep ? h->g0all.p + (size_t)u0 * nv * BLKD : h->post[5].p;
    double *d_ai = h->post[2].p, *d_bi = d_ai + (size_t)n * BLKC, *d_a0 = d_bi + (size_t)n * BLKC, *d_b0 = d_a0 + n;
    TRY(d_terminf(h, h->ahist.p, h->post[1].p, n, lld, d_ai, d_bi, d_a0, d_b0, true));
    TRY(d_bgreen(h, h->ahist.p, h->post[1].p, lld, n, h->post[4].p, nv, 0, nv, d_ai, d_bi, 0.0, 0.0, sym_term, d_g0));
    if (g0) TRY(to_host(h, g0 + (size_t)u0 * nv * BLKC, d_g0, (size_t)n * nv * BLKD));
    CUDA_TRY(cudaStreamSynchronize(h->st));
  }
  return RSREC_OK;
}

int rsrec_recur_b_green(rsrec_handle h, int nunits, const int32_t *site_i, int lld, const double *ene, int nv, int sym_term,
                        cplx *a_b, cplx *b2_b, cplx *g0) {
  return recur_b_green_impl(h, nunits, site_i, nullptr, nullptr, nullptr, lld, ene, nv, sym_term, a_b, b2_b, g0);
}
// the same for pair start vectors (recur_b_ij, recursion.f90:1655-1737, + block_green_ij, green.f90:356-384)
int rsrec_recur_b_ij_green(rsrec_handle h, int nunits, const int32_t *site_i, const int32_t *site_j, const cplx *asign,
                           const cplx *bsign, int lld, const double *ene, int nv, int sym_term, cplx *a_b, cplx *b2_b, cplx *g0) {
  return recur_b_green_impl(h, nunits, site_i, site_j, asign, bsign, lld, ene, nv, sym_term, a_b, b2_b, g0);
}

static int cheb_recur_green_impl(rsrec_handle h, int nunits, const int32_t *site_i, const int32_t *site_j, const cplx *asign,
                                 const cplx *bsign, int lld, double energy_min, double energy_max, const double *ene, int nv,
                                 cplx *mu_n, cplx *mu_ng, cplx *g0);
// chebyshev_recur + chebyshev_green (recursion.f90:3057-3130 + green.f90:1030-1108): mu_n, mu_ng may be NULL.
int rsrec_cheb_recur_green(rsrec_handle h, int nunits, const int32_t *site_i, int lld, double energy_min, double energy_max,
                           const double *ene, int nv, cplx *mu_n, cplx *mu_ng, cplx *g0) {
  return cheb_recur_green_impl(h, nunits, site_i, nullptr, nullptr, nullptr, lld, energy_min, energy_max, ene, nv, mu_n, mu_ng, g0);
}
// chebyshev_recur_ij (recursion.f90:2376-2487) + chebyshev_green_ij (green.f90:892-952)
int rsrec_cheb_recur_ij_green(rsrec_handle h, int nunits, const int32_t *site_i, const int32_t *site_j, const cplx *asign,
                              const cplx *bsign, int lld, double energy_min, double energy_max, const double *ene, int nv,
                              cplx *mu_n, cplx *mu_ng, cplx *g0) {
  return cheb_recur_green_impl(h, nunits, site_i, site_j, asign, bsign, lld, energy_min, energy_max, ene, nv, mu_n, mu_ng, g0);
}
static int cheb_recur_green_impl(rsrec_handle h, int nunits, const int32_t *site_i, const int32_t *site_j, const cplx *asign,
                                 const cplx *bsign, int lld, double energy_min, double energy_max, const double *ene, int nv,
                                 cplx *mu_n, cplx *mu_ng, cplx *g0) {
  if (!h || nunits < 0 || lld < 0 || nv < 1 || !ene || energy_max == energy_min || (nunits > 0 && !site_i))
    return fail(RSREC_EINVAL, "rsrec_cheb_recur_green: bad argument");
  if (nunits == 0) return RSREC_OK;
  CUDA_TRY(cudaSetDevice(h->dev));
  TRY(ensure_ready(h));
  const bool keep = g0_fits(h, nunits, nv);
  if (!keep && !g0) return fail(RSREC_ENOMEM, "rsrec_cheb_recur_green: g0 of all units does not fit on the device; pass a host g0");
  if (keep) TRY(g0_reserve(h, nunits, nv)); else { h->g0_units = 0; h->g0_nv = 0; }
  TRY(to_dev(h, h->post[4], ene, nv));
  const double a = (energy_max - energy_min) / (2 - 0.3), b = (energy_max + energy_min) / 2;  // recursion.f90:3078-3079
  const int ub = unit_batch(h, nunits, cheb_nvec(h), (size_t)2 * (2 * lld + 2) * BLKD * sizeof(double) + (keep ? 0 : (size_t)nv * BLKD * sizeof(double)));
  const size_t ms = (size_t)(2 * lld + 2) * BLKD;
  std::vector<cplx> mu_tmp;
  int rc_all = RSREC_OK;
  for (int u0 = 0; u0 < nunits; u0 += ub) {
    const int n = std::min(ub, nunits - u0);
    TRY(rsrec_cheb_begin_sites(h, n, site_i + u0, site_j ? site_j + u0 : nullptr, asign ? asign + u0 : nullptr,
                               bsign ? bsign + u0 : nullptr, lld, a, b));
    TRY(cheb_steps(h, lld));
    cplx *mu_out = mu_n ? mu_n + (size_t)u0 * (2 * lld + 2) * BLKC : nullptr;
    if (!mu_out) { mu_tmp.resize((size_t)n * (2 * lld + 2) * BLKC); mu_out = mu_tmp.data(); }  // the divergence guard reads them
    int rc = cheb_finish(h, mu_out);
    if (rc == RSREC_EDIVERGED) rc_all = rc; else TRY(rc);
    TRY(dev_alloc(h->post[1], (size_t)n * ms, false));
    if (!keep) TRY(dev_alloc(h->post[5], (size_t)n * nv * BLKD, false));
    double *d_g0 = keep ? h->g0all.p + (size_t)u0 * nv * BLKD : h->post[5].p;
    TRY(d_cheb_green(h, h->mu.p, n, lld, h->post[4].p, nv, energy_min, energy_max, h->post[1].p, d_g0));
    if (mu_ng) TRY(to_host(h, mu_ng + (size_t)u0 * (2 * lld + 2) * BLKC, h->post[1].p, (size_t)n * ms));
    if (g0) TRY(to_host(h, g0 + (size_t)u0 * nv * BLKC, d_g0, (size_t)n * nv * BLKD));
    CUDA_TRY(cudaStreamSynchronize(h->st));
  }
  return rc_all;
}

// compute_moments_stochastic + calculate_gamma_nm + the integrand of calculate_conductivity_tensor: only the 18
// diagonals of every mu_nm block are kept (on the device); mu_nm (18,18,M,M,nstart) is downloaded only if non-NULL.
int rsrec_kubo_conductivity(rsrec_handle h, int nstart, int start_kind, const int32_t *start_sites, const double *phases, int M,
                            double energy_min, double energy_max, const double *ene, int nv, cplx *mu_nm, cplx *integrand,
                            cplx *integrand_at) {
  if (!h || nstart < 0 || (nstart == 0 && !(start_kind == 1 && h->comm)) || M < 1 || nv < 1 || !ene || !integrand || energy_max == energy_min)
    return fail(RSREC_EINVAL, "rsrec_kubo_conductivity: bad argument");
  CUDA_TRY(cudaSetDevice(h->dev));
  if (nstart == 0) {  // a rank without random vectors of its own still takes part in the exchange
    const size_t n = 2 * (size_t)NB * nv;
    TRY(dev_alloc(h->post[7], n, false));
    CUDA_TRY(cudaMemsetAsync(h->post[7].p, 0, n * sizeof(double), h->st));
    TRY(comm_allreduce_dev(h, h->post[7].p, n));
    TRY(to_host(h, integrand, h->post[7].p, n));
    CUDA_TRY(cudaStreamSynchronize(h->st));
    return RSREC_OK;
  }
  const double a = (energy_max - energy_min) / (2 - 0.3), b = (energy_max + energy_min) / 2;
  TRY(dev_alloc(h->post[11], 2 * (size_t)nstart * M * M * NB, false));
  {
    HostScope hs_(h, HP_RECUR);
    TRY(kubo_moments_impl(h, nstart, start_kind, start_sites, phases, M, a, b, mu_nm, h->post[11].p));
  }
  HostScope hs_(h, HP_EXCHANGE);  // Gamma contraction + all-reduce of the integrand + download
  return d_cond_integrand(h, nullptr, h->post[11].p, M, nstart, ene, nv, energy_min, energy_max, start_kind == 0, integrand, integrand_at,
                          start_kind == 1);
}

}  // extern "C"

extern "C" {

// create_ll_map (recursion.f90:3277-3303) for the start mask of chebyshev_recur (izeroll(site,1) = 1, 3086-3087):
// izeroll (0:kk, lld+1) int32 column-major.
int rsrec_create_ll_map(rsrec_handle h, int site, int lld, int32_t *izeroll) {
  if (!h || !izeroll || lld < 0 || site < 1 || site > h->kk) return fail(RSREC_EINVAL, "rsrec_create_ll_map: bad argument");
  CUDA_TRY(cudaSetDevice(h->dev));
  TRY(ensure_ready(h));
  const size_t ld = (size_t)h->kk + 1, n = ld * (lld + 1);
  TRY(dev_alloc(h->post[0], (n + 1) / 2, false));
  int32_t *m = (int32_t *)h->post[0].p;
  CUDA_TRY(cudaMemsetAsync(m, 0, n * sizeof(int32_t), h->st));
  const int32_t one = 1;
  CUDA_TRY(cudaMemcpyAsync(m + site, &one, sizeof(int32_t), cudaMemcpyHostToDevice, h->st));
  for (int ll = 0; ll < lld; ll++) {
    k_ll_map_step<<<grid_for(h->kk, 256, h->sms * 8), 256, 0, h->st>>>(h->d_nbr, h->ncols, h->kk, m + ld * ll, m + ld * (ll + 1));
    h->launches++;
  }
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpyAsync(izeroll, m, n * sizeof(int32_t), cudaMemcpyDeviceToHost, h->st));
  h->d2h_bytes += (long long)(n * sizeof(int32_t));
  CUDA_TRY(cudaStreamSynchronize(h->st));
  return RSREC_OK;
}

// chebyshev_orbital_mod, moment part (recursion.f90:2901-3008): mu_n_orb (18,18,lld) = sum over the given start sites
// of  L_r^H T_{n-1}(H~)|r>,  |L_r> = i (Y H~ X - X H~ Y)|r>  (the reference loops over all kk sites and divides by kk;
// the Jackson weighting and the trace integration stay with the caller).  cr (3,kk) = lattice%cr, alat = lattice%alat.
int rsrec_orbital_moments(rsrec_handle h, int nstart, const int32_t *start_sites, const double *cr, double alat, int lld,
                          double a, double b, cplx *mu_n_orb) {
  if (!h || nstart < 0 || lld < 1 || !cr || !mu_n_orb || a == 0.0 || (nstart > 0 && !start_sites)) return fail(RSREC_EINVAL, "rsrec_orbital_moments: bad argument");
  CUDA_TRY(cudaSetDevice(h->dev));
  TRY(ensure_ready(h));
  h->plan.on = false;  // this%izero(:) = 1: every site active
  const int nctas = nctas_for(h, 1), kk = h->kk;
  double *psiref, *tmp, *l1, *l2, *w0, *left, *v0, *v1;
  TRY(get_vec(h, 0, 1, &psiref)); TRY(get_vec(h, 1, 1, &tmp)); TRY(get_vec(h, 2, 1, &l1)); TRY(get_vec(h, 3, 1, &l2));
  TRY(get_vec(h, 4, 1, &w0)); TRY(get_vec(h, 5, 1, &left)); TRY(get_vec(h, 6, 1, &v0)); TRY(get_vec(h, 7, 1, &v1));
  double *vs[] = {psiref, tmp, l1, l2, w0, left, v0, v1};
  for (double *v : vs) TRY(zero_vec(h, v, 1));  // also the null-site block of every vector
  TRY(to_dev(h, h->post[0], cr, (size_t)3 * kk));
  TRY(dev_alloc(h->part, part_doubles(h, 1, nctas), false));
  TRY(dev_alloc(h->mu, (size_t)(lld + 1) * BLKD, false));
  CUDA_TRY(cudaMemsetAsync(h->mu.p, 0, (size_t)(lld + 1) * BLKD * sizeof(double), h->st));
  double *acc = h->mu.p, *one = h->mu.p + (size_t)lld * BLKD;
  const int g = grid_for((size_t)kk * BLKC, 256, h->sms * 8);
  for (int s = 0; s < nstart; s++) {
    TRY(zero_vec(h, psiref, 1));
    TRY(upload_units(h, 1, start_sites + s, nullptr, nullptr, nullptr));
    k_init_site_start<<<1, 32, 0, h->st>>>(psiref, vstride(h), h->d_si, h->d_sj, h->d_as, h->d_bs, 1);
    // Y H~ X |r>
    k_scale_by_pos<<<g, 256, 0, h->st>>>(psiref, l1, h->post[0].p, 0, alat, kk);
    h->launches += 2;
    TRY(apply_op(h, OP_HAM_NOHOH, l1, w0, nullptr, tmp, EPI_HAM, a, b, 1, nctas, nullptr));
    k_scale_by_pos<<<g, 256, 0, h->st>>>(w0, l1, h->post[0].p, 1, alat, kk);
    // X H~ Y |r>
    k_scale_by_pos<<<g, 256, 0, h->st>>>(psiref, l2, h->post[0].p, 1, alat, kk);
    h->launches += 2;
    TRY(apply_op(h, OP_HAM_NOHOH, l2, w0, nullptr, tmp, EPI_HAM, a, b, 1, nctas, nullptr));
    k_scale_by_pos<<<g, 256, 0, h->st>>>(w0, l2, h->post[0].p, 0, alat, kk);
    k_i_times_diff<<<g, 256, 0, h->st>>>(l1, l2, left, kk);
    h->launches += 2;
    double *c0 = v0, *c1 = v1, *cur = psiref;
    for (int n = 0; n < lld; n++) {
      if (n == 1) {
        TRY(apply_op(h, OP_HAM, psiref, c1, nullptr, tmp, EPI_HAM, a, b, 1, nctas, nullptr));
        CUDA_TRY(cudaMemcpyAsync(c0, psiref, vstride(h) * sizeof(double), cudaMemcpyDeviceToDevice, h->st));
        cur = c1;
      } else if (n > 1) {  // c0 <- 2 H~ c1 - c0, then swap
        TRY(apply_op(h, OP_HAM, c1, c0, c0, tmp, EPI_CHEB_NOGRAM, a, b, 1, nctas, nullptr));
        std::swap(c0, c1);
        cur = c1;
      }
      TRY(launch_gram(h, left, cur, 1, nctas, h->part.p));
      TRY(launch_reduce(h, 1, nctas, 0, one, nullptr, BLKD, nullptr, nullptr));
      k_add_block<<<(BLKD + 255) / 256, 256, 0, h->st>>>(acc + (size_t)n * BLKD, one);
      h->launches++;
    }
    CUDA_TRY(cudaGetLastError());
  }
  TRY(to_host(h, mu_n_orb, acc, (size_t)lld * BLKD));
  CUDA_TRY(cudaStreamSynchronize(h->st));
  return RSREC_OK;
}

// lattice%nncal + lattice%remd on the device (lattice.f90:3035-3123, 2823-2907); handle-free because the handle's
// sizes (ncols) are only known once the table exists.  See kernels_lattice.cuh for the algorithm.
int rsrec_build_nn(int device, int kk, const double *crd, const int32_t *no, int ntot, const int32_t *iu, double ct,
                   const int32_t *pbc, const int32_t *nrep, const double *a, double alat, int ncols, int32_t *nn, int *nm_out) {
  if (kk < 1 || !crd || !no || ntot < 1 || !iu || !(ct > 0.0) || !nm_out) return fail(RSREC_EINVAL, "rsrec_build_nn: bad argument");
  for (int t = 0; t < ntot; t++) if (iu[t] < 1 || iu[t] > kk) return fail(RSREC_EINVAL, "rsrec_build_nn: representative site out of range");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return fail(RSREC_ECUDA, "no CUDA device available (this library has no CPU path)"); }
  if (device < 0 || device >= ndev) return fail(RSREC_EINVAL, "device ordinal out of range");
  CUDA_TRY(cudaSetDevice(device));
  LatPbc P;
  memset(&P, 0, sizeof(P));
  P.use = pbc && (pbc[0] || pbc[1] || pbc[2]);
  P.alat = alat;
  for (int l = 0; l < 3; l++) { P.b[l] = P.use ? (pbc[l] != 0) : 0; P.n[l] = nrep ? nrep[l] : 1; }
  if (P.use && !a) return fail(RSREC_EINVAL, "rsrec_build_nn: lattice vectors are required with periodic boundaries");
  for (int e = 0; e < 9; e++) P.a[e] = a ? a[e] : 0.0;
  // bounding box of the sites (host, O(kk)); ghosts and the grid live in the box padded by ct
  double lo[3] = {crd[0], crd[1], crd[2]}, hi[3] = {crd[0], crd[1], crd[2]};
  for (int i = 1; i < kk; i++)
    for (int l = 0; l < 3; l++) { lo[l] = std::min(lo[l], crd[l + 3 * (size_t)i]); hi[l] = std::max(hi[l], crd[l + 3 * (size_t)i]); }
  LatGrid G;
  double cs = ct * (1.0 + 1e-9), vol = 1.0;
  for (int l = 0; l < 3; l++) vol *= (hi[l] - lo[l] + 2.0 * ct);
  const double max_cells = 3.2e7;
  if (vol / (cs * cs * cs) > max_cells) cs = std::cbrt(vol / max_cells);
  G.inv_cs = 1.0 / cs;
  size_t ncells = 1;
  for (int l = 0; l < 3; l++) { G.org[l] = lo[l] - ct; G.dim[l] = (int)std::floor((hi[l] - lo[l] + 2.0 * ct) / cs) + 1; ncells *= (size_t)G.dim[l]; }
  struct Bufs {
    double *crd = nullptr, *gpos = nullptr, *set = nullptr;
    int32_t *no = nullptr, *iu = nullptr, *gidx = nullptr, *cell_pts = nullptr, *rows = nullptr, *nn = nullptr;
    int *ctr = nullptr, *cell_cnt = nullptr, *cell_start = nullptr, *cnt = nullptr;
    ~Bufs() { void *p[] = {crd, gpos, set, no, iu, gidx, cell_pts, rows, nn, ctr, cell_cnt, cell_start, cnt}; for (void *q : p) if (q) cudaFree(q); }
  } B;
  cudaStream_t st = nullptr;  // default stream: this is a set-up call
  CUDA_TRY(cudaMalloc(&B.crd, sizeof(double) * 3 * (size_t)kk));
  CUDA_TRY(cudaMalloc(&B.no, sizeof(int32_t) * (size_t)kk));
  CUDA_TRY(cudaMalloc(&B.iu, sizeof(int32_t) * (size_t)ntot));
  CUDA_TRY(cudaMalloc(&B.ctr, sizeof(int) * 4));
  CUDA_TRY(cudaMalloc(&B.cnt, sizeof(int) * (size_t)kk));
  CUDA_TRY(cudaMemcpy(B.crd, crd, sizeof(double) * 3 * (size_t)kk, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(B.no, no, sizeof(int32_t) * (size_t)kk, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(B.iu, iu, sizeof(int32_t) * (size_t)ntot, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemset(B.ctr, 0, sizeof(int) * 4));
  const int tb = 128, gb = (kk + tb - 1) / tb;
  int hctr[4] = {0, 0, 0, 0};
  if (P.use) {  // 1. ghost images: count, allocate, write
    k_lat_ghosts<<<gb, tb, 0, st>>>(B.crd, kk, P, G, lo[0] - ct, lo[1] - ct, lo[2] - ct, hi[0] + ct, hi[1] + ct, hi[2] + ct, B.ctr, nullptr, nullptr);
    CUDA_TRY(cudaMemcpy(hctr, B.ctr, sizeof(int), cudaMemcpyDeviceToHost));
  }
  const int ng = hctr[0], npts = kk + ng;
  if (ng > 0) {
    CUDA_TRY(cudaMalloc(&B.gidx, sizeof(int32_t) * (size_t)ng));
    CUDA_TRY(cudaMalloc(&B.gpos, sizeof(double) * 3 * (size_t)ng));
    CUDA_TRY(cudaMemset(B.ctr, 0, sizeof(int)));
    k_lat_ghosts<<<gb, tb, 0, st>>>(B.crd, kk, P, G, lo[0] - ct, lo[1] - ct, lo[2] - ct, hi[0] + ct, hi[1] + ct, hi[2] + ct, B.ctr, B.gidx, B.gpos);
  }
  // 2. counting sort of all points into the cell grid
  CUDA_TRY(cudaMalloc(&B.cell_cnt, sizeof(int) * ncells));
  CUDA_TRY(cudaMalloc(&B.cell_start, sizeof(int) * (ncells + 1)));
  CUDA_TRY(cudaMalloc(&B.cell_pts, sizeof(int32_t) * (size_t)npts));
  CUDA_TRY(cudaMemset(B.cell_cnt, 0, sizeof(int) * ncells));
  const int gp = (npts + tb - 1) / tb;
  k_lat_cell_count<<<gp, tb, 0, st>>>(B.crd, B.gpos, kk, npts, G, B.cell_cnt);
  k_lat_scan<<<1, 1024, 0, st>>>(B.cell_cnt, B.cell_start, (int)ncells);
  CUDA_TRY(cudaMemset(B.cell_cnt, 0, sizeof(int) * ncells));
  k_lat_cell_fill<<<gp, tb, 0, st>>>(B.crd, B.gpos, kk, npts, G, B.cell_start, B.cell_cnt, B.cell_pts);
  // 3. nncal: count (upper bound), allocate the rows, fill (exact)
  k_lat_nncal<<<gb, tb, 0, st>>>(B.crd, B.gpos, B.gidx, kk, P, G, B.cell_start, B.cell_pts, ct, B.cnt, B.ctr + 1, nullptr, 0);
  CUDA_TRY(cudaMemcpy(hctr, B.ctr, sizeof(int) * 4, cudaMemcpyDeviceToHost));
  const int rowcap = std::max(1, hctr[1] - 1);
  CUDA_TRY(cudaMalloc(&B.rows, sizeof(int32_t) * (size_t)kk * rowcap));
  k_lat_nncal<<<gb, tb, 0, st>>>(B.crd, B.gpos, B.gidx, kk, P, G, B.cell_start, B.cell_pts, ct, B.cnt, B.ctr + 2, B.rows, rowcap);
  CUDA_TRY(cudaMemcpy(hctr, B.ctr, sizeof(int) * 4, cudaMemcpyDeviceToHost));
  CUDA_TRY(cudaGetLastError());
  const int nnmax = std::max(hctr[2], 1);  // NM = NNMAX (nn(:,1) counts the site itself)
  *nm_out = nnmax;
  if (!nn || ncols < nnmax + 1) return fail(RSREC_EINVAL, "rsrec_build_nn: nn needs nm+1 = " + std::to_string(nnmax + 1) + " columns (lattice.f90:1856)");
  // 4. remd
  const int nmcols = nnmax + 2;
  CUDA_TRY(cudaMalloc(&B.set, sizeof(double) * 3 * (size_t)ntot * nmcols));
  CUDA_TRY(cudaMemset(B.set, 0, sizeof(double) * 3 * (size_t)ntot * nmcols));
  CUDA_TRY(cudaMalloc(&B.nn, sizeof(int32_t) * (size_t)kk * ncols));
  CUDA_TRY(cudaMemset(B.nn, 0, sizeof(int32_t) * (size_t)kk * ncols));
  k_lat_set<<<ntot, 64, 0, st>>>(B.crd, kk, P, B.iu, ntot, B.cnt, B.rows, nmcols, B.set);
  k_lat_remd<<<gb, tb, 0, st>>>(B.crd, kk, P, B.no, B.iu, ntot, B.cnt, B.rows, nmcols, B.set, ncols, B.nn, B.ctr + 3);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpy(hctr, B.ctr, sizeof(int) * 4, cudaMemcpyDeviceToHost));
  if (hctr[3] == 2) return fail(RSREC_EINVAL, "rsrec_build_nn: VECTOR NOT FOUND (a site has a neighbour vector its type's representative lacks; lattice.f90:2896)");
  if (hctr[3] == 3) return fail(RSREC_EINVAL, "rsrec_build_nn: TYPE NO NOT FOUND (lattice.f90:2860)");
  CUDA_TRY(cudaMemcpy(nn, B.nn, sizeof(int32_t) * (size_t)kk * ncols, cudaMemcpyDeviceToHost));
  return RSREC_OK;
}

// Device-side assembly of the block sets: what build_bulkham / build_locham (+ chbar_nc's orbital
// part, ham0m_nc, hcpx, build_obarm, build_enim) compute on the host, from the structure-constant blocks and the
// potential parameters, straight into the device-resident sets the recursion kernels read.
int rsrec_build_hamiltonian(rsrec_handle h, const double *hhh, const int32_t *jt, const int32_t *it, const cplx *pot,
                            const double *mom, const cplx *lsham, int hoh, cplx *ee, cplx *eeo, cplx *hall, cplx *hallo,
                            cplx *enim, cplx *obarm) {
  if (!h || !hhh || !jt || !it || !pot || !mom || !lsham) return fail(RSREC_EINVAL, "rsrec_build_hamiltonian: null argument");
  CUDA_TRY(cudaSetDevice(h->dev));
  const int ncls = h->ncls, nslot = h->nslot, ntype = h->ntype;
  for (int c = 0; c < ncls; c++) {
    if (it[c] < 1 || it[c] > ntype) return fail(RSREC_EINVAL, "rsrec_build_hamiltonian: it out of range");
    for (int m = 0; m < nslot; m++)
      if (jt[m + nslot * c] < 0 || jt[m + nslot * c] > ntype) return fail(RSREC_EINVAL, "rsrec_build_hamiltonian: jt out of range");
  }
  const size_t nblk = (size_t)ncls * nslot;
  TRY(to_dev(h, h->post[0], hhh, 81 * nblk));
  TRY(dev_alloc(h->post[1], (nblk + ncls + 1) / 2 + 1, false));
  int32_t *d_jt = (int32_t *)h->post[1].p, *d_it = d_jt + nblk;
  CUDA_TRY(cudaMemcpyAsync(d_jt, jt, nblk * sizeof(int32_t), cudaMemcpyHostToDevice, h->st));
  CUDA_TRY(cudaMemcpyAsync(d_it, it, ncls * sizeof(int32_t), cudaMemcpyHostToDevice, h->st));
  h->h2d_bytes += (long long)((nblk + ncls) * sizeof(int32_t));
  TRY(to_dev(h, h->post[2], pot, (size_t)2 * 9 * POT_NPAR * ntype));
  TRY(to_dev(h, h->post[3], mom, (size_t)3 * ntype));
  TRY(to_dev(h, h->cLS, lsham, (size_t)ntype * BLKD));
  TRY(dev_alloc(h->cBLK, nblk * BLKD, false));
  TRY(dev_alloc(h->cBLKO, nblk * BLKD, false));
  TRY(dev_alloc(h->cENIM, (size_t)ntype * BLKD, false));
  TRY(dev_alloc(h->cOBARM, (size_t)ntype * BLKD, false));
  k_ham_blocks<<<dim3(nslot, ncls), 96, 0, h->st>>>(h->post[0].p, d_jt, d_it, (const double2 *)h->post[2].p, h->post[3].p, hoh ? 1 : 0, nslot, (double2 *)h->cBLK.p);
  k_ham_onsite18<<<dim3(ntype, 2), 96, 0, h->st>>>((const double2 *)h->post[2].p, h->post[3].p, (double2 *)h->cOBARM.p, (double2 *)h->cENIM.p);
  h->launches += 2;
  if (hoh) {
    k_ham_times_o<<<dim3(nslot, ncls), BLKC, 0, h->st>>>((const double2 *)h->cBLK.p, (const double2 *)h->cOBARM.p, d_jt, nslot, (double2 *)h->cBLKO.p);
    h->launches++;
  }
  CUDA_TRY(cudaGetLastError());
  const size_t ntb = (size_t)ntype * nslot;
  if (ee) TRY(to_host(h, ee, h->cBLK.p, ntb * BLKD));
  if (hall && h->nmax > 0) TRY(to_host(h, hall, h->cBLK.p + ntb * BLKD, (nblk - ntb) * BLKD));
  if (hoh && eeo) TRY(to_host(h, eeo, h->cBLKO.p, ntb * BLKD));
  if (hoh && hallo && h->nmax > 0) TRY(to_host(h, hallo, h->cBLKO.p + ntb * BLKD, (nblk - ntb) * BLKD));
  if (enim) TRY(to_host(h, enim, h->cENIM.p, (size_t)ntype * BLKD));
  if (obarm) TRY(to_host(h, obarm, h->cOBARM.p, (size_t)ntype * BLKD));
  CUDA_TRY(cudaStreamSynchronize(h->st));
  h->hoh = hoh ? 1 : 0;
  h->have_ham = true; h->dirty_ham = true; h->ham_on_device = true; h->have_glob = false;
  return RSREC_OK;
}

// ---- local spin axis (hamiltonian%rotate_to_local_axis / rotate_from_local_axis, hamiltonian.f90:2442-2484) -----------
// ROTMAT / DSs / car2sph of math.f90:2055-2192, evaluated on the host (an 18x18 matrix of closed-form entries)
static long long factint(int n) { if (n < 0) return 0; long long f = 1; for (int i = 1; i <= n; i++) f *= i; return f; }
static long long binom_ll(int x, int y) { if (y < 0 || y > x) return 0; return factint(x) / (factint(y) * factint(x - y)); }
static int nint_d(double x) { return (int)std::lround(x); }
static double wigner_dss(double J, double M, double Mp, double beta) {
  const int smin = std::max(0, nint_d(-Mp - M)), smax = std::min(nint_d(J - Mp), nint_d(J - M));
  double d = 0.0;
  for (int s = smin; s <= smax; s++) {
    const double ds = s * 1.0;
    const double dst = (double)binom_ll(nint_d(J + M), nint_d(J - Mp - ds)) * (double)binom_ll(nint_d(J - M), s) *
                       std::pow(-1.0, nint_d(J - Mp - ds));
    d += dst * std::pow(std::cos(0.5 * beta), 2 * ds + Mp + M) * std::pow(std::sin(0.5 * beta), 2 * J - 2 * ds - Mp - M);
  }
  return d * std::sqrt(1.0 * factint(nint_d(J + Mp)) * factint(nint_d(J - Mp)) / ((double)factint(nint_d(J - M)) * factint(nint_d(J + M))));
}
static void host_rotmat(const double mom[3], std::vector<std::complex<double>> &R) {
  typedef std::complex<double> cd;
  const double X = mom[0], Y = mom[1], Z = mom[2], D2 = X * X + Y * Y, R2 = X * X + Y * Y + Z * Z;
  const double alfa = D2 == 0.0 ? 0.0 : std::atan2(Y, X), beta = std::acos(Z / R2), gama = 0.0;  // car2sph (sic: z / r^2)
  const cd IM(0.0, 1.0);
  cd SM[2][2];
  SM[0][0] = wigner_dss(0.5, 0.5, 0.5, beta) * std::exp(-IM * (0.5 * alfa + 0.5 * gama));
  SM[0][1] = wigner_dss(0.5, 0.5, -0.5, beta) * std::exp(-IM * (0.5 * alfa - 0.5 * gama));
  SM[1][0] = wigner_dss(0.5, -0.5, 0.5, beta) * std::exp(-IM * (-0.5 * alfa + 0.5 * gama));
  SM[1][1] = wigner_dss(0.5, -0.5, -0.5, beta) * std::exp(-IM * (-0.5 * alfa - 0.5 * gama));
  cd M9[9][9];
  for (auto &row : M9) for (auto &e : row) e = 0.0;
  for (int J = 0; J <= 2; J++) {
    const int S = J * J + 1 + J;
    for (int M = -J; M <= J; M++)
      for (int Mp = -J; Mp <= J; Mp++)
        M9[S + M - 1][S + Mp - 1] = wigner_dss(J * 1.0, M * 1.0, Mp * 1.0, beta) * std::exp(-IM * ((double)M * alfa + (double)Mp * gama));
  }
  R.assign(BLKC, cd(0.0, 0.0));
  for (int M = 0; M < 9; M++)
    for (int Mp = 0; Mp < 9; Mp++) {
      R[Mp + NB * M] = M9[Mp][M] * SM[0][0];
      R[Mp + NB * (M + 9)] = M9[Mp][M] * SM[0][1];
      R[(Mp + 9) + NB * M] = M9[Mp][M] * SM[1][0];
      R[(Mp + 9) + NB * (M + 9)] = M9[Mp][M] * SM[1][1];
    }
}
static int rotate_sets(H *h, const double m_loc[3]) {
  TRY(ensure_ready(h));  // the current sets are staged on the device
  const size_t nb = (size_t)h->ncls * h->nslot;
  if (!h->have_glob) {  // *_glob = the sets as built (hamiltonian.f90:1609-1612, 1661-1664)
    TRY(dev_alloc(h->gBLK, nb * BLKD, false));
    CUDA_TRY(cudaMemcpyAsync(h->gBLK.p, h->cBLK.p, nb * BLKD * sizeof(double), cudaMemcpyDeviceToDevice, h->st));
    if (h->hoh) {
      TRY(dev_alloc(h->gBLKO, nb * BLKD, false));
      TRY(dev_alloc(h->gENIM, (size_t)h->ntype * BLKD, false));
      CUDA_TRY(cudaMemcpyAsync(h->gBLKO.p, h->cBLKO.p, nb * BLKD * sizeof(double), cudaMemcpyDeviceToDevice, h->st));
      CUDA_TRY(cudaMemcpyAsync(h->gENIM.p, h->cENIM.p, (size_t)h->ntype * BLKD * sizeof(double), cudaMemcpyDeviceToDevice, h->st));
    }
    h->have_glob = true;
  }
  std::vector<std::complex<double>> R;
  host_rotmat(m_loc, R);
  TRY(to_dev(h, h->post[10], R.data(), BLKD));
  k_rotmag<<<(unsigned)nb, BLKC, 0, h->st>>>((const double2 *)h->gBLK.p, (const double2 *)h->post[10].p, (double2 *)h->cBLK.p);
  h->launches++;
  if (h->hoh) {
    k_rotmag<<<(unsigned)nb, BLKC, 0, h->st>>>((const double2 *)h->gBLKO.p, (const double2 *)h->post[10].p, (double2 *)h->cBLKO.p);
    k_rotmag<<<(unsigned)h->ntype, BLKC, 0, h->st>>>((const double2 *)h->gENIM.p, (const double2 *)h->post[10].p, (double2 *)h->cENIM.p);
    h->launches += 2;
  }
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaStreamSynchronize(h->st));  // R goes out of scope
  h->ham_on_device = true;  // the rotated sets live only on the device
  h->dirty_ham = true;      // repack (lsham is NOT rotated, like the reference)
  return RSREC_OK;
}

int rsrec_rotate_to_local_axis(rsrec_handle h, const double *m_loc) {
  if (!h || !m_loc) return fail(RSREC_EINVAL, "rsrec_rotate_to_local_axis: bad argument");
  CUDA_TRY(cudaSetDevice(h->dev));
  return rotate_sets(h, m_loc);
}
int rsrec_rotate_from_local_axis(rsrec_handle h) {
  if (!h) return fail(RSREC_EINVAL, "null handle");
  if (!h->have_glob) return RSREC_OK;
  CUDA_TRY(cudaSetDevice(h->dev));
  const size_t nb = (size_t)h->ncls * h->nslot;
  CUDA_TRY(cudaMemcpyAsync(h->cBLK.p, h->gBLK.p, nb * BLKD * sizeof(double), cudaMemcpyDeviceToDevice, h->st));
  if (h->hoh) {
    CUDA_TRY(cudaMemcpyAsync(h->cBLKO.p, h->gBLKO.p, nb * BLKD * sizeof(double), cudaMemcpyDeviceToDevice, h->st));
    CUDA_TRY(cudaMemcpyAsync(h->cENIM.p, h->gENIM.p, (size_t)h->ntype * BLKD * sizeof(double), cudaMemcpyDeviceToDevice, h->st));
  }
  h->dirty_ham = true;
  return RSREC_OK;
}
// recur_b with hamiltonian%local_axis (recursion.f90:1826-1832): unit u runs on the sets rotated to mom(:,u); like the
// reference the sets stay rotated to the last unit's axis on return (rsrec_rotate_from_local_axis restores them).
int rsrec_lanczos_block_local_axis(rsrec_handle h, int nunits, const int32_t *site_i, const double *mom, int lld, cplx *a_b,
                                   cplx *b2_b) {
  if (!h || nunits < 0 || !a_b || !b2_b || lld < 1 || (nunits > 0 && (!site_i || !mom))) return fail(RSREC_EINVAL, "rsrec_lanczos_block_local_axis: bad argument");
  CUDA_TRY(cudaSetDevice(h->dev));
  for (int u = 0; u < nunits; u++) {
    TRY(rotate_sets(h, mom + 3 * (size_t)u));
    TRY(rsrec_lanczos_block(h, 1, site_i + u, nullptr, nullptr, nullptr, lld, a_b + (size_t)u * lld * BLKC, b2_b + (size_t)u * lld * BLKC));
  }
  return RSREC_OK;
}

// Optional: lattice%cr (3,kk).  Coordinates never enter the arithmetic; they only order the 8-site tiles along a
// space-filling curve so that the gathers of concurrently running CTAs hit L2 (results are bit-identical either way).
int rsrec_set_positions(rsrec_handle h, const double *cr) {
  if (!h) return fail(RSREC_EINVAL, "null handle");
  if (cr) h->pos.assign(cr, cr + (size_t)3 * h->kk); else h->pos.clear();
  h->dirty = true;
  return RSREC_OK;
}

int rsrec_synchronize(rsrec_handle h) {
  if (!h) return fail(RSREC_EINVAL, "null handle");
  CUDA_TRY(cudaStreamSynchronize(h->st));
  return RSREC_OK;
}
void *rsrec_stream(rsrec_handle h) { return h ? (void *)h->st : nullptr; }
long long rsrec_h2d_bytes(rsrec_handle h) { return h ? h->h2d_bytes : 0; }
long long rsrec_d2h_bytes(rsrec_handle h) { return h ? h->d2h_bytes : 0; }
int rsrec_profile(rsrec_handle h, int enable) {
  if (!h) return fail(RSREC_EINVAL, "null handle");
  h->profile = enable != 0;
  h->prof_used = 0;
  return RSREC_OK;
}
int rsrec_profile_read(rsrec_handle h, double *total_ms, int *nlaunches) {
  if (!h || !total_ms || !nlaunches) return fail(RSREC_EINVAL, "rsrec_profile_read: bad argument");
  CUDA_TRY(cudaStreamSynchronize(h->st));
  double tot = 0.0;
  for (size_t i = 0; i < h->prof_used; i++) {
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, h->prof_events[i].first, h->prof_events[i].second));
    tot += ms;
  }
  *total_ms = tot;
  *nlaunches = (int)h->prof_used;
  h->prof_used = 0;
  return RSREC_OK;
}

// ---- the exchange step of the unit-sharded path: NCCL on the handle's stream ---------------------
#define NCCL_TRY(x)                                                                                            \
  do {                                                                                                         \
    int r_ = (x);                                                                                              \
    if (r_ != RS_NCCL_SUCCESS) return fail(RSREC_ECUDA, std::string(#x) + ": " + nccl_api()->GetErrorString(r_)); \
  } while (0)

}  // extern "C"
// in-place sum of n doubles on the device over all ranks (no-op without a communicator)
static int comm_allreduce_dev(H *h, double *d, size_t n) {
  if (!h->comm || h->comm_size == 1 || n == 0) return RSREC_OK;
  NCCL_TRY(nccl_api()->AllReduce(d, d, n, RS_NCCL_FLOAT64, RS_NCCL_SUM, h->comm, h->st));
  return RSREC_OK;
}
// get_mpi_variables (mpi.f90:32-58): 0-based half-open unit range of a rank
static void shard_of(int rank, int nranks, long long n, long long *lo, long long *hi) {
  long long per = n / nranks, rem = n % nranks;
  if (rank < rem) { per += 1; *lo = rank * per; } else { *lo = rank * per + rem; }
  *hi = *lo + per;
}
// every rank's contiguous shard (block rule) of `full` [nunits][per_unit doubles] is broadcast from its owner: afterwards
// all ranks hold all units.  `full` is a device buffer; the caller has placed its own shard at its offset.
static int comm_allgather_units_dev(H *h, double *full, size_t per_unit, long long nunits) {
  if (!h->comm || h->comm_size == 1) return RSREC_OK;
  NcclApi *N = nccl_api();
  if (nunits % h->comm_size == 0) {  // equal shards: one in-place all-gather (the shard of rank r already sits at r * count)
    const size_t count = (size_t)(nunits / h->comm_size) * per_unit;
    NCCL_TRY(N->AllGather(full + (size_t)h->comm_rank * count, full, count, RS_NCCL_FLOAT64, h->comm, h->st));
    return RSREC_OK;
  }
  NCCL_TRY(N->GroupStart());
  for (int r = 0; r < h->comm_size; r++) {
    long long lo, hi;
    shard_of(r, h->comm_size, nunits, &lo, &hi);
    if (hi > lo) NCCL_TRY(N->Broadcast(full + (size_t)lo * per_unit, full + (size_t)lo * per_unit, (size_t)(hi - lo) * per_unit, RS_NCCL_FLOAT64, r, h->comm, h->st));
  }
  NCCL_TRY(N->GroupEnd());
  return RSREC_OK;
}
extern "C" {

int rsrec_comm_unique_id(unsigned char *id128) {
  if (!id128) return fail(RSREC_EINVAL, "rsrec_comm_unique_id: null argument");
  NcclApi *N = nccl_api();
  if (!N->dl) return fail(RSREC_ECUDA, N->err);
  rs_ncclUniqueId id;
  NCCL_TRY(N->GetUniqueId(&id));
  memcpy(id128, id.internal, RS_NCCL_ID_BYTES);
  return RSREC_OK;
}

int rsrec_comm_init(rsrec_handle h, int nranks, int rank, const unsigned char *id128) {
  if (!h || nranks < 1 || rank < 0 || rank >= nranks || !id128) return fail(RSREC_EINVAL, "rsrec_comm_init: bad argument");
  NcclApi *N = nccl_api();
  if (!N->dl) return fail(RSREC_ECUDA, N->err);
  CUDA_TRY(cudaSetDevice(h->dev));
  if (h->comm) { N->CommDestroy(h->comm); h->comm = nullptr; }
  rs_ncclUniqueId id;
  memcpy(id.internal, id128, RS_NCCL_ID_BYTES);
  NCCL_TRY(N->CommInitRank(&h->comm, nranks, id, rank));
  h->comm_rank = rank; h->comm_size = nranks;
  return RSREC_OK;
}

int rsrec_comm_destroy(rsrec_handle h) {
  if (!h) return fail(RSREC_EINVAL, "null handle");
  if (h->comm) {
    CUDA_TRY(cudaSetDevice(h->dev));
    CUDA_TRY(cudaStreamSynchronize(h->st));
    nccl_api()->CommDestroy(h->comm);
  }
  h->comm = nullptr; h->comm_rank = 0; h->comm_size = 1;
  return RSREC_OK;
}

int rsrec_comm_info(rsrec_handle h, int *nranks, int *rank, int *nccl_version) {
  if (!h) return fail(RSREC_EINVAL, "null handle");
  if (nranks) *nranks = h->comm ? h->comm_size : 1;
  if (rank) *rank = h->comm ? h->comm_rank : 0;
  if (nccl_version) { *nccl_version = 0; if (h->comm) nccl_api()->GetVersion(nccl_version); }
  return RSREC_OK;
}

int rsrec_shard_range(int rank, int nranks, int nunits, int *first, int *last) {
  if (nranks < 1 || rank < 0 || rank >= nranks || nunits < 0 || !first || !last) return fail(RSREC_EINVAL, "rsrec_shard_range: bad argument");
  long long lo, hi;
  shard_of(rank, nranks, nunits, &lo, &hi);
  *first = (int)lo + 1; *last = (int)hi;  // 1-based inclusive like start_atom / end_atom
  return RSREC_OK;
}

// MPI_ALLREDUCE(MPI_IN_PLACE, buf, count, .., MPI_SUM) of a HOST array (bands.f90:270-275): dtype 0 = real(rp),
// 1 = complex(rp) (count complex numbers), 2 = integer(4).  Staged through the device, reduced over NVLink.
int rsrec_allreduce(rsrec_handle h, void *buf, long long count, int dtype) {
  if (!h || count < 0 || (count > 0 && !buf) || dtype < 0 || dtype > 2) return fail(RSREC_EINVAL, "rsrec_allreduce: bad argument");
  if (!h->comm || h->comm_size == 1 || count == 0) return RSREC_OK;
  CUDA_TRY(cudaSetDevice(h->dev));
  const size_t bytes = (size_t)count * (dtype == 0 ? 8 : dtype == 1 ? 16 : 4);
  TRY(dev_alloc(h->comm_buf, (bytes + 7) / 8, false));
  CUDA_TRY(cudaMemcpyAsync(h->comm_buf.p, buf, bytes, cudaMemcpyHostToDevice, h->st)); h->h2d_bytes += (long long)bytes;
  if (dtype == 2) NCCL_TRY(nccl_api()->AllReduce(h->comm_buf.p, h->comm_buf.p, (size_t)count, RS_NCCL_INT32, RS_NCCL_SUM, h->comm, h->st));
  else TRY(comm_allreduce_dev(h, h->comm_buf.p, bytes / 8));
  CUDA_TRY(cudaMemcpyAsync(buf, h->comm_buf.p, bytes, cudaMemcpyDeviceToHost, h->st)); h->d2h_bytes += (long long)bytes;
  CUDA_TRY(cudaStreamSynchronize(h->st));
  return RSREC_OK;
}

// The MPI_Allgather of per-unit results the reference has commented out (recursion.f90:1788-1799): `local` holds this
// rank's units (block rule of get_mpi_variables over nunits_total) with doubles_per_unit reals each; `full` receives all
// units in global order on every rank.  HOST arrays.
int rsrec_allgather_units(rsrec_handle h, const void *local, void *full, long long doubles_per_unit, int nunits_total) {
  if (!h || !full || doubles_per_unit < 1 || nunits_total < 0) return fail(RSREC_EINVAL, "rsrec_allgather_units: bad argument");
  long long lo, hi;
  shard_of(h->comm ? h->comm_rank : 0, h->comm ? h->comm_size : 1, nunits_total, &lo, &hi);
  if (hi > lo && !local) return fail(RSREC_EINVAL, "rsrec_allgather_units: local shard missing");
  if (!h->comm || h->comm_size == 1) { if (local != full && hi > lo) memcpy(full, local, (size_t)(hi - lo) * doubles_per_unit * 8); return RSREC_OK; }
  CUDA_TRY(cudaSetDevice(h->dev));
  const size_t n = (size_t)nunits_total * doubles_per_unit;
  TRY(dev_alloc(h->comm_buf, n, false));
  if (hi > lo) { CUDA_TRY(cudaMemcpyAsync(h->comm_buf.p + (size_t)lo * doubles_per_unit, local, (size_t)(hi - lo) * doubles_per_unit * 8, cudaMemcpyHostToDevice, h->st)); h->h2d_bytes += (hi - lo) * doubles_per_unit * 8; }
  TRY(comm_allgather_units_dev(h, h->comm_buf.p, (size_t)doubles_per_unit, nunits_total));
  CUDA_TRY(cudaMemcpyAsync(full, h->comm_buf.p, n * 8, cudaMemcpyDeviceToHost, h->st)); h->d2h_bytes += (long long)(n * 8);
  CUDA_TRY(cudaStreamSynchronize(h->st));
  return RSREC_OK;
}

// recur_b for ALL nunits_total recursion sites of the job: this rank runs its block-rule shard, the coefficient histories
// are gathered on the device over NCCL and every rank receives a_b, b2_b (18,18,lld,nunits_total).
int rsrec_lanczos_block_sharded(rsrec_handle h, int nunits_total, const int32_t *site_i, const int32_t *site_j, const cplx *asign,
                                const cplx *bsign, int lld, cplx *a_b, cplx *b2_b) {
  if (!h || nunits_total < 0 || !a_b || !b2_b || lld < 1 || (nunits_total > 0 && !site_i)) return fail(RSREC_EINVAL, "rsrec_lanczos_block_sharded: bad argument");
  if (nunits_total == 0) return RSREC_OK;
  CUDA_TRY(cudaSetDevice(h->dev));
  { HostScope hs_(h, HP_TABLES); TRY(ensure_ready(h)); }
  long long lo, hi;
  shard_of(h->comm ? h->comm_rank : 0, h->comm ? h->comm_size : 1, nunits_total, &lo, &hi);
  const size_t hs = (size_t)lld * BLKD, nall = (size_t)nunits_total * hs;
  TRY(dev_alloc(h->comm_res, 2 * nall, false));   // gathered a_b | b2_b histories
  double *g_a = h->comm_res.p, *g_b = g_a + nall;
  const int nloc = (int)(hi - lo);
  const int ub = nloc > 0 ? unit_batch(h, nloc, lanczos_nvec(h, false), (size_t)3 * lld * BLKD * sizeof(double)) : 1;
  for (int u0 = 0; u0 < nloc; u0 += ub) {
    const int n = std::min(ub, nloc - u0);
    const size_t g0 = (size_t)lo + u0;
    {
      HostScope hs_(h, HP_PLAN);
      TRY(upload_units(h, n, site_i + g0, site_j ? site_j + g0 : nullptr, asign ? asign + g0 : nullptr, bsign ? bsign + g0 : nullptr));
      TRY(plan_build(h, n, site_i + g0, site_j ? site_j + g0 : nullptr));
    }
    HostScope hs_(h, HP_RECUR);
    TRY(lanczos_batch(h, n, lld, false, nullptr, nullptr));
    CUDA_TRY(cudaMemcpyAsync(g_a + g0 * hs, h->ahist.p, (size_t)n * hs * sizeof(double), cudaMemcpyDeviceToDevice, h->st));
    CUDA_TRY(cudaMemcpyAsync(g_b + g0 * hs, h->b2hist.p, (size_t)n * hs * sizeof(double), cudaMemcpyDeviceToDevice, h->st));
  }
  {
    HostScope hs_(h, HP_EXCHANGE);
    TRY(comm_allgather_units_dev(h, g_a, hs, nunits_total));
    TRY(comm_allgather_units_dev(h, g_b, hs, nunits_total));
  }
  HostScope hs_(h, HP_DOWNLOAD);
  TRY(to_host(h, a_b, g_a, nall));
  TRY(to_host(h, b2_b, g_b, nall));
  CUDA_TRY(cudaStreamSynchronize(h->st));
  return RSREC_OK;
}

// Stochastic-trace KPM moments: the local random vectors (phases (kk,nvec_local), this rank's shard) are run, their
// moments summed on the device, the sum all-reduced over the communicator, and mu_sum (18,18,2*lld+2) = the sum over ALL
// vectors of the job comes back on every rank -- one download of 2*lld+2 blocks instead of nvec of them plus a host sum.
int rsrec_cheb_moments_random_sum(rsrec_handle h, int nvec_local, const double *phases, int lld, double a, double b, cplx *mu_sum) {
  if (!h || nvec_local < 0 || !mu_sum || lld < 0 || (nvec_local > 0 && !phases)) return fail(RSREC_EINVAL, "rsrec_cheb_moments_random_sum: bad argument");
  CUDA_TRY(cudaSetDevice(h->dev));
  TRY(ensure_ready(h));
  const size_t ms = (size_t)(2 * lld + 2) * BLKD;
  TRY(dev_alloc(h->comm_res, ms, false));
  CUDA_TRY(cudaMemsetAsync(h->comm_res.p, 0, ms * sizeof(double), h->st));
  const int ub = nvec_local > 0 ? unit_batch(h, nvec_local, cheb_nvec(h), ms * sizeof(double)) : 1;
  for (int u0 = 0; u0 < nvec_local; u0 += ub) {
    const int n = std::min(ub, nvec_local - u0);
    TRY(rsrec_cheb_begin_random(h, n, phases + (size_t)u0 * h->kk, lld, a, b));
    TRY(cheb_steps(h, lld));
    h->cheb.active = false;
    k_sum_units<<<grid_for(ms, 256, h->sms * 8), 256, 0, h->st>>>(h->mu.p, ms, n, h->comm_res.p);  // fixed unit order
    h->launches++;
  }
  TRY(comm_allreduce_dev(h, h->comm_res.p, ms));
  TRY(to_host(h, mu_sum, h->comm_res.p, ms));
  CUDA_TRY(cudaStreamSynchronize(h->st));
  // the reference's divergence guard (recursion.f90:2594) on the summed moments
  const double *m = (const double *)mu_sum;
  const double lim = 1000.0 * std::max(1, nvec_local) * (h->comm ? h->comm_size : 1);
  for (int k = 3; k < 2 * lld + 2; k += 2) {
    double sre = 0.0;
    for (int e = 0; e < BLKC; e++) sre += m[(size_t)k * BLKD + 2 * e];
    if (!(sre <= lim)) return fail(RSREC_EDIVERGED, "Chebyshev moments did not converge. Check energy limits energy_min and energy_max");
  }
  return RSREC_OK;
}

int rsrec_phase_timing(rsrec_handle h, int enable) {
  if (!h) return fail(RSREC_EINVAL, "null handle");
  h->phase_on = enable != 0;
  h->phase_used = 0;
  return RSREC_OK;
}
int rsrec_host_phase_read(rsrec_handle h, double *seconds) {
  if (!h || !seconds) return fail(RSREC_EINVAL, "rsrec_host_phase_read: bad argument");
  for (int k = 0; k < HP_COUNT; k++) { seconds[k] = h->hphase[k]; h->hphase[k] = 0.0; }
  return RSREC_OK;
}
int rsrec_phase_count(void) { return PH_COUNT; }
const char *rsrec_phase_label(int idx) { return (idx >= 0 && idx < PH_COUNT) ? kPhaseLabel[idx] : nullptr; }
int rsrec_phase_read(rsrec_handle h, double *ms, long long *calls) {
  if (!h || !ms || !calls) return fail(RSREC_EINVAL, "rsrec_phase_read: bad argument");
  CUDA_TRY(cudaStreamSynchronize(h->st));
  for (int p = 0; p < PH_COUNT; p++) { ms[p] = 0.0; calls[p] = 0; }
  for (size_t i = 0; i < h->phase_used; i++) {
    float t = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&t, h->phase_events[i].a, h->phase_events[i].b));
    ms[h->phase_events[i].phase] += t;
    calls[h->phase_events[i].phase]++;
  }
  h->phase_used = 0;
  return RSREC_OK;
}
long long rsrec_launch_count(rsrec_handle h) { return h ? h->launches : 0; }
long long rsrec_spin_diag_launch_count(rsrec_handle h) { return h ? h->sd_launches : 0; }

}  // extern "C"

// tail of calculate_conductivity_tensor (conductivity.f90:300-372): the T = 0 Fermi-weighted Simpson integrals of the
// integrand up to every mesh energy.  integrand (18,nv), integrand_at (18,nv,nat) (nat = 0: none);
// sigma (2,19,nv,1+nat): (re|im, total|orbital l2, energy, summed|per type); the summed block is divided by loop_over.
extern "C" int rsrec_conductivity_cumulative(rsrec_handle h, const cplx *integrand, const cplx *integrand_at, int nv, int nv1, int nat,
                                             double wstep, int loop_over, double *sigma) {
  if (!h || !integrand || nv < 3 || nv1 < 1 || nv1 + 9 > nv + 1 || nat < 0 || (nat > 0 && !integrand_at) || loop_over < 1 || !sigma)
    return fail(RSREC_EINVAL, "rsrec_conductivity_cumulative: bad argument (need nv1 + 9 <= nv + 1)");
  CUDA_TRY(cudaSetDevice(h->dev));
  TRY(to_dev(h, h->post[0], integrand, (size_t)36 * nv));
  if (nat) TRY(to_dev(h, h->post[1], integrand_at, (size_t)36 * nv * nat));
  const size_t nout = (size_t)38 * nv * (1 + nat);
  TRY(dev_alloc(h->post[2], nout, false));
  CUDA_TRY(cudaMemsetAsync(h->post[2].p, 0, nout * sizeof(double), h->st));
  const int nthr = 38 * (1 + nat);
  k_cond_cumulative<<<(nthr + 63) / 64, 64, 0, h->st>>>((const double2 *)h->post[0].p, nat ? (const double2 *)h->post[1].p : nullptr, nv,
                                                       nv1 + 9, nat, wstep, (double)(float)loop_over, h->post[2].p);
  h->launches++;
  CUDA_TRY(cudaGetLastError());
  TRY(to_host(h, sigma, h->post[2].p, nout));
  CUDA_TRY(cudaStreamSynchronize(h->st));
  return RSREC_OK;
}

// ---- calculate_intersite_gf (green.f90:425-469) on the device-resident g0 of the pair units ----------------------
extern "C" int rsrec_intersite_gf(rsrec_handle h, int njij, const int32_t *pair_i, const int32_t *pair_j, int compact, cplx *gij,
                                  cplx *gji, cplx *gspin) {
  if (!h || njij < 0 || !pair_i || !pair_j || !gij || !gji) return fail(RSREC_EINVAL, "rsrec_intersite_gf: bad argument");
  if (njij == 0) return RSREC_OK;
  std::vector<int32_t> meta(2 * (size_t)njij);  // first unit, same-site flag
  int next = 0;
  for (int p = 0; p < njij; p++) {
    const int same = pair_i[p] == pair_j[p];
    meta[2 * p] = compact ? next : 4 * p;
    meta[2 * p + 1] = same;
    next += compact ? (same ? 1 : 4) : 4;
  }
  if (h->g0_units != next || h->g0_nv < 1)
    return fail(RSREC_EINVAL, "rsrec_intersite_gf: the device-resident g0 does not hold the units of these pairs (run a Green-function call on the pair units first)");
  CUDA_TRY(cudaSetDevice(h->dev));
  const int nv = h->g0_nv;
  const size_t nblk = (size_t)njij * nv;
  TRY(dev_alloc(h->post[5], 2 * nblk * BLKD, false));
  TRY(dev_alloc(h->post[6], gspin ? 8 * nblk * 162 : 1, false));
  TRY(dev_alloc(h->post[7], (2 * (size_t)njij + 1) / 2 + 1, false));
  CUDA_TRY(cudaMemcpyAsync(h->post[7].p, meta.data(), meta.size() * sizeof(int32_t), cudaMemcpyHostToDevice, h->st));
  double2 *d_gij = (double2 *)h->post[5].p, *d_gji = d_gij + nblk * BLKC;
  k_intersite_combine<<<grid_for(nblk * BLKC, 256, h->sms * 16), 256, 0, h->st>>>((const double2 *)h->g0all.p, nv, njij,
                                                                               (const int32_t *)h->post[7].p, d_gij, d_gji);
  h->launches++;
  if (gspin) {
    k_intersite_pauli<<<grid_for(nblk * 81, 256, h->sms * 16), 256, 0, h->st>>>(d_gij, d_gji, nv, njij, (double2 *)h->post[6].p);
    h->launches++;
  }
  CUDA_TRY(cudaGetLastError());
  TRY(to_host(h, gij, h->post[5].p, nblk * BLKD));
  TRY(to_host(h, gji, h->post[5].p + nblk * BLKD, nblk * BLKD));
  if (gspin) TRY(to_host(h, gspin, h->post[6].p, 8 * nblk * 162));
  CUDA_TRY(cudaStreamSynchronize(h->st));
  return RSREC_OK;
}

// ---- `type bands` (bands.f90): the consumers of g0 in the SCF loop, on the device-resident g0 ----------------------
static int bands_need_g0(H *h, const char *who) {
  if (!h) return fail(RSREC_EINVAL, "null handle");
  if (h->g0_units <= 0 || h->g0_nv <= 0) return fail(RSREC_EINVAL, std::string(who) + ": no on-site Green function on the device (call a Green-function entry point or rsrec_bands_set_g0 first)");
  return RSREC_OK;
}
// simpson_m over nint integrands x nord powers of the energy (d_y: (nv,nint)) -> host out (nord,nint)
static int bands_simpson(H *h, const double *d_y, int nv, int nint, int nord, int ord0, const double *ene, double edel, double fermi,
                         int nv1, double e1, double *out) {
  if (nv1 < 1 || nv1 + 2 > nv) return fail(RSREC_EINVAL, "bands: need 1 <= nv1 and nv1 + 2 <= nv (simpson_m reads Y(NPTS+2))");
  TRY(to_dev(h, h->post[4], ene, nv));
  const int n = nint * nord;
  TRY(dev_alloc(h->bands_out, std::max(n, 4), false));
  k_bands_simpson<<<(n * 32 + 127) / 128, 128, 0, h->st>>>(d_y, nv, nint, nord, ord0, h->post[4].p, edel, fermi, nv1, e1, h->bands_out.p);
  h->launches++;
  CUDA_TRY(cudaGetLastError());
  TRY(to_host(h, out, h->bands_out.p, n));
  CUDA_TRY(cudaStreamSynchronize(h->st));
  return RSREC_OK;
}

extern "C" {

int rsrec_bands_set_g0(rsrec_handle h, const cplx *g0, int nunits, int nv) {
  if (!h || !g0 || nunits < 1 || nv < 1) return fail(RSREC_EINVAL, "rsrec_bands_set_g0: bad argument");
  CUDA_TRY(cudaSetDevice(h->dev));
  TRY(g0_reserve(h, nunits, nv));
  CUDA_TRY(cudaMemcpyAsync(h->g0all.p, g0, (size_t)nunits * nv * BLKD * sizeof(double), cudaMemcpyHostToDevice, h->st));
  h->h2d_bytes += (long long)((size_t)nunits * nv * BLKD * sizeof(double));
  CUDA_TRY(cudaStreamSynchronize(h->st));
  return RSREC_OK;
}

int rsrec_bands_g0_shape(rsrec_handle h, int *nunits, int *nv) {
  if (!h || !nunits || !nv) return fail(RSREC_EINVAL, "rsrec_bands_g0_shape: bad argument");
  *nunits = h->g0_units; *nv = h->g0_nv;
  return RSREC_OK;
}

int rsrec_bands_get_g0(rsrec_handle h, cplx *g0) {
  TRY(bands_need_g0(h, "rsrec_bands_get_g0"));
  if (!g0) return fail(RSREC_EINVAL, "rsrec_bands_get_g0: bad argument");
  CUDA_TRY(cudaSetDevice(h->dev));
  TRY(to_host(h, g0, h->g0all.p, (size_t)h->g0_units * h->g0_nv * BLKD));
  CUDA_TRY(cudaStreamSynchronize(h->st));
  return RSREC_OK;
}

int rsrec_bands_dos(rsrec_handle h, double *dtot, double *dosia, double *dosial) {
  TRY(bands_need_g0(h, "rsrec_bands_dos"));
  if (!dtot) return fail(RSREC_EINVAL, "rsrec_bands_dos: bad argument");
  CUDA_TRY(cudaSetDevice(h->dev));
  const int nv = h->g0_nv, nu = h->g0_units;
  const size_t n1 = (size_t)nv * nu;
  TRY(dev_alloc(h->bands_y, nv + n1 * 19, false));
  double *d_dtot = h->bands_y.p, *d_ia = d_dtot + nv, *d_ial = d_ia + n1;
  k_bands_dtot<<<(nv + 127) / 128, 128, 0, h->st>>>((const double2 *)h->g0all.p, nv, nu, d_dtot);
  h->launches++;
  if (dosia || dosial) {
    k_bands_ldos<<<(unsigned)((n1 + 127) / 128), 128, 0, h->st>>>((const double2 *)h->g0all.p, nv, nu, dosia ? d_ia : nullptr, dosial ? d_ial : nullptr);
    h->launches++;
  }
  CUDA_TRY(cudaGetLastError());
  TRY(comm_allreduce_dev(h, d_dtot, nv));  // MPI_ALLREDUCE of dtot over the unit shards (bands.f90:270-276), on the device
  TRY(to_host(h, dtot, d_dtot, nv));
  if (dosia) TRY(to_host(h, dosia, d_ia, n1));
  if (dosial) TRY(to_host(h, dosial, d_ial, n1 * 18));
  CUDA_TRY(cudaStreamSynchronize(h->st));
  return RSREC_OK;
}

int rsrec_bands_fermi(rsrec_handle h, const double *dtot, int nv, double edel, double energy_min, double qqv, int fix_fermi,
                      double *fermi, int *nv1, double *e1, int *ifail) {
  if (!h || !fermi || !nv1 || !e1 || nv < 3 || edel == 0.0 || (!fix_fermi && !dtot)) return fail(RSREC_EINVAL, "rsrec_bands_fermi: bad argument");
  if (ifail) *ifail = 0;
  if (fix_fermi) {  // bands.f90:338-341
    const int ik1 = (int)std::llround((*fermi - energy_min) / edel);  // nint
    *nv1 = ik1;
    *e1 = energy_min + (ik1 - 1) * edel;
    return RSREC_OK;
  }
  CUDA_TRY(cudaSetDevice(h->dev));
  TRY(to_dev(h, h->post[4], dtot, nv));
  TRY(dev_alloc(h->bands_out, 4, false));
  const size_t fsm = (size_t)(nv / 2 + 1) * sizeof(double);  // one value per Simpson panel
  double *pan_g = nullptr;
  if (fsm > 48 * 1024) { TRY(dev_alloc(h->bands_y, nv / 2 + 1, false)); pan_g = h->bands_y.p; }
  k_bands_fermi<<<1, 256, pan_g ? 0 : fsm, h->st>>>(h->post[4].p, nv, edel, energy_min, qqv, *fermi, *nv1, h->bands_out.p, pan_g);
  h->launches++;
  CUDA_TRY(cudaGetLastError());
  double res[4];
  TRY(to_host(h, res, h->bands_out.p, 4));
  CUDA_TRY(cudaStreamSynchronize(h->st));
  *fermi = res[0]; *e1 = res[1]; *nv1 = (int)res[2];
  if (ifail) *ifail = (int)res[3];
  return RSREC_OK;
}

int rsrec_bands_magnetic_moments(rsrec_handle h, const double *ene, double edel, double fermi, int nv1, double e1, double *mom0,
                                 double *mom1) {
  TRY(bands_need_g0(h, "rsrec_bands_magnetic_moments"));
  if (!ene || !mom0 || !mom1) return fail(RSREC_EINVAL, "rsrec_bands_magnetic_moments: bad argument");
  CUDA_TRY(cudaSetDevice(h->dev));
  const int nv = h->g0_nv, nu = h->g0_units;
  const size_t n = (size_t)nv * nu * 3;
  TRY(dev_alloc(h->bands_y, n, false));
  k_bands_spin<<<(unsigned)((n + 127) / 128), 128, 0, h->st>>>((const double2 *)h->g0all.p, nv, nu, h->bands_y.p);
  h->launches++;
  CUDA_TRY(cudaGetLastError());
  std::vector<double> out((size_t)6 * nu);  // (order 0..1, direction, unit)
  TRY(bands_simpson(h, h->bands_y.p, nv, 3 * nu, 2, 0, ene, edel, fermi, nv1, e1, out.data()));
  for (int i = 0; i < 3 * nu; i++) { mom0[i] = out[2 * i]; mom1[i] = out[2 * i + 1]; }
  return RSREC_OK;
}

int rsrec_bands_moments(rsrec_handle h, int channels_ldos, const double *ene, double edel, double fermi, int nv1, double e1,
                        const double *mom, double *occ, double *lmom) {
  TRY(bands_need_g0(h, "rsrec_bands_moments"));
  if (!ene || !mom || !occ || !lmom || channels_ldos < 0 || channels_ldos > h->g0_nv) return fail(RSREC_EINVAL, "rsrec_bands_moments: bad argument");
  CUDA_TRY(cudaSetDevice(h->dev));
  const int nv = h->g0_nv, nu = h->g0_units;
  const size_t n = (size_t)nv * nu * 9;
  TRY(dev_alloc(h->bands_y, n, false));
  TRY(to_dev(h, h->post[3], mom, (size_t)3 * nu));
  k_bands_dspd<<<(unsigned)((n + 127) / 128), 128, 0, h->st>>>((const double2 *)h->g0all.p, nv, channels_ldos, nu, h->post[3].p, h->bands_y.p);
  h->launches++;
  CUDA_TRY(cudaGetLastError());
  std::vector<double> out((size_t)27 * nu);  // (order 0..2, q 0..8, unit)
  TRY(bands_simpson(h, h->bands_y.p, nv, 9 * nu, 3, 0, ene, edel, fermi, nv1, e1, out.data()));
  for (int u = 0; u < nu; u++) {
    for (int q = 0; q < 6; q++)
      for (int k = 0; k < 3; k++) occ[k + 3 * (q + 6 * u)] = out[k + 3 * (q + 9 * (size_t)u)];
    for (int d = 0; d < 3; d++) lmom[d + 3 * u] = -(out[3 * (6 + d + 9 * (size_t)u)] / PI_RP);  // bands.f90:1144-1146
  }
  return RSREC_OK;
}

int rsrec_bands_band_energy(rsrec_handle h, const double *dtot, int nv, const double *ene, double edel, double fermi, int nv1,
                            double e1, double *eband) {
  if (!h || !dtot || !ene || !eband || nv < 3) return fail(RSREC_EINVAL, "rsrec_bands_band_energy: bad argument");
  CUDA_TRY(cudaSetDevice(h->dev));
  TRY(to_dev(h, h->bands_y, dtot, nv));
  return bands_simpson(h, h->bands_y.p, nv, 1, 1, 1, ene, edel, fermi, nv1, e1, eband);
}

}  // extern "C"

// rsrec.hpp -- C++ host-side mirror of the reference's `type recursion` (source/recursion.f90:41-116) on top of the
// C ABI in rsrec.h.  Header-only; same procedure names, same result members (column-major, reference shapes), same
// error behaviour (a fatal condition throws where the reference calls g_logger%fatal).  The structs `lattice`,
// `hamiltonian`, `control`, `energy` carry exactly the members of the reference types that the recursion reads.
#pragma once
#include "rsrec.h"

#include <cmath>
#include <complex>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

namespace rsrec {

using cplx = std::complex<double>;

struct fatal : std::runtime_error {  // g_logger%fatal (logger.f90:186-193)
  int code;
  fatal(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};
inline void check(int rc) {
  if (rc != RSREC_OK) throw fatal(rc, rsrec_last_error());
}

struct lattice {  // lattice.f90:144-309 (members on the hot path)
  int kk = 0, ncols = 0, ntype = 1, nmax = 0;
  std::vector<int32_t> nn;      // (kk, ncols) column-major
  std::vector<int32_t> iz;      // (kk)
  std::vector<int32_t> irec;    // (nrec) recursion sites
  std::vector<int32_t> ijpair;  // (njij, 2) column-major
  std::vector<double> cr;       // optional (3, kk) coordinates: work ordering only
  int nslot() const {
    int m = 0;
    for (int i = 0; i < kk; i++) m = nn[i] > m ? nn[i] : m;
    return m + 1;  // hamiltonian.f90:294
  }
};
struct hamiltonian {  // hamiltonian.f90:52-66
  std::vector<cplx> ee, eeo, hall, hallo, lsham, enim, v_a, v_b, vo_a, vo_b;
  bool hoh = false;
};
struct control {  // control.f90:356-384
  int lld = 16, cond_ll = 200;
};
struct energy {  // energy.f90: energy_min/max, channels_ldos, fermi and the mesh e_mesh builds (175-208)
  double energy_min = -1.5, energy_max = 1.5, fermi = 0.0;
  int channels_ldos = 2500;
  double edel = 0.0;
  int nv1 = 0, ik1 = 0;
  bool fix_fermi = false;
  std::vector<double> ene;
  void e_mesh() {
    if (channels_ldos % 2 == 0) nv1 = channels_ldos + 1;
    else { nv1 = channels_ldos; channels_ldos -= 1; }
    ik1 = nv1;
    edel = (energy_max - energy_min) / channels_ldos;
    edel = (fermi - energy_min) / std::round((fermi - energy_min) / edel);  // nint
    ene.resize(channels_ldos + 10);
    for (int i = 0; i < channels_ldos + 10; i++) ene[i] = energy_min + edel * i;
  }
};
struct mpi_vars {  // mpi.f90:32-58 (get_mpi_variables)
  int start_atom = 1, end_atom = 0, atoms_per_process = 0;
  static mpi_vars get(int rank, int numprocs, int n) {
    mpi_vars v;
    v.atoms_per_process = n / numprocs;
    const int rem = n % numprocs;
    if (rank < rem) {
      v.atoms_per_process += 1;
      v.start_atom = rank * v.atoms_per_process + 1;
    } else {
      v.start_atom = rank * v.atoms_per_process + rem + 1;
    }
    v.end_atom = v.start_atom + v.atoms_per_process - 1;
    return v;
  }
};

class recursion {
 public:
  // results, reference shapes
  std::vector<cplx> a_b, b2_b;   // (18,18,lld,nunits)
  std::vector<double> a, b2;     // (lld,18,nunits)   [this%a(:,:,:,1)]
  std::vector<cplx> mu_n;        // (18,18,2*lld+2,nunits)
  std::vector<cplx> mu_nm_stochastic;

  recursion(const hamiltonian &h, const lattice &l, const control &c, const energy &e, int device = 0, int rank = 0,
            int numprocs = 1)
      : ham_(h), lat_(l), ctl_(c), en_(e), rank_(rank), np_(numprocs) {
    check(rsrec_create(&h_, device, l.kk, l.ncols, l.nslot(), l.ntype, l.nmax));
    upload();
  }
  ~recursion() { rsrec_destroy(h_); }
  recursion(const recursion &) = delete;

  void upload() {
    check(rsrec_set_lattice(h_, lat_.nn.data(), lat_.iz.data()));
    if (!lat_.cr.empty()) check(rsrec_set_positions(h_, lat_.cr.data()));
    check(rsrec_set_hamiltonian(h_, p(ham_.ee), p(ham_.eeo), p(ham_.hall), p(ham_.hallo), p(ham_.lsham), p(ham_.enim),
                                ham_.hoh));
    if (!ham_.v_a.empty()) {
      check(rsrec_set_operator(h_, 'a', p(ham_.v_a), p(ham_.vo_a)));
      check(rsrec_set_operator(h_, 'b', p(ham_.v_b), p(ham_.vo_b)));
    }
  }

  void recur_b() {  // recursion.f90:1807-1866
    const auto sites = local_sites();
    const int n = (int)sites.size(), lld = ctl_.lld;
    a_b.assign((size_t)324 * lld * n, 0.0);
    b2_b.assign((size_t)324 * lld * n, 0.0);
    check(rsrec_lanczos_block(h_, n, sites.data(), nullptr, nullptr, nullptr, lld, rc(a_b), rc(b2_b)));
    a.assign((size_t)lld * 18 * n, 0.0);
    b2.assign((size_t)lld * 18 * n, 0.0);
    for (int u = 0; u < n; u++)
      for (int l = 0; l < 18; l++)
        for (int ll = 0; ll < lld; ll++) {
          a[ll + (size_t)lld * (l + 18 * u)] = a_b[(l + 18 * l) + (size_t)324 * (ll + (size_t)lld * u)].real();
          b2[ll + (size_t)lld * (l + 18 * u)] = b2_b[(l + 18 * l) + (size_t)324 * (ll + (size_t)lld * u)].real();
        }
  }
  void recur() {  // recursion.f90:3485-3532
    const auto sites = local_sites();
    const int n = (int)sites.size(), lld = ctl_.lld;
    a.assign((size_t)lld * 18 * n, 0.0);
    b2.assign((size_t)lld * 18 * n, 0.0);
    check(rsrec_lanczos_scalar(h_, n, sites.data(), lld, a.data(), b2.data()));
  }
  void zsqr() {  // recursion.f90:1980-2023
    const int lld = ctl_.lld;
    check(rsrec_zsqr(h_, rc(b2_b), lld, (int)(b2_b.size() / ((size_t)324 * lld))));
  }
  void chebyshev_recur() {  // recursion.f90:3057-3130
    const auto sites = local_sites();
    const int n = (int)sites.size(), lld = ctl_.lld;
    mu_n.assign((size_t)324 * (2 * lld + 2) * n, 0.0);
    check(rsrec_cheb_moments(h_, n, sites.data(), nullptr, nullptr, nullptr, lld, scale(), shift(), rc(mu_n)));
  }
  // ---- the exchange step inside the library (one process per GPU = one MPI rank of the reference, mpi.f90:32-58) ----
  // rank 0 creates the id, the host broadcasts its 128 bytes (MPI_Bcast), every rank attaches its handle
  static std::vector<unsigned char> comm_unique_id() {
    std::vector<unsigned char> id(RSREC_COMM_ID_BYTES);
    check(rsrec_comm_unique_id(id.data()));
    return id;
  }
  void comm_init(int nranks, int rank, const std::vector<unsigned char> &id) {
    check(rsrec_comm_init(h_, nranks, rank, id.data()));
    rank_ = rank; np_ = nranks;
  }
  // recur_b over ALL recursion sites of the job: a_b, b2_b (18,18,lld,nrec) gathered on the device, on every rank
  void recur_b_sharded() {
    const int n = (int)lat_.irec.size(), lld = ctl_.lld;
    a_b.assign((size_t)324 * lld * n, 0.0);
    b2_b.assign((size_t)324 * lld * n, 0.0);
    check(rsrec_lanczos_block_sharded(h_, n, lat_.irec.data(), nullptr, nullptr, nullptr, lld, rc(a_b), rc(b2_b)));
  }
  // MPI_ALLREDUCE(MPI_IN_PLACE, x, n, MPI_DOUBLE_PRECISION, MPI_SUM) (bands.f90:270-275)
  void allreduce(std::vector<double> &x) { check(rsrec_allreduce(h_, x.data(), (long long)x.size(), 0)); }

  // ---- device time per phase under the reference's g_timer labels (recursion.f90:1902-1970, 3104-3127) ----
  void phase_timing(bool on) { check(rsrec_phase_timing(h_, on ? 1 : 0)); }
  std::vector<std::pair<std::string, double>> phase_read() {  // (label, milliseconds) of the phases that ran
    const int n = rsrec_phase_count();
    std::vector<double> ms(n);
    std::vector<long long> calls(n);
    check(rsrec_phase_read(h_, ms.data(), calls.data()));
    std::vector<std::pair<std::string, double>> out;
    for (int k = 0; k < n; k++)
      if (calls[k] > 0) out.emplace_back(rsrec_phase_label(k), ms[k]);
    return out;
  }

  // recur_b_ij / chebyshev_recur_ij (recursion.f90:1655-1737 / 2376-2487): slot ij_loc*4-4+reci
  void recur_b_ij() { pair_run(true); }
  void chebyshev_recur_ij() { pair_run(false); }

  double scale() const { return (en_.energy_max - en_.energy_min) / (2 - 0.3); }  // recursion.f90:3078
  double shift() const { return (en_.energy_max + en_.energy_min) / 2; }          // recursion.f90:3079
  rsrec_handle handle() const { return h_; }
  const control &ctl() const { return ctl_; }
  const energy &en() const { return en_; }
  int nunits_local() const { return (int)local_sites().size(); }
  std::vector<int32_t> sites_local() const { return local_sites(); }

 private:
  static const rsrec_cplx *p(const std::vector<cplx> &v) { return v.empty() ? nullptr : reinterpret_cast<const rsrec_cplx *>(v.data()); }
  static rsrec_cplx *rc(std::vector<cplx> &v) { return reinterpret_cast<rsrec_cplx *>(v.data()); }
  std::vector<int32_t> local_sites() const {
    const auto v = mpi_vars::get(rank_, np_, (int)lat_.irec.size());
    return std::vector<int32_t>(lat_.irec.begin() + (v.start_atom - 1), lat_.irec.begin() + v.end_atom);
  }
  void pair_run(bool lanczos) {
    const int njij = (int)lat_.ijpair.size() / 2, lld = ctl_.lld;
    const auto v = mpi_vars::get(rank_, np_, njij);
    const int nloc = v.atoms_per_process;
    std::vector<int32_t> si, sj, slots;
    std::vector<cplx> as, bs;
    const double s = 1.0 / std::sqrt(2.0);
    const cplx sg[4] = {{s, 0}, {-s, 0}, {0, s}, {0, -s}};
    for (int ij = v.start_atom; ij <= v.end_atom; ij++) {
      const int i = lat_.ijpair[ij - 1], j = lat_.ijpair[ij - 1 + njij];
      for (int reci = 0; reci < 4; reci++) {
        if (i == j && reci > 0) continue;
        si.push_back(i); sj.push_back(j);
        as.push_back(i == j ? cplx(1.0) : cplx(s));
        bs.push_back(i == j ? cplx(1.0) : sg[reci]);
        slots.push_back((ij - v.start_atom) * 4 + reci);
      }
    }
    const int n = (int)si.size();
    const size_t per = lanczos ? (size_t)324 * lld : (size_t)324 * (2 * lld + 2);
    std::vector<cplx> r1(per * n), r2(lanczos ? per * n : 0);
    if (lanczos)
      check(rsrec_lanczos_block(h_, n, si.data(), sj.data(), p(as), p(bs), lld, rc(r1), rc(r2)));
    else
      check(rsrec_cheb_moments(h_, n, si.data(), sj.data(), p(as), p(bs), lld, scale(), shift(), rc(r1)));
    std::vector<cplx> &o1 = lanczos ? a_b : mu_n;
    o1.assign(per * 4 * nloc, 0.0);
    if (lanczos) b2_b.assign(per * 4 * nloc, 0.0);
    for (int u = 0; u < n; u++) {
      std::copy(r1.begin() + per * u, r1.begin() + per * (u + 1), o1.begin() + per * slots[u]);
      if (lanczos) std::copy(r2.begin() + per * u, r2.begin() + per * (u + 1), b2_b.begin() + per * slots[u]);
    }
  }
  const hamiltonian &ham_;
  const lattice &lat_;
  control ctl_;
  energy en_;
  int rank_, np_;
  rsrec_handle h_ = nullptr;
};

// Mirror of the reference's `type green` (green.f90) for the procedures that consume the recursion results.
class green {
 public:
  std::vector<cplx> g0;  // (18,18,nv,nunits)
  bool sym_term = false;  // control%sym_term
  green(recursion &r, energy &e) : rec_(r), en_(e) {
    if (en_.ene.empty()) en_.e_mesh();
  }
  void block_green() {  // green.f90:588-621 (after recursion%zsqr, like self%run_dos)
    const int lld = rec_.ctl().lld, nv = (int)en_.ene.size(), na = (int)(rec_.a_b.size() / ((size_t)324 * lld));
    g0.assign((size_t)324 * nv * na, 0.0);
    check(rsrec_block_green(rec_.handle(), cp(rec_.a_b), cp(rec_.b2_b), na, lld, en_.ene.data(), nv, sym_term, mp(g0)));
  }
  void chebyshev_green(std::vector<cplx> *mu_ng = nullptr) {  // green.f90:1030-1108
    const int lld = rec_.ctl().lld, nv = (int)en_.ene.size(), na = (int)(rec_.mu_n.size() / ((size_t)324 * (2 * lld + 2)));
    g0.assign((size_t)324 * nv * na, 0.0);
    if (mu_ng) mu_ng->assign(rec_.mu_n.size(), 0.0);
    check(rsrec_chebyshev_green(rec_.handle(), cp(rec_.mu_n), na, lld, en_.ene.data(), nv, en_.energy_min, en_.energy_max,
                                mu_ng ? mp(*mu_ng) : nullptr, mp(g0)));
  }
  // fused self%run_recursion + self%run_dos (block path); download_g0 = false leaves g0 on the device only (for `bands`)
  void recur_b_green(bool download_g0 = true) {
    const auto sites = rec_.sites_local();
    const int n = (int)sites.size(), lld = rec_.ctl().lld, nv = (int)en_.ene.size();
    rec_.a_b.assign((size_t)324 * lld * n, 0.0);
    rec_.b2_b.assign((size_t)324 * lld * n, 0.0);
    g0.assign(download_g0 ? (size_t)324 * nv * n : 0, 0.0);
    check(rsrec_recur_b_green(rec_.handle(), n, sites.data(), lld, en_.ene.data(), nv, sym_term, mp(rec_.a_b), mp(rec_.b2_b),
                              download_g0 ? mp(g0) : nullptr));
  }
  recursion &rec() { return rec_; }
  energy &en() { return en_; }

 private:
  static const rsrec_cplx *cp(const std::vector<cplx> &v) { return reinterpret_cast<const rsrec_cplx *>(v.data()); }
  static rsrec_cplx *mp(std::vector<cplx> &v) { return reinterpret_cast<rsrec_cplx *>(v.data()); }
  recursion &rec_;
  energy &en_;
};

// `type bands` (bands.f90): consumes the g0 the last Green-function call left on the device
class bands {
 public:
  double qqv = 0.0, e1 = 0.0, eband = 0.0;
  int nv1 = 0, ifail = 0, nsp = 2;
  std::vector<double> dtot;             // (nv)
  std::vector<double> mom0, mom1, mom;  // (3,nunits): mx,my,mz; potential%mom1; unit vectors potential%mom
  std::vector<double> occ, lmom;        // (3,6,nunits) = sgef,pmef,smef per (l, spin); (3,nunits)
  bands(green &g, double valence) : qqv(valence), gr_(g) {}
  void calculate_fermi() {  // bands.f90:227-347 (single rank: no all-reduce of dtot)
    int nu = 0, nv = 0;
    check(rsrec_bands_g0_shape(h(), &nu, &nv));
    dtot.assign(nv, 0.0);
    check(rsrec_bands_dos(h(), dtot.data(), nullptr, nullptr));
    energy &en = gr_.en();
    double fermi = en.fermi, e1_new = 0.0;
    int n1 = en.ik1;
    check(rsrec_bands_fermi(h(), dtot.data(), nv, en.edel, en.energy_min, qqv, en.fix_fermi, &fermi, &n1, &e1_new, &ifail));
    if (ifail == 0) { en.fermi = fermi; nv1 = n1; e1 = e1_new; }
  }
  void calculate_magnetic_moments() {  // bands.f90:791-855
    const int nu = units();
    mom0.assign(3 * (size_t)nu, 0.0); mom1.assign(3 * (size_t)nu, 0.0); mom.assign(3 * (size_t)nu, 0.0);
    energy &en = gr_.en();
    check(rsrec_bands_magnetic_moments(h(), en.ene.data(), en.edel, en.fermi, nv1, e1, mom0.data(), mom1.data()));
    for (int u = 0; u < nu; u++) {
      const double *m = &mom0[3 * (size_t)u];
      const double mtot = std::sqrt(m[0] * m[0] + m[1] * m[1] + m[2] * m[2]) + 1.0e-15;
      for (int d = 0; d < 3; d++) mom[3 * (size_t)u + d] = nsp < 3 ? (d == 2 ? 1.0 : 0.0) : m[d] / mtot;
    }
  }
  void calculate_moments() {  // bands.f90:409-524 with calculate_orbital_moments 1075-1156
    const int nu = units();
    if (mom.size() != 3 * (size_t)nu) { mom.assign(3 * (size_t)nu, 0.0); for (int u = 0; u < nu; u++) mom[3 * (size_t)u + 2] = 1.0; }
    occ.assign(18 * (size_t)nu, 0.0); lmom.assign(3 * (size_t)nu, 0.0);
    energy &en = gr_.en();
    check(rsrec_bands_moments(h(), en.channels_ldos, en.ene.data(), en.edel, en.fermi, nv1, e1, mom.data(), occ.data(), lmom.data()));
  }
  // potential%ql(q, l, isp) of unit u (q = 1..3, l = 0..2, isp = 1..2), bands.f90:493-495
  double ql(int q, int l, int isp, int u) const {
    const double *o = &occ[3 * ((size_t)(l + 3 * (isp - 1)) + 6 * (size_t)u)];
    if (q == 1) return o[0];
    if (q == 2) return 0.0;
    const double cg = o[1] / o[0];
    return o[2] - 2.0 * cg * o[1] + cg * cg * o[0];
  }
  void calculate_band_energy() {  // bands.f90:354-359
    energy &en = gr_.en();
    check(rsrec_bands_band_energy(h(), dtot.data(), (int)dtot.size(), en.ene.data(), en.edel, en.fermi, nv1, e1, &eband));
  }

 private:
  rsrec_handle h() { return gr_.rec().handle(); }
  int units() { int nu = 0, nv = 0; check(rsrec_bands_g0_shape(h(), &nu, &nv)); return nu; }
  green &gr_;
};

}  // namespace rsrec

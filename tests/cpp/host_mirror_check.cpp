// host_mirror_check.cpp -- exercises include/rsrec.hpp (the C++ mirror of `type recursion`) against arrays dumped by
// the Python test (tests/test_gpu_cpp_host.py): reads lattice + Hamiltonian + expected oracle results from a flat
// binary file, runs recur_b / chebyshev_recur / recur_b_ij through the C ABI and prints the max relative errors.
#include "../../include/rsrec.hpp"
#include <cstdio>
#include <cstdlib>

template <class T> static std::vector<T> rd(FILE *f) {
  int64_t n = 0;
  if (fread(&n, 8, 1, f) != 1) { fprintf(stderr, "short read\n"); exit(2); }
  std::vector<T> v(n);
  if (n && fread(v.data(), sizeof(T), n, f) != (size_t)n) { fprintf(stderr, "short read\n"); exit(2); }
  return v;
}
static double relerr(const std::vector<rsrec::cplx> &x, const std::vector<rsrec::cplx> &r) {
  double e = 0, m = 0;
  if (x.size() != r.size()) return 1e300;
  for (size_t i = 0; i < x.size(); i++) { e = std::max(e, std::abs(x[i] - r[i])); m = std::max(m, std::abs(r[i])); }
  return e / m;
}
static double relerr_d(const std::vector<double> &x, const std::vector<double> &r) {
  double e = 0, m = 0;
  if (x.size() != r.size()) return 1e300;
  for (size_t i = 0; i < x.size(); i++) { e = std::max(e, std::abs(x[i] - r[i])); m = std::max(m, std::abs(r[i])); }
  return e / m;
}
int main(int argc, char **argv) {
  FILE *f = fopen(argv[1], "rb");
  if (!f) return 2;
  auto hdr = rd<int32_t>(f);  // kk ncols ntype nmax lld
  rsrec::lattice lat;
  lat.kk = hdr[0]; lat.ncols = hdr[1]; lat.ntype = hdr[2]; lat.nmax = hdr[3];
  rsrec::control ctl; ctl.lld = hdr[4];
  lat.nn = rd<int32_t>(f); lat.iz = rd<int32_t>(f); lat.irec = rd<int32_t>(f); lat.ijpair = rd<int32_t>(f);
  rsrec::hamiltonian ham;
  ham.ee = rd<rsrec::cplx>(f); ham.lsham = rd<rsrec::cplx>(f); ham.hall = rd<rsrec::cplx>(f);
  auto emm = rd<double>(f);
  rsrec::energy en; en.energy_min = emm[0]; en.energy_max = emm[1];
  auto ref_a = rd<rsrec::cplx>(f), ref_b2 = rd<rsrec::cplx>(f), ref_mu = rd<rsrec::cplx>(f), ref_aij = rd<rsrec::cplx>(f);
  auto mesh = rd<double>(f);   // channels_ldos, fermi
  auto ref_g0 = rd<rsrec::cplx>(f), ref_gk = rd<rsrec::cplx>(f);
  en.channels_ldos = (int)mesh[0]; en.fermi = mesh[1];
  auto bsc = rd<double>(f);   // qqv, fermi, nv1, e1, eband
  auto ref_occ = rd<double>(f), ref_m0 = rd<double>(f), ref_lmom = rd<double>(f);
  fclose(f);
  try {
    rsrec::recursion rec(ham, lat, ctl, en);
    rec.recur_b();
    printf("recur_b a_b %.3e b2_b %.3e\n", relerr(rec.a_b, ref_a), relerr(rec.b2_b, ref_b2));
    rsrec::green gr(rec, en);
    rec.zsqr();
    gr.block_green();
    printf("block_green g0 %.3e\n", relerr(gr.g0, ref_g0));
    const auto staged = gr.g0;
    gr.recur_b_green();
    printf("recur_b_green g0 %.3e\n", relerr(gr.g0, staged));
    {  // bands on the device-resident g0 of the fused call (never downloaded)
      gr.recur_b_green(false);
      rsrec::bands bd(gr, bsc[0]);
      bd.calculate_fermi(); bd.calculate_magnetic_moments(); bd.calculate_moments(); bd.calculate_band_energy();
      printf("bands fermi %.3e nv1 %.1f eband %.3e occ %.3e mom0 %.3e lmom %.3e\n", std::abs(en.fermi - bsc[1]),
             std::abs((double)bd.nv1 - bsc[2]), std::abs(bd.eband - bsc[4]) / std::abs(bsc[4]), relerr_d(bd.occ, ref_occ),
             relerr_d(bd.mom0, ref_m0), relerr_d(bd.lmom, ref_lmom));
      en.fermi = mesh[1];
    }
    rec.chebyshev_recur();
    printf("chebyshev_recur mu_n %.3e\n", relerr(rec.mu_n, ref_mu));
    gr.chebyshev_green();
    printf("chebyshev_green g0 %.3e\n", relerr(gr.g0, ref_gk));
    rec.recur_b_ij();
    printf("recur_b_ij a_b %.3e\n", relerr(rec.a_b, ref_aij));
    {  // exchange inside the library with a single-rank communicator, and the g_timer phase labels
      rec.comm_init(1, 0, rsrec::recursion::comm_unique_id());
      rec.phase_timing(true);
      rec.recur_b_sharded();
      const auto ph = rec.phase_read();
      std::vector<double> x = {1.5, 2.5};
      rec.allreduce(x);
      printf("sharded a_b %.3e phases %d first %s allreduce %.3e\n", relerr(rec.a_b, ref_a), (int)ph.size(),
             ph.empty() ? "-" : ph[0].first.c_str(), std::abs(x[0] - 1.5) + std::abs(x[1] - 2.5));
      rec.phase_timing(false);
    }
    rsrec::energy bad; bad.energy_min = -0.05; bad.energy_max = 0.05;
    rsrec::control c40; c40.lld = 40;
    rsrec::recursion rec2(ham, lat, c40, bad);
    try { rec2.chebyshev_recur(); printf("diverged NOT raised\n"); }
    catch (const rsrec::fatal &e) { printf("fatal %d %s\n", e.code, e.what()); }
  } catch (const rsrec::fatal &e) {
    printf("FATAL %d %s\n", e.code, e.what());
    return 1;
  }
  return 0;
}

#!/usr/bin/env python
"""Times BASELINE.json configs 1-4 (full sizes) through the reference-facing API (host arrays in, host arrays out),
with the CPU oracle beside it on a bounded number of steps, and checks parity where the oracle runs in seconds.
Output: one JSON object per config (stdout) -> profiles/rNN_configs.json."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rslmtoasa_b200 import Recursion, Control, Energy, Green, Dos, Conductivity, synthetic as S  # noqa: E402
from oracle import oracle as O  # noqa: E402

EMIN, EMAX = -2.0, 2.0


def relerr(x, r):
    return float(np.abs(x - r).max() / np.abs(r).max())


def timed(fn, reps=3):
    fn()
    best = 1e30
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        best = min(best, time.perf_counter() - t0)
    return best


def run(name, lat, ham, lld, what, cpu=True, **kw):
    ctl = Control(lld=lld, **{k: v for k, v in kw.items() if k in ("cond_ll", "cond_calctype")})
    extra = {k: v for k, v in kw.items() if k in ("ijpair", "atlist", "phases")}
    rec = Recursion(ham, lat, ctl, Energy(EMIN, EMAX), **extra)
    out = {"config": name, "kk": lat.kk, "ntype": lat.ntype, "nmax": lat.nmax, "nslots": lat.ncols, "hoh": bool(ham.hoh),
           "units": int(len(lat.irec)), "lld": lld, "what": what}
    a, b = O.cheb_scale(EMIN, EMAX)
    orc = O.Oracle(lat, ham) if cpu else None
    if what == "recur_b":
        t = timed(rec.recur_b)
        steps = (lld - 1) * len(lat.irec)
        if cpu:
            t0 = time.perf_counter(); ra, rb = orc.lanczos_block(lat.irec, lld); tc = time.perf_counter() - t0
            out["relerr_a_b"], out["relerr_b2_b"] = relerr(rec.a_b, ra), relerr(rec.b2_b, rb)
    elif what == "chebyshev_recur":
        t = timed(rec.chebyshev_recur)
        steps = lld * len(lat.irec)
        if cpu:
            t0 = time.perf_counter(); rm, _ = orc.cheb_moments(lat.irec, lld, a, b); tc = time.perf_counter() - t0
            out["relerr_mu_n"] = relerr(rec.mu_n, rm)
    elif what == "recur":
        t = timed(rec.recur)
        steps = (lld - 1) * 18 * len(lat.irec)   # 18 scalar recursions per site
        if cpu:
            t0 = time.perf_counter(); ra, rb = orc.lanczos_scalar(lat.irec, lld); tc = time.perf_counter() - t0
            out["relerr_a"], out["relerr_b2"] = relerr(rec.a[..., 0], ra), relerr(rec.b2[..., 0], rb)
    elif what == "kubo":
        M = kw["cond_ll"]
        t = timed(rec.compute_moments_stochastic, reps=1)
        nst = len(kw["atlist"]) if "atlist" in kw else kw["phases"].shape[1]
        steps = 3 * M * nst                       # SpMV-equivalents: M left + M right + M velocity
        if cpu:
            t0 = time.perf_counter()
            rm = orc.kubo_moments(M, a, b, start_sites=kw.get("atlist"), phases=kw.get("phases"))
            tc = time.perf_counter() - t0
            out["relerr_mu_nm"] = relerr(rec.mu_nm_stochastic, rm)
    out.update({"gpu_seconds": t, "gpu_steps_per_s": steps / t, "steps": steps, "launches": rec.launch_count})
    if cpu:
        out.update({"cpu_seconds": tc, "cpu_steps_per_s": steps / tc, "cpu_threads": O.lib().orc_get_max_threads(),
                    "speedup": tc / t})
    print(json.dumps(out), flush=True)
    rec.close()


def run_post(name, lat, ham, lld, what, channels=2500, cpu=True, **kw):
    """the consumers either side of the recursion at the reference's mesh size
    (channels_ldos + 10 energies), host arrays in / host arrays out, CPU oracle beside it."""
    ctl = Control(lld=lld, **{k: v for k, v in kw.items() if k in ("cond_ll", "cond_calctype")})
    en = Energy(EMIN, EMAX, channels_ldos=channels, fermi=0.0)
    rec = Recursion(ham, lat, ctl, en, **{k: v for k, v in kw.items() if k in ("atlist", "phases")})
    out = {"config": name, "kk": lat.kk, "units": int(len(lat.irec)), "lld": lld, "what": what, "nv": channels + 10}
    l0 = rec.launch_count
    if what == "block_green":
        rec.recur_b(); rec.zsqr()
        g = Green(rec)
        t = timed(g.block_green)
        if cpu:
            t0 = time.perf_counter(); ref = O.block_green(rec.a_b, rec.b2_b, g.ene); tc = time.perf_counter() - t0
            out["relerr_g0"] = relerr(g.g0, ref)
        out["work"] = "%d energies x %d levels x %d units: 18x18 complex LU inverse + 2 products each" % (channels + 10, lld - 1, len(lat.irec))
    elif what == "chebyshev_green":
        rec.chebyshev_recur()
        g = Green(rec)
        t = timed(g.chebyshev_green)
        if cpu:
            t0 = time.perf_counter(); _, ref = O.chebyshev_green(rec.mu_n, g.ene, EMIN, EMAX); tc = time.perf_counter() - t0
            out["relerr_g0"] = relerr(g.g0, ref)
    elif what == "sgreen":
        rec.recur()
        g = Green(rec)
        na = len(lat.irec)
        dw, cs = np.ones((18, na)), np.zeros((18, na))
        t = timed(lambda: g.sgreen(dw, cs, 1))
        if cpu:
            t0 = time.perf_counter(); ref = O.sgreen(rec.a, rec.b2, 1, g.ene, dw, cs); tc = time.perf_counter() - t0
            out["relerr_g0"] = relerr(g.g0, ref)
    elif what == "conductivity":
        M = kw["cond_ll"]
        rng = np.random.default_rng(1)
        nloop = kw.get("nloop", 1)
        mu = np.asfortranarray((rng.normal(size=(18, 18, M, M, nloop)) + 1j * rng.normal(size=(18, 18, M, M, nloop))) * 1e-3)
        rec.mu_nm_stochastic = mu
        c = Conductivity(rec)
        t = timed(c.calculate_conductivity_tensor, reps=2)
        out["M"] = M
        if cpu:
            t0 = time.perf_counter(); ri, _ = O.conductivity_integrand(mu, c.ene, EMIN, EMAX, True); tc = time.perf_counter() - t0
            ok = ~np.isnan(ri)
            out["relerr_integrand"] = relerr(c.integrand[ok], ri[ok])
    out.update({"gpu_seconds": t, "launches_per_call": (rec.launch_count - l0) // 4 if what != "conductivity" else (rec.launch_count - l0) // 3})
    if cpu:
        out.update({"cpu_seconds": tc, "cpu_threads": O.lib().orc_get_max_threads(), "speedup": tc / t})
    print(json.dumps(out), flush=True)
    rec.close()


def main_post():
    lat = S.sphere_cluster("bcc", 80.0)
    run_post("1 bulk bccFe run_dos block_green", lat, S.make_hamiltonian(lat, seed=20260101), 21, "block_green")
    run_post("1 bulk bccFe chebyshev_green lld=100", lat, S.make_hamiltonian(lat, seed=20260101), 100, "chebyshev_green")
    run_post("1 bulk bccFe sgreen (scalar)", lat, S.make_hamiltonian(lat, seed=20260101, spin_orbit=False), 21, "sgreen")
    lat = S.sphere_cluster("fcc", 100.0, ntype=7, type_rule="layer")
    lat.irec = np.array([1, 2, 3, 14, 15, 20], dtype=np.int32)
    run_post("2 surface fcc 6 units block_green", lat, S.make_hamiltonian(lat, seed=20260102), 21, "block_green")
    lat = S.periodic_bcc(4, 4, 3)
    ham = S.make_hamiltonian(lat, seed=20260104, velocity=True)
    run_post("4 conductivity integrand cond_ll=100", lat, ham, 21, "conductivity", cond_ll=100, cond_calctype="per_type", atlist=[1])
    run_post("4 conductivity integrand cond_ll=300", lat, ham, 21, "conductivity", cond_ll=300, cond_calctype="per_type", atlist=[1])


def run_scf_step(name, lat, ham, lld=21, channels=2500):
    """what self%run_recursion + self%run_dos do every SCF iteration on the block path (self.f90:799-856):
    recur_b -> zsqr -> get_terminf -> bgreen on the full mesh, host arrays in / g0 out; fused GPU call vs CPU oracle."""
    rec = Recursion(ham, lat, Control(lld=lld), Energy(EMIN, EMAX, channels_ldos=channels, fermi=0.0))
    g = Green(rec)
    # GPU timings first: the oracle's OpenMP threads keep spinning after a parallel region and steal the launching thread's core
    t = timed(g.recur_b_green, reps=15)
    g0_gpu = g.g0
    a_b_gpu = rec.a_b.copy()
    # the same step with g0 left on the device and the `bands` consumers (Fermi level, moments, band energy) run there
    from rslmtoasa_b200.bands import Bands
    nu = len(lat.irec)

    def with_bands():
        g.recur_b_green(download_g0=False)
        rec.en.fermi = 0.0
        b = Bands(g, qqv=6.0 * nu)
        b.calculate_fermi(); b.calculate_magnetic_moments(); b.calculate_moments(); b.calculate_band_energy()
        return b
    tb = timed(with_bands, reps=15)
    b = with_bands()
    orc = O.Oracle(lat, ham)
    t0 = time.perf_counter()
    a_b, b2_b = orc.lanczos_block(lat.irec, lld)
    t1 = time.perf_counter()
    ref = O.block_green(a_b, orc.zsqr(b2_b), g.ene)
    tc = time.perf_counter() - t0
    ok = np.isfinite(ref) & np.isfinite(g0_gpu)
    t2 = time.perf_counter()
    dtot = O.bands_dos(ref)[0]
    ef, nv1, e1, _ = O.bands_fermi(dtot, rec.en.edel, rec.en.energy_min, 6.0 * nu, 0.0, rec.en.ik1)
    occ, _ = O.bands_moments(ref, rec.en.channels_ldos, b.mom, g.ene, rec.en.edel, ef, nv1, e1)
    O.bands_magnetic_moments(ref, g.ene, rec.en.edel, ef, nv1, e1)
    tcb = time.perf_counter() - t2
    out = {"config": name, "what": "scf_step: recur_b + zsqr + terminator + bgreen (fused, device-resident coefficients)",
           "gpu_seconds_with_bands_g0_resident": tb, "cpu_bands_seconds": tcb, "err_fermi": abs(rec.en.fermi - ef),
           "err_charges": float(np.abs(b.occ - occ).max()),
           "kk": lat.kk, "units": int(len(lat.irec)), "lld": lld, "nv": channels + 10, "hoh": bool(ham.hoh),
           "gpu_seconds": t, "cpu_seconds": tc, "cpu_recursion_seconds": t1 - t0, "cpu_threads": O.lib().orc_get_max_threads(),
           "speedup": tc / t, "speedup_whole_step": (tc + tcb) / tb, "relerr_a_b": relerr(a_b_gpu, a_b), "relerr_g0": relerr(g0_gpu[ok], ref[ok]),
           "d2h_bytes_g0": int(g0_gpu.nbytes)}
    print(json.dumps(out), flush=True)
    rec.close()


def main_scf():
    lat = S.sphere_cluster("bcc", 80.0)
    run_scf_step("1 bulk bccFe", lat, S.make_hamiltonian(lat, seed=20260101))
    run_scf_step("1 bulk bccFe hoh", lat, S.make_hamiltonian(lat, seed=20260101, hoh=True))
    lat = S.sphere_cluster("fcc", 100.0, ntype=7, type_rule="layer")
    lat.irec = np.array([1, 2, 3, 14, 15, 20], dtype=np.int32)
    run_scf_step("2 surface fcc 6 units", lat, S.make_hamiltonian(lat, seed=20260102))
    lat = S.sphere_cluster("bcc", 60.0, ntype=3, nmax=15, type_rule="b2")
    run_scf_step("3 impurity B2 nmax=15", lat, S.make_hamiltonian(lat, seed=20260103))


def main():
    quick = "--quick" in sys.argv
    # config 1: bulk bcc Fe, rc = 80 -> kk = 5984, 1 type, 1 unit, lld = 21, block Lanczos (hoh F/T) + Chebyshev lld=100
    lat = S.sphere_cluster("bcc", 80.0)
    run("1 bulk bccFe block", lat, S.make_hamiltonian(lat, seed=20260101), 21, "recur_b")
    run("1 bulk bccFe block hoh", lat, S.make_hamiltonian(lat, seed=20260101, hoh=True), 21, "recur_b")
    run("1 bulk bccFe chebyshev", lat, S.make_hamiltonian(lat, seed=20260101), 100, "chebyshev_recur")
    ham = S.make_hamiltonian(lat, seed=20260101, spin_orbit=False)
    run("1 bulk bccFe lanczos (scalar, nsp=1)", lat, ham, 21, "recur")
    # config 2: surface: fcc sphere r2 = 100 (16756 sites), 7 layer types, 19 slots, 6 units batched
    lat = S.sphere_cluster("fcc", 100.0, ntype=7, type_rule="layer")
    lat.irec = np.array([1, 2, 3, 14, 15, 20], dtype=np.int32)
    run("2 surface fcc 6 units block", lat, S.make_hamiltonian(lat, seed=20260102), 21, "recur_b")
    run("2 surface fcc 6 units chebyshev", lat, S.make_hamiltonian(lat, seed=20260102), 100, "chebyshev_recur", cpu=not quick)
    # config 3: impurity: B2 sphere r2 = 60 (3838 sites), 3 types, 15 site-indexed sites
    lat = S.sphere_cluster("bcc", 60.0, ntype=3, nmax=15, type_rule="b2")
    run("3 impurity B2 nmax=15 block", lat, S.make_hamiltonian(lat, seed=20260103), 21, "recur_b")
    run("3 impurity B2 nmax=15 block hoh", lat, S.make_hamiltonian(lat, seed=20260103, hoh=True), 21, "recur_b")
    # config 4: conductivity: bcc PBC 20^3 cells (SURVEY: kk = 8000 -> 10x20x20 cells x 2), Kubo moments
    lat = S.periodic_bcc(10, 20, 20)
    ham = S.make_hamiltonian(lat, seed=20260104, velocity=True)
    run("4 conductivity per_type cond_ll=50", lat, ham, 21, "kubo", cond_ll=50, cond_calctype="per_type", atlist=[1], cpu=not quick)
    if not quick:
        run("4 conductivity random_vec cond_ll=300", lat, ham, 21, "kubo", cond_ll=300, cond_calctype="random_vec",
            phases=S.random_phases(lat.kk, 1), cpu=False)


if __name__ == "__main__":
    if "--post" in sys.argv:
        main_post()
    elif "--scf" in sys.argv:
        main_scf()
    else:
        main()

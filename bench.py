#!/usr/bin/env python
"""bench.py -- recursion steps/s of the block-Chebyshev (KPM) hot path on B200, with roofline and CPU baseline.

Workload (BASELINE.json configs[4], the largest single-GPU configuration and the one the 1/2/4/8-GPU metric is
quoted on): synthetic 1M-site bcc cluster (PBC 100x100x50 cells x 2 atoms), 15 neighbour slots, full 18x18 complex
spin-orbit blocks, KPM random-phase block vectors, one vector per GPU (weak scaling: R = N vectors on N GPUs,
no halo, the vectors are independent; moments are all-reduced once at the end of a recursion).

One *step* = one `chebyshev_recur_ll` (reference recursion.f90:2495-2597) over the whole cluster for one 18-column
block vector: fused H~ psi1 gather-SpMV + three-term update + the two 18x18 reductions = 2 Chebyshev moments.

  value : steps/s with everything resident in HBM (device-timed with CUDA events on the library's stream)
  e2e   : steps/s through the C ABI with HOST buffers: set_lattice + set_hamiltonian + cheb_moments_random, i.e.
          uploads of nn/iz/H/phases, table builds, the whole recursion and the download of mu_n inside the timed
          region (+ the NCCL all-reduce of the moments when N > 1)
  roofline : the binding bound of this path on B200 is the FP64 tensor pipe; `roofline_hbm` is the
          same launch against the HBM roofline for reference
  cpu_baseline / --impl reference : the CPU oracle (a C/OpenMP restatement of the reference loops; the Fortran
          reference itself cannot be built in this image) on all host cores, on a bounded sample of the workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from rslmtoasa_b200 import synthetic as S  # noqa: E402

EMIN, EMAX = -2.0, 2.0
SEED_H, SEED_PH = 20260105, 20260105
FLOPS_PER_SITE_STEP = (15 + 2) * 46656.0        # (nnb + 2) complex 18^3 products per site-step, SURVEY.md 8(d)
FLOPS_PER_SITE_SPMV = 15 * 46656.0              # the gather-SpMV kernel's share (the two Gram products are k_gram_dmma)
BYTES_PER_SITE_STEP = 3 * 5184.0 + 4 * 15 + 4   # r psi1, r psi0, w psi2 + nn + iz, SURVEY.md 8(d)
FP64_TENSOR_PEAK_TFLOPS = 37.1                  # measured here: tools/fp64_peak.cu, profiles/r01_fp64_peak_microbench.txt


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler:
    """samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        busy = [x for x in sm if x > 0.5 * (mx or 1)] or sm
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def workload(cells):
    nx, ny, nz = cells
    lat = S.periodic_bcc(nx, ny, nz)
    if not os.environ.get("RSREC_BENCH_NO_POSITIONS"):   # lattice%cr: orders the work for L2 locality, results unchanged
        lat.cr = S.periodic_bcc_positions(nx, ny, nz)
    ham = S.make_hamiltonian(lat, seed=SEED_H, spin_orbit=True)
    return lat, ham


def config_dict(cells, n_gpus, note=None):
    nx, ny, nz = cells
    kk = 2 * nx * ny * nz
    name = ("config5: synthetic 1M-site bcc" if kk == 1_000_000 else f"config5-scaled: synthetic {kk}-site bcc")
    cfg = {"workload": f"{name} PBC {nx}x{ny}x{nz}x2, 15 slots, 18x18 complex SO blocks, block-Chebyshev KPM "
                       f"(chebyshev_recur_ll), 1 random-phase block vector per GPU",
           "sites": kk, "nnb": 15, "vectors_per_gpu": 1, "vectors_total": n_gpus, "parallelism": f"vectors x{n_gpus}",
           "cache": "inputs larger than L2 (3 x %.2f GB of block vectors per GPU vs 126 MB L2)" % (kk * 5184 / 1e9)}
    if note:
        cfg["note"] = note
    return cfg


def cpu_sample_cells(cells):
    """bounded CPU sample: 1/8 of the lattice (each dimension halved), same stencil and blocks."""
    return tuple(max(2, c // 2) for c in cells)


import contextlib


@contextlib.contextmanager
def stdout_to_stderr():
    """stdout carries exactly one JSON line: whatever NCCL prints while a communicator comes up (its version banner when
    NCCL_DEBUG is set) goes to stderr by pointing fd 1 at fd 2 for that moment."""
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    try:
        yield
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(saved)


def host_cores():
    """cores this process may run on (the box's host cores); torchrun exports OMP_NUM_THREADS=1, which must not
    decide how many threads the CPU arm uses."""
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_cpu(cells, steps, warmup, threads=None):
    """oracle timing on a bounded sample -> dict(value = steps/s in units of FULL-workload steps, cores, sample, seconds,
    sample_fraction, sample_ms_per_step = what one timed step on the sample actually took)."""
    from oracle import oracle as O
    full_kk = 2 * cells[0] * cells[1] * cells[2]
    sc = cpu_sample_cells(cells)
    lat, ham = workload(sc)
    O.lib().orc_set_threads(threads or host_cores())
    cores = O.lib().orc_get_max_threads()
    orc = O.Oracle(lat, ham)
    a, b = O.cheb_scale(EMIN, EMAX)
    ph = S.random_phases(lat.kk, 1, seed=SEED_PH)[:, 0]
    if warmup > 0:
        orc.cheb_time_steps(ph, warmup, a, b)
    sec = orc.cheb_time_steps(ph, steps, a, b)
    frac = lat.kk / full_kk
    sample = (f"{steps} chebyshev_recur_ll steps (after {warmup} warm-up) on a {lat.kk}-site sub-lattice ({sc[0]}x{sc[1]}x{sc[2]}x2 "
              f"cells, sample_fraction {frac:.4f} of the workload's sites, same stencil and blocks), value scaled by sites; "
              f"C/OpenMP oracle (port of the reference loops), {cores} threads")
    return {"value": frac * steps / sec, "cores": cores, "sample": sample, "seconds": sec, "sample_fraction": frac,
            "sample_ms_per_step": 1e3 * sec / steps}


# ---------------------------------------------------------------------------------------------------------------------
# BASELINE.json configs 1-4 at SURVEY.md 8(d) size: steps/s through the reference-facing API (host arrays in and out),
# binding-roofline fraction from the 8(d) flop counts, parity against the CPU oracle, CPU steps/s beside it.
P18 = 46656.0  # flops of one complex 18x18x18 product


def _active_site_steps(lat, units, nsteps):
    """sum over operator applications k = 1..nsteps of the number of sites that can be non-zero after k applications
    (breadth-first levels from each unit's start site over nn) -- the work the reference's izero/irlist bookkeeping
    (recursion.f90:1604-1636) and the library's active-region plan actually touch."""
    nn = np.asarray(lat.nn)
    kk = lat.kk
    total = 0
    for site in units:
        reach = np.zeros(kk + 1, dtype=bool)   # index 0 = "no neighbour"
        reach[site] = True
        for _ in range(nsteps):
            nb = reach[nn[:, 1:]].any(axis=1)
            reach[1:] |= nb
            total += int(reach[1:].sum())
    return total


def _relerr(x, r):
    return float(np.abs(np.asarray(x) - np.asarray(r)).max() / np.abs(np.asarray(r)).max())


def _best(fn, reps=3, warm=1):
    """best of `reps` host-timed calls after `warm` untimed ones (a millisecond-sized call needs tens of repetitions: the GPU
    has idled through the CPU baseline that ran just before and takes a few calls to come back to its clocks)"""
    for _ in range(warm):
        fn()
    best = 1e30
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        best = min(best, time.perf_counter() - t0)
    return best


def bench_configs(device, with_cpu=True):
    from rslmtoasa_b200 import Recursion, Control, Energy
    from oracle import oracle as O
    O.lib().orc_set_threads(host_cores())
    cores = O.lib().orc_get_max_threads()
    out = []

    def lanczos_case(name, lat, ham, lld, tol=1e-10):
        nnb = lat.ncols
        rec = Recursion(ham, lat, Control(lld=lld), Energy(EMIN, EMAX), device=device)
        small = lat.kk * len(lat.irec) < 20000
        t = _best(rec.recur_b, reps=25 if small else 5, warm=5 if small else 1)
        units, steps = len(lat.irec), (lld - 1) * len(lat.irec)
        fl = (nnb + 1 + 4) * P18                      # SpMV + A, then psi A, pmn^H pmn, pmn B^-1, psi B  
        act = _active_site_steps(lat, [int(x) for x in lat.irec], lld - 1)
        e = {"config": name, "what": f"recur_b (block Lanczos), lld={lld}", "sites": lat.kk, "units": units, "nnb": nnb,
             "ntype": lat.ntype, "nmax": lat.nmax, "steps": steps, "gpu_seconds": t, "steps_per_s": steps / t,
             "roofline": {"bound": "tensor (FP64 DMMA)", "flops_per_site_step": fl, "peak_tflops": FP64_TENSOR_PEAK_TFLOPS,
                          "frac": fl * act / t / 1e12 / FP64_TENSOR_PEAK_TFLOPS,
                          "achieved_tflops": fl * act / t / 1e12,
                          "active_site_steps": act, "full_site_steps": lat.kk * steps,
                          "frac_if_every_site_counted": fl * lat.kk * steps / t / 1e12 / FP64_TENSOR_PEAK_TFLOPS,
                          "note": "flops = 8(d) per-site figure x the site-steps the recursion has reached (breadth-first levels "
                                  "from the start sites: what the reference's izero mask and the library's plan compute); "
                                  "frac_if_every_site_counted charges all kk sites at every step and can exceed 1"},
             "tolerance": tol}
        if with_cpu:
            orc = O.Oracle(lat, ham)
            t0 = time.perf_counter(); ra, rb = orc.lanczos_block(lat.irec, lld); tc = time.perf_counter() - t0
            e["parity_relerr"] = max(_relerr(rec.a_b, ra), _relerr(rec.b2_b, rb))
            e["parity"] = {"a_b": _relerr(rec.a_b, ra), "b2_b": _relerr(rec.b2_b, rb), "against": "CPU oracle, same inputs, full size"}
            e["cpu"] = {"seconds": tc, "steps_per_s": steps / tc, "cores": cores, "kind": "port"}
        rec.close()
        out.append(e)

    # 1 bulk bccFe: bcc sphere r2 = 80 -> 5984 sites, 1 type, 1 unit, lld = 21
    lat = S.sphere_cluster("bcc", 80.0)
    lanczos_case("1 bulk bccFe", lat, S.make_hamiltonian(lat, seed=20260101), 21)
    # 2 surface: fcc sphere r2 = 100 -> 16756 sites, 7 layer types, 19 slots, 6 units in one batch
    lat = S.sphere_cluster("fcc", 100.0, ntype=7, type_rule="layer")
    lat.irec = np.array([1, 2, 3, 14, 15, 20], dtype=np.int32)
    lanczos_case("2 surface fcc(001)-typed, 6 units", lat, S.make_hamiltonian(lat, seed=20260102), 21)
    # 3 impurity: B2 sphere r2 = 60 -> 3838 sites, 3 types, 15 site-indexed (hall) sites
    lat = S.sphere_cluster("bcc", 60.0, ntype=3, nmax=15, type_rule="b2")
    lanczos_case("3 impurity B2, nmax=15", lat, S.make_hamiltonian(lat, seed=20260103), 21)
    # 4 conductivity: bcc PBC 8000 sites, cond_ll = 50, R = 8 random vectors (compute_moments_stochastic, full 18x18 blocks)
    lat = S.periodic_bcc(10, 20, 20)
    ham = S.make_hamiltonian(lat, seed=20260104, velocity=True)
    M, R = 50, 8
    ph = S.random_phases(lat.kk, R, seed=20260104)
    rec = Recursion(ham, lat, Control(lld=21, cond_ll=M, cond_calctype="random_vec"), Energy(EMIN, EMAX), device=device, phases=ph)
    t = _best(rec.compute_moments_stochastic, reps=2)
    nnb = lat.ncols
    apps = 3 * M                                            # M left + M right Chebyshev applications + M velocity applications (+1)
    fl_vec = (apps + 1) * nnb * P18 * lat.kk + M * M * P18 * lat.kk   # SpMV products + the M x M block contractions per vector
    e = {"config": "4 conductivity bcc PBC", "what": f"compute_moments_stochastic, cond_ll={M}, random_vec R={R}", "sites": lat.kk,
         "units": R, "nnb": nnb, "steps": apps * R, "gpu_seconds": t, "steps_per_s": apps * R / t,
         "roofline": {"bound": "tensor (FP64 DMMA)", "flops_total": fl_vec * R, "peak_tflops": FP64_TENSOR_PEAK_TFLOPS,
                      "frac": fl_vec * R / t / 1e12 / FP64_TENSOR_PEAK_TFLOPS, "achieved_tflops": fl_vec * R / t / 1e12,
                      "note": "steps = SpMV applications (3 per moment index); flops = applications x nnb x 46656 + cond_ll^2 x 46656 "
                              "per site and vector (8d Kubo row); every site is active (random start)"},
         "tolerance": 1e-9}
    if with_cpu:
        orc = O.Oracle(lat, ham)
        a, b = O.cheb_scale(EMIN, EMAX)
        msel = np.array([1, 2, M // 2, M], dtype=np.int32)
        t0 = time.perf_counter(); orc.kubo_moments_cols(M, a, b, msel[:0], phases=ph[:, :1]); t_chain = time.perf_counter() - t0
        t0 = time.perf_counter(); ref = orc.kubo_moments_cols(M, a, b, msel, phases=ph[:, :1]); t_sel = time.perf_counter() - t0
        got = rec.mu_nm_stochastic[:, :, :, msel - 1, 0]
        e["parity_relerr"] = _relerr(got, ref[..., 0])
        e["parity"] = {"mu_nm": e["parity_relerr"], "against": f"CPU oracle, same inputs, full size, vector 1, all n, left indices m = {msel.tolist()}"}
        est = R * (t_chain + max(t_sel - t_chain, 0.0) * M / len(msel))
        e["cpu"] = {"seconds_extrapolated": est, "steps_per_s": apps * R / est, "cores": cores, "kind": "port",
                    "sample": f"chains of 1 vector timed ({t_chain:.2f} s) + 4 of {M} left-index contractions timed "
                              f"({max(t_sel - t_chain, 0.0):.2f} s), extrapolated to {M} indices and {R} vectors"}
    rec.close()
    out.append(e)
    return out


def bench_strong(comm_info, device, rank, world, barrier, max_over_ranks):
    """Fixed-total unit-sharded work on N GPUs (the reference's MPI rank = unit shard, mpi.f90:32-58), the
    exchange inside the timed region and inside the library (NCCL on device-resident results):
      A  config 2 (surface, 16756 sites) x 24 recursion sites: recur_b on this rank's shard + all-gather of a_b/b2_b;
      C  config 1 (bulk, 5984 sites) x 32 pair start vectors (recur_b_ij);  D  config 3 (impurity, 3838 sites) x 16 sites;
      B  config 4 (8000 sites, cond_ll = 50) x 64 random vectors: Kubo moments + Gamma contraction + all-reduce of the integrand.
    t1 = the same total on ONE GPU, measured in this very run (every rank does it on its own GPU at the same time, max taken),
    so speedup and efficiency come from one box and one build."""
    from rslmtoasa_b200 import Recursion, Control, Energy, Conductivity
    recs = []

    def timed(fn):
        barrier()
        t0 = time.perf_counter()
        fn()
        return max_over_ranks(time.perf_counter() - t0)

    def phases_of(rec, fn):
        rec.phase_timing(True); rec.host_phase_read()
        barrier(); fn()
        ph = rec.host_phase_read(); rec.phase_timing(False)
        return {k: max_over_ranks(v) for k, v in ph.items()}

    # --- A
    lat = S.sphere_cluster("fcc", 100.0, ntype=7, type_rule="layer")
    lat.irec = np.arange(1, 25, dtype=np.int32)
    ham = S.make_hamiltonian(lat, seed=20260102)
    rec = Recursion(ham, lat, Control(lld=21), Energy(EMIN, EMAX), device=device)
    rec.recur_b()                                    # warm-up, all 24 units on this GPU
    t1 = min(timed(rec.recur_b) for _ in range(3))
    ref_a = rec.a_b.copy()
    ph1 = phases_of(rec, rec.recur_b)
    with stdout_to_stderr():
        rec.comm_init_torch()
    rec.recur_b_sharded()
    tn = min(timed(rec.recur_b_sharded) for _ in range(3))
    same = float(np.abs(rec.a_b - ref_a).max() / np.abs(ref_a).max())
    phn = phases_of(rec, rec.recur_b_sharded)
    recs.append({"workload": "A: config 2 surface (16756 sites, 7 types, 19 slots) x 24 recursion sites, recur_b lld=21, "
                             "all-gather of a_b/b2_b on the device (rsrec_lanczos_block_sharded)",
                 "units": 24, "n_gpus": world, "t1_s": t1, "tN_s": tn, "speedup": t1 / tn, "efficiency": t1 / tn / world,
                 "relerr_vs_1gpu": same, "phases_1gpu_s": ph1, "phases_Ngpu_s": phn,
                 "phases_note": "host wall-clock per stage with a stream sync at every stage boundary (max over ranks)"})
    rec.close()
    # --- C: config 1 (bulk bccFe, 5984 sites) x 8 pairs = 32 pair start vectors (recur_b_ij, the exchange path)
    lat = S.sphere_cluster("bcc", 80.0)
    ham = S.make_hamiltonian(lat, seed=20260101)
    pairs = np.array([[1, j] for j in range(2, 10)], dtype=np.int32)
    rec = Recursion(ham, lat, Control(lld=21), Energy(EMIN, EMAX), device=device, ijpair=pairs)
    rec.recur_b_ij()
    t1 = min(timed(rec.recur_b_ij) for _ in range(3))
    ref_a = rec.a_b.copy()
    ph1 = phases_of(rec, rec.recur_b_ij)
    with stdout_to_stderr():
        rec.comm_init_torch()
    rec.recur_b_ij_sharded()
    tn = min(timed(rec.recur_b_ij_sharded) for _ in range(3))
    same = float(np.abs(rec.a_b - ref_a).max() / np.abs(ref_a).max())
    phn = phases_of(rec, rec.recur_b_ij_sharded)
    recs.append({"workload": "C: config 1 bulk bccFe (5984 sites) x 8 pairs = 32 pair start vectors, recur_b_ij lld=21, "
                             "all-gather of a_b/b2_b on the device",
                 "units": 32, "n_gpus": world, "t1_s": t1, "tN_s": tn, "speedup": t1 / tn, "efficiency": t1 / tn / world,
                 "relerr_vs_1gpu": same, "phases_1gpu_s": ph1, "phases_Ngpu_s": phn})
    rec.close()
    # --- D: config 3 (impurity B2, 3838 sites, 15 site-indexed sites) x 16 recursion sites (the local region + 1)
    lat = S.sphere_cluster("bcc", 60.0, ntype=3, nmax=15, type_rule="b2")
    lat.irec = np.arange(1, 17, dtype=np.int32)
    ham = S.make_hamiltonian(lat, seed=20260103)
    rec = Recursion(ham, lat, Control(lld=21), Energy(EMIN, EMAX), device=device)
    rec.recur_b()
    t1 = min(timed(rec.recur_b) for _ in range(3))
    ref_a = rec.a_b.copy()
    ph1 = phases_of(rec, rec.recur_b)
    with stdout_to_stderr():
        rec.comm_init_torch()
    rec.recur_b_sharded()
    tn = min(timed(rec.recur_b_sharded) for _ in range(3))
    same = float(np.abs(rec.a_b - ref_a).max() / np.abs(ref_a).max())
    phn = phases_of(rec, rec.recur_b_sharded)
    recs.append({"workload": "D: config 3 impurity B2 (3838 sites, nmax=15) x 16 recursion sites, recur_b lld=21, "
                             "all-gather of a_b/b2_b on the device",
                 "units": 16, "n_gpus": world, "t1_s": t1, "tN_s": tn, "speedup": t1 / tn, "efficiency": t1 / tn / world,
                 "relerr_vs_1gpu": same, "phases_1gpu_s": ph1, "phases_Ngpu_s": phn})
    rec.close()
    # --- B
    lat = S.periodic_bcc(10, 20, 20)
    ham = S.make_hamiltonian(lat, seed=20260104, velocity=True)
    R, M = 64, 50
    ph = S.random_phases(lat.kk, R, seed=20260104)
    rec = Recursion(ham, lat, Control(lld=21, cond_ll=M, cond_calctype="random_vec"), Energy(EMIN, EMAX, channels_ldos=1000),
                    device=device, phases=ph)
    con = Conductivity(rec)
    rec.phases = ph[:, :2]; con.compute_conductivity()      # warm-up
    rec.phases = ph
    t1 = min(timed(con.compute_conductivity) for _ in range(2))
    ref_i = con.integrand.copy()
    ph1 = phases_of(rec, con.compute_conductivity)
    with stdout_to_stderr():
        rec.comm_init_torch()
    tn = min(timed(con.compute_conductivity) for _ in range(2))
    err = float(np.abs(con.integrand - ref_i).max() / np.abs(ref_i).max())
    phn = phases_of(rec, con.compute_conductivity)
    recs.append({"workload": "B: config 4 conductivity (8000 sites, cond_ll=50) x 64 random vectors, Kubo moments + Gamma "
                             "contraction, integrand all-reduced on the device (rsrec_kubo_conductivity)",
                 "units": R, "n_gpus": world, "t1_s": t1, "tN_s": tn, "speedup": t1 / tn, "efficiency": t1 / tn / world,
                 "relerr_vs_1gpu": err, "phases_1gpu_s": ph1, "phases_Ngpu_s": phn,
                 "phases_note": "recursion = moment chains + diagonal contractions; exchange = Gamma contraction + all-reduce + download"})
    rec.close()
    return {"nccl_version": comm_info[2], "records": recs}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cells", type=int, nargs=3, default=[100, 100, 50], help="bcc cells nx ny nz (x2 atoms)")
    ap.add_argument("--e2e-lld", type=int, default=249, help="recursion depth of the e2e call (500 moments)")
    ap.add_argument("--family", type=int, default=1, help="0 = SIMT kernels, 1 = DMMA pipeline")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-steps", type=int, default=20)
    ap.add_argument("--no-configs", action="store_true", help="skip the configs 1-4 records (N = 1)")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling records (N > 1)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cells = tuple(args.cells)
    kk = 2 * cells[0] * cells[1] * cells[2]

    if args.impl == "reference":
        # the reference's own CPU path for this metric: not buildable here (Fortran), so the oracle port stands in
        if rank != 0:
            return
        c = run_cpu(cells, args.steps, args.warmup)
        val = c["value"]
        # ms_per_step is what one TIMED step took (a step of the sample); value is scaled to full-workload steps by the
        # sample fraction, so value = sample_fraction * 1000 / ms_per_step
        line = {"impl": "reference", "metric": "recursion_steps_per_s", "value": val, "unit": "steps/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": c["sample_ms_per_step"], "sample_fraction": c["sample_fraction"],
                "ms_per_full_workload_step": 1e3 / val, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": config_dict(cells, args.gpus),
                "cpu_baseline": {"value": val, "unit": "steps/s", "cores": c["cores"], "kind": "port", "sample": c["sample"],
                                 "seconds": c["seconds"], "sample_fraction": c["sample_fraction"]},
                "e2e": {"value": val, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0,
                "note": "CPU arm does not scale with --gpus: it is the same host-core run at every N (the reference shards units over "
                        "MPI ranks on the same cores); compare N-GPU values with N x nothing -- the ratio at N is value_N / this"}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    from rslmtoasa_b200 import Recursion, Control, Energy
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        with stdout_to_stderr():
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    lat, ham = workload(cells)
    lld = args.warmup + 2 * args.steps
    rec = Recursion(ham, lat, Control(lld=lld), Energy(EMIN, EMAX), device=local_rank, rank=rank, numprocs=world)
    rec.set_kernel_family(args.family)
    # pinned host staging for the start phases (one random-phase vector per GPU, sharded with the reference rule)
    ph_all = S.random_phases(kk, world, seed=SEED_PH)
    ph_pin = torch.empty((kk,), dtype=torch.float64).pin_memory()
    ph_pin.numpy()[:] = ph_all[:, rank]
    ph_local = ph_pin.numpy().reshape(kk, 1, order="F")

    # ---------------- device-resident steps: `value` ----------------
    stream = torch.cuda.ExternalStream(rec.stream, device=torch.device("cuda", local_rank))
    rec.cheb_begin_random(ph_local, lld)
    rec.cheb_run_steps(args.warmup)
    rec.synchronize()
    torch.cuda.synchronize()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = rec.launch_count
    with torch.cuda.stream(stream):
        e0.record(stream)
        rec.cheb_run_steps(args.steps)
        e1.record(stream)
    rec.synchronize()
    torch.cuda.synchronize()
    barrier()
    launches = rec.launch_count - l0
    ms = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop()
    # per-kernel timing of the dominant kernel (gather-SpMV) over another K steps, CUDA events on the same stream
    rec.profile(True)
    rec.cheb_run_steps(args.steps)
    k_ms, k_n = rec.profile_read()
    rec.profile(False)
    mu_dev = rec.cheb_end()
    assert np.isfinite(mu_dev).all()
    k_avg_ms = k_ms / max(k_n, 1)
    ms_per_step = ms / args.steps
    value = world * args.steps / (ms / 1e3)

    peaks, peak_src = measured_peaks()
    flops_launch = FLOPS_PER_SITE_SPMV * kk            # one launch = the 15-slot gather-SpMV of one vector, one step
    bytes_launch = BYTES_PER_SITE_STEP * kk
    traffic, traffic_source = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tj = json.load(f)
            if tj.get("sites") == kk and tj.get("family") == args.family:
                traffic = tj.get("dram_bytes_per_launch")
                traffic_source = "static ncu capture (%s, commit %s): not re-measured in this run" % (tj.get("source", "profiles/"), tj.get("commit", "?"))
    except Exception:
        pass
    ach_tf = flops_launch / (k_avg_ms / 1e3) / 1e12
    roofline = {"bound": "tensor", "achieved": ach_tf, "peak": FP64_TENSOR_PEAK_TFLOPS, "unit": "TFLOP/s",
                "frac": ach_tf / FP64_TENSOR_PEAK_TFLOPS, "traffic": traffic, "traffic_source": traffic_source,
                "kernel": "k_apply_dmma<EPI_CHEB_NOGRAM>" if args.family == 1 else "k_apply_simt",
                "kernel_ms": k_avg_ms, "kernel_share_of_step": k_avg_ms / ms_per_step,
                "step_achieved": FLOPS_PER_SITE_STEP * kk / (ms_per_step / 1e3) / 1e12,
                "step_frac": FLOPS_PER_SITE_STEP * kk / (ms_per_step / 1e3) / 1e12 / FP64_TENSOR_PEAK_TFLOPS,
                "peak_source": "FP64 DMMA m8n8k4 peak measured on this pool's B200 with tools/fp64_peak.cu "
                               "(profiles/r01_fp64_peak_microbench.txt); MEASURED_PEAKS.json holds no FP64 figure; "
                               "algorithmic flops of this kernel = 15 x 46656 per site (the nnb SpMV products); step_* = all 17 "
                               "products of SURVEY.md 8d over the whole step (SpMV + Gram + reduce kernels)"}
    ach_gbs = bytes_launch / (k_avg_ms / 1e3) / 1e9
    roofline_hbm = {"bound": "hbm", "achieved": ach_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": ach_gbs / peaks["hbm_gbs"], "traffic": traffic,
                    "note": f"peak {peak_src}; this variant is FP64-bound (51 flop/B vs ~5.7 flop/B machine balance), "
                            "so the HBM fraction is capped near 0.11"}

    # ---------------- e2e through the C ABI with host buffers ----------------
    if world > 1:
        with stdout_to_stderr():
            rec.comm_init_torch()          # NCCL communicator inside librsrec.so (the id travels over the torch group)
    e2e = None
    if not args.no_e2e:
        e2e_lld = args.e2e_lld
        rec.control.lld = e2e_lld
        # two complete passes, the faster one reported (a single 7 s shot through the host API showed 6.7 ... 7.5 s on one box)
        passes = []
        for _ in range(2):
            barrier()
            torch.cuda.synchronize()
            h0, d0 = rec.h2d_bytes, rec.d2h_bytes
            t0 = time.perf_counter()
            rec.upload()                       # set_lattice + set_hamiltonian (tables rebuilt and re-uploaded)
            # this rank's vector (its block-rule shard of the `world` columns): phases H2D, lld steps, moments summed over
            # vectors and all-reduced over NVLink on the device (rsrec_cheb_moments_random_sum), one D2H of the summed moments
            mu_sum = rec.chebyshev_recur_random_sum(ph_local, sharded=False)   # ph_local: pinned host memory
            t1 = time.perf_counter()
            assert np.isfinite(mu_sum).all()
            passes.append(max_over_ranks(t1 - t0))
            h2d_step, d2h_step = (rec.h2d_bytes - h0) / e2e_lld, (rec.d2h_bytes - d0) / e2e_lld
        e2e_s = min(passes)
        e2e = {"value": world * e2e_lld / e2e_s, "unit": "steps/s",
               "h2d_bytes_per_step": h2d_step, "d2h_bytes_per_step": d2h_step,
               "steps": e2e_lld, "seconds": e2e_s, "passes_s": passes,
               "what": "rsrec_set_lattice + rsrec_set_hamiltonian + rsrec_cheb_moments_random_sum(lld=%d) from host arrays, "
                       "faster of two complete passes"
                       "%s" % (e2e_lld, " (moments all-reduced by the library's NCCL communicator, %d ranks)" % world if world > 1 else "")}

    # ---------------- strong scaling of unit-sharded work (N > 1): fixed total, sharded by the reference's block rule ----
    strong = None
    if world > 1 and not args.no_strong:
        strong = bench_strong(rec.comm_info(), local_rank, rank, world, barrier, max_over_ranks)
    configs = None
    if rank == 0 and world == 1 and not args.no_configs:
        configs = bench_configs(local_rank, with_cpu=not args.no_cpu)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        c = run_cpu(cells, args.cpu_steps, 2)
        cpu = {"value": c["value"], "unit": "steps/s", "cores": c["cores"], "kind": "port", "sample": c["sample"],
               "seconds": c["seconds"], "sample_fraction": c["sample_fraction"], "sample_ms_per_step": c["sample_ms_per_step"]}

    if rank == 0:
        line = {"metric": "recursion_steps_per_s", "value": value, "unit": "steps/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": config_dict(cells, world), "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
                "roofline": roofline, "roofline_hbm": roofline_hbm, "cpu_baseline": cpu,
                "kernel_family": "dmma" if args.family == 1 else "simt"}
        if configs is not None:
            line["configs"] = configs
        if strong is not None:
            line["strong"] = strong
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

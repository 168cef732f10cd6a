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


def run_cpu(cells, steps, warmup, threads=None):
    """oracle timing: returns (steps/s in units of FULL-workload steps, cores, sample description, seconds)."""
    from oracle import oracle as O
    full_kk = 2 * cells[0] * cells[1] * cells[2]
    sc = cpu_sample_cells(cells)
    lat, ham = workload(sc)
    if threads:
        O.lib().orc_set_threads(threads)
    cores = O.lib().orc_get_max_threads()
    orc = O.Oracle(lat, ham)
    a, b = O.cheb_scale(EMIN, EMAX)
    ph = S.random_phases(lat.kk, 1, seed=SEED_PH)[:, 0]
    if warmup > 0:
        orc.cheb_time_steps(ph, warmup, a, b)
    sec = orc.cheb_time_steps(ph, steps, a, b)
    frac = lat.kk / full_kk
    sample = (f"{steps} chebyshev_recur_ll steps on a {lat.kk}-site sub-lattice ({sc[0]}x{sc[1]}x{sc[2]}x2, "
              f"{frac:.4f} of the workload's sites), scaled by sites; C/OpenMP oracle, {cores} threads")
    return frac * steps / sec, cores, sample, sec


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
    ap.add_argument("--cpu-steps", type=int, default=8)
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
        val, cores, sample, sec = run_cpu(cells, args.steps, args.warmup)
        line = {"impl": "reference", "metric": "recursion_steps_per_s", "value": val, "unit": "steps/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1e3 / val, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": config_dict(cells, args.gpus),
                "cpu_baseline": {"value": val, "unit": "steps/s", "cores": cores, "kind": "port", "sample": sample},
                "e2e": {"value": val, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    from rslmtoasa_b200 import Recursion, Control, Energy
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # stdout carries exactly one JSON line: whatever NCCL prints while the communicator comes up (its "NCCL version .."
        # banner when NCCL_DEBUG is set in the environment) is sent to stderr by pointing fd 1 at fd 2 for that moment
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)

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
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tj = json.load(f)
            if tj.get("sites") == kk and tj.get("family") == args.family:
                traffic = tj.get("dram_bytes_per_launch")
    except Exception:
        pass
    ach_tf = flops_launch / (k_avg_ms / 1e3) / 1e12
    roofline = {"bound": "tensor", "achieved": ach_tf, "peak": FP64_TENSOR_PEAK_TFLOPS, "unit": "TFLOP/s",
                "frac": ach_tf / FP64_TENSOR_PEAK_TFLOPS, "traffic": traffic,
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
    e2e = None
    if not args.no_e2e:
        e2e_lld = args.e2e_lld
        rec.control.lld = e2e_lld
        rec.phases = ph_local
        mu_sum = torch.zeros((18, 18, 2 * e2e_lld + 2), dtype=torch.complex128, device="cuda")
        barrier()
        torch.cuda.synchronize()
        h0, d0 = rec.h2d_bytes, rec.d2h_bytes
        t0 = time.perf_counter()
        rec.upload()                       # set_lattice + set_hamiltonian (tables rebuilt and re-uploaded)
        rec.numprocs, rec.rank = 1, 0      # this rank's shard is exactly its one vector
        rec.chebyshev_recur_random()       # uploads phases, runs lld steps, downloads mu_n
        if world > 1:                      # the one real exchange of the path: sum of the moments over vectors
            mu_sum.copy_(torch.from_numpy(np.ascontiguousarray(rec.mu_n[..., 0])))
            dist.all_reduce(mu_sum)
            torch.cuda.synchronize()
        t1 = time.perf_counter()
        e2e_s = max_over_ranks(t1 - t0)
        e2e = {"value": world * e2e_lld / e2e_s, "unit": "steps/s",
               "h2d_bytes_per_step": (rec.h2d_bytes - h0) / e2e_lld, "d2h_bytes_per_step": (rec.d2h_bytes - d0) / e2e_lld,
               "steps": e2e_lld, "seconds": e2e_s,
               "what": "rsrec_set_lattice + rsrec_set_hamiltonian + rsrec_cheb_moments_random(lld=%d) from host arrays"
                       % e2e_lld}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        val, cores, sample, sec = run_cpu(cells, args.cpu_steps, 1)
        cpu = {"value": val, "unit": "steps/s", "cores": cores, "kind": "port", "sample": sample, "seconds": sec}

    if rank == 0:
        line = {"metric": "recursion_steps_per_s", "value": value, "unit": "steps/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": config_dict(cells, world), "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
                "roofline": roofline, "roofline_hbm": roofline_hbm, "cpu_baseline": cpu,
                "kernel_family": "dmma" if args.family == 1 else "simt"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

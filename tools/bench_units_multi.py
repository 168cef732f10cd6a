#!/usr/bin/env python
"""Multi-GPU runs of the unit-sharded configurations: recursion sites / random vectors are split over
the ranks with get_mpi_variables' block rule, every rank holds the full {H, nn} replica, one all-gather (site-resolved
coefficients) or all-reduce (stochastic moments) at the end.  Launch with torchrun; rank 0 prints one JSON line per case.

  config 2x : surface fcc (16 756 sites), 24 recursion sites (4 replicas of the 6 layer sites), recur_b lld = 21
  config 4  : conductivity bcc PBC 8000 sites, 8 random vectors, Kubo-Bastin moments cond_ll = 48
Timing: CUDA-synchronised wall clock of the whole call on every rank (uploads, recursion, download, collective),
barrier on both sides, max over ranks.
"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rslmtoasa_b200 import Recursion, Control, Energy, synthetic as S, parallel as P  # noqa: E402


def timed(fn, dev, reps=3):
    best = 1e30
    for _ in range(reps + 1):
        dist.barrier(); torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        out = fn()
        torch.cuda.synchronize(dev)
        t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        best = min(best, float(t.item()))
    return best, out


def main():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    saved_fd = os.dup(1); os.dup2(2, 1)        # NCCL's start-up banner goes to stderr, stdout stays JSON lines only
    try:
        dist.init_process_group("nccl", device_id=dev)
        dist.barrier(); torch.cuda.synchronize(dev)
    finally:
        sys.stdout.flush(); os.dup2(saved_fd, 1); os.close(saved_fd)
    # ---- config 2 (replicated unit batch) ----
    lat = S.sphere_cluster("fcc", 100.0, ntype=7, type_rule="layer")
    base = [1, 2, 3, 14, 15, 20]
    lat.irec = np.array([s + 40 * r for r in range(4) for s in base], dtype=np.int32)
    ham = S.make_hamiltonian(lat, seed=20260102)
    rec = Recursion(ham, lat, Control(lld=21), Energy(-2.0, 2.0), device=local, rank=rank, numprocs=world)

    def run2():
        rec.recur_b()
        return P.allgather_units(rec.a_b, len(lat.irec), dev)
    t, a_all = timed(run2, dev)
    if rank == 0:
        print(json.dumps({"config": "2 surface fcc, 24 recursion sites, recur_b lld=21", "n_gpus": world, "units_total": 24,
                          "units_per_gpu": len(lat.irec) // world, "seconds": t, "steps_per_s": 24 * 20 / t,
                          "checksum": float(np.abs(a_all).sum())}), flush=True)
    # ---- the whole SCF iteration on the same 24 units: recursion -> terminator -> Green function (g0 stays on each GPU)
    #      -> dtot all-reduced -> Fermi level -> charges / moments per unit, gathered ----
    from rslmtoasa_b200 import Green
    from rslmtoasa_b200.bands import Bands
    rec.en.channels_ldos, rec.en.fermi, rec.en.ene = 2500, 0.0, None
    g = Green(rec)

    def run_scf():
        g.recur_b_green(download_g0=False)
        rec.en.fermi = 0.0
        b = Bands(g, qqv=6.0 * len(lat.irec), device=dev)
        b.calculate_fermi(); b.calculate_magnetic_moments(); b.calculate_moments(); b.calculate_band_energy()
        return b.en.fermi, b.eband, b.gather(b.occ)
    t, (ef, eb, occ) = timed(run_scf, dev)
    if rank == 0:
        print(json.dumps({"config": "2 surface fcc, 24 recursion sites, whole SCF iteration (recur_b + green nv=2510 + bands)",
                          "n_gpus": world, "units_per_gpu": len(lat.irec) // world, "seconds": t, "fermi": ef, "eband": eb,
                          "checksum_occ": float(np.abs(occ).sum()), "occ_shape": list(occ.shape)}), flush=True)
    rec.close()
    # ---- config 4 ----
    lat = S.periodic_bcc(10, 20, 20)
    lat.cr = S.periodic_bcc_positions(10, 20, 20)
    ham = S.make_hamiltonian(lat, seed=20260104, velocity=True)
    nvec, M = 8, 48
    ph = S.random_phases(lat.kk, nvec)
    lo, hi = P.shard_range(nvec, rank, world)
    rec = Recursion(ham, lat, Control(cond_ll=M, cond_calctype="random_vec"), Energy(-2.0, 2.0), device=local,
                    phases=np.asfortranarray(ph[:, lo:hi]))

    def run4():
        if hi > lo:
            rec.compute_moments_stochastic()
            loc = rec.mu_nm_stochastic.sum(axis=-1)
        else:
            loc = np.zeros((18, 18, M, M), complex)
        return P.allreduce_sum(np.asfortranarray(loc), dev)
    t, mu = timed(run4, dev, reps=2)
    if rank == 0:
        print(json.dumps({"config": "4 conductivity bcc PBC 8000 sites, 8 random vectors, cond_ll=48", "n_gpus": world,
                          "vectors_per_gpu": nvec // world, "seconds": t, "spmv_equiv_per_s": 3 * M * nvec / t,
                          "checksum": float(np.abs(mu).sum())}), flush=True)
    rec.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Where a recur_b call spends its time on configs 1-3: host wall time of the ABI call, device time per g_timer phase
(CUDA events), and host-side stage times.  Usage: python tools/time_recur_b.py [reps]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rslmtoasa_b200 import Recursion, Control, Energy, Green, synthetic as S  # noqa: E402


def run(name, lat, ham, lld=21, reps=10):
    rec = Recursion(ham, lat, Control(lld=lld), Energy(-2.0, 2.0, channels_ldos=2500, fermi=0.0))
    rec.recur_b()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); rec.recur_b(); ts.append(time.perf_counter() - t0)
    l0 = rec.launch_count
    rec.phase_timing(True); rec.host_phase_read()
    t0 = time.perf_counter(); rec.recur_b(); tp = time.perf_counter() - t0
    ph = rec.phase_read(); hp = rec.host_phase_read()
    rec.phase_timing(False)
    g = Green(rec)
    g.recur_b_green(download_g0=False)
    tg = []
    for _ in range(reps):
        t0 = time.perf_counter(); g.recur_b_green(download_g0=False); tg.append(time.perf_counter() - t0)
    out = {"config": name, "kk": lat.kk, "units": int(len(lat.irec)), "lld": lld, "recur_b_wall_ms_min": 1e3 * min(ts),
           "recur_b_wall_ms_median": 1e3 * float(np.median(ts)), "launches": rec.launch_count - l0,
           "device_ms_per_phase": {k: round(v[0], 4) for k, v in ph.items()}, "device_ms_phases_sum": round(sum(v[0] for v in ph.values()), 4),
           "wall_ms_with_phase_syncs": 1e3 * tp, "host_stage_ms": {k: round(1e3 * v, 4) for k, v in hp.items()},
           "recur_b_green_g0_resident_ms_min": 1e3 * min(tg)}
    print(json.dumps(out), flush=True)
    rec.close()


if __name__ == "__main__":
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    lat = S.sphere_cluster("bcc", 80.0)
    run("1 bulk bccFe (full blocks)", lat, S.make_hamiltonian(lat, seed=20260101), reps=reps)
    run("1 bulk bccFe collinear (spin-diagonal hoppings)", lat, S.make_hamiltonian(lat, seed=20260101, spin_orbit=False), reps=reps)
    lat = S.sphere_cluster("fcc", 100.0, ntype=7, type_rule="layer")
    lat.irec = np.array([1, 2, 3, 14, 15, 20], dtype=np.int32)
    run("2 surface fcc 6 units", lat, S.make_hamiltonian(lat, seed=20260102), reps=max(3, reps // 3))
    lat = S.sphere_cluster("bcc", 60.0, ntype=3, nmax=15, type_rule="b2")
    run("3 impurity B2 nmax=15", lat, S.make_hamiltonian(lat, seed=20260103), reps=reps)

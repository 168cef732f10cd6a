#!/usr/bin/env python
"""chebyshev_recur on configs 1-3 (lld = 100) with the reductions inside the SpMV kernel (rsrec_set_fusion cheb=1) and as a
separate Gram kernel (cheb=0): host API wall time, best of N."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rslmtoasa_b200 import Recursion, Control, Energy, synthetic as S  # noqa: E402


def run(name, lat, ham, lld=100, reps=6):
    out = {"config": name, "kk": lat.kk, "units": int(len(lat.irec)), "lld": lld}
    for fused in (0, 1):
        rec = Recursion(ham, lat, Control(lld=lld), Energy(-2.0, 2.0))
        rec.set_fusion(cheb=fused)
        rec.chebyshev_recur()
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter(); rec.chebyshev_recur(); ts.append(time.perf_counter() - t0)
        out["fused_cheb=%d_ms" % fused] = 1e3 * min(ts)
        rec.close()
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    lat = S.sphere_cluster("bcc", 80.0)
    run("1 bulk bccFe", lat, S.make_hamiltonian(lat, seed=20260101))
    run("1 bulk bccFe collinear", lat, S.make_hamiltonian(lat, seed=20260101, spin_orbit=False))
    lat = S.sphere_cluster("fcc", 100.0, ntype=7, type_rule="layer")
    lat.irec = np.array([1, 2, 3, 14, 15, 20], dtype=np.int32)
    run("2 surface fcc 6 units", lat, S.make_hamiltonian(lat, seed=20260102), reps=3)
    lat = S.sphere_cluster("bcc", 60.0, ntype=3, nmax=15, type_rule="b2")
    run("3 impurity B2 nmax=15", lat, S.make_hamiltonian(lat, seed=20260103))
    lat = S.periodic_bcc(10, 20, 20)
    run("4-size bcc PBC 8000 sites, site start", lat, S.make_hamiltonian(lat, seed=20260104))

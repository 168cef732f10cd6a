#!/usr/bin/env python
"""Repeats the Kubo-Bastin moment computation many times (fresh handle each time, other work in between) and reports
every run that differs from the first one bit for bit -- a detector for races / uninitialised reads."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rslmtoasa_b200 import Recursion, Control, Energy, synthetic as S
from tests.cases import case

n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
bad = 0
for name in ("pbc_hoh", "pbc"):
    lat, ham = case(name)
    ph = S.random_phases(lat.kk, 2)
    first = {}
    for it in range(n):
        for M in (9, 6, 4):
            rec = Recursion(ham, lat, Control(cond_ll=M, cond_calctype="random_vec"), Energy(-2, 2), phases=ph)
            if it % 3 == 1:      # perturb the allocator / leave garbage around
                rec.recur_b(); rec.chebyshev_recur()
            rec.compute_moments_stochastic()
            mu = rec.mu_nm_stochastic
            if (name, M) not in first:
                first[(name, M)] = mu.copy()
            elif not np.array_equal(mu, first[(name, M)]):
                d = np.abs(mu - first[(name, M)])
                bad += 1
                idx = np.unravel_index(np.nanargmax(d), d.shape)
                print("MISMATCH", name, "M", M, "iter", it, "max", float(np.nanmax(d)), "nan", int(np.isnan(mu).sum()), "at", idx,
                      "count", int((d > 0).sum()), flush=True)
            rec.close()
print("done, mismatches:", bad)

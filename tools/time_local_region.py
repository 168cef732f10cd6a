#!/usr/bin/env python
"""Timing point for a large site-indexed (hall) region: config-3 lattice (3838 sites, B2) with nmax = 15 / 500 / 2000,
recur_b lld = 21 through the host API.  Every site of the region has its own 15 Hamiltonian blocks (77.8 kB per site and
application, SURVEY.md 8d 'site-indexed H' row)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rslmtoasa_b200 import Recursion, Control, Energy, synthetic as S  # noqa: E402

for nmax in (0, 15, 500, 2000):
    lat = S.sphere_cluster("bcc", 60.0, ntype=3, nmax=nmax, type_rule="b2")
    ham = S.make_hamiltonian(lat, seed=20260103)
    rec = Recursion(ham, lat, Control(lld=21), Energy(-2.0, 2.0))
    rec.recur_b()
    ts = []
    for _ in range(8):
        t0 = time.perf_counter(); rec.recur_b(); ts.append(time.perf_counter() - t0)
    rec.phase_timing(True); rec.recur_b(); ph = rec.phase_read(); rec.phase_timing(False)
    spmv_ms = ph["H|PSI_n>"][0]
    # algorithmic bytes of the site-indexed part per full-lattice application: nmax sites x 15 blocks x 5184 B (H) -- the psi
    # traffic is the same as for type-indexed sites
    print(json.dumps({"kk": lat.kk, "nmax": nmax, "recur_b_ms_min": 1e3 * min(ts), "H|PSI_n>_device_ms": spmv_ms,
                      "local_H_bytes_per_application": nmax * 15 * 5184, "phases": {k: v[0] for k, v in ph.items()}}), flush=True)
    rec.close()

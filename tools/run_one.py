#!/usr/bin/env python
"""Runs one driver once on a named configuration (for ncu launch lists): python tools/run_one.py <config> <what> [lld] [hoh]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rslmtoasa_b200 import Recursion, Control, Energy, synthetic as S  # noqa: E402

cfg, what = sys.argv[1], sys.argv[2]
lld = int(sys.argv[3]) if len(sys.argv) > 3 else 21
hoh = len(sys.argv) > 4 and sys.argv[4] == "hoh"
kw = {}
if cfg == "bulk":
    lat = S.sphere_cluster("bcc", 80.0); ham = S.make_hamiltonian(lat, seed=20260101, hoh=hoh)
elif cfg == "surface":
    lat = S.sphere_cluster("fcc", 100.0, ntype=7, type_rule="layer"); lat.irec = np.array([1, 2, 3, 14, 15, 20], dtype=np.int32)
    ham = S.make_hamiltonian(lat, seed=20260102, hoh=hoh)
elif cfg == "impurity":
    lat = S.sphere_cluster("bcc", 60.0, ntype=3, nmax=15, type_rule="b2"); ham = S.make_hamiltonian(lat, seed=20260103, hoh=hoh)
elif cfg == "cond":
    lat = S.periodic_bcc(10, 20, 20); ham = S.make_hamiltonian(lat, seed=20260104, velocity=True)
    kw = dict(atlist=[1])
rec = Recursion(ham, lat, Control(lld=lld, cond_ll=lld, cond_calctype="per_type"), Energy(-2.0, 2.0), **kw)
getattr(rec, what)()
print("done", cfg, what, lld, "launches", rec.launch_count)

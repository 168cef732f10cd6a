#!/usr/bin/env python
"""config 1 with collinear (spin-diagonal) hopping blocks, for ncu captures of the spin-resolved SpMV kernel:
python tools/run_collinear.py [recur_b|chebyshev_recur] [lld]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rslmtoasa_b200 import Recursion, Control, Energy, synthetic as S  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "recur_b"
lld = int(sys.argv[2]) if len(sys.argv) > 2 else 21
lat = S.sphere_cluster("bcc", 80.0)
ham = S.make_hamiltonian(lat, seed=20260101, spin_orbit=False)
rec = Recursion(ham, lat, Control(lld=lld), Energy(-2.0, 2.0))
getattr(rec, what)()
print("done", what, lld, "launches", rec.launch_count, "spin-resolved SpMV launches", int(rec._L.rsrec_spin_diag_launch_count(rec._h)))

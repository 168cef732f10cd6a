#!/usr/bin/env python
"""start / end of every g_timer phase of one recur_b on config 1 (us from the first event): RSREC_LZ_TIMELINE=1 python tools/lz_timeline.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("RSREC_LZ_TIMELINE", "1")
from rslmtoasa_b200 import Recursion, Control, Energy, synthetic as S  # noqa: E402
lat = S.sphere_cluster("bcc", 80.0)
ham = S.make_hamiltonian(lat, seed=20260101, spin_orbit=(len(sys.argv) < 2 or sys.argv[1] != "collinear"))
rec = Recursion(ham, lat, Control(lld=21), Energy(-2.0, 2.0))
os.environ["RSREC_LZ_TIMELINE"] = ""
for _ in range(3):
    rec.recur_b()
rec.phase_timing(True)
rec.recur_b()

#!/usr/bin/env python
"""Runs the fused Kubo-Bastin pipeline once on config 4 (8000-site periodic bcc): python tools/run_cond.py [cond_ll] [channels]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rslmtoasa_b200 import Recursion, Control, Energy, Conductivity, synthetic as S
M = int(sys.argv[1]) if len(sys.argv) > 1 else 64
ch = int(sys.argv[2]) if len(sys.argv) > 2 else 2500
lat = S.periodic_bcc(10, 20, 20)
ham = S.make_hamiltonian(lat, seed=20260104, velocity=True)
rec = Recursion(ham, lat, Control(cond_ll=M, cond_calctype="random_vec"), Energy(-2.0, 2.0, channels_ldos=ch, fermi=0.0),
                phases=S.random_phases(lat.kk, 1))
c = Conductivity(rec)
import time
t0 = time.perf_counter(); integ, _ = c.compute_conductivity(); t1 = time.perf_counter()
print("done cond_ll", M, "nv", ch + 10, "seconds", round(t1 - t0, 4), "launches", rec.launch_count, "finite", int(np.isfinite(integ).sum()))

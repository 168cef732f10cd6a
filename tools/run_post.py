#!/usr/bin/env python
"""Runs the fused recur_b -> zsqr -> terminator -> bgreen call once on config 1 or 2 (for ncu launch lists):
python tools/run_post.py <bulk|surface> [channels]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rslmtoasa_b200 import Recursion, Control, Energy, Green, synthetic as S  # noqa: E402

cfg = sys.argv[1]
channels = int(sys.argv[2]) if len(sys.argv) > 2 else 2500
if cfg == "bulk":
    lat = S.sphere_cluster("bcc", 80.0); ham = S.make_hamiltonian(lat, seed=20260101)
else:
    lat = S.sphere_cluster("fcc", 100.0, ntype=7, type_rule="layer"); lat.irec = np.array([1, 2, 3, 14, 15, 20], dtype=np.int32)
    ham = S.make_hamiltonian(lat, seed=20260102)
rec = Recursion(ham, lat, Control(lld=21), Energy(-2.0, 2.0, channels_ldos=channels, fermi=0.0))
g = Green(rec)
if "--bands" in sys.argv:      # the whole SCF iteration: g0 stays on the device, the `bands` consumers run there
    from rslmtoasa_b200.bands import Bands
    import time
    for rep in range(2):
        t0 = time.perf_counter()
        g.recur_b_green(download_g0=False)
        rec.en.fermi = 0.0
        b = Bands(g, qqv=6.0 * len(lat.irec))
        b.calculate_fermi(); b.calculate_magnetic_moments(); b.calculate_moments(); b.calculate_band_energy()
        print("scf step", rep, time.perf_counter() - t0, "s  fermi", rec.en.fermi, "eband", b.eband)
else:
    g.recur_b_green()
    print("done", cfg, "launches", rec.launch_count, "g0", g.g0.shape)

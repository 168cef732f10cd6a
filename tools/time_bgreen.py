#!/usr/bin/env python
"""Times block_green (terminator + block continued fraction, nv = 2510) on configs 1 and 2 from host coefficient arrays:
python tools/time_bgreen.py   (A/B of k_bgreen geometries: rebuild with -DBG_WARPS=.. first)"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if len(sys.argv) > 1:                      # A/B: load an alternative build of the library
    from rslmtoasa_b200 import build as _B
    _B.LIB = os.path.abspath(sys.argv[1])
from rslmtoasa_b200 import Recursion, Control, Energy, Green, synthetic as S  # noqa: E402

for name in ("bccfe", "bulk", "surface"):
    if name == "bccfe":      # the reference's collinear bccFe regression case: hopping blocks are spin-diagonal
        from oracle import oracle as O, ref_bccfe as R
        lat, ham, _, _ = R.case_inputs(O, "Example_bulk_bccFe_nsp2_block")
    elif name == "bulk":
        lat = S.sphere_cluster("bcc", 80.0); ham = S.make_hamiltonian(lat, seed=20260101)
    else:
        lat = S.sphere_cluster("fcc", 100.0, ntype=7, type_rule="layer"); lat.irec = np.array([1, 2, 3, 14, 15, 20], dtype=np.int32)
        ham = S.make_hamiltonian(lat, seed=20260102)
    rec = Recursion(ham, lat, Control(lld=21), Energy(-2.0, 2.0, channels_ldos=2500, fermi=0.0))
    rec.recur_b(); rec.zsqr()
    g = Green(rec)
    g.block_green()
    best = 1e9
    for _ in range(7):
        t0 = time.perf_counter(); g.block_green(); best = min(best, time.perf_counter() - t0)
    print(f"{name}: block_green {best * 1e3:.3f} ms (units {len(lat.irec)}, checksum {np.nansum(np.abs(g.g0)):.12e})", flush=True)
    g.recur_b_green(download_g0=False)
    best = 1e9
    for _ in range(7):
        t0 = time.perf_counter(); g.recur_b_green(download_g0=False); best = min(best, time.perf_counter() - t0)
    print(f"{name}: fused recur_b + green, g0 resident {best * 1e3:.3f} ms  (spin-diagonal SpMV launches: "
          f"{rec._L.rsrec_spin_diag_launch_count(rec._h)})", flush=True)
    rec.close()

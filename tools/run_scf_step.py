#!/usr/bin/env python
"""One SCF-type iteration of the block path on the reference's bccFe regression case (5984 sites, lld = 20, the standard
energy mesh): recursion -> terminator -> Green function -> Fermi level -> moments, twice.  For ncu captures of the
small kernels of the step (k_bgreen, k_lz_eig, k_rmul_dmma, k_gram_dmma, k_bpopt_warp)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O, ref_bccfe as R  # noqa: E402  (case inputs only)
from rslmtoasa_b200 import Recursion, Control, Energy, Green  # noqa: E402
from rslmtoasa_b200.bands import Bands  # noqa: E402

name = "Example_bulk_bccFe_nsp2_block"
lat, ham, ene, g = R.case_inputs(O, name)
rec = Recursion(ham, lat, Control(lld=g["lld"]),
                Energy(g["energy_min"], g["energy_max"], channels_ldos=g["channels_ldos"], fermi=g["fermi"]))
for _ in range(2):
    gr = Green(rec)
    gr.recur_b_green(download_g0=False)
    b = Bands(gr, qqv=8.0)
    b.calculate_fermi()
    b.calculate_moments()
print("done: launches", rec.launch_count, "fermi", rec.en.fermi)
rec.close()

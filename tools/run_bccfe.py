#!/usr/bin/env python
"""Runs recur_b (lld = 20) once on the reference's collinear bccFe regression case (for ncu captures of the spin-diagonal SpMV)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O, ref_bccfe as R  # noqa: E402
from rslmtoasa_b200 import Recursion, Control, Energy  # noqa: E402

lat, ham, _, g = R.case_inputs(O, "Example_bulk_bccFe_nsp2_block")
rec = Recursion(ham, lat, Control(lld=20), Energy(g["energy_min"], g["energy_max"]))
rec.recur_b()
rec.recur_b()
print("done", rec._L.rsrec_spin_diag_launch_count(rec._h), rec.launch_count)

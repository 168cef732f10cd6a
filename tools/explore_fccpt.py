#!/usr/bin/env python
"""Exploration: which operator / hoh setting of the conductivity post-processing case reproduces the reference's stored
Pt_cond.out values (tests/postproc/references/Example_exchange_conductivity_fccPt*/ref.json)?  GPU run."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O, ref_fccpt as P, ham_oracle as HO, ref_bccfe as R  # noqa: E402
from rslmtoasa_b200 import Recursion, Control, Energy, Conductivity  # noqa: E402

inp = P.INPUT
for hoh in (False, True):
    t0 = time.time()
    lat, ham, ene, cr = P.build_case(O, hoh=hoh)
    print("case built", lat.kk, lat.nn.shape, time.time() - t0, flush=True)
    obarm = HO.build_obarm(R.build_pot(P.PT), P.PT["mom"].reshape(3, 1)) if hoh else None
    for out_op in ("charge", "spin_z", "spin_z_in"):
        ham.v_a, ham.vo_a = P.velocity_blocks(lat, ham, cr, inp["alat"], inp["v_alpha"], "z" if out_op == "spin_z" else None, obarm)
        ham.v_b, ham.vo_b = P.velocity_blocks(lat, ham, cr, inp["alat"], inp["v_beta"], "z" if out_op == "spin_z_in" else None, obarm)
        en = Energy(inp["energy_min"], inp["energy_max"], channels_ldos=inp["channels_ldos"], fermi=inp["fermi"])
        rec = Recursion(ham, lat, Control(lld=50, cond_ll=inp["cond_ll"], cond_calctype="per_type"), en, atlist=[1])
        c = Conductivity(rec)
        t0 = time.time()
        c.compute_conductivity()
        sig = c.integrate_conductivity()
        print(f"hoh={hoh} out={out_op}: {time.time() - t0:.2f} s", flush=True)
        for row in (500, 1000, 1500):
            print(f"   row {row}: E-Ef {c.ene[row - 1] - inp['fermi']:.7f}  re {sig[0, 0, row - 1, 1]:.7e}  im {sig[1, 0, row - 1, 1]:.7e}   golden {P.GOLDEN['Example_exchange_conductivity_fccPt' + ('_hoh' if hoh else '')][row]}", flush=True)
        rec.close()

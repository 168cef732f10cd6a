#!/usr/bin/env python
"""Run-to-run reproducibility of the recursion drivers (same handle, fresh handle): prints max |difference|."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rslmtoasa_b200 import Recursion, Control, Energy, synthetic as S
from tests.cases import case

for name in ("surface", "impurity_hoh", "bulk"):
    lat, ham = case(name)
    for fam in (0, 1):
        outs = []
        rec = Recursion(ham, lat, Control(lld=8), Energy(-2, 2))
        rec.set_kernel_family(fam)
        for rep in range(3):
            rec.recur_b(); outs.append((rec.a_b.copy(), rec.b2_b.copy()))
        rec2 = Recursion(ham, lat, Control(lld=8), Energy(-2, 2))
        rec2.set_kernel_family(fam)
        rec2.recur_b(); outs.append((rec2.a_b.copy(), rec2.b2_b.copy()))
        rec.chebyshev_recur(); m1 = rec.mu_n.copy(); rec.chebyshev_recur(); m2 = rec.mu_n.copy()
        rec2.chebyshev_recur(); m3 = rec2.mu_n.copy()
        print(name, "family", fam, "recur_b same-handle", [float(np.abs(outs[0][0] - o[0]).max()) for o in outs[1:3]],
              "fresh", float(np.abs(outs[0][0] - outs[3][0]).max()), "first differing ll",
              [int(l) for l in range(8) if not np.array_equal(outs[0][0][:, :, l], outs[3][0][:, :, l])][:1],
              "cheb same", float(np.abs(m1 - m2).max()), "fresh", float(np.abs(m1 - m3).max()))

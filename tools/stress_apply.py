#!/usr/bin/env python
"""Bitwise repeatability of single operator applications (H~ and velocity, hoh and plain, both kernel families)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rslmtoasa_b200 import Recursion, Control, Energy, synthetic as S
from tests.cases import case

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
rng = np.random.default_rng(0)
for name in ("pbc_hoh", "pbc", "impurity_hoh"):
    lat, ham = case(name)
    psi = np.asfortranarray(rng.normal(size=(18, 18, lat.kk)) + 1j * rng.normal(size=(18, 18, lat.kk)))
    for fam in (1, 0):
        rec = Recursion(ham, lat, Control(lld=5), Energy(-2, 2))
        rec.set_kernel_family(fam)
        ops = [("ham", lambda: rec.ham_vec_matmul(psi, 2.3, 0.1))]
        if getattr(ham, "v_a", None) is not None:
            ops += [("velo_a", lambda: rec.velo_vec_matmul("a", psi)), ("velo_b", lambda: rec.velo_vec_matmul("b", psi))]
        for label, fn in ops:
            first = fn()
            bad = 0
            for it in range(n):
                out = fn()
                if not np.array_equal(out, first):
                    bad += 1
                    d = np.abs(out - first)
                    sites = np.unique(np.nonzero(d.max(axis=(0, 1)) > 0)[0])
                    if bad <= 3:
                        print("MISMATCH", name, "family", fam, label, "iter", it, "max", float(d.max()), "sites", sites[:12], len(sites), flush=True)
            print(name, "family", fam, label, "mismatches", bad, "of", n, flush=True)
        rec.close()

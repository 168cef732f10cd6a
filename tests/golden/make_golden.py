"""Generates tests/golden/oracle_golden.npz from the CPU oracle (run from the repo root).

The reference cannot be executed in this image (no Fortran compiler) and stores no a_n/b_n/mu_n fixtures, so
these vectors are ORACLE outputs (validated against oracle/dense_check.py at generation time): they pin the
oracle against its own regressions and give the GPU tests fixed numbers to hit.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import oracle as O, dense_check as D  # noqa: E402
from tests.cases import case, EMIN, EMAX, relerr  # noqa: E402

a, b = O.cheb_scale(EMIN, EMAX)
out = {}
lat, ham = case("impurity_hoh")
orc = O.Oracle(lat, ham)
out["imp_a_b"], out["imp_b2_b"] = orc.lanczos_block([1], 6)
out["imp_mu"], _ = orc.cheb_moments([1], 6, a, b)
H = D.dense_hamiltonian(lat, ham)
da, db = D.block_lanczos(H, D.start_block(lat, 1), 6)
assert relerr(out["imp_a_b"][..., 0], da) < 1e-11 and relerr(out["imp_b2_b"][..., 0], db) < 1e-11
assert relerr(out["imp_mu"][..., 0], D.cheb_moments(H, D.start_block(lat, 1), 6, a, b)) < 1e-11
lat, ham = case("bulk")
out["bulk_sa"], out["bulk_sb"] = O.Oracle(lat, ham).lanczos_scalar([1], 8)
lat, ham = case("pbc")
out["pbc_kubo"] = O.Oracle(lat, ham).kubo_moments(4, a, b, start_sites=[1])
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_golden.npz"), **out)
print({k: v.shape for k, v in out.items()})

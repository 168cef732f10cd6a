"""Generates tests/golden/post_golden.npz: oracle outputs of the consumers either side of the recursion (SURVEY 8f
rows 1-3), computed FROM the committed coefficient/moment fixtures in oracle_golden.npz and validated against the
independent numpy restatement (oracle/dense_check_post.py) at generation time.  Run from the repo root."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import oracle as O, dense_check_post as DP  # noqa: E402
from tests.cases import EMIN, EMAX, relerr  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
g = np.load(os.path.join(HERE, "oracle_golden.npz"))
out = {}
ene = O.e_mesh(-1.6, 0.4, 20, 0.05)
out["ene"] = ene
a_b, b_b = g["imp_a_b"], O.Oracle.zsqr(None, g["imp_b2_b"])
out["b_b"] = b_b
out["a_inf"], out["b_inf"], out["a_inf0"], out["b_inf0"] = O.get_terminf(a_b, b_b)
out["g0_block"] = O.block_green(a_b, b_b, ene)
out["g0_block_sym"] = O.block_green(a_b, b_b, ene, True)
ref = DP.bgreen(a_b[..., 0], b_b[..., 0], ene, out["a_inf"][..., 0], out["b_inf"][..., 0], 0.0, False)
assert relerr(out["g0_block"][..., 0], ref) < 1e-10
out["mu_ng"], out["g0_cheb"] = O.chebyshev_green(g["imp_mu"], ene, EMIN, EMAX)
assert relerr(out["g0_cheb"], DP.chebyshev_green(g["imp_mu"], ene, EMIN, EMAX)) < 1e-12
rng = np.random.default_rng(8)
out["dw"] = 1.0 + 0.05 * rng.normal(size=(18, 1))
out["cs"] = 0.02 * rng.normal(size=(18, 1))
out["tdens"] = O.density(g["bulk_sa"][..., 0], g["bulk_sb"][..., 0], ene, out["dw"][:, 0], out["cs"][:, 0])
mu_nm = g["pbc_kubo"]
out["integrand"], out["integrand_at"] = O.conductivity_integrand(mu_nm, ene, EMIN, EMAX, True)
ri, _ = DP.conductivity_integrand(mu_nm, ene, EMIN, EMAX)
assert relerr(np.nan_to_num(out["integrand"]), np.nan_to_num(ri)) < 1e-12
np.savez_compressed(os.path.join(HERE, "post_golden.npz"), **out)
print({k: v.shape for k, v in out.items()})

"""Parity of the device-side Hamiltonian assembly (rsrec_build_hamiltonian) with oracle/ham_oracle.py, and of the
device HR36 packing: a recursion on device-assembled blocks equals the recursion on the same blocks uploaded from the host."""
import numpy as np
import pytest

from oracle import ham_oracle as HO
from tests.cases import case, relerr, EMIN, EMAX

pytestmark = pytest.mark.gpu


def _inputs(lat, seed, scale=0.06):
    """structure-constant-like blocks with S(-R) = S(R)^T so that the assembled operator is Hermitian"""
    from rslmtoasa_b200 import synthetic as S
    rng = np.random.default_rng(seed)
    nt, ns, nl = lat.ntype, lat.nslot, lat.nmax
    ncls = nt + nl
    opp = S._opposite_slots(lat.disp)
    base = rng.normal(size=(9, 9, ns)) * scale
    for m in range(1, len(lat.disp)):
        if opp[m] < m:
            base[:, :, m] = base[:, :, opp[m]].T
    base[:, :, 0] = 0.5 * (base[:, :, 0] + base[:, :, 0].T)
    hhh = np.repeat(base[:, :, :, None], ncls, axis=3).copy(order="F")
    it = np.array(list(range(1, nt + 1)) + [int(lat.iz[i]) for i in range(nl)], np.int32)
    jt = np.zeros((ns, ncls), np.int32, order="F")
    for c in range(ncls):
        site = int(np.nonzero(lat.iz == c + 1)[0][0]) if c < nt else c - nt
        jt[0, c] = lat.iz[site]
        for m in range(1, lat.nn[site, 0]):
            nb = lat.nn[site, m]
            jt[m, c] = lat.iz[nb - 1] if nb > 0 else 0
    pot = {k: (0.8 + 0.2 * rng.normal(size=(9, nt)) if k in ("wx0",) else 0.1 * rng.normal(size=(9, nt))).astype(complex)
           for k in HO.POT_KEYS}
    pot["cx"] = (0.2 * rng.normal(size=(9, 2, nt))).astype(complex)
    pot["cex"] = (0.1 * rng.normal(size=(9, 2, nt))).astype(complex)
    mom = rng.normal(size=(3, nt)); mom /= np.linalg.norm(mom, axis=0)
    return hhh, jt, it, pot, mom


@pytest.mark.parametrize("hoh", [False, True])
@pytest.mark.parametrize("name", ["bulk", "impurity", "surface"])
def test_build_hamiltonian_vs_oracle(name, hoh):
    from rslmtoasa_b200 import Recursion, Control, Energy
    lat, ham = case(name)
    hhh, jt, it, pot, mom = _inputs(lat, 11)
    rec = Recursion(ham, lat, Control(lld=5), Energy(EMIN, EMAX))
    out = rec.build_hamiltonian(hhh, jt, it, pot, mom, ham.lsham, hoh)
    blk, blko, obarm, enim = HO.build_blocks(hhh, jt, it, pot, mom, hoh)
    nt = lat.ntype
    assert relerr(out["ee"], blk[..., :nt]) < 1e-13
    assert relerr(out["obarm"], obarm) < 1e-13 and relerr(out["enim"], enim) < 1e-13
    if lat.nmax:
        assert relerr(out["hall"], blk[..., nt:]) < 1e-13
    if hoh:
        assert relerr(out["eeo"], blko[..., :nt]) < 1e-13
        if lat.nmax:
            assert relerr(out["hallo"], blko[..., nt:]) < 1e-13
    assert not out["ee"][:, :, jt[:, 0] == 0, 0].any()                 # empty slots stay zero blocks


@pytest.mark.parametrize("hoh", [False, True])
def test_recursion_on_device_built_blocks(oracle_mod, hoh):
    """device-assembled sets, never uploaded from the host == the same blocks passed through rsrec_set_hamiltonian
    (bitwise), and both match the oracle run on the oracle-assembled blocks"""
    import copy
    from rslmtoasa_b200 import Recursion, Control, Energy
    lat, ham = case("impurity")
    hhh, jt, it, pot, mom = _inputs(lat, 12)
    rec = Recursion(ham, lat, Control(lld=7), Energy(EMIN, EMAX))
    out = rec.build_hamiltonian(hhh, jt, it, pot, mom, ham.lsham, hoh)
    rec.recur_b(); rec.chebyshev_recur()
    a1, b1, m1 = rec.a_b.copy(), rec.b2_b.copy(), rec.mu_n.copy()
    ham2 = copy.copy(ham)
    ham2.ee, ham2.hall, ham2.hoh = out["ee"], out["hall"], hoh
    ham2.eeo, ham2.hallo, ham2.enim = out["eeo"], out["hallo"], out["enim"]
    rec2 = Recursion(ham2, lat, Control(lld=7), Energy(EMIN, EMAX))
    rec2.recur_b(); rec2.chebyshev_recur()
    assert np.array_equal(rec2.a_b, a1) and np.array_equal(rec2.b2_b, b1) and np.array_equal(rec2.mu_n, m1)
    blk, blko, obarm, enim = HO.build_blocks(hhh, jt, it, pot, mom, hoh)
    ham3 = copy.copy(ham2)
    nt = lat.ntype
    ham3.ee, ham3.hall = np.asfortranarray(blk[..., :nt]), np.asfortranarray(blk[..., nt:])
    ham3.eeo, ham3.hallo, ham3.enim = np.asfortranarray(blko[..., :nt]), np.asfortranarray(blko[..., nt:]), np.asfortranarray(enim)
    orc = oracle_mod.Oracle(lat, ham3)
    oa, ob = orc.lanczos_block(lat.irec, 7)
    assert relerr(a1, oa) < 1e-10 and relerr(b1, ob) < 1e-10


@pytest.mark.parametrize("name", ["surface", "impurity_hoh"])
def test_recur_b_local_axis(oracle_mod, name):
    """per-unit rotation of the block sets on the device (rotmag_loc) == the oracle run on numpy-rotated host arrays"""
    import copy
    from rslmtoasa_b200 import Recursion, Control, Energy
    lat, ham = case(name)
    if name == "impurity_hoh":
        lat.irec = np.array([1, 3], dtype=np.int32)
    rng = np.random.default_rng(9)
    mom = rng.normal(size=(3, len(lat.irec))); mom /= np.linalg.norm(mom, axis=0)
    mom[:, 0] = [0.0, 0.0, 1.0]                               # identity rotation for the first unit
    rec = Recursion(ham, lat, Control(lld=6), Energy(EMIN, EMAX))
    rec.recur_b()
    plain = rec.a_b.copy()
    rec.recur_b_local_axis(mom)
    assert relerr(rec.a_b[..., 0], plain[..., 0]) < 1e-13     # m = z: nothing changes
    for u, site in enumerate(lat.irec):
        h2 = copy.copy(ham)
        h2.ee = HO.rotmag_loc(ham.ee, mom[:, u])
        if lat.nmax:
            h2.hall = HO.rotmag_loc(ham.hall, mom[:, u])
        if ham.hoh:
            h2.eeo, h2.enim = HO.rotmag_loc(ham.eeo, mom[:, u]), HO.rotmag_loc(ham.enim, mom[:, u])
            if lat.nmax:
                h2.hallo = HO.rotmag_loc(ham.hallo, mom[:, u])
        oa, ob = oracle_mod.Oracle(lat, h2).lanczos_block([site], 6)
        assert relerr(rec.a_b[..., u], oa[..., 0]) < 1e-10 and relerr(rec.b2_b[..., u], ob[..., 0]) < 1e-10
    # the sets stay rotated to the last unit's axis (like the reference) until rotate_from_local_axis
    rec.recur_b()
    assert relerr(rec.a_b[..., -1], plain[..., -1]) > 1e-6
    rec.rotate_from_local_axis()
    rec.recur_b()
    assert np.array_equal(rec.a_b, plain)

"""CPU tests of the `bands` oracle (oracle/rsrec_oracle_bands.c): pinned by an independent numpy statement
(Pauli-matrix traces + vectorised Simpson weights, oracle/dense_check_post.py), by closed forms, and -- for the DOS --
by the reference's own golden totaldos.out values (tests/test_reference_golden.py uses the same g0 -> dtot map)."""
import numpy as np
import pytest

from oracle import dense_check_post as DP


def _g0(nv, nu, seed=0):
    rng = np.random.default_rng(seed)
    g = rng.standard_normal((18, 18, nv, nu)) + 1j * rng.standard_normal((18, 18, nv, nu))
    g[np.arange(18), np.arange(18)] -= 3j           # negative imaginary diagonal: a positive DOS
    return np.asfortranarray(g)


def test_simpson_m_closed_forms(oracle_mod):
    m = oracle_mod.e_mesh_full(-1.0, 1.0, 200, 0.137)
    ene, edel = m["ene"], m["edel"]
    # cubic integrand: Simpson is exact on the full panels; Fermi level on a mesh point (e1 == fermi: no end correction)
    nv1 = 101
    e1 = fermi = ene[nv1 - 1]
    y = 1.0 + ene + ene ** 2
    for nexp in (0, 1):
        got = oracle_mod.simpson_m(edel, fermi, nv1, y, e1, nexp, ene)
        F = lambda x: (x + x ** 2 / 2 + x ** 3 / 3) if nexp == 0 else (x ** 2 / 2 + x ** 3 / 3 + x ** 4 / 4)
        assert abs(got - (F(fermi) - F(ene[0]))) < 1e-13
    # Fermi level between mesh points: the end panel (ef - ea)/6 (f0 + 4 f1 + f2) as the reference writes it
    fermi2 = e1 + 0.4 * edel
    got = oracle_mod.simpson_m(edel, fermi2, nv1, y, e1, 2, ene)
    want = DP.simpson_to_fermi(y, ene, edel, fermi2, nv1, e1, 2)
    assert abs(got - want) < 1e-14 * max(1.0, abs(want))


def test_fermi_scan_finds_the_valence_charge(oracle_mod):
    m = oracle_mod.e_mesh_full(-1.0, 1.0, 400, 0.2)
    ene, edel = m["ene"], m["edel"]
    dtot = 3.0 + 0.0 * ene                           # constant DOS: N(E) = 3 (E - Emin)
    qqv = 3.0 * 0.7137
    ef, nv1, e1, ifail = oracle_mod.bands_fermi(dtot, edel, -1.0, qqv, 0.2, m["nv1"])
    assert ifail == 0 and abs(ef - (-1.0 + 0.7137)) < 1e-13
    assert nv1 % 2 == 1 and abs(e1 - ene[nv1 - 1]) < 1e-13 and e1 <= ef < e1 + 2 * edel
    # the charge integrated by simpson_m up to that level is the valence
    assert abs(oracle_mod.simpson_m(edel, ef, nv1, dtot, e1, 0, ene) - qqv) < 1e-13
    # not enough states: ifail = 1 and the inputs come back
    ef2, nv2, _, ifail2 = oracle_mod.bands_fermi(dtot, edel, -1.0, 1e3, 0.2, m["nv1"])
    assert ifail2 == 1 and ef2 == 0.2 and nv2 == m["nv1"]
    # fixed Fermi level (bands.f90:338-341)
    ef3, nv3, e3, _ = oracle_mod.bands_fermi(dtot, edel, -1.0, qqv, 0.2, m["nv1"], fix_fermi=True)
    assert ef3 == 0.2 and nv3 == round((0.2 + 1.0) / edel) and abs(e3 - (-1.0 + (nv3 - 1) * edel)) < 1e-15


@pytest.mark.parametrize("nu", [1, 3])
def test_projections_and_moments_match_the_numpy_statement(oracle_mod, nu):
    m = oracle_mod.e_mesh_full(-1.2, 0.9, 120, 0.05)
    ene, edel, nv = m["ene"], m["edel"], len(m["ene"])
    g0 = _g0(nv, nu, seed=nu)
    rng = np.random.default_rng(7)
    mom = rng.standard_normal((3, nu)); mom /= np.linalg.norm(mom, axis=0)
    P = DP.bands_projections(g0, mom)
    dtot, dosia, dosial = oracle_mod.bands_dos(g0)
    assert np.allclose(dosia, P["dos"], rtol=0, atol=1e-13)
    assert np.allclose(dtot, P["dos"].sum(axis=1), rtol=0, atol=1e-12)
    assert np.allclose(dosial.sum(axis=0), dosia, rtol=0, atol=1e-13)
    qqv = 0.45 * oracle_mod.simpson_m(edel, ene[m["nv1"] - 1], m["nv1"], dtot, ene[m["nv1"] - 1], 0, ene)
    ef, nv1, e1, ifail = oracle_mod.bands_fermi(dtot, edel, -1.2, qqv, 0.05, m["nv1"])
    assert ifail == 0
    m0, m1 = oracle_mod.bands_magnetic_moments(g0, ene, edel, ef, nv1, e1)
    for d in range(3):
        assert np.allclose(m0[d], DP.simpson_to_fermi(P["spin"][d], ene[:, None], edel, ef, nv1, e1, 0), rtol=1e-12, atol=1e-13)
        assert np.allclose(m1[d], DP.simpson_to_fermi(P["spin"][d], ene[:, None], edel, ef, nv1, e1, 1), rtol=1e-12, atol=1e-13)
    occ, lmom = oracle_mod.bands_moments(g0, m["channels_ldos"], mom, ene, edel, ef, nv1, e1)
    dspd = P["dspd"].copy(); dspd[:, m["channels_ldos"]:] = 0.0
    for q in range(6):
        for k in range(3):
            assert np.allclose(occ[k, q], DP.simpson_to_fermi(dspd[q], ene[:, None], edel, ef, nv1, e1, k), rtol=1e-12, atol=1e-13)
    for d in range(3):
        assert np.allclose(lmom[d], -DP.simpson_to_fermi(P["lorb"][d], ene[:, None], edel, ef, nv1, e1, 0) / np.pi, rtol=1e-12, atol=1e-13)
    # sum rule: the six channel occupations add up to the integrated local DOS (channels below the cut)
    tot = DP.simpson_to_fermi(np.where(np.arange(nv)[:, None] < m["channels_ldos"], P["dos"], 0.0), ene[:, None], edel, ef, nv1, e1, 0)
    assert np.allclose(occ[0].sum(axis=0), tot, rtol=1e-12)
    # spin moment along mom = occupation difference of the two spin channels
    mz = sum(mom[d] * m0[d] for d in range(3))
    assert np.allclose(occ[0, :3].sum(axis=0) - occ[0, 3:].sum(axis=0), mz, rtol=1e-10, atol=1e-12)


def test_orbital_moment_vanishes_without_spin_orbit_coupling(oracle_mod):
    """g0 = f(E) * identity has Tr(L g0) = 0 (the L matrices are traceless); a g0 ~ L_z gives lmom along z only"""
    m = oracle_mod.e_mesh_full(-1.0, 1.0, 60, 0.0)
    ene, nv = m["ene"], len(m["ene"])
    L = oracle_mod.l_spherical()
    g0 = np.zeros((18, 18, nv, 1), complex, order="F")
    f = -1j / (1.0 + ene ** 2)
    for s in (0, 9):
        g0[s:s + 9, s:s + 9, :, 0] = f[None, None, :] * (np.eye(9)[:, :, None] + 0.1 * L[:, :, 2, None])
    occ, lmom = oracle_mod.bands_moments(g0, m["channels_ldos"], np.array([[0.0], [0.0], [1.0]]), ene, m["edel"], 0.0, 31, ene[30])
    assert abs(lmom[0, 0]) < 1e-14 and abs(lmom[1, 0]) < 1e-14
    want = -0.1 * 2 * np.real(np.trace(L[:, :, 2] @ L[:, :, 2])) * DP.simpson_to_fermi(np.imag(f), ene, m["edel"], 0.0, 31, ene[30], 0) / np.pi
    assert abs(lmom[2, 0] - want) < 1e-13

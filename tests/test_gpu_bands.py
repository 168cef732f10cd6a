"""GPU parity of the `bands` consumers (rsrec_bands_*: bands.f90) against the CPU oracle, through the C ABI, and the
SCF-level quantities the north star names (band energy, charges) on the reference's bccFe regression case."""
import ctypes as C
import numpy as np
import pytest

from rslmtoasa_b200 import Recursion, Control, Energy, Green, RsrecError
from rslmtoasa_b200.bands import Bands
from tests.cases import case, EMIN, EMAX

pytestmark = pytest.mark.gpu


def _rec(lld=12, hoh=False, fermi=0.05, channels=160):
    lat, ham = case("pbc_hoh" if hoh else "surface")
    if hoh:
        lat.irec = np.array([1, 7, 20], dtype=np.int32)
    en = Energy(energy_min=EMIN, energy_max=EMAX, channels_ldos=channels, fermi=fermi)
    return Recursion(ham, lat, Control(lld=lld), en), lat, ham


def _g0_random(nv, nu, seed):
    rng = np.random.default_rng(seed)
    g = rng.standard_normal((18, 18, nv, nu)) + 1j * rng.standard_normal((18, 18, nv, nu))
    g[np.arange(18), np.arange(18)] -= 3j
    return np.asfortranarray(g)


def _mom(nu, seed=3):
    m = np.random.default_rng(seed).standard_normal((3, nu))
    return np.asfortranarray(m / np.linalg.norm(m, axis=0))


@pytest.mark.parametrize("nu", [1, 4])
def test_bands_on_uploaded_g0_matches_oracle(oracle_mod, nu):
    rec, lat, ham = _rec()
    gr = Green(rec)
    b = Bands(gr, qqv=0.0, nsp=4)
    en = rec.en
    nv = len(en.ene)
    g0 = _g0_random(nv, nu, seed=nu)
    b.set_g0(g0)
    dtot_o, dosia_o, dosial_o = oracle_mod.bands_dos(g0)
    b.qqv = 0.4 * oracle_mod.simpson_m(en.edel, en.ene[en.nv1 - 1], en.nv1, dtot_o, en.ene[en.nv1 - 1], 0, en.ene)
    ik1 = en.ik1
    b.calculate_fermi(ldos=True)
    assert np.array_equal(b.dtot, dtot_o)                                   # same summation order: bit-exact
    assert np.array_equal(b.dosia, dosia_o) and np.array_equal(b.dosial, dosial_o)
    ef, nv1, e1, ifail = oracle_mod.bands_fermi(dtot_o, en.edel, en.energy_min, b.qqv, 0.05, ik1)
    assert ifail == 0 and b.ifail == 0
    assert (en.fermi, b.nv1, b.e1) == (ef, nv1, e1)                         # branch decisions and arithmetic identical
    m0, m1 = oracle_mod.bands_magnetic_moments(g0, en.ene, en.edel, ef, nv1, e1)
    b.calculate_magnetic_moments()
    assert np.allclose(b.mom0, m0, rtol=1e-12, atol=1e-13) and np.allclose(b.mom1, m1, rtol=1e-12, atol=1e-13)
    assert np.allclose(np.linalg.norm(b.mom, axis=0), 1.0)
    b.mom = _mom(nu)
    occ, lmom = oracle_mod.bands_moments(g0, en.channels_ldos, b.mom, en.ene, en.edel, ef, nv1, e1)
    b.calculate_moments()
    assert np.allclose(b.occ, occ, rtol=1e-12, atol=1e-13) and np.allclose(b.lmom, lmom, rtol=1e-12, atol=1e-13)
    assert np.allclose(b.ql[0].transpose(1, 0, 2).reshape(6, nu), occ[0], rtol=1e-12)
    eb = oracle_mod.simpson_m(en.edel, ef, nv1, dtot_o, e1, 1, en.ene)
    assert abs(b.calculate_band_energy() - eb) <= 1e-12 * abs(eb)


@pytest.mark.parametrize("recur", ["block", "chebyshev"])
def test_fused_green_keeps_g0_on_device_for_bands(oracle_mod, recur):
    """recursion -> Green function -> bands with g0 never downloaded == the staged flow with g0 on the host"""
    rec, lat, ham = _rec(hoh=(recur == "block"))
    gr = Green(rec)
    before = rec._L.rsrec_d2h_bytes(rec._h)
    g0 = gr.recur_b_green(download_g0=False) if recur == "block" else gr.chebyshev_recur_green(keep_moments=False, download_g0=False)
    assert g0 is None
    moved = rec._L.rsrec_d2h_bytes(rec._h) - before
    nv, nu = len(rec.en.ene), len(lat.irec)
    assert moved < 18 * 18 * 16 * nv * nu / 4                                # no g0-sized download happened
    b = Bands(gr, qqv=5.0 * nu)
    b.calculate_fermi(); b.calculate_magnetic_moments(); ql = b.calculate_moments().copy(); eb = b.calculate_band_energy()
    # staged: download g0, check against the oracle chain
    g0h = np.zeros((18, 18, nv, nu), np.complex128, order="F")
    rec._L.rsrec_bands_get_g0(rec._h, g0h.ctypes.data_as(C.c_void_p))
    g0s = gr.recur_b_green() if recur == "block" else gr.chebyshev_recur_green(keep_moments=False)
    assert np.array_equal(g0h, g0s)
    dtot_o = oracle_mod.bands_dos(g0h)[0]
    assert np.array_equal(b.dtot, dtot_o)
    en = rec.en
    ef, nv1, e1, ifail = oracle_mod.bands_fermi(dtot_o, en.edel, en.energy_min, b.qqv, 0.05, en.ik1)
    assert ifail == 0 and (en.fermi, b.nv1, b.e1) == (ef, nv1, e1)
    occ, lmom = oracle_mod.bands_moments(g0h, en.channels_ldos, b.mom, en.ene, en.edel, ef, nv1, e1)
    assert np.allclose(b.occ, occ, rtol=1e-12, atol=1e-13)
    assert abs(eb - oracle_mod.simpson_m(en.edel, ef, nv1, dtot_o, e1, 1, en.ene)) <= 1e-12 * abs(eb)
    assert abs(ql[0].sum() - b.qqv) < 5e-2 * b.qqv                           # the channel charges add up to ~ the valence


def test_bands_argument_checks_and_fixed_fermi(oracle_mod):
    rec, lat, ham = _rec()
    gr = Green(rec)
    b = Bands(gr, qqv=1.0)
    with pytest.raises(RsrecError):
        b.calculate_fermi()                                                 # no g0 on the device yet
    nv = len(rec.en.ene)
    g0 = _g0_random(nv, 2, seed=9)
    b.set_g0(g0)
    rec.en.fix_fermi = True
    b.calculate_fermi()
    ef, nv1, e1, _ = oracle_mod.bands_fermi(b.dtot, rec.en.edel, rec.en.energy_min, 1.0, 0.05, rec.en.ik1, fix_fermi=True)
    assert (rec.en.fermi, b.nv1, b.e1) == (ef, nv1, e1)
    rec.en.fix_fermi = False
    b.qqv = 1e9                                                            # valence never reached: ifail, nothing updated
    b.calculate_fermi()
    assert b.ifail == 1 and rec.en.fermi == 0.05
    with pytest.raises(RsrecError):                                        # simpson_m reads Y(NPTS+2)
        b.nv1 = nv
        b.calculate_band_energy()


def test_scf_quantities_of_the_reference_bccfe_case(oracle_mod):
    """north star: 'SCF band energy and charges to 1e-8 Ry' -- the whole chain on the reference's own regression case
    (recursion -> terminator -> Green function -> Fermi level -> band moments), GPU against the CPU oracle"""
    from oracle import ref_bccfe as R
    name = "Example_bulk_bccFe_nsp2_block"
    lat, ham, ene, g = R.case_inputs(oracle_mod, name)
    en = Energy(energy_min=g["energy_min"], energy_max=g["energy_max"], channels_ldos=R.INPUT["channels_ldos"], fermi=R.INPUT["fermi"])
    rec = Recursion(ham, lat, Control(lld=g["lld"]), en)
    gr = Green(rec)
    gr.recur_b_green(download_g0=False)
    b = Bands(gr, qqv=8.0)                                                  # Fe: 8 valence electrons
    b.calculate_fermi(); b.calculate_magnetic_moments(); b.calculate_moments(); eb = b.calculate_band_energy()
    # oracle chain
    orc = oracle_mod.Oracle(lat, ham)
    a_b, b2_b = orc.lanczos_block(lat.irec, g["lld"])
    g0 = oracle_mod.block_green(a_b, orc.zsqr(b2_b), ene)
    dtot = oracle_mod.bands_dos(g0)[0]
    ef, nv1, e1, ifail = oracle_mod.bands_fermi(dtot, en.edel, en.energy_min, 8.0, R.INPUT["fermi"], en.ik1)
    assert ifail == 0 and b.nv1 == nv1
    assert abs(en.fermi - ef) < 1e-8 and abs(b.e1 - e1) < 1e-12
    occ, lmom = oracle_mod.bands_moments(g0, en.channels_ldos, b.mom, ene, en.edel, ef, nv1, e1)
    assert np.abs(b.occ - occ).max() < 1e-8                                # charges and band moments, Ry units
    assert abs(eb - oracle_mod.simpson_m(en.edel, ef, nv1, dtot, e1, 1, ene)) < 1e-8
    m0, _ = oracle_mod.bands_magnetic_moments(g0, ene, en.edel, ef, nv1, e1)
    assert np.abs(b.mom0 - m0).max() < 1e-8
    assert 1.5 < b.mom0[2, 0] < 3.0 and abs(occ[0].sum() - 8.0) < 1e-6     # bcc Fe: ~2.2 mu_B, 8 electrons


def test_g0_too_large_to_keep_falls_back_to_the_host_copy(oracle_mod, monkeypatch):
    """g0 of all units stays on the device only if it leaves room for the recursion; otherwise the fused call streams it to
    the host batch by batch as before (same values), refuses g0 = NULL, and the bands calls ask for rsrec_bands_set_g0"""
    rec, lat, ham = _rec()
    gr = Green(rec)
    want = gr.recur_b_green().copy()
    monkeypatch.setenv("RSREC_G0_RESIDENT_MAX_MB", "1")
    got = gr.recur_b_green()
    assert np.array_equal(got, want)
    b = Bands(gr, qqv=1.0)
    with pytest.raises(RsrecError):
        b.calculate_fermi()
    with pytest.raises(RsrecError):
        gr.recur_b_green(download_g0=False)
    monkeypatch.setenv("RSREC_G0_RESIDENT_MAX_MB", "100000")
    b.set_g0(want)
    b.qqv = 2.0
    b.calculate_fermi()
    assert np.array_equal(b.dtot, oracle_mod.bands_dos(want)[0])

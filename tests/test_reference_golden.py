"""The oracle against the REFERENCE'S OWN GOLDEN VALUES: `tests/scf/references/Example_bulk_bccFe_*/ref.json` of
rslmtoasa/rslmtoasa (totaldos.out rows 500 / 1000 / 1500, written by the Fortran program with 5 decimals).

oracle/ref_bccfe.py restates the reference's pipeline for that regression case from the input files up (cluster, neighbour
table, screened structure constants, potential parameters, L.S, Hamiltonian blocks) and the C oracle does the recursion
(hop_b / hop_b_hoh / crecal_b or the Chebyshev moments), zsqr, terminator + block continued fraction or chebyshev_green.
All twelve fixtures of the case are reproduced to every printed digit -- this is what pins the oracle."""
import numpy as np
import pytest

from oracle import ref_bccfe as R

NAMES = sorted(n for n in R.GOLDEN if "nsp4" not in n)     # nsp = 4 runs the same recursion as nsp = 2 (same fixtures)


def _check(ene, dos, g, tol=6e-6):
    for row, (e_ref, d_ref) in g["rows"].items():
        assert abs((ene[row - 1] - g["fermi"]) - e_ref) < 6e-6            # the mesh (energy%e_mesh) hits the same points
        assert abs(dos[row - 1] - d_ref) < tol, (row, dos[row - 1], d_ref)


def test_cluster_and_structure_constants_of_the_case(oracle_mod):
    lat, ham, ene, g = R.case_inputs(oracle_mod, "Example_bulk_bccFe_nsp2_block")
    assert lat.kk == 5984 and lat.nn.shape == (5984, 16) and (lat.nn[0, 1:15] > 0).all()      # kk = 5984
    assert len(ene) == 2510
    h_on = ham.ee[:, :, 0, 0]
    assert np.abs(h_on - h_on.conj().T).max() < 1e-14
    # cubic symmetry of the on-site block: p orbitals degenerate, d split into t2g / eg
    d = np.real(np.diag(h_on))[:9]
    assert np.ptp(d[1:4]) < 1e-12 and abs(d[4] - d[8]) < 1e-12 and abs(d[5] - d[7]) < 1e-12


@pytest.mark.parametrize("name", NAMES)
def test_oracle_reproduces_reference_golden(oracle_mod, name):
    ene, dos = R.oracle_total_dos(oracle_mod, name)
    g = R.case_inputs(oracle_mod, name)[3]
    _check(ene, dos, g)


def test_nsp4_fixtures_equal_nsp2():
    for n, g in R.GOLDEN.items():
        if "nsp4" in n:
            assert g["rows"] == R.GOLDEN[n.replace("nsp4", "nsp2")]["rows"]


def test_embedded_golden_table_is_the_reference_fixture_file():
    """oracle/ref_bccfe.py:GOLDEN == tests/golden/reference_bccfe_ref.json (copied from the reference tree by
    tests/golden/make_reference_golden.py), including the settings of every case"""
    import json
    import os
    ref = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_bccfe_ref.json")))["cases"]
    from oracle import ref_fccpt as P
    for name, rows in P.GOLDEN.items():               # the conductivity fixtures embedded in oracle/ref_fccpt.py
        for row, (e, v) in rows.items():
            assert ref[name]["Pt_cond.out"][str(row)] == {"1": e, "2": v}
    from oracle import ref_exchange as X
    for name, rows in X.GOLDEN.items():               # the exchange fixtures embedded in oracle/ref_exchange.py
        for row, (j, dist) in rows.items():
            assert ref[name]["jij.out"][str(row)] == {"6": j, "7": dist}
        assert ref[name]["namelists"]["control"]["lld"] == X.INPUT["lld"]
    ref = {k: v for k, v in ref.items() if "totaldos.out" in v}
    assert set(ref) == set(R.GOLDEN)
    for name, g in R.GOLDEN.items():
        nml = ref[name]["namelists"]
        assert nml["control"] == {"nsp": g["nsp"], "recur": g["recur"], "lld": g["lld"]}
        assert nml["hamiltonian"]["hoh"] == g["hoh"] and nml["self"]["nstep"] == 1
        assert ("energy" in nml) == ("window" in g)
        if "window" in g:
            assert (nml["energy"]["energy_min"], nml["energy"]["energy_max"]) == g["window"]
        for row, (e, d) in g["rows"].items():
            assert ref[name]["totaldos.out"][str(row)] == {"1": e, "2": d}


@pytest.mark.parametrize("hoh", [False, True])
def test_oracle_reproduces_reference_conductivity_golden(oracle_mod, hoh):
    """tests/postproc/references/Example_exchange_conductivity_fccPt{,_hoh}/ref.json: Pt_cond.out rows 500/1000/1500 of
    the reference's Kubo-Bastin post-processing (fcc Pt, 8000 sites, cond_ll = 50, spin-Hall operator pair) reproduced by
    the oracle chain kubo moments -> Gamma contraction -> Fermi-weighted Simpson tail.  This pins compute_moments_stochastic
    (with the hoh operator variants), the velocity products, calculate_gamma_nm and calculate_conductivity_tensor."""
    import os
    from oracle import ref_fccpt as P
    if hoh and not os.environ.get("RSREC_SLOW_TESTS"):
        pytest.skip("second 8000-site CPU Kubo run (~1.5 min); set RSREC_SLOW_TESTS=1 -- the GPU test covers both cases")
    ene, re = P.oracle_conductivity(oracle_mod, hoh)
    worst = P.check_rows(ene, re, "Example_exchange_conductivity_fccPt" + ("_hoh" if hoh else ""))
    assert worst < 1e-5


@pytest.mark.parametrize("hoh", [False, True])
def test_oracle_reproduces_reference_exchange_golden(oracle_mod, hoh):
    """tests/postproc/references/Example_exchange_bccFe{,_hoh}/ref.json: J_ij of the nearest and next-nearest neighbour pair
    (jij.out, f12.6) from the oracle chain recur_b_ij -> zsqr -> block_green_ij -> calculate_intersite_gf; pins the pair
    recursion (four start-vector combinations) and the inter-site Green functions, with and without hoh."""
    from oracle import ref_exchange as X
    jij, dist = X.oracle_jij(oracle_mod, hoh)
    assert X.check(jij, dist, "Example_exchange_bccFe" + ("_hoh" if hoh else "")) < 1.5e-6

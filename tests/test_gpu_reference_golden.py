"""The CUDA path against the REFERENCE'S OWN GOLDEN VALUES (tests/scf/references/Example_bulk_bccFe_*/ref.json of
rslmtoasa/rslmtoasa): the bccFe regression case, 5984 sites, run through the C ABI -- recursion + Green function fused on the
device -- must print the same totaldos.out rows as the Fortran program did (5 decimals)."""
import numpy as np
import pytest

from oracle import ref_bccfe as R

pytestmark = pytest.mark.gpu

NAMES = sorted(n for n in R.GOLDEN if "nsp4" not in n)


def _rec(lat, ham, g):
    from rslmtoasa_b200 import Recursion, Control, Energy
    en = Energy(g["energy_min"], g["energy_max"], channels_ldos=g["channels_ldos"], fermi=g["fermi"])
    return Recursion(ham, lat, Control(lld=g["lld"]), en)


def _check(ene, dos, g):
    for row, (e_ref, d_ref) in g["rows"].items():
        assert abs((ene[row - 1] - g["fermi"]) - e_ref) < 6e-6
        assert abs(dos[row - 1] - d_ref) < 6e-6, (row, dos[row - 1], d_ref)


@pytest.mark.parametrize("name", NAMES)
def test_gpu_reproduces_reference_golden(oracle_mod, name):
    from rslmtoasa_b200 import Green
    lat, ham, ene, g = R.case_inputs(oracle_mod, name)
    rec = _rec(lat, ham, g)
    gr = Green(rec)
    assert np.array_equal(gr.ene, ene)
    g0 = gr.recur_b_green() if g["recur"] == "block" else gr.chebyshev_recur_green(keep_moments=False)
    _check(ene, R.total_dos(g0), g)


def test_gpu_pipeline_from_coordinates_reproduces_reference_golden(oracle_mod):
    """nothing but coordinates, structure-constant blocks and potential parameters enter: the neighbour table
    (rsrec_build_nn) and the Hamiltonian blocks (rsrec_build_hamiltonian) are built on the device as well"""
    from rslmtoasa_b200 import Recursion, Control, Energy, Green, synthetic as S
    from rslmtoasa_b200.lattice import build_nn
    inp, name = R.INPUT, "Example_bulk_bccFe_nsp2_block_hoh"
    g = R.GOLDEN[name]
    cr = R.bravais_cluster(inp["rc"])
    crd = np.asfortranarray(cr * inp["alat"])
    kk = cr.shape[1]
    nn, nm = build_nn(crd, np.ones(kk, np.int32), [1], inp["ct"])
    sbar, _ = R.screened_structure_constants(crd, 1, inp["r2"], inp["wav"])
    nslot, nr = nn.shape[1], int(nn[0, 0])
    hhh = np.zeros((9, 9, nslot, 1), order="F")
    for m in range(nr):
        hhh[:, :, m, 0] = sbar[:, :, m].T
    jt = np.zeros((nslot, 1), np.int32); jt[:nr] = 1
    lat = S.Lattice(kk=kk, nn=nn, iz=np.ones(kk, np.int32), ntype=1, nmax=0, irec=np.array([1], np.int32), cr=cr)
    placeholder = S.Hamiltonian(ee=np.zeros((18, 18, nslot, 1), complex, order="F"), lsham=np.zeros((18, 18, 1), complex, order="F"))
    en = Energy(inp["energy_min"], inp["energy_max"], channels_ldos=inp["channels_ldos"], fermi=inp["fermi"])
    rec = Recursion(placeholder, lat, Control(lld=g["lld"]), en)
    rec.build_hamiltonian(hhh, jt, np.array([1], np.int32), R.build_pot(R.FE), R.FE["mom"].reshape(3, 1),
                          R.build_lsham(R.FE["xi_p"], R.FE["xi_d"]), hoh=True, download=False)
    gr = Green(rec)
    g0 = gr.recur_b_green()
    _check(gr.ene, R.total_dos(g0), dict(g, fermi=inp["fermi"]))


@pytest.mark.parametrize("hoh", [False, True])
def test_gpu_reproduces_reference_conductivity_golden(oracle_mod, hoh):
    """the reference's stored Pt_cond.out values (tests/postproc/references/Example_exchange_conductivity_fccPt{,_hoh}) from
    the CUDA path: fused Kubo moments + Gamma contraction (rsrec_kubo_conductivity) + the sigma(E_F) tail
    (rsrec_conductivity_cumulative), 8000 sites, cond_ll = 50"""
    from oracle import ref_fccpt as P
    from rslmtoasa_b200 import Recursion, Control, Energy, Conductivity
    lat, ham, ene, mesh = P.kubo_inputs(oracle_mod, hoh)
    inp = P.INPUT
    en = Energy(inp["energy_min"], inp["energy_max"], channels_ldos=inp["channels_ldos"], fermi=inp["fermi"])
    rec = Recursion(ham, lat, Control(lld=50, cond_ll=inp["cond_ll"], cond_calctype="per_type"), en, atlist=[1])
    c = Conductivity(rec)
    assert np.array_equal(c.ene, ene)
    c.compute_conductivity()
    sig = c.integrate_conductivity()
    worst = P.check_rows(ene, sig[0, 0, :, 1], "Example_exchange_conductivity_fccPt" + ("_hoh" if hoh else ""))
    assert worst < 1e-5
    # the staged path (moments on the host, then the integrand) gives the same curve
    rec.compute_moments_stochastic()
    c.calculate_conductivity_tensor()
    sig2 = c.integrate_conductivity()
    assert np.allclose(sig2[0, 0, :, 1], sig[0, 0, :, 1], rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("hoh", [False, True])
def test_gpu_reproduces_reference_exchange_golden(oracle_mod, hoh):
    """the reference's stored J_ij (tests/postproc/references/Example_exchange_bccFe{,_hoh}/ref.json) from the CUDA path: fused
    pair recursion + Green functions (rsrec_recur_b_ij_green) + rsrec_intersite_gf on the device-resident g0"""
    from oracle import ref_exchange as X
    from rslmtoasa_b200 import Recursion, Control, Energy, Green
    lat, ham, ene, mesh = X.case_inputs(oracle_mod, hoh)
    inp = X.INPUT
    en = Energy(inp["energy_min"], inp["energy_max"], channels_ldos=inp["channels_ldos"], fermi=inp["fermi"])
    rec = Recursion(ham, lat, Control(lld=inp["lld"], recur="block"), en, ijpair=X.PAIRS)
    gr = Green(rec)
    assert np.array_equal(gr.ene, ene)
    gr.calculate_intersite_gf(fused=True)
    gs = np.stack([gr.ginmag, gr.gix, gr.giy, gr.giz, gr.gjnmag, gr.gjx, gr.gjy, gr.gjz], axis=-1)
    jij = X.jij_from_spin_components(gs, ene, mesh["nv1"])
    dist = np.linalg.norm(lat.cr[:, X.PAIRS[:, 1] - 1] - lat.cr[:, X.PAIRS[:, 0] - 1], axis=0)
    assert X.check(jij, dist, "Example_exchange_bccFe" + ("_hoh" if hoh else "")) < 1.5e-6

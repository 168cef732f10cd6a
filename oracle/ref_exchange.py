"""TEST INFRASTRUCTURE (part of the CPU oracle): the reference's exchange post-processing case restated from its input
files -- tests/postproc/cases/exchange/bccFe (bcc Fe, rc = 80, two pairs: nearest and next-nearest neighbour of atom 1),
whose stored outputs are tests/postproc/references/Example_exchange_bccFe{,_hoh}/ref.json (jij.out column 6 = J_ij in mRy,
column 7 = |r_ij|).  It pins recur_b_ij (recursion.f90:1655-1737) and calculate_intersite_gf (green.f90:425-469).

    post_processing_exchange (calculation.f90:~860-950): recur_b_ij -> calculate_intersite_gf -> calculate_exchange
    symbolic_atom%predls / d_matrix (symbolic_atom.f90:205-265)   -> d_matrix
    exchange%calculate_exchange, dGdG_Jnc (exchange.f90:1437-1530, 933-959) -> jij
The Heisenberg formula itself is outside the B200 library's scope; it is restated here only to reach the stored numbers.
"""
import numpy as np

from . import ref_bccfe as R

# ---- tests/postproc/cases/exchange/bccFe/input.nml (lld patched to 20 by tests/postproc/cases.json) ---------------------
INPUT = dict(R.INPUT, fermi=-0.069291, energy_min=-1.0, energy_max=1.2, channels_ldos=2500, lld=20)
PAIRS = np.array([[1, 2634], [1, 2635]], np.int32)
# ---- Fe.nml (&par) entries d_matrix needs, index [l, spin]; the file is identical to the SCF case's Fe.nml ------------------
FE_X = dict(
    c=np.array([[-0.29300025894758364, -0.26602765430747233], [0.72509183510161801, 0.75374707098984417],
                [-0.21467013394120280, -6.6044636984103622E-002]]),
    srdel=np.array([[0.43123549046157306, 0.43517259745556613], [0.41466306625733867, 0.41686905695212784],
                    [0.11560545723205057, 0.12982554663219267]]),
    ws_r=2.6621999999999999, vmad=3.3815563284505739E-014)
ANG2AU = 1.8897259886                                                  # math.f90:88
# ---- tests/postproc/references/Example_exchange_bccFe{,_hoh}/ref.json: jij.out row -> (J_ij, |r_ij|) ----------------------
GOLDEN = {"Example_exchange_bccFe": {1: (0.71797, 0.866025), 2: (0.485359, 1.0)},
          "Example_exchange_bccFe_hoh": {1: (1.919175, 0.866025), 2: (0.371106, 1.0)}}


def d_matrix(e):
    """symbolic_atom%d_matrix (241-265) with dele from predls (205-239): wow = wav*ang2au / ws_r"""
    wow = INPUT["wav"] * ANG2AU / FE_X["ws_r"]
    dele = FE_X["srdel"] * wow ** (0.5 - np.arange(1, 4))[:, None]
    cu, cd = FE_X["c"][:, 0] + FE_X["vmad"], FE_X["c"][:, 1] + FE_X["vmad"]
    wu, wd = dele[:, 0], dele[:, 1]
    de = (cd * wu * wu - cu * wd * wd + (wd * wd - wu * wu) * e) / (wu * wd)
    return np.diag(np.repeat(de, [1, 3, 5])).astype(complex)


def jij_from_spin_components(gs, ene, nv1):
    """gs (9,9,nv,njij,8) = Ginmag, Gix, Giy, Giz, Gjnmag, Gjx, Gjy, Gjz -> J_ij (mRy) per pair:
    dGdG_Jnc + the T = 0 Fermi-weighted Simpson integral up to en%fermi (simpson_f) * 1e3 / 4 pi"""
    nv, njij = gs.shape[2], gs.shape[3]
    h = ene[1] - ene[0]
    ifermi = int(np.argmin(np.abs(ene - INPUT["fermi"])))               # the mesh hits the Fermi level (energy%e_mesh)
    assert abs(ene[ifermi] - INPUT["fermi"]) < 1e-12
    W = np.ones(nv); W[1::2] = 4.0; W[2::2] = 2.0                        # composite Simpson weights of simpson_f's panels
    with np.errstate(over="ignore"):                                     # fermifun (math.f90:994-1000) at kBT = 1e-15, literal:
        f = 1.0 / (np.exp((ene - INPUT["fermi"]) / 1.0e-15) + 1.0)       # the mesh point "at" E_F is off by ~1e-17 -> 0.4965, not 1/2
    out = np.zeros(njij)
    for p in range(njij):
        y = np.zeros(nv)
        for iv in range(nv):
            D = d_matrix(ene[iv])
            j = (D @ gs[:, :, iv, p, 0]) @ (D @ gs[:, :, iv, p, 4])
            for c in (1, 2, 3):
                j = j - (D @ gs[:, :, iv, p, c]) @ (D @ gs[:, :, iv, p, 4 + c])
            y[iv] = np.trace(j).imag
        out[p] = h / 3.0 * np.sum(W * f * y) * 1.0e3 / 4.0 / np.pi
    return out


def case_inputs(oracle_mod, hoh):
    lat, ham, ene = R.build_case(oracle_mod, hoh=hoh, inp=INPUT)
    mesh = oracle_mod.e_mesh_full(INPUT["energy_min"], INPUT["energy_max"], INPUT["channels_ldos"], INPUT["fermi"])
    return lat, ham, ene, mesh


def pair_units(pairs):
    """the four start-vector combinations of every pair (recursion.f90:1655-1737), packed"""
    s = 1.0 / np.sqrt(2.0)
    si, sj, asg, bsg = [], [], [], []
    for i, j in pairs:
        for sg in (1.0, -1.0, 1j, -1j):
            si.append(i); sj.append(j); asg.append(s); bsg.append(s * sg)
    return np.array(si, np.int32), np.array(sj, np.int32), np.array(asg, complex), np.array(bsg, complex)


def oracle_jij(oracle_mod, hoh):
    lat, ham, ene, mesh = case_inputs(oracle_mod, hoh)
    orc = oracle_mod.Oracle(lat, ham)
    si, sj, asg, bsg = pair_units(PAIRS)
    a_b, b2_b = orc.lanczos_block(si, INPUT["lld"], site_j=sj, asign=asg, bsign=bsg)
    g0 = oracle_mod.block_green(a_b, orc.zsqr(b2_b), ene)
    _, _, gs = oracle_mod.intersite_gf(g0, PAIRS)
    dist = np.linalg.norm(lat.cr[:, PAIRS[:, 1] - 1] - lat.cr[:, PAIRS[:, 0] - 1], axis=0)
    return jij_from_spin_components(gs, ene, mesh["nv1"]), dist


def check(jij, dist, name):
    """jij.out is written with f12.6; the harness tolerance is 1e-6 absolute or relative (tests/run_test.py:206-215)"""
    worst = 0.0
    for row, (j_ref, d_ref) in GOLDEN[name].items():
        assert abs(dist[row - 1] - d_ref) < 1e-6, (row, dist[row - 1], d_ref)
        worst = max(worst, abs(jij[row - 1] - j_ref))
    return worst

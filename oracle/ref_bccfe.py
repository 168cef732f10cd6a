"""CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE): the reference's own regression case `tests/scf/cases/bulk/bccFe`
(nsp = 2, recur = block | chebyshev, lld = 20 | 100, nstep = 1) restated end to end, so that the recursion path can be pinned
against the reference's GOLDEN values `tests/scf/references/Example_bulk_bccFe_nsp2_*/ref.json` (totaldos.out rows
500 / 1000 / 1500), which the reference produced with its Fortran program.

With nstep = 1 the density of states written by `bands%calculate_fermi` (bands.f90:256-286) depends only on the INPUT
potential file (cases/bulk/bccFe/Fe.nml), the lattice and the structure constants -- no atomic solver is involved
(self.f90:683-712: run_recursion, then run_dos).  Restated here, statement for statement (numpy, small arrays):

    lattice%bravais + cut          lattice.f90:1006-1111, 3236-3268   cluster r^2 <= rc, generation order, kk made even
    lattice%nncal / remd           -> oracle.build_nn (rsrec_oracle_lattice.c)
    lattice%dbar1/clusba/micha/streze/shldch/canso   lattice.f90:2176-2672   screened structure constants S-bar
    symbolic_atom%build_pot        symbolic_atom.f90:163-195          potential parameters -> cx, wx, ...
    hamiltonian%build_lsham        hamiltonian.f90:1370-1420          (L matrices: math.f90:133-165)
    hamiltonian%build_bulkham      -> ham_oracle.build_blocks (hmfind: hhh(ilm,jlm) = sbar(jlm,ilm,m), 2417-2421)
    recursion%recur_b, zsqr        -> rsrec_oracle.c
    green%block_green / chebyshev_green -> rsrec_oracle_post.c
    bands: dtot = -Im sum_j (g0(j,j) + g0(j+9,j+9)) / pi    bands.f90:262
    energy%e_mesh                  -> oracle.e_mesh

The case inputs (potential parameters, lattice constants, mesh) are the numbers of the reference's input files, embedded
below so that the tests do not read /root/reference at run time.
"""
import numpy as np

from . import ham_oracle as HO

# ---- tests/scf/cases/bulk/bccFe/input.nml --------------------------------------------------------------------------
INPUT = dict(rc=80.0, alat=2.86120, wav=1.40880, ct=3.0, r2=9.0, fermi=-0.042265, energy_min=-2.0, energy_max=0.8,
             channels_ldos=2500)
# ---- tests/scf/cases/bulk/bccFe/Fe.nml (&par), index [l, spin] -----------------------------------------------------
FE = dict(
    center_band=np.array([[-0.30514742661207822, -0.27864492541420444], [0.34233634594168177, 0.39646896272705467],
                          [-0.21434211411691506, -5.8149847044184126E-002]]),
    width_band=np.array([[0.39990076064881702, 0.40308477172565943], [0.26074746866140475, 0.26808667950112641],
                         [0.11763839722311017, 0.13693483067531095]]),
    shifted_band=np.array([[0.15504326682966943, 0.15851574024532303], [0.64846897620728838, 0.64381585530588414],
                           [1.8934758575734675E-002, 0.15194499502218797]]),
    obar=np.array([[-0.46860906872401609, -0.46511382559922465], [-0.57237283814674811, -0.55433193077624476],
                   [0.93104368668848847, 0.36069514680089260]]),
    mom=np.array([0.0, 0.0, 1.0]),
    xi_p=np.array([1.2378281144950869E-002, 1.2717009101469700E-002]),
    xi_d=np.array([4.4521948116027481E-003, 3.4878427600180673E-003]),
)
# ---- tests/scf/references/<name>/ref.json: totaldos.out {row: (E - E_F, DOS)}; settings from tests/scf/cases.json -------
# nsp = 2 and 4 build lsham (self.f90:780), nsp = 3 leaves it zero; the chebyshev_hoh cases widen the energy window.
def _g(r500, r1000, r1500, e=(-1.39886, -0.83887, -0.27888)):
    return {500: (e[0], r500), 1000: (e[1], r1000), 1500: (e[2], r1500)}


_EW = (-1.99935, -1.03905, -0.07874)
GOLDEN = {
    "Example_bulk_bccFe_nsp2_block": dict(recur="block", lld=20, hoh=False, nsp=2, rows=_g(0.0, 0.0, 21.60495)),
    "Example_bulk_bccFe_nsp2_block_hoh": dict(recur="block", lld=20, hoh=True, nsp=2, rows=_g(0.0, 0.0, 23.80755)),
    "Example_bulk_bccFe_nsp2_chebyshev": dict(recur="chebyshev", lld=100, hoh=False, nsp=2, rows=_g(2e-05, 0.00058, 21.59953)),
    "Example_bulk_bccFe_nsp2_chebyshev_hoh": dict(recur="chebyshev", lld=100, hoh=True, nsp=2, window=(-3.0, 1.8),
                                                  rows=_g(2e-05, 0.00046, 25.86365, _EW)),
    "Example_bulk_bccFe_nsp3_block": dict(recur="block", lld=20, hoh=False, nsp=3, rows=_g(0.0, 0.0, 21.88928)),
    "Example_bulk_bccFe_nsp3_block_hoh": dict(recur="block", lld=20, hoh=True, nsp=3, rows=_g(0.0, 0.0, 23.53861)),
    "Example_bulk_bccFe_nsp3_chebyshev": dict(recur="chebyshev", lld=100, hoh=False, nsp=3, rows=_g(2e-05, 0.00058, 21.58127)),
    "Example_bulk_bccFe_nsp3_chebyshev_hoh": dict(recur="chebyshev", lld=100, hoh=True, nsp=3, window=(-3.0, 1.8),
                                                  rows=_g(2e-05, 0.00046, 25.89247, _EW)),
    "Example_bulk_bccFe_nsp4_block": dict(recur="block", lld=20, hoh=False, nsp=4, rows=_g(0.0, 0.0, 21.60495)),
    "Example_bulk_bccFe_nsp4_block_hoh": dict(recur="block", lld=20, hoh=True, nsp=4, rows=_g(0.0, 0.0, 23.80755)),
    "Example_bulk_bccFe_nsp4_chebyshev": dict(recur="chebyshev", lld=100, hoh=False, nsp=4, rows=_g(2e-05, 0.00058, 21.59953)),
    "Example_bulk_bccFe_nsp4_chebyshev_hoh": dict(recur="chebyshev", lld=100, hoh=True, nsp=4, window=(-3.0, 1.8),
                                                  rows=_g(2e-05, 0.00046, 25.86365, _EW)),
}

BCC_A = np.array([[-0.5, 0.5, 0.5], [0.5, -0.5, 0.5], [0.5, 0.5, -0.5]]).T      # a(:, i) columns (lattice.f90:739-741)


def bravais_cluster(rc):
    """lattice%bravais for 'bcc' (ntot = 1): cr (3,kk) in units of alat, origin first, then the translations in the
    reference's loop order (nx outer, nz inner), cut at r^2 <= rc; kk made even by dropping the last site (1091)."""
    R = int(np.ceil(np.sqrt(2.0 * rc))) + 1
    rng = np.arange(-R, R + 1)
    p, q, s = np.meshgrid(rng, rng, rng, indexing="ij")          # nx - lc, ny - lc, nz - lc ; C order = nx outer, nz inner
    pts = (p.ravel()[None, :] * BCC_A[:, [0]] + q.ravel()[None, :] * BCC_A[:, [1]] + s.ravel()[None, :] * BCC_A[:, [2]])
    keep = ((pts ** 2).sum(0) <= rc) & ~((p.ravel() == 0) & (q.ravel() == 0) & (s.ravel() == 0))
    cr = np.concatenate([np.zeros((3, 1)), pts[:, keep]], axis=1)
    kk = cr.shape[1]
    if kk % 2:
        kk -= 1
    return np.asfortranarray(cr[:, :kk])


def canso(dr):
    """canonical structure constants between s, p, d orbitals (lattice.f90:2540-2672) for w = 1; 9x9"""
    r1, r2, r3 = dr
    rr = np.sqrt(r1 * r1 + r2 * r2 + r3 * r3)
    sc = np.zeros((10, 10))                                        # 1-based like the reference
    if rr <= 0.30:
        return sc[1:, 1:]
    sbyr = 1.0 / rr
    s2 = sbyr * sbyr; s3 = s2 * sbyr; s4 = s3 * sbyr; s5 = s4 * sbyr
    sq3, sq5 = np.sqrt(3.0), np.sqrt(5.0)
    el, em, en = r1 / rr, r2 / rr, r3 / rr
    el2, em2, en2 = el * el, em * em, en * en
    elem, elen, emen = el * em, el * en, em * en
    sc[1, 1] = -2.0 * sbyr
    sc[1, 2] = el * s2 * 2.0 * sq3
    sc[1, 3] = em * s2 * 2.0 * sq3
    sc[1, 4] = en * s2 * 2.0 * sq3
    sc[1, 5] = -2.0 * sq3 * sq5 * elem * s3
    sc[1, 6] = -2.0 * sq3 * sq5 * emen * s3
    sc[1, 7] = -2.0 * sq3 * sq5 * elen * s3
    sc[1, 8] = -sq3 * sq5 * s3 * (el2 - em2)
    sc[1, 9] = sq5 * s3 * (1.0 - 3.0 * en2)
    sc[2, 2] = (3.0 * el2 - 1.0) * 6.0 * s3
    sc[2, 3] = 18.0 * s3 * elem
    sc[2, 4] = 18.0 * s3 * elen
    sc[2, 5] = 6.0 * sq5 * s4 * em * (1.0 - 5.0 * el2)
    sc[2, 6] = -30.0 * sq5 * s4 * elem * en
    sc[2, 7] = 6.0 * sq5 * s4 * en * (1.0 - 5.0 * el2)
    sc[2, 8] = 6.0 * sq5 * s4 * el * (1.0 - 2.5 * el2 + 2.5 * em2)
    sc[2, 9] = 3.0 * sq3 * sq5 * s4 * el * (1.0 - 5.0 * en2)
    sc[3, 3] = 6.0 * s3 * (3.0 * em2 - 1.0)
    sc[3, 4] = 18.0 * s3 * emen
    sc[3, 5] = 6.0 * sq5 * s4 * el * (1.0 - 5.0 * em2)
    sc[3, 6] = 6.0 * sq5 * s4 * en * (1.0 - 5.0 * em2)
    sc[3, 7] = sc[2, 6]
    sc[3, 8] = -6.0 * sq5 * s4 * em * (1.0 - 2.5 * em2 + 2.5 * el2)
    sc[3, 9] = 3.0 * sq3 * sq5 * s4 * em * (1.0 - 5.0 * en2)
    sc[4, 4] = 6.0 * s3 * (3.0 * en2 - 1.0)
    sc[4, 5] = sc[2, 6]
    sc[4, 6] = 6.0 * sq5 * s4 * em * (1.0 - 5.0 * en2)
    sc[4, 7] = 6.0 * sq5 * s4 * el * (1.0 - 5.0 * en2)
    sc[4, 8] = -15.0 * sq5 * s4 * en * (el * el - em2)
    sc[4, 9] = 3.0 * sq3 * sq5 * s4 * en * (3.0 - 5.0 * en2)
    sc[5, 5] = 10.0 * s5 * (-35.0 * el2 * em2 - 5.0 * en2 + 4.0)
    sc[5, 6] = -50.0 * s5 * elen * (7.0 * em2 - 1.0)
    sc[5, 7] = -50.0 * s5 * emen * (7.0 * el2 - 1.0)
    sc[5, 8] = -175.0 * s5 * elem * (el2 - em2)
    sc[5, 9] = -25.0 * sq3 * s5 * elem * (7.0 * en2 - 1.0)
    sc[6, 6] = 10.0 * s5 * (-35.0 * em2 * en2 - 5.0 * el2 + 4.0)
    sc[6, 7] = -50.0 * s5 * elem * (7.0 * en2 - 1.0)
    sc[6, 8] = 50.0 * s5 * emen * (3.5 * em2 - 3.5 * el2 - 1.0)
    sc[6, 9] = -25.0 * sq3 * s5 * emen * (7.0 * en2 - 3.0)
    sc[7, 7] = 10.0 * s5 * (-35.0 * el2 * en2 - 5.0 * em2 + 4.0)
    sc[7, 8] = -50.0 * s5 * elen * (3.5 * el2 - 3.5 * em2 - 1.0)
    sc[7, 9] = -25.0 * sq3 * s5 * elen * (7.0 * en2 - 3.0)
    sc[8, 8] = 10.0 * s5 * (-8.75 * (el2 - em2) ** 2 - 5.0 * en2 + 4.0)
    sc[8, 9] = -12.5 * sq3 * s5 * (7.0 * en2 - 1.0) * (el2 - em2)
    sc[9, 9] = -7.5 * s5 * (35.0 * en2 * en2 - 30.0 * en2 + 3.0)
    for l in range(2, 10):
        for j in range(1, l):
            sc[l, j] = sc[j, l]
    for l in range(1, 4):
        sc[l + 1, 1] = -sc[l + 1, 1]
    for l in range(5, 10):
        for j in range(2, 5):
            sc[l, j] = -sc[l, j]
    return -0.5 * sc[1:, 1:]                                       # ip = identity: s(j,i) = -0.5 sc(j,i)


def screened_structure_constants(crd, ia, r2, wav, ncut=9):
    """dbar1 (lattice.f90:2176-2229): S-bar blocks of the neighbours of site `ia` (1-based) within r2, in ascending site
    order with the site itself first -- the slot order of nncal for the representative atom.  crd = cr*alat (3,kk).
    -> sbar (9,9,nr), vectors (3,nr)"""
    d = crd - crd[:, [ia - 1]]
    s1 = (d ** 2).sum(0)
    big = np.nonzero((s1 < ncut * r2) & (s1 > 0.0001))[0]           # clusba(ncut*r2): ascending index, self first
    vec = np.concatenate([np.zeros((3, 1)), d[:, big]], axis=1)
    nr = vec.shape[1]
    S = np.zeros((9 * nr, 9 * nr))
    for ir in range(nr):                                            # STREZE
        for jr in range(nr):
            if ir != jr:
                S[9 * ir:9 * ir + 9, 9 * jr:9 * jr + 9] = canso((vec[:, jr] - vec[:, ir]) / wav)
    q = 2.0 * np.array([0.3485, 0.05303, 0.010714])
    bet = np.tile(np.repeat(1.0 / q, [1, 3, 5]), nr)                # SHLDCH
    st = S + np.diag(bet)
    c = np.linalg.cholesky(st)                                      # DPOTRF / DPOTRS on the first nlm columns
    x = np.linalg.solve(c.T, np.linalg.solve(c, S[:, :9]))
    x = -bet[:, None] * x
    sb, vs = [], []
    for ir in range(nr):
        if abs((vec[:, ir] ** 2).sum() - (vec[:, 0] ** 2).sum()) <= (ncut * r2) / ncut:
            sb.append(2.0 * x[9 * ir:9 * ir + 9, :])
            vs.append(vec[:, ir])
    return np.stack(sb, axis=2), np.stack(vs, axis=1)


def _L_matrices():
    """L_x, L_y, L_z of math.f90:133-165: reshape(column-major) * (-i)"""
    s3 = np.sqrt(3.0)
    lx = [[0, 0, 0, 0, 0, 0, 0, 0, 0], [0, 0, 0, 0, 0, 0, 0, 0, 0], [0, 0, 0, -1, 0, 0, 0, 0, 0], [0, 0, 1, 0, 0, 0, 0, 0, 0],
          [0, 0, 0, 0, 0, 0, -1, 0, 0], [0, 0, 0, 0, 0, 0, 0, -1, -s3], [0, 0, 0, 0, 1, 0, 0, 0, 0], [0, 0, 0, 0, 0, 1, 0, 0, 0],
          [0, 0, 0, 0, 0, s3, 0, 0, 0]]
    ly = [[0, 0, 0, 0, 0, 0, 0, 0, 0], [0, 0, 0, 1, 0, 0, 0, 0, 0], [0, 0, 0, 0, 0, 0, 0, 0, 0], [0, -1, 0, 0, 0, 0, 0, 0, 0],
          [0, 0, 0, 0, 0, 1, 0, 0, 0], [0, 0, 0, 0, -1, 0, 0, 0, 0], [0, 0, 0, 0, 0, 0, 0, -1, s3], [0, 0, 0, 0, 0, 0, 1, 0, 0],
          [0, 0, 0, 0, 0, 0, -s3, 0, 0]]
    lz = [[0, 0, 0, 0, 0, 0, 0, 0, 0], [0, 0, -1, 0, 0, 0, 0, 0, 0], [0, 1, 0, 0, 0, 0, 0, 0, 0], [0, 0, 0, 0, 0, 0, 0, 0, 0],
          [0, 0, 0, 0, 0, 0, 0, 2, 0], [0, 0, 0, 0, 0, 0, 1, 0, 0], [0, 0, 0, 0, 0, -1, 0, 0, 0], [0, 0, 0, 0, -2, 0, 0, 0, 0],
          [0, 0, 0, 0, 0, 0, 0, 0, 0]]
    # each inner list is one COLUMN of the Fortran matrix (reshape fills column-major)
    return [np.array(m, dtype=float).T * (-1j) for m in (lx, ly, lz)]


def build_lsham(xi_p, xi_d):
    """hamiltonian%build_lsham (hamiltonian.f90:1370-1420), orb_pol = .false.; one type -> (18,18,1)"""
    Lx, Ly, Lz = (HO.hcpx_cart2sph(m) for m in _L_matrices())
    ls = np.zeros((18, 18), complex)
    sg = 0.5
    soc_p, soc_d = np.sqrt(xi_p[0] * xi_p[1]), np.sqrt(xi_d[0] * xi_d[1])
    prefac = 0.0
    for i in range(1, 10):
        for j in range(1, 10):
            if 2 <= i <= 4 and 2 <= j <= 4:
                prefac = sg * soc_p
            if 5 <= i <= 9 and 5 <= j <= 9:
                prefac = sg * soc_d
            ls[j - 1, i - 1] += prefac * Lz[j - 1, i - 1]
            ls[j - 1, i + 8] += prefac * (Lx[j - 1, i - 1] - 1j * Ly[j - 1, i - 1])
            ls[j + 8, i - 1] += prefac * (Lx[j - 1, i - 1] + 1j * Ly[j - 1, i - 1])
            ls[j + 8, i + 8] += -prefac * Lz[j - 1, i - 1]
    return np.asfortranarray(ls[:, :, None])


def build_pot(par):
    """symbolic_atom%build_pot: (l, spin) parameters -> per-orbital complex arrays of one type"""
    rep = [1, 3, 5]
    cx = np.repeat(par["center_band"], rep, axis=0).astype(complex)
    wx = np.repeat(par["width_band"], rep, axis=0).astype(complex)
    cex = np.repeat(par["shifted_band"], rep, axis=0).astype(complex)
    obx = np.repeat(par["obar"], rep, axis=0).astype(complex)
    half = lambda a, sgn: (0.5 * (a[:, 0] + sgn * a[:, 1]))[:, None]
    return {"cx0": half(cx, 1), "cx1": half(cx, -1), "wx0": half(wx, 1), "wx1": half(wx, -1), "cex0": half(cex, 1),
            "cex1": half(cex, -1), "obx0": half(obx, 1), "obx1": half(obx, -1), "cx": cx[:, :, None], "cex": cex[:, :, None]}


def build_case(oracle_mod, hoh=False, inp=INPUT, par=FE, cr=None):
    """-> (lattice, hamiltonian, ene) of the bccFe regression case in the containers the oracle / the GPU library take
    (cr: another cluster in units of alat, site 1 = the representative atom; default the bcc cut of `inp`)"""
    from rslmtoasa_b200.synthetic import Lattice, Hamiltonian
    if cr is None:
        cr = bravais_cluster(inp["rc"])
    kk = cr.shape[1]
    crd = np.asfortranarray(cr * inp["alat"])
    nn, nm, rc = oracle_mod.build_nn(crd, np.ones(kk, np.int32), [1], inp["ct"])
    assert rc == 0
    sbar, vecs = screened_structure_constants(crd, 1, inp["r2"], inp["wav"])
    nr = int(nn[0, 0])
    assert sbar.shape[2] == nr                                      # same neighbour set as nncal (r2 = ct^2)
    for m in range(1, nr):                                          # hmfind: the slot order IS the sbar order
        assert np.abs(crd[:, nn[0, m] - 1] - crd[:, 0] - vecs[:, m]).max() < 1e-9
    nslot = nn.shape[1]
    hhh = np.zeros((9, 9, nslot, 1), order="F")
    for m in range(nr):
        hhh[:, :, m, 0] = sbar[:, :, m].T                           # hhh(ilm,jlm) = sbar(jlm,ilm,m,num(ia))
    jt = np.zeros((nslot, 1), np.int32); jt[:nr] = 1
    pot = build_pot(par)
    mom = par["mom"].reshape(3, 1)
    blk, blko, obarm, enim = HO.build_blocks(hhh, jt, np.array([1], np.int32), pot, mom, hoh)
    lat = Lattice(kk=kk, nn=np.asfortranarray(nn), iz=np.ones(kk, np.int32), ntype=1, nmax=0,
                  irec=np.array([1], np.int32), cr=cr)
    ham = Hamiltonian(ee=np.asfortranarray(blk), lsham=build_lsham(par["xi_p"], par["xi_d"]), hoh=hoh)
    if hoh:
        ham.eeo, ham.enim = np.asfortranarray(blko), np.asfortranarray(enim)
    ene = oracle_mod.e_mesh(inp["energy_min"], inp["energy_max"], inp["channels_ldos"], inp["fermi"])
    return lat, ham, ene


def total_dos(g0):
    """bands.f90:262: dtot(i) = -sum_j Im(g0(j,j,i) + g0(j+9,j+9,i)) / pi, summed over the recursion atoms"""
    d = np.arange(18)
    return -g0[d, d].imag.sum(axis=(0, 2)) / np.pi


def case_inputs(oracle_mod, name, _cache={}):
    """-> (lattice, hamiltonian, ene, settings) of a GOLDEN case (lattice / structure constants cached per hoh flag)"""
    import copy
    g = GOLDEN[name]
    inp = dict(INPUT)
    if "window" in g:
        inp["energy_min"], inp["energy_max"] = g["window"]
    if g["hoh"] not in _cache:
        _cache[g["hoh"]] = build_case(oracle_mod, hoh=g["hoh"])[:2]
    lat, ham = _cache[g["hoh"]]
    ham = copy.copy(ham)
    if g["nsp"] == 3:
        ham.lsham = np.zeros_like(ham.lsham)
    ene = oracle_mod.e_mesh(inp["energy_min"], inp["energy_max"], inp["channels_ldos"], inp["fermi"])
    return lat, ham, ene, dict(g, energy_min=inp["energy_min"], energy_max=inp["energy_max"], fermi=inp["fermi"],
                               channels_ldos=inp["channels_ldos"])


def oracle_total_dos(oracle_mod, name):
    """the restated reference pipeline run by the CPU oracle -> dtot on the mesh"""
    lat, ham, ene, g = case_inputs(oracle_mod, name)
    orc = oracle_mod.Oracle(lat, ham)
    if g["recur"] == "block":
        a_b, b2_b = orc.lanczos_block(lat.irec, g["lld"])
        g0 = oracle_mod.block_green(a_b, orc.zsqr(b2_b), ene)
    else:
        a, b = oracle_mod.cheb_scale(g["energy_min"], g["energy_max"])
        mu, _ = orc.cheb_moments(lat.irec, g["lld"], a, b)
        _, g0 = oracle_mod.chebyshev_green(mu, ene, g["energy_min"], g["energy_max"])
    dtot = oracle_mod.bands_dos(g0)[0]                  # calculate_fermi's DOS loop (rsrec_oracle_bands.c) writes totaldos.out
    assert np.allclose(dtot, total_dos(g0), rtol=0, atol=1e-12)
    return ene, dtot

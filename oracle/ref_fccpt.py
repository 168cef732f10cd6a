"""TEST INFRASTRUCTURE (part of the CPU oracle): the reference's conductivity post-processing case restated from its
input files up -- tests/postproc/cases/conductivity/fccPt (fcc Pt, 20 x 20 x 20 primitive cells with open boundaries,
nsp = 2, Kubo-Bastin with cond_ll = 50, per_type), whose stored outputs are
tests/postproc/references/Example_exchange_conductivity_fccPt*/ref.json (Pt_cond.out rows 500/1000/1500).

    lattice%bravais, pbc branch (lattice.f90:1037-1081)     -> pbc_cluster
    hamiltonian%build_realspace_velocity_operators (1308-1363), build_realspace_spin_operators (490-560) -> velocity_blocks
    post_processing_conductivity (calculation.f90:954-1071) -> conductivity_curve (recursion + Gamma contraction + tail)
Everything else (structure constants, potential, Hamiltonian blocks, mesh) is shared with ref_bccfe.py.
"""
import numpy as np

from . import ref_bccfe as R

# ---- tests/postproc/cases/conductivity/fccPt/input.nml ---------------------------------------------------------------
INPUT = dict(alat=3.9264951876, wav=1.5344598728, ct=4.0, r2=16.0, n=(20, 20, 20), fermi=-0.085837, energy_min=-2.5,
             energy_max=1.2, channels_ldos=2500, cond_ll=50, v_alpha=(0.0, 1.0, 0.0), v_beta=(1.0, 0.0, 0.0))
# ---- tests/postproc/cases/conductivity/fccPt/Pt.nml (&par), index [l, spin] --------------------------------------------
PT = dict(
    center_band=np.array([[-0.538906329111028, -0.538906329120192], [0.173709317657989, 0.173709317640736],
                          [-0.330833706305651, -0.330833706333152]]),
    width_band=np.array([[0.370054366308596, 0.370054366308049], [0.232127401294661, 0.232127401292080],
                         [0.150003832791302, 0.150003832789769]]),
    shifted_band=np.array([[7.642537183770061E-002, 7.642537183472130E-002], [0.584475095153574, 0.584475095154372],
                           [2.814518066619173E-002, 2.814518062236988E-002]]),
    obar=np.array([[-0.548883075029816, -0.548883075029913], [-0.710647821807747, -0.710647821816684],
                   [-1.105141923974950E-002, -1.105141914327711E-002]]),
    mom=np.array([0.0, 0.0, 1.0]),
    xi_p=np.array([0.215392421806229, 0.215392421805572]),
    xi_d=np.array([4.425297328320889E-002, 4.425297328594407E-002]),
)
# ---- tests/postproc/references/Example_exchange_conductivity_fccPt{,_hoh}/ref.json: row -> (E - E_F, real_part) ---------
GOLDEN = {
    "Example_exchange_conductivity_fccPt": {500: (-1.675556, -4.982769e-05), 1000: (-0.9354697, 0.001629417), 1500: (-0.1953829, 0.1025866)},
    "Example_exchange_conductivity_fccPt_hoh": {500: (-1.675556, -0.0001002817), 1000: (-0.9354697, 0.0005598055), 1500: (-0.1953829, 0.0617979)},
}

FCC_A = np.array([[0.0, 0.5, 0.5], [0.5, 0.0, 0.5], [0.5, 0.5, 0.0]]).T          # a(:, i) columns (lattice.f90:801-803)


def pbc_cluster(n1, n2, n3, a=FCC_A):
    """lattice%bravais with pbc = .true. (ntot = 1): the atom of the central cell first, then every other cell of the
    n1 x n2 x n3 block in the loop order nx outer .. nz inner, no cut; kk made even (lattice.f90:1091)"""
    lc = [(n + 1) // 2 for n in (n1, n2, n3)]
    p, q, s = np.meshgrid(np.arange(1, n1 + 1) - lc[0], np.arange(1, n2 + 1) - lc[1], np.arange(1, n3 + 1) - lc[2], indexing="ij")
    p, q, s = p.ravel(), q.ravel(), s.ravel()
    keep = ~((p == 0) & (q == 0) & (s == 0))
    pts = p[None, keep] * a[:, [0]] + q[None, keep] * a[:, [1]] + s[None, keep] * a[:, [2]]
    cr = np.concatenate([np.zeros((3, 1)), pts], axis=1)
    kk = cr.shape[1] - cr.shape[1] % 2
    return np.asfortranarray(cr[:, :kk])


def build_case(oracle_mod, hoh=False):
    inp = dict(INPUT)
    cr = pbc_cluster(*inp["n"])
    return R.build_case(oracle_mod, hoh=hoh, inp=inp, par=PT, cr=cr) + (cr,)


def s_op(pol):
    """S_x, S_y, S_z of math.f90:167-230: Pauli matrix (x) 1_9, divided by 2"""
    sig = {"x": np.array([[0, 1], [1, 0]], complex), "y": np.array([[0, -1j], [1j, 0]]), "z": np.array([[1, 0], [0, -1]], complex)}[pol]
    return np.kron(sig, np.eye(9)) / 2.0


def velocity_blocks(lat, ham, cr, alat, direction, spin_pol=None, obarm=None):
    """v(:,:,m,1) = (1/i) (dir . (r_i - r_j)) ee(:,:,m,1) for the representative atom's slots m >= 2
    (hamiltonian.f90:1336-1348); spin_pol: j^S = 1/2 {S_pol, v} (490-560); vo = v * obarm (hoh, 1353-1357)"""
    d = np.asarray(direction, float)
    d = d / np.linalg.norm(d)
    nn = lat.nn
    nr = int(nn[0, 0])
    v = np.zeros_like(ham.ee)
    for m in range(1, nr):
        j = int(nn[0, m])
        if j == 0:
            continue
        rij = (cr[:, 0] - cr[:, j - 1]) * alat
        v[:, :, m, 0] = (1.0 / 1j) * float(d @ rij) * ham.ee[:, :, m, 0]
    vo = None
    if obarm is not None:
        vo = np.zeros_like(v)
        for m in range(1, nr):
            vo[:, :, m, 0] = v[:, :, m, 0] @ obarm[:, :, 0]
    if spin_pol is not None:
        S = s_op(spin_pol)
        for arr in (v, vo):
            if arr is not None:
                for m in range(1, nr):
                    arr[:, :, m, 0] = 0.5 * (S @ arr[:, :, m, 0] + arr[:, :, m, 0] @ S)
    return np.asfortranarray(v), (None if vo is None else np.asfortranarray(vo))


def kubo_inputs(oracle_mod, hoh):
    """-> (lattice, hamiltonian with the Kubo operator slots filled, ene, e_mesh dict): output operator = the spin current
    j^S_z = 1/2 {S_z, v_alpha} (`cond_type = 'spin'`, `js_alpha = 'z'` of the case file; in today's namelists
    linear_out = 'spin', pol_alpha = 'z', recursion.f90:236-262), input operator = the charge current v_beta"""
    from . import ham_oracle as HO
    lat, ham, ene, cr = build_case(oracle_mod, hoh=hoh)
    obarm = HO.build_obarm(R.build_pot(PT), PT["mom"].reshape(3, 1)) if hoh else None
    ham.v_a, ham.vo_a = velocity_blocks(lat, ham, cr, INPUT["alat"], INPUT["v_alpha"], "z", obarm)
    ham.v_b, ham.vo_b = velocity_blocks(lat, ham, cr, INPUT["alat"], INPUT["v_beta"], None, obarm)
    mesh = oracle_mod.e_mesh_full(INPUT["energy_min"], INPUT["energy_max"], INPUT["channels_ldos"], INPUT["fermi"])
    return lat, ham, ene, mesh


def check_rows(ene, real_part, name, abs_tol=1e-6, rel_tol=1e-6):
    """the comparison rule of the reference's harness (tests/run_test.py:206-215): a value fails only if it is off by more
    than abs_tol AND by more than rel_tol; the stored numbers carry 7 significant digits (es16.6)"""
    worst = 0.0
    for row, (e_ref, v_ref) in GOLDEN[name].items():
        for got, ref in ((ene[row - 1] - INPUT["fermi"], e_ref), (real_part[row - 1], v_ref)):
            ad = abs(got - ref)
            rd = ad / max(abs(ref), 1e-300)
            assert not (ad > abs_tol and rd > rel_tol), (name, row, got, ref)
            worst = max(worst, rd)
    return worst


def oracle_conductivity(oracle_mod, hoh):
    """post_processing_conductivity (calculation.f90:954-1071) run by the CPU oracle -> (ene, real_part of Pt_cond.out)"""
    lat, ham, ene, mesh = kubo_inputs(oracle_mod, hoh)
    orc = oracle_mod.Oracle(lat, ham)
    a, b = oracle_mod.cheb_scale(INPUT["energy_min"], INPUT["energy_max"])
    mu = orc.kubo_moments(INPUT["cond_ll"], a, b, start_sites=[1])
    integ, integ_at = oracle_mod.conductivity_integrand(mu, ene, INPUT["energy_min"], INPUT["energy_max"], True)
    sig = oracle_mod.conductivity_cumulative(integ, integ_at, mesh["nv1"], (ene - b) / a, 1)
    return ene, sig[0, 0, :, 1]

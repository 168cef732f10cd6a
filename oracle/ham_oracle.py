"""CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE) for the Hamiltonian block assembly that feeds the recursion
(SURVEY.md 8f row 4): numpy restatement, statement for statement, of

    hcpx 'cart2sph'            math.f90:1508-1577
    ham0m_nc                   hamiltonian.f90:2225-2303
    chbar_nc (orbital part)    hamiltonian.f90:2305-2369   (structure-constant lookup `hmfind` is an input: `hhh`)
    build_bulkham / build_locham   hamiltonian.f90:1553-1616 / 1618-1667 (spin-block composition, eeo = ee*obarm)
    build_obarm / build_enim   hamiltonian.f90:1481-1508 / 1510-1551

PARITY PINNED by the reference's bccFe golden fixtures: oracle/ref_bccfe.py feeds the reference's Fe potential and
screened structure constants through build_blocks (nsp 1/2/4, hoh on/off) and the resulting DOS matches every stored
totaldos.out value (tests/test_reference_golden.py).  Pinned further by `pauli_block` below -- an independent statement of the same physics, block = P_i(l) S(l,l') P_j(l') with
P = w0 + w1 (m.sigma) as 2x2 spin matrices -- and by invariants in tests/test_oracle_ham.py.

Inputs (all per CLASS c = atom type 1..ntype followed by local site 1..nmax, like the device library):
    hhh  (9,9,nslot,ncls) real   hhh(ilm,jlm) of hmfind for slot m of the class's atom (slot 1 = on-site)
    jt   (nslot,ncls) int        type of the atom in that slot (slot 1: the atom itself), 0 = no neighbour
    it   (ncls) int              type of the class's own atom
    pot  dict of (9,ntype) complex arrays wx0 wx1 cx0 cx1 cex0 cex1 obx0 obx1 and (9,2,ntype) cx, cex
    mom  (3,ntype) real
"""
import numpy as np

I = 1j
POT_KEYS = ("wx0", "wx1", "cx0", "cx1", "cex0", "cex1", "obx0", "obx1")


def _v_vc():
    c = 1.0 / np.sqrt(2.0)
    v = np.zeros((9, 9), complex); vc = np.zeros((9, 9), complex)

    def s(m, i, j, val):       # 1-based like the reference
        m[i - 1, j - 1] = val
    s(v, 1, 1, 1); s(vc, 1, 1, 1)
    s(v, 2, 4, -c); s(vc, 4, 2, -c); s(v, 2, 2, c); s(vc, 2, 2, c)
    s(v, 3, 4, I * c); s(vc, 4, 3, -I * c); s(v, 3, 2, I * c); s(vc, 2, 3, -I * c)
    s(v, 4, 3, 1); s(vc, 3, 4, 1)
    s(v, 5, 5, I * c); s(v, 5, 9, -I * c); s(v, 6, 6, I * c); s(v, 6, 8, I * c)
    s(v, 7, 6, c); s(v, 7, 8, -c); s(v, 8, 5, c); s(v, 8, 9, c); s(v, 9, 7, 1)
    s(vc, 5, 5, -I * c); s(vc, 9, 5, I * c); s(vc, 6, 6, -I * c); s(vc, 8, 6, -I * c)
    s(vc, 6, 7, c); s(vc, 8, 7, -c); s(vc, 5, 8, c); s(vc, 9, 8, c); s(vc, 7, 9, 1)
    return v, vc


V, VC = _v_vc()


def hcpx_cart2sph(ham):
    """htmp = matmul(ham, v); hesf = matmul(vc, htmp)"""
    return VC @ (ham @ V)


def ham0m_nc(it, jt, onsite, hhh, pot, mom, hoh):
    """-> hhmag (9,9,4); it, jt 1-based types; onsite = (norm2(vet) <= 0.01)"""
    hh = np.zeros((9, 9, 4), complex)
    mi, mj = mom[:, it - 1], mom[:, jt - 1]
    dot = complex(np.dot(mi, mj))
    cross = np.array([mi[1] * mj[2] - mi[2] * mj[1], mi[2] * mj[0] - mi[0] * mj[2], mi[0] * mj[1] - mi[1] * mj[0]], complex)
    wx0i, wx1i = pot["wx0"][:, it - 1], pot["wx1"][:, it - 1]
    wx0j, wx1j = pot["wx0"][:, jt - 1], pot["wx1"][:, jt - 1]
    hc = hhh.astype(complex)
    for ilm in range(9):
        for jlm in range(9):
            hh[ilm, jlm, 3] = wx0i[ilm] * hc[ilm, jlm] * wx0j[jlm] + wx1i[ilm] * hc[ilm, jlm] * wx1j[jlm] * dot
    if onsite:
        for ilm in range(9):
            hh[ilm, ilm, 3] = hh[ilm, ilm, 3] + (pot["cex0"] if hoh else pot["cx0"])[ilm, it - 1]
    for m in range(3):
        for jlm in range(9):
            for ilm in range(9):
                hh[ilm, jlm, m] = (wx1i[ilm] * hc[ilm, jlm] * wx0j[jlm]) * complex(mi[m]) + \
                                  (wx0i[ilm] * hc[ilm, jlm] * wx1j[jlm]) * complex(mj[m]) + \
                                  I * wx1i[ilm] * hc[ilm, jlm] * wx1j[jlm] * cross[m]
    if onsite:
        for m in range(3):
            for ilm in range(9):
                hh[ilm, ilm, m] = hh[ilm, ilm, m] + (pot["cex1"] if hoh else pot["cx1"])[ilm, it - 1] * complex(mi[m])
    return hh


def spin_block(h):
    """ee(j,i)=H0+Hz, ee(j+9,i+9)=H0-Hz, ee(j,i+9)=Hx-iHy, ee(j+9,i)=Hx+iHy with h(:,:,1..4) = Hx,Hy,Hz,H0"""
    b = np.zeros((18, 18), complex)
    b[:9, :9] = h[:, :, 3] + h[:, :, 2]
    b[9:, 9:] = h[:, :, 3] - h[:, :, 2]
    b[:9, 9:] = h[:, :, 0] - I * h[:, :, 1]
    b[9:, :9] = h[:, :, 0] + I * h[:, :, 1]
    return b


def _spin_diag_18(d0, d1, mom):
    """build_obarm / build_enim pattern for diagonal 9x9 d0, d1, then hcpx on the four quadrants"""
    o = np.zeros((18, 18), complex)
    m0, m1 = np.diag(d0), np.diag(d1)
    for m in range(9):
        for l in range(9):
            o[m, l] = m0[m, l] + m1[m, l] * mom[2]
            o[m + 9, l + 9] = m0[m, l] - m1[m, l] * mom[2]
            o[l, m + 9] = m1[m, l] * mom[0] - I * m1[m, l] * mom[1]
            o[l + 9, m] = m1[m, l] * mom[0] + I * m1[m, l] * mom[1]
    for r, c in ((slice(0, 9), slice(0, 9)), (slice(9, 18), slice(9, 18)), (slice(0, 9), slice(9, 18)), (slice(9, 18), slice(0, 9))):
        o[r, c] = hcpx_cart2sph(o[r, c])
    return o


def build_obarm(pot, mom):
    nt = mom.shape[1]
    return np.stack([_spin_diag_18(pot["obx0"][:, t], pot["obx1"][:, t], mom[:, t].astype(complex)) for t in range(nt)], axis=2)


def build_enim(pot, mom):
    nt = mom.shape[1]
    out = []
    for t in range(nt):
        eu = pot["cx"][:, 0, t] - pot["cex"][:, 0, t]
        ed = pot["cx"][:, 1, t] - pot["cex"][:, 1, t]
        out.append(_spin_diag_18(0.5 * (eu + ed), 0.5 * (eu - ed), mom[:, t].astype(complex)))
    return np.stack(out, axis=2)


def build_blocks(hhh, jt, it, pot, mom, hoh):
    """-> blocks (18,18,nslot,ncls), blocks_o (= block * obarm(type in the slot); zeros when not hoh), obarm, enim"""
    nslot, ncls = jt.shape
    blk = np.zeros((18, 18, nslot, ncls), complex, order="F")
    blko = np.zeros_like(blk)
    obarm = build_obarm(pot, mom)
    enim = build_enim(pot, mom)
    for c in range(ncls):
        for m in range(nslot):
            if jt[m, c] == 0:
                continue
            hh = ham0m_nc(int(it[c]), int(jt[m, c]), m == 0, hhh[:, :, m, c], pot, mom, hoh)
            for mdir in range(4):
                hh[:, :, mdir] = hcpx_cart2sph(hh[:, :, mdir])
            blk[:, :, m, c] = spin_block(hh)
            if hoh:
                blko[:, :, m, c] = blk[:, :, m, c] @ obarm[:, :, jt[m, c] - 1]
    return blk, blko, obarm, enim


# ---- independent statement (pins the restatement above) -------------------------------------------------------------
SIG = np.array([[[0, 1], [1, 0]], [[0, -I], [I, 0]], [[1, 0], [0, -1]]], complex)


def pauli_block(it, jt, onsite, hhh, pot, mom, hoh):
    """block[(s,l),(s',l')] = sum over spin of P_i(l) S(l,l') P_j(l') (+ on-site C), P = w0 + w1 m.sigma, then the
    orbital transformation U = kron(1_2, V): block -> kron(1, VC) block kron(1, V)."""
    mi, mj = mom[:, it - 1], mom[:, jt - 1]
    msi = sum(mi[k] * SIG[k] for k in range(3)); msj = sum(mj[k] * SIG[k] for k in range(3))
    b = np.zeros((2, 9, 2, 9), complex)
    for l in range(9):
        Pi = pot["wx0"][l, it - 1] * np.eye(2) + pot["wx1"][l, it - 1] * msi
        for lp in range(9):
            Pj = pot["wx0"][lp, jt - 1] * np.eye(2) + pot["wx1"][lp, jt - 1] * msj
            b[:, l, :, lp] = Pi @ Pj * hhh[l, lp]
        if onsite:
            c0 = (pot["cex0"] if hoh else pot["cx0"])[l, it - 1]
            c1 = (pot["cex1"] if hoh else pot["cx1"])[l, it - 1]
            b[:, l, :, l] += c0 * np.eye(2) + c1 * msi
    b = b.reshape(18, 18)
    return np.kron(np.eye(2), VC) @ b @ np.kron(np.eye(2), V)


# ---- local-axis rotation: rotmag_loc / ROTMAT / DSs / car2sph (math.f90:1981-2192) -------------------------------------
def _factint(n):
    if n < 0:
        return 0
    f = 1
    for i in range(1, n + 1):
        f *= i
    return f


def _binom(x, y):
    if y < 0 or y > x:
        return 0
    return _factint(x) // (_factint(y) * _factint(x - y))


def _nint(x):
    return int(np.floor(x + 0.5)) if x >= 0 else -int(np.floor(-x + 0.5))


def dss(J, M, Mp, beta):
    """Wigner small-d element as the reference sums it (math.f90:2107-2125)"""
    smin = max(0, _nint(-Mp - M))
    smax = min(_nint(J - Mp), _nint(J - M))
    out = 0.0
    for s in range(smin, smax + 1):
        ds = s * 1.0
        dst = _binom(_nint(J + M), _nint(J - Mp - ds)) * _binom(_nint(J - M), s) * (-1.0) ** _nint(J - Mp - ds)
        out = out + dst * np.cos(0.5 * beta) ** (2 * ds + Mp + M) * np.sin(0.5 * beta) ** (2 * J - 2 * ds - Mp - M)
    return out * np.sqrt(1.0 * _factint(_nint(J + Mp)) * _factint(_nint(J - Mp)) / (_factint(_nint(J - M)) * _factint(_nint(J + M))))


def rotmat(a, b, g):
    """ROTMAT (math.f90:2055-2105): 18x18 = orbital D-matrices (l = 0,1,2) x spin-1/2 D-matrix"""
    sm = np.zeros((2, 2), complex)
    sm[0, 0] = dss(0.5, 0.5, 0.5, b) * np.exp(-I * (0.5 * a + 0.5 * g))
    sm[0, 1] = dss(0.5, 0.5, -0.5, b) * np.exp(-I * (0.5 * a - 0.5 * g))
    sm[1, 0] = dss(0.5, -0.5, 0.5, b) * np.exp(-I * (-0.5 * a + 0.5 * g))
    sm[1, 1] = dss(0.5, -0.5, -0.5, b) * np.exp(-I * (-0.5 * a - 0.5 * g))
    m9 = np.zeros((9, 9), complex)
    for J in range(3):
        S = J * J + 1 + J
        for M in range(-J, J + 1):
            for Mp in range(-J, J + 1):
                m9[S + M - 1, S + Mp - 1] = dss(J * 1.0, M * 1.0, Mp * 1.0, b) * np.exp(-I * (M * a + Mp * g))
    mat = np.zeros((18, 18), complex)
    for M in range(9):
        for Mp in range(9):
            mat[Mp, M] = m9[Mp, M] * sm[0, 0]
            mat[Mp, M + 9] = m9[Mp, M] * sm[0, 1]
            mat[Mp + 9, M] = m9[Mp, M] * sm[1, 0]
            mat[Mp + 9, M + 9] = m9[Mp, M] * sm[1, 1]
    return mat


def car2sph(c):
    """(theta, phi, r2) with the reference's conventions: theta = atan2(y,x) (0 when x=y=0), phi = acos(z / r^2)"""
    x, y, z = c
    d2, r2 = x * x + y * y, x * x + y * y + z * z
    theta = 0.0 if d2 == 0.0 else np.arctan2(y, x)
    return theta, np.arccos(z / r2), r2


def local_axis_rmat(mom):
    alfa, beta, _ = car2sph(mom)
    return rotmat(alfa, beta, 0.0)


def rotmag_loc(mat_in, mom):
    """MATout(:,:,j) = R^H (MATin(:,:,j) R) for every 18x18 block of mat_in (any trailing shape)"""
    r = local_axis_rmat(mom)
    flat = mat_in.reshape(18, 18, -1, order="F")
    out = np.stack([r.conj().T @ (flat[:, :, j] @ r) for j in range(flat.shape[2])], axis=2)
    return np.asfortranarray(out.reshape(mat_in.shape, order="F"))

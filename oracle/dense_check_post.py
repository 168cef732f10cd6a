"""Independent numpy restatement of the consumers either side of the recursion, used to pin
oracle/rsrec_oracle_post.c.  TEST INFRASTRUCTURE.

Different algorithms on purpose, so that an error in the C loops cannot be shared:
  * bgreen: numpy's LAPACK inverse, whole-matrix expressions (green.f90:1191-1339);
  * chebyshev_green: one einsum over closed-form factors -i exp(-i n acos w) (green.f90:1030-1108);
  * bprldos/density: vectorised continued fraction over the whole mesh (density_of_states.f90:248-407);
  * terminator: Beer-Pettifor iteration with numpy eigvalsh for the extreme eigenvalues of the scaled tridiagonal
    instead of the Sturm bisection (recursion.f90:3540-3706) -- agrees to the bisection's own 1e-6 tolerance;
  * gamma_nm / conductivity integrand: closed forms T_n(w) = cos(n acos w), broadcast sums (conductivity.f90:158-306).
"""
from __future__ import annotations

import numpy as np

NB = 18


def jackson_kernel(n):
    ll = np.arange(n, dtype=np.float64)
    th = np.pi * ll / (n + 1)
    return ((n - ll + 1) * np.cos(th) + np.sin(th) / np.tan(np.pi / (n + 1))) / (n + 1)


def lorentz_kernel(n, lam):
    ll = np.arange(n, dtype=np.float32)
    s = np.float32(1.0) - ll / np.float32(n)          # single-precision quotient like the reference
    th = lam * s.astype(np.float64)
    return np.sinh(th) / np.sinh(lam)


def bpopt(a, rb, eps=1e-5):
    """Beer-Pettifor terminator (a, rb of length ll; uses the first n = ll-1 levels)."""
    a = np.asarray(a, float); rb = np.asarray(rb, float)
    n = len(a) - 1
    ainf = a[n - 1]
    for it in range(301):
        az = 0.5 * (a[:n] - ainf)
        az[n - 1] = a[n - 1] - ainf
        bz = 0.5 * rb[:n].copy()
        bz[n - 1] = rb[n - 1] / np.sqrt(2.0)
        T = np.diag(az) + np.diag(bz[1:], 1) + np.diag(bz[1:], -1)
        ev = np.linalg.eigvalsh(T)
        bmax, bmin = ev[-1], ev[0]
        ainf = ainf + (bmax + bmin)
        if abs(bmax + bmin) <= eps:
            break
    return ainf, (bmax - bmin) / 2.0


def get_terminf(a_b, b_b):
    na = a_b.shape[3]
    a_inf = np.zeros((NB, NB, na)); b_inf = np.zeros((NB, NB, na))
    for n in range(na):
        for i in range(NB):          # only the diagonal is consumed by bgreen
            a_inf[i, i, n], b_inf[i, i, n] = bpopt(a_b[i, i, :, n].real, b_b[i, i, :, n].real)
        b_inf[0, 0, n] *= 1.01
        b_inf[9, 9, n] *= 1.01
    return a_inf, b_inf


def bgreen(a_b, b_b, ene, a_inf, b_inf, eta=0.0, sym_term=False):
    ll = a_b.shape[2]
    g = np.zeros((NB, NB, len(ene)), complex)
    da, db = np.diag(a_inf).copy(), np.diag(b_inf).copy()
    if sym_term:
        da = np.full(NB, 0.5 * (a_inf[0, 0] + a_inf[9, 9]))
        db = np.full(NB, 0.5 * (b_inf[0, 0] + b_inf[9, 9]))
        wid = 2.0 * db
    else:
        wid = 2.0 * db
        wid[[0, 9]] *= 1.025
    for k, e in enumerate(ene):
        zoff = np.sqrt(((e - (da + wid)) * (e - (da - wid))).astype(complex))
        Q = np.diag((e + eta - da - zoff) * 0.5)
        P = (e + eta if e != 0.0 else 0.0) * np.eye(NB)
        for l in range(ll - 1, 0, -1):
            A, B = a_b[:, :, l - 1], b_b[:, :, l - 1]
            Q = B.conj().T @ np.linalg.inv(P - A - Q) @ B
        g[:, :, k] = Q
    return g


def chebyshev_green(mu_n, ene, emin, emax):
    a, b = (emax - emin) / 1.7, (emax + emin) / 2
    nk = mu_n.shape[2]
    kern = jackson_kernel(nk)
    wt = np.full(nk, 2.0); wt[0] = 1.0
    w = (ene - b) / a
    th = np.arccos(w)
    n = np.arange(nk)
    fac = -1j * np.exp(-1j * np.outer(th, n))                      # (nv, nk)
    g = np.einsum("lmin,ei->lmen", mu_n * (kern * wt)[None, None, :, None], fac)
    return g / np.sqrt(a * a - (ene - b) ** 2)[None, None, :, None]


def bprldos_mesh(e, a, b2, edges):
    e = np.asarray(e, complex)
    ebot, etop = edges
    emid = 0.5 * (etop + ebot)
    zoff = np.sqrt((e - etop) * (e - ebot))
    q = (e - emid - zoff) * 0.5
    q = np.where(q.imag > 0, (e - emid + zoff) * 0.5, q)
    for l in range(len(a) - 1, 0, -1):
        q = b2[l - 1] / (e - a[l - 1] - q)
    return -q.imag / np.pi


def density(a, b2, ene, dw_l, cshi):
    td = np.zeros((NB, len(ene)))
    for nl in range(NB):
        am, bm = bpopt(a[:, nl], np.sqrt(b2[:, nl]))
        if nl in (0, 9):
            bm *= 1.01
        edges = (am - 2 * bm, am + 2 * bm)
        td[nl] = bprldos_mesh(ene / dw_l[nl] - cshi[nl], a[:, nl], b2[:, nl], edges) / dw_l[nl]
    return td


def gamma_nm(ene, M, emin, emax):
    a, b = (emax - emin) / 1.7, (emax + emin) / 2
    w = (ene - b) / a
    th = np.arccos(w)
    n = np.arange(M)
    T = np.cos(np.outer(th, n))                                        # T_n(w)
    sq = np.sqrt(1 - w * w)[:, None]
    cn = (w[:, None] - 1j * n[None, :] * sq) * np.exp(1j * np.outer(th, n))
    cm = (w[:, None] + 1j * n[None, :] * sq) * np.exp(-1j * np.outer(th, n))
    gk = lorentz_kernel(M, 6.0)
    wt = np.ones(M); wt[0] = 0.5
    g = cn[:, :, None] * T[:, None, :] + cm[:, None, :] * T[:, :, None]
    g = g / ((1 - w * w) ** 2)[:, None, None]
    return g * (gk * wt)[None, :, None] * (gk * wt)[None, None, :]


def conductivity_integrand(mu_nm, ene, emin, emax):
    M = mu_nm.shape[2]
    g = gamma_nm(ene, M, emin, emax)
    factor = 16 / (np.pi * (emax - emin) ** 2)
    d = np.arange(NB)
    mud = mu_nm[d, d]                                                  # (18, M, M, nloop)
    at = factor * np.einsum("enm,lnmt->let", g, mud)
    return at.sum(-1), at


# ---- `type bands`: independent numpy statement (vectorised; Pauli-matrix form instead of the element-wise loops) ------
def bands_projections(g0, mom):
    """g0 (18,18,nv,nu), mom (3,nu) -> dict of energy-resolved quantities:
    dos (nv,nu), spin (3,nv,nu) = -Im Tr(sigma_d g0)/pi, dspd (6,nv,nu), lorb (3,nv,nu) = Im Tr(L_d g0)"""
    from .oracle import l_spherical
    g = np.asarray(g0).reshape(9, 2, 9, 2, g0.shape[2], g0.shape[3], order="F")   # axes (o, s, o', s', e, u): row = o + 9 s
    sig = np.array([[[0, 1], [1, 0]], [[0, -1j], [1j, 0]], [[1, 0], [0, -1]]], complex)
    one = np.eye(2)
    # orbital-diagonal 2x2 spin blocks G_o[s, s']
    go = np.einsum("osotev->ostev", g)
    dos = -np.imag(np.einsum("ossev->ev", go)) / np.pi
    # Tr_spin(sigma_d G): the reference's expressions are aimag of the combinations below (no conjugations)
    spin = np.stack([-np.imag(np.einsum("st,otsev->ev", sig[d], go)) / np.pi for d in range(3)])
    lsl = [slice(0, 1), slice(1, 4), slice(4, 9)]
    nv, nu = g0.shape[2], g0.shape[3]
    dspd = np.zeros((6, nv, nu))
    for isp in range(2):
        sgn = 1.0 - 2.0 * isp
        for l in range(3):
            G = go[lsl[l]]
            n = np.imag(np.einsum("ossev->ev", G))
            s = [np.imag(np.einsum("st,otsev->ev", sig[d], G)) for d in range(3)]
            dspd[l + 3 * isp] = (-n - sgn * (mom[0][None, :] * s[0] + mom[1][None, :] * s[1] + mom[2][None, :] * s[2])) * 0.5 / np.pi
    L = l_spherical()
    lorb = np.stack([np.imag(np.einsum("ik,ksisev->ev", L[:, :, d], g)) for d in range(3)])
    return {"dos": dos, "spin": spin, "dspd": dspd, "lorb": lorb}


def simpson_to_fermi(y, ene, edel, fermi, nv1, e1, nexp):
    """simpson_m in closed vector form: composite Simpson over points 1..nv1 plus the partial last panel"""
    f = y * ene ** nexp
    w = np.zeros(nv1)
    w[0:nv1 - 2:2] += 1.0; w[1:nv1 - 1:2] += 4.0; w[2:nv1:2] += 1.0
    val = edel / 3.0 * np.tensordot(w, f[:nv1], axes=(0, 0))
    if e1 != fermi:
        val = val + (fermi - e1) * (f[nv1 - 1] + 4.0 * f[nv1] + f[nv1 + 1]) / 6.0
    return val


def orbital_tail(mu_n_orb, kk, ene, fermi, emin, emax, nv1):
    """tail of chebyshev_orbital_mod (recursion.f90:3009-3049), loop for loop: -> rows (E - E_F, -lz/pi, -lzi/pi) of fort.50.
    Test infrastructure (the product's vectorised form is Recursion.chebyshev_orbital_tail)."""
    import cmath
    import math
    mu = np.array(mu_n_orb, dtype=np.complex128)
    lld = mu.shape[2]
    nv = len(ene)
    a = (emax - emin) / (2 - 0.3)
    b = (emax + emin) / 2
    kernel = jackson_kernel(lld)
    mu = mu / float(kk)
    for l in range(18):
        for m in range(18):
            mu[l, m, :] = mu[l, m, :] * kernel
    mu[:, :, 1:] = mu[:, :, 1:] * 2.0
    lzi = np.zeros(nv)
    for ie in range(nv):
        g0 = np.zeros((18, 18), np.complex128)
        wsc = (ene[ie] - b) / a
        for i in range(1, lld + 1):
            exp_factor = -1j * cmath.exp(-1j * (i - 1) * math.acos(wsc))   # |wsc| < 1 on the reference's meshes (10 points past energy_max)
            g0 += mu[:, :, i - 1] * exp_factor.imag
        g0 = g0 / math.sqrt(a * a - (ene[ie] - b) ** 2)
        lzi[ie] = sum(g0[k, k].real for k in range(18))
    h = ene[1] - ene[0]
    out = np.zeros((nv, 3))

    def fermifun(e, ef, kbt):
        x = (e - ef) / kbt
        return 0.0 if x > 700.0 else 1.0 / (math.exp(x) + 1.0)

    def y(i1, ef):   # Y(I) * fermifun(Ene(I)) with 1-based I; one past the mesh (reference reads out of bounds there) counts as zero
        return lzi[i1 - 1] * fermifun(ene[i1 - 1], ef, 1.0e-15) if i1 <= nv else 0.0

    for ie in range(nv):
        aint = 0.0
        for i1 in range(2, nv1 + 10, 2):
            aint += y(i1 - 1, ene[ie]) + 4.0 * y(i1, ene[ie]) + y(i1 + 1, ene[ie])
        lz = h * aint / 3.0
        out[ie] = (ene[ie] - fermi, -(lz / math.pi), -(1 / math.pi) * lzi[ie])
    return out

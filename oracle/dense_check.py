"""Independent dense-matrix restatement of the recursions (numpy), used to pin the C oracle on small clusters.

TEST INFRASTRUCTURE.  It assembles the full (18 kk) x (18 kk) operator from the reference's data structures
(`nn`, `iz`, `ee/hall`, `lsham`, `eeo/hallo`, `enim`) and runs textbook block-Lanczos / Chebyshev / Kubo formulas
with dense linear algebra -- no masks, no neighbour loops, numpy's LAPACK for the matrix square root -- so an
error in the oracle's loop/mask/index logic cannot be shared.
"""
from __future__ import annotations

import numpy as np

NB = 18


def _assemble(lat, htype, hsite, onsite_extra=None, skip_onsite=False, first_site=1):
    kk = lat.kk
    H = np.zeros((NB * kk, NB * kk), dtype=np.complex128)
    for i in range(first_site, kk + 1):
        t = lat.iz[i - 1]
        nr = lat.nn[i - 1, 0]
        blk = (lambda j: hsite[:, :, j - 1, i - 1]) if i <= lat.nmax and hsite is not None else (lambda j: htype[:, :, j - 1, t - 1])
        r = slice(NB * (i - 1), NB * i)
        if not skip_onsite:
            H[r, r] += blk(1)
            if onsite_extra is not None:
                H[r, r] += onsite_extra[:, :, t - 1]
        for j in range(2, nr + 1):
            nb = lat.nn[i - 1, j - 1]
            if nb != 0:
                H[r, NB * (nb - 1):NB * nb] += blk(j)
    return H


def dense_hamiltonian(lat, ham):
    """Full operator applied by hop_b / chebyshev_recur_ll / ham_vec_matmul (and their hoh variants)."""
    if not ham.hoh:
        return _assemble(lat, ham.ee, ham.hall, onsite_extra=ham.lsham)
    h = _assemble(lat, ham.ee, ham.hall)
    ho = _assemble(lat, ham.eeo, ham.hallo)
    kk = lat.kk
    E = np.zeros_like(h)
    for i in range(1, kk + 1):
        t = lat.iz[i - 1]
        r = slice(NB * (i - 1), NB * i)
        E[r, r] = ham.enim[:, :, t - 1] + ham.lsham[:, :, t - 1]
    return h - ho @ h + E


def start_block(lat, site_i, site_j=0, asign=1.0, bsign=1.0):
    W = np.zeros((NB * lat.kk, NB), dtype=np.complex128)
    W[NB * (site_i - 1):NB * site_i] = asign * np.eye(NB)
    if site_j > 0:
        W[NB * (site_j - 1):NB * site_j] = bsign * np.eye(NB)
    return W


def random_block(lat, u):
    W = np.zeros((NB * lat.kk, NB), dtype=np.complex128)
    for k in range(lat.kk):
        W[NB * k:NB * (k + 1)] = np.exp(2j * np.pi * u[k]) * np.eye(NB) / float(np.sqrt(np.float32(lat.kk)))
    return W


def _psd_sqrt(m):
    ev, u = np.linalg.eigh(0.5 * (m + m.conj().T))
    return (u * np.sqrt(ev)) @ u.conj().T, (u / np.sqrt(ev)) @ u.conj().T


def block_lanczos(H, W, lld):
    """A_n (n=1..lld-1, A_lld = 0) and B2_n (B2_1 = I) exactly as crecal_b stores them."""
    a_b = np.zeros((NB, NB, lld), dtype=np.complex128)
    b2_b = np.zeros((NB, NB, lld), dtype=np.complex128)
    b2 = np.eye(NB, dtype=np.complex128)
    Wprev_B = np.zeros_like(W)
    for ll in range(lld - 1):
        HW = H @ W
        A = W.conj().T @ HW
        a_b[:, :, ll] = A
        b2_b[:, :, ll] = b2
        R = HW - Wprev_B - W @ A
        b2 = R.conj().T @ R
        B, Bi = _psd_sqrt(b2)
        Wprev_B = W @ B
        W = R @ Bi
    b2_b[:, :, lld - 1] = b2
    return a_b, b2_b


def scalar_lanczos(lat, ham, site, lld):
    """recur (nsp=1): only the two 9x9 spin-diagonal sub-blocks of ee/hall act, no lsham."""
    def spin_diag(arr):
        if arr is None:
            return None
        out = np.zeros_like(arr)
        out[:9, :9] = arr[:9, :9]
        out[9:, 9:] = arr[9:, 9:]
        return out
    H = _assemble(lat, spin_diag(ham.ee), spin_diag(ham.hall))
    a = np.zeros((lld, NB))
    b2 = np.zeros((lld, NB))
    for l in range(NB):
        psi = np.zeros(NB * lat.kk, dtype=np.complex128)
        psi[NB * (site - 1) + l] = 1.0
        prev = np.zeros_like(psi)
        s = 1.0
        for ll in range(lld - 1):
            v = H @ psi
            an = np.real(np.vdot(psi, v))
            a[ll, l] = an
            b2[ll, l] = s
            r = v - prev - an * psi
            s = np.real(np.vdot(r, r))
            prev = psi * np.sqrt(s)
            psi = r / np.sqrt(s)
        b2[lld - 1, l] = s
    return a, b2


def cheb_moments(H, W, lld, a, b):
    """mu_n(:,:,1..2lld+2) with the reference's doubling formulas (no masks)."""
    n = H.shape[0]
    Ht = (H - b * np.eye(n)) / a
    mu = np.zeros((NB, NB, 2 * lld + 2), dtype=np.complex128)
    p0 = W
    p1 = Ht @ p0
    mu[:, :, 0] = W.conj().T @ p0
    mu[:, :, 1] = W.conj().T @ p1
    for ll in range(1, lld + 1):
        p2 = 2.0 * (Ht @ p1) - p0
        mu[:, :, 2 * ll] = 2.0 * (p1.conj().T @ p1) - mu[:, :, 0]
        mu[:, :, 2 * ll + 1] = 2.0 * (p2.conj().T @ p1) - mu[:, :, 1]
        p0, p1 = p1, p2
    return mu


def cheb_moments_direct(H, W, nmom, a, b):
    """mu_n = W^H T_n(H~) W without the doubling trick (equal to the doubled ones iff H is Hermitian)."""
    n = H.shape[0]
    Ht = (H - b * np.eye(n)) / a
    mu = np.zeros((NB, NB, nmom), dtype=np.complex128)
    t0, t1 = W, Ht @ W
    mu[:, :, 0] = W.conj().T @ t0
    mu[:, :, 1] = W.conj().T @ t1
    for k in range(2, nmom):
        t0, t1 = t1, 2.0 * (Ht @ t1) - t0
        mu[:, :, k] = W.conj().T @ t1
    return mu


def kubo_moments(lat, ham, W, M, a, b):
    """mu_nm(:,:,n,m) = (T_m W)^H v_a T_n v_b W (`recursion.f90:1145-1229`), no hoh."""
    H = dense_hamiltonian(lat, ham)
    n = H.shape[0]
    Ht = (H - b * np.eye(n)) / a
    Va = _assemble(lat, ham.v_a, None, first_site=lat.nmax + 1)
    Vb = _assemble(lat, ham.v_b, None, first_site=lat.nmax + 1)

    def cheb_list(x):
        out = [x, Ht @ x]
        for _ in range(2, M):
            out.append(2.0 * (Ht @ out[-1]) - out[-2])
        return out[:M]

    left = cheb_list(W)
    right = [Va @ t for t in cheb_list(Vb @ W)]
    mu = np.zeros((NB, NB, M, M), dtype=np.complex128)
    for nn_ in range(M):
        for m in range(M):
            mu[:, :, nn_, m] = left[m].conj().T @ right[nn_]
    return mu


def ll_map(lat, site, lld):
    """create_ll_map (`recursion.f90:3277-3303`) via powers of the adjacency pattern: column ll+1 marks the sites
    that gather from a site marked in column ll (or are marked already)."""
    kk = lat.kk
    A = np.zeros((kk, kk), dtype=bool)
    for i in range(1, kk + 1):
        for j in range(2, lat.nn[i - 1, 0] + 1):
            nb = lat.nn[i - 1, j - 1]
            if nb != 0:
                A[i - 1, nb - 1] = True
    m = np.zeros((kk + 1, lld + 1), dtype=np.int32)
    m[site, 0] = 1
    for ll in range(lld):
        cur = m[1:, ll].astype(bool)
        m[1:, ll + 1] = (cur | (A @ cur)).astype(np.int32)
    return m


def orbital_moments(lat, ham, start_sites, cr, alat, lld, a, b):
    """chebyshev_orbital_mod's moments (`recursion.f90:2901-3008`): mu_n = sum_r L_r^H T_{n-1}(H~) |r>,
    |L_r> = i (Y H0~ X - X H0~ Y)|r>, H0 = the non-hoh operator (ham_vec_matmul), H = the full one."""
    H = dense_hamiltonian(lat, ham)
    n = H.shape[0]
    Ht = (H - b * np.eye(n)) / a
    if ham.hoh:
        H0 = _assemble(lat, ham.ee, ham.hall, onsite_extra=ham.lsham)
        H0t = (H0 - b * np.eye(n)) / a
    else:
        H0t = Ht
    X = np.kron(np.diag(cr[0] * alat), np.eye(NB))
    Y = np.kron(np.diag(cr[1] * alat), np.eye(NB))
    mu = np.zeros((NB, NB, lld), dtype=np.complex128)
    for r in start_sites:
        W = start_block(lat, r)
        left = 1j * (Y @ H0t @ X @ W - X @ H0t @ Y @ W)
        t0, t1 = W, Ht @ W
        for k in range(lld):
            cur = t0 if k == 0 else t1
            mu[:, :, k] += left.conj().T @ cur
            if k >= 1:
                t0, t1 = t1, 2.0 * (Ht @ t1) - t0
    return mu

"""CPU tests of the oracle for the consumers either side of the recursion: terminator,
block / Chebyshev Green functions, scalar continued-fraction DOS, Kubo-Bastin integrand.

Pinned (the reference stores no fixtures at this boundary) by oracle/dense_check_post.py -- numpy/LAPACK, closed forms,
different algorithms -- and by physical invariants (sum rules, Herglotz sign, semicircle limit).
"""
import numpy as np
import pytest

from oracle import dense_check_post as DP
from tests.cases import case, relerr, EMIN, EMAX


@pytest.fixture(scope="module")
def coeffs(oracle_mod):
    """a_b, B (= zsqr(b2_b)) of two units on the small bulk cluster, lld = 9."""
    lat, ham = case("bulk")
    orc = oracle_mod.Oracle(lat, ham)
    a_b, b2_b = orc.lanczos_block([1, 4], 9)
    return a_b, orc.zsqr(b2_b)


@pytest.fixture(scope="module")
def mesh(oracle_mod):
    return oracle_mod.e_mesh(-1.2, 1.0, 160, 0.05)


def test_e_mesh_matches_reference_rule(oracle_mod):
    ene = oracle_mod.e_mesh(-1.0, 1.0, 101, 0.1)      # odd channels_ldos is decremented (energy.f90:184-190)
    assert len(ene) == 110 and ene[0] == -1.0
    assert abs(ene[np.rint((0.1 + 1.0) / (2.0 / 100)).astype(int)] - 0.1) < 1e-14   # the mesh hits the Fermi level


def test_emami_vs_eigvalsh(oracle_mod):
    rng = np.random.default_rng(3)
    for n in (2, 5, 19):
        a = rng.normal(size=n) * 0.3
        b = np.abs(rng.normal(size=n)) * 0.2 + 0.05
        emax, emin = oracle_mod.emami(a, b)
        ev = np.linalg.eigvalsh(np.diag(a) + np.diag(b[1:], 1) + np.diag(b[1:], -1))
        assert abs(emax - ev[-1]) < 2e-6 * max(1, abs(ev[-1])) and abs(emin - ev[0]) < 2e-6 * max(1, abs(ev[0]))


def test_bpopt_constant_chain_and_vs_numpy(oracle_mod):
    ainf, binf, ifail = oracle_mod.bpopt(np.full(20, 0.3), np.full(20, 0.25))
    assert ifail == 0 and abs(ainf - 0.3) < 1e-5 and abs(binf - 0.25) < 2e-3      # -> band centre / quarter width
    rng = np.random.default_rng(4)
    a = 0.1 + 0.05 * rng.normal(size=12)
    rb = 0.3 + 0.02 * rng.normal(size=12)
    ainf, binf, ifail = oracle_mod.bpopt(a, rb)
    da, db = DP.bpopt(a, rb)
    assert ifail == 0 and abs(ainf - da) < 2e-5 and abs(binf - db) < 2e-5


def test_get_terminf_vs_numpy_and_fixups(oracle_mod, coeffs):
    a_b, b_b = coeffs
    a_inf, b_inf, a0, b0 = oracle_mod.get_terminf(a_b, b_b)
    da, db = DP.get_terminf(a_b, b_b)
    d = np.arange(18)
    assert np.abs(a_inf[d, d] - da[d, d]).max() < 5e-5 and np.abs(b_inf[d, d] - db[d, d]).max() < 5e-5
    assert np.allclose(a0, a_inf[d, d].mean(0)) and np.allclose(b0, b_inf[d, d].mean(0))
    assert not np.isnan(a_inf).any() and (a_inf[d, d] != 0).all() and (b_inf[d, d] != 0).all()


@pytest.mark.parametrize("sym_term", [False, True])
@pytest.mark.parametrize("eta", [0.0, 0.02j])
def test_bgreen_vs_numpy(oracle_mod, coeffs, mesh, sym_term, eta):
    a_b, b_b = coeffs
    a_inf, b_inf, _, _ = oracle_mod.get_terminf(a_b, b_b)
    for u in range(2):
        g = oracle_mod.bgreen(a_b[..., u], b_b[..., u], mesh, a_inf[..., u], b_inf[..., u], eta, sym_term)
        ref = DP.bgreen(a_b[..., u], b_b[..., u], mesh, a_inf[..., u], b_inf[..., u], eta, sym_term)
        assert relerr(g, ref) < 1e-10


def test_bgreen_channel_window_and_block_green(oracle_mod, coeffs, mesh):
    a_b, b_b = coeffs
    a_inf, b_inf, _, _ = oracle_mod.get_terminf(a_b, b_b)
    full = oracle_mod.block_green(a_b, b_b, mesh)
    one = oracle_mod.bgreen(a_b[..., 1], b_b[..., 1], mesh, a_inf[..., 1], b_inf[..., 1], 0.0, False, ie_start=40, ie_len=3)
    assert np.array_equal(one[:, :, 39:42], full[:, :, 39:42, 1])
    assert not one[:, :, :39].any() and not one[:, :, 42:].any()


def test_block_green_ldos_sum_rule_and_sign(oracle_mod):
    """-Im Tr G / pi is a non-negative density holding (almost all of) the 18 orbitals of the site; the rest sits in
    poles outside the terminator band, which a real-energy mesh cannot see."""
    lat, ham = case("bulk")
    orc = oracle_mod.Oracle(lat, ham)
    a_b, b2_b = orc.lanczos_block([1], 9)
    ene = oracle_mod.e_mesh(-3.0, 3.0, 3000, 0.0)
    g0 = oracle_mod.block_green(a_b, orc.zsqr(b2_b), ene)
    d = np.arange(18)
    ldos = -g0[d, d, :, 0].imag.sum(0) / np.pi
    assert ldos.min() > -1e-8
    assert 17.0 < np.trapezoid(ldos, ene) < 18.0 + 1e-6


def test_jackson_and_lorentz_kernels(oracle_mod):
    assert relerr(oracle_mod.jackson_kernel(202), DP.jackson_kernel(202)) < 1e-14
    assert relerr(oracle_mod.lorentz_kernel(300, 6.0), DP.lorentz_kernel(300, 6.0)) < 1e-14
    k = oracle_mod.jackson_kernel(50)
    assert abs(k[0] - 1.0) < 1e-15 and (np.diff(k) < 0).all()


def test_chebyshev_green_vs_numpy(oracle_mod, mesh):
    lat, ham = case("bulk")
    a, b = oracle_mod.cheb_scale(EMIN, EMAX)
    mu, rc = oracle_mod.Oracle(lat, ham).cheb_moments([1, 3], 12, a, b)
    mu_ng, g0 = oracle_mod.chebyshev_green(mu, mesh, EMIN, EMAX)
    ref = DP.chebyshev_green(mu, mesh, EMIN, EMAX)
    assert relerr(g0, ref) < 1e-12
    k = oracle_mod.jackson_kernel(mu.shape[2])
    assert np.allclose(mu_ng[:, :, 0], mu[:, :, 0] * k[0]) and np.allclose(mu_ng[:, :, 5], 2 * mu[:, :, 5] * k[5])


def test_chebyshev_green_agrees_with_block_green_ldos(oracle_mod):
    """two independent routes to the same local DOS (KPM vs continued fraction) agree after broadening"""
    lat, ham = case("pbc")
    orc = oracle_mod.Oracle(lat, ham)
    a, b = oracle_mod.cheb_scale(EMIN, EMAX)
    mu, _ = orc.cheb_moments([1], 40, a, b)
    ene = oracle_mod.e_mesh(-1.6, 1.6, 400, 0.0)[:400]
    _, gk = oracle_mod.chebyshev_green(mu, ene, EMIN, EMAX)
    a_b, b2_b = orc.lanczos_block([1], 12)
    gb = oracle_mod.block_green(a_b, orc.zsqr(b2_b), ene)
    d = np.arange(18)
    nk = np.trapezoid(-gk[d, d, :, 0].imag.sum(0) / np.pi, ene)
    nb = np.trapezoid(-gb[d, d, :, 0].imag.sum(0) / np.pi, ene)
    assert abs(nk - nb) < 0.5 and abs(nk - 18) < 0.5


def test_bprldos_semicircle_limit(oracle_mod):
    """constant a, b2 with the exact band edges reproduces the semicircle DOS at every level count"""
    a0, b0 = 0.2, 0.3
    for ll in (2, 7, 30):
        for e in (-0.3, 0.2, 0.55):
            b2 = np.full(ll, b0 * b0)
            b2[0] = 1.0                         # b2(1) is the squared norm of the start state (recursion.f90:3504)
            d = oracle_mod.bprldos(e, np.full(ll, a0), b2, (a0 - 2 * b0, a0 + 2 * b0))
            exact = np.sqrt(4 * b0 * b0 - (e - a0) ** 2) / (2 * np.pi * b0 * b0)
            assert abs(d - exact) < 1e-12


def test_density_and_sgreen_vs_numpy(oracle_mod, mesh):
    lat, ham = case("bulk")
    from rslmtoasa_b200 import synthetic as S
    ham1 = S.make_hamiltonian(lat, seed=20260101, spin_orbit=False)
    a, b2 = oracle_mod.Oracle(lat, ham1).lanczos_scalar([1, 2], 10)
    rng = np.random.default_rng(8)
    dw = 1.0 + 0.05 * rng.normal(size=(18, 2))
    cs = 0.02 * rng.normal(size=(18, 2))
    for u in range(2):
        td = oracle_mod.density(a[..., u], b2[..., u], mesh, dw[:, u], cs[:, u])
        ref = DP.density(a[..., u], b2[..., u], mesh, dw[:, u], cs[:, u])
        assert np.abs(td - ref).max() < 2e-4 * np.abs(ref).max()      # band edges differ by the bisection tolerance
        assert td.min() > -1e-12
    a4 = np.zeros((10, 18, 2, 3), order="F"); b4 = np.ones((10, 18, 2, 3), order="F")
    a4[..., 0], b4[..., 0] = a, b2
    g0 = oracle_mod.sgreen(a4, b4, 1, mesh, dw, cs)
    td = oracle_mod.density(a[..., 1], b2[..., 1], mesh, dw[:, 1], cs[:, 1])
    d = np.arange(18)
    assert np.allclose(g0[d, d, :, 1], -1j * np.pi * td) and abs(g0.sum() - g0[d, d].sum()) < 1e-12
    # three quantisation directions: charge part on the diagonal, spin part via the Pauli pattern (green.f90:688-699)
    a4[..., 1], b4[..., 1] = a * 1.01, b2
    a4[..., 2], b4[..., 2] = a * 0.99, b2
    g3 = oracle_mod.sgreen(a4, b4, 3, mesh, dw, cs)
    t = [oracle_mod.density(a4[:, :, 0, m], b4[:, :, 0, m], mesh, dw[:, 0], cs[:, 0]) for m in range(3)]
    dfac = 1j * np.pi / 2
    j = 2
    up_dn = [(x[j] + x[j + 9], x[j] - x[j + 9]) for x in t]
    want_jj = -sum(c for c, _ in up_dn) * dfac / 3 - up_dn[2][1] * dfac
    assert np.allclose(g3[j, j, :, 0], want_jj)
    assert np.allclose(g3[j, j + 9, :, 0], -up_dn[0][1] * dfac - up_dn[1][1] * (-1j) * dfac)


def test_gamma_nm_and_integrand_vs_numpy(oracle_mod):
    lat, ham = case("pbc")
    a, b = oracle_mod.cheb_scale(EMIN, EMAX)
    M = 7
    mu = oracle_mod.Oracle(lat, ham).kubo_moments(M, a, b, start_sites=[1, 2])
    ene = oracle_mod.e_mesh(-1.5, 1.5, 60, 0.0)
    g = oracle_mod.gamma_nm(ene, M, EMIN, EMAX)
    assert relerr(g, DP.gamma_nm(ene, M, EMIN, EMAX)) < 1e-12
    integ, integ_at = oracle_mod.conductivity_integrand(mu, ene, EMIN, EMAX, True)
    ri, rat = DP.conductivity_integrand(mu, ene, EMIN, EMAX)
    assert relerr(integ, ri) < 1e-12 and relerr(integ_at, rat) < 1e-12
    integ2, at2 = oracle_mod.conductivity_integrand(mu, ene, EMIN, EMAX, False)
    assert np.array_equal(integ2, integ) and not at2.any()


def test_post_golden_vectors(oracle_mod):
    """committed oracle outputs computed from the committed coefficient fixtures (tests/golden/make_golden_post.py)"""
    import os
    here = os.path.join(os.path.dirname(__file__), "golden")
    g, p = np.load(os.path.join(here, "oracle_golden.npz")), np.load(os.path.join(here, "post_golden.npz"))
    ene = p["ene"]
    b_b = oracle_mod.Oracle.zsqr(None, g["imp_b2_b"])
    assert relerr(b_b, p["b_b"]) < 1e-12
    a_inf, b_inf, a0, b0 = oracle_mod.get_terminf(g["imp_a_b"], p["b_b"])
    assert np.array_equal(a_inf, p["a_inf"]) and np.array_equal(b_inf, p["b_inf"]) and np.array_equal(a0, p["a_inf0"])
    assert relerr(oracle_mod.block_green(g["imp_a_b"], p["b_b"], ene), p["g0_block"]) < 1e-12
    assert relerr(oracle_mod.block_green(g["imp_a_b"], p["b_b"], ene, True), p["g0_block_sym"]) < 1e-12
    mu_ng, g0 = oracle_mod.chebyshev_green(g["imp_mu"], ene, EMIN, EMAX)
    assert relerr(mu_ng, p["mu_ng"]) < 1e-14 and relerr(g0, p["g0_cheb"]) < 1e-13
    td = oracle_mod.density(g["bulk_sa"][..., 0], g["bulk_sb"][..., 0], ene, p["dw"][:, 0], p["cs"][:, 0])
    assert relerr(td, p["tdens"]) < 1e-13
    integ, integ_at = oracle_mod.conductivity_integrand(g["pbc_kubo"], ene, EMIN, EMAX, True)
    assert relerr(np.nan_to_num(integ), np.nan_to_num(p["integrand"])) < 1e-13


def test_intersite_gf_is_the_off_diagonal_green_function(oracle_mod):
    """calculate_intersite_gf (green.f90:425-469) pinned by physics: with the four pair start vectors (|i> + s|j>)/sqrt2,
    s = 1, -1, i, -i, the combination ((g1 - g2) + (g3 - g4)/i)/2 is linear in the moments and must equal the Chebyshev
    Green function of the OFF-DIAGONAL moments <i|T_n(H)|j> computed with dense matrices; gji likewise for <j|T_n|i>."""
    from oracle import dense_check as D
    from tests.cases import case, EMIN, EMAX
    lat, ham = case("bulk")
    orc = oracle_mod.Oracle(lat, ham)
    a, b = oracle_mod.cheb_scale(EMIN, EMAX)
    lld, pairs = 6, np.array([[1, 2], [3, 3], [2, 9]], np.int32)
    s = 1 / np.sqrt(2)
    si, sj, asg, bsg, slots = [], [], [], [], []
    for p, (i, j) in enumerate(pairs):
        for r, sg in enumerate((1, -1, 1j, -1j)):
            if i == j and r > 0:
                continue
            si.append(i); sj.append(j); slots.append(4 * p + r)
            asg.append(1.0 if i == j else s); bsg.append(1.0 if i == j else s * sg)
    mu_c, _ = orc.cheb_moments(si, lld, a, b, site_j=sj, asign=asg, bsign=bsg)
    mu = np.zeros((18, 18, 2 * lld + 2, 4 * len(pairs)), complex, order="F")
    mu[..., slots] = mu_c
    ene = oracle_mod.e_mesh(-1.5, 1.5, 60, 0.0)[:60]
    g0 = oracle_mod.chebyshev_green(mu, ene, EMIN, EMAX)[1]
    gij, gji, gs = oracle_mod.intersite_gf(g0, pairs)
    H = D.dense_hamiltonian(lat, ham)
    Ht = (H - b * np.eye(H.shape[0])) / a
    for p, (i, j) in enumerate(pairs):
        Wi, Wj = D.start_block(lat, i), D.start_block(lat, j)
        t0, t1 = Wj, Ht @ Wj
        m_ij = np.zeros((18, 18, 2 * lld + 2, 1), complex, order="F")
        m_ji = np.zeros_like(m_ij, order="F")
        u0, u1 = Wi, Ht @ Wi
        for k in range(2 * lld + 2):
            m_ij[:, :, k, 0] = Wi.conj().T @ t0
            m_ji[:, :, k, 0] = Wj.conj().T @ u0
            t0, t1 = t1, 2.0 * (Ht @ t1) - t0
            u0, u1 = u1, 2.0 * (Ht @ u1) - u0
        want_ij = oracle_mod.chebyshev_green(m_ij, ene, EMIN, EMAX)[1][..., 0]
        want_ji = oracle_mod.chebyshev_green(m_ji, ene, EMIN, EMAX)[1][..., 0]
        scale = np.abs(want_ij).max()
        assert np.abs(gij[..., p] - want_ij).max() < 1e-11 * scale, p
        assert np.abs(gji[..., p] - want_ji).max() < 1e-11 * scale, p
    # spin decomposition: G = Gnmag*1 + Gx*sx + Gy*sy + Gz*sz restores the 18x18 block (orbital part 9x9 each)
    sig = [np.eye(2), np.array([[0, 1], [1, 0]]), np.array([[0, -1j], [1j, 0]]), np.array([[1, 0], [0, -1]])]
    for w, g in enumerate((gij, gji)):
        rebuilt = sum(np.einsum("st,jiep->sjtiep", sig[c], gs[..., 4 * w + c]) for c in range(4))
        rebuilt = rebuilt.reshape(18, 18, g.shape[2], g.shape[3])
        assert np.abs(rebuilt - g).max() < 1e-13 * np.abs(g).max()


def test_conductivity_cumulative_is_a_running_simpson_sum(oracle_mod):
    """the literal simpson_f loop of calculate_conductivity_tensor (one exp per term, O(nv^2)) against the closed form:
    at T = 0 the Fermi factor is 1 below, 1/2 at and 0 above E_F, so sigma(i) = h/3 (sum_{j<i} W_j Y_j + W_i Y_i / 2) with
    the composite Simpson weights W = 1,4,2,4,..."""
    m = oracle_mod.e_mesh_full(-1.0, 1.0, 80, 0.1)
    ene, nv, nv1 = m["ene"], len(m["ene"]), m["nv1"]
    a, b = oracle_mod.cheb_scale(-1.0, 1.0)
    ws = (ene - b) / a
    rng = np.random.default_rng(5)
    integ = np.asfortranarray(rng.standard_normal((18, nv)) + 1j * rng.standard_normal((18, nv)))
    integ_at = np.asfortranarray(rng.standard_normal((18, nv, 2)) + 1j * rng.standard_normal((18, nv, 2)))
    sig = oracle_mod.conductivity_cumulative(integ, integ_at, nv1, ws, 3)
    W = np.ones(nv); W[1::2] = 4.0; W[2::2] = 2.0
    h = ws[1] - ws[0]

    def closed(y):
        wy = W * y
        return h / 3.0 * (np.concatenate([[0.0], np.cumsum(wy)[:-1]]) + 0.5 * wy)
    for g, src, div in ((0, integ, 3.0), (1, integ_at[..., 0], 1.0), (2, integ_at[..., 1], 1.0)):
        for c, part in ((0, src.real), (1, src.imag)):
            assert np.allclose(sig[c, 0, :, g], closed(part.sum(axis=0)) / div, rtol=1e-12, atol=1e-13)
            for l2 in (0, 7, 17):
                assert np.allclose(sig[c, 1 + l2, :, g], closed(part[l2]) / div, rtol=1e-12, atol=1e-13)
    # physical reading: for a smooth integrand sigma(E_F) is its integral from the bottom of the mesh up to E_F
    y = np.exp(-ws ** 2)
    smooth = np.asfortranarray(np.tile(y / 18.0, (18, 1)).astype(complex))
    s2 = oracle_mod.conductivity_cumulative(smooth, None, nv1, ws, 1)
    from math import erf, sqrt, pi
    exact = np.array([sqrt(pi) / 2 * (erf(w) - erf(ws[0])) for w in ws])
    assert np.abs(s2[0, 0, 2::2, 0] - exact[2::2]).max() < 1e-6       # odd (1-based) points close whole Simpson panels


@pytest.mark.parametrize("channels", [400, 401])
def test_orbital_tail_vectorised_form_equals_the_loop_restatement(channels):
    """tail of chebyshev_orbital_mod (recursion.f90:3009-3049: Jackson weights, Chebyshev sum, trace, Fermi-weighted Simpson
    integral, fort.50 rows): the host mirror's vectorised form against the loop-for-loop restatement, and sanity of the result"""
    from rslmtoasa_b200.recursion import orbital_tail, Energy
    from oracle import dense_check_post as D
    rng = np.random.default_rng(11)
    lld = 14
    mu = rng.normal(size=(18, 18, lld)) + 1j * rng.normal(size=(18, 18, lld))
    en = Energy(-1.0, 1.2, channels_ldos=channels, fermi=0.1)
    en.e_mesh()
    rows, lz, lzi = orbital_tail(mu, 77, en)
    ref = D.orbital_tail(mu, 77, en.ene, en.fermi, -1.0, 1.2, en.nv1)
    assert rows.shape == (len(en.ene), 3) and np.isfinite(rows).all()
    assert np.abs(rows - ref).max() <= 1e-13 * np.abs(ref).max()
    assert np.allclose(rows[:, 0], en.ene - en.fermi) and np.allclose(rows[:, 2], -lzi / np.pi)
    # the running integral at mesh point k (0-based) holds every point below k with its full Simpson weight and point k with
    # the Fermi function's value at its own energy, 1/2 (simpson_f with fermi = .true. at kBT = 1e-15)
    h = en.ene[1] - en.ene[0]
    W = np.ones(len(en.ene)); W[1::2] = 4.0; W[2::2] = 2.0; W[0] = 1.0
    k = 22
    assert abs(lz[k] - h / 3.0 * (np.sum(W[:k] * lzi[:k]) + 0.5 * W[k] * lzi[k])) < 1e-12 * max(1.0, np.abs(lzi).max())

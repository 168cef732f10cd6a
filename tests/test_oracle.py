"""CPU tests of the oracle: pinned against the independent dense restatement and against invariants.

The reference stores no golden a_n/b_n/mu_n vectors; its end-to-end golden DOS values pin the oracle in
tests/test_reference_golden.py.  Here, at the coefficient boundary, the oracle is pinned by
(1) oracle/dense_check.py -- dense-matrix algebra with numpy/LAPACK, no masks or neighbour loops -- and
(2) mathematical invariants of the recursions, and (3) committed golden vectors generated from the oracle
(tests/golden, guards against regressions of the oracle itself).
"""
import os

import numpy as np
import pytest

from oracle import dense_check as D
from rslmtoasa_b200 import synthetic as S
from tests.cases import case, relerr, EMIN, EMAX

GOLD = os.path.join(os.path.dirname(__file__), "golden", "oracle_golden.npz")


@pytest.mark.parametrize("name", ["bulk", "bulk_hoh", "surface", "impurity", "impurity_hoh", "pbc", "tiny"])
def test_block_lanczos_vs_dense(oracle_mod, name):
    lat, ham = case(name)
    lld = 7
    a_b, b2_b = oracle_mod.Oracle(lat, ham).lanczos_block([1], lld)
    H = D.dense_hamiltonian(lat, ham)
    da, db = D.block_lanczos(H, D.start_block(lat, 1), lld)
    assert relerr(a_b[..., 0], da) < 1e-11
    assert relerr(b2_b[..., 0], db) < 1e-11


def test_block_lanczos_pair_start_vs_dense(oracle_mod):
    lat, ham = case("bulk")
    s = 1 / np.sqrt(2)
    a_b, b2_b = oracle_mod.Oracle(lat, ham).lanczos_block([2], 6, site_j=[5], asign=[s], bsign=[1j * s])
    H = D.dense_hamiltonian(lat, ham)
    da, db = D.block_lanczos(H, D.start_block(lat, 2, 5, s, 1j * s), 6)
    assert relerr(a_b[..., 0], da) < 1e-11 and relerr(b2_b[..., 0], db) < 1e-11


@pytest.mark.parametrize("name", ["bulk", "impurity"])
def test_scalar_lanczos_vs_dense(oracle_mod, name):
    lat, ham = case(name)
    a, b2 = oracle_mod.Oracle(lat, ham).lanczos_scalar([1], 8)
    da, db = D.scalar_lanczos(lat, ham, 1, 8)
    assert relerr(a[..., 0], da) < 1e-11 and relerr(b2[..., 0], db) < 1e-11


@pytest.mark.parametrize("name", ["bulk", "bulk_hoh", "surface", "impurity", "impurity_hoh", "pbc"])
def test_chebyshev_vs_dense(oracle_mod, name):
    lat, ham = case(name)
    a, b = oracle_mod.cheb_scale(EMIN, EMAX)
    mu, rc = oracle_mod.Oracle(lat, ham).cheb_moments([1], 9, a, b)
    assert rc == 0
    dm = D.cheb_moments(D.dense_hamiltonian(lat, ham), D.start_block(lat, 1), 9, a, b)
    assert relerr(mu[..., 0], dm) < 1e-11


def test_chebyshev_random_vs_dense(oracle_mod):
    lat, ham = case("pbc")
    a, b = oracle_mod.cheb_scale(EMIN, EMAX)
    ph = S.random_phases(lat.kk, 2)
    mu, _ = oracle_mod.Oracle(lat, ham).cheb_moments_random(ph, 6, a, b)
    H = D.dense_hamiltonian(lat, ham)
    for v in range(2):
        assert relerr(mu[..., v], D.cheb_moments(H, D.random_block(lat, ph[:, v]), 6, a, b)) < 1e-11


def test_kubo_vs_dense(oracle_mod):
    lat, ham = case("pbc")
    a, b = oracle_mod.cheb_scale(EMIN, EMAX)
    orc = oracle_mod.Oracle(lat, ham)
    mu = orc.kubo_moments(5, a, b, start_sites=[3])
    assert relerr(mu[..., 0], D.kubo_moments(lat, ham, D.start_block(lat, 3), 5, a, b)) < 1e-11
    ph = S.random_phases(lat.kk, 1)
    mu = orc.kubo_moments(4, a, b, phases=ph)
    assert relerr(mu[..., 0], D.kubo_moments(lat, ham, D.random_block(lat, ph[:, 0]), 4, a, b)) < 1e-11


# ---- invariants --------------------------------------------------------------------------------------------
def test_kubo_selected_left_indices_are_columns_of_the_full_moments(oracle_mod):
    """orc_kubo_moments_cols (used by the full-size GPU parity test, where cond_ll^2 contractions over 8000 sites take minutes
    on a CPU) runs the same chains and contracts a subset of the left indices: bit-identical to those columns"""
    lat, ham = case("pbc")
    orc = oracle_mod.Oracle(lat, ham)
    a, b = oracle_mod.cheb_scale(EMIN, EMAX)
    ph = S.random_phases(lat.kk, 2)
    full = orc.kubo_moments(6, a, b, phases=ph)
    msel = np.array([1, 4, 6], dtype=np.int32)
    sel = orc.kubo_moments_cols(6, a, b, msel, phases=ph)
    assert sel.shape == (18, 18, 6, 3, 2)
    assert np.array_equal(sel, full[:, :, :, msel - 1, :])
    none = orc.kubo_moments_cols(6, a, b, msel[:0], phases=ph[:, :1])     # chains only
    assert none.shape == (18, 18, 6, 0, 1)
    full_s = orc.kubo_moments(5, a, b, start_sites=[3])
    assert np.array_equal(orc.kubo_moments_cols(5, a, b, [2], start_sites=[3])[:, :, :, 0], full_s[:, :, :, 1])


def test_mask_is_only_an_optimisation(oracle_mod):
    """inactive sites hold exact zeros: results with and without izero/idum/irlist are identical"""
    for name in ["bulk", "impurity_hoh"]:
        lat, ham = case(name)
        a, b = oracle_mod.cheb_scale(EMIN, EMAX)
        m1, _ = oracle_mod.Oracle(lat, ham, use_mask=True).cheb_moments([1], 6, a, b)
        m0, _ = oracle_mod.Oracle(lat, ham, use_mask=False).cheb_moments([1], 6, a, b)
        assert relerr(m1, m0) < 1e-14
        a1 = oracle_mod.Oracle(lat, ham, use_mask=True).lanczos_block([1], 5)
        a0 = oracle_mod.Oracle(lat, ham, use_mask=False).lanczos_block([1], 5)
        assert relerr(a1[0], a0[0]) < 1e-13 and relerr(a1[1], a0[1]) < 1e-13


def test_active_region_grows_one_shell_per_step(oracle_mod):
    lat, ham = case("bulk")
    orc = oracle_mod.Oracle(lat, ham)
    orc.lanczos_block([1], 2)          # one hop from the centre site
    assert orc.last_irnum() == 15      # centre + 8 NN + 6 NNN
    orc.lanczos_block([1], 9)
    assert orc.last_irnum() == lat.kk  # whole 136-site cluster reached


def test_moment_identities(oracle_mod):
    lat, ham = case("pbc")             # Hermitian operator
    a, b = oracle_mod.cheb_scale(EMIN, EMAX)
    mu, _ = oracle_mod.Oracle(lat, ham).cheb_moments([1], 8, a, b)
    assert relerr(mu[:, :, 0, 0], np.eye(18)) < 1e-15          # mu_0 = psi0^H psi0 = I
    H = D.dense_hamiltonian(lat, ham)
    assert np.abs(H - H.conj().T).max() == 0.0
    direct = D.cheb_moments_direct(H, D.start_block(lat, 1), 18, a, b)   # doubling trick == direct for Hermitian H
    assert relerr(mu[..., 0], direct) < 1e-11
    for k in range(18):                                                   # moments of a Hermitian H are Hermitian
        assert relerr(mu[:, :, k, 0], mu[:, :, k, 0].conj().T) < 1e-12


def test_lanczos_reproduces_block_tridiagonal_projection(oracle_mod):
    """A_n = W_n^H H W_n with orthonormal W_n, and B_n^2 is Hermitian positive definite"""
    lat, ham = case("pbc")
    a_b, b2_b = oracle_mod.Oracle(lat, ham).lanczos_block([1], 6)
    for ll in range(6):
        assert relerr(b2_b[:, :, ll, 0], b2_b[:, :, ll, 0].conj().T) < 1e-12
        assert np.linalg.eigvalsh(b2_b[:, :, ll, 0]).min() > 0
    for ll in range(5):
        assert relerr(a_b[:, :, ll, 0], a_b[:, :, ll, 0].conj().T) < 1e-12
    assert not a_b[:, :, 5, 0].any()                                     # atemp_b(:,:,lld) = 0
    assert np.array_equal(b2_b[:, :, 0, 0], np.eye(18))                  # b2temp_b(:,:,1) = I


def test_site_permutation_invariance(oracle_mod):
    lat, ham = case("bulk")
    rng = np.random.default_rng(3)
    perm = np.concatenate([[0], 1 + rng.permutation(lat.kk - 1)])        # keep the start site first
    inv = np.argsort(perm)
    nn2 = np.zeros_like(lat.nn)
    nn2[:, 0] = lat.nn[perm, 0]
    old = lat.nn[perm, 1:]
    nn2[:, 1:] = np.where(old > 0, inv[np.maximum(old, 1) - 1] + 1, 0)
    lat2 = S.Lattice(kk=lat.kk, nn=np.asfortranarray(nn2), iz=lat.iz[perm].copy(), ntype=lat.ntype, nmax=0,
                     irec=lat.irec, disp=lat.disp)
    a, b = oracle_mod.cheb_scale(EMIN, EMAX)
    m1, _ = oracle_mod.Oracle(lat, ham).cheb_moments([1], 6, a, b)
    m2, _ = oracle_mod.Oracle(lat2, ham).cheb_moments([1], 6, a, b)
    assert relerr(m2, m1) < 1e-12


def test_zsqr_and_heev(oracle_mod):
    lat, ham = case("bulk")
    orc = oracle_mod.Oracle(lat, ham)
    _, b2_b = orc.lanczos_block([1], 5)
    b = orc.zsqr(b2_b)
    for ll in range(5):
        assert relerr(b[:, :, ll, 0] @ b[:, :, ll, 0], b2_b[:, :, ll, 0]) < 1e-12
    rng = np.random.default_rng(0)
    m = rng.normal(size=(18, 18)) + 1j * rng.normal(size=(18, 18))
    m = m @ m.conj().T
    ev, u = oracle_mod.heev18(m)
    assert np.abs(np.sort(ev) - np.linalg.eigvalsh(m)).max() < 1e-12 * np.abs(ev).max()
    assert relerr(u @ np.diag(ev) @ u.conj().T, m) < 1e-13


def test_divergence_guard(oracle_mod):
    lat, ham = case("bulk")
    a, b = oracle_mod.cheb_scale(-0.05, 0.05)
    _, rc = oracle_mod.Oracle(lat, ham).cheb_moments([1], 40, a, b)
    assert rc == -2


def test_single_precision_normalisation_of_random_start(oracle_mod):
    """`sqrt(real(kk))` is single precision in the reference (recursion.f90:1142): mu_0 = (1/sqrtf(kk))^2 kk I"""
    lat, ham = case("pbc")     # kk = 96, not a perfect square
    a, b = oracle_mod.cheb_scale(EMIN, EMAX)
    mu, _ = oracle_mod.Oracle(lat, ham).cheb_moments_random(S.random_phases(lat.kk, 1), 1, a, b)
    nrm = float(np.sqrt(np.float32(lat.kk)))
    assert abs(mu[0, 0, 0, 0].real - lat.kk / nrm ** 2) < 1e-14
    assert abs(mu[0, 0, 0, 0].real - 1.0) > 1e-9


# ---- golden vectors ------------------------------------------------------------------------------------------
def test_golden_vectors(oracle_mod):
    g = np.load(GOLD)
    lat, ham = case("impurity_hoh")
    orc = oracle_mod.Oracle(lat, ham)
    a, b = oracle_mod.cheb_scale(EMIN, EMAX)
    a_b, b2_b = orc.lanczos_block([1], 6)
    mu, _ = orc.cheb_moments([1], 6, a, b)
    assert relerr(a_b, g["imp_a_b"]) < 1e-12 and relerr(b2_b, g["imp_b2_b"]) < 1e-12 and relerr(mu, g["imp_mu"]) < 1e-12
    lat, ham = case("bulk")
    sa, sb = oracle_mod.Oracle(lat, ham).lanczos_scalar([1], 8)
    assert relerr(sa, g["bulk_sa"]) < 1e-12 and relerr(sb, g["bulk_sb"]) < 1e-12
    lat, ham = case("pbc")
    mk = oracle_mod.Oracle(lat, ham).kubo_moments(4, a, b, start_sites=[1])
    assert relerr(mk, g["pbc_kubo"]) < 1e-12


@pytest.mark.parametrize("name", ["bulk", "tiny", "pbc"])
def test_create_ll_map_vs_adjacency_powers(oracle_mod, name):
    lat, ham = case(name)
    orc = oracle_mod.Oracle(lat, ham)
    m = orc.create_ll_map(2, 6)
    assert np.array_equal(m, D.ll_map(lat, 2, 6))
    assert m[0].sum() == 0 and (np.diff(m.sum(0)) >= 0).all()          # index 0 stays 0; the region only grows


@pytest.mark.parametrize("name", ["bulk", "bulk_hoh", "impurity"])
def test_orbital_moments_vs_dense(oracle_mod, name):
    lat, ham = case(name)
    a, b = oracle_mod.cheb_scale(EMIN, EMAX)
    cr = 0.5 * lat.cr.astype(np.float64)
    mu = oracle_mod.Oracle(lat, ham).orbital_moments([1, 4], cr, 5.42, 7, a, b)
    dm = D.orbital_moments(lat, ham, [1, 4], cr, 5.42, 7, a, b)
    assert relerr(mu, dm) < 1e-11
